// kgpu_tile_table.hpp -- the tile bookkeeping of UpdateTiles.f90 / CheckIfNearBoundaries as a host-only value type
// (no CUDA in this file): groundwork for dynamic tile activation across ranks.
//
// A decomposed run cannot keep the tile table per rank: AddTile creates ghost tiles around the new tile
// (UpdateTiles.f90:389-481) and CheckIfNearBoundaries walks ONE ordered list of active tiles whose four passes
// read the trip count at loop entry while the list grows under them (quirk Q3, TimeStepper.f90:924-1150).  The
// plan: every rank holds this table for the WHOLE tile grid, the four flag bits of the
// active tiles are all-gathered, every rank replays the same mutation sequence, and the device work the replay
// asks for -- heights, ghost data, activation -- is executed by the rank that owns the tile (`ops`, in order).
// This file is the replay; it mirrors addTile / addGhostTiles / checkIfNearBoundaries of kestrel_gpu.cu
// statement for statement on global tile indices.  tests/test_tile_table.py checks it on CPU against the
// oracle's active and ghost sets, step by step, on the reference's dynamic-tile inputs (world size 2 over
// gloo: each rank contributes the flags of its own tiles).  kgpu_dyn_host.inl drives it in decomposed runs.
#pragma once
#include <algorithm>
#include <vector>

namespace kgpu {

enum TileOpKind { TOP_LOAD_HEIGHTS = 0, TOP_GHOST_DATA = 1, TOP_ACTIVATE_FRESH = 2, TOP_ACTIVATE_GHOST = 3 };
struct TileOp { int tile0, kind, mask; };   // 0-based global tile index; mask: seam bits of a TOP_LOAD_HEIGHTS (tile_vertices_kernel)

struct TileTable {
   int nXt = 0, nYt = 0;
   bool periodic = false, oneD = false, haltBc = true;
   std::vector<int> tstate;             // 0 untouched, 1 ghost, 2 active
   std::vector<char> loaded;
   std::vector<int> activeList, ghostList;   // 1-based ids; active ascending (utilities.f90:260)
   std::vector<TileOp> ops;             // device work requested so far, in order
   long ntilesAdded = 0;
   bool haltViolation = false;          // a tile outside the domain was requested with Boundary Conditions = halt

   void init(int nxt, int nyt, bool per, bool one, bool halt) {
      nXt = nxt; nYt = nyt; periodic = per; oneD = one; haltBc = halt;
      tstate.assign((size_t)nXt * nYt, 0); loaded.assign((size_t)nXt * nYt, 0);
      activeList.clear(); ghostList.clear(); ops.clear(); ntilesAdded = 0; haltViolation = false;
   }
   int nTiles() const { return nXt * nYt; }
   int W(int t0) const { int tx = t0 % nXt, ty = t0 / nXt; return tx == 0 ? (periodic ? (nXt - 1) + ty * nXt : -1) : t0 - 1; }
   int E(int t0) const { int tx = t0 % nXt, ty = t0 / nXt; return tx == nXt - 1 ? (periodic ? ty * nXt : -1) : t0 + 1; }
   int S(int t0) const { int tx = t0 % nXt, ty = t0 / nXt; return ty == 0 ? (periodic ? tx + (nYt - 1) * nXt : -1) : t0 - nXt; }
   int N(int t0) const { int tx = t0 % nXt, ty = t0 / nXt; return ty == nYt - 1 ? (periodic ? tx : -1) : t0 + nXt; }
   bool onDomainEdge(int t0) const {   // Grid.f90:322-335
      int tx = t0 % nXt, ty = t0 / nXt;
      bool on = (tx == 0 || tx == nXt - 1);
      return on || (nYt > 1 && (ty == 0 || ty == nYt - 1));
   }
   // GetHeights + the ghost tiles whose seam was refreshed (loadHeights of kestrel_gpu.cu)
   void loadHeights(int t0, bool given) {
      if (loaded[t0] && !given) return;
      // EqualiseTopographicBoundaryData (MorphodynamicRHS.f90:588-687): a shared vertex keeps the value of the tile for
      // which it is local index 1, so the east column / north row / north-east corner are only written while the tile
      // across them has no heights yet
      int tE = E(t0), tN = oneD ? -1 : N(t0);
      int tNE = (tE >= 0 && !oneD) ? N(tE) : -1;
      bool eL = tE >= 0 && tE != t0 && loaded[tE], nL = tN >= 0 && tN != t0 && loaded[tN], neL = tNE >= 0 && loaded[tNE];
      int mask = 0;
      if (!eL && !(periodic && nXt == 1)) mask |= 1;
      if (!nL && !(periodic && nYt == 1)) mask |= 2;
      if (!(eL || nL || neL) && (mask & 1) && (mask & 2)) mask |= 4;
      ops.push_back({t0, TOP_LOAD_HEIGHTS, mask});
      loaded[t0] = 1;
      int tW = W(t0), tS = oneD ? -1 : S(t0);
      int tSW = (tW >= 0 && !oneD) ? S(tW) : -1;
      for (int tt : {tW, tS, tSW})
         if (tt >= 0 && tt != t0 && loaded[tt] && tstate[tt] == 1) ops.push_back({tt, TOP_GHOST_DATA, 0});
   }
   // UpdateTiles.f90:389-481
   bool addGhostTiles(int t0) {
      int nb[8], n = 0;
      nb[n++] = W(t0); nb[n++] = E(t0);
      if (!oneD) {
         nb[n++] = N(t0); nb[n++] = S(t0);
         if (!periodic) {
            int s = S(t0), nn = N(t0);
            nb[n++] = s >= 0 ? W(s) : -1; nb[n++] = s >= 0 ? E(s) : -1;
            nb[n++] = nn >= 0 ? W(nn) : -1; nb[n++] = nn >= 0 ? E(nn) : -1;
         }
      }
      for (int k = 0; k < n; k++) {
         int tt = nb[k];
         if (tt < 0) return false;   // ghost tile out of bounds (UpdateTiles.f90:423)
         if (tstate[tt] != 0) continue;
         tstate[tt] = 1;
         ghostList.push_back(tt + 1);
         loadHeights(tt, false);
         ops.push_back({tt, TOP_GHOST_DATA, 0});
      }
      return true;
   }
   // AddTile (UpdateTiles.f90:56-78)
   bool addTile(int t0, bool countIt, bool heightsGiven = false) {
      if (t0 < 0 || t0 >= nTiles() || (onDomainEdge(t0) && !periodic)) {
         if (haltBc) { haltViolation = true; return false; }
         return true;
      }
      if (tstate[t0] == 2) return true;
      bool wasGhost = tstate[t0] == 1;
      tstate[t0] = 2;
      activeList.insert(std::upper_bound(activeList.begin(), activeList.end(), t0 + 1), t0 + 1);
      if (wasGhost) ghostList.erase(std::find(ghostList.begin(), ghostList.end(), t0 + 1));
      loadHeights(t0, heightsGiven);
      ops.push_back({t0, wasGhost ? TOP_ACTIVATE_GHOST : TOP_ACTIVATE_FRESH, 0});
      bool ok = addGhostTiles(t0);
      if (countIt) ntilesAdded++;
      return ok;
   }
   // CheckIfNearBoundaries (TimeStepper.f90:924-1150).  flags: per GLOBAL tile, bit 0 N (wet cell with jj > nY - buf),
   // bit 1 S, bit 2 E, bit 3 W; only the entries of active tiles are read.
   bool replay(const int *flags, int nX, int nY, int buf) {
      if ((int)activeList.size() == nTiles() || activeList.empty()) return true;
      for (int dir = (oneD ? 2 : 0); dir < 4; dir++) {
         int trip = (int)activeList.size();   // fixed at loop entry while the list grows (quirk Q3)
         for (int tt = 0; tt < trip; tt++) {
            int t0 = activeList[tt] - 1;
            int tx = t0 % nXt, ty = t0 / nXt;
            int nbr;
            switch (dir) {
               case 0: nbr = (periodic && ty == nYt - 1) ? tx : t0 + nXt; break;
               case 1: nbr = (periodic && ty == 0) ? tx + (nYt - 1) * nXt : t0 - nXt; break;
               case 2: nbr = (periodic && tx == nXt - 1) ? ty * nXt : t0 + 1; break;
               default: nbr = (periodic && tx == 0) ? (nXt - 1) + ty * nXt : t0 - 1; break;
            }
            if (nbr + 1 <= 0) continue;
            if (nbr < nTiles() && tstate[nbr] == 2) continue;
            int f = flags[t0];
            bool trig;
            switch (dir) {
               case 0: trig = (f & 1) || (0 > nY - buf); break;
               case 1: trig = (f & 2) || (nY <= buf); break;
               case 2: trig = (f & 4) || (0 > nX - buf); break;
               default: trig = (f & 8) || (nX <= buf); break;
            }
            if (trig && !addTile(nbr, true)) return false;
         }
      }
      return true;
   }
};

}  // namespace kgpu
