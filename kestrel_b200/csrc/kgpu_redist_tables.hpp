// kgpu_redist_tables.hpp -- host bookkeeping of RedistributeGrid across ranks (no CUDA in this file).
//
// Input: the redistribution lists of all ranks as gathered (rank-major, M slots per rank, counts[r] valid),
// with LOCAL cell indices.  Output: the global walk order of the reference (ascending excess, ties in scan
// order = ascending global tile, then j, then i; Redistribute.f90:69-101) and, for every entry, the offsets
// into the gathered patch array of the 4 x 4 vertices and 3 x 3 cells around its cell.  A vertex or cell
// that several patches hold (periodic images included) resolves to ONE canonical slot: the copy in the
// first patch of the walk that contains it.  Every rank runs this on identical input, so the tables -- and
// therefore the replicated walk -- are identical everywhere.
#pragma once
#include <algorithm>
#include <unordered_map>
#include <vector>

namespace kgpu {

constexpr int RT_V = 16, RT_C = 9, RT_DOUBLES = 93, RT_B0 = 0, RT_W0 = 48;   // patch layout (kgpu_morpho.cuh RP_*)

struct RedistTables {
   std::vector<int> patch;          // [n] gathered patch index (r * M + k) of the entries in walk order
   std::vector<int> gi, gj;         // [n] global cell of each entry
   std::vector<int> vslot, cslot;   // [n][16], [n][9] offsets into the patch array (doubles)
   std::vector<int> vkey, vbase;    // unique vertices: (gi, gj) pairs and their canonical offset
   std::vector<int> ckey, cbase;    // unique cells
};

struct RedistGeometry {
   int R, M;            // ranks, slots per rank in the gathered arrays
   int px;              // ranks per row of the decomposition (1 for a single device)
   int NX, NY;          // cells of one rank's block
   int nX, nY;          // cells per tile
   int gnXt, gnYt;      // tiles of the whole domain
   int oneD;
};

inline void buildRedistTables(const RedistGeometry &g, const int *counts, const double *excess, const int *li, const int *lj, RedistTables &T) {
   struct GEntry { double excess; int gi, gj, patch; };
   std::vector<GEntry> ge;
   const int NXg = g.nX * g.gnXt, NYg = g.oneD ? 1 : g.nY * g.gnYt;
   for (int r = 0; r < g.R; r++) {
      const int ox = (r % g.px) * g.NX, oy = (r / g.px) * g.NY;
      for (int k = 0; k < counts[r]; k++) {
         const size_t s = (size_t)r * g.M + k;
         ge.push_back({excess[s], li[s] + ox, g.oneD ? 0 : lj[s] + oy, (int)s});
      }
   }
   const int nX = g.nX, nY = g.nY, gnXt = g.gnXt;
   std::stable_sort(ge.begin(), ge.end(), [&](const GEntry &x, const GEntry &y) {
      if (x.excess != y.excess) return x.excess < y.excess;
      int tx_ = (x.gi / nX) + (x.gj / nY) * gnXt, ty_ = (y.gi / nX) + (y.gj / nY) * gnXt;
      if (tx_ != ty_) return tx_ < ty_;
      if (x.gj != y.gj) return x.gj < y.gj;
      return x.gi < y.gi;
   });
   const int n = (int)ge.size();
   T = RedistTables();
   T.patch.resize(n); T.gi.resize(n); T.gj.resize(n);
   T.vslot.resize((size_t)n * RT_V); T.cslot.resize((size_t)n * RT_C);
   std::unordered_map<long long, int> vmap, cmap;
   auto wrap = [](int i, int m) { return ((i % m) + m) % m; };
   for (int e = 0; e < n; e++) {
      T.patch[e] = ge[e].patch; T.gi[e] = ge[e].gi; T.gj[e] = ge[e].gj;
      const int pbase = ge[e].patch * RT_DOUBLES;
      for (int q = 0; q < RT_V; q++) {
         int vi = wrap(ge[e].gi - 1 + q % 4, NXg), vj = g.oneD ? 0 : wrap(ge[e].gj - 1 + q / 4, NYg);
         long long key = (long long)vj * NXg + vi;
         auto it = vmap.find(key);
         if (it == vmap.end()) {
            it = vmap.emplace(key, pbase + RT_B0 + q).first;
            T.vkey.push_back(vi); T.vkey.push_back(vj); T.vbase.push_back(it->second);
         }
         T.vslot[(size_t)e * RT_V + q] = it->second;
      }
      for (int q = 0; q < RT_C; q++) {
         int ci = wrap(ge[e].gi - 1 + q % 3, NXg), cj = g.oneD ? 0 : wrap(ge[e].gj - 1 + q / 3, NYg);
         long long key = (long long)cj * NXg + ci;
         auto it = cmap.find(key);
         if (it == cmap.end()) {
            it = cmap.emplace(key, pbase + RT_W0 + q).first;
            T.ckey.push_back(ci); T.ckey.push_back(cj); T.cbase.push_back(it->second);
         }
         T.cslot[(size_t)e * RT_C + q] = it->second;
      }
   }
}

}  // namespace kgpu
