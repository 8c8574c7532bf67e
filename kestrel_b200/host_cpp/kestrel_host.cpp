// kestrel_host.cpp -- see kestrel_host.hpp.  Compiled with -ffp-contract=off: every formula keeps
// the reference's operation order so that the tiles handed to kgpu_upload_tile carry the same
// bits as the Fortran host's.
#include "kestrel_host.hpp"

#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>

namespace kestrel {

static const double PI = 3.141592653589793238462643383279502884;
static const double VISC_W = 1.2e-6;  // Parameters.f90:76

// ------------------------------------------------------------------ small string helpers
static std::string strip(const std::string &s) {
   size_t a = 0, b = s.size();
   while (a < b && std::isspace((unsigned char)s[a])) a++;
   while (b > a && std::isspace((unsigned char)s[b - 1])) b--;
   return s.substr(a, b - a);
}
static std::string lower(std::string s) {
   for (auto &c : s) c = (char)std::tolower((unsigned char)c);
   return s;
}
static double readReal(const std::string &s0) {  // Fortran reals: 1.0d-3
   std::string s = strip(s0);
   for (auto &c : s)
      if (c == 'd' || c == 'D') c = 'e';
   char *end = nullptr;
   double v = std::strtod(s.c_str(), &end);
   if (end == s.c_str()) throw FatalError("Could not read a real number from '" + s0 + "'");
   return v;
}
static int readInt(const std::string &s) {
   try {
      return std::stoi(strip(s));
   } catch (...) {
      throw FatalError("Could not read an integer from '" + s + "'");
   }
}
static std::vector<double> readSet(const std::string &s0) {  // "(a, b, c)"
   std::string s = strip(s0);
   if (!s.empty() && s.front() == '(') s = s.substr(1);
   if (!s.empty() && s.back() == ')') s.pop_back();
   std::vector<double> out;
   std::stringstream ss(s);
   std::string item;
   while (std::getline(ss, item, ','))
      if (!strip(item).empty()) out.push_back(readReal(item));
   return out;
}

// ------------------------------------------------------------------ ReadInputFile
RunSet ReadInputFile(const std::string &path, std::vector<std::string> *warnings) {
   std::ifstream fh(path);
   if (!fh) throw FatalError("Could not open input file " + path);
   typedef std::vector<std::pair<std::string, std::string>> KV;
   std::map<std::string, KV> blocks;
   std::vector<std::map<std::string, std::string>> caps, cubes, srcs;
   const char *names[] = {"Domain", "Source", "Cap", "Cube", "Parameters", "Solver", "Output", "Topog"};
   std::string block, raw;
   auto warn = [&](const std::string &k) {
      if (warnings) warnings->push_back("Input label unrecognized: " + k);  // Messages.f90:299
   };
   while (std::getline(fh, raw)) {
      std::string line = strip(raw);
      size_t hash = line.find('#');
      if (hash != std::string::npos) line = strip(line.substr(0, hash));
      if (line.empty() || line[0] == '%' || line[0] == '#') continue;
      size_t colon = line.find(':');
      if (colon != std::string::npos) {
         std::string name = strip(line.substr(0, colon));
         for (const char *n : names)
            if (name == n) {
               block = name;
               if (name == "Cap") caps.emplace_back();
               else if (name == "Cube") cubes.emplace_back();
               else if (name == "Source") srcs.emplace_back();
            }
         continue;
      }
      size_t eq = line.find('=');
      if (eq == std::string::npos || block.empty()) continue;
      std::string key = strip(line.substr(0, eq)), val = strip(line.substr(eq + 1));
      if (block == "Cap") caps.back()[key] = val;
      else if (block == "Cube") cubes.back()[key] = val;
      else if (block == "Source") srcs.back()[key] = val;
      else blocks[block].emplace_back(lower(key), val);
   }

   RunSet rs;
   // ---- Domain (DomainSettings.f90:86-153)
   for (auto &kv : blocks["Domain"]) {
      const std::string &k = kv.first, &v = kv.second;
      if (k == "nxtiles") rs.nXtiles = readInt(v);
      else if (k == "nytiles") rs.nYtiles = readInt(v);
      else if (k == "nxpertile") rs.nXpertile = readInt(v);
      else if (k == "nypertile") rs.nYpertile = readInt(v);
      else if (k == "xtilesize") { rs.Xtilesize = readReal(v); rs.hasXtilesize = true; }
      else if (k == "ytilesize") { rs.Ytilesize = readReal(v); rs.hasYtilesize = true; rs.hasXtilesize = false; }
      else if (k == "boundary conditions") rs.bcs = lower(v);
      else if (k == "boundary hn") rs.bcsHnval = readReal(v);
      else if (k == "boundary u") rs.bcsuval = readReal(v);
      else if (k == "boundary v") rs.bcsvval = readReal(v);
      else if (k == "boundary psi") rs.bcspsival = readReal(v);
      else if (k == "lat" || k == "latitude" || k == "lon" || k == "longitude") {}
      else warn(k);
   }
   if (!rs.hasXtilesize) {
      rs.Xtilesize = rs.Ytilesize * double(rs.nXpertile) / double(rs.nYpertile);
      rs.hasXtilesize = true;
   }

   // ---- Parameters (Parameters.f90:171-467)
   const std::map<std::string, double RunSet::*> pmap = {
      {"g", &RunSet::g}, {"chezy co", &RunSet::ChezyCo}, {"manning co", &RunSet::ManningCo}, {"coulomb co", &RunSet::CoulombCo},
      {"pouliquen min", &RunSet::PouliquenMinSlope}, {"pouliquen max", &RunSet::PouliquenMaxSlope},
      {"pouliquen intermediate", &RunSet::PouliquenIntermediateSlope}, {"pouliquen beta", &RunSet::PouliquenBeta},
      {"edwards2019 betastar", &RunSet::Edwards2019betastar}, {"edwards2019 kappa", &RunSet::Edwards2019kappa},
      {"edwards2019 gamma", &RunSet::Edwards2019Gamma}, {"voellmy switch rate", &RunSet::VoellmySwitchRate},
      {"voellmy switch value", &RunSet::VoellmySwitchValue}, {"erosion rate", &RunSet::EroRate},
      {"granular erosion rate", &RunSet::EroRateGranular}, {"erosion depth", &RunSet::EroDepth},
      {"erosion critical height", &RunSet::EroCriticalHeight}, {"bed porosity", &RunSet::BedPorosity}, {"rhow", &RunSet::rhow},
      {"rhos", &RunSet::rhos}, {"maxpack", &RunSet::maxPack}, {"max pack", &RunSet::maxPack},
      {"solid diameter", &RunSet::SolidDiameter}, {"eddy viscosity", &RunSet::EddyViscosity}};
   for (auto &kv : blocks["Parameters"]) {
      const std::string &k = kv.first, &v = kv.second;
      auto it = pmap.find(k);
      if (it != pmap.end()) rs.*(it->second) = readReal(v);
      else if (k == "settling speed") { rs.ws0 = readReal(v); rs.hasWs0 = true; }
      else if (k == "drag") rs.drag = lower(v);
      else if (k == "erosion") rs.erosion = lower(v);
      else if (k == "deposition") rs.deposition = lower(v);
      else if (k == "erosion transition") rs.erosion_transition = lower(v);
      else if (k == "morphodynamic damping") rs.morpho_damp = lower(v);
      else if (k == "switch function") rs.fswitch = lower(v);
      else if (k == "iverson" || k == "geometric factors") {
         if (lower(v) == "off") rs.geometric_factors = false;
         else if (lower(v) == "on") rs.geometric_factors = true;
      } else warn(k);
   }

   // ---- Solver (SolverSettings.f90:94-173)
   for (auto &kv : blocks["Solver"]) {
      const std::string &k = kv.first, &v = kv.second;
      if (k == "t end") rs.tend = readReal(v);
      else if (k == "t start") rs.tstart = readReal(v);
      else if (k == "limiter") rs.limiter = lower(v);
      else if (k == "height threshold") rs.heightThreshold = readReal(v);
      else if (k == "tile buffer") rs.TileBuffer = readInt(v);
      else if (k == "cfl") { rs.cfl = readReal(v); rs.hasCfl = true; }
      else if (k == "max dt") rs.maxdt = readReal(v);
      else if (k == "sponge strength") rs.SpongeStrength = readReal(v);
      else if (k == "restart" || k == "initial condition") {}
      else warn(k);
   }
   // ---- Output (OutputSettings.f90), Topog (TopogSettings.f90:92-122)
   for (auto &kv : blocks["Output"]) {
      if (kv.first == "n out") rs.Nout = readInt(kv.second);
      else if (kv.first == "directory") rs.out_dir = kv.second;
   }
   for (auto &kv : blocks["Topog"]) {
      if (kv.first == "type") rs.topog_type = lower(kv.second);
      else if (kv.first == "topog function") rs.topog_func = lower(kv.second);
      else if (kv.first == "topog params") rs.topog_params = readSet(kv.second);
   }
   rs.Finalize();

   auto get = [](const std::map<std::string, std::string> &m, const char *k, const char *dflt) {
      auto it = m.find(k);
      return it == m.end() ? std::string(dflt) : it->second;
   };
   // ---- Caps (InitConds.f90:254-420)
   for (auto &c : caps) {
      Cap cap;
      cap.x = readReal(get(c, "capX", "0")); cap.y = readReal(get(c, "capY", "0"));
      cap.u = readReal(get(c, "capU", "0")); cap.v = readReal(get(c, "capV", "0"));
      cap.psi = readReal(get(c, "capConc", "0"));
      std::string shp = lower(get(c, "capShape", "para"));  // capShape_d = 'para', InitConds.f90:38
      cap.shape = (shp == "flat" || shp == "level") ? shp : "para";
      bool hasR = c.count("capRadius"), hasH = c.count("capHeight"), hasV = c.count("capVolume");
      if (hasR) cap.radius = readReal(c["capRadius"]);
      if (hasH) cap.height = readReal(c["capHeight"]);
      if (hasV && !(hasR && hasH)) cap.volume = readReal(c["capVolume"]);
      double f = cap.shape == "flat" ? 1.0 : 0.5;
      if (hasR && hasV && !hasH) cap.height = cap.volume / f / PI / cap.radius / cap.radius;
      if (hasH && hasV && !hasR) cap.radius = std::sqrt(cap.volume / f / PI / cap.height);
      rs.caps.push_back(cap);
   }
   // ---- Cubes (InitConds.f90:526-773)
   for (auto &c : cubes) {
      Cube cu;
      cu.x = readReal(get(c, "cubeX", "0")); cu.y = readReal(get(c, "cubeY", "0"));
      cu.length = readReal(get(c, "cubeLength", "0")); cu.width = readReal(get(c, "cubeWidth", "0"));
      cu.height = readReal(get(c, "cubeHeight", "0"));
      cu.u = readReal(get(c, "cubeU", "0")); cu.v = readReal(get(c, "cubeV", "0"));
      cu.psi = readReal(get(c, "cubeConc", "0"));
      cu.shape = lower(get(c, "cubeShape", "flat"));
      rs.cubes.push_back(cu);
   }
   // ---- Sources (InitConds.f90:46-200)
   for (auto &s : srcs) {
      FluxSource fs;
      fs.x = readReal(get(s, "sourceX", "0")); fs.y = readReal(get(s, "sourceY", "0"));
      if (!s.count("sourceRadius") || !s.count("sourceTime") || !s.count("sourceFlux") || !s.count("sourceConc"))
         throw FatalError("Source block needs sourceRadius, sourceTime, sourceFlux and sourceConc");
      fs.radius = readReal(s["sourceRadius"]);
      fs.time = readSet(s["sourceTime"]); fs.flux = readSet(s["sourceFlux"]); fs.psi = readSet(s["sourceConc"]);
      if (fs.time.size() != fs.flux.size() || fs.time.size() != fs.psi.size())
         throw FatalError("sourceTime/Flux/Conc sets differ in length");
      rs.sources.push_back(fs);
   }
   return rs;
}

// ------------------------------------------------------------------ derived constants
void RunSet::Finalize() {
   if (!hasYtilesize) Ytilesize = Xtilesize * double(nYpertile) / double(nXpertile);
   nTiles = nXtiles * nYtiles;
   xSize = nXtiles * Xtilesize;
   ySize = nYtiles * Ytilesize;
   NX = nXpertile * nXtiles;
   NY = nYpertile * nYtiles;
   isOneD = (nYtiles * nYpertile == 1);
   deltaX = Xtilesize / double(nXpertile);
   deltaY = Ytilesize / double(nYpertile);
   deltaXRecip = 1.0 / deltaX;
   deltaYRecip = 1.0 / deltaY;
   if (!hasCfl) cfl = isOneD ? 0.5 : 0.25;  // SolverSettings.f90:189-195
   MorphodynamicsOn = lower(erosion) != "off";
   SpongeLayer = bcs == "sponge";
   // Parameters.f90:629-648
   gred = (rhos / rhow - 1.0) * g;
   Rep = std::sqrt(g * SolidDiameter) * SolidDiameter / VISC_W;
   double R = std::pow(gred / VISC_W / VISC_W, 1.0 / 3.0) * SolidDiameter;
   if (!hasWs0) ws0 = VISC_W / SolidDiameter * (std::sqrt(10.36 * 10.36 + 1.048 * R * R * R) - 10.36);
   nsettling = (4.7 + 0.41 * std::pow(Rep, 0.75)) / (1.0 + 0.175 * std::pow(Rep, 0.75));
   CriticalShields = 0.3 / (1.0 + 1.2 * R) + 0.055 * (1.0 - std::exp(-0.02 * R));
   diffusiveTimeScale = std::numeric_limits<double>::max();
   if (EddyViscosity > 0.0) diffusiveTimeScale = std::min(deltaX * deltaX / EddyViscosity, deltaY * deltaY / EddyViscosity);
   DeltaT = (tend - tstart) / Nout;  // OutputSettings.f90:167
}

static int lookup(const std::map<std::string, int> &m, const std::string &key, const char *what) {
   auto it = m.find(lower(key));
   if (it == m.end()) throw FatalError(std::string("Unrecognised ") + what + " '" + key + "'");
   return it->second;
}

kgpu_params RunSet::ToParams(std::vector<kgpu_source> &srcs, kgpu_heights_fn cb, void *ctx) const {
   static const std::map<std::string, int> BCS = {{"halt", 0}, {"periodic", 1}, {"dirichlet", 2}, {"sponge", 3}};
   static const std::map<std::string, int> LIM = {{"minmod1", 0}, {"minmod2", 1}, {"none", 2}, {"van albada", 3}, {"albada", 3}, {"weno", 4}};
   static const std::map<std::string, int> DRAG = {{"chezy", 0}, {"coulomb", 1}, {"voellmy", 2}, {"pouliquen", 3},
                                                   {"edwards2019", 4}, {"variable", 5}, {"manning", 6}};
   static const std::map<std::string, int> ERO = {{"off", 0}, {"simple", 1}, {"fluid", 2}, {"granular", 3}, {"mixed", 4}, {"on", 4}};
   static const std::map<std::string, int> DEP = {{"none", 0}, {"simple", 1}, {"spearman manning", 2}};
   static const std::map<std::string, int> TRANS = {{"smooth", 0}, {"step", 1}, {"off", 2}};
   static const std::map<std::string, int> DAMP = {{"none", 0}, {"off", 0}, {"tanh", 1}, {"rat3", 2}};
   static const std::map<std::string, int> SW = {{"tanh", 0}, {"rat3", 1}, {"cos", 2}, {"linear", 3}, {"equal", 4}, {"0.5", 4},
                                                 {"off", 5}, {"0", 5}, {"zero", 5}, {"1", 6}, {"one", 6}, {"step", 7}};
   kgpu_params p;
   std::memset(&p, 0, sizeof(p));
   p.struct_bytes = (int32_t)sizeof(p);
   p.nXpertile = nXpertile; p.nYpertile = nYpertile; p.nXtiles = nXtiles; p.nYtiles = nYtiles;
   p.isOneD = isOneD ? 1 : 0;
   p.deltaX = deltaX; p.deltaY = deltaY; p.xSize = xSize; p.ySize = ySize;
   p.bcs = lookup(BCS, bcs, "boundary condition");
   p.bcsHnval = bcsHnval; p.bcsuval = bcsuval; p.bcsvval = bcsvval; p.bcspsival = bcspsival;
   p.geometric_factors = geometric_factors ? 1 : 0;
   p.MorphodynamicsOn = MorphodynamicsOn ? 1 : 0;
   p.g = g; p.rhow = rhow; p.rhos = rhos; p.gred = gred;
   p.ChezyCo = ChezyCo; p.ManningCo = ManningCo; p.CoulombCo = CoulombCo;
   p.PouliquenMinSlope = PouliquenMinSlope; p.PouliquenMaxSlope = PouliquenMaxSlope;
   p.PouliquenIntermediateSlope = PouliquenIntermediateSlope; p.PouliquenBeta = PouliquenBeta;
   p.Edwards2019betastar = Edwards2019betastar; p.Edwards2019kappa = Edwards2019kappa; p.Edwards2019Gamma = Edwards2019Gamma;
   p.VoellmySwitchRate = VoellmySwitchRate; p.VoellmySwitchValue = VoellmySwitchValue;
   p.EroRate = EroRate; p.EroRateGranular = EroRateGranular; p.CriticalShields = CriticalShields; p.EroDepth = EroDepth;
   p.EroCriticalHeight = EroCriticalHeight;
   p.BedPorosity = BedPorosity; p.maxPack = maxPack; p.SolidDiameter = SolidDiameter; p.ws0 = ws0; p.nsettling = nsettling;
   p.EddyViscosity = EddyViscosity;
   p.heightThreshold = heightThreshold; p.cfl = cfl; p.diffusiveTimeScale = diffusiveTimeScale; p.maxdt = maxdt; p.tstart = tstart;
   p.TileBuffer = TileBuffer; p.SpongeLayer = SpongeLayer ? 1 : 0; p.SpongeStrength = SpongeStrength;
   p.limiter = lookup(LIM, limiter, "limiter");
   p.drag = lookup(DRAG, drag, "drag");
   p.erosion = lookup(ERO, erosion, "erosion");
   p.deposition = lookup(DEP, deposition, "deposition");
   p.erosion_transition = lookup(TRANS, erosion_transition, "erosion transition");
   p.morpho_damp = lookup(DAMP, morpho_damp, "morphodynamic damping");
   p.fswitch = lookup(SW, fswitch, "switch function");
   srcs.clear();
   for (const auto &s : sources) {
      kgpu_source k;
      k.x = s.x; k.y = s.y; k.radius = s.radius;
      k.num_cells_in_src = s.numCellsInSrc;
      k.n_series = (int32_t)s.time.size();
      k.time = s.time.data(); k.flux = s.flux.data(); k.psi = s.psi.data();
      srcs.push_back(k);
   }
   p.n_sources = (int32_t)srcs.size();
   p.sources = srcs.empty() ? nullptr : srcs.data();
   p.heights = cb;
   p.heights_ctx = ctx;
   p.device = device;
   p.arithmetic = arithmetic;
   p.comm_rank = 0; p.comm_size = 1; p.comm_px = 1; p.comm_py = 1;
   return p;
}

// ------------------------------------------------------------------ coordinates and topography
void TileCoords(const RunSet &rs, int tileId, std::vector<double> &x, std::vector<double> &y, std::vector<double> &xv,
                std::vector<double> &yv) {
   const int gi = (tileId - 1) % rs.nXtiles + 1, gj = (tileId - 1) / rs.nXtiles + 1;
   x.resize(rs.nXpertile); y.resize(rs.nYpertile);
   for (int i = 1; i <= rs.nXpertile; i++) x[i - 1] = -0.5 * rs.xSize + rs.deltaX * ((gi - 1.0) * rs.nXpertile + (i - 0.5));
   for (int j = 1; j <= rs.nYpertile; j++) y[j - 1] = -0.5 * rs.ySize + rs.deltaY * ((gj - 1.0) * rs.nYpertile + (j - 0.5));
   xv.resize(rs.nXpertile + 1); yv.resize(rs.nYpertile + 1);
   for (int i = 0; i < rs.nXpertile; i++) xv[i] = x[i] - 0.5 * rs.deltaX;
   xv[rs.nXpertile] = x[rs.nXpertile - 1] + 0.5 * rs.deltaX;
   for (int j = 0; j < rs.nYpertile; j++) yv[j] = y[j] - 0.5 * rs.deltaY;
   yv[rs.nYpertile] = y[rs.nYpertile - 1] + 0.5 * rs.deltaY;
}

double TopogFunction(const RunSet &rs, double X, double Y) {  // TopogFuncs.f90
   const std::string &name = rs.topog_func;
   const std::vector<double> &p = rs.topog_params;
   auto need = [&](size_t n) {
      if (p.size() < n) throw FatalError("Topog function '" + name + "' needs " + std::to_string(n) + " Topog params");
   };
   if (name == "flat") return 0.0;
   if (name == "xslope") { need(1); return p[0] * X; }
   if (name == "yslope") { need(1); return p[0] * Y; }
   if (name == "xyslope") { need(2); return p[0] * X + p[1] * Y; }
   if (name == "xsinslope") { need(1); double Lx = rs.Xtilesize * rs.nXtiles; return p[0] * std::sin(X * (2.0 * PI / Lx)); }
   if (name == "xysinslope") {
      need(1);
      double Lx = rs.Xtilesize * rs.nXtiles, Ly = rs.Ytilesize * rs.nYtiles;
      return p[0] * std::sin(X * (2.0 * PI / Lx)) * std::sin(Y * (2.0 * PI / Ly));
   }
   if (name == "xhump") { need(2); double A = p[0], L = p[1]; return (X > -L && X < L) ? 0.5 * A * (1.0 + std::cos(PI * X / L)) : 0.0; }
   if (name == "xtanh") { need(3); return p[1] * (1.0 + std::tanh((X - p[0]) / p[2])); }
   if (name == "xparab") { need(1); return p[0] * X * X; }
   if (name == "xyparab") { need(2); return p[0] * X * X + p[1] * Y * Y; }
   if (name == "xbislope") {
      need(3);
      double phi1 = p[0] * PI / 180.0, phi2 = p[1] * PI / 180.0, lam = p[2];
      double a1 = std::tan(phi1), a2 = std::tan(phi2);
      return -0.5 * (a1 + a2) * X + 0.5 * (a1 - a2) * lam * std::log(std::cosh(X / lam));
   }
   if (name == "usgs" || name == "flume") {  // TopogFuncs.f90:145-242: two slopes joined by a cosh arc, tanh side walls
      double theta0 = 31.0, theta1 = 2.4, xwall = 8.5, wallW = 2.0, wallH, sigma;
      if (name == "usgs") { need(2); wallH = p[0]; sigma = p[1]; }
      else { need(6); theta0 = p[0]; theta1 = p[1]; xwall = p[2]; wallW = p[3]; wallH = p[4]; sigma = p[5]; }
      double alpha = 8.5 / (std::asinh(-std::tan(4.0 * PI / 180.0)) - std::asinh(-std::tan(theta0 * PI / 180.0)));
      double xc0 = -alpha * std::asinh(-std::tan(theta0 * PI / 180.0));
      double zc0 = -alpha * std::cosh((-xc0) / alpha);
      double x1 = xc0 + alpha * std::asinh(-std::tan(theta1 * PI / 180.0));
      double b;
      if (X < 0.0) b = -std::tan(theta0 * PI / 180.0) * X;
      else if (X > x1) b = zc0 + alpha * std::cosh((x1 - xc0) / alpha) - std::tan(theta1 * PI / 180.0) * (X - x1);
      else b = zc0 + alpha * std::cosh((X - xc0) / alpha);
      if (X < xwall)
         b = b + 0.5 * wallH * (std::tanh(sigma * (Y - 0.5 * wallW)) - std::tanh(sigma * (Y - 1.5 * wallW)) +
                                std::tanh(sigma * (Y + 1.5 * wallW)) - std::tanh(sigma * (Y + 0.5 * wallW)));
      return b;
   }
   if (name == "channel power law" || name == "channel_powerlaw") {  // TopogFuncs.f90:255-276
      need(3);
      double costheta = std::cos(std::atan(p[0]));
      return p[0] * X + costheta * std::pow(std::fabs(Y) / p[1], p[2]);
   }
   if (name == "channel trapezium" || name == "channel_trapezium") {  // TopogFuncs.f90:288-308
      need(3);
      double costheta = std::cos(std::atan(p[0]));
      return p[0] * X + costheta * std::max(0.0, p[2] * (std::fabs(Y) - 0.5 * p[1]));
   }
   if (name == "xtrislope") {  // TopogFuncs.f90:346-388
      need(6);
      double phi1 = p[0] * PI / 180.0, phi2 = p[1] * PI / 180.0, phi3 = p[2] * PI / 180.0, lam = p[3], x1 = p[4], x2 = p[5];
      double s1 = std::tan(phi1), s2 = std::tan(phi2), s3 = std::tan(phi3);
      double c2 = (x1 - 0.5 * lam) * 0.5 * (s1 - s2);
      double c3 = (x1 + 0.5 * lam) * 0.5 * (s1 - s2) + c2;
      double c4 = (x2 - 0.5 * lam) * 0.5 * (s2 - s3) + c3;
      double c5 = (x2 + 0.5 * lam) * 0.5 * (s2 - s3) + c4;
      if (X < x1 - 0.5 * lam) return s1 * X;
      if (X < x1 + 0.5 * lam) { double A = 0.5 * (s2 - s1) * lam / PI; return A * std::sin((X - x1) * PI / lam - 0.5 * PI) + 0.5 * (s1 + s2) * X + c2; }
      if (X < x2 - 0.5 * lam) return s2 * X + c3;
      if (X < x2 + 0.5 * lam) { double A = 0.5 * (s3 - s2) * lam / PI; return A * std::sin((X - x2) * PI / lam - 0.5 * PI) + 0.5 * (s2 + s3) * X + c4; }
      return s3 * X + c5;
   }
   if (name == "x2slopes") {
      need(3);
      double alpha = p[0], beta = p[1], R = p[2];
      double sa = std::sqrt(1.0 + alpha * alpha), sb = std::sqrt(1.0 + beta * beta);
      double xc0 = (sa - sb) * R / (alpha - beta), zc0 = (alpha * sb - beta * sa) * R / (alpha - beta);
      double x1 = xc0 - alpha * R / sa, x2 = xc0 - beta * R / sb;
      if (X < x1) return -alpha * X;
      if (X > x2) return -beta * X;
      return zc0 - std::sqrt(std::max(R * R - (X - xc0) * (X - xc0), 0.0));
   }
   throw FatalError("topography function '" + name + "' is not available (usgs, flume, channel_*, xtrislope and raster DEMs are out of scope)");
}

void GetHeights(const RunSet &rs, int tileId, double *b0) {  // dem.f90:360-415, Type = Function
   if (rs.topog_type != "function") throw FatalError("Topog Type '" + rs.topog_type + "' needs GDAL rasters, which stay with the Fortran host");
   std::vector<double> x, y, xv, yv;
   TileCoords(rs, tileId, x, y, xv, yv);
   const int nvx = rs.nXpertile + 1, nvy = rs.nYpertile + 1;
   for (int j = 0; j < nvy; j++)
      for (int i = 0; i < nvx; i++) b0[(size_t)j * nvx + i] = TopogFunction(rs, xv[i], yv[j]);
}

// ------------------------------------------------------------------ LoadSourceConditions
static double kahan(const double *t, int n) {  // utilities.f90:418-448
   double s = 0.0, c = 0.0;
   for (int k = 0; k < n; k++) {
      double y = t[k] - c;
      double tt = s + y;
      c = (tt - s) - y;
      s = tt;
   }
   return s;
}

static Tile makeTile(const RunSet &rs, int id) {  // AllocateTile + ActivateTile, UpdateTiles.f90:120-370
   const int nX = rs.nXpertile, nY = rs.nYpertile, nvx = nX + 1;
   Tile T;
   T.id = id;
   T.b0.assign((size_t)(nX + 1) * (nY + 1), 0.0);
   T.bt.assign((size_t)(nX + 1) * (nY + 1), 0.0);
   GetHeights(rs, id, T.b0.data());
   T.u.assign((size_t)13 * nX * nY, 0.0);
   T.maxima.assign((size_t)10 * nX * nY, 0.0);
   T.tfirst.assign((size_t)nX * nY, -1.0);
   // ComputeCellCentredTopographicData (MorphodynamicRHS.f90:308-368), bt = 0
   for (int j = 0; j < nY; j++)
      for (int i = 0; i < nX; i++) {
         double *q = &T.u[((size_t)j * nX + i) * 13];
         double b0c, bx, by;
         if (!rs.isOneD) {
            double a = T.b0[(size_t)j * nvx + i], b = T.b0[(size_t)j * nvx + i + 1], c = T.b0[(size_t)(j + 1) * nvx + i],
                   d = T.b0[(size_t)(j + 1) * nvx + i + 1];
            double t4[4] = {a, b, c, d};
            b0c = 0.25 * kahan(t4, 4);
            double tx[8] = {b, 0.0, -a, -0.0, d, 0.0, -c, -0.0}, ty[8] = {c, 0.0, -a, -0.0, d, 0.0, -b, -0.0};
            bx = 0.5 * rs.deltaXRecip * kahan(tx, 8);
            by = 0.5 * rs.deltaYRecip * kahan(ty, 8);
         } else {
            double a = T.b0[i], b = T.b0[i + 1];
            b0c = 0.5 * (a + b);
            double tx[4] = {b, 0.0, -a, -0.0};
            bx = rs.deltaXRecip * kahan(tx, 4);
            by = 0.0;
         }
         q[iRHO] = rs.rhow;  // AllocateU, UpdateTiles.f90:243
         q[iB0] = b0c; q[iBT] = 0.0; q[iBX] = bx; q[iBY] = by;
         q[iW] = b0c;        // ActivateTile, UpdateTiles.f90:368
      }
   return T;
}

static bool onDomainEdge(const RunSet &rs, int id) {  // Grid.f90:322-335
   int i = (id - 1) % rs.nXtiles + 1, j = (id - 1) / rs.nXtiles + 1;
   bool on = (i == 1 || i == rs.nXtiles);
   return on || (rs.nYtiles > 1 && (j == 1 || j == rs.nYtiles));
}

std::map<int, Tile> LoadSourceConditions(RunSet &rs) {
   std::map<int, Tile> tiles;
   for (auto &s : rs.sources) s.numCellsInSrc = 0;
   const double rhow = rs.rhow, rhos = rs.rhos;
   const int nX = rs.nXpertile, nY = rs.nYpertile;
   const size_t nc = (size_t)nX * nY;
   auto getTile = [&](int k) -> Tile * {
      auto it = tiles.find(k);
      if (it != tiles.end()) return &it->second;
      if (onDomainEdge(rs, k) && rs.bcs != "periodic") {
         if (rs.bcs == "halt") throw FatalError("Error: tried to add a tile outside the domain.");  // UpdateTiles.f90:63-65
         return nullptr;
      }
      return &tiles.emplace(k, makeTile(rs, k)).first->second;
   };
   auto gammaOf = [&](const double *q) { return rs.geometric_factors ? std::sqrt(1.0 + q[iBX] * q[iBX] + q[iBY] * q[iBY]) : 1.0; };
   const int ntx = rs.nXtiles, nty = rs.isOneD ? 1 : rs.nYtiles;
   std::vector<double> x, y, xv, yv;
   std::vector<char> m(nc);
   for (int i = 1; i <= ntx; i++)
      for (int j = 1; j <= nty; j++) {
         const int k = i + (j - 1) * rs.nXtiles;
         TileCoords(rs, k, x, y, xv, yv);
         for (const Cap &cap : rs.caps) {
            const double rho = rhow + (rhos - rhow) * cap.psi;
            bool any = false;
            for (int jj = 0; jj < nY; jj++)
               for (int ii = 0; ii < nX; ii++) {
                  double dx = x[ii] - cap.x, dy = y[jj] - cap.y;
                  double R2 = rs.isOneD ? dx * dx : dx * dx + dy * dy;
                  m[(size_t)jj * nX + ii] = R2 <= cap.radius * cap.radius;  // quirk Q9: <=
                  any = any || m[(size_t)jj * nX + ii];
               }
            if (!any) continue;
            Tile *T = getTile(k);
            if (!T) continue;
            double *Hnmax = &T->maxima[0], *psimax = &T->maxima[(size_t)8 * nc];
            for (int jj = 0; jj < nY; jj++)
               for (int ii = 0; ii < nX; ii++) {
                  const size_t c = (size_t)jj * nX + ii;
                  if (!m[c]) continue;
                  double *q = &T->u[c * 13];
                  const double dx = x[ii] - cap.x, dy = y[jj] - cap.y;
                  const double R2 = rs.isOneD ? dx * dx : dx * dx + dy * dy;
                  const double gam = gammaOf(q);
                  const double HnOrig = q[iHN], rhoOrig = q[iRHO];
                  double Hn = cap.height;
                  if (cap.shape == "flat") {
                     q[iW] += cap.height / gam;
                     q[iHN] += cap.height;
                     Hnmax[c] += cap.height;
                     q[iHU] += rho * cap.height * cap.u;
                     if (rs.isOneD) {  // quirk Q10
                        psimax[c] += cap.psi;
                        q[iU] += cap.u;
                        q[iPSI] += cap.psi;
                     } else q[iHV] += rho * cap.height * cap.v;
                     q[iHPSI] += cap.psi * cap.height;
                  } else if (cap.shape == "para") {
                     const double prof = cap.height * (1.0 - R2 / cap.radius / cap.radius);
                     q[iW] += prof / gam;
                     q[iHN] += prof;
                     Hnmax[c] += prof;
                     q[iHU] += rho * cap.height * cap.u;
                     if (rs.isOneD) psimax[c] += cap.psi;
                     else q[iHV] += rho * cap.height * cap.v;
                     q[iHPSI] += cap.psi * cap.height * (1.0 - R2 / cap.radius / cap.radius);
                  } else {  // level
                     const double hp = cap.height - q[iB0];
                     if (rs.isOneD) {
                        if (hp > 0.0) {
                           q[iW] += hp;
                           Hnmax[c] += hp * gam;
                           q[iHN] += hp * gam;
                           q[iHU] += rho * hp * gam * cap.u;
                           q[iHPSI] += cap.psi * hp * gam;
                           psimax[c] += cap.psi;
                        }
                     } else {
                        Hn = hp * gam;
                        if (Hn > 0.0) {
                           q[iW] += hp;
                           q[iHN] += Hn;
                           Hnmax[c] += Hn;
                           q[iHU] += rho * Hn * cap.u;
                           q[iHV] += rho * Hn * cap.v;
                           q[iHPSI] += cap.psi * Hn;
                        }
                     }
                  }
                  q[iRHO] = (rhoOrig * HnOrig + rho * Hn) / (HnOrig + Hn);  // SetSources.f90:159,206
               }
         }
         for (const Cube &cube : rs.cubes) {
            const double rho = rhow + (rhos - rhow) * cube.psi;
            bool any = false;
            for (int jj = 0; jj < nY; jj++)
               for (int ii = 0; ii < nX; ii++) {
                  bool in = std::fabs(x[ii] - cube.x) <= 0.5 * cube.length;
                  if (!rs.isOneD) in = in && std::fabs(y[jj] - cube.y) <= 0.5 * cube.width;
                  m[(size_t)jj * nX + ii] = in;
                  any = any || in;
               }
            if (!any) continue;
            Tile *T = getTile(k);
            if (!T) continue;
            double *Hnmax = &T->maxima[0];
            for (size_t c = 0; c < nc; c++) {
               if (!m[c]) continue;
               double *q = &T->u[c * 13];
               const double gam = gammaOf(q);
               const double HnOrig = q[iHN], rhoOrig = q[iRHO];
               double Hn = cube.height;
               if (cube.shape == "level") {
                  const double hp = cube.height - q[iB0];
                  Hn = hp * gam;
                  if (Hn > 0.0) {
                     q[iW] += hp;
                     q[iHN] += Hn;
                     Hnmax[c] += Hn;
                     q[iHPSI] += Hn * cube.psi;
                  }
               } else {  // flat
                  q[iW] += cube.height / gam;
                  q[iHN] += cube.height;
                  Hnmax[c] += cube.height;
                  q[iHPSI] += cube.psi * cube.height;
               }
               q[iRHO] = (rhoOrig * HnOrig + rho * Hn) / (HnOrig + Hn);  // SetSources.f90:299,351
            }
         }
         for (FluxSource &s : rs.sources) {
            int count = 0;
            for (int jj = 0; jj < nY; jj++)
               for (int ii = 0; ii < nX; ii++) {
                  double dx = x[ii] - s.x, dy = y[jj] - s.y;
                  double R2 = rs.isOneD ? dx * dx : dx * dx + dy * dy;
                  if (R2 <= s.radius * s.radius) count++;
               }
            if (!count) continue;
            Tile *T = getTile(k);
            if (!T) continue;
            T->containsSource = true;
            s.numCellsInSrc += count;  // SetSources.f90:227,372
         }
      }
   return tiles;
}

// ------------------------------------------------------------------ outputs
VolumeRow CalculateVolume(const RunSet &rs, double t, const std::map<int, Tile> &tiles) {  // Output.f90:617-735
   double s[4] = {0, 0, 0, 0}, c[4] = {0, 0, 0, 0};
   auto add = [&](int k, double x) {
      double yy = x - c[k];
      double tt = s[k] + yy;
      c[k] = (tt - s[k]) - yy;
      s[k] = tt;
   };
   const size_t nc = (size_t)rs.nXpertile * rs.nYpertile;
   for (const auto &kv : tiles) {
      const Tile &T = kv.second;
      for (size_t cell = 0; cell < nc; cell++) {
         const double *q = &T.u[cell * 13];
         const double gam = rs.geometric_factors ? std::sqrt(1.0 + q[iBX] * q[iBX] + q[iBY] * q[iBY]) : 1.0;
         add(0, (q[iW] - q[iB0] - q[iBT]) * gam * gam);
         add(1, q[iRHO] * q[iHN] * gam);
         add(2, q[iBT]);
         add(3, q[iHPSI] * gam);
      }
   }
   const double area = rs.isOneD ? rs.deltaX : rs.deltaX * rs.deltaY;
   const double vol = s[0] * area, mass = s[1] * area, bed = s[2] * area, solids = s[3] * area;
   const double rhob = rs.rhow * rs.BedPorosity + rs.rhos * (1.0 - rs.BedPorosity);
   return {t, vol, bed, mass, bed * rhob, solids * rs.rhos, bed * rs.rhos * (1.0 - rs.BedPorosity)};
}

void OutputSolutionDataTxt(const RunSet &rs, const std::string &path, const std::map<int, Tile> &tiles) {
   // column layout Output.f90:799-834; 1-D: Hn = 3, u = 5, Hnpsi = 9, bt = 12; 2-D: Hn = 6, u = 8, Hnpsi = 14, bt = 17
   FILE *fh = std::fopen(path.c_str(), "w");
   if (!fh) throw FatalError("Could not open " + path + " for writing");
   std::vector<double> x, y, xv, yv;
   const int nX = rs.nXpertile, nY = rs.nYpertile;
   for (const auto &kv : tiles) {
      const Tile &T = kv.second;
      TileCoords(rs, kv.first, x, y, xv, yv);
      for (int j = 0; j < nY; j++) {
         for (int i = 0; i < nX; i++) {
            const double *q = &T.u[((size_t)j * nX + i) * 13];
            const double spd = std::sqrt(q[iU] * q[iU] + q[iV] * q[iV]);
            std::fprintf(fh, "%8d", kv.first);
            if (rs.isOneD) {
               const double cols[13] = {x[i], q[iHN], q[iW], q[iU], spd, q[iRHO], q[iB0] + q[iBT], q[iHPSI], q[iPSI], q[iHU], q[iBT],
                                        q[iBX], q[iB0]};
               for (double v : cols) std::fprintf(fh, ", %18.10E", v);
            } else {
               const double cols[19] = {x[i], y[j], 0.0, 0.0, q[iHN], q[iW], q[iU], q[iV], spd, q[iRHO], q[iB0], q[iB0] + q[iBT],
                                        q[iHPSI], q[iPSI], q[iHU], q[iBT], q[iHV], q[iBX], q[iBY]};
               for (double v : cols) std::fprintf(fh, ", %18.10E", v);
            }
            std::fputc('\n', fh);
         }
         std::fputc('\n', fh);
      }
   }
   std::fclose(fh);
}

void OutputVolumeTxt(const std::string &path, const std::vector<VolumeRow> &rows) {
   FILE *fh = std::fopen(path.c_str(), "w");
   if (!fh) throw FatalError("Could not open " + path + " for writing");
   std::fputs("        time,                   volume,         total_bed_volume,               total_mass,"
              "                 bed_mass,        total_solids_mass,          bed_solids_mass\n", fh);
   for (const auto &r : rows)
      std::fprintf(fh, "%12.2f, %24.15E, %24.15E, %24.15E, %24.15E, %24.15E, %24.15E\n", r.t, r.vol, r.bed, r.mass, r.bedMass,
                   r.solidsMass, r.bedSolidsMass);
   std::fclose(fh);
}

// ------------------------------------------------------------------ Simulation
static int heightsCallback(void *ctx, int32_t tileId, double *b0) {  // never throws across the ABI
   try {
      GetHeights(*static_cast<const RunSet *>(ctx), tileId, b0);
      return 0;
   } catch (...) {
      return 1;
   }
}

Simulation::Simulation(RunSet rs) : rs_(std::move(rs)) {
   ic_ = LoadSourceConditions(rs_);  // also fills NumCellsInSrc
   kgpu_params p = rs_.ToParams(srcs_, heightsCallback, &rs_);
   int rc = kgpu_create(&p, &h_);
   if (rc != KGPU_OK) {
      std::string msg = h_ ? kgpu_last_error(h_) : "kgpu_create failed";
      throw FatalError("kgpu_create: " + msg + " (status " + std::to_string(rc) + "; there is no CPU path)");
   }
   // KESTREL_GPU_DEVICE_TOPOGRAPHY=1: tiles activated during the run get the heights of `Topog function` from a kernel
   // (kgpu_set_topography_function) instead of from the callback above; functions the library does not know keep the callback
   if (const char *e = std::getenv("KESTREL_GPU_DEVICE_TOPOGRAPHY")) {
      static const char *names[] = {"flat", "xslope", "yslope", "xyslope", "xsinslope", "xysinslope", "xhump", "xtanh", "xparab", "xyparab",
                                    "xbislope", "x2slopes", "usgs", "flume", "channel power law", "channel trapezium", "xtrislope"};
      if (e[0] == '1' && rs_.topog_type == "function")
         for (int f = 0; f < 17; f++)
            if (rs_.topog_func == names[f]) {
               Check(kgpu_set_topography_function(h_, f, rs_.topog_params.data(), (int32_t)rs_.topog_params.size()), "kgpu_set_topography_function");
               break;
            }
   }
   for (auto &kv : ic_) {
      Tile &T = kv.second;
      Check(kgpu_upload_tile(h_, kv.first, T.u.data(), T.b0.data(), nullptr, T.maxima.data(), T.tfirst.data(), T.containsSource ? 1 : 0),
            "kgpu_upload_tile");
   }
}

Simulation::~Simulation() {
   if (h_) kgpu_destroy(h_);
}

void Simulation::Check(int rc, const char *what) {
   if (rc == KGPU_OK) return;
   if (rc == KGPU_ERR_HALT_BC) throw FatalError("Error: tried to add a tile outside the domain.");  // UpdateTiles.f90:63-65
   const char *msg = h_ ? kgpu_last_error(h_) : nullptr;
   throw FatalError(std::string(what) + ": " + (msg ? msg : "error") + " (status " + std::to_string(rc) + ")");
}

std::map<int, Tile> Simulation::DownloadActive() {
   int32_t n = 0;
   Check(kgpu_active_tiles(h_, &n, nullptr), "kgpu_active_tiles");
   std::vector<int32_t> ids(n);
   if (n) Check(kgpu_active_tiles(h_, &n, ids.data()), "kgpu_active_tiles");
   const int nX = rs_.nXpertile, nY = rs_.nYpertile;
   std::map<int, Tile> out;
   for (int32_t id : ids) {
      Tile T;
      T.id = id;
      T.u.assign((size_t)13 * nX * nY, 0.0);
      T.b0.assign((size_t)(nX + 1) * (nY + 1), 0.0);
      T.bt.assign((size_t)(nX + 1) * (nY + 1), 0.0);
      T.maxima.assign((size_t)10 * nX * nY, 0.0);
      T.tfirst.assign((size_t)nX * nY, 0.0);
      Check(kgpu_download_tile(h_, id, T.u.data(), T.b0.data(), T.bt.data(), T.maxima.data(), T.tfirst.data()), "kgpu_download_tile");
      out.emplace(id, std::move(T));
   }
   return out;
}

void Simulation::Run(const std::string &outDir) {
   auto name = [&](int i) {
      char buf[32];
      std::snprintf(buf, sizeof buf, "%06d.txt", i);
      return outDir + "/" + buf;
   };
   volumeRows.push_back(CalculateVolume(rs_, rs_.tstart, ic_));
   if (!outDir.empty()) OutputSolutionDataTxt(rs_, name(0), ic_);
   for (int i = 1; i <= rs_.Nout; i++) {
      const double tk = rs_.tstart + i * rs_.DeltaT;
      kgpu_step_info info;
      Check(kgpu_integrate_to(h_, tk, 0, &info), "kgpu_integrate_to");
      infos.push_back(info);
      std::map<int, Tile> tiles = DownloadActive();
      volumeRows.push_back(CalculateVolume(rs_, tk, tiles));
      if (!outDir.empty()) OutputSolutionDataTxt(rs_, name(i), tiles);
   }
   if (!outDir.empty()) OutputVolumeTxt(outDir + "/Volume.txt", volumeRows);
}

}  // namespace kestrel
