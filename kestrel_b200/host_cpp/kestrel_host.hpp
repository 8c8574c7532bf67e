// kestrel_host.hpp -- C++ host side above the C-ABI of include/kestrel_gpu.h.
//
// The reference's host is Fortran (main.f90, Input.f90, *Settings.f90, SetSources.f90,
// TimeStepper.f90:73-113, Output.f90); the image has no Fortran compiler, so the part of it
// that surrounds IntegrateTo is mirrored here in C++ with the reference's names, argument
// meaning and error behaviour:
//
//   ReadInputFile          Input.f90:54-565 + DomainSettings / Parameters / SolverSettings /
//                          OutputSettings / TopogSettings / InitConds readers
//   RunSet::Finalize       DomainSettings.f90:173-225, Parameters.f90:629-648,
//                          SolverSettings.f90:189-203, OutputSettings.f90:167
//   TileCoords / GetHeights  Grid.f90:339-353, UpdateTiles.f90:288-325, TopogFuncs.f90,
//                          dem.f90:360-415 (Type = Function only; rasters need GDAL)
//   LoadSourceConditions   SetSources.f90:47-392
//   Run                    TimeStepper.f90:73-113: output, then per output interval
//                          IntegrateTo (= kgpu_integrate_to) + OutputSolutionData + CalculateVolume
//   CalculateVolume        Output.f90:617-735
//   OutputSolutionData_txt Output.f90:799-834 (column layout indexed by the reference's tests)
//
// kestrel_b200/host/ holds the same logic in Python over ctypes (it drives the parity tests);
// tests/test_host_cpp.py checks that both hosts produce the same files.
#pragma once

#include <cstdint>
#include <limits>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "kestrel_gpu.h"

namespace kestrel {

// FatalErrorMessage (Messages.f90): the reference prints and stops; the driver catches this,
// prints the message to stderr and exits with status 1.
struct FatalError : std::runtime_error {
   using std::runtime_error::runtime_error;
};

struct FluxSource {  // type Sources, RunSettings.f90:101-109
   double x = 0, y = 0, radius = 0;
   std::vector<double> time, flux, psi;
   int numCellsInSrc = 0;
};
struct Cap {  // InitConds.f90:254-460
   double x = 0, y = 0, radius = 0, height = 0, volume = 0, psi = 0, u = 0, v = 0;
   std::string shape = "flat";
};
struct Cube {  // InitConds.f90:526-773
   double x = 0, y = 0, length = 0, width = 0, height = 0, psi = 0, u = 0, v = 0;
   std::string shape = "flat";
};

struct RunSet {  // RunSettings.f90:168-285, the fields the time step reads
   // Domain
   int nXtiles = 1, nYtiles = 1, nXpertile = 1, nYpertile = 1;
   double Xtilesize = 1.0, Ytilesize = 0.0;
   bool hasXtilesize = true, hasYtilesize = false;
   std::string bcs = "halt";
   double bcsHnval = 0, bcsuval = 0, bcsvval = 0, bcspsival = 0;
   // Parameters (defaults Parameters.f90:41-77)
   bool geometric_factors = true;
   double g = 9.81, rhow = 1000.0, rhos = 2000.0;
   double ChezyCo = 0.01, ManningCo = 0.03, CoulombCo = 0.1;
   double PouliquenMinSlope = 0.1, PouliquenMaxSlope = 0.4, PouliquenIntermediateSlope = 0.2, PouliquenBeta = 0.136;
   double Edwards2019betastar = 0.136, Edwards2019kappa = 1.0, Edwards2019Gamma = 0.0;
   double VoellmySwitchRate = 3.0, VoellmySwitchValue = 0.2;
   double EroRate = 0.001, EroRateGranular = 4.0, EroDepth = 1.0, EroCriticalHeight = 0.01;
   double BedPorosity = 0.35, maxPack = 0.65, SolidDiameter = 1e-3, EddyViscosity = 0.0;
   double ws0 = 0.0;
   bool hasWs0 = false;
   std::string drag = "chezy", erosion = "mixed", deposition = "spearman manning", erosion_transition = "smooth",
               morpho_damp = "tanh", fswitch = "tanh";
   // Solver
   std::string limiter = "minmod2";
   double heightThreshold = 1e-6;
   int TileBuffer = 1;
   double cfl = 0.0;
   bool hasCfl = false;
   double maxdt = std::numeric_limits<double>::max();
   double tstart = 0.0, tend = 1.0, SpongeStrength = 0.2;
   // Output
   int Nout = 1;
   std::string out_dir = "results/";
   // Topog
   std::string topog_type = "function", topog_func = "flat";
   std::vector<double> topog_params;
   // Initial conditions
   std::vector<Cap> caps;
   std::vector<Cube> cubes;
   std::vector<FluxSource> sources;
   // library options
   int arithmetic = 0, device = -1;

   // derived
   int nTiles = 1, NX = 1, NY = 1;
   bool isOneD = false, MorphodynamicsOn = true, SpongeLayer = false;
   double xSize = 1, ySize = 1, deltaX = 1, deltaY = 1, deltaXRecip = 1, deltaYRecip = 1;
   double gred = 0, Rep = 0, nsettling = 0, CriticalShields = 0, diffusiveTimeScale = 0, DeltaT = 1;

   void Finalize();
   // the POD that crosses the ABI; `srcs` owns the source table the struct points into
   kgpu_params ToParams(std::vector<kgpu_source> &srcs, kgpu_heights_fn cb, void *ctx) const;
};

RunSet ReadInputFile(const std::string &path, std::vector<std::string> *warnings = nullptr);

// zero-based indices of u(d,:,:) (main.f90:76-101)
enum { iW = 0, iHU, iHV, iHPSI, iHN, iU, iV, iPSI, iRHO, iB0, iBT, iBX, iBY };

struct Tile {             // the fields of TileType (Grid.f90:57-121) that cross the ABI
   int id = 0;
   std::vector<double> u;        // (13, nX, nY): d fastest, then i, then j
   std::vector<double> b0, bt;   // (nX+1, nY+1), i fastest
   std::vector<double> maxima;   // 5 x (value plane, time plane) x (nX, nY): Hnmax, umax, emax, dmax, psimax
   std::vector<double> tfirst;   // (nX, nY)
   bool containsSource = false;
};

void TileCoords(const RunSet &rs, int tileId, std::vector<double> &x, std::vector<double> &y, std::vector<double> &xv,
                std::vector<double> &yv);
double TopogFunction(const RunSet &rs, double x, double y);
void GetHeights(const RunSet &rs, int tileId, double *b0Vertices);
std::map<int, Tile> LoadSourceConditions(RunSet &rs);

struct VolumeRow {
   double t, vol, bed, mass, bedMass, solidsMass, bedSolidsMass;
};
VolumeRow CalculateVolume(const RunSet &rs, double t, const std::map<int, Tile> &tiles);
void OutputSolutionDataTxt(const RunSet &rs, const std::string &path, const std::map<int, Tile> &tiles);
void OutputVolumeTxt(const std::string &path, const std::vector<VolumeRow> &rows);

// LoadSourceConditions + Run (main.f90:126-131, TimeStepper.f90:73-113) against libkestrel_gpu
class Simulation {
  public:
   explicit Simulation(RunSet rs);
   ~Simulation();
   Simulation(const Simulation &) = delete;
   Simulation &operator=(const Simulation &) = delete;
   // writes 000000.txt ... and Volume.txt into outDir when it is not empty
   void Run(const std::string &outDir);
   std::map<int, Tile> DownloadActive();
   const RunSet &settings() const { return rs_; }
   const std::map<int, Tile> &initialTiles() const { return ic_; }
   std::vector<VolumeRow> volumeRows;
   std::vector<kgpu_step_info> infos;

  private:
   void Check(int rc, const char *what);
   RunSet rs_;
   std::map<int, Tile> ic_;
   std::vector<kgpu_source> srcs_;
   kgpu_handle *h_ = nullptr;
};

}  // namespace kestrel
