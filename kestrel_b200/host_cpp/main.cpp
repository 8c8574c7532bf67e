// kestrel_gpu_run -- command-line driver: the reference's `kestrel <input file>` (main.f90) with
// the time step running in libkestrel_gpu.  Reads a Kestrel input file, builds the initial tiles
// (LoadSourceConditions), integrates to each output time through the C-ABI and writes the
// reference's text outputs (NNNNNN.txt, Volume.txt).
//
//   kestrel_gpu_run <input.txt> [-o DIR] [--arithmetic 0|1] [--tend T] [--nout N] [--init-only] [--quiet]
//
// --init-only stops after LoadSourceConditions and writes 000000.txt / Volume.txt without touching
// a GPU.  Fatal errors print the reference's message and exit with status 1; there is no CPU path.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <sys/stat.h>

#include "kestrel_host.hpp"

static void makeDirs(const std::string &path) {
   std::string cur;
   for (size_t i = 0; i <= path.size(); i++) {
      if (i == path.size() || path[i] == '/') {
         if (!cur.empty()) mkdir(cur.c_str(), 0777);
      }
      if (i < path.size()) cur.push_back(path[i]);
   }
}

int main(int argc, char **argv) {
   std::string input, outDir;
   int arithmetic = 0, nout = -1;
   double tend = 0.0;
   bool hasTend = false, initOnly = false, quiet = false;
   for (int a = 1; a < argc; a++) {
      std::string s = argv[a];
      auto next = [&]() -> const char * {
         if (a + 1 >= argc) { std::fprintf(stderr, "missing value after %s\n", s.c_str()); std::exit(2); }
         return argv[++a];
      };
      if (s == "-o" || s == "--out") outDir = next();
      else if (s == "--arithmetic") arithmetic = std::atoi(next());
      else if (s == "--tend") { tend = std::atof(next()); hasTend = true; }
      else if (s == "--nout") nout = std::atoi(next());
      else if (s == "--init-only") initOnly = true;
      else if (s == "--quiet") quiet = true;
      else if (s == "-h" || s == "--help") {
         std::printf("usage: %s <input.txt> [-o DIR] [--arithmetic 0|1] [--tend T] [--nout N] [--init-only] [--quiet]\n", argv[0]);
         return 0;
      } else if (input.empty()) input = s;
      else { std::fprintf(stderr, "unexpected argument %s\n", s.c_str()); return 2; }
   }
   if (input.empty()) { std::fprintf(stderr, "usage: %s <input.txt> [-o DIR] ...\n", argv[0]); return 2; }
   try {
      std::vector<std::string> warnings;
      kestrel::RunSet rs = kestrel::ReadInputFile(input, &warnings);
      if (!quiet) for (const auto &w : warnings) std::fprintf(stderr, "Warning: %s\n", w.c_str());
      if (hasTend) rs.tend = tend;
      if (nout > 0) rs.Nout = nout;
      rs.arithmetic = arithmetic;
      rs.Finalize();
      if (outDir.empty()) outDir = rs.out_dir;
      while (outDir.size() > 1 && outDir.back() == '/') outDir.pop_back();
      makeDirs(outDir);
      if (initOnly) {
         std::map<int, kestrel::Tile> ic = kestrel::LoadSourceConditions(rs);
         kestrel::OutputSolutionDataTxt(rs, outDir + "/000000.txt", ic);
         kestrel::OutputVolumeTxt(outDir + "/Volume.txt", {kestrel::CalculateVolume(rs, rs.tstart, ic)});
         if (!quiet) std::printf("initial conditions: %zu active tiles -> %s\n", ic.size(), outDir.c_str());
         return 0;
      }
      kestrel::Simulation sim(rs);
      sim.Run(outDir);
      if (!quiet) {
         const kgpu_step_info &last = sim.infos.back();
         long long steps = 0, refines = 0;
         for (const auto &i : sim.infos) { steps += i.nsteps; refines += i.nrefines; }
         std::printf("%s: t = %.6g, %lld steps (%lld rolled back), last dt = %.6g -> %s\n", kgpu_version(), last.t, steps, refines,
                     last.dt_last, outDir.c_str());
      }
   } catch (const kestrel::FatalError &e) {
      std::fprintf(stderr, "%s\n", e.what());
      return 1;
   }
   return 0;
}
