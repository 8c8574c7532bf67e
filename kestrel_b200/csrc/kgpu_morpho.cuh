// kgpu_morpho.cuh -- morphodynamic operator M (MorphodynamicRHS.f90, TimeStepper.f90:532-781).
#pragma once
#include "kgpu_device.cuh"

namespace kgpu {
struct RedistEntry {
   double excess;
   int i, j;
};
}  // namespace kgpu
