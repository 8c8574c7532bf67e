"""kestrel_b200 -- B200-native replacement for Kestrel's explicit finite-volume
time step (IntegrateTo, src/TimeStepper.f90:116), behind the C-ABI of
include/kestrel_gpu.h.  The package holds the CUDA library (csrc/ -> lib/), its
ctypes binding (capi) and a host-side mirror of the reference's setup code (host/).
There is no CPU execution path: kestrel_b200.capi.load_gpu() raises if the CUDA
library has not been built.
"""
__version__ = "0.1.0"
