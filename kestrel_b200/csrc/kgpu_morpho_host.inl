// Strang remainder M(2dt) H(dt) of one step (TimeStepper.f90:194-255); included by kestrel_gpu.cu.
static int strangRemainder(kgpu_handle *h, double t0, double &dt_hydro, bool &again) {
   (void)t0; (void)dt_hydro; again = false;
   h->err = "morphodynamics not available in this build";
   return KGPU_ERR_UNSUPPORTED;
}
