// sg_vehicle_rss1.cu -- sg_vehicle_kernel<RSS = 1> instantiations (see sg_vehicle.cuh).
#include "sg_vehicle.cuh"

cudaError_t sgi_launch_vehicle_rss1(int n_scen, cudaStream_t s, const SgScene& sc, const SgParams& p,
                                    const SgState& st, const SgInputs& in, const SgRngDev& rng, int act,
                                    int n_ticks, const GroupLayout& L) {
  return launch_vehicle_t<true>(n_scen, s, sc, p, st, in, rng, act, n_ticks, L);
}
