// sg_pcg.cuh -- numpy's PCG64 stream on the device (SgActionRng, include/sg_b200.h).
//
// numpy.random.default_rng(seed) is Generator(PCG64(seed)): a 128-bit LCG
//     state <- state * MULT + inc      (numpy/random/src/pcg64/pcg64.h, pcg_setseq_128_step_r)
// whose output is XSL-RR of the state AFTER the step (pcg_output_xsl_rr_128_64), and
//     random() = (next_uint64 >> 11) * 2^-53 ,  uniform(low, high) = low + (high - low) * random()
// (numpy/random/src/distributions/distributions.c).  Draw j therefore depends on the start state
// only through "the state advanced j + 1 times", and advancing by any fixed number of steps is
// itself an affine map  state <- A * state + C  (mod 2^128): a thread jumps to its own slot in
// O(log j) once per launch (pcg_advance) and moves from one tick's draw to the next with a single
// 128-bit multiply-add by the launch-uniform (A, C) of the tick stride.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

typedef unsigned __int128 sg_u128;

#define SG_PCG_MULT_HI 0x2360ED051FC65DA4ULL
#define SG_PCG_MULT_LO 0x4385DF649FCCF645ULL

__host__ __device__ __forceinline__ sg_u128 sg_u128_make(uint64_t hi, uint64_t lo) {
  return (((sg_u128)hi) << 64) | (sg_u128)lo;
}

// coefficients (A, C) of `delta` LCG steps: state_{k + delta} = A * state_k + C
__host__ __device__ inline void sg_pcg_jump_coeffs(sg_u128 inc, sg_u128 delta, sg_u128& A, sg_u128& C) {
  sg_u128 acc_mult = 1, acc_plus = 0, cur_mult = sg_u128_make(SG_PCG_MULT_HI, SG_PCG_MULT_LO), cur_plus = inc;
  while (delta > 0) {
    if (delta & 1) {
      acc_mult *= cur_mult;
      acc_plus = acc_plus * cur_mult + cur_plus;
    }
    cur_plus = (cur_mult + 1) * cur_plus;
    cur_mult *= cur_mult;
    delta >>= 1;
  }
  A = acc_mult;
  C = acc_plus;
}

__host__ __device__ inline sg_u128 sg_pcg_advance(sg_u128 state, sg_u128 inc, sg_u128 delta) {
  sg_u128 A, C;
  sg_pcg_jump_coeffs(inc, delta, A, C);
  return A * state + C;
}

// XSL-RR 128/64 output of a (post-step) state, as the double numpy's random() returns
__host__ __device__ __forceinline__ double sg_pcg_double(uint64_t hi, uint64_t lo) {
  const uint64_t x = hi ^ lo;
  const unsigned rot = (unsigned)(hi >> 58);
  const uint64_t r = (x >> rot) | (x << ((64u - rot) & 63u));
  return (double)(r >> 11) * (1.0 / 9007199254740992.0);
}

// Launch-uniform description of the two action streams, precomputed on the host (sg_api.cu):
// s[c] = the generator state advanced to row rng_tick0 of component c, slot 0 (i.e. by
// offset[c] + rng_tick0 * tick_stride steps); (a, c) = the affine map of one tick stride.
struct SgRngDev {
  uint64_t s_hi[2], s_lo[2];
  uint64_t inc_hi, inc_lo;
  uint64_t a_hi, a_lo, c_hi, c_lo;
  double low[2], scale[2];
};

// post-step state of the draw of slot index i in the first row: S advanced by i + 1
__device__ inline sg_u128 sg_rng_slot_state(const SgRngDev& r, int c, int64_t i) {
  return sg_pcg_advance(sg_u128_make(r.s_hi[c], r.s_lo[c]), sg_u128_make(r.inc_hi, r.inc_lo), (sg_u128)(i + 1));
}
