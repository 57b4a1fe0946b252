// sg_misc.cu -- the tick-parallel replay kernel (sg_replay.cuh), the FutureCollisionDetector look-ahead
// and the box-pair unit-test kernel.
#define SG_FLAT_BOXES 1  // boxes without area follow their own narrow-phase rules (sg_common.cuh)
#include "sg_common.cuh"
#include "sg_internal.h"

// ---------------------------------------------------------------------------------
__global__ void sg_box_pairs_kernel(const double* pa, const double* ba, const double* pb,
                                    const double* bb, uint8_t* out, int64_t n) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  double qa[8], qb[8];
  box_points(pa[3 * i], pa[3 * i + 1], pa[3 * i + 2], ba[4 * i], ba[4 * i + 1], ba[4 * i + 2], ba[4 * i + 3], qa);
  box_points(pb[3 * i], pb[3 * i + 1], pb[3 * i + 2], bb[4 * i], bb[4 * i + 1], bb[4 * i + 2], bb[4 * i + 3], qb);
  bool same = true;
  for (int f = 0; f < 8; ++f) same = same && (qa[f] == qb[f]);
  const Quad A = quad_from_array(qa), B = quad_from_array(qb);
  out[i] = !same && quads_intersect(A, quad_orientation(A), B, quad_orientation(B));
}

#include "sg_replay.cuh"

// ---------------------------------------------------------------------------------
// FutureCollisionDetector._step (reference sensor/common.py:88-105) for a batch: one warp per
// scenario; its lanes share the (look-ahead sample, other entity) pairs.  Every entity is placed
// at trajectory.position_at_t(time) (clamped), present or not; the pair test is the exact
// closed-set predicate of the collision path.
__global__ void sg_future_kernel(SgScene sc, const double* __restrict__ t, const int32_t* __restrict__ slot,
                                 double horizon, int n_samples, uint8_t* __restrict__ out) {
  const int n = (int)((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (n >= sc.n_scenarios) return;
  const int M = sc.n_slots;
  const int64_t nm = sc.plane_stride;
  const int es = slot ? slot[n] : sc.ego_slot[n];
  const int64_t ie = (int64_t)n * M + es;
  const int64_t re0 = sc.traj_off[ie];
  const int Ke = (int)(sc.traj_off[ie + 1] - re0);
  const double start = t[n], stop = t[n] + horizon;
  const double step = n_samples > 1 ? (stop - start) / (double)(n_samples - 1) : 0.0;  // numpy.linspace
  bool hit = false;
  if (Ke > 0)
    for (int w = lane; w < n_samples * M; w += 32) {
      const int k = w / M, j = w - k * M;
      const int64_t i = (int64_t)n * M + j;
      if (j == es || sc.kind[i] == SG_KIND_EMPTY) continue;
      const int64_t r0 = sc.traj_off[i];
      const int K = (int)(sc.traj_off[i + 1] - r0);
      if (K == 0) continue;
      double tk = (double)k * step + start;
      if (n_samples > 1 && k == n_samples - 1) tk = stop;
      double pe[6], po[6], qe[8], qo[8];
      int c0 = 0, c1 = 0;
      position_at_t(sc.traj_rows + re0 * 7, Ke, tk, EXT_CLAMP, c0, pe);
      position_at_t(sc.traj_rows + r0 * 7, K, tk, EXT_CLAMP, c1, po);
      box_points(pe[0], pe[1], pe[3], sc.box[ie], sc.box[nm + ie], sc.box[2 * nm + ie], sc.box[3 * nm + ie], qe);
      box_points(po[0], po[1], po[3], sc.box[i], sc.box[nm + i], sc.box[2 * nm + i], sc.box[3 * nm + i], qo);
      bool same = true;
#pragma unroll
      for (int f = 0; f < 8; ++f) same = same && (qe[f] == qo[f]);
      if (same) continue;  // `g != g_prime`, reference utils.py:58
      const Quad A = quad_from_array(qe), B = quad_from_array(qo);
      if (quads_intersect(A, quad_orientation(A), B, quad_orientation(B))) hit = true;
    }
  hit = __any_sync(0xffffffffu, hit);
  if (lane == 0) out[n] = hit ? 1 : 0;
}

// Rows of the action table an SgActionRng describes (sg_fill_random_actions): one thread per slot
// index walks the ticks with one 128-bit multiply-add per draw; stores are coalesced over the slots.
__global__ void sg_fill_actions_kernel(SgRngDev rng, int n_ticks, int64_t nm, double* __restrict__ out) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= nm) return;
  const sg_u128 A = sg_u128_make(rng.a_hi, rng.a_lo), C = sg_u128_make(rng.c_hi, rng.c_lo);
#pragma unroll 1
  for (int c = 0; c < 2; ++c) {
    sg_u128 q = sg_rng_slot_state(rng, c, i);
    for (int k = 0; k < n_ticks; ++k) {
      const double u = sg_pcg_double((uint64_t)(q >> 64), (uint64_t)q);
      out[((int64_t)k * 2 + c) * nm + i] = rng.low[c] + rng.scale[c] * u;
      q = A * q + C;
    }
  }
}

cudaError_t sgi_launch_fill_actions(cudaStream_t s, const SgRngDev& rng, int n_ticks, int64_t nm, double* out) {
  if (nm <= 0 || n_ticks <= 0) return cudaSuccess;
  sg_fill_actions_kernel<<<(unsigned)((nm + 255) / 256), 256, 0, s>>>(rng, n_ticks, nm, out);
  return cudaGetLastError();
}

// State.get_entities_in_radius for a batch (sg_entities_in_radius): one thread per slot
__global__ void sg_radius_kernel(SgState st, int n_scen, int M, const double* __restrict__ x, const double* __restrict__ y,
                                 const double* __restrict__ r, uint8_t* __restrict__ out) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, nm = (int64_t)n_scen * M;
  if (i >= nm) return;
  const int n = (int)(i / M);
  uint8_t hit = 0;
  if (r[n] > 0.0 && st.present[i]) hit = in_buffer(x[n], y[n], r[n], st.pose[i], st.pose[nm + i]) ? 1 : 0;
  out[i] = hit;
}

cudaError_t sgi_launch_radius(cudaStream_t s, const SgState& st, int n_scen, int M, const double* x, const double* y,
                              const double* r, uint8_t* out) {
  const int64_t nm = (int64_t)n_scen * M;
  sg_radius_kernel<<<(unsigned)((nm + 255) / 256), 256, 0, s>>>(st, n_scen, M, x, y, r, out);
  return cudaGetLastError();
}

// BatchReplayEntity union-knot table on the device (sg_build_union_x; reference entity/batch.py:80-112):
// one thread per (union row, slot).  Every replayed slot's trajectory is resampled, clamped, at the
// scenario's union knot times with scipy's `_call_linear` arithmetic -- the device's position_at_t in
// clamped mode; a single control point is duplicated 0.1 s later (batch.py:94-96).  The host uploads the
// 8-byte knot times only, not the 48 M-byte rows.
__global__ void sg_union_kernel(SgScene sc) {
  const int M = sc.n_slots;
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  // (a scenario window's rows are union_off[0] .. union_off[n_scenarios] of the batch's table)
  const int64_t row0 = __ldg(sc.union_off), row1 = __ldg(sc.union_off + sc.n_scenarios);
  if (idx >= (row1 - row0) * M) return;
  const int64_t r = row0 + idx / M;
  const int s = (int)(idx - (r - row0) * M);
  int lo = 0, hi = sc.n_scenarios;  // scenario of row r: union_off[n] <= r < union_off[n + 1]
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(sc.union_off + mid) <= r) lo = mid; else hi = mid;
  }
  const int64_t i = (int64_t)lo * M + s;
  double out[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  if (sc.kind[i] == SG_KIND_REPLAY) {
    const int64_t r0 = sc.traj_off[i];
    const int K = (int)(sc.traj_off[i + 1] - r0);
    const double* rows = sc.traj_rows + r0 * 7;
    const double t = sc.union_t[r];
    if (K == 1) {
      const double x_lo = __ldg(rows), x_hi = x_lo + 1e-1;
      const bool inside = !(t < x_lo) && !(t > x_hi);
      const double w1 = (t - x_lo) / (x_hi - x_lo), w0 = (x_hi - t) / (x_hi - x_lo);
#pragma unroll
      for (int f = 0; f < 6; ++f) {
        const double y = __ldg(rows + 1 + f);
        out[f] = inside ? w1 * y + w0 * y : y;
      }
    } else if (K > 1) {
      int cur = 0;
      position_at_t(rows, K, t, EXT_CLAMP, cur, out);
    }
  }
  double* X = (double*)sc.union_x + r * 6 * M + s;
#pragma unroll
  for (int f = 0; f < 6; ++f) X[(int64_t)f * M] = out[f];
}

cudaError_t sgi_launch_union(cudaStream_t s, const SgScene& sc, int64_t rows) {
  const int64_t n = rows * sc.n_slots;
  if (n <= 0) return cudaSuccess;
  sg_union_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(sc);
  return cudaGetLastError();
}

// Trajectory.position_at_t / velocity_at_t truth tables (sg_test_trajectory): one thread per query time
__global__ void sg_traj_kernel(const double* __restrict__ rows, int K, const double* __restrict__ t, int64_t n,
                               int mode, double* __restrict__ pos, uint8_t* __restrict__ ok, double* __restrict__ vel) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  double out[6] = {0, 0, 0, 0, 0, 0};
  int cur = 0;
  const bool present = position_at_t(rows, K, t[i], mode, cur, out);
  ok[i] = present ? 1 : 0;
  for (int f = 0; f < 6; ++f) pos[6 * i + f] = out[f];
  if (vel) {
    velocity_at_t(rows, K, t[i], out);
    for (int f = 0; f < 6; ++f) vel[6 * i + f] = out[f];
  }
}

cudaError_t sgi_launch_traj(cudaStream_t s, const double* rows, int K, const double* t, int64_t n, int mode,
                            double* pos, uint8_t* ok, double* vel) {
  if (n <= 0) return cudaSuccess;
  sg_traj_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(rows, K, t, n, mode, pos, ok, vel);
  return cudaGetLastError();
}

// FP64 pipe micro-benchmark (sg_measure_fp64_peak): 8 independent DFMA chains per thread
__global__ void sg_dfma_kernel(double* __restrict__ out, int iters, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
#pragma unroll 1
  for (int k = 0; k < iters; ++k) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      x0 = __fma_rn(x0, a, b); x1 = __fma_rn(x1, a, b); x2 = __fma_rn(x2, a, b); x3 = __fma_rn(x3, a, b);
      x4 = __fma_rn(x4, a, b); x5 = __fma_rn(x5, a, b); x6 = __fma_rn(x6, a, b); x7 = __fma_rn(x7, a, b);
    }
  }
  const double sum = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
  if (sum == 12345.678) out[0] = sum;  // keeps the chains alive
}

cudaError_t sgi_measure_fp64(cudaStream_t s, double* inst_per_s) {
  int dev = 0, sms = 0;
  cudaError_t err = cudaGetDevice(&dev);
  if (err != cudaSuccess) return err;
  err = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (err != cudaSuccess) return err;
  double* scratch = nullptr;
  err = cudaMalloc(&scratch, sizeof(double));
  if (err != cudaSuccess) return err;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int blocks = sms * 8, threads = 256, iters = 4096;
  double best = 0.0;
  for (int rep = 0; rep < 4 && err == cudaSuccess; ++rep) {
    cudaEventRecord(e0, s);
    sg_dfma_kernel<<<blocks, threads, 0, s>>>(scratch, iters, 0.999999, 1e-9);
    cudaEventRecord(e1, s);
    err = cudaEventSynchronize(e1);
    float ms = 0.f;
    if (err == cudaSuccess) err = cudaEventElapsedTime(&ms, e0, e1);
    const double inst = (double)blocks * threads * iters * 32.0;
    if (err == cudaSuccess && rep > 0 && ms > 0.f) best = fmax(best, inst / (ms * 1e-3));
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(scratch);
  *inst_per_s = best;
  return err;
}

cudaError_t sgi_launch_replay(cudaStream_t s, const SgScene& sc, const SgParams& p, const SgState& st,
                              int n_ticks) {
  const size_t rsm = replay_smem_bytes(sc.n_slots);
  auto rk = (p.features & SG_FEAT_COLL_MATRIX) ? sg_replay_kernel<true> : sg_replay_kernel<false>;
  if (rsm > 48 * 1024) {
    cudaError_t err = cudaFuncSetAttribute(rk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rsm);
    if (err != cudaSuccess) return err;
  }
  rk<<<sc.n_scenarios, SG_RP_BLOCK, rsm, s>>>(sc, p, st, n_ticks);
  return cudaGetLastError();
}

cudaError_t sgi_launch_box_pairs(cudaStream_t s, const double* pa, const double* ba, const double* pb,
                                 const double* bb, uint8_t* out, int64_t n) {
  sg_box_pairs_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(pa, ba, pb, bb, out, n);
  return cudaGetLastError();
}

cudaError_t sgi_launch_future(cudaStream_t s, const SgScene& sc, const double* t, const int32_t* slot,
                              double horizon, int n_samples, uint8_t* out) {
  const int64_t threads = (int64_t)sc.n_scenarios * 32;
  sg_future_kernel<<<(unsigned)((threads + 127) / 128), 128, 0, s>>>(sc, t, slot, horizon, n_samples, out);
  return cudaGetLastError();
}
