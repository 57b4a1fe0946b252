// sg_vehicle.cuh -- the fused tick loop of vehicle-only scenes (C3 / C5).  Instantiated by
// sg_vehicle_rss0.cu / sg_vehicle_rss1.cu (one translation unit per RSS flavour: parallel builds).
#pragma once
#include "sg_common.cuh"
#include "sg_internal.h"
#include "sg_pcg.cuh"

SG_DEV void cp_async4(unsigned smem_addr, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_addr), "l"(gsrc) : "memory");
}

// ---------------------------------------------------------------------------------
// Vehicle-only scenes (every live slot has a VehicleController; C3 / C5): a lean tick that
// keeps x, y, h, their velocities, cos/sin of the heading, distance and speed in registers.
// z, p, r never change under VehicleController._step (controller.py:122-131), so their
// velocities are 0 after the first tick and they stay in global memory.
// ---------------------------------------------------------------------------------
// MC > 0: the kernel is specialised for scenes of exactly MC slots -- the shared-memory layout, the
// group size and every index stride are compile-time constants (immediates instead of address
// arithmetic on the layout struct in the constant bank: 5 % of the tick's instructions).
template <bool RSS, int MC>
struct VehLayout {
  static constexpr GroupLayout value = make_layout(MC > 0 ? MC : 64, false, RSS, true, false);
};
template <bool RSS, int MAXT, int MINB, bool SORTED, bool LEAN = false, int ACT = ACT_F64, int MC = 0>
__global__ void __launch_bounds__(MAXT, MINB)
sg_vehicle_kernel(SgScene sc, SgParams p, SgState st, SgInputs in, int n_ticks, GroupLayout Lrt, SgRngDev rng) {
  extern __shared__ __align__(16) unsigned char smem[];
  const GroupLayout L = MC > 0 ? VehLayout<RSS, MC>::value : Lrt;
  const int G = L.G, M = MC > 0 ? MC : sc.n_slots, W = L.W;
  const int gpb = blockDim.x / G;
  const int gl = threadIdx.x / G;
  const int s = threadIdx.x - gl * G;
  const int n = blockIdx.x * gpb + gl;
  if (gl >= gpb || n >= sc.n_scenarios) return;
  Grp c;
  setup_group(c, sc, L, smem, gl, s, n, M);
  const bool live = s < M && sc.kind[c.i] == SG_KIND_VEHICLE;
  const int ego_slot = sc.ego_slot[n], first_slot = sc.first_slot[n];
  // (LEAN is only launched with collisions on, no trace and no pair matrix; RSS only with the feature on)
  const bool need_coll = LEAN || (p.features & SG_FEAT_COLLISIONS) ||
                         (p.terminal & (SG_TERM_COLLISION | SG_TERM_EGO_COLLISION));
  const bool feat_rss = RSS;
  const bool matrix = !LEAN && (p.features & SG_FEAT_COLL_MATRIX) != 0;
  // lagged tick tail (sg_common.cuh): whole-warp groups of two or more warps, no terminal condition on collisions
  // (scenarios of several warps with the sorted sweep: C5 + 7 %; two-warp scenarios lose 8 % with it, the ego's
  // warp stays their critical path)
  const bool lag = LEAN && SORTED && !(p.terminal & (SG_TERM_COLLISION | SG_TERM_EGO_COLLISION));

  // hot per-entity state in registers; everything that is only read back at the end (safe
  // distances, ratios, heading rate) or is uniform per scenario (tick times, length, origin)
  // lives in shared memory to keep the register footprint of the tick loop small
  double x = 0, y = 0, h = 0, vx = 0, vy = 0, dist = 0, speed = 0, cs = 1, sn = 0;
  int orient_hint = 0;
  bool present = false;
  uint8_t collided = 0, rss_state = 0, rss_last = SG_RSS_NONE;
  bool rss_evald = false;  // RSSDistances ran for this hazard in the last executed tick
  double* tc = c.tcold + s;  // [0..3] safe dist / ratios, [4] heading rate, [5] 1 / wheelbase
  {
    const int64_t i = c.i, nm = c.nm;
    if (live) {
      x = st.pose[i]; y = st.pose[nm + i]; h = st.pose[3 * nm + i];
      vx = st.vel[i]; vy = st.vel[nm + i];
      dist = st.dist[i]; speed = st.speed[i];
      present = st.present[i] != 0;
      collided = st.collided[i];
      sincos_fast(h, sn, cs);
      const double bw = sc.box[i], bl = sc.box[nm + i];
      c.boxp[s] = bw; c.boxp[G + s] = bl;
      c.boxp[2 * G + s] = sc.box[2 * nm + i]; c.boxp[3 * G + s] = sc.box[3 * nm + i];
      tc[4 * G] = st.vel[3 * nm + i];
      tc[5 * G] = 1.0 / bl;
      orient_hint = box_orientation_hint(bw, bl);
      if (RSS) {
        rss_state = st.rss_state[i]; rss_last = st.rss_last[i];
        tc[0] = st.safe_dist[i]; tc[G] = st.safe_dist[nm + i];
        tc[2 * G] = st.safe_ratio[i]; tc[3 * G] = st.safe_ratio[nm + i];
      }
    }
    if (s == 0) {
      double* U = c.cold_d;
      U[COLD_T0] = st.t[n]; U[COLD_PT0] = st.prev_t[n];
      U[COLD_LEN] = sc.length[n];
      const int64_t er = sc.traj_off[(int64_t)n * M + ego_slot];  // origin of the fp32 bounds
      U[COLD_OX] = __ldg(sc.traj_rows + er * 7 + 1);
      U[COLD_OY] = __ldg(sc.traj_rows + er * 7 + 2);
      U[COLD_R2MINA] = 1.0 / (2 * p.rss_min_long_accel);
    }
  }
  int tick = st.tick[n];
  bool done = st.done[n] != 0;
  load_cold(st, c, n, s, W, ego_slot);
  if (RSS && s == ego_slot) publish_ego_box(c);
  if (SORTED) sorted_setup(c);
  int sort_round = 0;
  group_sync(c);

  int limit = n_ticks < 0 ? p.max_ticks : n_ticks;
  if (limit > in.n_action_ticks) limit = in.n_action_ticks;
  int parity = 0;
  const int tick0 = tick;
  // VehicleAction rows are staged one tick ahead with cp.async (no registers held); with the
  // device-side action source the same shared words hold the slot's two PCG64 states instead
  double* ab = c.actbuf + s;
  const unsigned ab_sh = (unsigned)__cvta_generic_to_shared(ab);  // converted once, not per tick
  const double* act = in.actions + c.i;
  const float* act32 = in.actions_f32 + c.i;
  unsigned long long* rq = (unsigned long long*)c.actbuf + s;  // [4][G]: accel lo, hi, steer lo, hi
  if (ACT == ACT_RNG) {
    if (live && limit > 0) {
#pragma unroll 1
      for (int cc = 0; cc < 2; ++cc) {
        const sg_u128 q = sg_rng_slot_state(rng, cc, c.i + (int64_t)sc.scenario_base * M);
        rq[(2 * cc) * G] = (unsigned long long)q;
        rq[(2 * cc + 1) * G] = (unsigned long long)(q >> 64);
      }
    }
  } else if (ACT == ACT_F32) {
    if (live && limit > 0 && (!done || in.step_done)) { cp_async4(ab_sh, act32); cp_async4(ab_sh + G * 8, act32 + c.nm); }
    cp_async_commit();
  } else {
    if (live && limit > 0 && (!done || in.step_done)) { cp_async8(ab_sh, act); cp_async8(ab_sh + G * 8, act + c.nm); }
    cp_async_commit();
  }

  double* const corners0 = c.corners;
  const unsigned corners_sh0 = c.corners_sh;
  for (int k = 0; k < limit && (!done || in.step_done); ++k) {
    if (LEAN && lag) {  // this tick's corners: the previous tick's narrow phase may still be reading the other set
      c.corners = corners0 + parity * 8 * G;
      c.corners_sh = corners_sh0 + (unsigned)(parity * 8 * G * 8);
    }
    const double* U = c.cold_d;
    const double t = U[COLD_T0 + parity];
    const double next_t = t + p.timestep;  // scenario_gym.py:229
    const double dt = next_t - t;          // controller.py:123 and State.dt after the step
    if (ACT != ACT_RNG) cp_async_wait_all();
    if (live && present) {  // VehicleController._step, controller.py:105-140
      double a_in, s_in;
      if (ACT == ACT_RNG) {  // numpy: low + scale * random(); then one tick stride further
        const sg_u128 A = sg_u128_make(rng.a_hi, rng.a_lo), Cc = sg_u128_make(rng.c_hi, rng.c_lo);
        unsigned long long lo = rq[0], hi = rq[G];
        a_in = rng.low[0] + rng.scale[0] * sg_pcg_double(hi, lo);
        sg_u128 q = A * sg_u128_make(hi, lo) + Cc;
        rq[0] = (unsigned long long)q; rq[G] = (unsigned long long)(q >> 64);
        lo = rq[2 * G]; hi = rq[3 * G];
        s_in = rng.low[1] + rng.scale[1] * sg_pcg_double(hi, lo);
        q = A * sg_u128_make(hi, lo) + Cc;
        rq[2 * G] = (unsigned long long)q; rq[3 * G] = (unsigned long long)(q >> 64);
      } else if (ACT == ACT_F32) {
        a_in = (double)*(const float*)(ab + parity * 2 * G);
        s_in = (double)*(const float*)(ab + parity * 2 * G + G);
      } else {
        a_in = ab[parity * 2 * G];
        s_in = ab[parity * 2 * G + G];
      }
      const double accel = clipd(a_in, -p.veh_max_accel, p.veh_max_accel);
      const double steer = clipd(s_in, -p.veh_max_steer, p.veh_max_steer);
      const double dx = speed * cs, dy = speed * sn;
      const double tn = fabs(steer) <= 0.78 ? tan_small(steer) : tan_lib(steer);
      const double dh = div_r(speed * tn, c.boxp[G + s], tc[5 * G]);
      const double nx = x + dx * dt, ny = y + dy * dt, nh = h + dh * dt;
      double ns = speed + accel * dt;
      if (!p.veh_allow_reverse) ns = ns < 0.0 ? 0.0 : ns;
      if (p.veh_max_speed == p.veh_max_speed) ns = ns > p.veh_max_speed ? p.veh_max_speed : ns;
      speed = ns;
      // State.update_statistics (state.py:230-239)
      const double rdt = fast_rcp(dt);
      const double ex = nx - x, ey = ny - y;
      vx = div_r(ex, dt, rdt); vy = div_r(ey, dt, rdt);
      tc[4 * G] = div_r(nh - h, dt, rdt);
      dist += fnorm2(ex, ey);
      x = nx; y = ny; h = nh;
      sincos_fast(h, sn, cs);
    }
    if (ACT == ACT_F64) {
      if (live && k + 1 < limit) {
        act += 2 * c.nm;
        cp_async8(ab_sh + (parity ^ 1) * 2 * G * 8, act);
        cp_async8(ab_sh + ((parity ^ 1) * 2 * G + G) * 8, act + c.nm);
      }
      cp_async_commit();
    } else if (ACT == ACT_F32) {
      if (live && k + 1 < limit) {
        act32 += 2 * c.nm;
        cp_async4(ab_sh + (parity ^ 1) * 2 * G * 8, act32);
        cp_async4(ab_sh + ((parity ^ 1) * 2 * G + G) * 8, act32 + c.nm);
      }
      cp_async_commit();
    }
    tick += 1;
    if (!LEAN && st.trace_cap > 0 && tick < st.trace_cap && s < M) {
      const int64_t i = c.i, nm = c.nm;
      st.trace_present[(int64_t)tick * nm + i] = present;
      double* tp = st.trace_pose + (int64_t)tick * 6 * nm + i;
      tp[0] = x; tp[nm] = y; tp[2 * nm] = st.pose[2 * nm + i]; tp[3 * nm] = h;
      tp[4 * nm] = st.pose[4 * nm + i]; tp[5 * nm] = st.pose[5 * nm + i];
      if (s == 0) st.trace_t[(int64_t)tick * sc.n_scenarios + n] = next_t;
    }
    if (s < M) {
      if (need_coll || feat_rss)
        publish_box<RSS, SORTED>(c, present, x, y, cs, sn, orient_hint, U[COLD_OX], U[COLD_OY]);
      if (!LEAN && matrix) {
        uint32_t* row = st.coll_mask + ((int64_t)n * M + s) * W;
        for (int w = 0; w < W; ++w) row[w] = 0;
      }
    }
    if (RSS && feat_rss && s == ego_slot) publish_ego(c, present, x, y, cs, sn, vx, vy);
    if (s == 0) {  // the next tick reads its times from the other parity
      c.cold_d[COLD_T0 + (parity ^ 1)] = next_t;
      c.cold_d[COLD_PT0 + (parity ^ 1)] = t;
    }
    group_sync(c);
    bool resort = false;
    if (SORTED && need_coll) {
      resort = c.sflag[3] != 0;  // (uniform: set before the barrier above or at launch)
      if (!resort) {
        if (s < M) sorted_scatter(c);
        group_sync(c);
      }
    }
    // ---- phase B1: callbacks (RSS) + broad phase
    if (LEAN && lag && k > 0)  // the previous tick's books (its narrow phase is complete: it ran before this tick's barrier)
      lagged_epilogue(p, st, c, n + sc.scenario_base, s, W, G, ego_slot, parity ^ 1, tick - 1, t);
    if (live && present) {
      if (RSS && feat_rss) {  // RSSDistances.__call__, callback.py:57-122
        rss_last = SG_RSS_NONE;
        rss_evald = false;
        if (next_t != 0.0 && s != ego_slot && c.egop[EGO_PRESENT] != 0.0) {
          RssConst KR;
          KR.CLR = p.rss_min_safe_clearance; KR.RT = p.rss_response_time;
          KR.MAXA = p.rss_max_long_accel; KR.MINA = p.rss_min_long_accel;
          KR.r2mina = c.cold_d[COLD_R2MINA];
          // (lean rollouts: the safe ratios are pure outputs - computed once after the last tick)
          rss_last = (uint8_t)rss_hazard<!LEAN>(KR, c, x, y, vx, vy, rss_state, tc, G);
          rss_evald = true;
          const int found = (rss_state >> 2) & 3;  // RSS metric latch, rss.py:71-103
          if (found) atomicOr(&c.acc[parity * ACC_N + ACC_RSS], found == 2 ? 1 : 2);
        }
      }
      if (need_coll && !SORTED) broad_phase(c, parity);
    }
    if (SORTED && need_coll && !resort && s < M) broad_phase_sorted<true>(c, parity);
    group_sync(c);
    if (SORTED && need_coll && (resort || c.sflag[3] != 0)) {
      // the one-pass re-ranking did not verify (or the launch starts unsorted): rebuild the order from
      // the boxes in slot order with full transposition rounds and redo the sweep
      group_sync(c);  // everyone has read the flag
      if (s == 0) { c.sflag[3] = 0; c.acc[parity * ACC_N + ACC_QCOUNT] = 0; }
      if (s < M) { c.aabb[s] = c.tmpbox[s]; c.sid[s] = (uint16_t)s; }
      group_sync(c);
      sort_positions(c, sort_round);
      group_sync(c);
      if (s < M) broad_phase_sorted<false>(c, parity);
      group_sync(c);
    }
    if (LEAN && lag) {
      narrow_phase_lagged(p, st, c, s, G, ego_slot, first_slot, parity, live && present);
      if (c.acc[parity * ACC_N + ACC_QCOUNT] > c.QCAP) group_sync(c);  // (queue overflow: the redo reads this tick's AABBs)
      const double ta = c.cold_d[COLD_T0 + (parity ^ 1)], dta = ta - c.cold_d[COLD_PT0 + (parity ^ 1)];
      done = (p.terminal & SG_TERM_MAX_LENGTH) && (ta + dta > c.cold_d[COLD_LEN]);  // state.py:397-398
      if (s == ego_slot && (p.features & SG_FEAT_EGO_METRICS)) {  // metrics/trajectory.py:20-24,39-42,58-60
        double* m = c.cold_d;
        const double sp = fast_sqrt(vx * vx + vy * vy);
        const double w = ta > 0.0 ? div_r(m[COLD_AVG_T], ta, fast_rcp(ta)) : m[COLD_AVG_T] / ta;
        m[COLD_AVG] += (1.0 - w) * (sp - m[COLD_AVG]);
        m[COLD_AVG_T] = ta;
        m[COLD_MAX] = fmax(sp, m[COLD_MAX]);
        m[COLD_EGOD] = dist;
      }
    } else {
      done = finish_tick<true>(p, st, c, n + sc.scenario_base, s, W, G, ego_slot, first_slot, parity, tick,
                         c.cold_d[COLD_T0 + (parity ^ 1)], c.cold_d[COLD_T0 + (parity ^ 1)] - c.cold_d[COLD_PT0 + (parity ^ 1)],
                         c.cold_d[COLD_LEN], live, live && present, collided, vx, vy, 0.0, dist);
    }
    parity ^= 1;
  }
  if (LEAN && lag && tick > tick0) {  // the last tick's books, and the slots that collided during the launch
    group_sync(c);
    lagged_epilogue(p, st, c, n + sc.scenario_base, s, W, G, ego_slot, parity ^ 1, tick, c.cold_d[COLD_T0 + parity]);
    if (live && ((c.bits[s >> 5] >> (s & 31)) & 1)) collided = 1;
  }

  if (RSS && LEAN && live && present && rss_evald) rss_ratios(c, x, y, tc, G);
  if (live) {
    const int64_t i = c.i, nm = c.nm;
    st.pose[i] = x; st.pose[nm + i] = y; st.pose[3 * nm + i] = h;
    st.vel[i] = vx; st.vel[nm + i] = vy; st.vel[3 * nm + i] = tc[4 * G];
    if (tick > tick0 && present) { st.vel[2 * nm + i] = 0.0; st.vel[4 * nm + i] = 0.0; st.vel[5 * nm + i] = 0.0; }
    st.dist[i] = dist;
    st.speed[i] = speed;
    st.collided[i] = collided;
    if (RSS) {
      st.rss_state[i] = rss_state; st.rss_last[i] = rss_last;
      st.safe_dist[i] = tc[0]; st.safe_dist[nm + i] = tc[G];
      st.safe_ratio[i] = tc[2 * G]; st.safe_ratio[nm + i] = tc[3 * G];
    }
  }
  if (s == 0) {
    st.t[n] = c.cold_d[COLD_T0 + parity]; st.prev_t[n] = c.cold_d[COLD_PT0 + parity];
    st.tick[n] = tick; st.done[n] = done;
  }
  store_cold(st, c, n, s, W, ego_slot);
}

template <bool RSS>
static cudaError_t launch_vehicle_t(int n_scen, cudaStream_t s, const SgScene& sc, const SgParams& p,
                                    const SgState& st, const SgInputs& in, const SgRngDev& rng, int act,
                                    int n_ticks, const GroupLayout& L) {
  void (*kern)(SgScene, SgParams, SgState, SgInputs, int, GroupLayout, SgRngDev);
  int threads;
  const bool lean = sgi_vehicle_lean(p, st);
  if (!lean && act != ACT_F64) return cudaErrorNotSupported;  // sg_api.cu rejects this combination first
  // lean: collisions on, no trace, no pair matrix - those code paths are compiled out; the action
  // source (fp64 table / device PCG64 stream / fp32 table) is a compile-time variant of the lean kernels
#define SG_VEH_PICK(T, B, S)                                                                   \
  (!lean ? sg_vehicle_kernel<RSS, T, B, S, false, ACT_F64>                                     \
         : act == ACT_RNG ? sg_vehicle_kernel<RSS, T, B, S, true, ACT_RNG>                     \
         : act == ACT_F32 ? sg_vehicle_kernel<RSS, T, B, S, true, ACT_F32>                     \
                          : sg_vehicle_kernel<RSS, T, B, S, true, ACT_F64>)
  // lean kernels specialised for a slot count (compile-time layout): the device-side action source and the fp64 table
#define SG_VEH_PICK_M(T, B, S, MC_)                                                            \
  (act == ACT_RNG ? sg_vehicle_kernel<RSS, T, B, S, true, ACT_RNG, MC_>                        \
                  : sg_vehicle_kernel<RSS, T, B, S, true, ACT_F64, MC_>)
  const int M = sc.n_slots;
  const bool spec = lean && (act == ACT_RNG || act == ACT_F64);
  // (the sorted sweep is a compile-time variant too: scenes of up to 128 slots carry none of its code)
  if (L.G <= SG_VEH_THREADS) {
    kern = (spec && M == 64) ? SG_VEH_PICK_M(SG_VEH_THREADS, SG_VEH_MINB, false, 64)
         : (spec && M == 128) ? SG_VEH_PICK_M(SG_VEH_THREADS, SG_VEH_MINB, false, 128)
                              : SG_VEH_PICK(SG_VEH_THREADS, SG_VEH_MINB, false);
    threads = SG_VEH_THREADS;
  } else if (L.G <= SG_THREADS) {
    kern = (spec && M == 256 && L.sorted) ? SG_VEH_PICK_M(SG_THREADS, 2, true, 256)
         : L.sorted ? SG_VEH_PICK(SG_THREADS, 2, true) : SG_VEH_PICK(SG_THREADS, 2, false);
    threads = SG_THREADS;
  }
  else { kern = L.sorted ? SG_VEH_PICK(1024, 1, true) : SG_VEH_PICK(1024, 1, false); threads = L.G; }
#undef SG_VEH_PICK
#undef SG_VEH_PICK_M
  const int gpb = threads / L.G;
  const int blocks = (n_scen + gpb - 1) / gpb;
  const size_t smem = (size_t)gpb * L.bytes;
  if (smem > 48 * 1024) {
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
  }
  kern<<<blocks, threads, smem, s>>>(sc, p, st, in, n_ticks, L, rng);
  return cudaGetLastError();
}
