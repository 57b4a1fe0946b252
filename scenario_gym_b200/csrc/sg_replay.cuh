// sg_replay.cuh -- tick-parallel rollout of replay-only scenes (C1 / C2).
//
// When every slot of a scene is a BatchReplayEntity (entity/batch.py:34-128) or a
// ReplayTrajectoryAgent (agent.py:118-128), nothing a tick computes depends on the previous tick's
// *poses*: poses and presence are pure functions of the tick times, and the only cross-tick
// state is a handful of running reductions (tick times, distance, EgoAvgSpeed, the previous
// tick's ego collision row).  So instead of walking the ticks of a scenario one after another
// (latency bound: ~700 dependent ticks, a third of the lanes active), one CTA takes a scenario and
// its threads take 124 consecutive TICKS at a time:
//   1  one thread extends the tick times by repeated addition (scenario_gym.py:229 - the running
//      fp64 sum is part of the reference's results) while the previous chunk is committed;
//      `max_length` is then decided for all ticks of the chunk at once;
//   2  every thread evaluates its time: presence and poses of all slots (uniform loops over the
//      slots: no divergence between entity kinds; the previous pose comes from the neighbouring
//      lane), distance increments, ego speed, staged conservative fp32 AABBs, the pair sweep and the
//      narrow phase of its survivors from the poses staged in shared memory (separating-axis
//      filter, exact predicate for knife edges);
//   3  the chunk is committed: ordered reductions over the ticks (distance, EgoAvgSpeed recurrence
//      with the same operation order as metrics/trajectory.py:20-24, collision rising edges,
//      first collision, terminal conditions) and the State rows of the entities that left / the
//      final tick.
// Results equal the sequential kernel's bit for bit except the accumulated distances, which are
// summed as a fixed tree over the ticks (<= 1e-15 relative, deterministic).
#pragma once

#define SG_RP_THREADS 128
#define SG_RP_WARPS (SG_RP_THREADS / 32)
// Lane 0 of every warp only supplies the pose "before" the warp's first tick (what lane l - 1
// computed is the previous pose of lane l), so a warp advances 31 ticks and a chunk 124.
#define SG_RP_TPC (SG_RP_WARPS * 31)
// A fifth warp runs the two serial recurrences off the critical path: while the workers evaluate
// chunk c, it extends the tick times for chunk c + 1 and folds chunk c - 1 into EgoAvgSpeed.
#define SG_RP_BLOCK (SG_RP_THREADS + 32)

struct RpUnion {  // position of a time in the scenario's union-knot table
  int mode;       // 0: before the first knot, 1: after the last, 2: interpolate rows cur-1, cur
  int cur;
  double w1, w0;
};

// min(max(search_left(x, stride, K, t), 1), K - 1) for K >= 2, started from where t would sit if the knots
// were evenly spaced (x0 = x[0], xN = x[K - 1]): recorded trajectories mostly are, and then two probes replace
// the log2(K) dependent loads of the bisection.  A wrong guess costs three more probes and then bisects the
// side of the guess the answer is on -- the result is the bisection's for any knot sequence.
SG_DEV int rp_search_guess(const double* __restrict__ x, int stride, int K, double t, double x0, double xN) {
  const float span = (float)(xN - x0);
  int g = 1;
  if (span > 0.f) g = 1 + (int)((float)(t - x0) * (float)(K - 1) / span);
  g = min(max(g, 1), K - 1);
  if (__ldg(x + (int64_t)(g - 1) * stride) < t) {  // the answer is >= g
#pragma unroll 1
    for (int k = 0; k < 3; ++k) {
      if (g >= K - 1 || !(__ldg(x + (int64_t)g * stride) < t)) return g;
      ++g;
    }
    int lo = g, hi = K;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (__ldg(x + (int64_t)mid * stride) < t) lo = mid + 1; else hi = mid;
    }
    return min(lo, K - 1);
  }
  int lo = 0, hi = g - 1;  // x[g - 1] >= t: the first such index is in [0, g - 1]
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(x + (int64_t)mid * stride) < t) lo = mid + 1; else hi = mid;
  }
  return max(lo, 1);
}

SG_DEV RpUnion rp_union_weights(const double* __restrict__ ts, int UK, double t) {
  RpUnion u;
  u.mode = 0; u.cur = 1; u.w1 = 0.0; u.w0 = 0.0;
  if (UK <= 0) return u;
  const double t_first = __ldg(ts), t_last = __ldg(ts + UK - 1);
  if (t < t_first) return u;  // fill_value = (X[0], X[-1]), entity/batch.py:120-127
  if (t > t_last) { u.mode = 1; return u; }
  u.mode = 2;
  u.cur = UK >= 2 ? rp_search_guess(ts, 1, UK, t, t_first, t_last) : min(max(search_left(ts, 1, UK, t), 1), UK - 1);
  const double x_lo = __ldg(ts + u.cur - 1), x_hi = __ldg(ts + u.cur);
  u.w1 = (t - x_lo) / (x_hi - x_lo);
  u.w0 = (x_hi - t) / (x_hi - x_lo);
  return u;
}
SG_DEV double rp_union_value(const RpUnion& u, const double* __restrict__ X, int UK, int M, int s, int f) {
  if (u.mode == 0) return __ldg(X + f * M + s);
  if (u.mode == 1) return __ldg(X + ((int64_t)(UK - 1) * 6 + f) * M + s);
  const double* lo = X + (int64_t)(u.cur - 1) * 6 * M + s;
  return u.w1 * __ldg(lo + (6 + f) * M) + u.w0 * __ldg(lo + f * M);
}

struct RpSlot {
  int kind, K;
  const double* rows;
  double tmin, tmax;
};
SG_DEV RpSlot rp_slot(const SgScene& sc, int64_t i) {
  RpSlot e;
  e.kind = sc.kind[i];
  const int64_t r0 = sc.traj_off[i];
  e.K = (int)(sc.traj_off[i + 1] - r0);
  e.rows = sc.traj_rows + r0 * 7;
  e.tmin = e.K ? __ldg(e.rows) : 0.0;
  e.tmax = e.K ? __ldg(e.rows + (int64_t)(e.K - 1) * 7) : 0.0;
  return e;
}
// presence after a tick to time tau (scenario_gym.py:233-245, entity/batch.py:47-52); for an
// agent slot `agent_pres` = present at launch or inserted by the first tick (its trajectory
// starts at or after the launch time), which then holds for every later tick
SG_DEV bool rp_present(const RpSlot& e, int persist, double tau, bool agent_pres) {
  if (e.kind == SG_KIND_AGENT_REPLAY) return agent_pres;
  return persist || e.K == 1 || (tau >= e.tmin && tau <= e.tmax);
}
template <int NF>
SG_DEV void rp_pose(const RpSlot& e, const RpUnion& u, const double* __restrict__ X, int UK, int M, int s,
                    double tau, double out[6]) {
  if (e.kind == SG_KIND_AGENT_REPLAY) {  // agent.py:125-128, extrapolate=(False, False)
    int cur = 0;
    double full[6];
    position_at_t(e.rows, e.K, tau, EXT_CLAMP, cur, full);
#pragma unroll
    for (int f = 0; f < NF; ++f) out[f] = full[f];
  } else {
#pragma unroll
    for (int f = 0; f < NF; ++f) out[f] = rp_union_value(u, X, UK, M, s, f);
  }
}

struct RpCtx {  // per-scenario constants
  int n, M, ego_slot, first_slot, UK;
  const double* ts;
  const double* X;
  double t_launch, ox, oy;
};

// full pose of slot s after the tick (tkm -> tk) and the pose State.update_poses differences it
// against (state.py:203-239): last tick's pose, or for a newcomer its trajectory extrapolated
// to the previous time.  `first`: the previous tick is the launch state (read from State).
SG_DEV void rp_pose_and_prev(const SgScene& sc, const SgParams& p, const SgState& st, const RpCtx& c,
                             const RpSlot& e, int s, bool agent_pres, double tk, double tkm, bool first,
                             double pk[6], double prev[6]) {
  const int64_t i = (int64_t)c.n * c.M + s, nm = sc.plane_stride;
  const RpUnion uk = rp_union_weights(c.ts, c.UK, tk);
  rp_pose<6>(e, uk, c.X, c.UK, c.M, s, tk, pk);
  bool ppres;
  if (first) {
    ppres = st.present[i] != 0;
    if (ppres) {
#pragma unroll
      for (int f = 0; f < 6; ++f) prev[f] = st.pose[f * nm + i];
    }
  } else {
    ppres = rp_present(e, p.persist, tkm, agent_pres);
    if (ppres) {
      const RpUnion um = rp_union_weights(c.ts, c.UK, tkm);
      rp_pose<6>(e, um, c.X, c.UK, c.M, s, tkm, prev);
    }
  }
  if (!ppres) {  // state.py:219-222
    int cur = 0;
    position_at_t(e.rows, e.K, tkm, EXT_TRUE, cur, prev);
  }
}

// State rows of slot s as they stand after the tick (tkm -> tk); rare (once per slot per launch)
static __device__ __noinline__ void rp_write_rows(const SgScene* sc, const SgParams* p, const SgState* st,
                                           RpCtx c, int s, bool agent_pres, double tk, double tkm,
                                           bool first) {
  const int64_t i = (int64_t)c.n * c.M + s, nm = sc->plane_stride;
  const RpSlot e = rp_slot(*sc, i);
  double pk[6], prev[6];
  rp_pose_and_prev(*sc, *p, *st, c, e, s, agent_pres, tk, tkm, first, pk, prev);
  const double dt = tk - tkm;
#pragma unroll
  for (int f = 0; f < 6; ++f) {
    st->pose[f * nm + i] = pk[f];
    st->vel[f * nm + i] = (pk[f] - prev[f]) / dt;  // state.py:234-236
  }
}

// out-of-line so that the tick loop carries one copy of the control-point search
static __device__ __noinline__ double4 rp_agent_pose4(const double* rows, int K, double t, int mode) {
  int cur = 0;  // (position_at_t verifies the cursor it is given: two probes when it is the answer)
  if (K >= 2) cur = rp_search_guess(rows, 7, K, t, __ldg(rows), __ldg(rows + (int64_t)(K - 1) * 7));
  double full[6];
  position_at_t(rows, K, t, mode, cur, full);
  return make_double4(full[0], full[1], full[2], full[3]);
}
static __device__ __noinline__ RpUnion rp_union_weights_ool(const double* ts, int UK, double t) {
  return rp_union_weights(ts, UK, t);
}

struct RpDesc {  // per-slot constants staged in shared memory
  const double* rows;
  double tmin, tmax, bw, bl, bcx, bcy;
  int K, kind;
};
SG_DEV bool rp_present_d(const RpDesc& e, int persist, double tau, bool agent_pres) {
  if (e.kind == SG_KIND_AGENT_REPLAY) return agent_pres;
  return persist || e.K == 1 || (tau >= e.tmin && tau <= e.tmax);
}

// fp64 corners of a slot from its staged pose (Entity.get_bounding_box_points, entity/base.py:100-138; the
// same expression as publish_box, on the cos / sin the tick's own libm call returned)
SG_DEV void rp_corners_staged(const RpDesc& e, double2 xy, double2 hc, double q[8]) {
  const double x = xy.x, y = xy.y, cs = hc.x, sn = hc.y;
  const double hx0 = e.bcx - 0.5 * e.bl, hx1 = e.bcx + 0.5 * e.bl;
  const double hy0 = e.bcy + 0.5 * e.bw, hy1 = e.bcy - 0.5 * e.bw;
  q[0] = x + (hx0 * cs + hy0 * -sn); q[1] = y + (hx0 * sn + hy0 * cs);
  q[2] = x + (hx1 * cs + hy0 * -sn); q[3] = y + (hx1 * sn + hy0 * cs);
  q[4] = x + (hx1 * cs + hy1 * -sn); q[5] = y + (hx1 * sn + hy1 * cs);
  q[6] = x + (hx0 * cs + hy1 * -sn); q[7] = y + (hx0 * sn + hy1 * cs);
}

SG_DEV Obb rp_obb(const RpDesc& e, double2 xy, double2 hc) {
  Obb o;
  const double cs = hc.x, sn = hc.y, hl = 0.5 * e.bl, hw = 0.5 * e.bw;
  o.cx = xy.x + (e.bcx * cs - e.bcy * sn); o.cy = xy.y + (e.bcx * sn + e.bcy * cs);
  o.ux = hl * cs; o.uy = hl * sn;
  o.vx = -hw * sn; o.vy = hw * cs;
  return o;
}
// narrow phase of one AABB-surviving pair of a tick (state/utils.py:10-49, utils.py:28-62) from the poses the
// slot loop staged in shared memory: nothing is re-read from global memory, no control-point search, no
// second sincos.  The branch-free separating-axis filter decides all but knife-edge contacts; those (and
// coincident boxes, `g != g_prime` utils.py:58) build the fp64 corners and go to the exact predicate.
static __device__ __noinline__ bool rp_pair_exact_staged(const RpDesc* desc, const double2* xy, const double2* hc,
                                                         int a, int b, int col) {
  const RpDesc ea = desc[a], eb = desc[b];
  double qa[8], qb[8];
  rp_corners_staged(ea, xy[a * SG_RP_THREADS + col], hc[a * SG_RP_THREADS + col], qa);
  rp_corners_staged(eb, xy[b * SG_RP_THREADS + col], hc[b * SG_RP_THREADS + col], qb);
  bool same = true;
#pragma unroll
  for (int f = 0; f < 8; ++f) same = same && qa[f] == qb[f];
  if (same) return false;
  const Quad A = quad_from_array(qa), B = quad_from_array(qb);
  const int ha = box_orientation_hint(ea.bw, ea.bl), hb = box_orientation_hint(eb.bw, eb.bl);
  return quads_intersect(A, ha ? ha : quad_orientation(A), B, hb ? hb : quad_orientation(B));
}
SG_DEV bool rp_pair_staged(const RpDesc* desc, const double2* xy, const double2* hc, int a, int b, int col) {
  const Obb A = rp_obb(desc[a], xy[a * SG_RP_THREADS + col], hc[a * SG_RP_THREADS + col]);
  const Obb B = rp_obb(desc[b], xy[b * SG_RP_THREADS + col], hc[b * SG_RP_THREADS + col]);
  const int v = sat_classify_obb(A, B);
  if (v != 0) return v > 0;
  return rp_pair_exact_staged(desc, xy, hc, a, b, col);
}

struct RpCarry {  // cross-chunk state of the scenario (shared memory)
  double avg, avg_t, mx, last_sp;
  long long pair_ticks;
  int tick, executed, done, cnt, cnt_next, nv, nv_prev, first_tick, fp0, fp1, end_here, buf;
  uint32_t collided, chunk_first, term_first;
};

// EgoAvgSpeed / EgoMaxSpeed over one chunk in tick order (metrics/trajectory.py:20-24, 39-42);
// t_last = time after the chunk's last tick
SG_DEV void rp_ego_metrics(RpCarry* car, const double* spv, const double* cwv, int nv, double t_last) {
  double avg = car->avg, mx = car->mx, last = car->last_sp;
  for (int q = 1; q <= nv; ++q) {
    double sp = spv[q];
    if (sp != sp) sp = last; else last = sp;  // ego absent: its velocity is unchanged
    avg += cwv[q] * (sp - avg);
    mx = fmax(sp, mx);
  }
  car->avg = avg; car->mx = mx; car->last_sp = last; car->avg_t = t_last;
}

template <bool MATRIX>  // MATRIX: also write the pair matrix of the final tick (SG_FEAT_COLL_MATRIX)
#ifndef SG_RP_MINB
#define SG_RP_MINB 3  // (the staged poses bring a 9-slot scene to 64 KB of shared memory: three CTAs per SM, 128 registers)
#endif
__global__ void __launch_bounds__(SG_RP_BLOCK, SG_RP_MINB)
sg_replay_kernel(const __grid_constant__ SgScene sc, const __grid_constant__ SgParams p,
                 const __grid_constant__ SgState st, int n_ticks) {
  extern __shared__ __align__(16) unsigned char smem[];
  constexpr int CH = SG_RP_THREADS, TPC = SG_RP_TPC, TS = TPC + 2;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int j = 31 * warp + lane;  // time index inside the chunk: T[j]; lanes >= 1 own tick j
  RpCtx c;
  c.n = blockIdx.x; c.M = sc.n_slots;
  const int M = c.M, n = c.n;
  const int64_t nm = sc.plane_stride, i0 = (int64_t)n * M;
  c.ego_slot = sc.ego_slot[n]; c.first_slot = sc.first_slot[n];
  {
    const int64_t u0 = sc.union_off[n];
    c.UK = (int)(sc.union_off[n + 1] - u0);
    c.ts = sc.union_t + u0;
    c.X = sc.union_x + u0 * 6 * M;
    const int64_t er = sc.traj_off[i0 + c.ego_slot];  // origin of the fp32 bounds
    c.ox = __ldg(sc.traj_rows + er * 7 + 1);
    c.oy = __ldg(sc.traj_rows + er * 7 + 2);
  }
  c.t_launch = st.t[n];
  const bool need_coll = (p.features & SG_FEAT_COLLISIONS) ||
                         (p.terminal & (SG_TERM_COLLISION | SG_TERM_EGO_COLLISION));
  const bool coll_terminal = (p.terminal & (SG_TERM_COLLISION | SG_TERM_EGO_COLLISION)) != 0;
  const int limit = n_ticks < 0 ? p.max_ticks : n_ticks;
  const double length = sc.length[n];

  // shared layout
  double* Tb = (double*)smem;                        // [2][TS] tick times (double buffered); T[0] = time before the chunk
  double* spb = Tb + 2 * TS;                         // [2][TS] ego speed after tick j (NaN: ego absent)
  double* cwb = spb + 2 * TS;                        // [2][TS] 1 - t_prev / t of EgoAvgSpeed
  double* part = cwb + 2 * TS;                       // [SG_RP_WARPS][M] distance partial sums
  double* cdist = part + SG_RP_WARPS * M;            // [M] accumulated distance
  RpDesc* desc = (RpDesc*)(cdist + M + (M & 1));     // [M]
  float4* aabb = (float4*)(desc + M);                // [M][CH]
  double2* sxy = (double2*)(aabb + (size_t)M * CH);  // [M][CH] position of a present slot after the thread's tick
  double2* shc = sxy + (size_t)M * CH;               // [M][CH] cos / sin of its heading
  uint32_t* egonow = (uint32_t*)(shc + (size_t)M * CH);  // [TS]; [0] = ego row before the chunk
  uint32_t* cbits = egonow + TS;                     // [TS]
  uint32_t* fpair = cbits + TS;                      // [TS]
  int* npairs = (int*)(fpair + TS);                  // [TS]
  uint8_t* apres = (uint8_t*)(npairs + TS);          // [M] agent slots: present from the first tick on
  RpCarry* car = (RpCarry*)(apres + ((M + 15) / 16) * 16);

  if (tid < M) {
    const int64_t i = i0 + tid;
    cdist[tid] = st.dist[i];
    const RpSlot e = rp_slot(sc, i);
    apres[tid] = (st.present[i] != 0) || (e.K > 0 && e.tmin >= c.t_launch);  // scenario_gym.py:240-244
    RpDesc d;
    d.rows = e.rows; d.tmin = e.tmin; d.tmax = e.tmax; d.K = e.K; d.kind = e.kind;
    d.bw = sc.box[i]; d.bl = sc.box[nm + i]; d.bcx = sc.box[2 * nm + i]; d.bcy = sc.box[3 * nm + i];
    desc[tid] = d;
  }
  if (tid == 0) {
    car->avg = st.ego_avg_speed[n]; car->avg_t = st.ego_avg_t[n]; car->mx = st.ego_max_speed[n];
    {
      const int64_t ie = i0 + c.ego_slot;
      car->last_sp = norm3(st.vel[ie], st.vel[nm + ie], st.vel[2 * nm + ie]);
    }
    car->pair_ticks = st.n_pair_ticks[n];
    car->tick = st.tick[n]; car->executed = 0; car->done = st.done[n] != 0;
    car->first_tick = st.first_coll_tick[n];
    car->fp0 = st.first_coll_pair[2 * n]; car->fp1 = st.first_coll_pair[2 * n + 1];
    egonow[0] = st.ego_hits[n];
    car->collided = 0;
    car->buf = 0;
    car->nv_prev = 0;
    // tick times of the first chunk by repeated addition (scenario_gym.py:229)
    int cnt = 0;
    double t = c.t_launch;
    Tb[0] = t;
    if (!car->done)
      while (cnt < TPC && cnt < limit) { t = t + p.timestep; Tb[++cnt] = t; }
    Tb[cnt + 1] = t + p.timestep;
    car->cnt = cnt;
  }
  for (int q = tid; q < SG_RP_WARPS * M; q += SG_RP_BLOCK) part[q] = 0.0;
  __syncthreads();
  // slots past the last non-empty one take no part in any loop below (scenes padded to the batch's slot count)
  int ML = 0;
  for (int s = 0; s < M; ++s)
    if (desc[s].kind != SG_KIND_EMPTY) ML = s + 1;

  for (;;) {
    const int buf = car->buf;
    const double* T = Tb + buf * TS;
    double* spv = spb + buf * TS;
    double* cwv = cwb + buf * TS;
    const int cnt = car->cnt;
    if (cnt == 0) break;
    const bool first_chunk = car->executed == 0;
    // ---- 1: `max_length` (state.py:397-398) decided per tick in parallel ------------------------
    if (tid == 0) { car->nv = cnt; car->term_first = 0xffffffffu; car->chunk_first = 0xffffffffu; }
    __syncthreads();
    if (p.terminal & SG_TERM_MAX_LENGTH) {
      const bool over = lane >= 1 && j <= cnt && (T[j] + (T[j] - T[j - 1]) > length);
      const uint32_t b = __ballot_sync(0xffffffffu, over);
      if (b && lane == 0) atomicMin(&car->nv, 31 * warp + __ffs(b) - 1);
    }
    __syncthreads();
    int nv = car->nv;
    bool done = nv < cnt || ((p.terminal & SG_TERM_MAX_LENGTH) && (T[nv] + (T[nv] - T[nv - 1]) > length));

    // ---- helper warp: serial recurrences, overlapped with the workers' pass ----------------------
    if (tid == SG_RP_THREADS && (p.features & SG_FEAT_EGO_METRICS) && car->nv_prev > 0)
      rp_ego_metrics(car, spb + (buf ^ 1) * TS, cwb + (buf ^ 1) * TS, car->nv_prev, T[0]);
    if (tid == SG_RP_THREADS + 1) {  // tick times of the next chunk by repeated addition (scenario_gym.py:229)
      double* Tn = Tb + (buf ^ 1) * TS;
      int cn = 0;
      double t = T[cnt];
      Tn[0] = t;
      const int left = limit - (car->executed + cnt);
      if (cnt == TPC)
        while (cn < TPC && cn < left) { t = t + p.timestep; Tn[++cn] = t; }
      Tn[cn + 1] = t + p.timestep;
      car->cnt_next = cn;
    }

    uint32_t leave = 0;  // slots present after this thread's tick but not after the next one
    for (int attempt = 0; attempt < 2; ++attempt) {
      // ---- 2: every thread evaluates its time ----------------------------------------------------
      if (warp < SG_RP_WARPS) {
        leave = 0;
        const bool live_t = j <= nv;             // T[j] is a time of this chunk
        const bool valid = lane >= 1 && live_t;  // this thread owns tick j
        const double tj = live_t ? T[j] : T[0], tkm = valid ? T[j - 1] : T[0];
        const double dt = tj - tkm;
        const bool first = first_chunk && j == 1;
        const RpUnion uk = rp_union_weights_ool(c.ts, c.UK, tj);
        uint32_t pm = 0;  // present slots after this tick
        double sp = NAN;
        for (int s = 0; s < ML; ++s) {
          const RpDesc e = desc[s];
          if (e.kind == SG_KIND_EMPTY) {
            if (need_coll) aabb[(size_t)s * CH + tid] = make_float4(INFINITY, INFINITY, -INFINITY, -INFINITY);
            if (lane == 0) part[warp * M + s] = 0.0;
            continue;  // uniform over the CTA
          }
          const bool ap = apres[s] != 0;
          const bool pres = live_t && rp_present_d(e, p.persist, tj, ap);
          if (valid && pres && e.kind != SG_KIND_AGENT_REPLAY && !rp_present_d(e, p.persist, T[j + 1], ap))
            leave |= 1u << s;
          double4 pk = make_double4(0.0, 0.0, 0.0, 0.0);
          if (pres) {
            if (e.kind == SG_KIND_AGENT_REPLAY) pk = rp_agent_pose4(e.rows, e.K, tj, EXT_CLAMP);  // agent.py:125-128
            else {
              pk.x = rp_union_value(uk, c.X, c.UK, M, s, 0); pk.y = rp_union_value(uk, c.X, c.UK, M, s, 1);
              pk.z = rp_union_value(uk, c.X, c.UK, M, s, 2); pk.w = rp_union_value(uk, c.X, c.UK, M, s, 3);
            }
          }
          // what lane l - 1 computed is the pose before this lane's tick
          bool ppres = __shfl_up_sync(0xffffffffu, (int)pres, 1) != 0;
          double px = __shfl_up_sync(0xffffffffu, pk.x, 1), py = __shfl_up_sync(0xffffffffu, pk.y, 1),
                 pz = __shfl_up_sync(0xffffffffu, pk.z, 1);
          float4 bb = make_float4(INFINITY, INFINITY, -INFINITY, -INFINITY);
          double inc = 0.0;
          if (valid && pres) {
            pm |= 1u << s;
            if (first) {  // the previous tick is the launch state
              const int64_t i = i0 + s;
              ppres = st.present[i] != 0;
              if (ppres) { px = st.pose[i]; py = st.pose[nm + i]; pz = st.pose[2 * nm + i]; }
            }
            if (!ppres) {  // newcomer: state.py:219-222
              const double4 q = rp_agent_pose4(e.rows, e.K, tkm, EXT_TRUE);
              px = q.x; py = q.y; pz = q.z;
            }
            const double d0 = pk.x - px, d1 = pk.y - py, d2 = pk.z - pz;
            inc = norm3(d0, d1, d2);  // state.py:237-239
            if (s == c.ego_slot) sp = norm3(d0 / dt, d1 / dt, d2 / dt);  // metrics/trajectory.py:21
            if (need_coll) {
              double sn, cs;
              sincos(pk.w, &sn, &cs);
              bb = make_aabb_box(pk.x, pk.y, cs, sn, e.bw, e.bl, e.bcx, e.bcy, c.ox, c.oy);
              sxy[(size_t)s * CH + tid] = make_double2(pk.x, pk.y);  // for the narrow phase (rp_pair_staged)
              shc[(size_t)s * CH + tid] = make_double2(cs, sn);
            }
          }
          if (need_coll) aabb[(size_t)s * CH + tid] = bb;
          // distance: fixed tree over the ticks of the warp, then over the warps in order
  #pragma unroll
          for (int off = 16; off > 0; off >>= 1) inc += __shfl_xor_sync(0xffffffffu, inc, off);
          if (lane == 0) part[warp * M + s] = inc;
        }
        // pair sweep on the conservative AABBs, narrow phase on the survivors (from the staged poses)
        uint32_t cb = 0, en = 0, fp = 0x7fffffffu, term = 0;
        int np = 0;
        if (need_coll && valid) {
          for (int a = 0; a < ML; ++a) {
            if (!((pm >> a) & 1)) continue;
            const float4 A = aabb[(size_t)a * CH + tid];
            for (int b = a + 1; b < ML; ++b) {
              if (!((pm >> b) & 1)) continue;
              const float4 B = aabb[(size_t)b * CH + tid];
              if (!(A.x <= B.z && B.x <= A.z && A.y <= B.w && B.y <= A.w)) continue;
              if (!rp_pair_staged(desc, sxy, shc, a, b, tid)) continue;
              ++np;
              fp = min(fp, ((uint32_t)a << 16) | (uint32_t)b);
              cb |= (1u << a) | (1u << b);
              if (a == c.ego_slot) en |= 1u << b;
              if (b == c.ego_slot) en |= 1u << a;
              if (a == c.first_slot || b == c.first_slot) term |= 2u;
            }
          }
          if ((p.terminal & SG_TERM_COLLISION) && np > 0) term |= 1u;
          if ((p.terminal & SG_TERM_EGO_COLLISION) && (term & 2u)) term |= 1u;
        }
        if (valid) {
          spv[j] = sp;
          cwv[j] = 1.0 - (first ? car->avg_t : tkm) / tj;  // metrics/trajectory.py:20-24
          egonow[j] = en; cbits[j] = cb; fpair[j] = fp; npairs[j] = np;
          if (coll_terminal && (term & 1u)) atomicMin(&car->term_first, (uint32_t)j);
        }
      }
      __syncthreads();
      if (!coll_terminal || car->term_first >= (uint32_t)nv) {
        if (coll_terminal && car->term_first == (uint32_t)nv) done = true;
        break;
      }
      // a collision ends the rollout inside the chunk: redo the reductions over the shorter range
      nv = (int)car->term_first;
      done = true;
      __syncthreads();
      if (tid == 0) car->term_first = 0xffffffffu;
      __syncthreads();
    }
    const bool end_here = done || car->executed + nv >= limit;
    const bool valid = warp < SG_RP_WARPS && lane >= 1 && j <= nv;

    // ---- 3: commit the chunk ----------------------------------------------------------------------
    if (tid >= 64 && tid < 64 + M) {
      const int s = tid - 64;
      double d = cdist[s];
#pragma unroll
      for (int w = 0; w < SG_RP_WARPS; ++w) d += part[w * M + s];
      cdist[s] = d;
    }
    if (need_coll) {
      uint32_t cb = valid ? cbits[j] : 0u;
      cb = __reduce_or_sync(0xffffffffu, cb);
      int np = valid ? npairs[j] : 0;
      const uint32_t hit = __ballot_sync(0xffffffffu, np > 0);
      np = __reduce_add_sync(0xffffffffu, np);
      if (lane == 0) {
        if (cb) atomicOr(&car->collided, cb);
        if (np) atomicAdd((unsigned long long*)&car->pair_ticks, (unsigned long long)np);
        if (hit) atomicMin(&car->chunk_first, (uint32_t)(31 * warp + __ffs(hit) - 1));
      }
      if ((p.features & SG_FEAT_COLLISIONS) && valid) {  // CollisionMetric._step, metrics/collision.py:70-75
        uint32_t fresh = egonow[j] & ~egonow[j - 1];
        while (fresh) {
          const int b = __ffs(fresh) - 1;
          fresh &= fresh - 1;
          const int slot = atomicAdd(st.event_count, 1);
          if (slot < st.event_cap) {
            SgEvent ev;
            ev.scenario = n + sc.scenario_base; ev.tick = car->tick + j; ev.slot = b; ev._pad = 0; ev.t = T[j];
            st.events[slot] = ev;
          }
        }
      }
    }
    // State rows of a slot that is present after tick j but not after the next one
    if (valid && !(end_here && j == nv)) {
      const double tj = T[j], tkm = T[j - 1];
      for (uint32_t lv = leave; lv; lv &= lv - 1) {
        const int s = __ffs(lv) - 1;
        rp_write_rows(&sc, &p, &st, c, s, apres[s] != 0, tj, tkm, first_chunk && j == 1);
      }
    }
    if (end_here) {
      const double tj = T[nv], tkm = T[nv - 1];
      if (tid >= 96 && tid < 96 + M) {  // final tick: one thread per slot
        const int s = tid - 96;
        const int64_t i = i0 + s;
        const RpDesc e = desc[s];
        const bool ap = apres[s] != 0;
        const bool pres = e.kind != SG_KIND_EMPTY && rp_present_d(e, p.persist, tj, ap);
        if (pres) rp_write_rows(&sc, &p, &st, c, s, ap, tj, tkm, first_chunk && nv == 1);
        st.present[i] = pres;
        st.cur_own[i] = 1;
      }
      if (MATRIX && valid && j == nv) {  // pair matrix of the final tick
        uint32_t* rows = st.coll_mask + i0;  // W = 1
        for (int s = 0; s < M; ++s) rows[s] = 0;
        uint32_t pm = 0;
        for (int s = 0; s < M; ++s) {
          const RpDesc e = desc[s];
          if (e.kind != SG_KIND_EMPTY && rp_present_d(e, p.persist, tj, apres[s] != 0)) pm |= 1u << s;
        }
        if (need_coll)
          for (int a = 0; a < M; ++a) {
            if (!((pm >> a) & 1)) continue;
            const float4 A = aabb[(size_t)a * CH + tid];
            for (int b = a + 1; b < M; ++b) {
              if (!((pm >> b) & 1)) continue;
              const float4 B = aabb[(size_t)b * CH + tid];
              if (!(A.x <= B.z && B.x <= A.z && A.y <= B.w && B.y <= A.w)) continue;
              if (!rp_pair_staged(desc, sxy, shc, a, b, tid)) continue;
              rows[a] |= 1u << b;
              rows[b] |= 1u << a;
            }
          }
      }
    }
    __syncthreads();
    if (tid == 0) {
      if (need_coll && car->first_tick < 0 && car->chunk_first != 0xffffffffu) {
        const int q = (int)car->chunk_first;
        car->first_tick = car->tick + q;
        car->fp0 = (int)(fpair[q] >> 16); car->fp1 = (int)(fpair[q] & 0xffffu);
      }
      if (p.features & SG_FEAT_COLLISIONS) egonow[0] = egonow[nv];
      car->tick += nv; car->executed += nv;
      car->done = done; car->end_here = end_here;
      car->buf ^= 1;
      car->nv_prev = nv;
      car->cnt = end_here ? 0 : car->cnt_next;
    }
    __syncthreads();
    if (end_here) {
      if (tid == 0) {  // times after the last executed tick
        st.t[n] = T[nv]; st.prev_t[n] = T[nv - 1];
        if (p.features & SG_FEAT_EGO_METRICS) rp_ego_metrics(car, spv, cwv, nv, T[nv]);
      }
      break;
    }
  }

  // ---- per-scenario results ---------------------------------------------------------------------
  __syncthreads();
  if (car->executed > 0) {
    if (tid < M) {
      const int64_t i = i0 + tid;
      st.dist[i] = cdist[tid];
      if ((car->collided >> tid) & 1) st.collided[i] = 1;
    }
    if (tid == 0) {
      st.tick[n] = car->tick; st.done[n] = car->done;
      st.cur_union[n] = 1;
      if (p.features & SG_FEAT_EGO_METRICS) {
        st.ego_avg_speed[n] = car->avg; st.ego_avg_t[n] = car->avg_t; st.ego_max_speed[n] = car->mx;
        st.ego_dist[n] = cdist[c.ego_slot];
      }
      st.first_coll_tick[n] = car->first_tick;
      st.first_coll_pair[2 * n] = car->fp0; st.first_coll_pair[2 * n + 1] = car->fp1;
      st.n_pair_ticks[n] = car->pair_ticks;
      if (p.features & SG_FEAT_COLLISIONS) st.ego_hits[n] = egonow[0];
    }
  }
}

static size_t replay_smem_bytes(int M) {
  const int CH = SG_RP_THREADS, TS = SG_RP_TPC + 2;
  size_t o = (size_t)(6 * TS + SG_RP_WARPS * M + M + (M & 1)) * sizeof(double);
  o += (size_t)M * sizeof(RpDesc);
  o += (size_t)M * CH * (sizeof(float4) + 2 * sizeof(double2));
  o += (size_t)TS * 4 * sizeof(uint32_t);
  o += (size_t)((M + 15) / 16) * 16;
  o += sizeof(RpCarry) + 16;
  return o;
}
