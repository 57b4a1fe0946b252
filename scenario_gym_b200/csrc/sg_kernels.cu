// sg_kernels.cu -- the fused per-tick rollout kernel and the C ABI (include/sg_b200.h).
//
// Mapping: one thread per entity slot; the G threads of a scenario ("group") are a
// sub-warp (M <= 32: G = next pow2, several scenarios per warp), or G/32 whole warps
// (M > 32: G = M rounded up to 32).  A 256-thread CTA carries 256/G scenarios.  Each
// thread keeps its entity's State row (pose, velocity, distance, controller speed, RSS
// history bits) in registers for all ticks of the call; per tick the group stages every
// entity's fp64 box corners + a conservative fp32 AABB in shared memory, synchronises
// (sub-warp: __syncwarp(mask); multi-warp: a named barrier per scenario) and every
// thread sweeps the scenario's boxes: fp32 AABB reject, then the exact closed-set test
// on the fp64 corners.  Collision rows are built 32 slots at a time as bit words; the
// ego row feeds CollisionMetric's rising-edge detection.  n_ticks = 1 is
// ScenarioGym.step(); n_ticks < 0 is ScenarioGym.rollout().
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <string.h>

#include "sg_device.cuh"

#define SG_THREADS 256

__constant__ double c_ngon[64][2];  // (cos, sin)(-k * 2pi/64): GEOS Point.buffer vertices

static thread_local char g_err[512];
static int set_err(const char* what, cudaError_t e) {
  snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
  return -2;
}
static int set_msg(const char* what) {
  snprintf(g_err, sizeof(g_err), "%s", what);
  return -1;
}

// ---------------------------------------------------------------------------------
struct GroupLayout {
  int G;            // threads (slots incl. padding) per scenario
  int W;            // 32-bit words per collision row
  int off_aabb;     // byte offsets inside a group's shared-memory block
  int off_ped;
  int off_flags;
  int off_orient;
  int off_ego;
  int off_hits;
  int off_acc;
  int bytes;
};

static GroupLayout make_layout(int M, bool ped) {
  GroupLayout L;
  int G;
  if (M <= 32) { G = 1; while (G < M) G <<= 1; } else { G = (M + 31) / 32 * 32; }
  L.G = G;
  L.W = (M + 31) / 32;
  int o = 8 * G * (int)sizeof(double);               // corners[8][G]
  L.off_ped = o;     o += ped ? 4 * G * (int)sizeof(double) : 0;  // x,y,vx,vy of pedestrians (old state)
  L.off_ego = o;     o += 8 * (int)sizeof(double);   // ego x,y,h,vx,vy
  L.off_aabb = o;    o += G * (int)sizeof(float4);
  L.off_hits = o;    o += 2 * L.W * (int)sizeof(uint32_t);  // ego_now[W], ego_last[W]
  L.off_acc = o;     o += 8 * (int)sizeof(int);      // 2 parities x {npairs, first_pair, first_hit, rss}
  L.off_flags = o;   o += G + 16;                    // old present|etype (ped neighbour filter); [G] = ego present
  L.off_orient = o;  o += G;                         // ring orientation of each box
  L.bytes = (o + 15) / 16 * 16;
  return L;
}

SG_DEV void group_sync(int G, int bar_id, unsigned mask) {
  if (G <= 32) __syncwarp(mask);
  else asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(G) : "memory");
}

struct TickCtx {
  int n, s, M, G, W;
  int64_t i, nm;
  double* corners;
  double* pedbuf;
  double* egobuf;
  float4* aabb;
  uint32_t* ego_now;
  uint32_t* ego_last;
  int* acc;
  uint8_t* flags;
  int8_t* orient;
};

// conservative fp32 AABB of the fp64 corners, relative to the scenario origin (ox, oy)
SG_DEV float4 make_aabb(const double* c, double ox, double oy) {
  const double lox = fmin(fmin(c[0], c[2]), fmin(c[4], c[6]));
  const double hix = fmax(fmax(c[0], c[2]), fmax(c[4], c[6]));
  const double loy = fmin(fmin(c[1], c[3]), fmin(c[5], c[7]));
  const double hiy = fmax(fmax(c[1], c[3]), fmax(c[5], c[7]));
  float4 b;
  b.x = __double2float_rd(__dsub_rd(lox, ox));
  b.y = __double2float_rd(__dsub_rd(loy, oy));
  b.z = __double2float_ru(__dsub_ru(hix, ox));
  b.w = __double2float_ru(__dsub_ru(hiy, oy));
  return b;
}

// strict interior of Point(x, y).buffer(r): GEOS 64-gon (reference state/state.py:352-372)
SG_DEV bool in_buffer(double x, double y, double r, double qx, double qy) {
  const double dx = qx - x, dy = qy - y, d2 = dx * dx + dy * dy;
  const double rin = r * 0.99879545620517241 * (1.0 - 1e-9);  // cos(pi/64): inscribed circle
  if (d2 < rin * rin) return true;
  const double rout = r * (1.0 + 1e-9);
  if (d2 > rout * rout) return false;
  for (int k = 0; k < 64; ++k) {
    const int k1 = (k + 1) & 63;
    const double ax = x + r * c_ngon[k][0], ay = y + r * c_ngon[k][1];
    const double bx = x + r * c_ngon[k1][0], by = y + r * c_ngon[k1][1];
    if (orient_sign(ax, ay, bx, by, qx, qy) >= 0) return false;
  }
  return true;
}

template <bool PED>
SG_DEV void pedestrian_step(const SgScene& sc, const SgParams& p, const TickCtx& c,
                            const double pose[6], const double vel[6], double t, double prev_t,
                            double next_t, double sight_cos, int& goal, double force[2],
                            double& speed_io, double out[6]) {
  if (!PED) return;
  const int64_t r0 = sc.route_off[c.i];
  const int R = (int)(sc.route_off[c.i + 1] - r0);
  const double* route = sc.route_xy + 2 * r0;
  double speed, heading;
  if (goal <= R - 1) {  // pedestrian/agent.py:60-62
    const double sarc = route_project(route, R, pose[0], pose[1]);
    double arc = 0.0;
    int last = 0;
    for (int k = 0; k < R; ++k) {
      if (k > 0)
        arc += norm2(__ldg(route + 2 * k) - __ldg(route + 2 * k - 2),
                     __ldg(route + 2 * k + 1) - __ldg(route + 2 * k - 1));
      if (arc <= sarc) last = k;
    }
    goal = last + 1;
  }
  if (goal <= R - 1) {
    const double speed_desired = sc.ped_speed_desired[c.i];
    // SocialForce._force_to_goal, pedestrian/social_force.py:119-138
    const double dvx = __ldg(route + 2 * goal) - pose[0], dvy = __ldg(route + 2 * goal + 1) - pose[1];
    double dn = norm2(dvx, dvy);
    if (dn == 0) dn += 0.000000001;
    const double ux = dvx / dn, uy = dvy / dn;
    const double k = 1 / p.sf_relaxation_time;
    double F0 = k * (speed_desired * ux - vel[0]), F1 = k * (speed_desired * uy - vel[1]);
    const double thr = p.ped_distance_threshold;
    double sh, ch;
    sincos(p.ped_head_rot_angle, &sh, &ch);  // viewer/utils.py:6-17
    const double* bx = c.pedbuf;
    const double* by = c.pedbuf + c.G;
    const double* bvx = c.pedbuf + 2 * c.G;
    const double* bvy = c.pedbuf + 3 * c.G;
    for (int o = 0; o < c.M; ++o) {  // state.poses order == slot order
      const uint8_t fl = c.flags[o];
      if (o == c.s || !(fl & 1) || (fl >> 1) != SG_ETYPE_PEDESTRIAN) continue;
      const double ox = bx[o], oy = by[o];
      if (!in_buffer(pose[0], pose[1], thr, ox, oy)) continue;
      const double ovx = bvx[o], ovy = bvy[o];
      const double vdx = ovx * ch + ovy * -sh, vdy = ovx * sh + ovy * ch;
      const double vn = norm2(vdx, vdy) + 0.0000000001;
      const double view0 = vdx / vn, view1 = vdy / vn;
      // _force_pedestrian_repulsion :140-176
      const double rx = pose[0] - ox, ry = pose[1] - oy, rn = norm2(rx, ry);
      const double vmag = norm2(ovx, ovy) + 0.0000000001;
      const double uox = ovx / vmag, uoy = ovy / vmag;
      const double other_step = vmag * (next_t - t);
      const double r2x = rx - other_step * uox, r2y = ry - other_step * uoy;
      const double r2n = norm2(r2x, r2y) + 0.0000000001;
      const double b = (1.0 / 2) * sqrt((rn + r2n) * (rn + r2n) - other_step * other_step);
      const double c0 = (1.0 / 4) * (1 / b) * (rn + r2n);
      const double dbx = c0 * (rx / rn + r2x / r2n), dby = c0 * (ry / rn + r2y / r2n);
      const double g = p.sf_ped_repulse_V / p.sf_ped_repulse_sigma * exp(-b / p.sf_ped_repulse_sigma);
      const double Fr0 = g * dbx, Fr1 = g * dby;
      const double Fa0 = 2 * p.sf_ped_attract_C * rx, Fa1 = 2 * p.sf_ped_attract_C * ry;
      if (p.sf_sight_weight_use) {  // _sight_weight :213-222
        double dd = dot2(view0, view1, Fr0, Fr1) / (norm2(Fr0, Fr1) + 0.0000000001);
        double w = dd >= sight_cos ? 1.0 : p.sf_sight_weight;
        F0 += w * Fr0; F1 += w * Fr1;
        dd = dot2(view0, view1, Fa0, Fa1) / (norm2(Fa0, Fa1) + 0.0000000001);
        w = dd >= sight_cos ? 1.0 : p.sf_sight_weight;
        F0 += w * Fa0; F1 += w * Fa1;
      } else {
        F0 += Fa0; F1 += Fa1;
        F0 += Fr0; F1 += Fr1;
      }
    }
    speed = py_min(norm2(F0, F1) + p.sf_bias_lon, speed_desired * p.sf_max_speed_factor);
    heading = atan2(F1, F0) + p.sf_bias_lat;
    force[0] = F0;
    force[1] = F1;
  } else {  // agent.py:65-68
    speed = 0;
    heading = 0;
    force[0] = 0.0;
    force[1] = 0.0;
  }
  // PedestrianController._step, pedestrian/controller.py:38-46 (uses state.dt)
  const double sp = np_clip(speed, -p.ped_max_speed, p.ped_max_speed);
  const double dt = t - prev_t;
  speed_io = sp;
  double sh2, ch2;
  sincos(heading, &sh2, &ch2);
#pragma unroll
  for (int f = 0; f < 6; ++f) out[f] = pose[f];
  out[0] = pose[0] + sp * dt * ch2;
  out[1] = pose[1] + sp * dt * sh2;
  out[3] = heading;
}

// per-thread view of one entity's mutable state
struct Ent {
  double pose[6], vel[6], dist, speed;
  double force[2];
  double sd[2], ratio[2];
  int cur_own, goal;
  uint8_t present, rss_state, rss_last, collided;
};

SG_DEV void load_ent(const SgState& st, int64_t i, int64_t nm, Ent& e) {
#pragma unroll
  for (int f = 0; f < 6; ++f) { e.pose[f] = st.pose[f * nm + i]; e.vel[f] = st.vel[f * nm + i]; }
  e.dist = st.dist[i];
  e.speed = st.speed[i];
  e.force[0] = st.force[i]; e.force[1] = st.force[nm + i];
  e.sd[0] = st.safe_dist[i]; e.sd[1] = st.safe_dist[nm + i];
  e.ratio[0] = st.safe_ratio[i]; e.ratio[1] = st.safe_ratio[nm + i];
  e.cur_own = st.cur_own[i];
  e.goal = st.goal_idx[i];
  e.present = st.present[i];
  e.rss_state = st.rss_state[i];
  e.rss_last = st.rss_last[i];
  e.collided = st.collided[i];
}
SG_DEV void store_ent(const SgState& st, int64_t i, int64_t nm, const Ent& e) {
#pragma unroll
  for (int f = 0; f < 6; ++f) { st.pose[f * nm + i] = e.pose[f]; st.vel[f * nm + i] = e.vel[f]; }
  st.dist[i] = e.dist;
  st.speed[i] = e.speed;
  st.force[i] = e.force[0]; st.force[nm + i] = e.force[1];
  st.safe_dist[i] = e.sd[0]; st.safe_dist[nm + i] = e.sd[1];
  st.safe_ratio[i] = e.ratio[0]; st.safe_ratio[nm + i] = e.ratio[1];
  st.cur_own[i] = e.cur_own;
  st.goal_idx[i] = e.goal;
  st.present[i] = e.present;
  st.rss_state[i] = e.rss_state;
  st.rss_last[i] = e.rss_last;
  st.collided[i] = e.collided;
}

// stage this entity's box for the group: fp64 corners, ring orientation, fp32 AABB
SG_DEV void publish_box(const TickCtx& c, const Ent& e, const double bw, const double bl,
                        const double bcx, const double bcy, double ox, double oy, double my[8],
                        int& my_or) {
  if (e.present) {
    box_points(e.pose[0], e.pose[1], e.pose[3], bw, bl, bcx, bcy, my);
#pragma unroll
    for (int f = 0; f < 8; ++f) c.corners[f * c.G + c.s] = my[f];
    my_or = quad_orientation(my);
    c.orient[c.s] = (int8_t)my_or;
    c.aabb[c.s] = make_aabb(my, ox, oy);
  } else {
    c.aabb[c.s] = make_float4(INFINITY, INFINITY, -INFINITY, -INFINITY);
  }
}

// RSSDistances.__call__ for one hazard entity (reference metrics/rss/callback.py:57-122)
SG_DEV void rss_entity(const SgScene& sc, const SgParams& p, const TickCtx& c, Ent& e,
                       const double my[8], double bw, double bl, int ego_slot, double t, int parity) {
  e.rss_last = SG_RSS_NONE;
  if (t == 0.0) return;  // :72
  if (c.s == ego_slot || !e.present) return;
  if (!(c.flags[c.G + 0] & 1)) return;  // ego absent (reference raises KeyError)
  const double ex = c.egobuf[0], ey = c.egobuf[1], ehd = c.egobuf[2];
  const double evx = c.egobuf[3], evy = c.egobuf[4];
  double es, ec;
  sincos(ehd, &es, &ec);
  const double eh[2] = {ec, es};
  double einv[2];
  inverse_direction(eh, einv);
  const double epos[2] = {ex, ey};
  double ecorn[8];
#pragma unroll
  for (int f = 0; f < 8; ++f) ecorn[f] = c.corners[f * c.G + ego_slot];
  const int64_t ei = (int64_t)c.n * c.M + ego_slot;
  RssEnt ego, haz;
  rss_entity_params(ex, ey, ehd, evx, evy, ecorn, sc.box[ei], sc.box[c.nm + ei], eh, einv, epos, ego);
  rss_entity_params(e.pose[0], e.pose[1], e.pose[3], e.vel[0], e.vel[1], my, bw, bl, eh, einv, epos, haz);
  double sd[2];
  sd[1] = fabs(safe_longitudinal_distance(p, ego, haz));  // :101-103
  sd[0] = fabs(safe_lateral_distance(p, ego, haz));
  e.sd[0] = sd[0];
  e.sd[1] = sd[1];
  safe_ratios(ego, haz, e.ratio);
  e.rss_last = (uint8_t)unsafe_distance(ego, haz, e.rss_state, sd);
  const int found = (e.rss_state >> 2) & 3;  // RSS metric latch, rss.py:71-103
  if (found) atomicOr(&c.acc[parity * 4 + 3], found == 2 ? 1 : 2);
}

// exact narrow phase for one AABB-surviving pair; both quads are read from the group's
// staged corners so the caller's registers stay free (out of line: it is the rare path)
__device__ __noinline__ bool pair_collides(const double* corners, const int8_t* orient, int G,
                                           int a, int b) {
  double qa[8], qb[8];
  bool same = true;
#pragma unroll
  for (int f = 0; f < 8; ++f) {
    qa[f] = corners[f * G + a];
    qb[f] = corners[f * G + b];
    same = same && (qa[f] == qb[f]);
  }
  if (same) return false;  // `g != g_prime`, reference utils.py:58
  return quads_intersect(qa, orient[a], qb, orient[b]);
}

// state.collisions() row of this entity (reference state/utils.py:10-49, utils.py:28-62)
SG_DEV void collision_sweep(const SgParams& p, const SgState& st, const TickCtx& c, bool present,
                            uint8_t& collided, int ego_slot, int first_slot, int parity) {
  const bool matrix = (p.features & SG_FEAT_COLL_MATRIX) != 0;
  uint32_t* row = matrix ? st.coll_mask + ((int64_t)c.n * c.M + c.s) * c.W : nullptr;
  if (!present) {
    if (matrix) for (int w = 0; w < c.W; ++w) row[w] = 0;
    return;
  }
  const float4 mb = c.aabb[c.s];
  int npairs = 0, first_j = -1;
  bool any = false;
  for (int w = 0; w < c.W; ++w) {
    uint32_t word = 0;
    const int j0 = w * 32, j1 = min(j0 + 32, c.M);
#pragma unroll 4
    for (int j = j0; j < j1; ++j) {
      const float4 ob = c.aabb[j];
      // closed-interval overlap of conservative bounds (STRtree's envelope filter is closed too)
      if (mb.x <= ob.z && ob.x <= mb.z && mb.y <= ob.w && ob.y <= mb.w && j != c.s) {
        if (pair_collides(c.corners, c.orient, c.G, c.s, j)) word |= 1u << (j - j0);
      }
    }
    if (matrix) row[w] = word;
    if (word) {
      any = true;
      if (c.s == ego_slot) c.ego_now[w] = word;
      // pairs (s, j) with j > s are counted by the lower slot
      uint32_t gt = word;
      if (c.s >= j0 + 31) gt = 0;
      else if (c.s >= j0) gt &= ~((2u << (c.s - j0)) - 1u);
      if (gt) {
        npairs += __popc(gt);
        if (first_j < 0) first_j = j0 + __ffs(gt) - 1;
      }
    }
  }
  if (any) {
    collided = 1;
    if (c.s == first_slot) c.acc[parity * 4 + 2] = 1;
  }
  if (npairs) {
    atomicAdd(&c.acc[parity * 4 + 0], npairs);
    atomicMin(&c.acc[parity * 4 + 1], (c.s << 16) | first_j);
  }
}

template <bool PED, bool RSS, int MAXT>
__global__ void __launch_bounds__(MAXT)
sg_rollout_kernel(SgScene sc, SgParams p, SgState st, SgInputs in, int n_ticks, GroupLayout L,
                  int reset) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int G = L.G, M = sc.n_slots, W = L.W;
  const int gpb = blockDim.x / G;  // scenario groups per CTA
  const int gl = threadIdx.x / G;
  const int s = threadIdx.x - gl * G;
  const int n = blockIdx.x * gpb + gl;
  if (gl >= gpb || n >= sc.n_scenarios) return;
  unsigned mask = 0xffffffffu;
  if (G < 32) mask = ((G == 32 ? 0u : (1u << G)) - 1u) << ((threadIdx.x & 31) / G * G);
  const int bar_id = 1 + gl;

  unsigned char* base = smem + (size_t)gl * L.bytes;
  TickCtx c;
  c.n = n; c.s = s; c.M = M; c.G = G; c.W = W;
  c.nm = (int64_t)sc.n_scenarios * M;
  c.i = (int64_t)n * M + s;
  c.corners = (double*)base;
  c.pedbuf = (double*)(base + L.off_ped);
  c.egobuf = (double*)(base + L.off_ego);
  c.aabb = (float4*)(base + L.off_aabb);
  c.ego_now = (uint32_t*)(base + L.off_hits);
  c.ego_last = c.ego_now + W;
  c.acc = (int*)(base + L.off_acc);
  c.flags = (uint8_t*)(base + L.off_flags);
  c.orient = (int8_t*)(base + L.off_orient);

  const bool live = s < M;  // padding threads only take part in barriers
  const int64_t i = c.i, nm = c.nm;
  const int kind = live ? sc.kind[i] : SG_KIND_EMPTY;
  const int etype = live ? sc.etype[i] : 0;
  const int ego_slot = sc.ego_slot[n], first_slot = sc.first_slot[n];
  const bool need_coll = (p.features & SG_FEAT_COLLISIONS) ||
                         (p.terminal & (SG_TERM_COLLISION | SG_TERM_EGO_COLLISION));
  const bool feat_rss = RSS && (p.features & SG_FEAT_RSS);

  double bw = 1, bl = 1, bcx = 0, bcy = 0;
  const double* rows = nullptr;
  int K = 0;
  if (kind != SG_KIND_EMPTY) {
    bw = sc.box[i]; bl = sc.box[nm + i]; bcx = sc.box[2 * nm + i]; bcy = sc.box[3 * nm + i];
    const int64_t r0 = sc.traj_off[i];
    K = (int)(sc.traj_off[i + 1] - r0);
    rows = sc.traj_rows + r0 * 7;
  }
  const double traj_min_t = K ? __ldg(rows) : 0.0, traj_max_t = K ? __ldg(rows + (int64_t)(K - 1) * 7) : 0.0;
  // scenario origin for the fp32 bounds: the ego's start position
  double ox, oy;
  {
    const int64_t er = sc.traj_off[(int64_t)n * M + ego_slot];
    ox = __ldg(sc.traj_rows + er * 7 + 1);
    oy = __ldg(sc.traj_rows + er * 7 + 2);
  }
  const double sight_cos = PED ? cos(p.sf_sight_angle / 2 * M_PI / 180) : 0.0;

  Ent e;
  double t, prev_t;
  int tick, cur_union;
  bool done;
  double avg = 0, avg_t = 0, mx = 0, egod = 0;  // ego metrics (ego thread only)
  int first_tick = -1, fp0 = -1, fp1 = -1;      // leader only
  long long pair_ticks = 0;
  int rss_flags = 0;

  if (reset) {
    // State.reset(t0), reference state/state.py:106-143 (+ controller/metric resets)
    t = sc.t0[n];
    memset(&e, 0, sizeof(e));
    e.cur_own = 1;
    e.rss_last = SG_RSS_NONE;
    e.ratio[0] = INFINITY; e.ratio[1] = INFINITY;  // callback.py:53-55
    if (kind != SG_KIND_EMPTY) {
      const int mode = (K == 1) ? EXT_TRUE : (p.persist ? EXT_CLAMP : EXT_NONE);  // :123-129
      int cur = 0;
      e.present = position_at_t(rows, K, t, mode, cur, e.pose);
      if (e.present) velocity_at_t(rows, K, t, e.vel);  // :132
      else {
#pragma unroll
        for (int f = 0; f < 6; ++f) e.pose[f] = 0.0;
      }
      if (kind == SG_KIND_VEHICLE) e.speed = norm2(e.vel[0], e.vel[1]);  // controller.py:100-103
    }
    prev_t = t - 0.1;  // :135
    tick = 0;
    done = false;
    cur_union = 1;
    if (s < W) c.ego_last[s] = 0;
  } else {
    t = st.t[n];
    prev_t = st.prev_t[n];
    tick = st.tick[n];
    done = st.done[n] != 0;
    cur_union = st.cur_union[n];
    if (live) load_ent(st, i, nm, e); else memset(&e, 0, sizeof(e));
    if (s == ego_slot) {
      avg = st.ego_avg_speed[n]; avg_t = st.ego_avg_t[n]; mx = st.ego_max_speed[n]; egod = st.ego_dist[n];
    }
    if (s == 0) {
      first_tick = st.first_coll_tick[n]; fp0 = st.first_coll_pair[2 * n]; fp1 = st.first_coll_pair[2 * n + 1];
      pair_ticks = st.n_pair_ticks[n];
      rss_flags = st.rss_flags[n];
    }
    if (s < W) c.ego_last[s] = st.ego_hits[(int64_t)n * W + s];
  }
  if (s < W) c.ego_now[s] = 0;
  if (s == 0) {
    for (int q = 0; q < 8; ++q) c.acc[q] = 0;
    c.acc[1] = 0x7fffffff; c.acc[5] = 0x7fffffff;
  }
  // stage the "old" state the pedestrians' sensors read in the first tick
  if (live) {
    c.flags[s] = (uint8_t)((e.present ? 1 : 0) | (etype << 1));
    if (PED) {
      c.pedbuf[s] = e.pose[0]; c.pedbuf[G + s] = e.pose[1];
      c.pedbuf[2 * G + s] = e.vel[0]; c.pedbuf[3 * G + s] = e.vel[1];
    }
  }
  double my[8];
  int my_or = 0;
  int parity = 0;

  if (reset) {
    // update_callbacks() at reset (state.py:137-139) and Metric.reset (metrics/trajectory.py:13-18)
    if (feat_rss) {
      if (live) publish_box(c, e, bw, bl, bcx, bcy, ox, oy, my, my_or);
      if (s == ego_slot) {
        c.egobuf[0] = e.pose[0]; c.egobuf[1] = e.pose[1]; c.egobuf[2] = e.pose[3];
        c.egobuf[3] = e.vel[0]; c.egobuf[4] = e.vel[1];
        c.flags[G] = e.present;
      }
      group_sync(G, bar_id, mask);
      if (live) rss_entity(sc, p, c, e, my, bw, bl, ego_slot, t, 0);
      group_sync(G, bar_id, mask);
      if (s == 0) { rss_flags |= c.acc[3]; c.acc[3] = 0; }
    }
    if (s == ego_slot) {
      const double sp = norm3(e.vel[0], e.vel[1], e.vel[2]);
      avg = sp; avg_t = 0.0; mx = sp; egod = 0.0;
    }
    if (st.trace_cap > 0 && live) {
      st.trace_present[i] = e.present;
#pragma unroll
      for (int f = 0; f < 6; ++f) st.trace_pose[f * nm + i] = e.pose[f];
      if (s == 0) st.trace_t[n] = t;
    }
    n_ticks = 0;
  }
  group_sync(G, bar_id, mask);

  int limit = n_ticks < 0 ? p.max_ticks : n_ticks;
  if (in.actions && limit > in.n_action_ticks) limit = in.n_action_ticks;

  for (int k = 0; k < limit && !done; ++k) {
    // ---------------- phase A: agents / batch replay produce the new poses ----------------
    const double next_t = t + p.timestep;  // scenario_gym.py:229
    double np_[6];
    bool newpres = false;
    double newspeed = e.speed;
    if (kind >= SG_KIND_AGENT_REPLAY) {  // scenario_gym.py:233-244
      if (e.present) {
        if (kind == SG_KIND_AGENT_REPLAY) {  // agent.py:125-128, extrapolate=(False, False)
          position_at_t(rows, K, next_t, EXT_CLAMP, e.cur_own, np_);
          newpres = true;
        } else if (kind == SG_KIND_VEHICLE) {  // VehicleController._step, controller.py:105-140
          double accel = in.actions[((int64_t)k * 2 + 0) * nm + i];
          double steer = in.actions[((int64_t)k * 2 + 1) * nm + i];
          accel = np_clip(accel, -p.veh_max_accel, p.veh_max_accel);
          steer = np_clip(steer, -p.veh_max_steer, p.veh_max_steer);
          const double dt = next_t - t;
          double sh, ch;
          sincos(e.pose[3], &sh, &ch);
          const double dx = e.speed * ch, dy = e.speed * sh, dh = e.speed * tan(steer) / bl;
#pragma unroll
          for (int f = 0; f < 6; ++f) np_[f] = e.pose[f];
          np_[0] += dx * dt;
          np_[1] += dy * dt;
          np_[3] += dh * dt;
          double ns = e.speed + accel * dt;
          if (!p.veh_allow_reverse) ns = fmax(0.0, ns);
          if (!isnan(p.veh_max_speed)) ns = fmin(p.veh_max_speed, ns);
          newspeed = ns;
          newpres = true;
        } else if (kind == SG_KIND_PEDESTRIAN) {
          pedestrian_step<PED>(sc, p, c, e.pose, e.vel, t, prev_t, next_t, sight_cos, e.goal,
                               e.force, newspeed, np_);
          newpres = true;
        } else {  // SG_KIND_HOST
          if (in.host_present && in.host_present[i]) {
#pragma unroll
            for (int f = 0; f < 6; ++f) np_[f] = in.host_pose[f * nm + i];
            newpres = true;
          } else if (p.persist) {  // scenario_gym.py:238-239
#pragma unroll
            for (int f = 0; f < 6; ++f) np_[f] = e.pose[f];
            newpres = true;
          }
        }
      } else if (traj_min_t >= t) {  // :240-244 agent initialised at its start position
        position_at_t(rows, K, next_t, EXT_CLAMP, e.cur_own, np_);
        newpres = true;
      }
    } else if (kind == SG_KIND_REPLAY) {  // BatchReplayEntity.step, entity/batch.py:34-53
      // every thread advances the scenario's shared cursor identically
      const int64_t u0 = sc.union_off[n];
      const int UK = (int)(sc.union_off[n + 1] - u0);
      const double* ts = sc.union_t + u0;
      if (p.persist || K == 1 || (next_t >= traj_min_t && next_t <= traj_max_t)) {
        const double* X = sc.union_x + u0 * 6 * M;
        if (next_t < __ldg(ts)) {  // fill_value = (X[0], X[-1]), entity/batch.py:120-127
#pragma unroll
          for (int f = 0; f < 6; ++f) np_[f] = __ldg(X + f * M + s);
        } else if (next_t > __ldg(ts + UK - 1)) {
#pragma unroll
          for (int f = 0; f < 6; ++f) np_[f] = __ldg(X + ((int64_t)(UK - 1) * 6 + f) * M + s);
        } else {
          cur_union = search_left_cursor(ts, 1, UK, next_t, cur_union);
          const double x_lo = __ldg(ts + cur_union - 1), x_hi = __ldg(ts + cur_union);
          const double w1 = (next_t - x_lo) / (x_hi - x_lo), w0 = (x_hi - next_t) / (x_hi - x_lo);
          const double* lo = X + (int64_t)(cur_union - 1) * 6 * M + s;
#pragma unroll
          for (int f = 0; f < 6; ++f) np_[f] = w1 * __ldg(lo + (6 + f) * M) + w0 * __ldg(lo + f * M);
        }
        newpres = true;
      }
    }
    // ---------------- State.step: update_poses / update_statistics (state.py:203-239) -------
    prev_t = t;
    t = next_t;
    tick += 1;
    const double dt = t - prev_t;
    if (newpres) {
      double prev[6];
      if (e.present) {
#pragma unroll
        for (int f = 0; f < 6; ++f) prev[f] = e.pose[f];
      } else {  // :219-222 newcomer: previous pose extrapolated from its trajectory
        int cur = 0;
        position_at_t(rows, K, prev_t, EXT_TRUE, cur, prev);
      }
      double d[6];
#pragma unroll
      for (int f = 0; f < 6; ++f) {
        d[f] = np_[f] - prev[f];
        e.vel[f] = d[f] / dt;
        e.pose[f] = np_[f];
      }
      e.dist += norm3(d[0], d[1], d[2]);
      e.speed = newspeed;
    }
    e.present = newpres;
    if (st.trace_cap > 0 && tick < st.trace_cap && live) {
      st.trace_present[(int64_t)tick * nm + i] = e.present;
#pragma unroll
      for (int f = 0; f < 6; ++f) st.trace_pose[((int64_t)tick * 6 + f) * nm + i] = e.pose[f];
      if (s == 0) st.trace_t[(int64_t)tick * sc.n_scenarios + n] = t;
    }
    if ((need_coll || feat_rss) && live) publish_box(c, e, bw, bl, bcx, bcy, ox, oy, my, my_or);
    if (s == ego_slot) {
      c.egobuf[0] = e.pose[0]; c.egobuf[1] = e.pose[1]; c.egobuf[2] = e.pose[3];
      c.egobuf[3] = e.vel[0]; c.egobuf[4] = e.vel[1];
      c.flags[G] = e.present;
    }
    group_sync(G, bar_id, mask);
    // ---------------- phase B: callbacks (RSS) + collisions --------------------------------
    if (live) {
      c.flags[s] = (uint8_t)((e.present ? 1 : 0) | (etype << 1));
      if (PED) {
        c.pedbuf[s] = e.pose[0]; c.pedbuf[G + s] = e.pose[1];
        c.pedbuf[2 * G + s] = e.vel[0]; c.pedbuf[3 * G + s] = e.vel[1];
      }
      if (feat_rss) rss_entity(sc, p, c, e, my, bw, bl, ego_slot, t, parity);
      if (need_coll) collision_sweep(p, st, c, e.present != 0, e.collided, ego_slot, first_slot, parity);
    }
    group_sync(G, bar_id, mask);
    // ---------------- phase C: terminal check + metrics ------------------------------------
    const int npairs = c.acc[parity * 4 + 0];
    const int first_pair = c.acc[parity * 4 + 1];
    const int first_hit = c.acc[parity * 4 + 2];
    const int rss_now = c.acc[parity * 4 + 3];
    bool dn = false;  // state.py:268-270, 397-408
    if ((p.terminal & SG_TERM_MAX_LENGTH) && (t + dt > sc.length[n])) dn = true;
    if ((p.terminal & SG_TERM_COLLISION) && npairs > 0) dn = true;
    if ((p.terminal & SG_TERM_EGO_COLLISION) && first_hit) dn = true;
    done = dn;
    if (s == 0) {
      pair_ticks += npairs;
      if (npairs > 0 && first_tick < 0) { first_tick = tick; fp0 = first_pair >> 16; fp1 = first_pair & 0xffff; }
      rss_flags |= rss_now;
      const int q = (parity ^ 1) * 4;  // reset the other parity for the next tick
      c.acc[q] = 0; c.acc[q + 1] = 0x7fffffff; c.acc[q + 2] = 0; c.acc[q + 3] = 0;
    }
    if (s < W) {  // CollisionMetric._step, metrics/collision.py:70-75
      const uint32_t now = c.ego_now[s];
      if (p.features & SG_FEAT_COLLISIONS) {
        uint32_t fresh = now & ~c.ego_last[s];
        while (fresh) {
          const int b = __ffs(fresh) - 1;
          fresh &= fresh - 1;
          const int slot = atomicAdd(st.event_count, 1);
          if (slot < st.event_cap) {
            SgEvent ev;
            ev.scenario = n; ev.tick = tick; ev.slot = s * 32 + b; ev._pad = 0; ev.t = t;
            st.events[slot] = ev;
          }
        }
        c.ego_last[s] = now;
      }
      c.ego_now[s] = 0;
    }
    if (s == ego_slot && (p.features & SG_FEAT_EGO_METRICS)) {  // metrics/trajectory.py:20-24,39-42,58-60
      const double sp = norm3(e.vel[0], e.vel[1], e.vel[2]);
      const double w = avg_t / t;
      avg += (1.0 - w) * (sp - avg);
      avg_t = t;
      mx = fmax(sp, mx);
      egod = e.dist;
    }
    parity ^= 1;
  }

  // ---------------- write the State rows back ----------------------------------------------
  if (live) store_ent(st, i, nm, e);
  if (s == ego_slot) {
    st.ego_avg_speed[n] = avg; st.ego_avg_t[n] = avg_t; st.ego_max_speed[n] = mx; st.ego_dist[n] = egod;
  }
  if (s == 0) {
    st.t[n] = t; st.prev_t[n] = prev_t; st.tick[n] = tick; st.done[n] = done; st.cur_union[n] = cur_union;
    st.first_coll_tick[n] = first_tick; st.first_coll_pair[2 * n] = fp0; st.first_coll_pair[2 * n + 1] = fp1;
    st.n_pair_ticks[n] = pair_ticks;
    st.rss_flags[n] = (uint8_t)rss_flags;
  }
  if (s < W) st.ego_hits[(int64_t)n * W + s] = c.ego_last[s];
}

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
  out[i] = !same && quads_intersect(qa, quad_orientation(qa), qb, quad_orientation(qb));
}

// ---------------------------------------------------------------------------------
static bool g_ngon_ready[64] = {false};

static int ensure_constants(int device) {
  if (device < 0 || device >= 64) return set_msg("device index out of range");
  if (g_ngon_ready[device]) return 0;
  double cs[64][2];
  const double inc = (2.0 * M_PI) / 64;
  for (int k = 0; k < 64; ++k) {  // GEOS addDirectedFillet: clockwise from angle 0 (libm on the host)
    const double ang = 0.0 + -1.0 * k * inc;
    cs[k][0] = cos(ang);
    cs[k][1] = sin(ang);
  }
  cudaError_t err = cudaMemcpyToSymbol(c_ngon, cs, sizeof(cs));
  if (err != cudaSuccess) return set_err("cudaMemcpyToSymbol", err);
  g_ngon_ready[device] = true;
  return 0;
}

static bool scene_has_kind(const SgParams* p) { (void)p; return true; }

static int launch(const SgScene* sc, const SgParams* p, SgState* st, const SgInputs* in, int n_ticks,
                  int device, void* stream, int reset) {
  if (!sc || !p || !st) return set_msg("null argument");
  if (sc->n_slots < 1 || sc->n_slots > 1024) return set_msg("n_slots must be in 1..1024");
  if (sc->n_scenarios < 1) return set_msg("n_scenarios must be >= 1");
  cudaError_t err = cudaSetDevice(device);
  if (err != cudaSuccess) return set_err("cudaSetDevice", err);
  int rc = ensure_constants(device);
  if (rc) return rc;
  const bool ped = sc->route_off != nullptr && sc->n_route_pts > 0;
  const bool rss = (p->features & SG_FEAT_RSS) != 0;
  GroupLayout L = make_layout(sc->n_slots, ped);
  const int threads = L.G <= SG_THREADS ? SG_THREADS : L.G;
  const int gpb = threads / L.G;
  const int blocks = (sc->n_scenarios + gpb - 1) / gpb;
  const size_t smem = (size_t)gpb * L.bytes;
  SgInputs none;
  memset(&none, 0, sizeof(none));
  const SgInputs inp = in ? *in : none;
  void (*kern)(SgScene, SgParams, SgState, SgInputs, int, GroupLayout, int);
  const bool big = threads > SG_THREADS;
  if (ped && rss) kern = big ? sg_rollout_kernel<true, true, 1024> : sg_rollout_kernel<true, true, SG_THREADS>;
  else if (ped) kern = big ? sg_rollout_kernel<true, false, 1024> : sg_rollout_kernel<true, false, SG_THREADS>;
  else if (rss) kern = big ? sg_rollout_kernel<false, true, 1024> : sg_rollout_kernel<false, true, SG_THREADS>;
  else kern = big ? sg_rollout_kernel<false, false, 1024> : sg_rollout_kernel<false, false, SG_THREADS>;
  if (smem > 48 * 1024) {
    err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return set_err("cudaFuncSetAttribute", err);
  }
  kern<<<blocks, threads, smem, (cudaStream_t)stream>>>(*sc, *p, *st, inp, n_ticks, L, reset);
  err = cudaGetLastError();
  if (err != cudaSuccess) return set_err("sg_rollout_kernel launch", err);
  (void)scene_has_kind;
  return 0;
}

extern "C" {

int sg_abi_version(void) { return SG_ABI_VERSION; }

int64_t sg_sizeof(int which) {
  switch (which) {
    case 0: return sizeof(SgParams);
    case 1: return sizeof(SgScene);
    case 2: return sizeof(SgState);
    case 3: return sizeof(SgInputs);
    case 4: return sizeof(SgEvent);
  }
  return -1;
}

const char* sg_last_error(void) { return g_err; }

void sg_default_params(SgParams* p) {
  memset(p, 0, sizeof(*p));
  p->timestep = 1.0 / 30.0;
  p->terminal = SG_TERM_MAX_LENGTH;
  p->features = SG_FEAT_COLLISIONS | SG_FEAT_EGO_METRICS;
  p->max_ticks = 1 << 20;
  p->veh_max_steer = 0.7;
  p->veh_max_accel = 5.0;
  p->veh_max_speed = NAN;
  p->ped_max_speed = 5.0;
  p->ped_distance_threshold = 1.0;
  p->sf_max_speed_factor = 1.3;
  p->sf_sight_weight = 0.5;
  p->sf_sight_weight_use = 1;
  p->sf_sight_angle = 200.0;
  p->sf_relaxation_time = 1.5;
  p->sf_ped_repulse_V = 1.0;
  p->sf_ped_repulse_sigma = 1.0;
  p->rss_response_time = 0.6;
  p->rss_min_long_accel = 1.2 * 9.81;
  p->rss_max_long_accel = 1.2 * 9.81;
  p->rss_min_safe_clearance = 0.1;
}

int sg_reset(const SgScene* scene, const SgParams* params, SgState* state, int device, void* stream) {
  if (state && state->event_count) {
    cudaError_t err = cudaSetDevice(device);
    if (err != cudaSuccess) return set_err("cudaSetDevice", err);
    err = cudaMemsetAsync(state->event_count, 0, sizeof(int32_t), (cudaStream_t)stream);
    if (err != cudaSuccess) return set_err("cudaMemsetAsync", err);
  }
  return launch(scene, params, state, nullptr, 0, device, stream, 1);
}

int sg_rollout(const SgScene* scene, const SgParams* params, SgState* state, const SgInputs* inputs,
               int n_ticks, int device, void* stream) {
  return launch(scene, params, state, inputs, n_ticks, device, stream, 0);
}

int sg_test_box_pairs(const double* pose_a, const double* box_a, const double* pose_b,
                      const double* box_b, uint8_t* out, int64_t n, int device, void* stream) {
  cudaError_t err = cudaSetDevice(device);
  if (err != cudaSuccess) return set_err("cudaSetDevice", err);
  if (n <= 0) return 0;
  sg_box_pairs_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      pose_a, box_a, pose_b, box_b, out, n);
  err = cudaGetLastError();
  if (err != cudaSuccess) return set_err("sg_box_pairs_kernel launch", err);
  return 0;
}

// ---- host-buffer path ------------------------------------------------------------------
struct CopyItem { const void* src; void* dst; size_t bytes; };

static int scene_copy_list(const SgScene* h, const SgScene* d, CopyItem* items) {
  const int64_t N = h->n_scenarios, M = h->n_slots, NM = N * M;
  int k = 0;
#define ITEM(field, bytes_) items[k++] = CopyItem{h->field, (void*)d->field, (size_t)(bytes_)}
  ITEM(kind, NM);
  ITEM(etype, NM);
  ITEM(box, 4 * NM * 8);
  ITEM(traj_off, (NM + 1) * 8);
  ITEM(traj_rows, h->n_traj_rows * 7 * 8);
  ITEM(union_off, (N + 1) * 8);
  ITEM(union_t, h->n_union_rows * 8);
  ITEM(union_x, h->n_union_rows * 6 * M * 8);
  ITEM(t0, N * 8);
  ITEM(length, N * 8);
  ITEM(ego_slot, N * 4);
  ITEM(first_slot, N * 4);
  ITEM(ped_speed_desired, NM * 8);
  ITEM(route_off, (NM + 1) * 8);
  ITEM(route_xy, h->n_route_pts * 2 * 8);
#undef ITEM
  return k;
}

int64_t sg_host_h2d_bytes(const SgScene* h, const SgInputs* in, int copy_static) {
  CopyItem items[16];
  SgScene dummy = *h;
  int k = scene_copy_list(h, &dummy, items);
  int64_t total = 0;
  if (copy_static)
    for (int q = 0; q < k; ++q)
      if (items[q].src) total += (int64_t)items[q].bytes;
  if (in && in->actions) total += (int64_t)in->n_action_ticks * 2 * h->n_scenarios * h->n_slots * 8;
  return total;
}

int64_t sg_host_d2h_bytes(const SgScene* h) {
  const int64_t N = h->n_scenarios;
  return N * (8 * 3 + 4 + 8 + 8 + 1 + 4 + 8) + 4;
}

int sg_rollout_host(const SgScene* hs, const SgScene* ds, const SgParams* params, SgState* dst,
                    const SgInputs* hin, const SgInputs* din, SgHostResults* res, int copy_static,
                    int device, void* stream) {
  if (!hs || !ds || !params || !dst || !res) return set_msg("null argument");
  cudaError_t err = cudaSetDevice(device);
  if (err != cudaSuccess) return set_err("cudaSetDevice", err);
  cudaStream_t s = (cudaStream_t)stream;
  if (copy_static) {
    CopyItem items[16];
    const int k = scene_copy_list(hs, ds, items);
    for (int q = 0; q < k; ++q) {
      if (!items[q].src || !items[q].bytes) continue;
      if (!items[q].dst) return set_msg("device scene mirror is missing an array");
      err = cudaMemcpyAsync(items[q].dst, items[q].src, items[q].bytes, cudaMemcpyHostToDevice, s);
      if (err != cudaSuccess) return set_err("cudaMemcpyAsync H2D scene", err);
    }
  }
  int rc = sg_reset(ds, params, dst, device, stream);
  if (rc) return rc;
  const int64_t NM = (int64_t)hs->n_scenarios * hs->n_slots;
  if (hin && hin->actions) {
    if (!din || !din->actions) return set_msg("device action buffer missing");
    // stream the action table in chunks of ticks: the copy of chunk c+1 overlaps the kernel
    // of chunk c (copies on a second stream, ordered with events)
    const int T = hin->n_action_ticks;
    const int chunk = T < 16 ? T : 16;
    static thread_local cudaStream_t copy_stream = nullptr;
    static thread_local cudaEvent_t ev_copy[2] = {nullptr, nullptr}, ev_done = nullptr;
    if (!copy_stream) {
      cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking);
      cudaEventCreateWithFlags(&ev_copy[0], cudaEventDisableTiming);
      cudaEventCreateWithFlags(&ev_copy[1], cudaEventDisableTiming);
      cudaEventCreateWithFlags(&ev_done, cudaEventDisableTiming);
    }
    cudaEventRecord(ev_done, s);
    cudaStreamWaitEvent(copy_stream, ev_done, 0);
    int c = 0;
    for (int k0 = 0; k0 < T; k0 += chunk, ++c) {
      const int kt = (T - k0) < chunk ? (T - k0) : chunk;
      const size_t off = (size_t)k0 * 2 * NM;
      err = cudaMemcpyAsync((double*)din->actions + off, hin->actions + off,
                            (size_t)kt * 2 * NM * 8, cudaMemcpyHostToDevice, copy_stream);
      if (err != cudaSuccess) return set_err("cudaMemcpyAsync H2D actions", err);
      cudaEventRecord(ev_copy[c & 1], copy_stream);
      cudaStreamWaitEvent(s, ev_copy[c & 1], 0);
      SgInputs part = *din;
      part.actions = din->actions + off;
      part.n_action_ticks = kt;
      rc = sg_rollout(ds, params, dst, &part, kt, device, stream);
      if (rc) return rc;
    }
  } else {
    rc = sg_rollout(ds, params, dst, din, -1, device, stream);
    if (rc) return rc;
  }
  const int64_t N = hs->n_scenarios;
#define BACK(field, bytes_)                                                                   \
  if (res->field) {                                                                           \
    err = cudaMemcpyAsync(res->field, dst->field, (size_t)(bytes_), cudaMemcpyDeviceToHost, s); \
    if (err != cudaSuccess) return set_err("cudaMemcpyAsync D2H " #field, err);               \
  }
  BACK(ego_avg_speed, N * 8);
  BACK(ego_max_speed, N * 8);
  BACK(ego_dist, N * 8);
  BACK(first_coll_tick, N * 4);
  BACK(first_coll_pair, N * 8);
  BACK(n_pair_ticks, N * 8);
  BACK(rss_flags, N);
  BACK(tick, N * 4);
  BACK(t, N * 8);
  BACK(event_count, 4);
#undef BACK
  return 0;
}

}  // extern "C"
