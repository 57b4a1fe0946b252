/*
 * sg_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C, single-threaded, scalar restatement of the reference's per-tick rollout
 * path (driskai/scenario_gym v0.3.1).  It exists to CHECK the CUDA engine
 * (tests/, __graft_entry__.smoke) and to be timed as the CPU baseline
 * (bench.py cpu_baseline / --impl reference).  The product never links or calls it.
 *
 * It exports the same entry points as include/sg_b200.h with the prefix sgo_ and
 * HOST pointers.  Every function cites the reference file:line it follows (paths
 * relative to the reference root, package dir scenario_gym/).
 *
 * Pinning: oracle/gen_golden.py runs the unmodified Python reference in the
 * authoring container (through oracle/refshim) and commits golden vectors under
 * tests/golden/; tests/test_oracle_golden.py checks this file against them.
 * Shapely/GEOS arithmetic is absent from the reference tree (pyproject.toml:59,
 * "shapely>=2.0.0" unpinned): its published semantics are restated here
 * (closed-set intersects decided with exact orientation signs; Point.buffer = 64-gon;
 * vectorized.contains = strict interior; LineString.project) and are pinned only
 * through the reference's own known-answer tests -- see DESIGN.md "Oracle".
 *
 * Build: oracle/Makefile (gcc -O2 -ffp-contract=off: no FMA contraction, numpy has none).
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/sg_b200.h"

#define NM ((int64_t)sc->n_scenarios * sc->n_slots)

static __thread char g_err[256];
const char* sgo_last_error(void) { return g_err; }
int sgo_abi_version(void) { return SG_ABI_VERSION; }

int64_t sgo_sizeof(int which) {
  switch (which) {
    case 0: return sizeof(SgParams);
    case 1: return sizeof(SgScene);
    case 2: return sizeof(SgState);
    case 3: return sizeof(SgInputs);
    case 4: return sizeof(SgEvent);
    case 5: return sizeof(SgActionRng);
    case 6: return sizeof(SgHostResults);
  }
  return -1;
}

void sgo_default_params(SgParams* p) {
  memset(p, 0, sizeof(*p));
  p->timestep = 1.0 / 30.0; /* scenario_gym.py:31 */
  p->persist = 0;
  p->terminal = SG_TERM_MAX_LENGTH; /* scenario_gym.py:77-78 */
  p->features = SG_FEAT_COLLISIONS | SG_FEAT_EGO_METRICS;
  p->max_ticks = 1 << 20;
  p->veh_max_steer = 0.7; /* controller.py:67-70 */
  p->veh_max_accel = 5.0;
  p->veh_max_speed = NAN;
  p->veh_allow_reverse = 0;
  p->ped_max_speed = 5.0;         /* pedestrian/agent.py:24 */
  p->ped_head_rot_angle = 0.0;    /* pedestrian/agent.py:25 */
  p->ped_distance_threshold = 1.0;/* pedestrian/agent.py:26 */
  p->sf_max_speed_factor = 1.3;   /* pedestrian/behaviour.py:11 */
  p->sf_bias_lon = 0.0;           /* pedestrian/random_walk.py:16-17 */
  p->sf_bias_lat = 0.0;
  p->sf_sight_weight = 0.5;       /* pedestrian/social_force.py:19-30 */
  p->sf_sight_weight_use = 1;
  p->sf_sight_angle = 200.0;
  p->sf_relaxation_time = 1.5;
  p->sf_ped_repulse_V = 1.0;
  p->sf_ped_repulse_sigma = 1.0;
  p->sf_ped_attract_C = 0.0;
  p->rss_response_time = 0.6;     /* metrics/rss/callback.py:24-27 */
  p->rss_min_long_accel = 1.2 * 9.81;
  p->rss_max_long_accel = 1.2 * 9.81;
  p->rss_min_safe_clearance = 0.1;
  p->pid_steer_Kp = 0.03054; /* controller.py:154-161 */
  p->pid_steer_Kd = 1.5709;
  p->pid_accel_Kp = 0.3753;
  p->pid_accel_Kd = 1.8970;
  p->pid_accel_Ki = 0.0204;
  p->sf_boundary_repulse_U = 10.0; /* pedestrian/social_force.py:26-29 */
  p->sf_boundary_repulse_R = 0.2;
  p->sf_imp_boundary_repulse_U = 2.0;
  p->sf_imp_boundary_repulse_R = 0.1;
}

/* ------------------------------------------------------------------------- */
/* scipy.interpolate.interp1d(kind="linear")._call_linear, scipy 1.18.1
 * _interpolate.py:491-518 (third-party, called from trajectory.py:178-184 and
 * entity/batch.py:99-127):
 *   i = clip(searchsorted(x, t, 'left'), 1, K-1)
 *   y = ((t-x_lo)/(x_hi-x_lo))*y_hi + ((x_hi-t)/(x_hi-x_lo))*y_lo              */
static int64_t searchsorted_left(const double* x, int64_t stride, int64_t K, double t) {
  int64_t lo = 0, hi = K;
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if (x[mid * stride] < t) lo = mid + 1; else hi = mid;
  }
  return lo;
}

static void call_linear_rows(const double* rows, int64_t K, double t, double out[6]) {
  /* rows: [K][7] = t,x,y,z,h,p,r ; K >= 2 */
  int64_t i = searchsorted_left(rows, 7, K, t);
  if (i < 1) i = 1;
  if (i > K - 1) i = K - 1;
  const double* lo = rows + (i - 1) * 7;
  const double* hi = rows + i * 7;
  double w1 = (t - lo[0]) / (hi[0] - lo[0]);
  double w0 = (hi[0] - t) / (hi[0] - lo[0]);
  for (int f = 0; f < 6; ++f) out[f] = w1 * hi[1 + f] + w0 * lo[1 + f];
}

/* Trajectory.position_at_t for scalar t, trajectory.py:142-197.
 * mode: 0 = extrapolate False (None outside), 1 = (False, False) clamped,
 *       2 = extrapolate True.  Returns 0 when the result is None. */
enum { EXT_NONE = 0, EXT_CLAMP = 1, EXT_TRUE = 2 };

static int position_at_t(const double* rows, int64_t K, double t, int mode, double out[6]) {
  double min_t = rows[0], max_t = rows[(K - 1) * 7];
  if (mode == EXT_NONE && (t < min_t || t > max_t)) return 0;            /* :191-192 */
  if (mode != EXT_TRUE && t < min_t) { memcpy(out, rows + 1, 48); return 1; }            /* :193-194 */
  if (mode != EXT_TRUE && t > max_t) { memcpy(out, rows + (K - 1) * 7 + 1, 48); return 1; } /* :195-196 */
  if (K == 1) { /* :175-177: single control point duplicated with t + 1e-3 */
    double two[14];
    memcpy(two, rows, 56);
    memcpy(two + 7, rows, 56);
    two[7] += 1e-3;
    call_linear_rows(two, 2, t, out);
  } else {
    call_linear_rows(rows, K, t, out);
  }
  return 1;
}

/* Trajectory.velocity_at_t, trajectory.py:243-273 (eps = 1e-4) */
static void velocity_at_t(const double* rows, int64_t K, double t, double out[6]) {
  const double eps = 1e-4;
  double min_t = rows[0], max_t = rows[(K - 1) * 7];
  int inside = (min_t <= t) && (t <= max_t);
  double a[6], b[6];
  position_at_t(rows, K, t + eps / 2, EXT_TRUE, a);
  position_at_t(rows, K, t - eps / 2, EXT_TRUE, b);
  for (int f = 0; f < 6; ++f) out[f] = inside ? (a[f] - b[f]) / eps : 0.0;
}

/* BatchReplayEntity.fn(t): interp1d(ts, X.T, bounds_error=False, fill_value=(X[0], X[-1]))
 * entity/batch.py:120-128 + scipy _evaluate/_check_bounds (_interpolate.py:561-606) */
static void union_interp(const SgScene* sc, int n, int slot, double t, double out[6]) {
  int64_t r0 = sc->union_off[n], K = sc->union_off[n + 1] - r0;
  const double* ts = sc->union_t + r0;
  int64_t M = sc->n_slots;
  const double* X = sc->union_x + r0 * 6 * M;
  if (t < ts[0]) { for (int f = 0; f < 6; ++f) out[f] = X[f * M + slot]; return; }
  if (t > ts[K - 1]) { for (int f = 0; f < 6; ++f) out[f] = X[((K - 1) * 6 + f) * M + slot]; return; }
  int64_t i = searchsorted_left(ts, 1, K, t);
  if (i < 1) i = 1;
  if (i > K - 1) i = K - 1;
  double w1 = (t - ts[i - 1]) / (ts[i] - ts[i - 1]);
  double w0 = (ts[i] - t) / (ts[i] - ts[i - 1]);
  for (int f = 0; f < 6; ++f)
    out[f] = w1 * X[(i * 6 + f) * M + slot] + w0 * X[((i - 1) * 6 + f) * M + slot];
}

/* ------------------------------------------------------------------------- */
/* Exact orientation sign.  GEOS decides `intersects` with robust orientation
 * predicates (Shapely call sites utils.py:51-62, rss/callback.py:191-196,317-328);
 * here the sign of |b-a, c-a| is exact for the fp64 inputs: static filter
 * (Shewchuk's ccwerrboundA) then exact expansion arithmetic.                     */
static inline void two_sum(double a, double b, double* s, double* e) {
  double x = a + b, bv = x - a, av = x - bv;
  *s = x;
  *e = (a - av) + (b - bv);
}
static inline void two_prod(double a, double b, double* p, double* e) {
  *p = a * b;
  *e = fma(a, b, -*p);
}
static int grow(double* h, int n, double b) { /* Shewchuk grow-expansion */
  double q = b;
  for (int i = 0; i < n; ++i) { double s, e; two_sum(q, h[i], &s, &e); h[i] = e; q = s; }
  h[n] = q;
  return n + 1;
}
static int orient_exact(double ax, double ay, double bx, double by, double cx, double cy) {
  double d[4][2]; /* (ax-cx), (by-cy), (ay-cy), (bx-cx) as head+tail */
  two_sum(ax, -cx, &d[0][0], &d[0][1]);
  two_sum(by, -cy, &d[1][0], &d[1][1]);
  two_sum(ay, -cy, &d[2][0], &d[2][1]);
  two_sum(bx, -cx, &d[3][0], &d[3][1]);
  double h[40];
  int n = 0;
  for (int i = 0; i < 2; ++i)
    for (int j = 0; j < 2; ++j) {
      double p, e;
      two_prod(d[0][i], d[1][j], &p, &e);
      n = grow(h, n, e); n = grow(h, n, p);
      two_prod(d[2][i], d[3][j], &p, &e);
      n = grow(h, n, -e); n = grow(h, n, -p);
    }
  for (int i = n - 1; i >= 0; --i) {
    if (h[i] > 0) return 1;
    if (h[i] < 0) return -1;
  }
  return 0;
}
static int orient_sign(double ax, double ay, double bx, double by, double cx, double cy) {
  const double eps = 1.1102230246251565e-16;
  const double errbound = (3.0 + 16.0 * eps) * eps;
  double l = (ax - cx) * (by - cy), r = (ay - cy) * (bx - cx);
  double det = l - r;
  if (fabs(det) > errbound * (fabs(l) + fabs(r))) return det > 0 ? 1 : -1;
  return orient_exact(ax, ay, bx, by, cx, cy);
}

/* ring orientation of a convex quad: +1 ccw, -1 cw, 0 degenerate */
static int quad_orientation(const double q[8]) {
  int s = orient_sign(q[0], q[1], q[2], q[3], q[4], q[5]);
  if (s == 0) s = orient_sign(q[2], q[3], q[4], q[5], q[6], q[7]);
  return s;
}

/* is every point of pts strictly outside edge k of convex polygon q (n verts)?  A quad without
   area (o == 0: a segment or a point) has no inside: its edge separates when the points lie strictly
   on one side of it, whichever side. */
static int edge_separates(const double* q, int n, int o, int k, const double* pts, int npts) {
  double ax = q[2 * k], ay = q[2 * k + 1];
  double bx = q[2 * ((k + 1) % n)], by = q[2 * ((k + 1) % n) + 1];
  if (o == 0) {
    int pos = 0, neg = 0;
    for (int m = 0; m < npts; ++m) {
      int sg = orient_sign(ax, ay, bx, by, pts[2 * m], pts[2 * m + 1]);
      pos += sg > 0; neg += sg < 0;
    }
    return pos == npts || neg == npts;
  }
  for (int m = 0; m < npts; ++m)
    if (orient_sign(ax, ay, bx, by, pts[2 * m], pts[2 * m + 1]) * o >= 0) return 0;
  return 1;
}

/* closed segments (a, b) and (c, d) share a point (exact) */
static int on_segment(double ax, double ay, double bx, double by, double px, double py) {
  return orient_sign(ax, ay, bx, by, px, py) == 0 && fmin(ax, bx) <= px && px <= fmax(ax, bx) &&
         fmin(ay, by) <= py && py <= fmax(ay, by);
}
static int segments_meet(const double* a, const double* b, const double* c, const double* d) {
  int o1 = orient_sign(a[0], a[1], b[0], b[1], c[0], c[1]), o2 = orient_sign(a[0], a[1], b[0], b[1], d[0], d[1]);
  int o3 = orient_sign(c[0], c[1], d[0], d[1], a[0], a[1]), o4 = orient_sign(c[0], c[1], d[0], d[1], b[0], b[1]);
  if (o1 * o2 < 0 && o3 * o4 < 0) return 1;
  return on_segment(a[0], a[1], b[0], b[1], c[0], c[1]) || on_segment(a[0], a[1], b[0], b[1], d[0], d[1]) ||
         on_segment(c[0], c[1], d[0], d[1], a[0], a[1]) || on_segment(c[0], c[1], d[0], d[1], b[0], b[1]);
}

/* closed-set intersection of two convex quads */
static int quads_intersect(const double a[8], const double b[8]) {
  int oa = quad_orientation(a), ob = quad_orientation(b);
  if (oa == 0 && ob == 0) { /* two quads without area (segments / points): they meet iff their rings do */
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 4; ++j)
        if (segments_meet(a + 2 * i, a + 2 * ((i + 1) % 4), b + 2 * j, b + 2 * ((j + 1) % 4))) return 1;
    return 0;
  }
  for (int k = 0; k < 4; ++k) if (edge_separates(a, 4, oa, k, b, 4)) return 0;
  for (int k = 0; k < 4; ++k) if (edge_separates(b, 4, ob, k, a, 4)) return 0;
  return 1;
}

/* closed-set intersection of a convex quad and a segment */
static int quad_intersects_segment(const double q[8], const double s[4]) {
  int o = quad_orientation(q);
  for (int k = 0; k < 4; ++k) if (edge_separates(q, 4, o, k, s, 2)) return 0;
  int pos = 0, neg = 0;
  for (int m = 0; m < 4; ++m) {
    int sg = orient_sign(s[0], s[1], s[2], s[3], q[2 * m], q[2 * m + 1]);
    pos += sg > 0; neg += sg < 0;
  }
  if (pos == 4 || neg == 4) return 0;
  return 1;
}

/* Entity.get_bounding_box_points, entity/base.py:100-138:
 * corners [(cx-L/2, cy+W/2), (cx+L/2, cy+W/2), (cx+L/2, cy-W/2), (cx-L/2, cy-W/2)]
 * times R=[[cos h, sin h], [-sin h, cos h]] (einsum "ij,jk->ik") plus (x, y). */
static void box_points(double x, double y, double h, double W, double L, double cx, double cy,
                       double out[8]) {
  double c = cos(h), s = sin(h);
  double R00 = c, R01 = s, R10 = -s, R11 = c;
  double px[4] = {cx - 0.5 * L, cx + 0.5 * L, cx + 0.5 * L, cx - 0.5 * L};
  double py[4] = {cy + 0.5 * W, cy + 0.5 * W, cy - 0.5 * W, cy - 0.5 * W};
  for (int i = 0; i < 4; ++i) {
    out[2 * i] = x + (px[i] * R00 + py[i] * R10);
    out[2 * i + 1] = y + (px[i] * R01 + py[i] * R11);
  }
}

/* ------------------------------------------------------------------------- */
static inline double norm2(double a, double b) { return sqrt(a * a + b * b); }
static inline double norm3(double a, double b, double c) { return sqrt(a * a + b * b + c * c); }
static inline double dot2(double a0, double a1, double b0, double b1) { return a0 * b0 + a1 * b1; }
static inline double py_max(double a, double b) { return b > a ? b : a; } /* Python max(a, b) */
static inline double py_min(double a, double b) { return b < a ? b : a; } /* Python min(a, b) */
static inline double np_sign(double v) { return v > 0 ? 1.0 : (v < 0 ? -1.0 : (v == 0 ? 0.0 : NAN)); }
static inline double np_clip(double v, double lo, double hi) { return fmin(fmax(v, lo), hi); }

#define IDX(n, s) ((int64_t)(n) * sc->n_slots + (s))

static int is_agent_kind(int k) { return k >= SG_KIND_AGENT_REPLAY; }

/* ------------------------------------------------------------------------- */
/* RSS: metrics/rss/callback.py + rss_utils.py                                 */
typedef struct {
  double position[2], heading[2], velocity[2], box[8], length, width;
} RssEnt;

/* rss_utils.py:7-21 inverse_direction (normalised) */
static void inverse_direction(const double v[2], double out[2]) {
  double n = norm2(v[1], v[0]);
  out[0] = v[1] / n;
  out[1] = -v[0] / n;
}
/* rss_utils.py:24-45 coord_change */
static void coord_change(const double v[2], const double dir[2], const double c[2], double out[2]) {
  double inv[2];
  inverse_direction(dir, inv);
  double d0 = v[0] - c[0], d1 = v[1] - c[1];
  out[0] = dot2(d0, d1, inv[0], inv[1]);
  out[1] = dot2(d0, d1, dir[0], dir[1]);
}
/* callback.py:340-386 get_entity_parameters */
static void rss_entity_params(const double pose[6], const double vel[6], const double bx[4],
                              const double eh[2], const double einv[2], const double epos[2],
                              RssEnt* o) {
  double dir[2] = {cos(pose[3]), sin(pose[3])}; /* rss_utils.py:95-103 */
  coord_change(pose, eh, epos, o->position);
  o->heading[0] = dot2(dir[0], dir[1], einv[0], einv[1]);
  o->heading[1] = dot2(dir[0], dir[1], eh[0], eh[1]);
  o->velocity[0] = dot2(vel[0], vel[1], einv[0], einv[1]);
  o->velocity[1] = dot2(vel[0], vel[1], eh[0], eh[1]);
  double pts[8];
  box_points(pose[0], pose[1], pose[3], bx[0], bx[1], bx[2], bx[3], pts);
  for (int i = 0; i < 4; ++i) coord_change(pts + 2 * i, eh, epos, o->box + 2 * i);
  o->length = bx[1];
  o->width = bx[0];
}
/* callback.py:454-472 */
static double long_dist_same_direction(double vf, double vr, double a, double RT, double MINA) {
  double v = vr * RT + py_min(vf * vf / (2 * a), 0.5 * a * (RT * RT)) +
             ((vr + RT * a) * (vr + RT * a)) / (2 * MINA) - vf * vf / (2 * a);
  return py_max(0, v);
}
/* callback.py:474-492 */
static double long_dist_opp_direction(double v1, double v2, double a, double RT, double MINA) {
  double av2 = fabs(v2);
  double v = (2 * v1 + RT * a) * RT / 2 + ((v1 + RT * a) * (v1 + RT * a)) / (2 * MINA) +
             (2 * av2 + RT * a) * RT / 2 + ((av2 + RT * a) * (av2 + RT * a)) / (2 * MINA);
  return py_max(0, v);
}
/* callback.py:494-505 */
static double lat_dist(double v, double amax, double amin, double RT) {
  double x = 0.5 * RT * (2 * v + RT * amax) + ((v + RT * amax) * (v + RT * amax)) / (2 * amin) -
             0.5 * (RT * RT) * amax - ((RT * amax) * (RT * amax)) / (2 * amin);
  return py_max(0, x);
}
/* callback.py:230-269 */
static double safe_longitudinal_distance(const SgParams* p, const RssEnt* ego, const RssEnt* haz) {
  double CLR = p->rss_min_safe_clearance, RT = p->rss_response_time;
  double dp = dot2(ego->heading[0], ego->heading[1], haz->heading[0], haz->heading[1]);
  double a = fabs(p->rss_max_long_accel * dp);
  double d0;
  if (dp > 0) {
    double vf, vr;
    if (ego->position[1] > haz->position[1]) { /* rss_utils.py:79-92 ahead */
      vf = norm2(ego->velocity[0], ego->velocity[1]);
      vr = dot2(haz->velocity[0], haz->velocity[1], ego->heading[0], ego->heading[1]);
    } else {
      vf = dot2(haz->velocity[0], haz->velocity[1], ego->heading[0], ego->heading[1]);
      vr = norm2(ego->velocity[0], ego->velocity[1]);
    }
    if (vr == 0.0) return CLR + 0.5 * ego->length;
    d0 = long_dist_same_direction(vf, vr, a, RT, p->rss_min_long_accel);
  } else {
    double v1 = fabs(dot2(ego->velocity[0], ego->velocity[1], ego->heading[0], ego->heading[1]));
    double v2 = -fabs(dot2(haz->velocity[0], haz->velocity[1], ego->heading[0], ego->heading[1]));
    if (np_sign(haz->position[1]) == np_sign(haz->velocity[1])) return CLR + 0.5 * ego->length;
    d0 = long_dist_opp_direction(v1, v2, a, RT, p->rss_min_long_accel);
  }
  return d0 + CLR + 0.5 * ego->length;
}
/* callback.py:271-302 */
static double safe_lateral_distance(const SgParams* p, const RssEnt* ego, const RssEnt* haz) {
  double CLR = p->rss_min_safe_clearance, RT = p->rss_response_time;
  double v = haz->velocity[0];
  double inv[2];
  inverse_direction(ego->heading, inv);
  double k = fabs(dot2(inv[0], inv[1], haz->heading[0], haz->heading[1]));
  double amax = p->rss_max_long_accel * k, amin = p->rss_min_long_accel * k;
  double d0;
  if (np_sign(-haz->position[0]) == np_sign(v)) {
    v = fabs(v);
    if (v == 0.0) return CLR + 0.5 * ego->width;
    d0 = lat_dist(v, amax, amin, RT);
  } else {
    d0 = 0;
  }
  return d0 + CLR + 0.5 * ego->width;
}
/* callback.py:124-166 safe_ratios (the `ego_entity in safe_distances` test is always False) */
static void safe_ratios(const RssEnt* ego, const RssEnt* haz, double out[2]) {
  double safe_lat = 0.5 * ego->width, safe_long = 0.5 * ego->length;
  double inv[2];
  inverse_direction(haz->heading, inv);
  double wl_inv = fabs(dot2(haz->width, haz->length, inv[0], inv[1]));
  double wl_dir = fabs(dot2(haz->width, haz->length, haz->heading[0], haz->heading[1]));
  double actual_lat = py_max(1e-6, fabs(haz->position[0]) - 0.5 * ego->width - 0.5 * wl_inv);
  double actual_long = py_max(1e-6, fabs(haz->position[1]) - 0.5 * ego->length - 0.5 * wl_dir);
  out[0] = fabs(actual_lat / safe_lat);
  out[1] = fabs(actual_long / safe_long);
}
/* callback.py:168-228 unsafe_distance + :304-338 write_intersections + :388-452 generate_buffer.
 * st: bits0-1 last marker, bits2-3 found.  Returns the SgRssRecord appended. */
static int unsafe_distance(const RssEnt* ego, const RssEnt* haz, uint8_t* st, const double sd[2]) {
  if ((*st >> 2) & 3) return SG_RSS_FOUND; /* :186-189 */
  double slat = sd[0], slong = sd[1];
  double buffer[8] = {slat, slong, -slat, slong, -slat, -slong, slat, -slong}; /* :423-428 */
  if (quads_intersect(haz->box, buffer)) { /* :196 */
    int marker = *st & 3;
    if (marker == 1) { *st |= 2 << 2; return SG_RSS_UNSAFE_LONGITUDINAL; } /* :201-203 */
    if (marker == 2) { *st |= 1 << 2; return SG_RSS_UNSAFE_LATERAL; }      /* :204-206 */
    /* :210-226 default when no previous single-direction marker */
    double ed[2] = {ego->width, ego->length}, inv[2];
    inverse_direction(ed, inv);
    double lhs = fabs(fabs(haz->position[0]) - fabs(dot2(haz->position[0], haz->position[1], ed[0], ed[1]))) / slat;
    double rhs = fabs(fabs(haz->position[1] - dot2(haz->position[0], haz->position[1], inv[0], inv[1])) / slong);
    if (lhs > rhs) { *st |= 2 << 2; return SG_RSS_UNSAFE_LONGITUDINAL; }
    *st |= 1 << 2;
    return SG_RSS_UNSAFE_LATERAL;
  }
  /* write_intersections: lengths are the diagonals b0->b2, b1->b3 scaled x100 in y;
     widths are b0->b1, b2->b3 scaled x100 in x (:429-451) */
  double len0[4] = {slat, 100 * slong, -slat, 100 * -slong};
  double len1[4] = {-slat, 100 * slong, slat, 100 * -slong};
  double wid0[4] = {100 * slat, slong, 100 * -slat, slong};
  double wid1[4] = {100 * -slat, -slong, 100 * slat, -slong};
  int lat = quad_intersects_segment(haz->box, len0) || quad_intersects_segment(haz->box, len1);
  int lon = quad_intersects_segment(haz->box, wid0) || quad_intersects_segment(haz->box, wid1);
  if (lat && lon) return SG_RSS_BOTH;
  if (lat) { *st = (uint8_t)((*st & ~3) | 1); return SG_RSS_LATERAL; }
  if (lon) { *st = (uint8_t)((*st & ~3) | 2); return SG_RSS_LONGITUDINAL; }
  return SG_RSS_SAFE;
}

/* RSSDistances.__call__ (callback.py:57-122) + RSS._step latch (rss.py:132-159) */
static void rss_update(const SgScene* sc, const SgParams* p, SgState* st, int n) {
  int64_t nm = NM;
  int M = sc->n_slots;
  for (int s = 0; s < M; ++s) st->rss_last[IDX(n, s)] = SG_RSS_NONE;
  if (st->t[n] == 0.0) return; /* :72 */
  int es = sc->ego_slot[n];
  int64_t ei = IDX(n, es);
  if (!st->present[ei]) return; /* reference would raise KeyError */
  double epose[6], evel[6], ebox[4];
  for (int f = 0; f < 6; ++f) { epose[f] = st->pose[f * nm + ei]; evel[f] = st->vel[f * nm + ei]; }
  for (int f = 0; f < 4; ++f) ebox[f] = sc->box[f * nm + ei];
  double eh[2] = {cos(epose[3]), sin(epose[3])}, einv[2], epos[2] = {epose[0], epose[1]};
  inverse_direction(eh, einv);
  RssEnt ego;
  rss_entity_params(epose, evel, ebox, eh, einv, epos, &ego);
  for (int s = 0; s < M; ++s) {
    int64_t i = IDX(n, s);
    if (s == es || !st->present[i]) continue;
    double pose[6], vel[6], bx[4];
    for (int f = 0; f < 6; ++f) { pose[f] = st->pose[f * nm + i]; vel[f] = st->vel[f * nm + i]; }
    for (int f = 0; f < 4; ++f) bx[f] = sc->box[f * nm + i];
    RssEnt haz;
    rss_entity_params(pose, vel, bx, eh, einv, epos, &haz);
    double sd[2];
    sd[1] = fabs(safe_longitudinal_distance(p, &ego, &haz)); /* :101-103 */
    sd[0] = fabs(safe_lateral_distance(p, &ego, &haz));
    st->safe_dist[i] = sd[0];
    st->safe_dist[nm + i] = sd[1];
    double ratio[2];
    safe_ratios(&ego, &haz, ratio);
    st->safe_ratio[i] = ratio[0];
    st->safe_ratio[nm + i] = ratio[1];
    int rec = unsafe_distance(&ego, &haz, &st->rss_state[i], sd);
    st->rss_last[i] = (uint8_t)rec;
    int found = (st->rss_state[i] >> 2) & 3;
    if (found == 2) st->rss_flags[n] |= 1; /* rss.py:71-84 */
    if (found == 1) st->rss_flags[n] |= 2; /* rss.py:86-103 */
  }
}

/* ------------------------------------------------------------------------- */
/* collisions: state/utils.py:10-49 + utils.py:28-62 + metrics/collision.py:70-75 */
static void collisions_update(const SgScene* sc, const SgParams* p, SgState* st, int n,
                              int record_metric, int* out_any, int* out_first_hit) {
  int64_t nm = NM;
  int M = sc->n_slots, W = (M + 31) / 32;
  double* pts = (double*)malloc(sizeof(double) * 12 * M);
  double* env = pts + 8 * M;
  for (int s = 0; s < M; ++s) {
    int64_t i = IDX(n, s);
    if (!st->present[i]) continue;
    box_points(st->pose[i], st->pose[nm + i], st->pose[3 * nm + i], sc->box[i], sc->box[nm + i],
               sc->box[2 * nm + i], sc->box[3 * nm + i], pts + 8 * s);
    const double* q = pts + 8 * s;
    double* e = env + 4 * s;
    e[0] = fmin(fmin(q[0], q[2]), fmin(q[4], q[6]));
    e[1] = fmin(fmin(q[1], q[3]), fmin(q[5], q[7]));
    e[2] = fmax(fmax(q[0], q[2]), fmax(q[4], q[6]));
    e[3] = fmax(fmax(q[1], q[3]), fmax(q[5], q[7]));
  }
  uint32_t* ego_now = (uint32_t*)calloc(W, sizeof(uint32_t));
  int es = sc->ego_slot[n], fs = sc->first_slot[n];
  int any = 0, fa = -1, fb = -1, first_hit = 0;
  int64_t npairs = 0;
  if (p->features & SG_FEAT_COLL_MATRIX)
    memset(st->coll_mask + (int64_t)n * M * W, 0, sizeof(uint32_t) * M * W);
  for (int a = 0; a < M; ++a) {
    if (!st->present[IDX(n, a)]) continue;
    for (int b = a + 1; b < M; ++b) {
      if (!st->present[IDX(n, b)]) continue;
      const double *qa = pts + 8 * a, *qb = pts + 8 * b;
      /* STRtree.query first filters by envelope intersection (closed, exact on fp64) */
      const double *ea = env + 4 * a, *eb = env + 4 * b;
      if (ea[0] > eb[2] || eb[0] > ea[2] || ea[1] > eb[3] || eb[1] > ea[3]) continue;
      if (memcmp(qa, qb, 64) == 0) continue; /* `g != g_prime`, utils.py:58 */
      if (!quads_intersect(qa, qb)) continue;
      ++npairs;
      if (!any) { any = 1; fa = a; fb = b; }
      if (a == fs || b == fs) first_hit = 1;
      st->collided[IDX(n, a)] = 1;
      st->collided[IDX(n, b)] = 1;
      if (a == es) ego_now[b >> 5] |= 1u << (b & 31);
      if (b == es) ego_now[a >> 5] |= 1u << (a & 31);
      if (p->features & SG_FEAT_COLL_MATRIX) {
        st->coll_mask[((int64_t)n * M + a) * W + (b >> 5)] |= 1u << (b & 31);
        st->coll_mask[((int64_t)n * M + b) * W + (a >> 5)] |= 1u << (a & 31);
      }
    }
  }
  st->n_pair_ticks[n] += npairs;
  if (any && st->first_coll_tick[n] < 0) {
    st->first_coll_tick[n] = st->tick[n];
    st->first_coll_pair[2 * n] = fa;
    st->first_coll_pair[2 * n + 1] = fb;
  }
  if (record_metric) { /* CollisionMetric._step */
    uint32_t* last = st->ego_hits + (int64_t)n * W;
    for (int s = 0; s < M; ++s) {
      uint32_t bit = 1u << (s & 31);
      if ((ego_now[s >> 5] & bit) && !(last[s >> 5] & bit)) {
        int32_t k = st->event_count[0]++;
        if (k < st->event_cap) {
          SgEvent ev = {n, st->tick[n], s, 0, st->t[n]};
          st->events[k] = ev;
        }
      }
    }
    memcpy(last, ego_now, sizeof(uint32_t) * W);
  }
  free(ego_now);
  free(pts);
  *out_any = any;
  *out_first_hit = first_hit;
}

/* ------------------------------------------------------------------------- */
static void record_trace(const SgScene* sc, SgState* st, int n) {
  int k = st->tick[n];
  if (k >= st->trace_cap) return;
  int64_t nm = NM;
  for (int s = 0; s < sc->n_slots; ++s) {
    int64_t i = IDX(n, s);
    st->trace_present[(int64_t)k * nm + i] = st->present[i];
    for (int f = 0; f < 6; ++f) st->trace_pose[((int64_t)k * 6 + f) * nm + i] = st->pose[f * nm + i];
  }
  st->trace_t[(int64_t)k * sc->n_scenarios + n] = st->t[n];
}

static const double* slot_rows(const SgScene* sc, int64_t i, int64_t* K) {
  *K = sc->traj_off[i + 1] - sc->traj_off[i];
  return sc->traj_rows + sc->traj_off[i] * 7;
}

/* State.reset, state/state.py:106-143 */
int sgo_reset(const SgScene* sc, const SgParams* p, SgState* st, int device, void* stream) {
  (void)device; (void)stream;
  int64_t nm = NM;
  int M = sc->n_slots, W = (M + 31) / 32;
  st->event_count[0] = 0;
  for (int n = 0; n < sc->n_scenarios; ++n) {
    double t0 = sc->t0[n];
    for (int s = 0; s < M; ++s) {
      int64_t i = IDX(n, s);
      int kind = sc->kind[i];
      double pose[6] = {0, 0, 0, 0, 0, 0}, vel[6] = {0, 0, 0, 0, 0, 0};
      int pres = 0;
      if (kind != SG_KIND_EMPTY) {
        int64_t K;
        const double* rows = slot_rows(sc, i, &K);
        int mode = (K == 1) ? EXT_TRUE : (p->persist ? EXT_CLAMP : EXT_NONE); /* :123-129 */
        pres = position_at_t(rows, K, t0, mode, pose);
        if (pres) velocity_at_t(rows, K, t0, vel); /* :132 */
      }
      st->present[i] = (uint8_t)pres;
      for (int f = 0; f < 6; ++f) {
        st->pose[f * nm + i] = pres ? pose[f] : 0.0;
        st->vel[f * nm + i] = pres ? vel[f] : 0.0;
      }
      st->dist[i] = 0.0;
      /* VehicleController._reset controller.py:100-103 ; PedestrianController._reset :23 */
      st->speed[i] = (kind == SG_KIND_VEHICLE || kind == SG_KIND_PID) ? norm2(vel[0], vel[1]) : 0.0;
      for (int f = 0; f < 3; ++f) st->pid_err[f * nm + i] = 0.0; /* controller.py:198-203 */
      st->goal_idx[i] = 0;
      st->force[i] = 0.0;
      st->force[nm + i] = 0.0;
      st->cur_own[i] = 1;
      st->collided[i] = 0;
      st->rss_state[i] = 0;
      st->rss_last[i] = SG_RSS_NONE;
      st->safe_dist[i] = 0.0; st->safe_dist[nm + i] = 0.0;           /* callback.py:51 */
      st->safe_ratio[i] = INFINITY; st->safe_ratio[nm + i] = INFINITY; /* callback.py:53-55 */
    }
    st->t[n] = t0;
    st->prev_t[n] = t0 - 0.1; /* :135 */
    st->tick[n] = 0;
    st->done[n] = 0;
    st->cur_union[n] = 1;
    st->first_coll_tick[n] = -1;
    st->first_coll_pair[2 * n] = -1;
    st->first_coll_pair[2 * n + 1] = -1;
    st->n_pair_ticks[n] = 0;
    st->rss_flags[n] = 0;
    for (int w = 0; w < W; ++w) st->ego_hits[(int64_t)n * W + w] = 0;
    if (p->features & SG_FEAT_RSS) rss_update(sc, p, st, n); /* :137-139 update_callbacks */
    /* Metric.reset: metrics/trajectory.py:13-18, 33-37 */
    int64_t ei = IDX(n, sc->ego_slot[n]);
    double sp = norm3(st->vel[ei], st->vel[nm + ei], st->vel[2 * nm + ei]);
    st->ego_avg_speed[n] = sp;
    st->ego_avg_t[n] = 0.0;
    st->ego_max_speed[n] = sp;
    st->ego_dist[n] = 0.0;
    if (st->trace_cap > 0) record_trace(sc, st, n);
  }
  return 0;
}

/* PedestrianAgent._step + SocialForce._step + PedestrianController._step
 * pedestrian/agent.py:49-69, social_force.py:44-222, controller.py:25-46, sensor.py:54-64,
 * state/state.py:340-372 */
static double g_ngon_cs[64][2];
static int g_ngon_init = 0;
static void ngon_init(void) {
  if (g_ngon_init) return;
  double inc = (2.0 * M_PI) / 64;
  for (int i = 0; i < 64; ++i) {
    double ang = 0.0 + -1.0 * i * inc;
    g_ngon_cs[i][0] = cos(ang);
    g_ngon_cs[i][1] = sin(ang);
  }
  g_ngon_init = 1;
}
/* strict interior of Point(x, y).buffer(r) (64-gon, clockwise ring) */
static int in_buffer(double x, double y, double r, double qx, double qy) {
  double v[64][2];
  for (int k = 0; k < 64; ++k) { v[k][0] = x + r * g_ngon_cs[k][0]; v[k][1] = y + r * g_ngon_cs[k][1]; }
  for (int k = 0; k < 64; ++k) {
    int k1 = (k + 1) & 63;
    if (orient_sign(v[k][0], v[k][1], v[k1][0], v[k1][1], qx, qy) >= 0) return 0;
  }
  return 1;
}
/* LineString(route).project(Point) -- GEOS LengthIndexedLine semantics */
static double route_project(const double* xy, int64_t R, double px, double py) {
  double best_d = INFINITY, best_s = 0.0, s0 = 0.0;
  for (int64_t i = 0; i + 1 < R; ++i) {
    double ax = xy[2 * i], ay = xy[2 * i + 1], bx = xy[2 * i + 2], by = xy[2 * i + 3];
    double dx = bx - ax, dy = by - ay, seg2 = dx * dx + dy * dy, seglen = sqrt(seg2);
    double r = seg2 == 0.0 ? 0.0 : ((px - ax) * dx + (py - ay) * dy) / seg2;
    double qx, qy, sl;
    if (r <= 0.0) { qx = ax; qy = ay; sl = 0.0; }
    else if (r >= 1.0) { qx = bx; qy = by; sl = seglen; }
    else { qx = ax + r * dx; qy = ay + r * dy; sl = r * seglen; }
    double d = hypot(px - qx, py - qy);
    if (d < best_d) { best_d = d; best_s = s0 + sl; }
    s0 += seglen;
  }
  return best_s;
}

static double sight_weight(const SgParams* p, const double F[2], const double view[2]) {
  /* social_force.py:213-222 */
  double dd = dot2(view[0], view[1], F[0], F[1]) / (norm2(F[0], F[1]) + 0.0000000001);
  if (dd >= cos(p->sf_sight_angle / 2 * M_PI / 180)) return 1.0;
  return p->sf_sight_weight;
}

/* ---- road-network surfaces (SgScene.rn_*): restated Shapely semantics.  The reference builds
   driveable / walkable / impenetrable_surface with shapely.ops.unary_union
   (road_network/road_network.py:306-328) and asks `surface.contains(Point)` (state/state.py:401-407,
   pedestrian/social_force.py:87,96) and `nearest_points(surface, Point)` (social_force.py:204).
   Restated on the polygon soup: contained = strictly inside one member polygon (crossing parity over
   all of its rings with exact orientation signs; on a ring = boundary); nearest point = the point
   itself when it lies in the closed surface, else the closest point on the rings
   (GEOS LineSegment::closestPoint).  GEOS itself is absent: parity unpinned for points exactly on an
   edge shared by two member polygons (the dissolved union has no edge there). */
static int polygon_side(const double* edges, int64_t e0, int64_t e1, double px, double py) {
  int inside = 0;
  for (int64_t e = e0; e < e1; ++e) {
    double ax = edges[4 * e], ay = edges[4 * e + 1], bx = edges[4 * e + 2], by = edges[4 * e + 3];
    int straddles = (ay > py) != (by > py);
    int in_box = px >= fmin(ax, bx) && px <= fmax(ax, bx) && py >= fmin(ay, by) && py <= fmax(ay, by);
    if (!straddles && !in_box) continue;
    int o = orient_sign(ax, ay, bx, by, px, py);
    if (o == 0 && in_box) return 0;
    if (straddles && ((o > 0) == (by > ay))) inside = !inside;
  }
  return inside ? 1 : -1;
}
static int surface_network(const SgScene* sc, int n) {
  if (!sc->rn_of || sc->n_networks <= 0) return -1;
  return sc->rn_of[n];
}
static int surface_has_area(const SgScene* sc, int n, int k) {
  int r = surface_network(sc, n);
  return r >= 0 && sc->rn_has_area[3 * r + k] != 0;
}
static int surface_contains(const SgScene* sc, int n, int k, double px, double py) {
  int r = surface_network(sc, n);
  if (r < 0) return 0;
  for (int64_t q = sc->rn_poly_off[3 * r + k]; q < sc->rn_poly_off[3 * r + k + 1]; ++q)
    if (polygon_side(sc->rn_edges, sc->rn_edge_off[q], sc->rn_edge_off[q + 1], px, py) > 0) return 1;
  return 0;
}
static void surface_nearest(const SgScene* sc, int n, int k, double px, double py, double out[2]) {
  int r = surface_network(sc, n);
  out[0] = px; out[1] = py;
  int64_t q0 = sc->rn_poly_off[3 * r + k], q1 = sc->rn_poly_off[3 * r + k + 1];
  for (int64_t q = q0; q < q1; ++q)
    if (polygon_side(sc->rn_edges, sc->rn_edge_off[q], sc->rn_edge_off[q + 1], px, py) >= 0) return;
  double best = INFINITY;
  for (int64_t e = sc->rn_edge_off[q0]; e < sc->rn_edge_off[q1]; ++e) {
    const double* E = sc->rn_edges + 4 * e;
    double ax = E[0], ay = E[1], bx = E[2], by = E[3];
    double dx = bx - ax, dy = by - ay, len2 = dx * dx + dy * dy, cx = ax, cy = ay;
    if (len2 > 0.0) {
      double f = ((px - ax) * dx + (py - ay) * dy) / len2;
      if (f > 0.0 && f < 1.0) { cx = ax + f * dx; cy = ay + f * dy; }
      else {
        double da = (px - ax) * (px - ax) + (py - ay) * (py - ay);
        double db = (px - bx) * (px - bx) + (py - by) * (py - by);
        if (db < da) { cx = bx; cy = by; }
      }
    }
    double d2 = (px - cx) * (px - cx) + (py - cy) * (py - cy);
    if (d2 < best) { best = d2; out[0] = cx; out[1] = cy; }
  }
}
/* SocialForce._force_boundary, pedestrian/social_force.py:190-211 */
static void boundary_force(const SgScene* sc, int n, int k, double px, double py, double U, double R, double out[2]) {
  double c[2];
  surface_nearest(sc, n, k, px, py, c);
  double r0 = px - c[0], r1 = py - c[1];
  double rn = sqrt(r0 * r0 + r1 * r1);
  double u0 = r0 / (rn + 0.0000000001), u1 = r1 / (rn + 0.0000000001);
  double ex = exp(-rn / R);
  out[0] = U / R * u0 * ex;
  out[1] = U / R * u1 * ex;
}

/* SocialForce noise with std != 0: the ENGINE's counter-based stream (include/sg_b200.h, SgParams.sf_std_*);
   the reference draws from numpy's global generator (social_force.py:106-108), which nothing can reproduce. */
static uint64_t splitmix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ULL;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}
static void noise2(uint64_t seed, int64_t i, int tick, double z[2]) {
  uint64_t a = splitmix64(seed ^ splitmix64((uint64_t)i * 0x9E3779B97F4A7C15ULL + (uint64_t)(unsigned)tick));
  uint64_t b = splitmix64(a);
  double u1 = ((double)(a >> 11) + 1.0) * (1.0 / 9007199254740992.0);
  double u2 = (double)(b >> 11) * (1.0 / 9007199254740992.0);
  double r = sqrt(-2.0 * log(u1));
  z[0] = r * cos(2.0 * M_PI * u2);
  z[1] = r * sin(2.0 * M_PI * u2);
}

static void pedestrian_step(const SgScene* sc, const SgParams* p, SgState* st, int n, int s,
                            double next_t, double out[6]) {
  int64_t nm = NM, i = IDX(n, s);
  int M = sc->n_slots;
  double pose[6], vel[6];
  for (int f = 0; f < 6; ++f) { pose[f] = st->pose[f * nm + i]; vel[f] = st->vel[f * nm + i]; }
  int64_t r0 = sc->route_off[i], R = sc->route_off[i + 1] - r0;
  const double* route = sc->route_xy + 2 * r0;
  int goal = st->goal_idx[i];
  double speed, heading;
  if (goal <= R - 1) { /* agent.py:60-62 */
    double sarc = route_project(route, R, pose[0], pose[1]);
    double arc = 0.0;
    int last = 0;
    for (int64_t k = 0; k < R; ++k) {
      if (k > 0) arc += norm2(route[2 * k] - route[2 * k - 2], route[2 * k + 1] - route[2 * k - 1]);
      if (arc <= sarc) last = (int)k;
    }
    goal = last + 1;
  }
  if (goal <= R - 1) {
    double speed_desired = sc->ped_speed_desired[i];
    /* _force_to_goal, social_force.py:119-138 */
    double dvx = route[2 * goal] - pose[0], dvy = route[2 * goal + 1] - pose[1];
    double dn = norm2(dvx, dvy);
    if (dn == 0) dn += 0.000000001;
    double ux = dvx / dn, uy = dvy / dn;
    double k = 1 / p->sf_relaxation_time;
    double F[2] = {k * (speed_desired * ux - vel[0]), k * (speed_desired * uy - vel[1])};
    double thr = p->ped_distance_threshold;
    double ch = cos(p->ped_head_rot_angle), sh = sin(p->ped_head_rot_angle); /* viewer/utils.py:6-17 */
    for (int o = 0; o < M; ++o) { /* state.poses order = slot order */
      int64_t j = IDX(n, o);
      if (o == s || !st->present[j] || sc->etype[j] != SG_ETYPE_PEDESTRIAN) continue;
      double ox = st->pose[j], oy = st->pose[nm + j];
      if (!in_buffer(pose[0], pose[1], thr, ox, oy)) continue;
      double ovx = st->vel[j], ovy = st->vel[nm + j];
      double vdx = ovx * ch + ovy * -sh, vdy = ovx * sh + ovy * ch; /* X.dot(R.T) */
      double vn = norm2(vdx, vdy) + 0.0000000001;
      double view[2] = {vdx / vn, vdy / vn};
      /* _force_pedestrian_repulsion :140-176 */
      double rx = pose[0] - ox, ry = pose[1] - oy, rn = norm2(rx, ry);
      double vmag = norm2(ovx, ovy) + 0.0000000001;
      double uox = ovx / vmag, uoy = ovy / vmag;
      double other_step = vmag * (next_t - st->t[n]);
      double r2x = rx - other_step * uox, r2y = ry - other_step * uoy;
      double r2n = norm2(r2x, r2y) + 0.0000000001;
      double b = (1.0 / 2) * sqrt((rn + r2n) * (rn + r2n) - other_step * other_step);
      double c0 = (1.0 / 4) * (1 / b) * (rn + r2n);
      double dbx = c0 * (rx / rn + r2x / r2n), dby = c0 * (ry / rn + r2y / r2n);
      double g = p->sf_ped_repulse_V / p->sf_ped_repulse_sigma * exp(-b / p->sf_ped_repulse_sigma);
      double Frep[2] = {g * dbx, g * dby};
      /* _force_pedestrian_attraction :178-188 */
      double Fatt[2] = {2 * p->sf_ped_attract_C * rx, 2 * p->sf_ped_attract_C * ry};
      if (p->sf_sight_weight_use) {
        double w = sight_weight(p, Frep, view);
        F[0] += w * Frep[0]; F[1] += w * Frep[1];
        w = sight_weight(p, Fatt, view);
        F[0] += w * Fatt[0]; F[1] += w * Fatt[1];
      } else {
        F[0] += Fatt[0]; F[1] += Fatt[1];
        F[0] += Frep[0]; F[1] += Frep[1];
      }
    }
    /* boundary forces, social_force.py:83-104 (skipped for surfaces without area, e.g. an empty network) */
    if (surface_has_area(sc, n, 1) && surface_contains(sc, n, 1, pose[0], pose[1])) {
      double fb[2];
      boundary_force(sc, n, 1, pose[0], pose[1], p->sf_boundary_repulse_U, p->sf_boundary_repulse_R, fb);
      F[0] += fb[0]; F[1] += fb[1];
    }
    if (surface_has_area(sc, n, 2)) {
      double sign = 1 - 2 * surface_contains(sc, n, 2, pose[0], pose[1]);
      double fb[2];
      boundary_force(sc, n, 2, pose[0], pose[1], p->sf_imp_boundary_repulse_U, p->sf_imp_boundary_repulse_R, fb);
      F[0] += sign * fb[0]; F[1] += sign * fb[1];
    }
    /* random fluctuations :106-108; np.random.normal(bias, 0) == bias */
    double speed_rand = p->sf_bias_lon, heading_rand = p->sf_bias_lat;
    if (p->sf_std_lon != 0.0 || p->sf_std_lat != 0.0) {
      double z[2];
      noise2(p->sf_noise_seed, i, st->tick[n], z);
      speed_rand = p->sf_bias_lon + p->sf_std_lon * z[0];
      heading_rand = p->sf_bias_lat + p->sf_std_lat * z[1];
    }
    speed = py_min(norm2(F[0], F[1]) + speed_rand, speed_desired * p->sf_max_speed_factor);
    heading = atan2(F[1], F[0]) + heading_rand;
    st->force[i] = F[0];
    st->force[nm + i] = F[1];
  } else { /* agent.py:65-68 reached goal */
    speed = 0; heading = 0;
    st->force[i] = 0.0;
    st->force[nm + i] = 0.0;
  }
  st->goal_idx[i] = goal;
  /* PedestrianController._step controller.py:38-46 (uses state.dt) */
  double sp = np_clip(speed, -p->ped_max_speed, p->ped_max_speed);
  double dt = st->t[n] - st->prev_t[n];
  st->speed[i] = sp;
  memcpy(out, pose, 48);
  out[0] = pose[0] + sp * dt * cos(heading);
  out[1] = pose[1] + sp * dt * sin(heading);
  out[3] = heading;
}

/* one ScenarioGym.step() for scenario n, scenario_gym.py:227-254 */
static void tick_scenario(const SgScene* sc, const SgParams* p, SgState* st, const SgInputs* in,
                          int n, int k_action, double* newpose, uint8_t* newpres, double* newspeed) {
  int64_t nm = NM;
  int M = sc->n_slots;
  double t = st->t[n];
  double next_t = t + p->timestep; /* :229 */
  for (int s = 0; s < M; ++s) {
    int64_t i = IDX(n, s);
    int kind = sc->kind[i];
    double* np_ = newpose + 6 * s;
    newpres[s] = 0;
    newspeed[s] = st->speed[i];
    if (kind == SG_KIND_EMPTY) continue;
    int64_t K;
    const double* rows = slot_rows(sc, i, &K);
    if (is_agent_kind(kind)) { /* :233-244 */
      if (st->present[i]) {
        if (kind == SG_KIND_AGENT_REPLAY) { /* agent.py:125-128, default extrapolate=(False, False) */
          position_at_t(rows, K, next_t, EXT_CLAMP, np_);
          newpres[s] = 1;
        } else if (kind == SG_KIND_VEHICLE || kind == SG_KIND_PID) {
          double accel, steer;
          double pose[6];
          for (int f = 0; f < 6; ++f) pose[f] = st->pose[f * nm + i];
          double h = pose[3], spd = st->speed[i], l = sc->box[nm + i];
          if (kind == SG_KIND_PID) {
            /* PIDAgent._step agent.py:144-148 + PIDController._step controller.py:205-258 */
            double tgt[6];
            position_at_t(rows, K, next_t, EXT_CLAMP, tgt);
            double e0 = tgt[0] - pose[0], e1 = tgt[1] - pose[1];
            double ch = cos(h), sh = sin(h);
            double e_lon = ch * e0 + sh * e1, e_lat = -sh * e0 + ch * e1;
            double gain_adj;
            if (spd > 5.0 && spd <= 15) gain_adj = 1.0 - 0.9 * ((spd - 5.0)) / 10.0;
            else if (spd > 15) gain_adj = 0.1;
            else gain_adj = 1.0;
            double sdt = st->t[n] - st->prev_t[n]; /* state.dt */
            double e_lat_D = (e_lat - st->pid_err[2 * nm + i]) / sdt;
            steer = (p->pid_steer_Kp * gain_adj) * e_lat + (p->pid_steer_Kd * gain_adj) * e_lat_D;
            double e_lon_D = (e_lon - st->pid_err[i]) / sdt;
            double e_lon_I = st->pid_err[nm + i] + e_lon * sdt;
            if (fabs(e_lon) > 0.1)
              accel = p->pid_accel_Kp * e_lon + p->pid_accel_Kd * e_lon_D + p->pid_accel_Ki * e_lon_I;
            else
              accel = 0.0;
            st->pid_err[2 * nm + i] = e_lat;
            st->pid_err[i] = e_lon;
            st->pid_err[nm + i] = e_lon_I;
          } else {
            if (in->actions) {
              accel = in->actions[((int64_t)k_action * 2 + 0) * nm + i];
              steer = in->actions[((int64_t)k_action * 2 + 1) * nm + i];
            } else { /* fp32 policy outputs: widening is exact */
              accel = (double)in->actions_f32[((int64_t)k_action * 2 + 0) * nm + i];
              steer = (double)in->actions_f32[((int64_t)k_action * 2 + 1) * nm + i];
            }
          }
          /* VehicleController._step controller.py:105-140; the limits are per controller instance (:64-98) */
          double max_steer = p->veh_max_steer, max_accel = p->veh_max_accel, max_speed = p->veh_max_speed;
          int allow_reverse = p->veh_allow_reverse;
          if (sc->veh_limits) {
            max_steer = sc->veh_limits[i]; max_accel = sc->veh_limits[nm + i];
            max_speed = sc->veh_limits[2 * nm + i]; allow_reverse = sc->veh_limits[3 * nm + i] != 0.0;
          }
          accel = np_clip(accel, -max_accel, max_accel);
          steer = np_clip(steer, -max_steer, max_steer);
          double dt = next_t - t;
          double dx = spd * cos(h), dy = spd * sin(h), dh = spd * tan(steer) / l;
          pose[0] += dx * dt;
          pose[1] += dy * dt;
          pose[3] += dh * dt;
          double ns = spd + accel * dt;
          if (!allow_reverse) ns = fmax(0.0, ns);
          if (!isnan(max_speed)) ns = fmin(max_speed, ns);
          newspeed[s] = ns;
          memcpy(np_, pose, 48);
          newpres[s] = 1;
        } else if (kind == SG_KIND_PEDESTRIAN) {
          pedestrian_step(sc, p, st, n, s, next_t, np_);
          newspeed[s] = st->speed[i];
          newpres[s] = 1;
        } else { /* SG_KIND_HOST */
          if (in && in->host_present && in->host_present[i]) {
            for (int f = 0; f < 6; ++f) np_[f] = in->host_pose[f * nm + i];
            newpres[s] = 1;
          } else if (p->persist) { /* :238-239 */
            for (int f = 0; f < 6; ++f) np_[f] = st->pose[f * nm + i];
            newpres[s] = 1;
          }
        }
      } else if (rows[0] >= t) { /* :240-244 agent initialised at its start position */
        position_at_t(rows, K, next_t, EXT_CLAMP, np_);
        newpres[s] = 1;
      }
    } else { /* BatchReplayEntity.step entity/batch.py:34-53 */
      double min_t = rows[0], max_t = rows[(K - 1) * 7];
      if (p->persist || K == 1 || (next_t >= min_t && next_t <= max_t)) {
        union_interp(sc, n, s, next_t, np_);
        newpres[s] = 1;
      }
    }
  }
  /* State.step -> update_poses / update_statistics, state/state.py:165-171, 203-239 */
  st->prev_t[n] = t;
  st->t[n] = next_t;
  st->tick[n] += 1;
  double dt = st->t[n] - st->prev_t[n];
  for (int s = 0; s < M; ++s) {
    int64_t i = IDX(n, s);
    if (!newpres[s]) {
      st->present[i] = 0;
      continue;
    }
    double prev[6];
    if (st->present[i]) {
      for (int f = 0; f < 6; ++f) prev[f] = st->pose[f * nm + i];
    } else { /* :219-222 newcomer: extrapolated previous pose */
      int64_t K;
      const double* rows = slot_rows(sc, i, &K);
      position_at_t(rows, K, st->prev_t[n], EXT_TRUE, prev);
    }
    double d[6];
    for (int f = 0; f < 6; ++f) {
      d[f] = newpose[6 * s + f] - prev[f];
      st->vel[f * nm + i] = d[f] / dt;
      st->pose[f * nm + i] = newpose[6 * s + f];
    }
    st->dist[i] += norm3(d[0], d[1], d[2]);
    st->present[i] = 1;
    st->speed[i] = newspeed[s];
  }
  if (st->trace_cap > 0) record_trace(sc, st, n);
  /* update_callbacks (state.py:263-266) */
  if (p->features & SG_FEAT_RSS) rss_update(sc, p, st, n);
  /* collisions are needed by terminal conditions and by CollisionMetric */
  int need_coll = (p->features & SG_FEAT_COLLISIONS) ||
                  (p->terminal & (SG_TERM_COLLISION | SG_TERM_EGO_COLLISION));
  int any_coll = 0, first_hit = 0;
  if (need_coll)
    collisions_update(sc, p, st, n, (p->features & SG_FEAT_COLLISIONS) != 0, &any_coll, &first_hit);
  /* check_terminal, state.py:268-270, 397-408 */
  int done = 0;
  if ((p->terminal & SG_TERM_MAX_LENGTH) && (st->t[n] + dt > sc->length[n])) done = 1;
  if ((p->terminal & SG_TERM_COLLISION) && any_coll) done = 1;
  if ((p->terminal & SG_TERM_EGO_COLLISION) && first_hit) done = 1; /* collisions()[entities[0]] */
  if (p->terminal & SG_TERM_EGO_OFF_ROAD) { /* state.py:401-407 */
    int64_t fi = IDX(n, sc->first_slot[n]);
    if (!(st->present[fi] && surface_contains(sc, n, 0, st->pose[fi], st->pose[nm + fi]))) done = 1;
  }
  st->done[n] = (uint8_t)done;
  /* metrics, scenario_gym.py:251-252 ; metrics/trajectory.py:20-24, 39-42, 58-60 */
  if (p->features & SG_FEAT_EGO_METRICS) {
    int64_t ei = IDX(n, sc->ego_slot[n]);
    double sp = norm3(st->vel[ei], st->vel[nm + ei], st->vel[2 * nm + ei]);
    double w = st->ego_avg_t[n] / st->t[n];
    st->ego_avg_speed[n] += (1.0 - w) * (sp - st->ego_avg_speed[n]);
    st->ego_avg_t[n] = st->t[n];
    st->ego_max_speed[n] = fmax(sp, st->ego_max_speed[n]);
    st->ego_dist[n] = st->dist[ei];
  }
}

/* ---- SgActionRng: numpy's PCG64 stream restated (third-party: numpy 2.3.5,
   numpy/random/src/pcg64/pcg64.h: pcg_setseq_128_step_r, pcg_output_xsl_rr_128_64, pcg64_advance;
   numpy/random/_common.pxd / distributions.c: next_double = (next_uint64 >> 11) * (1.0 / 2^53),
   random_uniform = off + rng * next_double).  Pinned by tests/test_action_rng.py against
   numpy.random.default_rng itself. ------------------------------------------------------------ */
typedef unsigned __int128 u128;
#define PCG_MULT ((((u128)0x2360ED051FC65DA4ULL) << 64) | (u128)0x4385DF649FCCF645ULL)

static u128 pcg_advance(u128 state, u128 inc, u128 delta) {
  u128 acc_mult = 1, acc_plus = 0, cur_mult = PCG_MULT, cur_plus = inc;
  while (delta > 0) {
    if (delta & 1) {
      acc_mult *= cur_mult;
      acc_plus = acc_plus * cur_mult + cur_plus;
    }
    cur_plus = (cur_mult + 1) * cur_plus;
    cur_mult *= cur_mult;
    delta >>= 1;
  }
  return acc_mult * state + acc_plus;
}
static uint64_t pcg_output(u128 state) {
  uint64_t hi = (uint64_t)(state >> 64), lo = (uint64_t)state, x = hi ^ lo;
  unsigned rot = (unsigned)(hi >> 58);
  return (x >> rot) | (x << ((-rot) & 63));
}

int sgo_fill_random_actions(const SgActionRng* r, int tick0, int n_ticks, int64_t nm, double* out,
                            int device, void* stream) {
  (void)device; (void)stream;
  const u128 s0 = (((u128)r->state_hi) << 64) | r->state_lo, inc = (((u128)r->inc_hi) << 64) | r->inc_lo;
  for (int k = 0; k < n_ticks; ++k)
    for (int c = 0; c < 2; ++c) {
      u128 st = pcg_advance(s0, inc, (u128)(r->offset[c] + (int64_t)(tick0 + k) * r->tick_stride));
      double* row = out + ((int64_t)k * 2 + c) * nm;
      for (int64_t i = 0; i < nm; ++i) {
        st = st * PCG_MULT + inc;
        const double u = (double)(pcg_output(st) >> 11) * (1.0 / 9007199254740992.0);
        row[i] = r->low[c] + r->scale[c] * u;
      }
    }
  return 0;
}

static int scenario_has_vehicle(const SgScene* sc, int n) {
  for (int s = 0; s < sc->n_slots; ++s)
    if (sc->kind[(int64_t)n * sc->n_slots + s] == SG_KIND_VEHICLE) return 1;
  return 0;
}

int sgo_rollout(const SgScene* sc, const SgParams* p, SgState* st, const SgInputs* in, int n_ticks,
                int device, void* stream) {
  (void)device; (void)stream;
  ngon_init();
  int M = sc->n_slots;
  double* newpose = (double*)malloc(sizeof(double) * 6 * M);
  double* newspeed = (double*)malloc(sizeof(double) * M);
  uint8_t* newpres = (uint8_t*)malloc(M);
  const int limit = n_ticks < 0 ? p->max_ticks : n_ticks;
  SgInputs inp;
  memset(&inp, 0, sizeof(inp));
  if (in) inp = *in;
  double* table = NULL;
  const int have_rows = inp.actions || inp.actions_f32 || inp.use_rng;
  if (!inp.actions && !inp.actions_f32 && inp.use_rng) { /* materialise the rows this call can consume */
    int rows = inp.n_action_ticks < limit ? inp.n_action_ticks : limit;
    if (rows < 0) rows = 0;
    table = (double*)malloc(sizeof(double) * 2 * NM * (rows > 0 ? rows : 1));
    sgo_fill_random_actions(&inp.rng, inp.rng_tick0, rows, NM, table, 0, NULL);
    inp.actions = table;
  }
  for (int n = 0; n < sc->n_scenarios; ++n) {
    /* only scenarios with VehicleController slots are bounded by the action rows */
    int lim = limit;
    if (have_rows && lim > inp.n_action_ticks && scenario_has_vehicle(sc, n)) lim = inp.n_action_ticks;
    for (int k = 0; k < lim; ++k) {
      if (st->done[n] && !inp.step_done) break;
      tick_scenario(sc, p, st, &inp, n, k, newpose, newpres, newspeed);
    }
  }
  free(table);
  free(newpose);
  free(newspeed);
  free(newpres);
  return 0;
}

int sgo_test_box_pairs(const double* pose_a, const double* box_a, const double* pose_b,
                       const double* box_b, uint8_t* out, int64_t n, int device, void* stream) {
  (void)device; (void)stream;
  for (int64_t i = 0; i < n; ++i) {
    double qa[8], qb[8];
    box_points(pose_a[3 * i], pose_a[3 * i + 1], pose_a[3 * i + 2], box_a[4 * i], box_a[4 * i + 1],
               box_a[4 * i + 2], box_a[4 * i + 3], qa);
    box_points(pose_b[3 * i], pose_b[3 * i + 1], pose_b[3 * i + 2], box_b[4 * i], box_b[4 * i + 1],
               box_b[4 * i + 2], box_b[4 * i + 3], qb);
    out[i] = (memcmp(qa, qb, 64) != 0) && quads_intersect(qa, qb);
  }
  return 0;
}

/* FutureCollisionDetector._step, reference sensor/common.py:88-105: ten (n_samples) look-ahead
   times numpy.linspace(t, t + horizon, n) = start + i * step with step = (stop - start) / (n - 1)
   and the last sample set to stop exactly (numpy/_core/function_base.py); every entity at
   trajectory.position_at_t(time) with the default extrapolate=(False, False) (clamped,
   trajectory.py:185-197); detect_collisions({ego: pose}, others) (state/utils.py:10-49). */
int sgo_future_collisions(const SgScene* sc, const double* t, const int32_t* slot, double horizon,
                          int n_samples, uint8_t* out, int device, void* stream) {
  (void)device; (void)stream;
  const int M = sc->n_slots;
  const int64_t nm = (int64_t)sc->n_scenarios * M;
  for (int n = 0; n < sc->n_scenarios; ++n) {
    const int es = slot ? slot[n] : sc->ego_slot[n];
    const double start = t[n], stop = t[n] + horizon;
    const double step = n_samples > 1 ? (stop - start) / (double)(n_samples - 1) : 0.0;
    int hit = 0;
    for (int k = 0; k < n_samples && !hit; ++k) {
      double tk = (double)k * step + start;
      if (n_samples > 1 && k == n_samples - 1) tk = stop;
      const int64_t ie = (int64_t)n * M + es;
      int64_t Ke;
      const double* re = slot_rows(sc, ie, &Ke);
      if (Ke == 0) break;
      double pe[6], qe[8];
      position_at_t(re, Ke, tk, EXT_CLAMP, pe);
      box_points(pe[0], pe[1], pe[3], sc->box[ie], sc->box[nm + ie], sc->box[2 * nm + ie],
                 sc->box[3 * nm + ie], qe);
      for (int j = 0; j < M && !hit; ++j) {
        const int64_t i = (int64_t)n * M + j;
        if (j == es || sc->kind[i] == SG_KIND_EMPTY) continue;
        int64_t K;
        const double* rows = slot_rows(sc, i, &K);
        if (K == 0) continue;
        double po[6], qo[8];
        position_at_t(rows, K, tk, EXT_CLAMP, po);
        box_points(po[0], po[1], po[3], sc->box[i], sc->box[nm + i], sc->box[2 * nm + i],
                   sc->box[3 * nm + i], qo);
        if (memcmp(qe, qo, 64) != 0 && quads_intersect(qe, qo)) hit = 1;
      }
    }
    out[n] = (uint8_t)hit;
  }
  return 0;
}

/* helpers exported for unit tests of the restated predicates */
int sgo_orient_sign(double ax, double ay, double bx, double by, double cx, double cy) {
  return orient_sign(ax, ay, bx, by, cx, cy);
}
int sgo_quads_intersect(const double* a, const double* b) { return quads_intersect(a, b); }
int sgo_quad_intersects_segment(const double* q, const double* s) { return quad_intersects_segment(q, s); }
void sgo_box_points(double x, double y, double h, double W, double L, double cx, double cy, double* out) {
  box_points(x, y, h, W, L, cx, cy, out);
}
int sgo_position_at_t(const double* rows, int64_t K, double t, int mode, double* out) {
  return position_at_t(rows, K, t, mode, out);
}
void sgo_velocity_at_t(const double* rows, int64_t K, double t, double* out) { velocity_at_t(rows, K, t, out); }
/* State.get_entities_in_radius, state/state.py:352-372 */
int sgo_entities_in_radius(const SgState* st, int n_scenarios, int n_slots, const double* x, const double* y,
                           const double* r, uint8_t* out, int device, void* stream) {
  (void)device; (void)stream;
  ngon_init();
  int64_t nm = (int64_t)n_scenarios * n_slots;
  for (int64_t i = 0; i < nm; ++i) {
    int n = (int)(i / n_slots);
    out[i] = (uint8_t)(r[n] > 0.0 && st->present[i] && in_buffer(x[n], y[n], r[n], st->pose[i], st->pose[nm + i]));
  }
  return 0;
}
/* BatchReplayEntity.add_entities, entity/batch.py:80-112: every replayed slot resampled (clamped)
   at the scenario's union knot times; a single control point is duplicated 0.1 s later (:94-96) */
int sgo_build_union_x(const SgScene* sc, int device, void* stream) {
  (void)device; (void)stream;
  int M = sc->n_slots;
  double* X = (double*)sc->union_x;
  for (int n = 0; n < sc->n_scenarios; ++n)
    for (int64_t r = sc->union_off[n]; r < sc->union_off[n + 1]; ++r)
      for (int s = 0; s < M; ++s) {
        int64_t i = (int64_t)n * M + s;
        double out[6] = {0, 0, 0, 0, 0, 0};
        if (sc->kind[i] == SG_KIND_REPLAY) {
          int64_t r0 = sc->traj_off[i], K = sc->traj_off[i + 1] - r0;
          const double* rows = sc->traj_rows + r0 * 7;
          double t = sc->union_t[r];
          if (K == 1) {
            double x_lo = rows[0], x_hi = x_lo + 1e-1;
            int inside = !(t < x_lo) && !(t > x_hi);
            double w1 = (t - x_lo) / (x_hi - x_lo), w0 = (x_hi - t) / (x_hi - x_lo);
            for (int f = 0; f < 6; ++f) out[f] = inside ? w1 * rows[1 + f] + w0 * rows[1 + f] : rows[1 + f];
          } else if (K > 1) {
            position_at_t(rows, K, t, EXT_CLAMP, out);
          }
        }
        for (int f = 0; f < 6; ++f) X[(r * 6 + f) * M + s] = out[f];
      }
  return 0;
}
int sgo_test_trajectory(const double* rows, int64_t K, const double* t, int64_t n, int mode, double* pos,
                        uint8_t* ok, double* vel, int device, void* stream) {
  (void)device; (void)stream;
  for (int64_t i = 0; i < n; ++i) {
    double out[6] = {0, 0, 0, 0, 0, 0};
    ok[i] = (uint8_t)position_at_t(rows, K, t[i], mode, out);
    memcpy(pos + 6 * i, out, 48);
    if (vel) velocity_at_t(rows, K, t[i], vel + 6 * i);
  }
  return 0;
}
int sgo_polygon_side(const double* edges, int64_t n_edges, double px, double py) {
  return polygon_side(edges, 0, n_edges, px, py);
}
int sgo_in_buffer(double x, double y, double r, double qx, double qy) {
  ngon_init();
  return in_buffer(x, y, r, qx, qy);
}
double sgo_route_project(const double* xy, int64_t R, double px, double py) {
  return route_project(xy, R, px, py);
}
