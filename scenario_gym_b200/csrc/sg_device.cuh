// sg_device.cuh -- device-side building blocks of the rollout engine (sm_100a).
//
// Everything is fp64 and keeps the reference's operation order; the translation unit
// is compiled with -fmad=false so a*b+c is never contracted (numpy has no FMA on this
// path).  fma() is used explicitly only inside the exact-arithmetic predicates, where
// it is an error-free transformation, not an approximation.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/sg_b200.h"

#define SG_DEV __device__ __forceinline__

// ----------------------------------------------------------------------------------
// Linear interpolation with scipy's interp1d._call_linear operation order (third-party
// arithmetic reached from reference trajectory.py:178-184, entity/batch.py:99-127):
//   w1 = (t - x_lo) / (x_hi - x_lo);  w0 = (x_hi - t) / (x_hi - x_lo);  y = w1*y_hi + w0*y_lo
// ----------------------------------------------------------------------------------
enum { EXT_NONE = 0, EXT_CLAMP = 1, EXT_TRUE = 2 };

// first index i in [0, K) with x[i*stride] >= t (numpy searchsorted side='left')
SG_DEV int search_left(const double* __restrict__ x, int stride, int K, double t) {
  int lo = 0, hi = K;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (__ldg(x + (int64_t)mid * stride) < t) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// Monotone cursor version: `cur` is the previous clipped index (in [1, K-1]); the tick
// times only grow, so the control-point search is O(1) amortised.  Falls back to a
// binary search when t moved backwards.
SG_DEV int search_left_cursor(const double* __restrict__ x, int stride, int K, double t, int cur) {
  if (cur < 1 || cur > K - 1 || __ldg(x + (int64_t)(cur - 1) * stride) >= t) {
    int i = search_left(x, stride, K, t);
    return min(max(i, 1), K - 1);
  }
  while (cur < K - 1 && __ldg(x + (int64_t)cur * stride) < t) ++cur;
  return cur;
}

// Trajectory.position_at_t for scalar t (reference trajectory.py:142-197) on the slot's
// own control points rows[K][7].  Returns false when the reference returns None.
SG_DEV bool position_at_t(const double* __restrict__ rows, int K, double t, int mode, int& cur,
                          double out[6]) {
  const double min_t = __ldg(rows), max_t = __ldg(rows + (int64_t)(K - 1) * 7);
  if (mode == EXT_NONE && (t < min_t || t > max_t)) return false;  // :191-192
  if (mode != EXT_TRUE && t < min_t) {                               // :193-194
#pragma unroll
    for (int f = 0; f < 6; ++f) out[f] = __ldg(rows + 1 + f);
    return true;
  }
  if (mode != EXT_TRUE && t > max_t) {  // :195-196
#pragma unroll
    for (int f = 0; f < 6; ++f) out[f] = __ldg(rows + (int64_t)(K - 1) * 7 + 1 + f);
    return true;
  }
  if (K == 1) {  // :175-177 single control point duplicated at t + 1e-3
    const double x_lo = min_t, x_hi = min_t + 1e-3;
    const double w1 = (t - x_lo) / (x_hi - x_lo), w0 = (x_hi - t) / (x_hi - x_lo);
#pragma unroll
    for (int f = 0; f < 6; ++f) {
      const double y = __ldg(rows + 1 + f);
      out[f] = w1 * y + w0 * y;
    }
    return true;
  }
  cur = search_left_cursor(rows, 7, K, t, cur);
  const double* lo = rows + (int64_t)(cur - 1) * 7;
  const double* hi = lo + 7;
  const double x_lo = __ldg(lo), x_hi = __ldg(hi);
  const double w1 = (t - x_lo) / (x_hi - x_lo), w0 = (x_hi - t) / (x_hi - x_lo);
#pragma unroll
  for (int f = 0; f < 6; ++f) out[f] = w1 * __ldg(hi + 1 + f) + w0 * __ldg(lo + 1 + f);
  return true;
}

// Trajectory.velocity_at_t (reference trajectory.py:243-273), eps = 1e-4
SG_DEV void velocity_at_t(const double* __restrict__ rows, int K, double t, double out[6]) {
  const double eps = 1e-4;
  const double min_t = __ldg(rows), max_t = __ldg(rows + (int64_t)(K - 1) * 7);
  const bool inside = (min_t <= t) && (t <= max_t);
  double a[6], b[6];
  int c0 = 0, c1 = 0;
  position_at_t(rows, K, t + eps / 2, EXT_TRUE, c0, a);
  position_at_t(rows, K, t - eps / 2, EXT_TRUE, c1, b);
#pragma unroll
  for (int f = 0; f < 6; ++f) out[f] = inside ? (a[f] - b[f]) / eps : 0.0;
}

// ----------------------------------------------------------------------------------
// Exact orientation sign (GEOS decides `intersects` with robust orientation predicates;
// Shapely call sites: reference utils.py:51-62, metrics/rss/callback.py:191-196,317-328).
// Static filter with Shewchuk's bound, exact expansion arithmetic behind it.
// ----------------------------------------------------------------------------------
SG_DEV void two_sum(double a, double b, double& s, double& e) {
  s = __dadd_rn(a, b);
  const double bv = __dsub_rn(s, a), av = __dsub_rn(s, bv);
  e = __dadd_rn(__dsub_rn(a, av), __dsub_rn(b, bv));
}
SG_DEV void two_prod(double a, double b, double& p, double& e) {
  p = __dmul_rn(a, b);
  e = __fma_rn(a, b, -p);
}

__device__ __noinline__ int orient_exact(double ax, double ay, double bx, double by, double cx,
                                         double cy) {
  double d[4][2];
  two_sum(ax, -cx, d[0][0], d[0][1]);
  two_sum(by, -cy, d[1][0], d[1][1]);
  two_sum(ay, -cy, d[2][0], d[2][1]);
  two_sum(bx, -cx, d[3][0], d[3][1]);
  double h[20];
  int n = 0;
  auto grow = [&](double b) {  // Shewchuk grow-expansion
    double q = b;
    for (int i = 0; i < n; ++i) {
      double s, e;
      two_sum(q, h[i], s, e);
      h[i] = e;
      q = s;
    }
    h[n++] = q;
  };
  for (int i = 0; i < 2; ++i)
    for (int j = 0; j < 2; ++j) {
      double p, e;
      two_prod(d[0][i], d[1][j], p, e);
      grow(e);
      grow(p);
      two_prod(d[2][i], d[3][j], p, e);
      grow(-e);
      grow(-p);
    }
  for (int i = n - 1; i >= 0; --i) {
    if (h[i] > 0) return 1;
    if (h[i] < 0) return -1;
  }
  return 0;
}

SG_DEV int orient_sign(double ax, double ay, double bx, double by, double cx, double cy) {
  const double errbound = (3.0 + 16.0 * 1.1102230246251565e-16) * 1.1102230246251565e-16;
  const double l = (ax - cx) * (by - cy), r = (ay - cy) * (bx - cx);
  const double det = l - r;
  if (fabs(det) > errbound * (fabs(l) + fabs(r))) return det > 0 ? 1 : -1;
  return orient_exact(ax, ay, bx, by, cx, cy);
}

// ring orientation of a convex quad q[8] = x0,y0,..,x3,y3 : +1 ccw, -1 cw, 0 degenerate
SG_DEV int quad_orientation(const double* q) {
  int s = orient_sign(q[0], q[1], q[2], q[3], q[4], q[5]);
  if (s == 0) s = orient_sign(q[2], q[3], q[4], q[5], q[6], q[7]);
  return s;
}

// true if all `npts` points are strictly outside edge k of the convex quad q (orientation o)
SG_DEV bool edge_separates(const double* q, int o, int k, const double* pts, int npts) {
  const double ax = q[2 * k], ay = q[2 * k + 1];
  const double bx = q[2 * ((k + 1) & 3)], by = q[2 * ((k + 1) & 3) + 1];
  for (int m = 0; m < npts; ++m)
    if (orient_sign(ax, ay, bx, by, pts[2 * m], pts[2 * m + 1]) * o >= 0) return false;
  return true;
}

// closed-set intersection of two convex quads (touching counts, as GEOS `intersects`)
SG_DEV bool quads_intersect(const double* a, int oa, const double* b, int ob) {
  for (int k = 0; k < 4; ++k)
    if (edge_separates(a, oa, k, b, 4)) return false;
  for (int k = 0; k < 4; ++k)
    if (edge_separates(b, ob, k, a, 4)) return false;
  return true;
}

// closed-set intersection of a convex quad and a segment s = x0,y0,x1,y1
SG_DEV bool quad_intersects_segment(const double* q, int o, const double* s) {
  for (int k = 0; k < 4; ++k)
    if (edge_separates(q, o, k, s, 2)) return false;
  int pos = 0, neg = 0;
  for (int m = 0; m < 4; ++m) {
    const int sg = orient_sign(s[0], s[1], s[2], s[3], q[2 * m], q[2 * m + 1]);
    pos += sg > 0;
    neg += sg < 0;
  }
  return !(pos == 4 || neg == 4);
}

// Entity.get_bounding_box_points (reference entity/base.py:100-138)
SG_DEV void box_points(double x, double y, double h, double W, double L, double cx, double cy,
                       double out[8]) {
  double s, c;
  sincos(h, &s, &c);
  const double hx0 = cx - 0.5 * L, hx1 = cx + 0.5 * L;
  const double hy0 = cy + 0.5 * W, hy1 = cy - 0.5 * W;
  const double px[4] = {hx0, hx1, hx1, hx0};
  const double py[4] = {hy0, hy0, hy1, hy1};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    out[2 * i] = x + (px[i] * c + py[i] * -s);
    out[2 * i + 1] = y + (px[i] * s + py[i] * c);
  }
}

SG_DEV double norm2(double a, double b) { return sqrt(a * a + b * b); }
SG_DEV double norm3(double a, double b, double c) { return sqrt(a * a + b * b + c * c); }
SG_DEV double dot2(double a0, double a1, double b0, double b1) { return a0 * b0 + a1 * b1; }
SG_DEV double py_max(double a, double b) { return b > a ? b : a; }
SG_DEV double py_min(double a, double b) { return b < a ? b : a; }
SG_DEV double np_sign(double v) { return v > 0 ? 1.0 : (v < 0 ? -1.0 : (v == 0 ? 0.0 : NAN)); }
SG_DEV double np_clip(double v, double lo, double hi) { return fmin(fmax(v, lo), hi); }

// ----------------------------------------------------------------------------------
// RSS (reference metrics/rss/callback.py, rss_utils.py)
// ----------------------------------------------------------------------------------
struct RssEnt {
  double position[2], heading[2], velocity[2], box[8], length, width;
};

SG_DEV void inverse_direction(const double v[2], double out[2]) {  // rss_utils.py:7-21
  const double n = norm2(v[1], v[0]);
  out[0] = v[1] / n;
  out[1] = -v[0] / n;
}
SG_DEV void coord_change(const double v[2], const double dir[2], const double c[2],
                         double out[2]) {  // rss_utils.py:24-45
  double inv[2];
  inverse_direction(dir, inv);
  const double d0 = v[0] - c[0], d1 = v[1] - c[1];
  out[0] = dot2(d0, d1, inv[0], inv[1]);
  out[1] = dot2(d0, d1, dir[0], dir[1]);
}
// callback.py:340-386; `pts` are the entity's world-frame corners
SG_DEV void rss_entity_params(double x, double y, double h, double vx, double vy,
                              const double* pts, double W, double L, const double eh[2],
                              const double einv[2], const double epos[2], RssEnt& o) {
  double s, c;
  sincos(h, &s, &c);
  const double dir[2] = {c, s};
  const double xy[2] = {x, y};
  coord_change(xy, eh, epos, o.position);
  o.heading[0] = dot2(dir[0], dir[1], einv[0], einv[1]);
  o.heading[1] = dot2(dir[0], dir[1], eh[0], eh[1]);
  o.velocity[0] = dot2(vx, vy, einv[0], einv[1]);
  o.velocity[1] = dot2(vx, vy, eh[0], eh[1]);
  for (int i = 0; i < 4; ++i) coord_change(pts + 2 * i, eh, epos, o.box + 2 * i);
  o.length = L;
  o.width = W;
}
SG_DEV double long_dist_same_direction(double vf, double vr, double a, double RT, double MINA) {
  const double v = vr * RT + py_min(vf * vf / (2 * a), 0.5 * a * (RT * RT)) +
                   ((vr + RT * a) * (vr + RT * a)) / (2 * MINA) - vf * vf / (2 * a);
  return py_max(0, v);  // callback.py:454-472
}
SG_DEV double long_dist_opp_direction(double v1, double v2, double a, double RT, double MINA) {
  const double av2 = fabs(v2);
  const double v = (2 * v1 + RT * a) * RT / 2 + ((v1 + RT * a) * (v1 + RT * a)) / (2 * MINA) +
                   (2 * av2 + RT * a) * RT / 2 + ((av2 + RT * a) * (av2 + RT * a)) / (2 * MINA);
  return py_max(0, v);  // callback.py:474-492
}
SG_DEV double lat_dist(double v, double amax, double amin, double RT) {  // callback.py:494-505
  const double x = 0.5 * RT * (2 * v + RT * amax) +
                   ((v + RT * amax) * (v + RT * amax)) / (2 * amin) - 0.5 * (RT * RT) * amax -
                   ((RT * amax) * (RT * amax)) / (2 * amin);
  return py_max(0, x);
}
SG_DEV double safe_longitudinal_distance(const SgParams& p, const RssEnt& ego, const RssEnt& haz) {
  const double CLR = p.rss_min_safe_clearance, RT = p.rss_response_time;  // callback.py:230-269
  const double dp = dot2(ego.heading[0], ego.heading[1], haz.heading[0], haz.heading[1]);
  const double a = fabs(p.rss_max_long_accel * dp);
  double d0;
  if (dp > 0) {
    double vf, vr;
    if (ego.position[1] > haz.position[1]) {
      vf = norm2(ego.velocity[0], ego.velocity[1]);
      vr = dot2(haz.velocity[0], haz.velocity[1], ego.heading[0], ego.heading[1]);
    } else {
      vf = dot2(haz.velocity[0], haz.velocity[1], ego.heading[0], ego.heading[1]);
      vr = norm2(ego.velocity[0], ego.velocity[1]);
    }
    if (vr == 0.0) return CLR + 0.5 * ego.length;
    d0 = long_dist_same_direction(vf, vr, a, RT, p.rss_min_long_accel);
  } else {
    const double v1 = fabs(dot2(ego.velocity[0], ego.velocity[1], ego.heading[0], ego.heading[1]));
    const double v2 = -fabs(dot2(haz.velocity[0], haz.velocity[1], ego.heading[0], ego.heading[1]));
    if (np_sign(haz.position[1]) == np_sign(haz.velocity[1])) return CLR + 0.5 * ego.length;
    d0 = long_dist_opp_direction(v1, v2, a, RT, p.rss_min_long_accel);
  }
  return d0 + CLR + 0.5 * ego.length;
}
SG_DEV double safe_lateral_distance(const SgParams& p, const RssEnt& ego, const RssEnt& haz) {
  const double CLR = p.rss_min_safe_clearance, RT = p.rss_response_time;  // callback.py:271-302
  double v = haz.velocity[0];
  double inv[2];
  inverse_direction(ego.heading, inv);
  const double k = fabs(dot2(inv[0], inv[1], haz.heading[0], haz.heading[1]));
  const double amax = p.rss_max_long_accel * k, amin = p.rss_min_long_accel * k;
  double d0;
  if (np_sign(-haz.position[0]) == np_sign(v)) {
    v = fabs(v);
    if (v == 0.0) return CLR + 0.5 * ego.width;
    d0 = lat_dist(v, amax, amin, RT);
  } else {
    d0 = 0;
  }
  return d0 + CLR + 0.5 * ego.width;
}
SG_DEV void safe_ratios(const RssEnt& ego, const RssEnt& haz, double out[2]) {  // callback.py:124-166
  const double safe_lat = 0.5 * ego.width, safe_long = 0.5 * ego.length;
  double inv[2];
  inverse_direction(haz.heading, inv);
  const double wl_inv = fabs(dot2(haz.width, haz.length, inv[0], inv[1]));
  const double wl_dir = fabs(dot2(haz.width, haz.length, haz.heading[0], haz.heading[1]));
  const double actual_lat = py_max(1e-6, fabs(haz.position[0]) - 0.5 * ego.width - 0.5 * wl_inv);
  const double actual_long = py_max(1e-6, fabs(haz.position[1]) - 0.5 * ego.length - 0.5 * wl_dir);
  out[0] = fabs(actual_lat / safe_lat);
  out[1] = fabs(actual_long / safe_long);
}
// callback.py:168-228 (+ :304-338 write_intersections, :388-452 generate_buffer)
__device__ __noinline__ int unsafe_distance(const RssEnt& ego, const RssEnt& haz, uint8_t& st,
                                            const double sd[2]) {
  if ((st >> 2) & 3) return SG_RSS_FOUND;
  const double slat = sd[0], slong = sd[1];
  const double buffer[8] = {slat, slong, -slat, slong, -slat, -slong, slat, -slong};
  const int oh = quad_orientation(haz.box), ob = quad_orientation(buffer);
  if (quads_intersect(haz.box, oh, buffer, ob)) {
    const int marker = st & 3;
    if (marker == 1) { st |= 2 << 2; return SG_RSS_UNSAFE_LONGITUDINAL; }
    if (marker == 2) { st |= 1 << 2; return SG_RSS_UNSAFE_LATERAL; }
    const double ed[2] = {ego.width, ego.length};
    double inv[2];
    inverse_direction(ed, inv);
    const double lhs =
        fabs(fabs(haz.position[0]) - fabs(dot2(haz.position[0], haz.position[1], ed[0], ed[1]))) / slat;
    const double rhs =
        fabs(fabs(haz.position[1] - dot2(haz.position[0], haz.position[1], inv[0], inv[1])) / slong);
    if (lhs > rhs) { st |= 2 << 2; return SG_RSS_UNSAFE_LONGITUDINAL; }
    st |= 1 << 2;
    return SG_RSS_UNSAFE_LATERAL;
  }
  const double len0[4] = {slat, 100 * slong, -slat, 100 * -slong};
  const double len1[4] = {-slat, 100 * slong, slat, 100 * -slong};
  const double wid0[4] = {100 * slat, slong, 100 * -slat, slong};
  const double wid1[4] = {100 * -slat, -slong, 100 * slat, -slong};
  const bool lat = quad_intersects_segment(haz.box, oh, len0) || quad_intersects_segment(haz.box, oh, len1);
  const bool lon = quad_intersects_segment(haz.box, oh, wid0) || quad_intersects_segment(haz.box, oh, wid1);
  if (lat && lon) return SG_RSS_BOTH;
  if (lat) { st = (uint8_t)((st & ~3) | 1); return SG_RSS_LATERAL; }
  if (lon) { st = (uint8_t)((st & ~3) | 2); return SG_RSS_LONGITUDINAL; }
  return SG_RSS_SAFE;
}

// ----------------------------------------------------------------------------------
// pedestrians
// ----------------------------------------------------------------------------------
// LineString(route).project(Point) (GEOS LengthIndexedLine semantics; reference
// pedestrian/agent.py:61)
SG_DEV double route_project(const double* __restrict__ xy, int R, double px, double py) {
  double best_d = INFINITY, best_s = 0.0, s0 = 0.0;
  for (int i = 0; i + 1 < R; ++i) {
    const double ax = __ldg(xy + 2 * i), ay = __ldg(xy + 2 * i + 1);
    const double bx = __ldg(xy + 2 * i + 2), by = __ldg(xy + 2 * i + 3);
    const double dx = bx - ax, dy = by - ay, seg2 = dx * dx + dy * dy, seglen = sqrt(seg2);
    const double r = seg2 == 0.0 ? 0.0 : ((px - ax) * dx + (py - ay) * dy) / seg2;
    double qx, qy, sl;
    if (r <= 0.0) { qx = ax; qy = ay; sl = 0.0; }
    else if (r >= 1.0) { qx = bx; qy = by; sl = seglen; }
    else { qx = ax + r * dx; qy = ay + r * dy; sl = r * seglen; }
    const double d = hypot(px - qx, py - qy);
    if (d < best_d) { best_d = d; best_s = s0 + sl; }
    s0 += seglen;
  }
  return best_s;
}
