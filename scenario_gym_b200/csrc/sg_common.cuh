// sg_common.cuh -- device code shared by the tick-loop kernels (sg_vehicle.cu, sg_general.cu, sg_crowd.cu):
// group descriptor, staging of boxes, broad / narrow phase, RSS, social force, tick epilogue.
//
//
// Mapping: one thread per entity slot; the G threads of a scenario ("group") are a
// sub-warp (M <= 32: G = next pow2, several scenarios per warp) or G/32 whole warps
// (M > 32: G = M rounded up to 32).  A 256-thread CTA carries 256/G scenarios.  Each
// thread keeps its entity's State row (pose, velocity, distance, controller speed, RSS
// history bits) in registers for all ticks of the call (n_ticks = 1 is
// ScenarioGym.step(); n_ticks < 0 is ScenarioGym.rollout()).  Per tick:
//   A  every thread produces its entity's new pose (replay interpolation / kinematic
//      bicycle / social force), updates velocity + distance, and stages its fp64 box
//      corners and a conservative fp32 AABB in shared memory;
//   B1 RSS per hazard against the ego published in shared memory; broad phase as a
//      circular half sweep: thread s tests slots s+1 .. s+M/2 (each unordered pair once,
//      branch-free bit accumulation, AABB array duplicated so the wrap is an immediate
//      offset) and pushes the rare AABB survivors to a per-scenario queue;
//   B2 the queue is drained cooperatively with the exact closed-set predicate on the
//      fp64 corners; hits set bits in the scenario's collision words;
//   C  terminal conditions, CollisionMetric rising edges (ego row bit words), ego metrics.
// Groups synchronise with __syncwarp(mask) (sub-warp) or a named barrier per scenario.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <string.h>

#include "sg_device.cuh"
#include "sg_layout.h"

// (cos, sin)(-k * 2pi/64): GEOS Point.buffer vertices
static __constant__ double c_ngon[64][2] = {
#include "sg_ngon.inc"
};

struct Grp {
  int n, s, M, G, W, H, QCAP, bar_id;
  unsigned mask;
  int64_t i, nm;
  double* corners;
  double* actbuf;
  double* rbox;
  unsigned corners_sh, rbox_sh;  // shared-window addresses for the ld.shared predicates
  double* tcold;
  double* hcs;
  double* pedbuf;
  float4* pednb;
  uint16_t* nblist;
  double* boxp;
  double* egop;
  double* cold_d;
  int* cold_i;
  float4* aabb;
  uint32_t* queue;
  uint32_t* ego_now;
  uint32_t* ego_last;
  uint32_t* bits;
  int* acc;
  uint8_t* flags;
  int8_t* orient;
  uint16_t* sid;
  uint16_t* posof;
  int* sflag;
  float* skey;
  float4* tmpbox;
  int sorted;
  uint32_t* gstart;
  uint16_t* gsorted;
  uint16_t* glarge;
  int* gmisc;
};

SG_DEV void cp_async8(unsigned smem_addr, const void* gsrc) {  // smem_addr: shared-window address
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_addr), "l"(gsrc) : "memory");
}
SG_DEV void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
SG_DEV void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

SG_DEV void group_sync(const Grp& g) {
  if (g.G <= 32) __syncwarp(g.mask);
  else asm volatile("bar.sync %0, %1;" ::"r"(g.bar_id), "r"(g.G) : "memory");
}

SG_DEV void setup_group(Grp& g, const SgScene& sc, const GroupLayout& L, unsigned char* smem,
                        int gl, int s, int n, int M = -1) {
  if (M < 0) M = sc.n_slots;  // (kernels specialised for a slot count pass it as a constant)
  unsigned char* base = smem + (size_t)gl * L.bytes;
  g.n = n; g.s = s; g.M = M; g.G = L.G; g.W = L.W; g.H = L.H; g.QCAP = L.QCAP;
  g.bar_id = 1 + gl;
  g.mask = 0xffffffffu;
  if (L.G < 32) g.mask = ((1u << L.G) - 1u) << ((threadIdx.x & 31) / L.G * L.G);
  g.nm = sc.plane_stride;  // (sg_api.cu fills it in for whole batches)
  g.i = (int64_t)n * M + s;
  g.corners = (double*)base;
  g.actbuf = (double*)(base + L.off_act);
  g.rbox = (double*)(base + L.off_rbox);
  g.corners_sh = (unsigned)__cvta_generic_to_shared(base);
  g.rbox_sh = g.corners_sh + (unsigned)L.off_rbox;
  g.tcold = (double*)(base + L.off_tcold);
  g.hcs = (double*)(base + L.off_hcs);
  g.pedbuf = (double*)(base + L.off_ped);
  g.pednb = (float4*)(base + L.off_pednb);
  g.nblist = (uint16_t*)(base + L.off_nbl);
  g.boxp = (double*)(base + L.off_box);
  g.egop = (double*)(base + L.off_ego);
  g.cold_d = (double*)(base + L.off_cold);
  g.cold_i = (int*)(base + L.off_cold + COLD_ND * sizeof(double));
  g.aabb = (float4*)(base + L.off_aabb);
  g.queue = (uint32_t*)(base + L.off_queue);
  g.ego_now = (uint32_t*)(base + L.off_hits);
  g.ego_last = g.ego_now + L.W;
  g.bits = (uint32_t*)(base + L.off_bits);
  g.acc = (int*)(base + L.off_acc);
  g.flags = (uint8_t*)(base + L.off_flags);
  g.orient = (int8_t*)(base + L.off_orient);
  g.sid = (uint16_t*)(base + L.off_sid);
  g.posof = (uint16_t*)(base + L.off_posof);
  g.sflag = (int*)(base + L.off_sflag);
  g.skey = (float*)(base + L.off_skey) + SG_SORT_WIN + 1;  // index -WIN-1 .. M+WIN are valid
  g.tmpbox = (float4*)(base + L.off_tmpbox);
  g.sorted = L.sorted;
  g.gstart = (uint32_t*)(base + L.off_gstart);
  g.gsorted = (uint16_t*)(base + L.off_gsorted);
  g.glarge = (uint16_t*)(base + L.off_glarge);
  g.gmisc = (int*)(base + L.off_gmisc);
}

// 1/d for a finite, normal, non-zero d: hardware seed (rcp.approx.ftz.f64, ~2^-23) refined by
// two Newton steps to within an ulp; no special-case paths, ~6 instructions instead of ~25
SG_DEV double fast_rcp(double d) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
  double e = __fma_rn(-d, r, 1.0);
  r = __fma_rn(r, e, r);
  e = __fma_rn(-d, r, 1.0);
  return __fma_rn(r, e, r);
}

// sqrt(x) for finite x >= 0: hardware rsqrt seed, two Newton steps on 1/sqrt and one residual
// correction on the root (within an ulp); ~12 instructions, no special-case paths
SG_DEV double fast_sqrt(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double hx = 0.5 * x;
  double e = __fma_rn(-hx * y, y, 0.5);
  y = __fma_rn(y, e, y);
  e = __fma_rn(-hx * y, y, 0.5);
  y = __fma_rn(y, e, y);
  double r = x * y;
  r = __fma_rn(__fma_rn(-r, r, x), 0.5 * y, r);
  return x > 0.0 ? r : 0.0;
}
SG_DEV double fnorm2(double a, double b) { return fast_sqrt(a * a + b * b); }

// n / d with a shared reciprocal r = 1/d (correctly rounded): one Newton correction on the
// quotient (Markstein); equals the IEEE quotient except in rare last-bit cases
SG_DEV double div_r(double n, double d, double r) {
  const double q = n * r;
  const double rem = __fma_rn(-q, d, n);
  return __fma_rn(rem, r, q);
}

// Conservative fp32 AABB (relative to the scenario origin) of the box whose fp64 corners
// publish_box stages: centre +- (|L/2 c| + |W/2 s|, |L/2 s| + |W/2 c|), widened by 32 ulp of the
// coordinate magnitudes (covers the rounding of the reference's corner formula, ~6 ulp) and
// rounded outwards at every step.  Costs ~20 flops instead of 12 fp64 min/max sequences.
SG_DEV float4 make_aabb_box(double x, double y, double cs, double sn, double bw, double bl,
                            double bcx, double bcy, double ox, double oy) {
  const double xc = x + (bcx * cs - bcy * sn), yc = y + (bcx * sn + bcy * cs);
  const double hl = 0.5 * bl, hw = 0.5 * bw;
  const double m = 3.552713678800501e-15 *
                   (fabs(x) + fabs(y) + fabs(bl) + fabs(bw) + fabs(bcx) + fabs(bcy));
  const double ex = fabs(hl * cs) + fabs(hw * sn) + m, ey = fabs(hl * sn) + fabs(hw * cs) + m;
  float4 b;
  b.x = __double2float_rd(__dsub_rd(__dsub_rd(xc, ex), ox));
  b.y = __double2float_rd(__dsub_rd(__dsub_rd(yc, ey), oy));
  b.z = __double2float_ru(__dsub_ru(__dadd_ru(xc, ex), ox));
  b.w = __double2float_ru(__dsub_ru(__dadd_ru(yc, ey), oy));
  return b;
}

SG_DEV double min2(double a, double b) { return a < b ? a : b; }
SG_DEV double max2(double a, double b) { return a > b ? a : b; }
SG_DEV double clipd(double v, double lo, double hi) { return v < lo ? lo : (v > hi ? hi : v); }

// sin / cos kernels on |x| <= pi/4 (fdlibm k_sin.c / k_cos.c minimax polynomials, < 1 ulp),
// coefficients read from the constant bank instead of 64-bit immediates
static __constant__ double c_trig[12] = {
    -1.66666666666666324348e-01, 8.33333333332248946124e-03,  -1.98412698298579493134e-04,
    2.75573137070700676789e-06,  -2.50507602534068634195e-08, 1.58969099521155010221e-10,
    4.16666666666666019037e-02,  -1.38888888888741095749e-03, 2.48015872894767294178e-05,
    -2.75573143513906633035e-07, 2.08757232129817482790e-09,  -1.13596475577881948265e-11};
SG_DEV void sincos_kernel(double x, double& sn, double& cs) {
  const double z = x * x;
  const double* S = c_trig;
  const double* C = c_trig + 6;
  double r = __fma_rn(z, S[5], S[4]);
  r = __fma_rn(z, r, S[3]);
  r = __fma_rn(z, r, S[2]);
  r = __fma_rn(z, r, S[1]);
  r = __fma_rn(z, r, S[0]);
  sn = __fma_rn(x * z, r, x);
  double q = __fma_rn(z, C[5], C[4]);
  q = __fma_rn(z, q, C[3]);
  q = __fma_rn(z, q, C[2]);
  q = __fma_rn(z, q, C[1]);
  q = __fma_rn(z, q, C[0]);
  const double hz = 0.5 * z, w = 1.0 - hz;
  cs = w + (((1.0 - w) - hz) + z * (z * q));
}
// library fall-backs kept out of line so the tick loop does not carry their code
static __device__ __noinline__ double2 sincos_lib(double x) {  // by value: no address-taken locals in the callers
  double2 r;
  sincos(x, &r.x, &r.y);
  return r;
}
static __device__ __noinline__ double tan_lib(double x) { return tan(x); }

// tan on |x| <= pi/4 (steering angles are clipped to +-max_steer)
SG_DEV double tan_small(double x) {
  double sn, cs;
  sincos_kernel(x, sn, cs);
  return div_r(sn, cs, fast_rcp(cs));
}
// sincos with a 3-term Cody-Waite reduction (exact under FMA for |x| < 1e9)
SG_DEV void sincos_fast(double x, double& sn, double& cs) {
  if (!(fabs(x) < 1.0e9)) { const double2 r = sincos_lib(x); sn = r.x; cs = r.y; return; }
  const double kd = rint(x * 0.63661977236758138);
  double r = __fma_rn(-kd, 1.5707963267948966, x);
  r = __fma_rn(-kd, 6.123233995736766e-17, r);
  r = __fma_rn(-kd, -1.4973849048591698e-33, r);
  double a, b;
  sincos_kernel(r, a, b);
  const int k = (int)kd;
  const double s0 = (k & 1) ? b : a, c0 = (k & 1) ? a : b;
  sn = (k & 2) ? -s0 : s0;
  cs = ((k + 1) & 2) ? -c0 : c0;
}

// strict interior of Point(x, y).buffer(r): GEOS 64-gon (reference state/state.py:352-372).
// Between the inscribed and the circumscribed circle the polygon decides: the query point lies in
// the angular sector of one edge, and only that edge (and, against rounding of the sector index, its
// two neighbours) can have the point on its outer side -- every other edge's line is further than
// 0.009 r away from any point of the band in this sector.  The three edges are tested with the exact
// orientation sign, so the answer equals the all-edges test.
static __device__ __noinline__ bool in_buffer_band(double x, double y, double r, double qx, double qy) {
  const double a = atan2(-(qy - y), qx - x);  // clockwise angle: vertex k sits at k * 2 pi / 64
  double kk = a * (64.0 / (2.0 * M_PI));
  if (kk < 0.0) kk += 64.0;
  const int k0 = (int)kk;
#pragma unroll 1
  for (int j = -1; j <= 1; ++j) {
    const int k = (k0 + j) & 63, k1 = (k + 1) & 63;
    const double ax = x + r * c_ngon[k][0], ay = y + r * c_ngon[k][1];
    const double bx = x + r * c_ngon[k1][0], by = y + r * c_ngon[k1][1];
    if (orient_sign(ax, ay, bx, by, qx, qy) >= 0) return false;
  }
  return true;
}
SG_DEV bool in_buffer(double x, double y, double r, double qx, double qy) {
  const double dx = qx - x, dy = qy - y, d2 = dx * dx + dy * dy;
  const double rin = r * 0.99879545620517241 * (1.0 - 1e-9);  // cos(pi/64): inscribed circle
  if (d2 < rin * rin) return true;
  const double rout = r * (1.0 + 1e-9);
  if (d2 > rout * rout) return false;
  return in_buffer_band(x, y, r, qx, qy);
}

// closed-interval overlap of two conservative AABBs as one predicate chain; sets bit `bit`
#define SG_AABB_TEST(hits, mb, ob, bit)                                                         \
  asm("{ .reg .pred p;\n\t"                                                                     \
      "setp.le.f32 p, %1, %2;\n\t"                                                              \
      "setp.le.and.f32 p, %3, %4, p;\n\t"                                                       \
      "setp.le.and.f32 p, %5, %6, p;\n\t"                                                       \
      "setp.le.and.f32 p, %7, %8, p;\n\t"                                                       \
      "@p or.b32 %0, %0, %9; }"                                                                  \
      : "+r"(hits)                                                                              \
      : "f"(mb.x), "f"(ob.z), "f"(ob.x), "f"(mb.z), "f"(mb.y), "f"(ob.w), "f"(ob.y), "f"(mb.w), \
        "r"(bit))

// old pose / velocity of every slot as seen by the pedestrians' sensors in the next tick, plus a
// conservative fp32 box of half-size r/2 around pedestrians: two pedestrians can only be within
// the sensor's 64-gon (circumradius r) if those boxes overlap
SG_DEV void stage_ped_state(const Grp& c, bool present, int etype, double x, double y, double vx,
                            double vy, double r, double ox, double oy) {
  c.flags[c.s] = (uint8_t)((present ? 1 : 0) | (etype << 1));
  c.pedbuf[c.s] = x; c.pedbuf[c.G + c.s] = y;
  c.pedbuf[2 * c.G + c.s] = vx; c.pedbuf[3 * c.G + c.s] = vy;
  float4 b = make_float4(INFINITY, INFINITY, -INFINITY, -INFINITY);
  if (present && etype == SG_ETYPE_PEDESTRIAN) {
    const double h = 0.5 * r * (1.0 + 1e-9);
    b.x = __double2float_rd(__dsub_rd(__dsub_rd(x, h), ox));
    b.y = __double2float_rd(__dsub_rd(__dsub_rd(y, h), oy));
    b.z = __double2float_ru(__dsub_ru(__dadd_ru(x, h), ox));
    b.w = __double2float_ru(__dsub_ru(__dadd_ru(y, h), oy));
  }
  c.pednb[c.s] = b;
}


// ---------------------------------------------------------------------------------
// Cell grid of a crowd scenario (one CTA per scenario, G > 256).  Every present entity is binned
// by its pose position into a 64 x 64 toroidal grid of cell size cs = r (1 + 1e-6):
//   * pedestrian sensors: a neighbour strictly inside the 64-gon of circumradius r lies in the
//     3 x 3 cells around the pedestrian;
//   * broad phase: entities whose AABB reaches at most cs/2 from their position ("small") can
//     only overlap small entities in the 3 x 3 cells around them; the few large ones are kept
//     in a list that every entity tests against.
// Aliasing through the wrap only adds candidates, and the fp32 box tests applied to the
// candidates are the ones the exhaustive sweeps use, so the outcomes are identical.  O(M) per
// tick instead of O(M^2).
// ---------------------------------------------------------------------------------
struct GridPos {
  int ix, iy;
  bool large;
};

SG_DEV uint32_t grid_start(const uint32_t* gs, int cell) {  // cell in [0, SG_GRID_CELLS]
  const uint32_t w = gs[cell >> 1];
  return (cell & 1) ? (w >> 16) : (w & 0xffffu);
}

// up to two index ranges of gsorted that hold the cells (ix-1 .. ix+1, row iy + dy)
SG_DEV int grid_row_ranges(const uint32_t* gs, int ix, int iy, int dy, int beg[2], int end[2]) {
  const int row = ((iy + dy) & (SG_GRID_DIM - 1)) << SG_GRID_BITS;
  const int a = (ix - 1) & (SG_GRID_DIM - 1), b = (ix + 1) & (SG_GRID_DIM - 1);
  if (a < b) {
    beg[0] = (int)grid_start(gs, row | a); end[0] = (int)grid_start(gs, (row | b) + 1);
    return 1;
  }
  beg[0] = (int)grid_start(gs, row | a); end[0] = (int)grid_start(gs, row + SG_GRID_DIM);
  beg[1] = (int)grid_start(gs, row);     end[1] = (int)grid_start(gs, (row | b) + 1);
  return 2;
}

// All G threads of the scenario's CTA call this (it synchronises the group).
SG_DEV GridPos grid_build(const Grp& c, bool present, bool have_box, double x, double y, double ox,
                          double oy, double cs, double inv_cs) {
  uint32_t* gs = c.gstart;
  const int nwords = SG_GRID_CELLS / 2;  // two 16-bit cells per word; word nwords holds the end
  for (int q = c.s; q <= nwords; q += c.G) gs[q] = 0;
  if (c.s == 0) c.gmisc[0] = 0;
  group_sync(c);
  GridPos g;
  g.ix = 0; g.iy = 0; g.large = false;
  int cell = 0;
  uint32_t rank = 0;
  if (present) {
    const double px = x - ox, py = y - oy;
    g.ix = __double2int_rd(px * inv_cs);
    g.iy = __double2int_rd(py * inv_cs);
    cell = ((g.iy & (SG_GRID_DIM - 1)) << SG_GRID_BITS) | (g.ix & (SG_GRID_DIM - 1));
    if (have_box) {
      const float4 bb = c.aabb[c.s];
      const double reach = fmax(fmax(px - (double)bb.x, (double)bb.z - px),
                                fmax(py - (double)bb.y, (double)bb.w - py));
      g.large = !(reach <= 0.5 * cs * (1.0 - 1e-6));
      if (g.large) {
        const int k = atomicAdd(&c.gmisc[0], 1);
        if (k < SG_GRID_LCAP) c.glarge[k] = (uint16_t)c.s;
      }
    }
    const uint32_t old = atomicAdd(&gs[cell >> 1], (cell & 1) ? 0x10000u : 1u);
    rank = (cell & 1) ? (old >> 16) : (old & 0xffffu);
  }
  group_sync(c);
  // exclusive scan of the packed counts: thread s owns wpt consecutive words
  const int wpt = (nwords + c.G - 1) / c.G;
  const int w0 = c.s * wpt;
  uint32_t local = 0;
  for (int q = 0; q < wpt; ++q)
    if (w0 + q < nwords) { const uint32_t v = gs[w0 + q]; local += (v & 0xffffu) + (v >> 16); }
  const int lane = threadIdx.x & 31, wid = c.s >> 5;
  uint32_t incl = local;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += o;
  }
  if (lane == 31) c.gmisc[8 + wid] = (int)incl;
  group_sync(c);
  uint32_t wt = lane < (c.G >> 5) ? (uint32_t)c.gmisc[8 + lane] : 0u, wi = wt;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t o = __shfl_up_sync(0xffffffffu, wi, d);
    if (lane >= d) wi += o;
  }
  const uint32_t wbase = __shfl_sync(0xffffffffu, wi - wt, wid);
  const uint32_t total = __shfl_sync(0xffffffffu, wi, 31);
  uint32_t run = wbase + incl - local;
  for (int q = 0; q < wpt; ++q)
    if (w0 + q < nwords) {
      const uint32_t v = gs[w0 + q], lo = v & 0xffffu, hi = v >> 16;
      gs[w0 + q] = run | ((run + lo) << 16);
      run += lo + hi;
    }
  if (c.s == 0) gs[nwords] = total;
  group_sync(c);
  if (present)
    c.gsorted[grid_start(gs, cell) + rank] = (uint16_t)((uint32_t)c.s | (g.large ? SG_GRID_LARGE : 0u));
  group_sync(c);
  return g;
}

// broad phase through the grid: every unordered pair whose conservative AABBs overlap is queued
// exactly once (small-small by the lower slot, small-large by the small one, large-large by the
// lower slot).  Falls back to the exhaustive half sweep when the large list overflowed.
SG_DEV void broad_phase(const Grp& c, int parity);
SG_DEV void broad_phase_grid(const Grp& c, int parity, const GridPos& g) {
  const int nl = c.gmisc[0];
  if (nl > SG_GRID_LCAP) { broad_phase(c, parity); return; }
  const float4 mb = c.aabb[c.s];
  int* acc = c.acc + parity * ACC_N;
  auto test = [&](int o) {
    const float4 ob = c.aabb[o];
    if (mb.x <= ob.z && ob.x <= mb.z && mb.y <= ob.w && ob.y <= mb.w) {
      const int q = atomicAdd(&acc[ACC_QCOUNT], 1);
      if (q < c.QCAP) c.queue[q] = ((uint32_t)c.s << 16) | (uint32_t)o;
    }
  };
  if (!g.large) {
    for (int dy = -1; dy <= 1; ++dy) {
      int beg[2], end[2];
      const int nr = grid_row_ranges(c.gstart, g.ix, g.iy, dy, beg, end);
      for (int r = 0; r < nr; ++r)
        for (int idx = beg[r]; idx < end[r]; ++idx) {
          const uint32_t o = c.gsorted[idx];
          if (o > (uint32_t)c.s && o < SG_GRID_LARGE) test((int)o);
        }
    }
  }
  for (int k = 0; k < nl; ++k) {
    const int o = c.glarge[k];
    if (o == c.s || (g.large && o < c.s)) continue;
    test(o);
  }
}

// Contribution of neighbour slot o to the pedestrian standing at (px, py), as the two vector terms
// SocialForce adds for it in application order (reference social_force.py:51-80, 140-188,
// 213-222): with sight weights w_r F_rep then w_a F_att, without them F_att then F_rep.
struct NbForce {
  int valid;
  double a0, a1, b0, b1;
};
SG_DEV NbForce neighbour_force(const SgParams& p, const Grp& c, double px, double py, int o, double step_dt,
                               double sh, double ch, double sight_cos) {
  NbForce f;
  f.valid = 0; f.a0 = 0.0; f.a1 = 0.0; f.b0 = 0.0; f.b1 = 0.0;
  const double* bx = c.pedbuf;
  const double* by = c.pedbuf + c.G;
  const double* bvx = c.pedbuf + 2 * c.G;
  const double* bvy = c.pedbuf + 3 * c.G;
  const double ox = bx[o], oy = by[o];
  if (!in_buffer(px, py, p.ped_distance_threshold, ox, oy)) return f;
  f.valid = 1;
  const double ovx = bvx[o], ovy = bvy[o];
  // (tolerance path: square roots and quotients through the rsqrt / rcp seeded helpers, < 1 ulp)
  const double vdx = ovx * ch + ovy * -sh, vdy = ovx * sh + ovy * ch;
  const double vn = fnorm2(vdx, vdy) + 0.0000000001, rvn = fast_rcp(vn);
  const double view0 = div_r(vdx, vn, rvn), view1 = div_r(vdy, vn, rvn);
  // _force_pedestrian_repulsion :140-176
  const double rx = px - ox, ry = py - oy, rn = fnorm2(rx, ry), rrn = fast_rcp(rn);
  const double vmag = fnorm2(ovx, ovy) + 0.0000000001, rvm = fast_rcp(vmag);
  const double uox = div_r(ovx, vmag, rvm), uoy = div_r(ovy, vmag, rvm);
  const double other_step = vmag * step_dt;
  const double r2x = rx - other_step * uox, r2y = ry - other_step * uoy;
  const double r2n = fnorm2(r2x, r2y) + 0.0000000001, rr2 = fast_rcp(r2n);
  const double b = (1.0 / 2) * fast_sqrt((rn + r2n) * (rn + r2n) - other_step * other_step);
  const double c0 = (1.0 / 4) * fast_rcp(b) * (rn + r2n);
  const double dbx = c0 * (div_r(rx, rn, rrn) + div_r(r2x, r2n, rr2)),
               dby = c0 * (div_r(ry, rn, rrn) + div_r(r2y, r2n, rr2));
  const double g = p.sf_ped_repulse_V / p.sf_ped_repulse_sigma * exp(-b / p.sf_ped_repulse_sigma);
  const double Fr0 = g * dbx, Fr1 = g * dby;
  const double Fa0 = 2 * p.sf_ped_attract_C * rx, Fa1 = 2 * p.sf_ped_attract_C * ry;
  if (p.sf_sight_weight_use) {  // _sight_weight :213-222
    const double nr = fnorm2(Fr0, Fr1) + 0.0000000001, na = fnorm2(Fa0, Fa1) + 0.0000000001;
    double dd = div_r(dot2(view0, view1, Fr0, Fr1), nr, fast_rcp(nr));
    double w = dd >= sight_cos ? 1.0 : p.sf_sight_weight;
    f.a0 = w * Fr0; f.a1 = w * Fr1;
    dd = div_r(dot2(view0, view1, Fa0, Fa1), na, fast_rcp(na));
    w = dd >= sight_cos ? 1.0 : p.sf_sight_weight;
    f.b0 = w * Fa0; f.b1 = w * Fa1;
  } else {
    f.a0 = Fa0; f.a1 = Fa1;
    f.b0 = Fr0; f.b1 = Fr1;
  }
  return f;
}

// PedestrianAgent._step + SocialForce._step + PedestrianController._step.  Called by every lane
// (is_ped = this lane steps a pedestrian this tick).  With `coop` (whole warps per scenario) the
// (pedestrian, neighbour) terms of a warp are pooled and spread evenly over its 32 lanes - a lane
// no longer waits for the pedestrian with the most neighbours - and handed back to their owners
// with shuffles in list order, so every force is the same sum in the same order.
template <bool PED>
SG_DEV void pedestrian_step(const SgScene& sc, const SgParams& p, const Grp& c, bool is_ped, bool coop,
                            const double pose[6], const double vel[6], double t, double prev_t,
                            double next_t, double sight_cos, int& goal, double force[2],
                            double& speed_io, double out[6], bool use_grid, double ox, double oy,
                            double inv_cs, int tick) {
  if (!PED) return;
  const double* route = nullptr;
  int R = 0, ncand = 0;
  bool walking = false;
  double F0 = 0.0, F1 = 0.0, speed_desired = 0.0;
  double sh, ch;
  sincos(p.ped_head_rot_angle, &sh, &ch);  // viewer/utils.py:6-17
  const float4 mb = c.pednb[c.s];
  uint16_t* nbl = c.nblist + c.s;
  if (is_ped) {
    const int64_t r0 = sc.route_off[c.i];
    R = (int)(sc.route_off[c.i + 1] - r0);
    route = sc.route_xy + 2 * r0;
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
    walking = goal <= R - 1;
  }
  if (walking) {
    speed_desired = sc.ped_speed_desired[c.i];
    // SocialForce._force_to_goal, pedestrian/social_force.py:119-138
    const double dvx = __ldg(route + 2 * goal) - pose[0], dvy = __ldg(route + 2 * goal + 1) - pose[1];
    double dn = norm2(dvx, dvy);
    if (dn == 0) dn += 0.000000001;
    const double ux = dvx / dn, uy = dvy / dn;
    const double k = 1 / p.sf_relaxation_time;
    F0 = k * (speed_desired * ux - vel[0]);
    F1 = k * (speed_desired * uy - vel[1]);
    // sensor: fp32 box prefilter; candidates are collected per pedestrian in slot (= state.poses) order
    if (use_grid) {  // from the 3 x 3 cells around the pedestrian, then put in slot order
      const int ix = __double2int_rd((pose[0] - ox) * inv_cs), iy = __double2int_rd((pose[1] - oy) * inv_cs);
      for (int dy = -1; dy <= 1; ++dy) {
        int beg[2], end[2];
        const int nr = grid_row_ranges(c.gstart, ix, iy, dy, beg, end);
        for (int r = 0; r < nr; ++r)
          for (int idx = beg[r]; idx < end[r]; ++idx) {
            const int o = (int)(c.gsorted[idx] & (SG_GRID_LARGE - 1u));
            const float4 ob = c.pednb[o];
            if (o == c.s || !(mb.x <= ob.z && ob.x <= mb.z && mb.y <= ob.w && ob.y <= mb.w)) continue;
            if (ncand < SG_NBCAP) nbl[ncand * c.G] = (uint16_t)o;
            ++ncand;
          }
      }
      for (int a = 1; a < min(ncand, SG_NBCAP); ++a) {  // insertion sort (a handful of entries)
        const uint16_t v = nbl[a * c.G];
        int b = a - 1;
        while (b >= 0 && nbl[b * c.G] > v) { nbl[(b + 1) * c.G] = nbl[b * c.G]; --b; }
        nbl[(b + 1) * c.G] = v;
      }
    } else {
      for (int o0 = 0; o0 < c.M; o0 += 32) {
        uint32_t cand = 0;
#pragma unroll
        for (int oo = 0; oo < 32; ++oo) {  // pednb is padded with empty boxes beyond M
          const float4 ob = c.pednb[o0 + oo];
          SG_AABB_TEST(cand, mb, ob, 1u << oo);
        }
        if (c.s >= o0 && c.s < o0 + 32) cand &= ~(1u << (c.s - o0));
        while (cand) {
          const int o = o0 + __ffs(cand) - 1;
          cand &= cand - 1;
          if (ncand < SG_NBCAP) nbl[ncand * c.G] = (uint16_t)o;
          ++ncand;
        }
      }
    }
  }
  // neighbour terms of the listed candidates (grid mode: a list that overflowed is not used)
  const int nlist = !walking ? 0 : (use_grid ? (ncand <= SG_NBCAP ? ncand : 0) : min(ncand, SG_NBCAP));
  const double step_dt = next_t - t;
  if (coop) {
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, wslot0 = c.s - lane;
    int incl = nlist;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int v = __shfl_up_sync(FULL, incl, d);
      if (lane >= d) incl += v;
    }
    const int off = incl - nlist, total = __shfl_sync(FULL, incl, 31);
    __syncwarp();  // the candidate lists of the warp's lanes are complete
    for (int base = 0; base < total; base += 32) {
      const int it = base + lane;
      int q = 0;  // owner lane of item `it`: the first lane whose inclusive count exceeds it
#pragma unroll
      for (int stp = 16; stp > 0; stp >>= 1) {
        const int tq = __shfl_sync(FULL, incl, (q + stp - 1) & 31);
        if (tq <= it) q += stp;
      }
      q &= 31;
      const int offq = __shfl_sync(FULL, off, q);
      NbForce nf;
      nf.valid = 0; nf.a0 = 0.0; nf.a1 = 0.0; nf.b0 = 0.0; nf.b1 = 0.0;
      if (it < total) {
        const int slot = wslot0 + q;
        const int o = c.nblist[(it - offq) * c.G + slot];
        nf = neighbour_force(p, c, c.pedbuf[slot], c.pedbuf[c.G + slot], o, step_dt, sh, ch, sight_cos);
      }
      // hand the terms back: this lane owns items [lo, hi) of the round, in list order
      const int lo = max(off, base), hi = min(off + nlist, base + 32);
      const int cnt = max(hi - lo, 0);
      const int maxc = __reduce_max_sync(FULL, cnt);
      for (int jj = 0; jj < maxc; ++jj) {
        const int src = (lo - base + jj) & 31;
        const int v = __shfl_sync(FULL, nf.valid, src);
        const double a0 = __shfl_sync(FULL, nf.a0, src), a1 = __shfl_sync(FULL, nf.a1, src);
        const double b0 = __shfl_sync(FULL, nf.b0, src), b1 = __shfl_sync(FULL, nf.b1, src);
        if (jj < cnt && v) { F0 += a0; F1 += a1; F0 += b0; F1 += b1; }
      }
    }
  } else {
    for (int k = 0; k < nlist; ++k) {
      const NbForce nf = neighbour_force(p, c, pose[0], pose[1], nbl[k * c.G], step_dt, sh, ch, sight_cos);
      if (nf.valid) { F0 += nf.a0; F1 += nf.a1; F0 += nf.b0; F1 += nf.b1; }
    }
  }
  if (walking && ncand > SG_NBCAP) {  // very dense crowd: the candidates beyond / instead of the list
    int seen = 0;
    for (int o = 0; o < c.M; ++o) {
      const float4 ob = c.pednb[o];
      if (o == c.s || !(mb.x <= ob.z && ob.x <= mb.z && mb.y <= ob.w && ob.y <= mb.w)) continue;
      if (use_grid || seen++ >= SG_NBCAP) {
        const NbForce nf = neighbour_force(p, c, pose[0], pose[1], o, step_dt, sh, ch, sight_cos);
        if (nf.valid) { F0 += nf.a0; F1 += nf.a1; F0 += nf.b0; F1 += nf.b1; }
      }
    }
  }
  if (!is_ped) return;
  if (walking && sc.n_networks > 0) {  // pedestrian/social_force.py:83-104
    if (surface_has_area(sc, c.n, 1) && surface_contains(sc, c.n, 1, pose[0], pose[1])) {
      double fb[2];
      boundary_force(sc, c.n, 1, pose[0], pose[1], p.sf_boundary_repulse_U, p.sf_boundary_repulse_R, fb);
      F0 += fb[0]; F1 += fb[1];
    }
    if (surface_has_area(sc, c.n, 2)) {
      const double sign = 1 - 2 * (surface_contains(sc, c.n, 2, pose[0], pose[1]) ? 1 : 0);
      double fb[2];
      boundary_force(sc, c.n, 2, pose[0], pose[1], p.sf_imp_boundary_repulse_U, p.sf_imp_boundary_repulse_R, fb);
      F0 += sign * fb[0]; F1 += sign * fb[1];
    }
  }
  double speed, heading;
  if (walking) {
    double speed_rand = p.sf_bias_lon, heading_rand = p.sf_bias_lat;  // np.random.normal(bias, 0) == bias
    if (p.sf_std_lon != 0.0 || p.sf_std_lat != 0.0) {
      const double2 z = sg_noise2(p.sf_noise_seed, c.i + (int64_t)sc.scenario_base * sc.n_slots, tick);
      speed_rand = p.sf_bias_lon + p.sf_std_lon * z.x;
      heading_rand = p.sf_bias_lat + p.sf_std_lat * z.y;
    }
    speed = py_min(norm2(F0, F1) + speed_rand, speed_desired * p.sf_max_speed_factor);
    heading = atan2(F1, F0) + heading_rand;
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

// ---------------------------------------------------------------------------------
// staging
// ---------------------------------------------------------------------------------
// box ring orientation is invariant under the rigid motion of entity/base.py:100-138:
// the local ring (-,+),(+,+),(+,-),(-,-) is clockwise for W*L > 0.  Zero-area boxes are
// decided exactly per tick.
SG_DEV int box_orientation_hint(double bw, double bl) {
  const double a = bw * bl;
  return a > 0 ? -1 : (a < 0 ? 1 : 0);
}

// Entity.get_bounding_box_points (reference entity/base.py:100-138) from cos/sin of the heading
template <bool RSS, bool SORTED = false>
SG_DEV void publish_box(const Grp& c, bool present, double x, double y, double cs, double sn,
                        int orient_hint, double ox, double oy) {
  float4 bb = make_float4(INFINITY, INFINITY, -INFINITY, -INFINITY);
  if (present) {
    const double bw = c.boxp[c.s], bl = c.boxp[c.G + c.s];
    const double bcx = c.boxp[2 * c.G + c.s], bcy = c.boxp[3 * c.G + c.s];
    const double hx0 = bcx - 0.5 * bl, hx1 = bcx + 0.5 * bl;
    const double hy0 = bcy + 0.5 * bw, hy1 = bcy - 0.5 * bw;
    double my[8];
    my[0] = x + (hx0 * cs + hy0 * -sn); my[1] = y + (hx0 * sn + hy0 * cs);
    my[2] = x + (hx1 * cs + hy0 * -sn); my[3] = y + (hx1 * sn + hy0 * cs);
    my[4] = x + (hx1 * cs + hy1 * -sn); my[5] = y + (hx1 * sn + hy1 * cs);
    my[6] = x + (hx0 * cs + hy1 * -sn); my[7] = y + (hx0 * sn + hy1 * cs);
#pragma unroll
    for (int f = 0; f < 8; ++f) c.corners[f * c.G + c.s] = my[f];
    if (RSS) { c.hcs[c.s] = cs; c.hcs[c.G + c.s] = sn; }
    c.orient[c.s] = (int8_t)(orient_hint ? orient_hint : quad_orientation(quad_from_array(my)));
    bb = make_aabb_box(x, y, cs, sn, bw, bl, bcx, bcy, ox, oy);
  }
  if (SORTED) {  // publish the new key at the position of the last tick; sorted_scatter re-ranks the box
    if (!(bb.x <= bb.z)) bb = make_float4(INFINITY, INFINITY, -INFINITY, -INFINITY);  // NaN pose: sorts last
    c.skey[c.posof[c.s]] = bb.x;
    c.tmpbox[c.s] = bb;
    return;
  }
  c.aabb[c.s] = bb;
  if (c.s < c.H + 1) c.aabb[c.M + c.s] = bb;
}

// ---------------------------------------------------------------------------------
// Sorted sweep (vehicle scenes, M >= 128).  The conservative AABBs are kept sorted by their lower
// x bound across ticks: entities move a few metres per tick, so an odd-even transposition pass or
// two restores the order (three rotating flags tell every thread whether a round swapped
// anything).  Boxes that overlap box r in x then sit directly behind it, so the thread of position
// r tests a fixed, branch-free window of SG_SWEEP_WIN successors (and walks on only while a successor still
// starts inside its x-range): O(M * window) tests instead of O(M^2 / 2), each unordered pair once.
// ---------------------------------------------------------------------------------
SG_DEV void sorted_setup(const Grp& c) {  // all G threads, once per launch (followed by a group_sync)
  if (!c.sorted) return;
  if (c.s < c.M) { c.posof[c.s] = (uint16_t)c.s; c.sid[c.s] = (uint16_t)c.s; }
  for (int q = c.M + c.s; q < c.M + c.H + 1; q += c.G)
    c.aabb[q] = make_float4(INFINITY, INFINITY, -INFINITY, -INFINITY);
  if (c.s <= SG_SORT_WIN) { c.skey[-1 - c.s] = -INFINITY; c.skey[c.M + c.s] = INFINITY; }
  if (c.s < 3) c.sflag[c.s] = 0;
  if (c.s == 3) c.sflag[3] = 1;  // the launch starts from slot order: the first tick sorts from scratch
}
// full repair: odd-even transposition rounds until a round swaps nothing (any disorder; used for the
// first tick of a launch and whenever the one-pass re-ranking below does not verify)
SG_DEV void sort_positions(const Grp& c, int& round) {  // all G threads
  const int r = c.s;
  for (;;) {
    int* flag = c.sflag + round % 3;
    if (r == 0) c.sflag[(round + 1) % 3] = 0;  // last read two rounds ago, next written in the next round
#pragma unroll
    for (int ph = 0; ph < 2; ++ph) {
      if ((r & 1) == ph && r + 1 < c.M) {
        const float4 a = c.aabb[r], b = c.aabb[r + 1];
        if (a.x > b.x) {
          c.aabb[r] = b; c.aabb[r + 1] = a;
          const uint16_t t = c.sid[r];
          c.sid[r] = c.sid[r + 1]; c.sid[r + 1] = t;
          *flag = 1;
        }
      }
      group_sync(c);
    }
    const int swapped = *flag;
    ++round;
    if (!swapped) break;
  }
  if (r < c.M) c.posof[c.sid[r]] = (uint16_t)r;
}
// One-pass re-ranking (owner threads, after the barrier that follows publish_box): the boxes were
// sorted before the tick and move a few metres per tick, so a box's new rank differs from its old
// position p only by the boxes within SG_SORT_WIN positions that it overtook or that overtook it:
//   rank = p - #(earlier boxes in the window with a larger key) + #(later ones with a smaller key)
// (ties keep their order).  The result is verified by broad_phase_sorted; when the window was too
// small the tick falls back to sort_positions.
SG_DEV void sorted_scatter(const Grp& c) {
  const int p = c.posof[c.s];
  const float k = c.skey[p];
  int np = p;
#pragma unroll
  for (int d = 1; d <= SG_SORT_WIN; ++d) {
    np -= c.skey[p - d] > k ? 1 : 0;
    np += c.skey[p + d] < k ? 1 : 0;
  }
  c.aabb[np] = c.tmpbox[c.s];
  c.sid[np] = (uint16_t)c.s;
  c.posof[c.s] = (uint16_t)np;
}
template <bool VERIFY = false>
SG_DEV void broad_phase_sorted(const Grp& c, int parity) {  // thread = sorted position
  const int r = c.s;
  const float4 mb = c.aabb[r];
  if (VERIFY) {  // every position written by exactly one slot, keys in order
    const float nx = c.aabb[r + 1].x;  // (position M holds an empty box: key +inf)
    if (c.posof[c.sid[r]] != r || mb.x > nx) c.sflag[3] = 1;
  }
  if (!(mb.x <= mb.z)) return;  // empty box (entity absent)
  const float4* nb = c.aabb + r + 1;
  int* acc = c.acc + parity * ACC_N;
  auto push = [&](int d) {
    const int q = atomicAdd(&acc[ACC_QCOUNT], 1);
    if (q < c.QCAP) c.queue[q] = ((uint32_t)c.sid[r] << 16) | (uint32_t)c.sid[r + d];
  };
  uint32_t hits = 0;
#pragma unroll
  for (int dd = 0; dd < SG_SWEEP_WIN; ++dd) {  // positions beyond M hold empty boxes
    const float4 ob = nb[dd];
    SG_AABB_TEST(hits, mb, ob, 1u << dd);
  }
  while (hits) {
    const int dd = __ffs(hits) - 1;
    hits &= hits - 1;
    push(dd + 1);
  }
  for (int d = SG_SWEEP_WIN + 1; r + d < c.M && c.aabb[r + d].x <= mb.z; ++d) {  // a run longer than the window
    const float4 ob = c.aabb[r + d];
    if (mb.x <= ob.z && ob.x <= mb.z && mb.y <= ob.w && ob.y <= mb.w) push(d);
  }
}

// the ego's box dimensions and the reciprocals safe_ratios divides by (once per launch)
SG_DEV void publish_ego_box(const Grp& c) {
  const double eW = c.boxp[c.s], eL = c.boxp[c.G + c.s];
  c.egop[EGO_W] = eW; c.egop[EGO_L] = eL;
  c.egop[EGO_RHW] = 1.0 / (0.5 * eW); c.egop[EGO_RHL] = 1.0 / (0.5 * eL);
}

// The norm of a direction that is a unit vector up to rounding (s2 = a^2 + b^2 within 1e-8 of 1) and its
// reciprocal, without the square-root / reciprocal iterations: sqrt(1 + d) = 1 + d/2 - d^2/8 ..., so 1 + d/2 is
// the correctly rounded root for |d| < 1e-8 (d = s2 - 1 is exact), and one Newton step on 2 - n gives 1/n to
// the last bit.  publish_ego is one lane's serial work on the tick's critical path (every other warp of the
// scenario waits for it at the barrier); the general case falls back to fnorm2 / fast_rcp.
SG_DEV void unit_norm(double a, double b, double& nn, double& rn) {
  const double s2 = a * a + b * b, d = s2 - 1.0;
  if (fabs(d) < 1e-8) {
    nn = 1.0 + 0.5 * d;
    const double r0 = 2.0 - nn, e = __fma_rn(-nn, r0, 1.0);
    rn = __fma_rn(r0, e, r0);
  } else {
    nn = fast_sqrt(s2);
    rn = fast_rcp(nn);
  }
}

// ego parameters in its own frame (reference metrics/rss/callback.py:73-97, 340-386)
SG_DEV void publish_ego(const Grp& c, bool present, double x, double y, double ec, double es,
                        double vx, double vy) {
  double einv[2];
  {  // inverse_direction((cos h, sin h)), rss_utils.py:7-21
    double nn, rn;
    unit_norm(es, ec, nn, rn);
    einv[0] = div_r(es, nn, rn);
    einv[1] = div_r(-ec, nn, rn);
  }
  double* E = c.egop;
  E[EGO_X] = x; E[EGO_Y] = y;
  E[EGO_C] = ec; E[EGO_S] = es;
  E[EGO_INV0] = einv[0]; E[EGO_INV1] = einv[1];
  const double hd[2] = {dot2(ec, es, einv[0], einv[1]), dot2(ec, es, ec, es)};
  E[EGO_HD0] = hd[0]; E[EGO_HD1] = hd[1];
  double hinv[2];  // inverse_direction(ego-frame heading): used by safe_lateral_distance
  {
    double nn, rn;
    unit_norm(hd[1], hd[0], nn, rn);
    hinv[0] = div_r(hd[1], nn, rn);
    hinv[1] = div_r(-hd[0], nn, rn);
  }
  E[EGO_HINV0] = hinv[0]; E[EGO_HINV1] = hinv[1];
  const double v0 = dot2(vx, vy, einv[0], einv[1]), v1 = dot2(vx, vy, ec, es);
  E[EGO_V0] = v0; E[EGO_V1] = v1;
  E[EGO_VNORM] = fnorm2(v0, v1);
  E[EGO_VLONG] = fabs(dot2(v0, v1, hd[0], hd[1]));
  E[EGO_PRESENT] = present ? 1.0 : 0.0;
}

// ---------------------------------------------------------------------------------
// RSS (reference metrics/rss/callback.py)
// ---------------------------------------------------------------------------------
// The hazard's ego-frame corners are staged in shared memory (rbox) for these; `qa` is the shared
// address of the thread's first coordinate, `qs` the byte stride between coordinates.

// (These read the corners on the fly with ld.shared inside rolled loops: few registers, so the
// caller saves little around the call.)
SG_DEV double qx(unsigned qa, unsigned qs, int k) { return lds_f64(qa + 2u * (unsigned)(k & 3) * qs); }
SG_DEV double qy(unsigned qa, unsigned qs, int k) { return lds_f64(qa + (2u * (unsigned)(k & 3) + 1u) * qs); }
SG_DEV int quad_orientation_sh(unsigned qa, unsigned qs) {
  int s = orient_sign(qx(qa, qs, 0), qy(qa, qs, 0), qx(qa, qs, 1), qy(qa, qs, 1), qx(qa, qs, 2), qy(qa, qs, 2));
  if (s == 0)
    s = orient_sign(qx(qa, qs, 1), qy(qa, qs, 1), qx(qa, qs, 2), qy(qa, qs, 2), qx(qa, qs, 3), qy(qa, qs, 3));
  return s;
}

// Does the infinite line through (ax, ay), (bx, by) meet the closed convex quad?  (Exact.)
static __device__ __noinline__ bool line_hits_quad(unsigned qa, unsigned qs, double ax, double ay, double bx,
                                            double by) {
  int pos = 0, neg = 0;
#pragma unroll 1
  for (int m = 0; m < 4; ++m) {
    const int sg = orient_sign(ax, ay, bx, by, qx(qa, qs, m), qy(qa, qs, m));
    pos += sg > 0;
    neg += sg < 0;
  }
  return !(pos == 4 || neg == 4);
}

// Closed intersection of a convex quad with the rectangle [-a, a] x [-b, b] whose bounding
// ranges already overlap: separating axes are the quad's edges; per edge only the rectangle
// corner that is extreme towards the quad's inside has to be tested.  (Exact.)
static __device__ __noinline__ bool quad_hits_centered_rect(unsigned qa, unsigned qs, double a, double b) {
  const int o = quad_orientation_sh(qa, qs);
#pragma unroll 1
  for (int k = 0; k < 4; ++k) {
    const double x0 = qx(qa, qs, k), y0 = qy(qa, qs, k), x1 = qx(qa, qs, k + 1), y1 = qy(qa, qs, k + 1);
    // orient(p0, p1, p) = dx*(py-y0) - dy*(px-x0): o*orient is largest for
    // py = b*sign(o*dx), px = -a*sign(o*dy)
    const double dx = x1 - x0, dy = y1 - y0;
    const double py = ((o > 0) == (dx > 0) || dx == 0) ? b : -b;
    const double px = ((o > 0) == (dy > 0) && dy != 0) ? -a : a;
    if (orient_sign(x0, y0, x1, y1, px, py) * o < 0) return false;  // every corner strictly outside
  }
  return true;
}

// Closed intersection of a convex quad with the horizontal segment y = c, |x| <= w, when the
// quad's y-range already contains c: the quad's corners are then not strictly on one side of
// the segment's line, so the segment misses the quad iff some quad edge has both segment
// endpoints strictly outside.  (Exact.)
static __device__ __noinline__ bool quad_hits_hsegment(unsigned qa, unsigned qs, double w, double c) {
  const int o = quad_orientation_sh(qa, qs);
#pragma unroll 1
  for (int k = 0; k < 4; ++k) {
    const double x0 = qx(qa, qs, k), y0 = qy(qa, qs, k), x1 = qx(qa, qs, k + 1), y1 = qy(qa, qs, k + 1);
    if (orient_sign(x0, y0, x1, y1, -w, c) * o < 0 && orient_sign(x0, y0, x1, y1, w, c) * o < 0)
      return false;
  }
  return true;
}

static __device__ __noinline__ bool rss_box_hits_segment(unsigned qa, unsigned qs, double x0, double y0,
                                                  double x1, double y1) {
  const Quad q = load_quad_shared(qa, qs);
  return quad_intersects_segment(q, quad_orientation(q), x0, y0, x1, y1);
}

struct RssConst {  // uniform per launch
  double CLR, RT, MAXA, MINA, r2mina;
};

// RSSDistances.__call__ for one hazard (callback.py:57-122).  Geometric predicates are
// answered by exact comparisons where those decide and by exact orientation signs otherwise,
// so every record equals the reference's; scalar quotients with a launch-uniform or shared
// denominator use div_r.
// safe_ratios (callback.py:124-166) of one hazard from the ego published in shared memory; pure
// outputs (no decision reads them), so a fused rollout only needs them after its last tick
SG_DEV void rss_ratios(const Grp& c, double x, double y, double* out, int ost) {
  const double* E = c.egop;
  const double eh0 = E[EGO_C], eh1 = E[EGO_S], ei0 = E[EGO_INV0], ei1 = E[EGO_INV1];
  const double dirc = c.hcs[c.s], dirs = c.hcs[c.G + c.s];
  const double bw = c.boxp[c.s], bl = c.boxp[c.G + c.s];
  const double d0 = x - E[EGO_X], d1 = y - E[EGO_Y];
  const double pos0 = dot2(d0, d1, ei0, ei1), pos1 = dot2(d0, d1, eh0, eh1);
  const double hd0 = dot2(dirc, dirs, ei0, ei1), hd1 = dot2(dirc, dirs, eh0, eh1);
  const double eW = E[EGO_W], eL = E[EGO_L];
  const double hn = fnorm2(hd1, hd0), rhn = fast_rcp(hn);  // inverse_direction(haz heading)
  const double inv0 = div_r(hd1, hn, rhn), inv1 = div_r(-hd0, hn, rhn);
  const double wl_inv = fabs(dot2(bw, bl, inv0, inv1));
  const double wl_dir = fabs(dot2(bw, bl, hd0, hd1));
  const double actual_lat = py_max(1e-6, fabs(pos0) - 0.5 * eW - 0.5 * wl_inv);
  const double actual_long = py_max(1e-6, fabs(pos1) - 0.5 * eL - 0.5 * wl_dir);
  out[2 * ost] = fabs(div_r(actual_lat, 0.5 * eW, E[EGO_RHW]));
  out[3 * ost] = fabs(div_r(actual_long, 0.5 * eL, E[EGO_RHL]));
}

template <bool RATIOS = true>
SG_DEV int rss_hazard(const RssConst& K, const Grp& c, double x, double y, double vx, double vy,
                      uint8_t& state, double* out, int ost) {  // out: lat, long, ratio lat, ratio long
  const double* E = c.egop;
  const double eh0 = E[EGO_C], eh1 = E[EGO_S], ei0 = E[EGO_INV0], ei1 = E[EGO_INV1];
  const double dirc = c.hcs[c.s], dirs = c.hcs[c.G + c.s];
  const double bw = c.boxp[c.s], bl = c.boxp[c.G + c.s];
  // get_entity_parameters (callback.py:340-386) with coord_change (rss_utils.py:24-45)
  const double d0 = x - E[EGO_X], d1 = y - E[EGO_Y];
  const double pos0 = dot2(d0, d1, ei0, ei1), pos1 = dot2(d0, d1, eh0, eh1);
  const double hd0 = dot2(dirc, dirs, ei0, ei1), hd1 = dot2(dirc, dirs, eh0, eh1);
  const double v0 = dot2(vx, vy, ei0, ei1), v1 = dot2(vx, vy, eh0, eh1);
  const double eW = E[EGO_W], eL = E[EGO_L];
  const double eHD0 = E[EGO_HD0], eHD1 = E[EGO_HD1];
  // safe_longitudinal_distance (callback.py:230-269); the ego's position is [0, 0].  Both
  // branches of the reference (same / opposite direction) are evaluated and selected: lanes of
  // a warp take both anyway, and straight-line code lets the two chains overlap.
  double slong;
  {
    const double dp = dot2(eHD0, eHD1, hd0, hd1);
    const double a = fabs(K.MAXA * dp);
    const double hv = dot2(v0, v1, eHD0, eHD1);
    const double rta = K.RT * a;
    // same direction: long_dist_same_direction, callback.py:243-256, 454-472
    const bool ahead = 0.0 > pos1;
    const double vf = ahead ? E[EGO_VNORM] : hv, vr = ahead ? hv : E[EGO_VNORM];
    const double a2 = 2 * a;
    const double vf2a = div_r(vf * vf, a2, fast_rcp(a2));
    const double u = vr + rta;
    const double dd_s = py_max(0, vr * K.RT + py_min(vf2a, 0.5 * a * (K.RT * K.RT)) +
                                      div_r(u * u, 2 * K.MINA, K.r2mina) - vf2a);
    // opposite direction: long_dist_opp_direction, callback.py:257-268, 474-492
    const double v1e = E[EGO_VLONG], av2 = fabs(hv);
    const double u1 = v1e + rta, u2 = av2 + rta;
    const double dd_o = py_max(0, (2 * v1e + rta) * K.RT / 2 + div_r(u1 * u1, 2 * K.MINA, K.r2mina) +
                                      (2 * av2 + rta) * K.RT / 2 + div_r(u2 * u2, 2 * K.MINA, K.r2mina));
    const bool same = dp > 0;
    const bool early = same ? (vr == 0.0) : (np_sign(pos1) == np_sign(v1));
    const double dd = same ? dd_s : dd_o;
    slong = fabs(early ? K.CLR + 0.5 * eL : dd + K.CLR + 0.5 * eL);
  }
  // safe_lateral_distance (callback.py:271-302), lat_dist :494-505
  double slat;
  {
    const double k = fabs(dot2(E[EGO_HINV0], E[EGO_HINV1], hd0, hd1));
    const double amax = K.MAXA * k, amin = K.MINA * k;
    const double v = fabs(v0);
    const double den = 2 * amin, rden = fast_rcp(den);
    const double w = K.RT * amax, u = v + w;
    const double dl = py_max(0, 0.5 * K.RT * (2 * v + w) + div_r(u * u, den, rden) -
                                    0.5 * (K.RT * K.RT) * amax - div_r(w * w, den, rden));
    const bool conv = np_sign(-pos0) == np_sign(v0);  // lateral convergence
    const bool early = conv && v == 0.0;
    const double dd = (conv && !early) ? dl : 0.0;
    slat = fabs(early ? K.CLR + 0.5 * eW : dd + K.CLR + 0.5 * eW);
  }
  out[0] = slat;
  out[ost] = slong;
  // safe_ratios (callback.py:124-166)
  if (RATIOS) {
    const double hn = fnorm2(hd1, hd0), rhn = fast_rcp(hn);  // inverse_direction(haz heading)
    const double inv0 = div_r(hd1, hn, rhn), inv1 = div_r(-hd0, hn, rhn);
    const double wl_inv = fabs(dot2(bw, bl, inv0, inv1));
    const double wl_dir = fabs(dot2(bw, bl, hd0, hd1));
    const double actual_lat = py_max(1e-6, fabs(pos0) - 0.5 * eW - 0.5 * wl_inv);
    const double actual_long = py_max(1e-6, fabs(pos1) - 0.5 * eL - 0.5 * wl_dir);
    out[2 * ost] = fabs(div_r(actual_lat, 0.5 * eW, E[EGO_RHW]));
    out[3 * ost] = fabs(div_r(actual_long, 0.5 * eL, E[EGO_RHL]));
  }
  // unsafe_distance (callback.py:168-228)
  if ((state >> 2) & 3) return SG_RSS_FOUND;
  double box[8];  // hazard corners in the ego frame
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const double c0 = c.corners[(2 * q) * c.G + c.s] - E[EGO_X];
    const double c1 = c.corners[(2 * q + 1) * c.G + c.s] - E[EGO_Y];
    box[2 * q] = dot2(c0, c1, ei0, ei1);
    box[2 * q + 1] = dot2(c0, c1, eh0, eh1);
  }
  double* rbp = c.rbox + c.s;  // the same corners, staged for the out-of-line exact predicates
#pragma unroll
  for (int q = 0; q < 8; ++q) rbp[q * c.G] = box[q];
  const unsigned rb = c.rbox_sh + (unsigned)c.s * 8u, st = (unsigned)c.G * 8u;
  const double bxmin = min2(min2(box[0], box[2]), min2(box[4], box[6]));
  const double bxmax = max2(max2(box[0], box[2]), max2(box[4], box[6]));
  const double bymin = min2(min2(box[1], box[3]), min2(box[5], box[7]));
  const double bymax = max2(max2(box[1], box[3]), max2(box[5], box[7]));
  bool inter = false;
  if (!(bxmin > slat || bxmax < -slat || bymin > slong || bymax < -slong)) {
    bool corner_in = false;  // a hazard corner inside the closed buffer decides at once
#pragma unroll
    for (int q = 0; q < 4; ++q)
      corner_in = corner_in || (fabs(box[2 * q]) <= slat && fabs(box[2 * q + 1]) <= slong);
    inter = corner_in || quad_hits_centered_rect(rb, st, slat, slong);
  }
  if (inter) {
    const int marker = state & 3;
    if (marker == 1) { state |= 2 << 2; return SG_RSS_UNSAFE_LONGITUDINAL; }
    if (marker == 2) { state |= 1 << 2; return SG_RSS_UNSAFE_LATERAL; }
    const double ed[2] = {eW, eL};
    double inv[2];
    inverse_direction(ed, inv);
    const double lhs = fabs(fabs(pos0) - fabs(dot2(pos0, pos1, ed[0], ed[1]))) / slat;
    const double rhs = fabs(fabs(pos1 - dot2(pos0, pos1, inv[0], inv[1])) / slong);
    if (lhs > rhs) { state |= 2 << 2; return SG_RSS_UNSAFE_LONGITUDINAL; }
    state |= 1 << 2;
    return SG_RSS_UNSAFE_LATERAL;
  }
  // write_intersections (callback.py:304-338) against generate_buffer's segments (:429-451):
  // "lengths" are the diagonals (+-slat, 100 slong) -> (-+slat, -100 slong), "widths" the
  // horizontal segments y = +-slong, |x| <= 100 slat.
  const double L100 = 100 * slong, W100 = 100 * slat;
  bool lat = false, lon = false;
  if (!(bxmin > slat || bxmax < -slat || bymin > L100 || bymax < -L100)) {
    if (bymax < L100 && bymin > -L100) {  // segment spans the box's y-range: segment <=> line
      // Both diagonals pass through the origin (their end points are exact negatives), so
      // orient(a, -a, c) = 2 (a_y c_x - a_x c_y): the side of corner c is the sign of
      // L100 c_x -+ slat c_y, decided here whenever it clears the rounding bound of that
      // expression (3 roundings) and by the exact predicate otherwise.
      int pos1 = 0, neg1 = 0, pos2 = 0, neg2 = 0;
      bool unsure = false;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const double pp = L100 * box[2 * q], rr = slat * box[2 * q + 1];
        const double d1 = pp - rr, d2 = pp + rr, tol = 4e-16 * (fabs(pp) + fabs(rr));
        unsure = unsure || !(fabs(d1) > tol) || !(fabs(d2) > tol);
        pos1 += d1 > 0; neg1 += d1 < 0;
        pos2 += d2 > 0; neg2 += d2 < 0;
      }
      if (unsure)
        lat = line_hits_quad(rb, st, slat, L100, -slat, 100 * -slong) ||
              line_hits_quad(rb, st, -slat, L100, slat, 100 * -slong);
      else
        lat = !(pos1 == 4 || neg1 == 4) || !(pos2 == 4 || neg2 == 4);
    } else
      lat = rss_box_hits_segment(rb, st, slat, L100, -slat, 100 * -slong) ||
            rss_box_hits_segment(rb, st, -slat, L100, slat, 100 * -slong);
  }
  if (!(bxmin > W100 || bxmax < -W100)) {
    if (bxmin > -W100 && bxmax < W100) {  // box inside the segments' x-range: segment <=> line y = c
      lon = (bymin <= slong && slong <= bymax) || (bymin <= -slong && -slong <= bymax);
    } else {
      if (!(bymin > slong || bymax < slong)) lon = quad_hits_hsegment(rb, st, W100, slong);
      if (!lon && !(bymin > -slong || bymax < -slong)) lon = quad_hits_hsegment(rb, st, W100, -slong);
    }
  }
  if (lat && lon) return SG_RSS_BOTH;
  if (lat) { state = (uint8_t)((state & ~3) | 1); return SG_RSS_LATERAL; }
  if (lon) { state = (uint8_t)((state & ~3) | 2); return SG_RSS_LONGITUDINAL; }
  return SG_RSS_SAFE;
}

// ---------------------------------------------------------------------------------
// collisions (reference state/utils.py:10-49, utils.py:28-62)
// ---------------------------------------------------------------------------------
#ifdef SG_FLAT_BOXES
// two staged quads without area (segments / points): they meet iff their rings do
static __device__ __noinline__ bool flat_pair_meets(unsigned pa, unsigned pb, unsigned qs) {
#pragma unroll 1
  for (int i = 0; i < 4; ++i) {
    const double ax = qx(pa, qs, i), ay = qy(pa, qs, i), bx = qx(pa, qs, i + 1), by = qy(pa, qs, i + 1);
#pragma unroll 1
    for (int j = 0; j < 4; ++j)
      if (segments_meet(ax, ay, bx, by, qx(pb, qs, j), qy(pb, qs, j), qx(pb, qs, j + 1), qy(pb, qs, j + 1))) return true;
  }
  return false;
}
#endif

// exact narrow phase for one AABB-surviving pair; both quads are read on the fly from the staged
// corners with ld.shared inside rolled loops (`csh`: shared address of corners[0][0]), so the
// routine needs few registers and its callers save little around the call.
// Boxes without area (orientation 0: segments / points) follow their own rules -- such a quad has no
// inside: its edge separates when the other quad lies strictly on one side of it, whichever side, and
// two of them meet iff their rings do.  Those rules are compiled in with SG_FLAT_BOXES only (the
// general, replay and crowd kernels): they cost registers at every call site, so the vehicle kernels
// leave them out and sg_api.cu routes scenes with such boxes (SG_SCENE_FLAT_BOXES) to the general kernel.
static __device__ __noinline__ bool pair_collides(unsigned csh, const int8_t* orient, int G, int a, int b) {
  const unsigned qs = (unsigned)G * 8u;
  unsigned pa = csh + (unsigned)a * 8u, pb = csh + (unsigned)b * 8u;
  bool same = true;
#pragma unroll 1
  for (unsigned f = 0; f < 8; ++f) same = same && (lds_f64(pa + f * qs) == lds_f64(pb + f * qs));
  if (same) return false;  // `g != g_prime`, reference utils.py:58
  int oa = orient[a], ob = orient[b];
#ifdef SG_FLAT_BOXES
  if (oa == 0 && ob == 0) return flat_pair_meets(pa, pb, qs);
#endif
  // closed-set intersection of two convex quads (touching counts, as GEOS `intersects`):
  // disjoint iff some edge of either has all four corners of the other strictly outside
#pragma unroll 1
  for (int pass = 0; pass < 2; ++pass) {
#pragma unroll 1
    for (int k = 0; k < 4; ++k) {
      const double ax = qx(pa, qs, k), ay = qy(pa, qs, k), bx = qx(pa, qs, k + 1), by = qy(pa, qs, k + 1);
      bool sep = true;
#ifdef SG_FLAT_BOXES
      int want = -oa;
#pragma unroll 1
      for (int m = 0; m < 4 && sep; ++m) {
        const int sg = orient_sign(ax, ay, bx, by, qx(pb, qs, m), qy(pb, qs, m));
        if (want == 0) want = sg;  // flat quad: the side of the first corner decides which side is "outside"
        sep = sg != 0 && sg == want;
      }
#else
#pragma unroll 1
      for (int m = 0; m < 4 && sep; ++m)
        sep = orient_sign(ax, ay, bx, by, qx(pb, qs, m), qy(pb, qs, m)) * oa < 0;
#endif
      if (sep) return false;
    }
    const unsigned tp = pa; pa = pb; pb = tp;
    const int to = oa; oa = ob; ob = to;
  }
  return true;
}

// Where a colliding pair is booked: passed BY VALUE to the out-of-line routines so that the
// group descriptor itself never has its address taken (it then lives in registers / is
// rematerialised from the constant bank instead of being re-read from local memory).
struct PairSink {
  int* acc;            // this tick's accumulators
  uint32_t* bits;      // this tick's collided bits
  uint32_t* ego_now;   // ego row of this tick
  uint32_t* rows;      // optional pair matrix of this scenario (SG_FEAT_COLL_MATRIX)
  int W, ego_slot, first_slot;
};
SG_DEV PairSink make_sink(int features, uint32_t* coll_mask, const Grp& c, int ego_slot, int first_slot,
                          int parity) {
  PairSink k;
  k.acc = c.acc + parity * ACC_N;
  k.bits = c.bits + parity * c.W;
  k.ego_now = c.ego_now;
  k.rows = (features & SG_FEAT_COLL_MATRIX) ? coll_mask + (int64_t)c.n * c.M * c.W : nullptr;
  k.W = c.W; k.ego_slot = ego_slot; k.first_slot = first_slot;
  return k;
}

// book-keeping for one colliding pair (scenario-level shared atomics)
static __device__ __noinline__ void commit_pair(PairSink k, int a, int b) {
  const int lo = min(a, b), hi = max(a, b);
  atomicAdd(&k.acc[ACC_NPAIRS], 1);
  atomicMin(&k.acc[ACC_FIRST_PAIR], (lo << 16) | hi);
  if (lo == k.first_slot || hi == k.first_slot) k.acc[ACC_FIRST_HIT] = 1;
  atomicOr(&k.bits[lo >> 5], 1u << (lo & 31));
  atomicOr(&k.bits[hi >> 5], 1u << (hi & 31));
  if (lo == k.ego_slot) atomicOr(&k.ego_now[hi >> 5], 1u << (hi & 31));
  if (hi == k.ego_slot) atomicOr(&k.ego_now[lo >> 5], 1u << (lo & 31));
  if (k.rows) {
    atomicOr(&k.rows[(int64_t)lo * k.W + (hi >> 5)], 1u << (hi & 31));
    atomicOr(&k.rows[(int64_t)hi * k.W + (lo >> 5)], 1u << (lo & 31));
  }
}

// Conservative separating-axis classification of two boxes from their fp64 corners (ring order of
// entity/base.py:117-136: corner1 - corner0 runs along the box's length, corner0 - corner3 along its
// width): +1 they intersect, -1 they are disjoint, 0 too close to call -- the exact predicate decides.
// Rectangles are disjoint iff one of their four edge directions separates them.  Every gap
//   |D.a| - (|E1A.a| + |E2A.a| + |E1B.a| + |E2B.a|)      (D = 2 x centre difference, E = full edges)
// is compared with a tolerance of 1e-9 relative to the coordinates involved, seven orders of
// magnitude above its rounding error, so a +-1 answer is the exact closed-set answer for these
// corners.  Branch-free, ~130 flops: the exact edge tests cost ~8x that and diverge.
SG_DEV int sat_classify(const double* A, const double* B) {
  const double e1ax = A[2] - A[0], e1ay = A[3] - A[1], e2ax = A[0] - A[6], e2ay = A[1] - A[7];
  const double e1bx = B[2] - B[0], e1by = B[3] - B[1], e2bx = B[0] - B[6], e2by = B[1] - B[7];
  const double dx = (B[0] + B[4]) - (A[0] + A[4]), dy = (B[1] + B[5]) - (A[1] + A[5]);
  const double scale = fabs(A[0]) + fabs(A[1]) + fabs(B[0]) + fabs(B[1]) + fabs(e1ax) + fabs(e1ay) + fabs(e2ax) +
                       fabs(e2ay) + fabs(e1bx) + fabs(e1by) + fabs(e2bx) + fabs(e2by);
  const double rel = 1e-9 * scale;
  bool apart = false, inside = true;
#define SG_SAT_AXIS(ax, ay)                                                                             \
  {                                                                                                     \
    const double gap = fabs(dx * (ax) + dy * (ay)) -                                                    \
                       (fabs(e1ax * (ax) + e1ay * (ay)) + fabs(e2ax * (ax) + e2ay * (ay)) +             \
                        fabs(e1bx * (ax) + e1by * (ay)) + fabs(e2bx * (ax) + e2by * (ay)));             \
    const double tol = rel * (fabs(ax) + fabs(ay));                                                     \
    apart = apart || gap > tol;                                                                         \
    inside = inside && gap < -tol;                                                                      \
  }
  SG_SAT_AXIS(e1ax, e1ay)
  SG_SAT_AXIS(e2ax, e2ay)
  SG_SAT_AXIS(e1bx, e1by)
  SG_SAT_AXIS(e2bx, e2by)
#undef SG_SAT_AXIS
  if (!(fabs(dx) + fabs(dy) > rel)) return 0;  // (nearly) coincident boxes: `g != g_prime` is the exact path's call
  return apart ? -1 : (inside ? 1 : 0);
}

// Separating-axis classification of two boxes from centre + half-edge vectors (the quantities sat_classify
// derives from the corners, taken from the staged pose instead): +1 intersect, -1 disjoint, 0 too close to
// call.  Every gap is compared with a tolerance of 1e-9 relative to the coordinates involved -- seven orders
// of magnitude above the rounding error of these expressions AND of the reference's corner formula -- so a
// +-1 answer is the exact closed-set answer for the fp64 corners; 0 sends the pair to the exact predicate.
struct Obb {
  double cx, cy, ux, uy, vx, vy;  // centre, half-length vector, half-width vector
};
SG_DEV int sat_classify_obb(const Obb& A, const Obb& B) {
  const double dx = B.cx - A.cx, dy = B.cy - A.cy;
  const double scale = fabs(A.cx) + fabs(A.cy) + fabs(B.cx) + fabs(B.cy) + fabs(A.ux) + fabs(A.uy) + fabs(A.vx) +
                       fabs(A.vy) + fabs(B.ux) + fabs(B.uy) + fabs(B.vx) + fabs(B.vy);
  const double rel = 1e-9 * scale;
  bool apart = false, inside = true;
#define SG_OBB_AXIS(ax, ay)                                                                              \
  {                                                                                                     \
    const double gap = fabs(dx * (ax) + dy * (ay)) -                                                    \
                       (fabs(A.ux * (ax) + A.uy * (ay)) + fabs(A.vx * (ax) + A.vy * (ay)) +             \
                        fabs(B.ux * (ax) + B.uy * (ay)) + fabs(B.vx * (ax) + B.vy * (ay)));             \
    const double tol = rel * (fabs(ax) + fabs(ay));                                                     \
    apart = apart || gap > tol;                                                                         \
    inside = inside && gap < -tol;                                                                      \
  }
  SG_OBB_AXIS(A.ux, A.uy)
  SG_OBB_AXIS(A.vx, A.vy)
  SG_OBB_AXIS(B.ux, B.uy)
  SG_OBB_AXIS(B.vx, B.vy)
#undef SG_OBB_AXIS
  if (!(fabs(dx) + fabs(dy) > rel)) return 0;  // (nearly) coincident boxes: `g != g_prime` is the exact path's call
  return apart ? -1 : (inside ? 1 : 0);
}

SG_DEV void record_pair(const PairSink& k, unsigned corners_sh, const int8_t* orient, int G, int a, int b) {
  if (pair_collides(corners_sh, orient, G, a, b)) commit_pair(k, a, b);
}

// one queued pair: separating-axis filter on the staged corners, exact predicate when it cannot tell
static __device__ __noinline__ void decide_pair(PairSink k, const double* corners, unsigned corners_sh,
                                                const int8_t* orient, int G, int a, int b) {
  double qa[8], qb[8];
#pragma unroll
  for (int f = 0; f < 8; ++f) { qa[f] = corners[f * G + a]; qb[f] = corners[f * G + b]; }
  const int v = sat_classify(qa, qb);
  if (v > 0) commit_pair(k, a, b);
  else if (v == 0) record_pair(k, corners_sh, orient, G, a, b);
}

// circular half sweep over the staged AABBs (STRtree's envelope filter is closed too);
// survivors go to the scenario's queue
SG_DEV void broad_phase(const Grp& c, int parity) {
  const float4 mb = c.aabb[c.s];
  const float4* nb = c.aabb + c.s + 1;
  int* acc = c.acc + parity * ACC_N;
  for (int d0 = 0; d0 < c.H; d0 += 32) {
    const int dn = min(32, c.H - d0);
    uint32_t hits = 0;
    if (dn == 32) {
#pragma unroll
      for (int dd = 0; dd < 32; ++dd) {
        const float4 ob = nb[d0 + dd];
        SG_AABB_TEST(hits, mb, ob, 1u << dd);
      }
    } else {
      for (int dd = 0; dd < dn; ++dd) {
        const float4 ob = nb[d0 + dd];
        SG_AABB_TEST(hits, mb, ob, 1u << dd);
      }
    }
    while (hits) {
      const int dd = __ffs(hits) - 1;
      hits &= hits - 1;
      const int d = d0 + dd + 1;
      int j = c.s + d;
      if (j >= c.M) j -= c.M;
      if (2 * d == c.M && c.s > j) continue;  // the antipodal pair is seen from both ends
      const int q = atomicAdd(&acc[ACC_QCOUNT], 1);
      if (q < c.QCAP) c.queue[q] = ((uint32_t)c.s << 16) | (uint32_t)j;
    }
  }
}

// queue overflow (very dense scenes): redo the sweep and test every survivor in place
static __device__ __noinline__ void broad_phase_direct(PairSink k, const float4* aabb, unsigned corners_sh,
                                                const int8_t* orient, int G, int M, int H, int s) {
  const float4 mb = aabb[s];
  for (int d = 1; d <= H; ++d) {
    const float4 ob = aabb[s + d];
    if (!(mb.x <= ob.z && ob.x <= mb.z && mb.y <= ob.w && ob.y <= mb.w)) continue;
    int j = s + d;
    if (j >= M) j -= M;
    if (2 * d == M && s > j) continue;
    record_pair(k, corners_sh, orient, G, s, j);
  }
}

static __device__ __noinline__ void broad_phase_direct_sorted(PairSink k, const float4* aabb, const uint16_t* sid,
                                                       unsigned corners_sh, const int8_t* orient, int G, int M,
                                                       int r) {
  const float4 mb = aabb[r];
  if (!(mb.x <= mb.z)) return;
  for (int d = 1; r + d < M && aabb[r + d].x <= mb.z; ++d) {
    const float4 ob = aabb[r + d];
    if (!(mb.x <= ob.z && ob.x <= mb.z && mb.y <= ob.w && ob.y <= mb.w)) continue;
    record_pair(k, corners_sh, orient, G, sid[r], sid[r + d]);
  }
}

// CollisionMetric rising edges of one ego-row word (rare: out of line)
static __device__ __noinline__ void emit_events(SgEvent* events, int32_t* event_count, int event_cap, uint32_t fresh,
                                         int n, int tick, int slot0, double t) {
  while (fresh) {
    const int b = __ffs(fresh) - 1;
    fresh &= fresh - 1;
    const int slot = atomicAdd(event_count, 1);
    if (slot < event_cap) {
      SgEvent ev;
      ev.scenario = n; ev.tick = tick; ev.slot = slot0 + b; ev._pad = 0; ev.t = t;
      events[slot] = ev;
    }
  }
}

// phases B2 + C, shared by both kernel flavours.  Returns state.is_done.  (`n`: the scenario's number in
// the whole batch, for the event records; the arrays are addressed through the group descriptor)
template <bool FAST>
SG_DEV bool finish_tick(const SgParams& p, const SgState& st, const Grp& c, int n, int s, int W,
                        int G, int ego_slot, int first_slot, int parity, int tick, double t,
                        double dt, double length, bool live, bool present, uint8_t& collided,
                        double vx, double vy, double vz, double dist) {
  int* acc = c.acc + parity * ACC_N;
  const int nq = acc[ACC_QCOUNT];
  if (nq > 0) {  // phase B2: exact narrow phase on the queued pairs
    const PairSink sink = make_sink(p.features, st.coll_mask, c, ego_slot, first_slot, parity);
    if (nq <= c.QCAP) {
      // a pair per lane: the branch-free separating-axis filter decides all but knife-edge contacts,
      // which go to the exact predicate
      for (int q = s; q < nq; q += G) {
        const uint32_t pr = c.queue[q];
        decide_pair(sink, c.corners, c.corners_sh, c.orient, G, (int)(pr >> 16), (int)(pr & 0xffff));
      }
    } else if (c.sorted) {
      if (s < c.M) broad_phase_direct_sorted(sink, c.aabb, c.sid, c.corners_sh, c.orient, G, c.M, s);
    } else if (present) {
      broad_phase_direct(sink, c.aabb, c.corners_sh, c.orient, G, c.M, c.H, s);
    }
    group_sync(c);
  }
  // phase C: terminal check + metrics
  const int npairs = acc[ACC_NPAIRS];
  if (live && ((c.bits[parity * W + (s >> 5)] >> (s & 31)) & 1)) collided = 1;
  bool dn = false;  // state.py:268-270, 397-408
  if ((p.terminal & SG_TERM_MAX_LENGTH) && (t + dt > length)) dn = true;
  if ((p.terminal & SG_TERM_COLLISION) && npairs > 0) dn = true;
  if ((p.terminal & SG_TERM_EGO_OFF_ROAD) && acc[ACC_OFFROAD]) dn = true;
  if ((p.terminal & SG_TERM_EGO_COLLISION) && acc[ACC_FIRST_HIT]) dn = true;
  // the single-lane book-keeping runs in another warp than the ego's metrics when there is one,
  // so that neither warp of the scenario carries all the serial work of the tick
  const int bs = s - ((G > 32 && ego_slot < 32) ? 32 : 0);
  if (bs == 0) {
    int* cold = c.cold_i;
    if (npairs > 0) {
      *(long long*)(cold + COLD_PAIR_TICKS) += npairs;
      if (cold[COLD_FIRST_TICK] < 0) {
        const int fp = acc[ACC_FIRST_PAIR];
        cold[COLD_FIRST_TICK] = tick; cold[COLD_FP0] = fp >> 16; cold[COLD_FP1] = fp & 0xffff;
      }
    }
    cold[COLD_RSS] |= acc[ACC_RSS];
    int* nx = c.acc + (parity ^ 1) * ACC_N;  // reset the other parity for the next tick
    nx[ACC_NPAIRS] = 0; nx[ACC_FIRST_PAIR] = 0x7fffffff; nx[ACC_FIRST_HIT] = 0; nx[ACC_RSS] = 0;
    nx[ACC_QCOUNT] = 0; nx[ACC_OFFROAD] = 0;
  }
  if (bs >= 0 && bs < W) {  // CollisionMetric._step, metrics/collision.py:70-75
    const uint32_t now = c.ego_now[bs];
    if (p.features & SG_FEAT_COLLISIONS) {
      const uint32_t fresh = now & ~c.ego_last[bs];
      if (fresh) emit_events(st.events, st.event_count, st.event_cap, fresh, n, tick, bs * 32, t);
      c.ego_last[bs] = now;
    }
    c.ego_now[bs] = 0;
    c.bits[(parity ^ 1) * W + bs] = 0;
  }
  if (s == ego_slot && (p.features & SG_FEAT_EGO_METRICS)) {  // metrics/trajectory.py:20-24,39-42,58-60
    double* m = c.cold_d;
    // (FAST: the vehicle tick's tolerance-level square root / quotient; the general kernel keeps
    // the IEEE operations so that replayed scenes give the same bits on every path)
    const double sp = FAST ? fast_sqrt(vx * vx + vy * vy + vz * vz) : norm3(vx, vy, vz);
    const double w = (FAST && t > 0.0) ? div_r(m[COLD_AVG_T], t, fast_rcp(t)) : m[COLD_AVG_T] / t;
    m[COLD_AVG] += (1.0 - w) * (sp - m[COLD_AVG]);
    m[COLD_AVG_T] = t;
    m[COLD_MAX] = fmax(sp, m[COLD_MAX]);
    m[COLD_EGOD] = dist;
  }
  return dn;
}

// ---------------------------------------------------------------------------------
// The lean vehicle kernels' tick tail when no terminal condition reads the tick's collisions: the narrow
// phase runs on the warps that do NOT hold the ego (whose warp carries the scenario's other serial job,
// publish_ego), nobody waits for it, and what finish_tick's phase C books -- pair counts, the first
// collision, CollisionMetric rising edges, the RSS latch -- is booked one barrier later, at the start of
// the next tick's callback phase (lagged_epilogue; once more after the last tick).  The `collided` bits
// are sticky over the launch and read once at its end.  Same results, a shorter critical path per tick.
// ---------------------------------------------------------------------------------
SG_DEV void narrow_phase_lagged(const SgParams& p, const SgState& st, const Grp& c, int s, int G, int ego_slot,
                                int first_slot, int parity, bool present) {
  const int nq = c.acc[parity * ACC_N + ACC_QCOUNT];
  if (nq <= 0) return;
  PairSink sink = make_sink(p.features, st.coll_mask, c, ego_slot, first_slot, parity);
  sink.bits = c.bits;  // (parity 0's words: sticky)
  if (nq <= c.QCAP) {
    int qs = s - (((ego_slot >> 5) + 1) << 5);  // the lanes behind the ego's warp take the first entries
    if (qs < 0) qs += G;
    for (int q = qs; q < nq; q += G) {
      const uint32_t pr = c.queue[q];
      decide_pair(sink, c.corners, c.corners_sh, c.orient, G, (int)(pr >> 16), (int)(pr & 0xffff));
    }
  } else if (c.sorted) {
    if (s < c.M) broad_phase_direct_sorted(sink, c.aabb, c.sid, c.corners_sh, c.orient, G, c.M, s);
  } else if (present) {
    broad_phase_direct(sink, c.aabb, c.corners_sh, c.orient, G, c.M, c.H, s);
  }
}
// books the tick whose accumulators are `parity` (tick number `tick`, time after it `t`) and clears them
SG_DEV void lagged_epilogue(const SgParams& p, const SgState& st, const Grp& c, int n, int s, int W, int G,
                            int ego_slot, int parity, int tick, double t) {
  int* acc = c.acc + parity * ACC_N;
  const int bs = s - ((G > 32 && ego_slot < 32) ? 32 : 0);
  if (bs == 0) {
    int* cold = c.cold_i;
    const int npairs = acc[ACC_NPAIRS];
    if (npairs > 0) {
      *(long long*)(cold + COLD_PAIR_TICKS) += npairs;
      if (cold[COLD_FIRST_TICK] < 0) {
        const int fp = acc[ACC_FIRST_PAIR];
        cold[COLD_FIRST_TICK] = tick; cold[COLD_FP0] = fp >> 16; cold[COLD_FP1] = fp & 0xffff;
      }
    }
    cold[COLD_RSS] |= acc[ACC_RSS];
    acc[ACC_NPAIRS] = 0; acc[ACC_FIRST_PAIR] = 0x7fffffff; acc[ACC_FIRST_HIT] = 0; acc[ACC_RSS] = 0;
    acc[ACC_QCOUNT] = 0; acc[ACC_OFFROAD] = 0;
  }
  if (bs >= 0 && bs < W) {  // CollisionMetric._step, metrics/collision.py:70-75
    const uint32_t now = c.ego_now[bs];
    if (p.features & SG_FEAT_COLLISIONS) {
      const uint32_t fresh = now & ~c.ego_last[bs];
      if (fresh) emit_events(st.events, st.event_count, st.event_cap, fresh, n, tick, bs * 32, t);
      c.ego_last[bs] = now;
    }
    c.ego_now[bs] = 0;
  }
}

// per-scenario accumulators live in shared memory between ticks
SG_DEV void load_cold(const SgState& st, const Grp& c, int n, int s, int W, int ego_slot) {
  if (s == ego_slot) {
    c.cold_d[COLD_AVG] = st.ego_avg_speed[n]; c.cold_d[COLD_AVG_T] = st.ego_avg_t[n];
    c.cold_d[COLD_MAX] = st.ego_max_speed[n]; c.cold_d[COLD_EGOD] = st.ego_dist[n];
  }
  if (s == 0) {
    c.cold_i[COLD_FIRST_TICK] = st.first_coll_tick[n];
    c.cold_i[COLD_FP0] = st.first_coll_pair[2 * n]; c.cold_i[COLD_FP1] = st.first_coll_pair[2 * n + 1];
    *(long long*)(c.cold_i + COLD_PAIR_TICKS) = st.n_pair_ticks[n];
    c.cold_i[COLD_RSS] = st.rss_flags[n];
    c.cold_i[COLD_HAS_VEH] = 0;
    for (int q = 0; q < 2 * ACC_N; ++q) c.acc[q] = 0;
    c.acc[ACC_FIRST_PAIR] = 0x7fffffff; c.acc[ACC_N + ACC_FIRST_PAIR] = 0x7fffffff;
  }
  if (s < W) {
    c.ego_last[s] = st.ego_hits[(int64_t)n * W + s];
    c.ego_now[s] = 0;
    c.bits[s] = 0; c.bits[W + s] = 0;
  }
}
SG_DEV void store_cold(const SgState& st, const Grp& c, int n, int s, int W, int ego_slot) {
  if (s == ego_slot) {
    st.ego_avg_speed[n] = c.cold_d[COLD_AVG]; st.ego_avg_t[n] = c.cold_d[COLD_AVG_T];
    st.ego_max_speed[n] = c.cold_d[COLD_MAX]; st.ego_dist[n] = c.cold_d[COLD_EGOD];
  }
  // written back by the threads that keep them in finish_tick (no barrier after the last tick)
  const int bs = s - ((c.G > 32 && ego_slot < 32) ? 32 : 0);
  if (bs == 0) {
    st.first_coll_tick[n] = c.cold_i[COLD_FIRST_TICK];
    st.first_coll_pair[2 * n] = c.cold_i[COLD_FP0]; st.first_coll_pair[2 * n + 1] = c.cold_i[COLD_FP1];
    st.n_pair_ticks[n] = *(long long*)(c.cold_i + COLD_PAIR_TICKS);
    st.rss_flags[n] = (uint8_t)c.cold_i[COLD_RSS];
  }
  if (bs >= 0 && bs < W) st.ego_hits[(int64_t)n * W + bs] = c.ego_last[bs];
}

SG_DEV RssConst make_rss_const(const SgParams& p) {
  RssConst K;
  K.CLR = p.rss_min_safe_clearance; K.RT = p.rss_response_time;
  K.MAXA = p.rss_max_long_accel; K.MINA = p.rss_min_long_accel;
  K.r2mina = 1.0 / (2 * p.rss_min_long_accel);
  return K;
}
