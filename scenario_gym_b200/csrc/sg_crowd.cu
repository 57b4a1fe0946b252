// sg_crowd.cu -- the fused tick loop of crowd scenes (C4: ~1000 social-force pedestrians per
// scenario; reference pedestrian/*.py, state/state.py:340-372, state/utils.py:10-49).
//
// Scenes it takes: more than 256 slots, every slot a PedestrianAgent with SocialForce, a
// ReplayTrajectoryAgent (the ego driving past the crowd) or padding; no RSS, no trace.  Everything
// else goes to the general kernel, whose results this kernel reproduces bit for bit (it calls the
// same device functions for the arithmetic; tests/test_gpu_parity.py::test_crowd_cell_grid).
//
// Mapping.  One scenario = one CTA: 1024 threads, a slot each, one CTA per SM at 64 registers (scenes of
// up to 512 slots: 512 threads at 128 registers).  The State rows other threads read (x, y, vx, vy), the
// flags and the heads of the cell grid are double-buffered in shared memory, so a tick needs two CTA
// barriers:
//   step phase       sensors + behaviour -- neighbours listed by last tick's candidate phase, social force
//                    (neighbour terms pooled per warp, summed in state.poses order) -- then controller +
//                    State.step on the thread's own row (written to the other buffer), cos / sin of the
//                    heading; the owner inserts its slot into the cell grid of the NEW positions (linked
//                    lists) while the grid the last candidate phase read is cleared
//   -- barrier: rows, flags and grid of the new positions complete --
//   candidate phase  every slot gathers the ids listed in the 3 x 3 cells around it into a private strip of
//                    shared memory (a light, ragged walk), then runs the candidate tests over the strip:
//                    next tick's sensor list (pedestrians within r) and this tick's collision broad phase by
//                    inscribed / circumscribed circles -- colliding for sure, apart for sure, or queued on
//                    the warp's own pair queue, which the same warp then drains (separating-axis filter /
//                    exact closed-set predicate, corners recomputed from the rows with the reference's
//                    expression, entity/base.py:100-138)
//   -- barrier: pairs booked --
//   epilogue         terminal conditions, CollisionMetric rising edges, ego metrics.
#define SG_FLAT_BOXES 1  // boxes without area follow their own narrow-phase rules (sg_common.cuh)
#include "sg_common.cuh"
#include "sg_internal.h"

#define CR_NBCAP 8       // per-pedestrian neighbour candidate list
#define CR_CAP 16        // ids a slot gathers from its 3 x 3 cells before the tests (more: tested while walking)

struct CrowdLayout {
  int G, W, QCAP;
  int off_state, off_fcs, off_pool, off_rad, off_nbl, off_gstart, off_gsorted, off_glarge, off_queue, off_flags,
      off_ncnt, off_orient, off_hits, off_bits, off_acc, off_cold, off_gmisc, off_unif, off_egoc, off_ring;
  int bytes;
};

// (constexpr: the kernel's shared-memory views are compile-time offsets, no registers or address arithmetic)
__host__ __device__ constexpr CrowdLayout crowd_layout(int ept, int threads) {
  CrowdLayout L{};
  const int G = threads * ept;
  L.G = G;
  L.W = G / 32;
  L.QCAP = 2 * G;
  int o = 0;
  L.off_state = o;   o += 2 * 4 * G * (int)sizeof(double);      // x, y, vx, vy: the State of this tick and of the next
  L.off_fcs = o;     o += 2 * G * (int)sizeof(double);          // cos / sin of the heading
  L.off_pool = o;    o += (threads / 32) * CR_CAP * 32 * (int)sizeof(uint16_t);  // [warp][k][lane]: gathered ids
  L.off_rad = o;     o += G * (int)sizeof(float2);              // inscribed / circumscribed radius of each box
  L.off_nbl = o;     o += CR_NBCAP * G * (int)sizeof(uint16_t);
  o = (o + 15) / 16 * 16;
  L.off_gstart = o;  o += 2 * SG_GRID_CELLS * (int)sizeof(uint16_t);  // head slot of every cell's list (0xffff: empty), 2 grids
  L.off_gsorted = o; o += G * (int)sizeof(uint16_t);                // next slot in the cell's list
  L.off_glarge = o;  o += SG_GRID_LCAP * (int)sizeof(uint16_t);
  o = (o + 15) / 16 * 16;
  L.off_queue = o;   o += L.QCAP * (int)sizeof(uint32_t);
  L.off_flags = o;   o += 2 * G;                                // present | etype << 1 | large << 3, this tick / next
  L.off_ncnt = o;    o += G;                                    // sensor candidates of each slot (next tick)
  L.off_orient = o;  o += G;
  o = (o + 15) / 16 * 16;
  L.off_hits = o;    o += 2 * L.W * (int)sizeof(uint32_t);      // ego_now, ego_last
  L.off_bits = o;    o += L.W * (int)sizeof(uint32_t);          // slots that collided during this launch
  L.off_acc = o;     o += 2 * ACC_N * (int)sizeof(int);
  L.off_cold = o;    o += COLD_ND * (int)sizeof(double) + COLD_NI * (int)sizeof(int);
  L.off_gmisc = o;   o += 48 * (int)sizeof(int);
  L.off_unif = o;    o += 8 * (int)sizeof(double);            // launch-uniform doubles (kept out of the registers)
  L.off_egoc = o;    o += 12 * (int)sizeof(double);           // the ego's cached row (crowd_replay_cached)
  L.off_ring = o;    o += 32 * 18 * (int)sizeof(double);      // the ego's next 32 steps (crowd_ego_ring)
  L.bytes = (o + 15) / 16 * 16;
  return L;
}

struct Crowd {  // shared-memory views of one scenario
  int G, W, QCAP, M;
  double* state;   // [4][G]
  double* fcs;     // [2][G]
  uint16_t* pool;
  float2* rad;
  uint16_t* nbl;
  uint16_t* ghead;
  uint16_t* gnext;
  uint16_t* glarge;
  uint32_t* queue;
  uint8_t* flags;
  uint8_t* ncnt;
  int8_t* orient;
  uint32_t* ego_now;
  uint32_t* ego_last;
  uint32_t* bits;
  int* acc;
  double* cold_d;
  int* cold_i;
  int* gmisc;
  double* unif;
  double* egoc;
  double* ring;
};

SG_DEV void crowd_views(Crowd& c, const CrowdLayout& L, unsigned char* base, int M) {
  c.G = L.G; c.W = L.W; c.QCAP = L.QCAP; c.M = M;
  c.state = (double*)(base + L.off_state);
  c.fcs = (double*)(base + L.off_fcs);
  c.pool = (uint16_t*)(base + L.off_pool);
  c.rad = (float2*)(base + L.off_rad);
  c.nbl = (uint16_t*)(base + L.off_nbl);
  c.ghead = (uint16_t*)(base + L.off_gstart);
  c.gnext = (uint16_t*)(base + L.off_gsorted);
  c.glarge = (uint16_t*)(base + L.off_glarge);
  c.queue = (uint32_t*)(base + L.off_queue);
  c.flags = base + L.off_flags;
  c.ncnt = base + L.off_ncnt;
  c.orient = (int8_t*)(base + L.off_orient);
  c.ego_now = (uint32_t*)(base + L.off_hits);
  c.ego_last = c.ego_now + L.W;
  c.bits = (uint32_t*)(base + L.off_bits);
  c.acc = (int*)(base + L.off_acc);
  c.cold_d = (double*)(base + L.off_cold);
  c.cold_i = (int*)(base + L.off_cold + COLD_ND * sizeof(double));
  c.gmisc = (int*)(base + L.off_gmisc);
  c.unif = (double*)(base + L.off_unif);
  c.egoc = (double*)(base + L.off_egoc);
  c.ring = (double*)(base + L.off_ring);
}

SG_DEV void cta_sync() { __syncthreads(); }

// ---------------------------------------------------------------------------------
// Cell grid over the scenario's present entities (same cells, cell size and "large" rule as the
// general kernel's grid): every cell heads a linked list of the slots whose position falls into it
// (`ghead[cell]`, `gnext[slot]`; 0xffff ends a list).  Inserting is one atomic exchange on the head,
// so the grid needs no scan / scatter passes and no barriers of its own: the heads are cleared at
// the start of a tick (nobody reads the grid in phase A), every owner inserts its slots at the end of
// phase B, and the barrier that ends phase B publishes the lists.
#define CR_NIL 0xffffu

SG_DEV void crowd_grid_clear(uint16_t* ghead) {
  uint4* w = (uint4*)ghead;
  for (int q = threadIdx.x; q < SG_GRID_CELLS / 8; q += blockDim.x) w[q] = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
}

// 16-bit exchange on a shared array through the containing 32-bit word
SG_DEV uint32_t crowd_exch16(uint16_t* arr, int idx, uint32_t v) {
  uint32_t* w = (uint32_t*)arr + (idx >> 1);
  const int sh = (idx & 1) * 16;
  uint32_t old = *w, assumed;
  do {
    assumed = old;
    old = atomicCAS(w, assumed, (assumed & ~(0xffffu << sh)) | (v << sh));
  } while (old != assumed);
  return (old >> sh) & 0xffffu;
}

// insert slot s (present) at its position
SG_DEV void crowd_grid_insert(const Crowd& c, int s, double gox, double goy, double inv_cs) {
  const double px = c.state[s] - gox, py = c.state[c.G + s] - goy;
  const int ix = __double2int_rd(px * inv_cs), iy = __double2int_rd(py * inv_cs);
  const int cell = ((iy & (SG_GRID_DIM - 1)) << SG_GRID_BITS) | (ix & (SG_GRID_DIM - 1));
  c.gnext[s] = (uint16_t)crowd_exch16(c.ghead, cell, (uint32_t)s);
}

// corners of slot s as entity/base.py:100-138 computes them, from the State rows
SG_DEV void crowd_corners(const double* __restrict__ state, const double* __restrict__ hcs, const double* __restrict__ box,
                          int64_t nm, int64_t i0, int G, int s, double q[8]) {
  const double x = state[s], y = state[G + s], cs = hcs[s], sn = hcs[G + s];
  const double bw = __ldg(box + i0 + s), bl = __ldg(box + nm + i0 + s);
  const double bcx = __ldg(box + 2 * nm + i0 + s), bcy = __ldg(box + 3 * nm + i0 + s);
  const double hx0 = bcx - 0.5 * bl, hx1 = bcx + 0.5 * bl;
  const double hy0 = bcy + 0.5 * bw, hy1 = bcy - 0.5 * bw;
  q[0] = x + (hx0 * cs + hy0 * -sn); q[1] = y + (hx0 * sn + hy0 * cs);
  q[2] = x + (hx1 * cs + hy0 * -sn); q[3] = y + (hx1 * sn + hy0 * cs);
  q[4] = x + (hx1 * cs + hy1 * -sn); q[5] = y + (hx1 * sn + hy1 * cs);
  q[6] = x + (hx0 * cs + hy1 * -sn); q[7] = y + (hx0 * sn + hy1 * cs);
}

// centre and half-edge vectors of slot s's box for the separating-axis filter (sat_classify_obb)
SG_DEV Obb crowd_obb(const double* __restrict__ state, const double* __restrict__ hcs, const double* __restrict__ box,
                     int64_t nm, int64_t i0, int G, int s) {
  const double x = state[s], y = state[G + s], cs = hcs[s], sn = hcs[G + s];
  const double hw = 0.5 * __ldg(box + i0 + s), hl = 0.5 * __ldg(box + nm + i0 + s);
  const double bcx = __ldg(box + 2 * nm + i0 + s), bcy = __ldg(box + 3 * nm + i0 + s);
  Obb o;
  o.cx = x + (bcx * cs - bcy * sn); o.cy = y + (bcx * sn + bcy * cs);
  o.ux = hl * cs; o.uy = hl * sn;
  o.vx = -hw * sn; o.vy = hw * cs;
  return o;
}

// exact narrow phase of one pair (`g != g_prime` exclusion of reference utils.py:58, then the closed-set
// predicate on the corners); only reached for knife-edge contacts
static __device__ __noinline__ bool crowd_pair_exact(const double* __restrict__ state, const double* __restrict__ hcs,
                                                     const double* __restrict__ box, int64_t nm, int64_t i0,
                                                     const int8_t* __restrict__ orient, int G, int a, int b) {
  double qa[8], qb[8];
  crowd_corners(state, hcs, box, nm, i0, G, a, qa);
  crowd_corners(state, hcs, box, nm, i0, G, b, qb);
  bool same = true;
#pragma unroll
  for (int f = 0; f < 8; ++f) same = same && (qa[f] == qb[f]);
  if (same) return false;
  const Quad A = quad_from_array(qa), B = quad_from_array(qb);
  const int oa = orient[a] ? orient[a] : quad_orientation(A), ob = orient[b] ? orient[b] : quad_orientation(B);
  return quads_intersect(A, oa, B, ob);
}

// one queued pair: the branch-free separating-axis filter decides all but knife-edge contacts
static __device__ __noinline__ bool crowd_pair_collides(const double* __restrict__ state, const double* __restrict__ hcs,
                                                        const double* __restrict__ box, int64_t nm, int64_t i0,
                                                        const int8_t* __restrict__ orient, int G, int a, int b) {
  const Obb A = crowd_obb(state, hcs, box, nm, i0, G, a), B = crowd_obb(state, hcs, box, nm, i0, G, b);
  const int v = sat_classify_obb(A, B);
  if (v != 0) return v > 0;
  return crowd_pair_exact(state, hcs, box, nm, i0, orient, G, a, b);
}

struct CrowdSink {  // where colliding pairs are booked (shared atomics)
  int* acc;
  uint32_t* bits;
  uint32_t* ego_now;
  uint32_t* rows;
  uint32_t* wqueue;  // this warp's strip of the pair queue and its fill count
  int* wqcount;
  int wqcap;
  int W, ego_slot, first_slot;
  bool need_first;  // no tick of this scenario had a collision yet: the smallest pair of the tick is wanted
};

// rare parts of booking a pair: the ego's row, the optional pair matrix
static __device__ __noinline__ void crowd_commit_rare(CrowdSink k, int lo, int hi) {
  if (lo == k.first_slot || hi == k.first_slot) k.acc[ACC_FIRST_HIT] = 1;
  if (lo == k.ego_slot) atomicOr(&k.ego_now[hi >> 5], 1u << (hi & 31));
  if (hi == k.ego_slot) atomicOr(&k.ego_now[lo >> 5], 1u << (lo & 31));
  if (k.rows) {
    atomicOr(&k.rows[(int64_t)lo * k.W + (hi >> 5)], 1u << (hi & 31));
    atomicOr(&k.rows[(int64_t)hi * k.W + (lo >> 5)], 1u << (lo & 31));
  }
}
// a colliding pair (crowds have hundreds per tick).  The pair count is kept by the calling thread and
// added to the scenario's once per tick; the `collided` bits are sticky over the launch (a slot's
// bit is set by its first collision only); the smallest pair is only wanted until the first tick
// with a collision is on record.
SG_DEV void crowd_commit(const CrowdSink& k, int a, int b, int& mypairs) {
  const int lo = min(a, b), hi = max(a, b);
  ++mypairs;
  if (k.need_first) atomicMin(&k.acc[ACC_FIRST_PAIR], (lo << 16) | hi);
  if (!((k.bits[lo >> 5] >> (lo & 31)) & 1u)) atomicOr(&k.bits[lo >> 5], 1u << (lo & 31));
  if (!((k.bits[hi >> 5] >> (hi & 31)) & 1u)) atomicOr(&k.bits[hi >> 5], 1u << (hi & 31));
  if (k.rows || lo == k.ego_slot || hi == k.ego_slot || lo == k.first_slot || hi == k.first_slot)
    crowd_commit_rare(k, lo, hi);
}

struct CrowdBoxes {  // what the narrow phase needs to rebuild corners (DIRECT mode only)
  const double* box;
  int64_t nm, i0;
};

// One pair (a < b) whose squared centre distance is d2, decided by circles where they decide --
// centres closer than the sum of the inscribed radii: the boxes intersect; further apart than the sum
// of the circumscribed radii: they cannot -- with a 1e-6 relative margin, nine orders of magnitude
// above the rounding of the corner coordinates; everything else is queued for the separating-axis /
// exact predicate (DIRECT: decided in place -- the redo after a queue overflow, where the pairs the
// circles decide are already on record).
template <bool DIRECT>
SG_DEV void crowd_pair(const Crowd& c, const CrowdSink& sink, const CrowdBoxes& bx, int a, int b, double d2, int& mypairs) {
  const float2 ra = c.rad[a], rb = c.rad[b];
  const double rin = (double)ra.x + (double)rb.x, rout = (double)ra.y + (double)rb.y;
  if (d2 > rout * rout) return;
  if (d2 < rin * rin && d2 > 0.0) {
    if (!DIRECT) crowd_commit(sink, a, b, mypairs);
    return;
  }
  if (!DIRECT) {
    const int q = atomicAdd(sink.wqcount, 1);
    if (q < sink.wqcap) sink.wqueue[q] = ((uint32_t)a << 16) | (uint32_t)b;
  } else if (crowd_pair_collides(c.state, c.fcs, bx.box, bx.nm, bx.i0, c.orient, c.G, a, b)) {
    crowd_commit(sink, a, b, mypairs);
  }
}

// fn(o) for every slot listed in the 3 x 3 cells around cell (ix, iy)
template <typename F>
SG_DEV void crowd_cells(const Crowd& c, int ix, int iy, F&& fn) {
  const int cx0 = (ix - 1) & (SG_GRID_DIM - 1), cx1 = ix & (SG_GRID_DIM - 1), cx2 = (ix + 1) & (SG_GRID_DIM - 1);
  uint32_t o = c.ghead[(((iy - 1) & (SG_GRID_DIM - 1)) << SG_GRID_BITS) | cx0];
#pragma unroll 1
  for (int q = 0; q < 9; ++q) {
    // the head of the next cell is fetched while this cell's list is walked
    uint32_t nxt = CR_NIL;
    if (q < 8) {
      const int q1 = q + 1, row = (iy + q1 / 3 - 1) & (SG_GRID_DIM - 1), col = q1 % 3;
      nxt = c.ghead[(row << SG_GRID_BITS) | (col == 0 ? cx0 : (col == 1 ? cx1 : cx2))];
    }
    for (; o != CR_NIL; o = c.gnext[o]) fn(o);
    o = nxt;
  }
}

// the tests of slot s (at px, py) against listed slot o: next tick's sensor list and this tick's pair
struct CrowdSlot {
  double px, py, r2;
  int s, ncand;
  bool sensor, coll;
};
template <bool DIRECT>
SG_DEV void crowd_item(const Crowd& c, const CrowdSink& sink, const CrowdBoxes& bx, CrowdSlot& w, uint32_t o, int& mypairs) {
  const int G = c.G;
  const double ddx = c.state[o] - w.px, ddy = c.state[G + o] - w.py, d2 = ddx * ddx + ddy * ddy;
  const uint8_t fo = c.flags[o];
  if (!DIRECT && w.sensor && ((fo >> 1) & 3) == SG_ETYPE_PEDESTRIAN && !(d2 > w.r2)) {
    if (w.ncand < CR_NBCAP) c.nbl[w.ncand * G + w.s] = (uint16_t)o;
    ++w.ncand;
  }
  if (w.coll && o > (uint32_t)w.s && !(fo & 8)) crowd_pair<DIRECT>(c, sink, bx, w.s, (int)o, d2, mypairs);
}

// ---------------------------------------------------------------------------------
// One pass over the 3 x 3 cells around present slot s on the grid of the positions just computed serves
//   * the pedestrian's sensor of the NEXT tick (neighbours strictly inside the 64-gon of
//     circumradius r lie in these cells): pedestrians within r (1 + 1e-9) go to the slot's
//     candidate list, which is then put in slot (= state.poses) order, and
//   * this tick's collision broad phase: pairs (s, o), o > s, of entities small enough for the grid
//     (circumscribed radius <= half a cell: their boxes can only meet from neighbouring cells).
// The ragged part -- chasing the cells' lists -- only copies ids into the thread's strip of `pool`;
// the tests then run over the strip in a plain counted loop.
// `r2` = (r (1 + 1e-9))^2; sensor = the slot is a present pedestrian; coll = it takes part in the
// cell broad phase (present, not large).
// Called by all 32 lanes of a warp (`active`: the slot is present and has something to test): the
// lanes leave the ragged walk at different times and are brought together again before the counted
// loop and before the sorting network, which would otherwise run once per straggler group.
template <bool DIRECT>
SG_DEV void crowd_scan(const Crowd& c, const CrowdSink& sink, const CrowdBoxes& bx, int s, bool active, bool sensor,
                       bool coll, double r2, double gox, double goy, double inv_cs, uint16_t* mypool, int& mypairs) {
  const int G = c.G;
  CrowdSlot w;
  w.px = c.state[s]; w.py = c.state[G + s]; w.r2 = r2;
  w.s = s; w.ncand = 0; w.sensor = sensor; w.coll = coll;
  const int ix = __double2int_rd((w.px - gox) * inv_cs), iy = __double2int_rd((w.py - goy) * inv_cs);
  int k = 0;
  if (active)
    crowd_cells(c, ix, iy, [&](uint32_t o) {
      if (o != (uint32_t)s) {
        if (k < CR_CAP) mypool[k * 32] = (uint16_t)o;
        ++k;
      }
    });
  __syncwarp();
  const int kk = min(k, CR_CAP), kmax = __reduce_max_sync(0xffffffffu, kk);
  for (int j = 0; j < kmax; ++j)
    if (j < kk) crowd_item<DIRECT>(c, sink, bx, w, (uint32_t)mypool[j * 32], mypairs);
  if (k > CR_CAP) {  // (a very dense neighbourhood: the ids beyond the strip are tested while walking)
    int j = 0;
    crowd_cells(c, ix, iy, [&](uint32_t o) {
      if (o != (uint32_t)s && j++ >= CR_CAP) crowd_item<DIRECT>(c, sink, bx, w, o, mypairs);
    });
  }
  __syncwarp();
  if (!DIRECT && sensor) {
    c.ncnt[s] = (uint8_t)min(w.ncand, 255);
    // slot (= state.poses) order: a sorting network over the (at most CR_NBCAP) listed ids
    const int n = min(w.ncand, CR_NBCAP);
    if (n > 1) {
      uint32_t v[CR_NBCAP];
#pragma unroll
      for (int a = 0; a < CR_NBCAP; ++a) v[a] = a < n ? (uint32_t)c.nbl[a * G + s] : 0xffffu;
#define CR_CE(a, b) { const uint32_t lo_ = min(v[a], v[b]), hi_ = max(v[a], v[b]); v[a] = lo_; v[b] = hi_; }
      CR_CE(0, 2) CR_CE(1, 3) CR_CE(4, 6) CR_CE(5, 7)
      CR_CE(0, 4) CR_CE(1, 5) CR_CE(2, 6) CR_CE(3, 7)
      CR_CE(0, 1) CR_CE(2, 3) CR_CE(4, 5) CR_CE(6, 7)
      CR_CE(2, 4) CR_CE(3, 5)
      CR_CE(1, 4) CR_CE(3, 6)
      CR_CE(1, 2) CR_CE(3, 4) CR_CE(5, 6)
#undef CR_CE
#pragma unroll
      for (int a = 0; a < CR_NBCAP; ++a)
        if (a < n) c.nbl[a * G + s] = (uint16_t)v[a];
    }
  }
}

// the pairs of present slot s this tick: through the grid (crowd_scan) when s is small, against the
// list of large entities, or -- too many large entities for the list -- against every later slot
// (called by all lanes of a warp; fl = the slot's flags, bit 0: present)
template <bool DIRECT>
SG_DEV void crowd_slot_pairs(const Crowd& c, const CrowdSink& sink, const CrowdBoxes& bx, int s, uint8_t fl,
                             bool need_coll, bool exhaustive, double r2, double gox, double goy, double inv_cs,
                             uint16_t* mypool, int& mypairs) {
  const int G = c.G;
  const bool pres = (fl & 1) != 0, large = (fl & 8) != 0;
  const bool sensor = !DIRECT && pres && ((fl >> 1) & 3) == SG_ETYPE_PEDESTRIAN;
  const bool coll = pres && need_coll && !large && !exhaustive;
  crowd_scan<DIRECT>(c, sink, bx, s, sensor || coll, sensor, coll, r2, gox, goy, inv_cs, mypool, mypairs);
  if (!need_coll || !pres) return;
  const double px = c.state[s], py = c.state[G + s];
  if (!exhaustive) {  // the large entities are kept in a list everyone tests against
    const int nl = c.gmisc[0];
    for (int k = 0; k < nl; ++k) {
      const int o = c.glarge[k];
      if (o == s || (large && o < s) || !(c.flags[o] & 1)) continue;
      const double ddx = c.state[o] - px, ddy = c.state[G + o] - py;
      crowd_pair<DIRECT>(c, sink, bx, min(s, o), max(s, o), ddx * ddx + ddy * ddy, mypairs);
    }
  } else {
    for (int o = s + 1; o < c.M; ++o) {
      if (!(c.flags[o] & 1)) continue;
      const double ddx = c.state[o] - px, ddy = c.state[G + o] - py;
      crowd_pair<DIRECT>(c, sink, bx, s, o, ddx * ddx + ddy * ddy, mypairs);
    }
  }
}

// ---------------------------------------------------------------------------------
// ReplayTrajectoryAgent slot (the ego): rare kind, rows kept in global memory
struct CrowdReplayOut { double x, y, h, vx, vy, speed3; bool present; };
static __device__ __noinline__ CrowdReplayOut crowd_replay_step(const SgScene sc, const SgState st, int64_t i, int64_t nm,
                                                               bool present, double t, double next_t) {
  CrowdReplayOut o;
  const int64_t r0 = sc.traj_off[i];
  const int K = (int)(sc.traj_off[i + 1] - r0);
  const double* rows = sc.traj_rows + r0 * 7;
  double np_[6], prev[6];
  int cur = st.cur_own[i];
  o.present = false;
  o.x = st.pose[i]; o.y = st.pose[nm + i]; o.h = st.pose[3 * nm + i];
  o.vx = st.vel[i]; o.vy = st.vel[nm + i]; o.speed3 = 0.0;
  if (!present && !(K > 0 && __ldg(rows) >= t)) return o;  // scenario_gym.py:240-244
  position_at_t(rows, K, next_t, EXT_CLAMP, cur, np_);      // agent.py:125-128
  if (present) {
#pragma unroll
    for (int f = 0; f < 6; ++f) prev[f] = st.pose[f * nm + i];
  } else {  // state.py:219-222
    int c2 = 0;
    position_at_t(rows, K, t, EXT_TRUE, c2, prev);
  }
  const double dt = next_t - t;
  double d[6], v[6];
#pragma unroll
  for (int f = 0; f < 6; ++f) { d[f] = np_[f] - prev[f]; v[f] = d[f] / dt; }
#pragma unroll
  for (int f = 0; f < 6; ++f) { st.pose[f * nm + i] = np_[f]; st.vel[f * nm + i] = v[f]; }
  st.dist[i] += norm3(d[0], d[1], d[2]);
  st.cur_own[i] = cur;
  o.present = true;
  o.x = np_[0]; o.y = np_[1]; o.h = np_[3]; o.vx = v[0]; o.vy = v[1];
  o.speed3 = norm3(v[0], v[1], v[2]);
  return o;
}

// The ego's ReplayTrajectoryAgent step with its State row cached on chip: x, y in the shared State rows,
// heading / distance / control-point cursor in the owner's registers, z / p / r and the six velocities in
// shared memory (`eg`); the row goes back to global memory once, after the last tick.  Same arithmetic as
// crowd_replay_step; one lane of one warp runs it, so every L2 round trip it saves is taken off the tick's
// critical path (the other warps wait for that warp at the barrier).
enum { EG_Z = 0, EG_P, EG_R, EG_H, EG_V0, EG_N = EG_V0 + 6 };
SG_DEV CrowdReplayOut crowd_replay_cached(const double* __restrict__ rows, int K, double* eg, bool present, double px,
                                          double py, double& h, double& dist, int& cur, double t, double next_t) {
  CrowdReplayOut o;
  o.present = false;
  o.x = px; o.y = py; o.h = h; o.vx = eg[EG_V0]; o.vy = eg[EG_V0 + 1]; o.speed3 = 0.0;
  if (!present && !(K > 0 && __ldg(rows) >= t)) return o;  // scenario_gym.py:240-244
  double np_[6], prev[6];
  position_at_t(rows, K, next_t, EXT_CLAMP, cur, np_);  // agent.py:125-128
  if (present) {
    prev[0] = px; prev[1] = py; prev[2] = eg[EG_Z]; prev[3] = h; prev[4] = eg[EG_P]; prev[5] = eg[EG_R];
  } else {  // state.py:219-222
    int c2 = 0;
    position_at_t(rows, K, t, EXT_TRUE, c2, prev);
  }
  const double dt = next_t - t;
  double d[6], v[6];
#pragma unroll
  for (int f = 0; f < 6; ++f) { d[f] = np_[f] - prev[f]; v[f] = d[f] / dt; eg[EG_V0 + f] = v[f]; }
  eg[EG_Z] = np_[2]; eg[EG_P] = np_[4]; eg[EG_R] = np_[5]; eg[EG_H] = np_[3];
  dist += norm3(d[0], d[1], d[2]);
  o.present = true;
  o.x = np_[0]; o.y = np_[1]; o.h = np_[3]; o.vx = v[0]; o.vy = v[1];
  o.speed3 = norm3(v[0], v[1], v[2]);
  h = np_[3];
  return o;
}

// The ego's next 32 steps at once: lane l of one warp evaluates the tick l + 1 ticks ahead -- the tick times
// by the same repeated addition as the tick loop, the pose at each, and against the pose one lane below
// (lane 0: the ego's current row, `pose`) the velocities, the distance increment, |v| and cos / sin of the
// heading, exactly as crowd_replay_step computes them tick by tick.  The owner of the ego's row then only
// copies an entry per tick: a replay step run by one lane of one warp kept the scenario's other 31 warps
// waiting at the barrier for 8 % of the tick.  Valid while the ego stays present (it does, once it is).
enum { ER_X = 0, ER_Y, ER_Z, ER_H, ER_P, ER_R, ER_V0, ER_SPEED3 = ER_V0 + 6, ER_INC, ER_CS, ER_SN, ER_CUR, ER_N = 18 };
static __device__ __noinline__ void crowd_ego_ring(const double* __restrict__ rows, int K, const double* xy, int G, int ego,
                                                   const double* eg, double t, double timestep, double* ring, int lane) {
  double told = t;
  for (int q = 0; q < lane; ++q) told = told + timestep;  // scenario_gym.py:229, tick by tick
  const double tnew = told + timestep;
  double np_[6], prev[6];
  int cur = 0;
  position_at_t(rows, K, tnew, EXT_CLAMP, cur, np_);  // agent.py:125-128
#pragma unroll
  for (int f = 0; f < 6; ++f) prev[f] = __shfl_up_sync(0xffffffffu, np_[f], 1);
  if (lane == 0) {
    prev[0] = xy[ego]; prev[1] = xy[G + ego]; prev[2] = eg[EG_Z]; prev[3] = eg[EG_H]; prev[4] = eg[EG_P]; prev[5] = eg[EG_R];
  }
  const double dt = tnew - told;
  double* e = ring + lane * ER_N;
  double d[6], v[6];
#pragma unroll
  for (int f = 0; f < 6; ++f) { d[f] = np_[f] - prev[f]; v[f] = d[f] / dt; e[ER_X + f] = np_[f]; e[ER_V0 + f] = v[f]; }
  e[ER_INC] = norm3(d[0], d[1], d[2]);
  e[ER_SPEED3] = norm3(v[0], v[1], v[2]);
  double sn, cs;
  sincos(np_[3], &sn, &cs);
  e[ER_CS] = cs; e[ER_SN] = sn;
  e[ER_CUR] = (double)cur;
}

// a pedestrian that is not in the scene yet is inserted at its first control point when its
// trajectory starts at or after t (scenario_gym.py:240-244); returns false otherwise
static __device__ __noinline__ bool crowd_ped_appears(const SgScene sc, int64_t i, double t, double next_t,
                                                      double np_[6], double prev[6]) {
  const int64_t r0 = sc.traj_off[i];
  const int K = (int)(sc.traj_off[i + 1] - r0);
  const double* rows = sc.traj_rows + r0 * 7;
  if (!(K > 0 && __ldg(rows) >= t)) return false;
  int cur = 1, c2 = 0;
  position_at_t(rows, K, next_t, EXT_CLAMP, cur, np_);
  position_at_t(rows, K, t, EXT_TRUE, c2, prev);  // state.py:219-222
  return true;
}

// ---------------------------------------------------------------------------------
// per-slot values only the owner thread needs between ticks (registers; the loops over a thread's
// slots copy one PerSlot in and out with constant indices, so the array never goes to local memory)
struct PerSlot {
  int kind, goal;
  double h, dist;
  uint32_t bits;  // bit 0: was in a collision so far, bit 1: moved during this launch
};
enum { CR_EGO_SPEED = COLD_T0, CR_EGO_DIST = COLD_T1 };
enum { U_INV_CS = 0, U_SIGHT, U_SH, U_CH, U_OX, U_OY, U_LEN };  // the ego's |v| and distance, for the metrics

#ifndef CR_MINB
#define CR_MINB 1  // one scenario per SM at 128 registers: no spills, and shared memory for the double buffers
#endif
template <int EPT, int CR_THREADS>
__global__ void __launch_bounds__(CR_THREADS, CR_MINB)
sg_crowd_kernel(SgScene sc, SgParams p, SgState st, SgInputs in, int n_ticks) {
  extern __shared__ __align__(16) unsigned char smem[];
  constexpr CrowdLayout L = crowd_layout(EPT, CR_THREADS);
  constexpr int CR_WARPS = CR_THREADS / 32;
  const int n = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
  const int M = sc.n_slots, G = L.G, W = L.W, WM = (M + 31) / 32;
  Crowd c;
  crowd_views(c, L, smem, M);
  Grp g;  // the view neighbour_force reads: old x, y, vx, vy of every slot
  g.pedbuf = c.state; g.G = G;
  const int64_t nm = sc.plane_stride, i0 = (int64_t)n * M;
  const int ego_slot = sc.ego_slot[n], first_slot = sc.first_slot[n];
  const bool need_coll = (p.features & SG_FEAT_COLLISIONS) || (p.terminal & (SG_TERM_COLLISION | SG_TERM_EGO_COLLISION));
  const bool matrix = (p.features & SG_FEAT_COLL_MATRIX) != 0;
  const bool ego_cached = sc.kind[i0 + ego_slot] == SG_KIND_AGENT_REPLAY;  // its row lives on chip
  const double grid_cs = p.ped_distance_threshold * (1.0 + 1e-6);
  if (tid == 0) {  // launch-uniform doubles live in shared memory: 64 registers per thread are all there is
    double sh, ch;
    sincos(p.ped_head_rot_angle, &sh, &ch);  // viewer/utils.py:6-17
    const int64_t er = sc.traj_off[i0 + ego_slot];  // origin of the fp32 bounds / the grid: the ego's first control point
    c.unif[U_INV_CS] = 1.0 / grid_cs;
    c.unif[U_SIGHT] = cos(p.sf_sight_angle / 2 * M_PI / 180);
    c.unif[U_SH] = sh; c.unif[U_CH] = ch;
    c.unif[U_OX] = __ldg(sc.traj_rows + er * 7 + 1);
    c.unif[U_OY] = __ldg(sc.traj_rows + er * 7 + 2);
    c.unif[U_LEN] = sc.length[n];
    c.gmisc[0] = 0;  // entities too large for the grid (a property of the box: listed once, below)
    c.gmisc[6] = -1; // chunk of 32 ticks the ego's ring holds (none)
  }
  __syncthreads();
#define grid_inv_cs (c.unif[U_INV_CS])
#define sight_cos (c.unif[U_SIGHT])
#define sh_rot (c.unif[U_SH])
#define ch_rot (c.unif[U_CH])
#define ox (c.unif[U_OX])
#define oy (c.unif[U_OY])
#define length (c.unif[U_LEN])

  // ---- load the rows ----------------------------------------------------------------------
  PerSlot ent[EPT];
#pragma unroll
  for (int e = 0; e < EPT; ++e) {
    const int s = tid + e * CR_THREADS;
    PerSlot me;
    me.kind = SG_KIND_EMPTY; me.goal = 0; me.h = 0; me.dist = 0; me.bits = 0;
    uint8_t fl = 0;
    double x = 0, y = 0, vx = 0, vy = 0;
    float2 rd = make_float2(0.f, 0.f);
    int8_t oh = 0;
    if (s < M) {
      const int64_t i = i0 + s;
      me.kind = sc.kind[i];
      if (me.kind != SG_KIND_EMPTY) {
        x = st.pose[i]; y = st.pose[nm + i]; me.h = st.pose[3 * nm + i];
        vx = st.vel[i]; vy = st.vel[nm + i];
        me.dist = st.dist[i];
        me.goal = st.goal_idx[i];
        me.bits = st.collided[i] ? 1u : 0u;
        fl = (uint8_t)((st.present[i] ? 1 : 0) | (sc.etype[i] << 1));
        const double bw = sc.box[i], bl = sc.box[nm + i], bcx = sc.box[2 * nm + i], bcy = sc.box[3 * nm + i];
        oh = (int8_t)box_orientation_hint(bw, bl);
        // circles about the pose position: inscribed only when the box is centred on it
        const double aw = fabs(bw), al = fabs(bl), off = sqrt(bcx * bcx + bcy * bcy);
        const double rin = (bcx == 0.0 && bcy == 0.0) ? 0.5 * fmin(aw, al) * (1.0 - 1e-6) : 0.0;
        const double rout = (0.5 * sqrt(aw * aw + al * al) + off) * (1.0 + 1e-6);
        rd = make_float2(__double2float_rd(rin), __double2float_ru(rout));
        // the box stays within rout of the pose position: up to half a cell, boxes only meet from neighbouring cells
        if (!((double)rd.y <= 0.5 * grid_cs * (1.0 - 1e-6))) {
          fl |= 8;
          const int k = atomicAdd(&c.gmisc[0], 1);
          if (k < SG_GRID_LCAP) c.glarge[k] = (uint16_t)s;
        }
        if (s == ego_slot) {
          c.cold_d[CR_EGO_SPEED] = norm3(vx, vy, st.vel[2 * nm + i]);
          c.cold_d[CR_EGO_DIST] = me.dist;
          if (me.kind == SG_KIND_AGENT_REPLAY) {  // its row stays on chip (crowd_replay_cached)
            me.goal = st.cur_own[i];
            c.egoc[EG_Z] = st.pose[2 * nm + i]; c.egoc[EG_P] = st.pose[4 * nm + i]; c.egoc[EG_R] = st.pose[5 * nm + i];
            c.egoc[EG_H] = me.h;
#pragma unroll
            for (int f = 0; f < 6; ++f) c.egoc[EG_V0 + f] = st.vel[f * nm + i];
          }
        }
      }
    }
    c.state[s] = x; c.state[G + s] = y; c.state[2 * G + s] = vx; c.state[3 * G + s] = vy;
    c.flags[s] = fl;
    c.rad[s] = rd;
    c.orient[s] = oh;
    ent[e] = me;
  }
  double t = st.t[n], prev_t = st.prev_t[n];
  int tick = st.tick[n];
  bool done = st.done[n] != 0;
  if (tid == 0) {
    c.cold_i[COLD_FIRST_TICK] = st.first_coll_tick[n];
    c.cold_i[COLD_FP0] = st.first_coll_pair[2 * n]; c.cold_i[COLD_FP1] = st.first_coll_pair[2 * n + 1];
    *(long long*)(c.cold_i + COLD_PAIR_TICKS) = st.n_pair_ticks[n];
    for (int q = 0; q < 2 * ACC_N; ++q) c.acc[q] = 0;
    c.acc[ACC_FIRST_PAIR] = 0x7fffffff; c.acc[ACC_N + ACC_FIRST_PAIR] = 0x7fffffff;
    c.cold_d[COLD_AVG] = st.ego_avg_speed[n]; c.cold_d[COLD_AVG_T] = st.ego_avg_t[n];
    c.cold_d[COLD_MAX] = st.ego_max_speed[n]; c.cold_d[COLD_EGOD] = st.ego_dist[n];
  }
  if (tid < W) {
    c.ego_last[tid] = tid < WM ? st.ego_hits[(int64_t)n * WM + tid] : 0u;
    c.ego_now[tid] = 0;
    c.bits[tid] = 0;
  }
  crowd_grid_clear(c.ghead);
  crowd_grid_clear(c.ghead + SG_GRID_CELLS);
  cta_sync();
#pragma unroll 1
  for (int e = 0; e < EPT; ++e) {  // the grid of the loaded positions: the sensors of the first tick
    const int s = tid + e * CR_THREADS;
    if (c.flags[s] & 1) crowd_grid_insert(c, s, ox, oy, grid_inv_cs);
  }
  cta_sync();
  uint16_t* const mypool = c.pool + (tid >> 5) * (CR_CAP * 32) + lane;  // this thread's strip: mypool[k * 32]
  CrowdBoxes bx;
  bx.box = sc.box; bx.nm = nm; bx.i0 = i0;
  const double rr_sensor = p.ped_distance_threshold * (1.0 + 1e-9);
  constexpr int QW = L.QCAP / CR_WARPS;  // pair queue strip of a warp
  CrowdSink sink;
  sink.bits = c.bits; sink.ego_now = c.ego_now;
  sink.rows = matrix ? st.coll_mask + (int64_t)n * M * WM : nullptr;
  sink.wqueue = c.queue + (tid >> 5) * QW; sink.wqcount = c.gmisc + 8 + (tid >> 5); sink.wqcap = QW;
  sink.W = WM; sink.ego_slot = ego_slot; sink.first_slot = first_slot;
  sink.acc = c.acc; sink.need_first = false;
  {
    int nopairs = 0;
#pragma unroll 1
    for (int e = 0; e < EPT; ++e) {
      const int s = tid + e * CR_THREADS;
      const uint8_t fl = c.flags[s];
      crowd_slot_pairs<false>(c, sink, bx, s, fl, false, false, rr_sensor * rr_sensor, ox, oy, grid_inv_cs, mypool, nopairs);
    }
  }
  cta_sync();

  const int limit = n_ticks < 0 ? p.max_ticks : n_ticks;
  int parity = 0;
  // The State rows / flags other threads read are double-buffered (`cur`: the buffer holding the current
  // State), and so are the heads of the cell grid (`gp`: the grid the last candidate phase read, cleared
  // during this tick while the owners insert into the other one): a tick needs two CTA barriers --
  // new rows + grid complete, pairs booked -- instead of one per phase.
  int cur = 0, gp = 0;

  for (int k = 0; k < limit && (!done || in.step_done); ++k) {
    const double next_t = t + p.timestep;  // scenario_gym.py:229
    const double step_dt = next_t - t;
    Crowd co = c, cn = c;  // views of the State before / after this tick's step
    co.state = c.state + cur * 4 * G;       co.flags = c.flags + cur * G;
    cn.state = c.state + (cur ^ 1) * 4 * G; cn.flags = c.flags + (cur ^ 1) * G;
    cn.ghead = c.ghead + (gp ^ 1) * SG_GRID_CELLS;
    g.pedbuf = co.state;
    crowd_grid_clear(c.ghead + gp * SG_GRID_CELLS);
    const double dt_state = t - prev_t;  // state.dt: PedestrianController uses the previous interval
    const double t_old = t;
    prev_t = t;
    t = next_t;
    tick += 1;
    const double dt = t - prev_t;
    const double rdt = 1.0 / dt;
#pragma unroll 1
    for (int e = 0; e < EPT; ++e) {
      const int s = tid + e * CR_THREADS;
      PerSlot me = ent[0];
      if (EPT > 1 && e == 1) me = ent[EPT - 1];
      const int64_t i = i0 + s;
      const uint8_t fl_old = co.flags[s];
      const bool present = (fl_old & 1) != 0;
      const double px = co.state[s], py = co.state[G + s];
      // ============ sensors + behaviour (reads the rows of other slots as they were before the tick) =====
      const bool is_ped = me.kind == SG_KIND_PEDESTRIAN && present;
      bool walking = false;
      double F0 = 0.0, F1 = 0.0;
      int ncand = 0;
      if (is_ped) {
        const int64_t r0 = sc.route_off[i];
        const int R = (int)(sc.route_off[i + 1] - r0);
        const double* route = sc.route_xy + 2 * r0;
        if (me.goal <= R - 1) {  // pedestrian/agent.py:60-62
          const double sarc = route_project(route, R, px, py);
          double arc = 0.0;
          int last = 0;
          for (int q = 0; q < R; ++q) {
            if (q > 0)
              arc += norm2(__ldg(route + 2 * q) - __ldg(route + 2 * q - 2),
                           __ldg(route + 2 * q + 1) - __ldg(route + 2 * q - 1));
            if (arc <= sarc) last = q;
          }
          me.goal = last + 1;
        }
        walking = me.goal <= R - 1;
        if (walking) {
          // SocialForce._force_to_goal, pedestrian/social_force.py:119-138
          const double speed_desired = sc.ped_speed_desired[i];
          const double dvx = __ldg(route + 2 * me.goal) - px, dvy = __ldg(route + 2 * me.goal + 1) - py;
          double dn = norm2(dvx, dvy);
          if (dn == 0) dn += 0.000000001;
          const double ux = dvx / dn, uy = dvy / dn;
          const double kk = 1 / p.sf_relaxation_time;
          F0 = kk * (speed_desired * ux - co.state[2 * G + s]);
          F1 = kk * (speed_desired * uy - co.state[3 * G + s]);
        }
      }
      // the sensor's candidates were listed by the last candidate phase, on the grid of these positions
      if (walking) ncand = c.ncnt[s];
      // neighbour terms, pooled over the warp's 32 pedestrians of this pass and handed back to their
      // owners in list order (same sums, same order as one lane walking its own list)
      const int nlist = (walking && ncand <= CR_NBCAP) ? ncand : 0;
      {
        const unsigned FULL = 0xffffffffu;
        const int wslot0 = s - lane;
        int incl = nlist;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const int v = __shfl_up_sync(FULL, incl, d);
          if (lane >= d) incl += v;
        }
        const int off = incl - nlist, total = __shfl_sync(FULL, incl, 31);
        __syncwarp();
        for (int base = 0; base < total; base += 32) {
          const int it = base + lane;
          int q = 0;
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
            const int o = c.nbl[(it - offq) * G + slot];
            nf = neighbour_force(p, g, co.state[slot], co.state[G + slot], o, step_dt, sh_rot, ch_rot, sight_cos);
          }
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
      }
      if (walking && ncand > CR_NBCAP) {  // very dense crowd: every present pedestrian in slot order
        for (int o = 0; o < M; ++o) {
          if (o == s || (co.flags[o] & 7) != (1 | (SG_ETYPE_PEDESTRIAN << 1))) continue;
          const NbForce nf = neighbour_force(p, g, px, py, o, step_dt, sh_rot, ch_rot, sight_cos);
          if (nf.valid) { F0 += nf.a0; F1 += nf.a1; F0 += nf.b0; F1 += nf.b1; }
        }
      }
      // ============ controller + State.step on the own row (written to the other buffer) =================
      bool newpres = false;
      double cs = 1.0, sn = 0.0, nx = px, ny = py, nvx = co.state[2 * G + s], nvy = co.state[3 * G + s];
      if (me.kind == SG_KIND_PEDESTRIAN) {
        double prevx = nx, prevy = ny, prevh = me.h, nh = me.h;
        bool moved = false;
        if (present) {
          double speed, heading;
          if (walking) {
            double speed_rand = p.sf_bias_lon, heading_rand = p.sf_bias_lat;
            if (p.sf_std_lon != 0.0 || p.sf_std_lat != 0.0) {  // engine-defined noise stream (sg_device.cuh)
              const double2 z = sg_noise2(p.sf_noise_seed, i + (int64_t)sc.scenario_base * M, tick - 1);
              speed_rand = p.sf_bias_lon + p.sf_std_lon * z.x;
              heading_rand = p.sf_bias_lat + p.sf_std_lat * z.y;
            }
            speed = py_min(norm2(F0, F1) + speed_rand, sc.ped_speed_desired[i] * p.sf_max_speed_factor);
            heading = atan2(F1, F0) + heading_rand;
          } else {  // agent.py:65-68
            speed = 0; heading = 0; F0 = 0.0; F1 = 0.0;
          }
          // PedestrianController._step, pedestrian/controller.py:38-46 (uses state.dt)
          const double sp = np_clip(speed, -p.ped_max_speed, p.ped_max_speed);
          st.speed[i] = sp;                            // pure outputs: written through, not kept
          st.force[i] = F0; st.force[nm + i] = F1;
          sincos(heading, &sn, &cs);
          nx = prevx + sp * dt_state * cs;
          ny = prevy + sp * dt_state * sn;
          nh = heading;
          moved = true;
          me.dist += norm3(nx - prevx, ny - prevy, 0.0);
        } else {
          double np_[6], pv[6];
          if (crowd_ped_appears(sc, i, t_old, t, np_, pv)) {
            // (rare) first appearance: all six components go through global memory
#pragma unroll
            for (int f = 0; f < 6; ++f) {
              st.pose[f * nm + i] = np_[f];
              st.vel[f * nm + i] = div_r(np_[f] - pv[f], dt, rdt);
            }
            me.dist += norm3(np_[0] - pv[0], np_[1] - pv[1], np_[2] - pv[2]);
            prevx = pv[0]; prevy = pv[1]; prevh = pv[3];
            nx = np_[0]; ny = np_[1]; nh = np_[3];
            sincos(nh, &sn, &cs);
            moved = true;
            if (s == ego_slot)
              c.cold_d[CR_EGO_SPEED] = norm3(div_r(nx - prevx, dt, rdt), div_r(ny - prevy, dt, rdt),
                                             div_r(np_[2] - pv[2], dt, rdt));
          }
        }
        if (moved) {
          const double ex = nx - prevx, ey = ny - prevy;
          nvx = div_r(ex, dt, rdt); nvy = div_r(ey, dt, rdt);
          st.vel[3 * nm + i] = div_r(nh - prevh, dt, rdt);
          me.h = nh;
          if (present) {
            me.bits |= 2u;  // z, p, r did not change: their velocities are 0 from now on
            if (s == ego_slot) c.cold_d[CR_EGO_SPEED] = norm3(nvx, nvy, 0.0);
          }
          if (s == ego_slot) c.cold_d[CR_EGO_DIST] = me.dist;
          newpres = true;
        }
      } else if (me.kind == SG_KIND_AGENT_REPLAY) {
        CrowdReplayOut r;
        bool have_cs = false;
        if (s == ego_slot && present && k > 0 && c.gmisc[6] == ((k - 1) >> 5)) {  // precomputed (crowd_ego_ring)
          const double* e = c.ring + ((k - 1) & 31) * ER_N;
          r.present = true;
          r.x = e[ER_X]; r.y = e[ER_Y]; r.h = e[ER_H]; r.vx = e[ER_V0]; r.vy = e[ER_V0 + 1]; r.speed3 = e[ER_SPEED3];
          me.dist += e[ER_INC];
          me.goal = (int)e[ER_CUR];
          me.bits |= 4u;
          cs = e[ER_CS]; sn = e[ER_SN];
          have_cs = true;
          c.egoc[EG_Z] = e[ER_Z]; c.egoc[EG_P] = e[ER_P]; c.egoc[EG_R] = e[ER_R]; c.egoc[EG_H] = e[ER_H];
#pragma unroll
          for (int f = 0; f < 6; ++f) c.egoc[EG_V0 + f] = e[ER_V0 + f];
        } else if (s == ego_slot) {  // cached rows: no round trips to global memory on the tick's critical path
          const int64_t r0 = sc.traj_off[i];
          r = crowd_replay_cached(sc.traj_rows + r0 * 7, (int)(sc.traj_off[i + 1] - r0), c.egoc, present, px, py, me.h,
                                  me.dist, me.goal, t_old, t);
          if (r.present) me.bits |= 4u;
        } else {
          r = crowd_replay_step(sc, st, i, nm, present, t_old, t);
        }
        newpres = r.present;
        if (newpres) {
          nx = r.x; ny = r.y; nvx = r.vx; nvy = r.vy;
          me.h = r.h;
          if (!have_cs) sincos(r.h, &sn, &cs);
          if (s == ego_slot) { c.cold_d[CR_EGO_SPEED] = r.speed3; c.cold_d[CR_EGO_DIST] = me.dist; }
        }
      }
      cn.state[s] = nx; cn.state[G + s] = ny; cn.state[2 * G + s] = nvx; cn.state[3 * G + s] = nvy;
      cn.flags[s] = (uint8_t)((fl_old & ~1) | (newpres ? 1 : 0));
      if (me.kind != SG_KIND_EMPTY) {
        if (newpres && need_coll) { c.fcs[s] = cs; c.fcs[G + s] = sn; }
        if (newpres) crowd_grid_insert(cn, s, ox, oy, grid_inv_cs);
        if (matrix) {
          uint32_t* row = st.coll_mask + ((int64_t)n * M + s) * WM;
          for (int w = 0; w < WM; ++w) row[w] = 0;
        }
      }
      if (e == 0) ent[0] = me; else ent[EPT - 1] = me;
    }
    if (lane == 0) *sink.wqcount = 0;
    cta_sync();  // the rows, the flags and the grid of the new positions are complete
    int* acc = c.acc + parity * ACC_N;
    sink.acc = acc;
    sink.need_first = c.cold_i[COLD_FIRST_TICK] < 0;  // (written in the tick epilogue only, behind the barriers)
    const bool exhaustive = need_coll && c.gmisc[0] > SG_GRID_LCAP;  // too many large entities for the list
    int mypairs = 0;
    if (ego_cached && (k & 31) == 0 && (tid >> 5) == ((k >> 5) & (CR_WARPS - 1)) && (cn.flags[ego_slot] & 1)) {
      // one warp (a different one every time) lays out the ego's next 32 steps
      const int64_t ie = i0 + ego_slot, r0 = sc.traj_off[ie];
      crowd_ego_ring(sc.traj_rows + r0 * 7, (int)(sc.traj_off[ie + 1] - r0), cn.state, G, ego_slot, c.egoc, t, p.timestep,
                     c.ring, lane);
      if (lane == 0) c.gmisc[6] = k >> 5;
    }
    {
      // ============ next tick's sensor candidates + broad phase; the warp then decides the pairs it queued =====
      const double r2s = rr_sensor * rr_sensor;
#pragma unroll 1
      for (int e = 0; e < EPT; ++e) {
        const int s = tid + e * CR_THREADS;
        crowd_slot_pairs<false>(cn, sink, bx, s, cn.flags[s], need_coll, exhaustive, r2s, ox, oy, grid_inv_cs, mypool, mypairs);
      }
      __syncwarp();
      const int nq = *sink.wqcount;
      if (nq <= QW) {
        for (int q = lane; q < nq; q += 32) {
          const uint32_t pr = sink.wqueue[q];
          const int a = (int)(pr >> 16), b = (int)(pr & 0xffff);
          if (crowd_pair_collides(cn.state, c.fcs, sc.box, nm, i0, c.orient, G, a, b)) crowd_commit(sink, a, b, mypairs);
        }
      } else {  // the warp's strip overflowed (very dense scenes): go over its candidates again and decide in place
#pragma unroll 1
        for (int e = 0; e < EPT; ++e) {
          const int s = tid + e * CR_THREADS;
          crowd_slot_pairs<true>(cn, sink, bx, s, cn.flags[s], need_coll, exhaustive, r2s, ox, oy, grid_inv_cs, mypool, mypairs);
        }
      }
      __syncwarp();
      {  // the colliding pairs this thread booked: one shared atomic per warp
        const int wsum = __reduce_add_sync(0xffffffffu, mypairs);
        if (lane == 0 && wsum) atomicAdd(&acc[ACC_NPAIRS], wsum);
      }
      cta_sync();
    }
    // ============ terminal check + metrics ==========================================================
    const int npairs = acc[ACC_NPAIRS];
    bool dn = false;  // state.py:268-270, 397-408
    if ((p.terminal & SG_TERM_MAX_LENGTH) && (t + dt > length)) dn = true;
    if ((p.terminal & SG_TERM_COLLISION) && npairs > 0) dn = true;
    if ((p.terminal & SG_TERM_EGO_COLLISION) && acc[ACC_FIRST_HIT]) dn = true;
    done = dn;
    if (tid == 32) {
      int* cold = c.cold_i;
      if (npairs > 0) {
        *(long long*)(cold + COLD_PAIR_TICKS) += npairs;
        if (cold[COLD_FIRST_TICK] < 0) {
          const int fp = acc[ACC_FIRST_PAIR];
          cold[COLD_FIRST_TICK] = tick; cold[COLD_FP0] = fp >> 16; cold[COLD_FP1] = fp & 0xffff;
        }
      }
      int* nx = c.acc + (parity ^ 1) * ACC_N;  // the other parity is reset for the next tick
      nx[ACC_NPAIRS] = 0; nx[ACC_FIRST_PAIR] = 0x7fffffff; nx[ACC_FIRST_HIT] = 0; nx[ACC_RSS] = 0; nx[ACC_QCOUNT] = 0;
    }
    if (tid >= 64 && tid < 64 + W) {  // CollisionMetric._step, metrics/collision.py:70-75
      const int w = tid - 64;
      const uint32_t now = c.ego_now[w];
      if (p.features & SG_FEAT_COLLISIONS) {
        const uint32_t fresh = now & ~c.ego_last[w];
        if (fresh) emit_events(st.events, st.event_count, st.event_cap, fresh, n + sc.scenario_base, tick, w * 32, t);
        c.ego_last[w] = now;
      }
      c.ego_now[w] = 0;
    }
    // (by the thread that owns the ego's row: it writes the ego's speed / distance in the next step phase,
    // which other warps may already have entered)
    if (tid == (ego_slot & (CR_THREADS - 1)) && (p.features & SG_FEAT_EGO_METRICS)) {  // metrics/trajectory.py:20-24,39-42,58-60
      double* m = c.cold_d;
      const double sp = m[CR_EGO_SPEED];
      const double w = m[COLD_AVG_T] / t;
      m[COLD_AVG] += (1.0 - w) * (sp - m[COLD_AVG]);
      m[COLD_AVG_T] = t;
      m[COLD_MAX] = fmax(sp, m[COLD_MAX]);
      m[COLD_EGOD] = m[CR_EGO_DIST];
    }
    parity ^= 1;
    cur ^= 1;
    gp ^= 1;
  }
  cta_sync();

  // ---------------- write the rows back --------------------------------------------------------------
  const double* rows = c.state + cur * 4 * G;
  const uint8_t* flags = c.flags + cur * G;
#pragma unroll
  for (int e = 0; e < EPT; ++e) {
    const int s = tid + e * CR_THREADS;
    const PerSlot me = ent[e];
    if (s >= M || me.kind == SG_KIND_EMPTY) continue;
    const int64_t i = i0 + s;
    st.present[i] = (flags[s] & 1) != 0;
    st.collided[i] = (uint8_t)((me.bits | (c.bits[s >> 5] >> (s & 31))) & 1u);  // before the launch or during it
    if (me.kind == SG_KIND_AGENT_REPLAY && (me.bits & 4u)) {  // the ego's cached row
      st.pose[i] = rows[s]; st.pose[nm + i] = rows[G + s]; st.pose[2 * nm + i] = c.egoc[EG_Z];
      st.pose[3 * nm + i] = me.h; st.pose[4 * nm + i] = c.egoc[EG_P]; st.pose[5 * nm + i] = c.egoc[EG_R];
#pragma unroll
      for (int f = 0; f < 6; ++f) st.vel[f * nm + i] = c.egoc[EG_V0 + f];
      st.dist[i] = me.dist;
      st.cur_own[i] = me.goal;
    }
    if (me.kind == SG_KIND_PEDESTRIAN) {
      st.pose[i] = rows[s]; st.pose[nm + i] = rows[G + s]; st.pose[3 * nm + i] = me.h;
      st.vel[i] = rows[2 * G + s]; st.vel[nm + i] = rows[3 * G + s];
      if (me.bits & 2u) { st.vel[2 * nm + i] = 0.0; st.vel[4 * nm + i] = 0.0; st.vel[5 * nm + i] = 0.0; }
      st.dist[i] = me.dist;
      st.goal_idx[i] = me.goal;
    }
  }
  if (tid == 0) {
    st.t[n] = t; st.prev_t[n] = prev_t; st.tick[n] = tick; st.done[n] = done;
    st.ego_avg_speed[n] = c.cold_d[COLD_AVG]; st.ego_avg_t[n] = c.cold_d[COLD_AVG_T];
    st.ego_max_speed[n] = c.cold_d[COLD_MAX]; st.ego_dist[n] = c.cold_d[COLD_EGOD];
    st.first_coll_tick[n] = c.cold_i[COLD_FIRST_TICK];
    st.first_coll_pair[2 * n] = c.cold_i[COLD_FP0]; st.first_coll_pair[2 * n + 1] = c.cold_i[COLD_FP1];
    st.n_pair_ticks[n] = *(long long*)(c.cold_i + COLD_PAIR_TICKS);
  }
  if (tid < WM) st.ego_hits[(int64_t)n * WM + tid] = c.ego_last[tid];
#undef grid_inv_cs
#undef sight_cos
#undef sh_rot
#undef ch_rot
#undef ox
#undef oy
#undef length
}

#ifndef CR_WIDE
#define CR_WIDE 1
#endif
cudaError_t sgi_launch_crowd(cudaStream_t s, const SgScene& sc, const SgParams& p, const SgState& st,
                             const SgInputs& in, int n_ticks) {
  // up to 512 slots: 512 threads, a slot each; up to 1024: 1024 threads at 64 registers (CR_WIDE) or 512
  // threads owning two slots each at 128 registers
  const bool big = sc.n_slots > 512;
  const int threads = (big && CR_WIDE) ? 1024 : 512;
  const CrowdLayout L = crowd_layout(big && !CR_WIDE ? 2 : 1, threads);
  auto kern = !big ? sg_crowd_kernel<1, 512> : (CR_WIDE ? sg_crowd_kernel<1, 1024> : sg_crowd_kernel<2, 512>);
  cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L.bytes);
  if (err != cudaSuccess) return err;
  kern<<<sc.n_scenarios, threads, L.bytes, s>>>(sc, p, st, in, n_ticks);
  return cudaGetLastError();
}
