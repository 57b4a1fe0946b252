// sg_layout.h -- per-scenario shared-memory layout of the tick-loop kernels (host + device).
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

#define SG_THREADS 256
#define SG_NBCAP 12  // per-pedestrian neighbour candidate list kept in shared memory
#ifndef SG_SORT_MIN_M
#define SG_SORT_MIN_M 129  // vehicle scenes with at least this many slots use the sorted sweep
#endif
#ifndef SG_SWEEP_WIN
#define SG_SWEEP_WIN 4  // successors tested branch-free by the sorted sweep (a longer run is walked); 3..5 measure alike on C5, 8: -2 %
#endif
#ifndef SG_SORT_WIN
#define SG_SORT_WIN 4  // positions either side within which the sorted sweep re-ranks a box in one pass
#endif
#ifndef SG_WARP_PAIRS
#define SG_WARP_PAIRS 6  // queued pairs per warp up to which the narrow phase runs warp-cooperatively
#endif
// uniform cell grid of a crowd scenario (one CTA per scenario): 64 x 64 cells, toroidal
#define SG_GRID_BITS 6
#define SG_GRID_DIM (1 << SG_GRID_BITS)
#define SG_GRID_CELLS (SG_GRID_DIM * SG_GRID_DIM)
#define SG_GRID_LCAP 64        // entities too large for the grid are kept in a list
#define SG_GRID_LARGE 0x8000u  // flag on a sorted slot id
#ifndef SG_VEH_THREADS
#define SG_VEH_THREADS 128  // CTA size of the vehicle kernel for scenarios of up to that many slots
#endif
#ifndef SG_VEH_MINB
#define SG_VEH_MINB 4  // resident CTAs per SM the vehicle kernel is compiled for
#endif

// ---------------------------------------------------------------------------------
// per-scenario shared-memory block
struct GroupLayout {
  int G;        // threads (slots incl. padding) per scenario
  int W;        // 32-bit words per collision row
  int H;        // half-sweep length M/2
  int QCAP;     // candidate-pair queue capacity
  int off_act, off_rbox, off_tcold, off_hcs, off_ped, off_pednb, off_nbl, off_box, off_ego, off_cold, off_aabb, off_queue, off_hits, off_bits,
      off_acc, off_flags, off_orient;
  int sorted;   // vehicle scenes with M >= 128: boxes kept sorted by their lower x bound, windowed sweep
  int off_sid, off_posof, off_sflag, off_skey, off_tmpbox;
  int grid;     // crowd scenario with a shared-memory cell grid (sensor + broad phase)
  int off_gstart, off_gsorted, off_glarge, off_gmisc;
  int bytes;
};

enum { EGO_X = 0, EGO_Y, EGO_C, EGO_S, EGO_INV0, EGO_INV1, EGO_HD0, EGO_HD1, EGO_HINV0, EGO_HINV1,
       EGO_V0, EGO_V1, EGO_VNORM, EGO_VLONG, EGO_W, EGO_L, EGO_RHW, EGO_RHL, EGO_PRESENT, EGO_N = 20 };
enum { COLD_AVG = 0, COLD_AVG_T, COLD_MAX, COLD_EGOD, COLD_T0, COLD_T1, COLD_PT0, COLD_PT1, COLD_LEN,
       COLD_OX, COLD_OY, COLD_R2MINA, COLD_ND = 12 };  // doubles (T/PT: tick time, 2 parities)
enum { COLD_FIRST_TICK = 0, COLD_FP0, COLD_FP1, COLD_RSS, COLD_PAIR_TICKS = 4, COLD_HAS_VEH = 6, COLD_NI = 8 };  // ints
enum { ACC_NPAIRS = 0, ACC_FIRST_PAIR, ACC_FIRST_HIT, ACC_RSS, ACC_QCOUNT, ACC_OFFROAD, ACC_N = 8 };

// (constexpr: kernels specialised for a slot count take their shared-memory offsets as immediates)
__host__ __device__ constexpr GroupLayout make_layout(int M, bool ped, bool rss, bool veh, bool grid = false) {
  GroupLayout L{};
  int G = 1;
  if (M <= 32) { while (G < M) G <<= 1; } else { G = (M + 31) / 32 * 32; }
  L.G = G;
  L.W = (M + 31) / 32;
  L.H = M / 2;
  L.QCAP = 4 * G;
  L.sorted = (veh && M >= SG_SORT_MIN_M && G > SG_VEH_THREADS) ? 1 : 0;
  int o = (L.sorted ? 2 : 1) * 8 * G * (int)sizeof(double);       // corners[8][G] (sorted vehicle kernels: one per tick parity)
  L.off_act = o;    o += veh ? 4 * G * (int)sizeof(double) : 0;             // VehicleAction rows, 2 stages x (accel, steer)
  L.off_rbox = o;   o += rss ? 8 * G * (int)sizeof(double) : 0;   // hazard corners in the ego frame
  L.off_tcold = o;  o += veh ? 6 * G * (int)sizeof(double) : 0;   // per-thread cold values: sd[2], ratio[2], vh, 1/length
  L.off_hcs = o;    o += rss ? 2 * G * (int)sizeof(double) : 0;   // cos/sin of each heading
  L.off_ped = o;    o += ped ? 4 * G * (int)sizeof(double) : 0;   // old x,y,vx,vy (pedestrian sensors)
  L.off_pednb = o;  o += ped ? (G + 32) * (int)sizeof(float4) : 0; // fp32 sensor boxes of the pedestrians (old state)
  L.off_nbl = o;    o += ped ? SG_NBCAP * G * (int)sizeof(uint16_t) : 0;
  o = (o + 15) / 16 * 16;
  L.off_box = o;    o += 4 * G * (int)sizeof(double);             // width, length, center_x, center_y
  L.off_ego = o;    o += EGO_N * (int)sizeof(double);
  L.off_cold = o;   o += COLD_ND * (int)sizeof(double) + COLD_NI * (int)sizeof(int);
  L.off_aabb = o;   o += (M + L.H + 1) * (int)sizeof(float4);     // duplicated head: no wrap in the sweep
  L.off_queue = o;  o += L.QCAP * (int)sizeof(uint32_t);
  L.off_hits = o;   o += 2 * L.W * (int)sizeof(uint32_t);         // ego_now[W], ego_last[W]
  L.off_bits = o;   o += 2 * L.W * (int)sizeof(uint32_t);         // collided bits, 2 parities
  L.off_acc = o;    o += 2 * ACC_N * (int)sizeof(int);            // 2 parities
  L.off_flags = o;  o += G + 16;                                  // old present|etype<<1
  L.off_orient = o; o += G;                                       // ring orientation of each box
  o = (o + 15) / 16 * 16;
  L.off_sid = o;    o += L.sorted ? (M + 64) * (int)sizeof(uint16_t) : 0;   // slot id at each sorted position
  L.off_posof = o;  o += L.sorted ? G * (int)sizeof(uint16_t) : 0;          // sorted position of each slot
  o = (o + 15) / 16 * 16;
  L.off_sflag = o;  o += L.sorted ? 4 * (int)sizeof(int) : 0;
  L.off_skey = o;   o += L.sorted ? (M + 2 * SG_SORT_WIN + 2) * (int)sizeof(float) : 0;  // keys by old position (+ sentinels)
  o = (o + 15) / 16 * 16;
  L.off_tmpbox = o; o += L.sorted ? G * (int)sizeof(float4) : 0;   // the owner's new box between publish and scatter
  L.grid = (grid && ped && G > SG_THREADS) ? 1 : 0;
  L.off_gstart = o;  o += L.grid ? (SG_GRID_CELLS / 2 + 4) * (int)sizeof(uint32_t) : 0;  // packed 16-bit cell starts (+ end)
  L.off_gsorted = o; o += L.grid ? G * (int)sizeof(uint16_t) : 0;                         // slot ids sorted by cell
  L.off_glarge = o;  o += L.grid ? SG_GRID_LCAP * (int)sizeof(uint16_t) : 0;
  L.off_gmisc = o;   o += L.grid ? 40 * (int)sizeof(int) : 0;                             // counters + 32 warp totals
  L.bytes = (o + 15) / 16 * 16;
  return L;
}
