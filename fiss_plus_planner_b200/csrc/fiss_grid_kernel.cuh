// fiss_grid_kernel.cuh -- the lattice ("grid") kernel: the hot path when the end states form the
// product lattice d_end x v_end x T that FrenetOptimalPlanner / FopPlusPlanner / FissPlanner sample
// (frenet_optimal_planner.py:72-101, fiss_planner.py:33-99).
//
// The reference evaluates every (d, v, T) candidate from scratch.  On the lattice most of that work
// is shared: the longitudinal quartic s(t) -- and with it the reference-line frame (spline segment
// search, position, unit tangent), the speed / acceleration masks, the truncation length n' and the
// longitudinal cost terms -- depends on (v_end, T) only; the lateral quintic d(t) and its cost
// terms depend on (d_end, T) only.  One work item = `slots` consecutive (ego state b, horizon T_k) pairs
// [or one pair and a chunk of its lateral rows when the batch is small]; their rows share the stages, the barriers,
// the per-step bounding boxes and the item bookkeeping -- per-item costs that one 54-candidate pair amortises
// badly.  A persistent CTA
//
//   stage 0  (once per CTA) bulk-TMA the spline table [mbarrier 0] and the centres of the checked obstacle rows
//            [mbarrier 1] into shared memory; stage A starts as soon as the spline has landed
//   stage A  one warp per row (dealt through a shared counter), two time steps per lane in lockstep:
//              longitudinal rows (slots * nv): quartic solve + evaluation, masks, cost terms, frame -> smem
//                                      tables P2 = (px, py), U2 = unit tangent, SD          polynomial.py:5-41
//              lateral rows (slots * rows)   : quintic solve + evaluation, cost terms -> smem table D polynomial.py:45-84
//            each longitudinal row folds its frame points into per-step bounding boxes (native 32-bit atomics)
//   stage A' proximity masks: for every (longitudinal row, checked step) a bit per obstacle whose
//            centre is within (max|d| + r_ego + r_obs) of the FRAME point -- a superset of the
//            obstacles any candidate on that row can touch at that step; obstacles are first tested against the
//            step's box, the survivors against the rows; the (row, step) pairs with any bit set are appended to a
//            compact work list
//   stage B  collision: one LANE per (lateral row, listed pair): ego pose from the tables, then the exact
//            predicate (circle reject + closed-set SAT, :168-195) for the listed obstacles only
//            materialisation (when asked): lanes = flattened (longitudinal row, step) elements, a task = (block of
//            31 elements, group of 3 lateral rows): x = PX - D*UY, y = PY + D*UX, heading / ds / kappa by finite
//            differences (frenet_optimal_planner.py:121-134) with the three rows' FP64 chains in lockstep, the five
//            output rows streamed to HBM
//   stage C  one lane per candidate: cost = (lon + lat terms)/n, flags word; reset of the per-item state, fetch of
//            the next item's ego states
//
// Work items are dealt with a grid stride or drawn through a device counter (big items first, single pairs last), and a
// launch can be CHAINED to the previous kernel of its stream -- scheduled as that kernel's CTAs retire, waiting for it only
// before its first output writes (GridArgs::dynamic, ::chained; fiss_abi.cu eval_grid says when).
//
// so a candidate costs ~2 table reads per step instead of two polynomial solves, a 7-step segment
// search and an M x n/2 obstacle sweep.  All arithmetic is FP64 with the same expressions as the
// generic kernel in fiss_kernels.cuh (masks, n' and winners are bit-identical between the two).
#pragma once

#include "fiss_kernels.cuh"
#include "fiss_math.cuh"

#ifdef FISS_PHASE_TIMING
// debug build only (tools/phase_timing.py): thread 0 of every CTA accumulates the cycles between stage boundaries
#define FISS_PHASE(k) do { if (threadIdx.x == 0) { const long long now_ = clock64(); phase_acc[k] += now_ - phase_t; phase_t = now_; } } while (0)
__device__ long long g_fiss_phase[16];
#else
#define FISS_PHASE(k) do { } while (0)
#endif

#ifdef FISS_TRACE
// debug build only (tools/warp_trace.py): lane 0 of every warp stamps clock64() at the stage boundaries of its first
// kTraceItems items -- [CTA][warp][item][stamp]; stamp 14 of item 0 is the kernel entry, stamp 15 the SM id
constexpr int kTraceCtas = 512, kTraceWarps = 16, kTraceItems = 8, kTraceStamps = 16;
__device__ long long g_fiss_trace[kTraceCtas * kTraceWarps * kTraceItems * kTraceStamps];
#define FISS_STAMP(k) do { if ((threadIdx.x & 31) == 0 && trace_item < kTraceItems && blockIdx.x < kTraceCtas) \
    g_fiss_trace[((blockIdx.x * kTraceWarps + (threadIdx.x >> 5)) * kTraceItems + trace_item) * kTraceStamps + (k)] = clock64(); } while (0)
#else
#define FISS_STAMP(k) do { } while (0)
#endif

// Output rows are written once and never read by the kernel: streaming stores (st.global.cs, evict-first) keep them
// from displacing the tables and the spilled registers in L1 / L2.
#ifdef FISS_PLAIN_STORES
#define FISS_ST(ptr, val) (*(ptr) = (val))
#else
#define FISS_ST(ptr, val) __stcs((ptr), (val))
#endif

namespace fiss {

// CTA shape per kernel variant (measured on the B200, cfg4): the winner-only kernel runs best as 8 warps x 3 CTAs per SM
// with two slots per item; the materialising kernel as 12 warps x 2 CTAs with three slots (its items carry 60-90
// materialisation tasks, which deal evenly over 12 warps).  Both are register-capped at 80.
#ifndef FISS_GRID_WARPS
#define FISS_GRID_WARPS 8
#endif
#ifndef FISS_GRID_MIN_CTAS
#define FISS_GRID_MIN_CTAS 3
#endif
#ifndef FISS_GRID_WARPS_MAT
#define FISS_GRID_WARPS_MAT 12
#endif
#ifndef FISS_GRID_MIN_CTAS_MAT
#define FISS_GRID_MIN_CTAS_MAT 2
#endif
__host__ __device__ constexpr int grid_warps(bool yaw) { return yaw ? FISS_GRID_WARPS_MAT : FISS_GRID_WARPS; }
__host__ __device__ constexpr int grid_min_ctas(bool yaw) { return yaw ? FISS_GRID_MIN_CTAS_MAT : FISS_GRID_MIN_CTAS; }  // resident CTAs per SM the register budget is capped for
__host__ __device__ constexpr int grid_slots(bool yaw) { return yaw ? 3 : 2; }  // (ego, horizon) pairs per work item, at most
constexpr int kAxisMax = 64;  // lattice points per axis
// components of an obstacle row staged in shared memory: 2 = centres only (cos / sin of the few exact tests come from
// L2), 4 = the whole row (the exact predicate then never leaves the SM)
#ifndef FISS_OBS_COMP
#define FISS_OBS_COMP 2
#endif
constexpr int kObsComp = FISS_OBS_COMP;
constexpr int kMaxSlots = 4;  // (ego, horizon) pairs per work item
#ifndef FISS_MAT_GROUP
#define FISS_MAT_GROUP 3
#endif
#ifndef FISS_MAT_ILP
#define FISS_MAT_ILP 3
#endif
constexpr int kMatIlp = FISS_MAT_ILP;  // lateral rows whose heading chains advance in lockstep in one lane
constexpr int kMatRows = FISS_MAT_GROUP;  // lateral rows one materialisation task walks with the same frame points

// Shared-memory carve-up (byte offsets, 16-byte aligned), used by the host for the launch size too.
struct GridLayout {
  uint32_t spline, oc, obs, bbox, bbox_key, axes, slot_d, slot_base, slot_off, lon, th, lat, lon_cost, lat_cost, dmax, lon_viol,
      lon_ncart, lon_E, rowdesc, counters, pairs, cflags, masks, listed, near_list, bytes;
};

struct GridArgs {
  const double* ego;   // [B][6]
  const double* axes;  // [4][kAxisMax]: d_end, v_end, T, n (step count as a double)
  int32_t nd, nv, nt;
  int32_t sd, sv, st;  // candidate c = i_d*sd + j_v*sv + k_t*st
  int32_t B, C;
  int32_t d_chunk;     // lateral rows per work item
  int32_t n_chunks;
  int32_t slots;       // (ego, horizon) pairs per work item; > 1 only when the lateral axis is not chunked
  int64_t items;       // work items of the launch: n_big + (B * nt - n_big * slots) when the lateral axis is not chunked, else
                       // B * nt * n_chunks
  int32_t n_big;       // items [0, n_big) carry `slots` pairs each, the items behind them one pair (n_chunks == 1)
  int32_t dynamic;     // 1: the items are handed out through the device counter `work`; 0: with a grid stride
  // Chained launches: the kernel is launched as a programmatic dependent of the previous kernel of the stream (normally the
  // previous step's record kernel, itself a dependent of the previous lattice launch), so that the CTAs of this launch fill
  // the SMs that the previous launch's last items leave idle.  Inputs are read-only; before a CTA's first OUTPUT write the
  // earlier launches must be out of the way:
  //   materialised rows  the previous lattice launch may still be writing the same buffer: wait until its last CTA has
  //                      published `seq_prev` in work[4] (every lattice launch publishes its sequence number when all its CTAs
  //                      are done -- it is running or finished by the time a CTA of this launch exists, so the wait ends);
  //   cost / flags       the previous step's record kernel reads them: griddepcontrol.wait (the whole previous kernel) --
  //                      but not before the CTA's SECOND item: the first item's values go to a shadow block of the engine
  //                      (`shadow`: per CTA ids, costs, flags; two blocks alternate between launches) and are copied
  //                      into the volume behind the wait, by which time that kernel is long over.
  int32_t chained;
  unsigned char* shadow;   // [gridDim.x][shadow_stride * 20 B]
  int32_t shadow_stride;   // entries per CTA (>= candidates of an item)
  uint32_t seq;        // sequence number of this lattice launch among the directly issued ones of its handle (0: part of a graph)
  uint32_t seq_prev;   // ... and the number of the directly issued launch before it (0: none)
  uint32_t* work;      // [0] next item to hand out (minus gridDim.x), [1] CTAs that are done (both zero between launches;
                       // two such pairs alternate between launches: [0..1], [2..3]); [4] sequence number of the last
                       // lattice launch all of whose CTAs are done (work_done points at it)
  uint32_t* work_done;
  uint32_t* work_done_host;  // the same number for the HOST (a mapped, page-locked word): is the GPU still busy with this handle?
  int64_t total;       // B * C
  fiss_params p;
  const double* spline;  // [9][Kp]
  int32_t K, Kp, search_iters;   // search_iters: bisection steps (log2 K, or the bound over the index cells)
  int32_t lut_cells;        // cells of the knot index behind the spline table (0: none)
  int32_t lut_bytes;        // its size, a multiple of 16 (it travels in the spline's bulk copy)
  double lut_inv_h;         // cells / (knots[K-1] - knots[0])
  const double* obs_tab;    // [T_obs][4][Mp]
  const double* obs_const;  // [4][Mp]
  int32_t M, Mp, mp_shift, T_obs, final_time_step;
  int32_t E_stage;     // obstacle rows (centres only) staged in shared memory (0: read them from global / L2)
  int32_t words;       // 32-bit mask words per (row, step) = max(1, Mp/32)
  int32_t n_pad;       // table row length (>= max n + 2, odd: rows then start 4 / 2 banks apart, so that reads of one
                       // column across rows -- the masks, the collision stage -- do not conflict)
  int32_t e_pad;       // mask row length (>= max checked steps)
  double* cost;        // [B*C]
  uint32_t* flags;     // [B*C]
  double* mat;         // [5][B*C][n_stride] or NULL
  int32_t n_stride;
  int64_t mat_pitch;   // elements between two fields of the materialisation = B*C*n_stride
  // x / d == (x * magic(d)) >> 20 for x * d < 2^20 (host-computed: an integer division costs ~20 issue slots per warp)
  uint32_t nv_magic, ns_magic, ng_magic;  // d = nv, n_stride, groups of a full chunk
  uint32_t nt_magic;   // x / nt == umulhi(x, nt_magic) for x < 2^32 / nt, nt > 1
  double kap_limit;    // max_curvature when the optional curvature mask is on, +inf otherwise
  GridLayout lay;      // shared-memory carve-up (host-computed: the offsets are then constant-bank operands)
  int32_t row_len;     // slots * nv * n_pad: elements of one longitudinal table
};

__host__ __device__ inline uint32_t grid_align16(uint32_t v) { return (v + 15u) & ~15u; }

__host__ __device__ inline GridLayout grid_layout(int Kp, int Mp, int E_stage, int nv, int d_chunk, int n_pad, int e_pad,
                                                  int words, int slots, int lut_bytes, bool yaw = false) {
  GridLayout L;
  const uint32_t lon_rows = (uint32_t)slots * nv, lat_rows = (uint32_t)slots * d_chunk;
  uint32_t o = 16;  // two mbarriers
  L.spline = o;     o += 9u * Kp * 8u + (uint32_t)lut_bytes;
  L.oc = o;         o += 4u * Mp * 8u;
  L.obs = o;        o += (uint32_t)E_stage * (uint32_t)kObsComp * Mp * 8u;
  L.bbox = o;       o += (uint32_t)e_pad * 4u * 8u;
  L.bbox_key = o;   o += (uint32_t)e_pad * 4u * 4u;
  L.axes = o;       o += 4u * kAxisMax * 8u;
  L.slot_d = o;     o += 2u * kMaxSlots * 8u * 8u;   // [2][kMaxSlots][8]: ego state (6), T, n -- double-buffered by item parity
  L.slot_base = o;  o += 2u * kMaxSlots * 8u;        // [2][kMaxSlots] id of the slot's candidate (i0, 0, k)
  L.slot_off = o;   o += 2u * kMaxSlots * 4u;        // [2][kMaxSlots] the same relative to slot 0, in output elements
  L.lon = o;        o += grid_align16(5u * lon_rows * n_pad * 8u);
  L.th = o;         o += yaw ? grid_align16(lon_rows * n_pad * 8u) : 0u;  // reference-line heading per (row, step): materialising kernel only
  L.lat = o;        o += grid_align16(lat_rows * n_pad * 8u);
  L.lon_cost = o;   o += grid_align16(lon_rows * 8u);
  L.lat_cost = o;   o += grid_align16(lat_rows * 8u);
  L.dmax = o;       o += 16;
  L.lon_viol = o;   o += grid_align16(lon_rows * 4u);
  L.lon_ncart = o;  o += grid_align16(lon_rows * 4u);
  L.lon_E = o;      o += grid_align16(lon_rows * 4u);
  L.rowdesc = o;    o += yaw ? lon_rows * 16u : 0u;  // per longitudinal row: what a materialisation task needs, one LDS.128
  L.counters = o;   o += 32;
  L.pairs = o;      o += grid_align16(lon_rows * e_pad * 4u);
  L.cflags = o;     o += grid_align16(lat_rows * nv * 4u);
  L.masks = o;      o += grid_align16(lon_rows * e_pad * words * 4u);
  L.listed = o;     o += grid_align16(lon_rows * e_pad * 4u);
  L.near_list = o;  o += grid_align16((uint32_t)e_pad * Mp * 4u);
  L.bytes = o;
  return L;
}

// Floats as unsigned 32-bit keys whose integer order is the numeric order: the per-step bounding boxes are accumulated
// with NATIVE shared-memory atomicMin / atomicMax (64-bit ones are compare-and-swap loops).  The bounds are rounded
// outward to float first, so the box only grows: it stays a conservative filter.
__device__ __forceinline__ uint32_t order_key(float v) {
  const int32_t b = __float_as_int(v);
  return (uint32_t)(b ^ ((b >> 31) | (int32_t)0x80000000));
}
__device__ __forceinline__ float order_value(uint32_t k) {
  const int32_t b = (int32_t)k;
  return __int_as_float(b ^ (((b >> 31) ^ -1) | (int32_t)0x80000000));
}

// Squared distance, one expression for the box test and the per-row test of stage A' (so that rounding is monotone:
// |bx| <= |dx| and |by| <= |dy| imply near2(bx, by) <= near2(dx, dy) bit for bit).
__device__ __forceinline__ double near2(double dx, double dy) { return fma(dx, dx, dy * dy); }

// Candidate position at step m from the row tables (same expression as the generic kernel:
// x = px - d*(ty*r), y = py + d*(tx*r) with the unit tangent stored already multiplied out).
// P2[m] = (px, py) and U2[m] = (ux, uy) are 16-byte pairs: one LDS.128 each, conflict-free for consecutive m.
__device__ __forceinline__ void grid_pos(const double2* __restrict__ P2, const double2* __restrict__ U2,
                                         const double* __restrict__ D, int m, double& x, double& y) {
  const double2 P = P2[m], U = U2[m];
  const double d = D[m];
  x = P.x - d * U.y;
  y = P.y + d * U.x;
}

// Materialisation of R lateral rows at one (longitudinal row, step) lane: position from the frame points, heading,
// 1/ds and curvature (calc_global_paths, frenet_optimal_planner.py:121-134), five stores.  The R rows are independent
// instruction chains in ONE basic block (no vote or branch between them), which is what lets the scheduler overlap
// their fixed FP64 latencies: a warp's issue rate, not its instruction count, bounds this stage.
struct MatOut {
  double* f_x;       // x field of the item's first candidate (NULL: nothing is written)
  bool has_seg, at_seg, has_kap, writes;
};

template <int R>
__device__ __forceinline__ void mat_rows(const GridArgs& a, const MatOut& mo, const double2 Pa, const double2 Ua, const double2 Pb,
                                         const double2 Ub, double th, double sd_v, const double* __restrict__ Dr, int n_pad,
                                         int off, int lat_pitch, uint32_t* cf, int nv) {
  double dx[R], dy[R], yaw[R], inv_ds[R], kap[R];
  bool ok[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const double da = Dr[r * n_pad], db = Dr[r * n_pad + 1];
    const double xa = Pa.x - da * Ua.y, ya = Pa.y + da * Ua.x;
    const double xb = Pb.x - db * Ub.y, yb = Pb.y + db * Ub.x;
    dx[r] = xb - xa;
    dy[r] = yb - ya;
    if (mo.writes) {  // position and speed leave first: their registers are free for the heading chains
      double* o = mo.f_x + (off + r * lat_pitch);
      FISS_ST(o, mo.at_seg ? xa : xb);
      FISS_ST(o + a.mat_pitch, mo.at_seg ? ya : yb);
      FISS_ST(o + 3 * a.mat_pitch, sd_v);
    }
  }
#ifdef FISS_EXP_NOMATH
#pragma unroll
  for (int r = 0; r < R; ++r) { yaw[r] = dx[r]; inv_ds[r] = dy[r]; ok[r] = true; }
#elif defined(FISS_NO_FRAME_HEADING)
  segment_fast_n<R>(dx, dy, yaw, inv_ds, ok);
#else
  // heading relative to the reference line's tangent: the short polynomial when every segment of the task stays within
  // atan(0.3) of it (lanes without a segment -- NaN frame points -- do not vote); else the full-octant path
  bool narrow;
  segment_frame_n<R>(dx, dy, Ua.x, Ua.y, th, yaw, inv_ds, ok, narrow);
  if (__any_sync(kFull, mo.has_seg && !narrow)) segment_fast_n<R>(dx, dy, yaw, inv_ds, ok);
#endif
  // a zero-length / non-finite segment takes the library (its special cases are the reference's)
  bool any_odd = false;
#pragma unroll
  for (int r = 0; r < R; ++r) any_odd |= !ok[r] && mo.has_seg;
  // the next step of the same row is the next lane; the last step of a row never looks at its neighbour
  // (m >= n' - 1), and lane 31 writes nothing.  c = dyaw / ds with ds = hypot(dx, dy) (:128,132; no unwrap; the
  // last element is 0/ds)
  if (__any_sync(kFull, any_odd)) {  // rare, warp-uniform
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const bool odd = !ok[r] && mo.has_seg;
      if (odd) yaw[r] = atan2_library(dy[r], dx[r]);
      const double yaw_next = __shfl_down_sync(kFull, yaw[r], 1);
      kap[r] = (yaw_next - yaw[r]) * inv_ds[r];
      if (odd) kap[r] = div_hypot_library(yaw_next - yaw[r], dx[r], dy[r]);
    }
  } else {
#pragma unroll
    for (int r = 0; r < R; ++r) kap[r] = (__shfl_down_sync(kFull, yaw[r], 1) - yaw[r]) * inv_ds[r];
  }
#pragma unroll
  for (int r = 0; r < R; ++r) kap[r] = mo.has_kap ? kap[r] : CUDART_NAN;
  if (mo.writes) {
#pragma unroll
    for (int r = 0; r < R; ++r) {
      double* o = mo.f_x + (off + r * lat_pitch);
      FISS_ST(o + 2 * a.mat_pitch, yaw[r]);
      FISS_ST(o + 4 * a.mat_pitch, kap[r]);
    }
  }
  if (a.p.check_curvature) {  // the optional curvature mask is on (uniform over the launch; NaN never exceeds)
#pragma unroll
    for (int r = 0; r < R; ++r)
      if (fabs(kap[r]) > a.kap_limit) atomicOr(cf + r * nv, FISS_FLAG_CURVATURE);
  }
}

// kYaw: heading / curvature are needed (materialisation and/or the optional curvature mask).
template <bool kYaw>
__global__ void __launch_bounds__(grid_warps(kYaw) * 32, grid_min_ctas(kYaw)) fiss_grid_kernel(const GridArgs a) {
#ifdef FISS_PHASE_TIMING
  long long phase_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  long long phase_t = clock64();
#endif
#ifdef FISS_TRACE
  int trace_item = 0;
  FISS_STAMP(14);  // kernel entry
  if ((threadIdx.x & 31) == 0 && blockIdx.x < kTraceCtas) {
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    g_fiss_trace[((blockIdx.x * kTraceWarps + (threadIdx.x >> 5)) * kTraceItems) * kTraceStamps + 15] = smid;
  }
#endif
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const GridLayout& L = a.lay;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);  // [0] spline, [1] obstacles
  double* sp = reinterpret_cast<double*>(smem_raw + L.spline);
  double* oc = reinterpret_cast<double*>(smem_raw + L.oc);
  double* obs_s = reinterpret_cast<double*>(smem_raw + L.obs);  // [E_stage][2][Mp]: centres of the checked steps
  // [e_pad][4]: x_min, x_max, y_min, y_max of the frame points of the rows that check a step: accumulated as float keys
  // (rounded outward), decoded to doubles once per item
  double* bbox = reinterpret_cast<double*>(smem_raw + L.bbox);
  uint32_t* bbox_key = reinterpret_cast<uint32_t*>(smem_raw + L.bbox_key);
  double* ax = reinterpret_cast<double*>(smem_raw + L.axes);
  double* slot_d = reinterpret_cast<double*>(smem_raw + L.slot_d);
  long long* slot_base = reinterpret_cast<long long*>(smem_raw + L.slot_base);
  int32_t* slot_off = reinterpret_cast<int32_t*>(smem_raw + L.slot_off);
  double* lon = reinterpret_cast<double*>(smem_raw + L.lon);
  double* lat = reinterpret_cast<double*>(smem_raw + L.lat);  // [slots * d_chunk][n_pad] lateral offsets d(t)
  double* lon_cost = reinterpret_cast<double*>(smem_raw + L.lon_cost);
  double* lat_cost = reinterpret_cast<double*>(smem_raw + L.lat_cost);
  uint32_t* dmax_bits = reinterpret_cast<uint32_t*>(smem_raw + L.dmax);  // max |d| of the item's lateral rows, a float rounded up
  uint32_t* lon_viol = reinterpret_cast<uint32_t*>(smem_raw + L.lon_viol);
  int32_t* lon_ncart = reinterpret_cast<int32_t*>(smem_raw + L.lon_ncart);
  int32_t* lon_E = reinterpret_cast<int32_t*>(smem_raw + L.lon_E);        // checked steps of the row
  // (n', output offset of (slot, j) relative to the item's first candidate, cflags index of (slot, row 0, j), lat index of
  // (slot, row 0, step 0)): written by the row warp, read once per materialisation task (kYaw)
  int4* rowdesc = reinterpret_cast<int4*>(smem_raw + L.rowdesc);
  uint32_t* npairs = reinterpret_cast<uint32_t*>(smem_raw + L.counters);  // length of the work list
  uint32_t* n_near = npairs + 1;     // length of near_list
  uint32_t* rows_done = npairs + 2;  // longitudinal rows of the item that have folded their frame points into the boxes
  uint32_t* row_task = npairs + 3;   // next row task of stage A
  uint32_t* next_item = npairs + 4;  // [2] the CTA's next work item, by item parity (fetched one item ahead)
  uint32_t* n_shadow = npairs + 6;   // [2] entries in the CTA's shadow block, by item parity (chained launches)
  uint32_t* pairs = reinterpret_cast<uint32_t*>(smem_raw + L.pairs);      // (row << 16 | checked step) with any proximity bit
  uint32_t* cflags = reinterpret_cast<uint32_t*>(smem_raw + L.cflags);    // per candidate: collision / curvature bits
  uint32_t* masks = reinterpret_cast<uint32_t*>(smem_raw + L.masks);
  uint32_t* listed = reinterpret_cast<uint32_t*>(smem_raw + L.listed);        // (row, step) already on the work list
  uint32_t* near_list = reinterpret_cast<uint32_t*>(smem_raw + L.near_list);  // (step << 16 | obstacle) that passed the box

  const fiss_params& p = a.p;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int wpc = blockDim.x >> 5;
  const int n_pad = a.n_pad;
  const int Mp = a.Mp;
  const int nv = a.nv;
  const int G = a.slots;
  const int dc = a.d_chunk;
  // "longitudinal row" jj = g * nv + j and "lateral row" ll = g * d_chunk + ii number the rows of all slots of an item
  const int row_len = a.row_len;  // one longitudinal table
  double2* P2 = reinterpret_cast<double2*>(lon);                // [slots * nv][n_pad] frame point (px, py)
  double2* U2 = reinterpret_cast<double2*>(lon + 2 * row_len);  // [slots * nv][n_pad] unit tangent (ux, uy)
  double* SD = lon + 4 * row_len;                               // [slots * nv][n_pad] longitudinal speed
  double* TH = reinterpret_cast<double*>(smem_raw + L.th);      // [slots * nv][n_pad] heading of the reference line (kYaw)

  // ---- stage 0: tables.  The spline is needed first (stage A); the obstacle rows only in stage A'.
  if (threadIdx.x == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    fence_mbar_init();
  }
  __syncthreads();
  const uint32_t const_bytes = 4u * Mp * 8u;
  const uint32_t row_bytes = (uint32_t)kObsComp * Mp * 8u;  // cx, cy (and cos, sin) of one step: the leading components of a table row
  int rows_live = 0;  // staged rows that exist in the table (time < T_obs)
  for (int e = 0; e < a.E_stage; ++e)
    if (p.time_step_now + e * p.check_res < a.T_obs) rows_live = e + 1;
  if (threadIdx.x == 0) {
    const uint32_t spline_bytes = 9u * a.Kp * 8u + (uint32_t)a.lut_bytes;  // the knot index sits behind the table
    mbar_expect_tx(&bar[0], spline_bytes);
    bulk_g2s(sp, a.spline, spline_bytes, &bar[0]);
    mbar_expect_tx(&bar[1], (Mp > 0 ? const_bytes : 0u) + (uint32_t)rows_live * row_bytes);
    if (Mp > 0) bulk_g2s(oc, a.obs_const, const_bytes, &bar[1]);
    for (int e = 0; e < rows_live; ++e)
      bulk_g2s(obs_s + (int64_t)e * kObsComp * Mp, a.obs_tab + (int64_t)(p.time_step_now + e * p.check_res) * 4 * Mp,
               row_bytes, &bar[1]);
  }
  // staged rows past the end of the predictions: nobody has a state there (state_at_time -> None)
  for (int64_t q = (int64_t)rows_live * kObsComp * Mp + threadIdx.x; q < (int64_t)a.E_stage * kObsComp * Mp; q += blockDim.x)
    obs_s[q] = kObsFar;
  for (int q = threadIdx.x; q < 4 * kAxisMax; q += blockDim.x) ax[q] = a.axes[q];

  // centres (cx, cy) of checked step e: staged rows, or the global table
  const double* obs = a.E_stage > 0 ? obs_s : a.obs_tab + (int64_t)p.time_step_now * 4 * Mp;
  const int obs_pitch = a.E_stage > 0 ? kObsComp * Mp : p.check_res * 4 * Mp;
  // (cos, sin) of the obstacles at checked step e (only the exact predicate reads them): staged with the centres
  // (kObsComp == 4) or the global table
  const bool cs_staged = kObsComp == 4 && a.E_stage > 0;
  const double* obs_cs = cs_staged ? obs_s + 2 * Mp : a.obs_tab + (int64_t)p.time_step_now * 4 * Mp + 2 * Mp;
  const int cs_pitch = cs_staged ? 4 * Mp : p.check_res * 4 * Mp;
  const double hle = 0.5 * p.ego_length, hwe = 0.5 * p.ego_width;
  const double re = sqrt(hle * hle + hwe * hwe);

  const uint32_t nv_magic = a.nv_magic;  // c / nv == (c * magic) >> 20 for c < 2^20 / nv
  const int e_pad = a.e_pad;
  const int words = a.words;
  const int res = p.check_res;
  const int t_left = a.final_time_step - p.time_step_now;
  const uint32_t n_items = (uint32_t)a.items;
  const uint32_t n_chunks = (uint32_t)a.n_chunks, nt = (uint32_t)a.nt;
  const uint32_t n_bk = (uint32_t)a.B * nt;  // (ego, horizon) pairs of the launch

  // Per-item state of stage A' / B: clear masks, list marks and counters.  Runs before the first item and inside stage C of
  // every item (which is the only reader of cflags and clears them itself), so that an item costs one barrier less.
  auto reset_item_state = [&]() {
    if (threadIdx.x == 0) {
      *dmax_bits = 0u;
      *npairs = 0u;
      *n_near = 0u;
      *rows_done = 0u;
      *row_task = 0u;
    }
    // empty boxes (+inf, -inf): infinitely far from everything
    for (int q = threadIdx.x; q < 4 * e_pad; q += blockDim.x) bbox_key[q] = order_key((q & 1) ? -CUDART_INF_F : CUDART_INF_F);
    for (int q = threadIdx.x; q < G * nv * e_pad; q += blockDim.x) listed[q] = 0u;
    for (int q = threadIdx.x; q < G * nv * e_pad * words; q += blockDim.x) masks[q] = 0u;
  };
  // The slots of an item: ego state, horizon, step count, output ids -- fetched one item ahead (the global loads are
  // in flight while the previous item finishes), double-buffered by the parity of the CTA's item count.
  // Work items: dealt with a grid stride, or (a.dynamic) handed out through a global counter -- the time of an item follows
  // its collision stage, which varies with the scene around the ego state; the big items then come first and single pairs
  // last, so that the CTAs run out of work together.
  const uint32_t n_big = (uint32_t)a.n_big;
  auto item_decode = [&](uint32_t item, uint32_t& bk0, uint32_t& chunk, int& Gv) {
    if (n_chunks == 1u) {
      const bool big = item < n_big;
      bk0 = big ? item * (uint32_t)G : n_big * (uint32_t)G + (item - n_big);
      Gv = big ? (int)min((uint32_t)G, n_bk - bk0) : 1;
      chunk = 0u;
    } else {
      bk0 = item / n_chunks;
      chunk = item - bk0 * n_chunks;
      Gv = 1;
    }
  };
  auto load_slots = [&](uint32_t item, int par) {
    if (threadIdx.x < (unsigned)(8 * G) && item < n_items) {
      const int g = threadIdx.x >> 3, q = threadIdx.x & 7;
      uint32_t bk0, chunk;
      int Gv_;
      item_decode(item, bk0, chunk, Gv_);
      const uint32_t bk = bk0 + (uint32_t)g;
      if (bk < n_bk) {
        const uint32_t b = nt == 1u ? bk : __umulhi(bk, a.nt_magic), k = bk - b * nt;
        double v;
        if (q < 6) v = a.ego[6 * (int64_t)b + q];
        else v = a.axes[(q - 4) * kAxisMax + k];  // q = 6: T, q = 7: n
        slot_d[(par * kMaxSlots + g) * 8 + q] = v;
        if (q == 0) {
          const uint32_t b0 = nt == 1u ? bk0 : __umulhi(bk0, a.nt_magic), k0 = bk0 - b0 * nt;
          const long long i_off = (long long)chunk * dc * a.sd;
          const long long base = (long long)b * a.C + i_off + (long long)k * a.st;
          const long long base0 = (long long)b0 * a.C + i_off + (long long)k0 * a.st;
          slot_base[par * kMaxSlots + g] = base;
          slot_off[par * kMaxSlots + g] = (int32_t)((base - base0) * a.n_stride);
        }
      }
    }
  };
  reset_item_state();
  for (int q = threadIdx.x; q < G * dc * nv; q += blockDim.x) cflags[q] = 0u;
  load_slots(blockIdx.x, 0);
  mbar_wait(&bar[0], 0);  // the obstacle rows are waited for where stage A' first needs them
  __syncthreads();
  FISS_PHASE(0);

  // Chained launches keep no state in registers across items (the kernel is register-bound): a CTA's first item is item
  // blockIdx.x, and the number of entries in its shadow block sits in shared memory.
  if (threadIdx.x == 0) n_shadow[0] = 0u;
  // shadow block of this CTA: ids, costs, flags; the thread that wrote an entry copies it into the volume
  auto flush_shadow = [&](uint32_t n) {
    int64_t* sh_id = reinterpret_cast<int64_t*>(a.shadow + (size_t)blockIdx.x * a.shadow_stride * 20);
    double* sh_cost = reinterpret_cast<double*>(sh_id + a.shadow_stride);
    uint32_t* sh_flags = reinterpret_cast<uint32_t*>(sh_cost + a.shadow_stride);
    for (uint32_t cidx = threadIdx.x; cidx < n; cidx += blockDim.x) {
      const int64_t id = sh_id[cidx];
      a.cost[id] = sh_cost[cidx];
      a.flags[id] = sh_flags[cidx];
    }
  };
  int par = 0;
  for (uint32_t item = blockIdx.x; item < n_items; item = next_item[par ^ 1], par ^= 1) {
    uint32_t bk0, chunk_u;
    int Gv;  // slots of this item
    item_decode(item, bk0, chunk_u, Gv);
    const int chunk = (int)chunk_u;
    const int i0 = chunk * dc;
    const int rows_i = min(dc, a.nd - i0);
    const int n_lon = Gv * nv;
    const double* sl = slot_d + par * kMaxSlots * 8;

    // the per-item state was reset before the loop / by the previous item's stage C, which also fetched this item's
    // slots; this barrier also means the previous item's readers are done with the tables
    __syncthreads();
    // Warp 0 draws the CTA's next item (the first one is blockIdx.x) and fetches its slots: the other buffer is free from
    // here on, and the global loads complete under stage A.  Everybody reads next_item[] at the end of the item, four
    // barriers from here.
    if (warp == 0) {
      uint32_t nxt = item + gridDim.x;
      if (a.dynamic) {
        if (lane == 0) nxt = gridDim.x + atomicAdd(&a.work[0], 1u);
        nxt = __shfl_sync(kFull, nxt, 0);
      }
      if (lane == 0) next_item[par ^ 1] = nxt;
      // last item of this CTA: a dependent launch (the record kernel, fiss_pick_winners_dev) may start staging its tables
      if (nxt >= n_items) pdl_launch_dependents();
      load_slots(nxt, par ^ 1);  // (threads < 8 * slots: all in warp 0)
    }
    FISS_STAMP(0);
    FISS_PHASE(1);

    // ---- stage A: one warp per row, rows dealt dynamically (the longitudinal rows, ~3x the work of a lateral row,
    // go first)
    for (;;) {
      int task = 0;
      if (lane == 0) task = (int)atomicAdd(row_task, 1u);
      task = __shfl_sync(kFull, task, 0);
      if (task >= n_lon + Gv * rows_i) break;
      if (task < n_lon) {
        // longitudinal quartic, end (v_end, 0)                      polynomial.py:5-19 (SURVEY A.2 closed form)
        const int jj = task;
        const int g = (int)(((uint32_t)jj * nv_magic) >> 20);
        const int j = jj - g * nv;
        const double s0 = sl[8 * g + 0], v0 = sl[8 * g + 1], a0 = sl[8 * g + 2];
        const double T = sl[8 * g + 6];
        const int n = (int)sl[8 * g + 7];
        const double T2 = T * T, T3 = T2 * T;
        const double v_end = ax[kAxisMax + j];
        const double qa2 = 0.5 * a0;
        const double Vq = v_end - v0 - 2.0 * qa2 * T;
        const double Aq = -2.0 * qa2;
        const double qa3 = (3.0 * Vq - Aq * T) / (3.0 * T2);
        const double qa4 = (Aq * T - 2.0 * Vq) / (4.0 * T3);
        double acc = 0.0;
        unsigned viol = 0;
        int first_bad = n;
        const int base = jj * n_pad;
        // Two time steps per lane (m and m + 32) in lockstep: two independent dependency chains through the
        // polynomials, the segment search and the frame -- a row is one warp's serial work, so its latency is the
        // stage's.  The accumulation order (m, then m + 32, then m + 64 ...) is the one-step-per-pass order.
        for (int m0 = lane; m0 < n; m0 += 64) {
          const int m1 = m0 + 32;
          const bool has1 = m1 < n;
          const double tA = m0 * p.tick_t, tB = m1 * p.tick_t;  // np.arange: start + m*step
          const double tA2 = tA * tA, tA3 = tA2 * tA, tA4 = tA3 * tA;
          const double tB2 = tB * tB, tB3 = tB2 * tB, tB4 = tB3 * tB;
          const double sA = s0 + v0 * tA + qa2 * tA2 + qa3 * tA3 + qa4 * tA4;
          const double sB = s0 + v0 * tB + qa2 * tB2 + qa3 * tB3 + qa4 * tB4;
          const double sdA = v0 + 2.0 * qa2 * tA + 3.0 * qa3 * tA2 + 4.0 * qa4 * tA3;
          const double sdB = v0 + 2.0 * qa2 * tB + 3.0 * qa3 * tB2 + 4.0 * qa4 * tB3;
          const double sddA = 2.0 * qa2 + 6.0 * qa3 * tA + 12.0 * qa4 * tA2;
          const double sddB = 2.0 * qa2 + 6.0 * qa3 * tB + 12.0 * qa4 * tB2;
          const double sdddA = 6.0 * qa3 + 24.0 * qa4 * tA;
          const double sdddB = 6.0 * qa3 + 24.0 * qa4 * tB;
          const double dvA = sdA - p.target_speed, dvB = sdB - p.target_speed;
          acc += p.w_speed * (dvA * dvA) + p.w_accel * (sddA * sddA) + p.w_jerk * (sdddA * sdddA);  // cost_function.py:43-46
          if (has1) acc += p.w_speed * (dvB * dvB) + p.w_accel * (sddB * sddB) + p.w_jerk * (sdddB * sdddB);
          if (sdA > p.max_speed || (has1 && sdB > p.max_speed)) viol |= FISS_FLAG_SPEED;               // frenet_optimal_planner.py:152
          if (fabs(sddA) > p.max_accel || (has1 && fabs(sddB) > p.max_accel)) viol |= FISS_FLAG_ACCEL;  // :155
          bool okA, okB;
          double pxA, pyA, txA, tyA, pxB, pyB, txB, tyB;
          spline_frame2(sp, a.K, a.Kp, a.search_iters, sA, sB, okA, okB, pxA, pyA, txA, tyA, pxB, pyB, txB, tyB,
                        reinterpret_cast<const int32_t*>(sp + 9 * a.Kp), a.lut_cells, a.lut_inv_h);
          const double rA = rsqrt(txA * txA + tyA * tyA), rB = rsqrt(txB * txB + tyB * tyB);
          if (okA) {
            P2[base + m0] = make_double2(pxA, pyA);
            U2[base + m0] = make_double2(txA * rA, tyA * rA);
            if (kYaw) TH[base + m0] = atan2_finite(tyA, txA);
          } else {
            first_bad = min(first_bad, m0);
          }
          SD[base + m0] = sdA;
          if (has1) {
            if (okB) {
              P2[base + m1] = make_double2(pxB, pyB);
              U2[base + m1] = make_double2(txB * rB, tyB * rB);
              if (kYaw) TH[base + m1] = atan2_finite(tyB, txB);
            } else {
              first_bad = min(first_bad, m1);
            }
            SD[base + m1] = sdB;
          }
        }
        acc = warp_sum(acc);
        viol = warp_or(viol);
        const int n_cart = warp_min(first_bad);  // n' (:112-113)
        // checked steps i = e*check_res < t_step_max = min(n', final_time_step - now) (:173-176);
        // none when the collision stage is skipped for the row (:257-259) or cannot run (n' < 2)
        const int horizon = min(n_cart, t_left);
        const bool do_coll = a.M > 0 && horizon > 0 && n_cart >= 2 && (p.collide_all || viol == 0);
        const int E_row = do_coll ? (horizon + res - 1) / res : 0;
        if (lane == 0) {
          lon_cost[jj] = acc;
          lon_viol[jj] = viol;
          lon_ncart[jj] = n_cart;
          lon_E[jj] = E_row;
          if (kYaw)
            rowdesc[jj] = make_int4(n_cart, slot_off[par * kMaxSlots + g] + j * (a.sv * a.n_stride), g * dc * nv + j, g * dc * n_pad);
        }
        // frame points from the truncation on (and the pad) are NaN: the materialisation reads them without a test and
        // x, y, heading and curvature of the steps outside the Cartesian part come out NaN by propagation; the speed row is
        // NaN from its own length n on (the output rows are NaN-padded up to n_stride)
        if (kYaw)
          for (int m = n_cart + lane; m < n_pad; m += 32) {
            P2[base + m] = make_double2(CUDART_NAN, CUDART_NAN);
            if (m >= n) SD[base + m] = CUDART_NAN;
          }
        // this row's frame points into the per-step bounding boxes of stage A' (min / max commute, so the boxes do not
        // depend on the order the rows arrive in)
        __syncwarp();
        for (int e = lane; e < E_row; e += 32) {
          const double2 fp = P2[base + e * res];
          atomicMin(&bbox_key[4 * e + 0], order_key(__double2float_rd(fp.x)));
          atomicMax(&bbox_key[4 * e + 1], order_key(__double2float_ru(fp.x)));
          atomicMin(&bbox_key[4 * e + 2], order_key(__double2float_rd(fp.y)));
          atomicMax(&bbox_key[4 * e + 3], order_key(__double2float_ru(fp.y)));
        }
        // the last row to arrive turns the keys into doubles, once, for the box test of stage A'
        __syncwarp();
        unsigned arrived = 0;
        if (lane == 0) {
          __threadfence_block();
          arrived = atomicAdd(rows_done, 1u);
        }
        if (__shfl_sync(kFull, arrived, 0) == (unsigned)n_lon - 1u) {
          __threadfence_block();
          for (int q = lane; q < 4 * e_pad; q += 32) bbox[q] = (double)order_value(bbox_key[q]);
        }
      } else {
        // lateral quintic, end (d_end, 0, 0)                        polynomial.py:45-62
        const int lt = task - n_lon;
        const int g = lt / rows_i;
        const int ii = lt - g * rows_i;
        const int ll = g * dc + ii;
        const double d0 = sl[8 * g + 3], dv0 = sl[8 * g + 4], da0 = sl[8 * g + 5];
        const double T = sl[8 * g + 6];
        const int n = (int)sl[8 * g + 7];
        const double T2 = T * T;
        const double d_end = ax[i0 + ii];
        const double iT = 1.0 / T;
        const double iT2 = iT * iT, iT3 = iT2 * iT;
        const double la2 = 0.5 * da0;
        const double Dl = d_end - d0 - dv0 * T - la2 * T2;
        const double Vl = -dv0 - 2.0 * la2 * T;
        const double Al = -2.0 * la2;
        const double la3 = (10.0 * Dl - 4.0 * Vl * T + 0.5 * Al * T2) * iT3;
        const double la4 = (-15.0 * Dl + 7.0 * Vl * T - Al * T2) * (iT3 * iT);
        const double la5 = (6.0 * Dl - 3.0 * Vl * T + 0.5 * Al * T2) * (iT3 * iT2);
        double acc = 0.0, dmax = 0.0;
        const int base = ll * n_pad;
        for (int m0 = lane; m0 < n; m0 += 64) {  // two steps per lane in lockstep, same accumulation order
          const int m1 = m0 + 32;
          const bool has1 = m1 < n;
          const double tA = m0 * p.tick_t, tB = m1 * p.tick_t;
          const double tA2 = tA * tA, tA3 = tA2 * tA, tA4 = tA3 * tA, tA5 = tA4 * tA;
          const double tB2 = tB * tB, tB3 = tB2 * tB, tB4 = tB3 * tB, tB5 = tB4 * tB;
          const double dA = d0 + dv0 * tA + la2 * tA2 + la3 * tA3 + la4 * tA4 + la5 * tA5;
          const double dB = d0 + dv0 * tB + la2 * tB2 + la3 * tB3 + la4 * tB4 + la5 * tB5;
          const double ddA = 2.0 * la2 + 6.0 * la3 * tA + 12.0 * la4 * tA2 + 20.0 * la5 * tA3;
          const double ddB = 2.0 * la2 + 6.0 * la3 * tB + 12.0 * la4 * tB2 + 20.0 * la5 * tB3;
          const double dddA = 6.0 * la3 + 24.0 * la4 * tA + 60.0 * la5 * tA2;
          const double dddB = 6.0 * la3 + 24.0 * la4 * tB + 60.0 * la5 * tB2;
          acc += p.w_accel * (ddA * ddA) + p.w_jerk * (dddA * dddA) + p.w_offset * (dA * dA);  // cost_function.py:45-47
          dmax = fabs(dA) > dmax ? fabs(dA) : dmax;  // NaN-ignoring: a NaN row has no frame anyway
          lat[base + m0] = dA;
          if (has1) {
            acc += p.w_accel * (ddB * ddB) + p.w_jerk * (dddB * dddB) + p.w_offset * (dB * dB);
            dmax = fabs(dB) > dmax ? fabs(dB) : dmax;
            lat[base + m1] = dB;
          }
        }
        acc = warp_sum(acc);
        // max |d| of the row as a float rounded up (the reach only grows); non-negative floats order as integers, so the
        // warp maximum is one integer reduction
        const uint32_t dmax_row = __reduce_max_sync(kFull, __float_as_uint(__double2float_ru(dmax)));
        if (lane == 0) {
          lat_cost[ll] = acc;
          atomicMax(dmax_bits, dmax_row);
        }
      }
    }
    FISS_STAMP(1);
    __syncthreads();
    FISS_STAMP(2);
    FISS_PHASE(2);

    // ---- stage A': proximity masks.  masks[jj][e] gets a bit per obstacle whose centre is within
    // (max|d| + r_ego + r_obs) of the frame point of longitudinal row jj at checked step e -- a superset of the
    // obstacles any candidate on that row can touch there (max|d| over the lateral rows of the whole item).  The
    // per-step bounding boxes of the frame points were accumulated by the row warps of stage A; here
    //   (1) one lane per (step, obstacle) tests the obstacle against the box (the same squared-distance expression
    //       as the per-row test, so the box can only over-accept); the few survivors go to a compact list;
    //   (2) one lane per (survivor, row) runs the per-row test, sets the mask bit and -- the first one to touch a
    //       (row, step) -- appends it to the work list of stage B.
    if (a.M > 0) {
      mbar_wait(&bar[1], 0);  // obstacle rows landed (phase 0 completes once: later items pass straight through)
      const double dmax = (double)__uint_as_float(*dmax_bits);
      const double reach0 = (dmax + re) * (1.0 + 1.0e-9) + 1.0e-9;
      // lanes = obstacles (their reach is per-lane state), warps stride over the checked steps; with fewer than 32
      // obstacle slots a warp packs 32 / Mp steps per pass
      const int jl = lane & (min(Mp, 32) - 1);
      const int e_sub = lane >> a.mp_shift;
      const int steps_per_pass = 32 >> a.mp_shift;
      for (int jb = 0; jb < Mp; jb += 32) {
        const int jo = jb + jl;
        const double reach = reach0 + oc[2 * Mp + jo] * (1.0 + 1.0e-9);
        const double reach2 = reach * reach;
        for (int e = warp * steps_per_pass + e_sub; e < e_pad; e += wpc * steps_per_pass) {
          const bool in_tab = a.E_stage > 0 ? e < a.E_stage : p.time_step_now + e * res < a.T_obs;
          if (jo < a.M && in_tab) {
            const double* slot = obs + e * obs_pitch + jo;
            const double ox = slot[0], oy = slot[Mp];
            const double2 box_x = *reinterpret_cast<const double2*>(bbox + 4 * e);
            const double2 box_y = *reinterpret_cast<const double2*>(bbox + 4 * e + 2);
            // distance to the box per axis: max(lo - o, o - hi, 0), each value one of the differences (or 0)
            const double xl = box_x.x - ox, xh = ox - box_x.y, yl = box_y.x - oy, yh = oy - box_y.y;
            double bx = xh > xl ? xh : xl, by = yh > yl ? yh : yl;
            bx = bx > 0.0 ? bx : 0.0;
            by = by > 0.0 ? by : 0.0;
            if (near2(bx, by) <= reach2) near_list[atomicAdd(n_near, 1u)] = ((uint32_t)e << 16) | (uint32_t)jo;
          }
        }
      }
      FISS_STAMP(3);
      __syncthreads();
      FISS_STAMP(4);
      FISS_PHASE(4);
      const uint32_t n_cand_pairs = *n_near * (uint32_t)n_lon;
      for (uint32_t q = threadIdx.x; q < n_cand_pairs; q += blockDim.x) {
        const uint32_t c = q / (uint32_t)n_lon;
        const int jj = (int)(q - c * (uint32_t)n_lon);
        const uint32_t ej = near_list[c];
        const int e = (int)(ej >> 16), jo = (int)(ej & 0xffffu);
        if (e < lon_E[jj]) {
          const double* slot = obs + e * obs_pitch + jo;
          const double2 fp = P2[jj * n_pad + e * res];
          const double reach = reach0 + oc[2 * Mp + jo] * (1.0 + 1.0e-9);
          if (near2(slot[0] - fp.x, slot[Mp] - fp.y) <= reach * reach) {
            atomicOr(&masks[(jj * e_pad + e) * words + (jo >> 5)], 1u << (jo & 31));
            if (atomicExch(&listed[jj * e_pad + e], 1u) == 0u)
              pairs[atomicAdd(npairs, 1u)] = ((uint32_t)jj << 16) | (uint32_t)e;
          }
        }
      }
    }
    FISS_STAMP(5);
    __syncthreads();
    FISS_STAMP(6);
    FISS_PHASE(5);

    // ---- stage B, collision (has_collision, frenet_optimal_planner.py:168-195): one lane per
    // (lateral row of the pair's slot, listed (row, step) pair); exact predicate on the listed obstacles only
    {
      const uint32_t n_pairs = *npairs;
      const uint32_t n_work = (uint32_t)rows_i * n_pairs;
      // consecutive lanes = consecutive lateral rows of ONE pair: the four 16-byte frame-point reads of a warp are then
      // broadcasts of ~4 addresses (with the pairs fastest they were 32 scattered ones, ~4-way bank conflicts each)
      const uint32_t rows_magic = 0xffffffffu / (uint32_t)rows_i + 1u;  // w / rows_i == umulhi(w, magic), w < 2^32 / rows_i
      for (uint32_t wk = threadIdx.x; wk < n_work; wk += blockDim.x) {
        const uint32_t q = rows_i == 1 ? wk : __umulhi(wk, rows_magic);
        const int ii = (int)(wk - q * (uint32_t)rows_i);
        const uint32_t pr = pairs[q];
        const int jj = (int)(pr >> 16), e = (int)(pr & 0xffffu);
        const int g = (int)(((uint32_t)jj * nv_magic) >> 20);
        const int ll = g * dc + ii;
        const double* Dr = lat + ll * n_pad;
        const double2* pP = P2 + jj * n_pad;
        const double2* pU = U2 + jj * n_pad;
        // ego pose at checked step i = e*check_res: centre (x_i, y_i), heading of segment min(i, n'-2)
        const int i = e * res;
        const int seg = min(i, lon_ncart[jj] - 2);
        double xa, ya, xb, yb;
        grid_pos(pP, pU, Dr, seg, xa, ya);
        grid_pos(pP, pU, Dr, seg + 1, xb, yb);
        const double dxs = xb - xa, dys = yb - ya;
        const double h2 = dxs * dxs + dys * dys;
        double c, s;
        if (h2 > 0.0 && h2 < 1.0e300) {
          const double r = rsqrt(h2);  // cos/sin of atan2(dy, dx) without the round trip
          c = dxs * r;
          s = dys * r;
        } else {
          sincos(atan2(dys, dxs), &s, &c);
        }
        const double ex = i == seg ? xa : xb, ey = i == seg ? ya : yb;
        const double* orow = obs + e * obs_pitch;
        const double* crow = obs_cs + (int64_t)e * cs_pitch;
        const uint32_t* mw = masks + (jj * e_pad + e) * words;
        bool h = false;
        for (int w = 0; w < words && !h; ++w) {
          uint32_t bits = mw[w];
          while (bits && !h) {
            const int jo = w * 32 + __ffs(bits) - 1;
            bits &= bits - 1;
            const double dx = orow[jo] - ex, dy = orow[Mp + jo] - ey;
            const double thr = re + oc[2 * Mp + jo];
            if (dx * dx + dy * dy <= thr * thr)
              h = rect_sat(dx, dy, c, s, hle, hwe, crow[jo], crow[Mp + jo], oc[jo], oc[Mp + jo]);
          }
        }
        if (h) atomicOr(&cflags[ll * nv + (jj - g * nv)], FISS_FLAG_COLLISION);
      }
    }

#ifdef FISS_PHASE_TIMING
    __syncthreads();
    FISS_PHASE(6);
#endif
    FISS_STAMP(7);
    // ---- stage B, materialisation (calc_global_paths, frenet_optimal_planner.py:121-134).
    // Lanes = flattened (longitudinal row jj, time step m) elements f = jj*n_stride + m of the frame tables, cut into
    // blocks of 31 outputs: the 32nd lane of a block only supplies the heading of the next step (kappa_m needs
    // yaw_{m+1}), so blocks are independent.  A task = (block, group of kMatRows lateral rows of the lane's slot): the
    // lane loads its two frame points (4 x LDS.128) and advances the heading chains of the group's rows in lockstep
    // (mat_rows); five coalesced stores per row.
    if (kYaw) {
      // first output write of this CTA (materialised rows): the previous launch -- which may still be writing the same
      // buffers, or reading this launch's cost / flags volume -- has to be over
#ifndef FISS_TEST_NO_GATE  // (negative control of tests/test_gpu_chained.py: built without the guard, the tests must fail)
      if (a.chained && item == blockIdx.x) {  // (the CTA's first item; later ones are behind this wait anyway)
        if (lane == 0) {
          uint32_t done;
          do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(done) : "l"(a.work_done) : "memory");
          } while ((int32_t)(done - a.seq_prev) < 0);
        }
        __syncwarp();
      }
#endif
      const int ns = a.n_stride;
      const uint32_t ns_magic = a.ns_magic;  // f / ns for f < 2^20 / ns
      const int n_blocks = (n_lon * ns + 30) / 31;
      const int n_groups = (rows_i + kMatRows - 1) / kMatRows;
      const uint32_t ng_magic = rows_i == dc ? a.ng_magic : (1u << 20) / (uint32_t)n_groups + 1u;
      MatOut mo;
      // the x field of slot 0's candidate (i0, j = 0, k): uniform per item; the lanes add 32-bit element offsets
      // (the host checks that slots * C * n_stride < 2^31 elements)
      mo.f_x = a.mat ? a.mat + slot_base[par * kMaxSlots] * ns : nullptr;
      const int lat_pitch = a.sd * ns;
      // tasks t = blk * n_groups + grp.  A task recomputes its lane set-up from t (~35 issue slots against ~400 for its
      // rows): state carried from task to task was spilled to local memory, and the loads of the spilled loop state
      // were the longest stalls of the kernel.
      // (Dealing the tasks through a shared counter instead of this static stride measured 4 % slower.)  The stride
      // starts from the last warp: the low warps carry the extra pass of the collision stage above.
      const int n_tasks = n_groups * n_blocks;
      for (int t = wpc - 1 - warp; t < n_tasks; t += wpc) {
        const int blk = (int)(((uint32_t)t * ng_magic) >> 20);
        const int grp = t - blk * n_groups;
        const int f = blk * 31 + lane;
        const int jj = min((int)(((uint32_t)f * ns_magic) >> 20), n_lon - 1);
        const int m = f - jj * ns;  // >= ns for the lanes past the end of the table
        const int4 rd = rowdesc[jj];  // (n', output offset, cflags index, lat index) of the row
        const int n_cart = rd.x;
        const bool in_cart = m < n_cart;  // (n' <= n <= ns)
        // yaw_m = atan2 of segment min(m, n'-2): the last point repeats the previous heading (:127-130).  Steps
        // outside the Cartesian part read NaN frame points (stage A), and so does the neighbour of a lone point
        // (n' == 1: yaw / ds / c stay empty, :121).
        const int seg = in_cart ? max(min(m, n_cart - 2), 0) : min(m, n_pad - 2);
        const int sidx = jj * n_pad + seg;
        const double2* fp = P2 + sidx;
        mo.has_seg = in_cart && n_cart >= 2;
        mo.at_seg = m == seg;
        mo.has_kap = m < n_cart - 1 && lane < 31;  // lane 31 only supplies the next heading
        mo.writes = mo.f_x != nullptr && lane < 31 && m < ns;
#ifdef FISS_EXP_NOSTORE
        mo.writes = mo.writes && a.B < 0;
#endif
        const double sd_v = SD[sidx - seg + min(m, n_pad - 1)];  // NaN from the row's length on (stage A)
        const double2 Pa = fp[0], Pb = fp[1];
        const double2 Ua = fp[row_len], Ub = fp[row_len + 1];  // U2 = P2 + row_len
        const double th = TH[sidx];
        const int i_first = grp * kMatRows;
        const double* Dr = lat + (rd.w + i_first * n_pad + seg);  // first lateral row of the group in the lane's slot
        const int off = rd.y + i_first * lat_pitch + m;
        uint32_t* cf = cflags + (rd.z + i_first * nv);
        const int rows_here = min(kMatRows, rows_i - i_first);  // warp-uniform
        int r = 0;
        for (; r + kMatIlp <= rows_here; r += kMatIlp)
          mat_rows<kMatIlp>(a, mo, Pa, Ua, Pb, Ub, th, sd_v, Dr + r * n_pad, n_pad, off + r * lat_pitch, lat_pitch, cf + r * nv, nv);
        for (; r < rows_here; ++r)
          mat_rows<1>(a, mo, Pa, Ua, Pb, Ub, th, sd_v, Dr + r * n_pad, n_pad, off + r * lat_pitch, lat_pitch, cf + r * nv, nv);
      }
    }
    FISS_STAMP(8);
    __syncthreads();
    FISS_STAMP(9);
    FISS_PHASE(7);

    // chained launches: the first item's cost / flags go to the CTA's shadow block; from the second item on the volume is
    // written directly, behind the wait for the previous kernel of the stream
#ifdef FISS_TEST_NO_SHADOW  // (negative control: the first item's cost / flags straight into the volume, no wait at all)
    const bool to_shadow = false;
#else
    const bool to_shadow = a.chained && item == blockIdx.x;
#endif
    if (a.chained && !to_shadow) {
#if !defined(FISS_TEST_NO_WAIT) && !defined(FISS_TEST_NO_SHADOW)  // (negative control, as above)
      pdl_wait_producer();  // (returns at once from the CTA's third item on)
#endif
      flush_shadow(n_shadow[par]);
    }
    // ---- stage C: one lane per candidate -- cost (cost_function.py:41-50) and the flags word; then the reset of the
    // per-item state and the fetch of the next item's slots
    {
      const int n_cand = Gv * rows_i * nv;
      int64_t* sh_id = reinterpret_cast<int64_t*>(a.shadow + (size_t)blockIdx.x * a.shadow_stride * 20);
      double* sh_cost = reinterpret_cast<double*>(sh_id + a.shadow_stride);
      uint32_t* sh_flags = reinterpret_cast<uint32_t*>(sh_cost + a.shadow_stride);
      for (int cidx = threadIdx.x; cidx < n_cand; cidx += blockDim.x) {
        const int lt = (int)(((uint32_t)cidx * nv_magic) >> 20);  // lateral row among the item's Gv * rows_i
        const int j = cidx - lt * nv;
        const int g = lt / rows_i;
        const int ii = lt - g * rows_i;
        const int ll = g * dc + ii, jj = g * nv + j;
        const int n = (int)sl[8 * g + 7];
        const double inv_n = 1.0 / (double)n;
        const double cost_time = p.cost_time_offset - (n - 1) * p.tick_t;  // (10 - t_last), cost_function.py:42
        const int n_cart = lon_ncart[jj];
        const unsigned viol = lon_viol[jj];
        unsigned extra = cflags[ll * nv + j];
        cflags[ll * nv + j] = 0u;
        // n' == 1 with obstacles: traj.yaw[0] raises inside the try => "collision" (:178-182)
        if (a.M > 0 && min(n_cart, t_left) > 0 && n_cart < 2 && (p.collide_all || viol == 0)) extra |= FISS_FLAG_COLLISION;
        const int64_t out_id = slot_base[par * kMaxSlots + g] + ii * a.sd + j * a.sv;
        const double cost_v = (cost_time + (lon_cost[jj] + lat_cost[ll])) * inv_n;
        const uint32_t flags_v = viol | extra | ((uint32_t)n_cart << FISS_FLAG_NCART_SHIFT);
        if (to_shadow) {
          sh_id[cidx] = out_id;
          sh_cost[cidx] = cost_v;
          sh_flags[cidx] = flags_v;
        } else {
          a.cost[out_id] = cost_v;
          a.flags[out_id] = flags_v;
        }
      }
      if (threadIdx.x == 0) n_shadow[par ^ 1] = to_shadow ? (uint32_t)n_cand : 0u;  // (read by the next item's stage C)
      reset_item_state();
    }
    FISS_STAMP(10);
#ifdef FISS_TRACE
    ++trace_item;
#endif
  }
  if (a.chained) {  // (a CTA with a single item still has its shadow block to copy)
    __syncthreads();  // n_shadow[] of the last item is thread 0's
    pdl_wait_producer();
    flush_shadow(n_shadow[par]);
  }
  // Every CTA has drawn its last (out-of-range) item by now and its output writes are issued: the last one to leave zeroes
  // the counters for the launch after the next and publishes this launch's sequence number (chained launches, above).
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(&a.work[1], 1u) == gridDim.x - 1u) {
      a.work[0] = 0u;
      a.work[1] = 0u;
      __threadfence();
      if (a.seq) {
        asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(a.work_done), "r"(a.seq) : "memory");
        *reinterpret_cast<volatile uint32_t*>(a.work_done_host) = a.seq;
      }
    }
  }
#ifdef FISS_PHASE_TIMING
  if (threadIdx.x == 0)
    for (int k = 0; k < 8; ++k) atomicAdd((unsigned long long*)&g_fiss_phase[k], (unsigned long long)phase_acc[k]);
#endif
}

}  // namespace fiss
