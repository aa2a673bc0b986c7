// fiss_kernels.cuh -- sm_100a kernels of the Frenet lattice engine (device side).
//
// One warp per candidate trajectory.  A persistent CTA stages the reference-line spline table and
// the slice of the obstacle table it needs into shared memory with 1-D bulk TMA
// (cp.async.bulk + mbarrier), then its warps stride over candidates:
//
//   P0  closed-form quintic (lateral) / quartic (longitudinal) coefficients      polynomial.py:5-19,45-62
//   P1  lanes stride over time steps: 8 polynomial evaluations, cost terms,       polynomial.py:21-41,64-84
//       speed/accel masks, spline segment search + Frenet->Cartesian              cost_function.py:41-50
//                                                                                 frenet_optimal_planner.py:106-119,140-160
//       -> warp-shuffle reductions: cost, mask bits, truncation length n'
//   P2  heading of each checked step from finite differences (and yaw/ds/kappa    frenet_optimal_planner.py:121-134
//       when materialising)
//   P3  collision: (checked step, obstacle) pairs spread over the 32 lanes;       frenet_optimal_planner.py:168-195
//       exact circumscribed-circle reject, then closed-set SAT of the two
//       oriented rectangles; early exit on the first hit (__any_sync)
//
// All arithmetic is FP64 (kappa = dyaw/ds is ill-conditioned: SURVEY A.9).  No tensor cores: the
// path is a scan with ~10^4 flop per candidate, not a contraction.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>
#include <math_constants.h>

#include "fiss_abi.h"

namespace fiss {

constexpr unsigned kFull = 0xffffffffu;
constexpr int kWarpsPerCta = 8;
constexpr int kThreads = kWarpsPerCta * 32;
constexpr double kObsFar = 1.0e150;  // centre of a "no state at this step" obstacle slot

struct EvalArgs {
  // candidates
  const double* ego;   // [B][6]
  const double* end;   // [C][4] (d_end, v_end, T, n)
  const int32_t* sel;  // optional candidate selection (full-record mode)
  int32_t B, C;
  int64_t total;       // number of candidates this launch evaluates
  int32_t per_problem; // full-record mode: record r -> ego[r], end[sel[r]]
  fiss_params p;
  // spline table in global memory: [9][Kp]
  const double* spline;
  int32_t K, Kp, search_iters;
  // obstacle tables in global memory
  const double* obs_tab;    // [T_obs][4][Mp]: rows cx, cy, cos, sin of every obstacle at one time step
  const double* obs_const;  // [4][Mp]: rows half_l, half_w, circumscribed radius, 0
  int32_t M, Mp, mp_shift, T_obs, final_time_step;
  int32_t E_max;            // checked steps staged per CTA
  int32_t obs_in_smem;
  // outputs
  double* cost;
  uint32_t* flags;
  double* mat;      // [5][total][n_stride] or NULL
  double* records;  // [total][16][n_stride] or NULL
  int32_t n_stride;
  int32_t n_pad;    // per-warp scratch row length (>= max n, multiple of 2)
  int32_t e_cap;    // per-warp capacity of the compact checked-step arrays (>= ceil(max n / check_res))
  // full-record mode with the pick fused in (one launch instead of pick + records): the warp of record r first
  // takes the argmin over problem r's C candidates of pick_cost / pick_flags, publishes it, then writes its record
  const double* pick_cost;     // [total][C] or NULL (then `sel` names the candidates)
  const uint32_t* pick_flags;  // [total][C]
  int32_t* pick_idx;           // [total] winner (-1: none)
  double* pick_best;           // [total] its cost (+inf: none)
  int32_t* pick_meta;          // [total][2] (n, n') of the winner, or NULL
  // launched with programmatic stream serialisation behind the kernel that writes pick_cost / pick_flags: the CTAs
  // stage their tables while that kernel drains and wait for its results only then
  int32_t after_producer;
};

// Programmatic dependent launch (PTX griddepcontrol): a producer signals that its dependents may be scheduled; a
// dependent waits until the producer grid has completed and its writes are visible.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait_producer() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------
// PTX wrappers: mbarrier + 1-D bulk TMA (SASS: SYNCS.* / UBLKCP)
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}
__device__ __forceinline__ int warp_min(int v) { return __reduce_min_sync(kFull, v); }
__device__ __forceinline__ unsigned warp_or(unsigned v) { return __reduce_or_sync(kFull, v); }

// argmin with the reference's tie rule: among feasible candidates the minimum cost, and among equal minima the
// LARGEST index (`min_cost >= cost` scan, frenet_optimal_planner.py:263-268).  NaN costs never win (SURVEY 5:
// NaN/Inf candidates are masked infeasible).  The selection is associative and commutative, so any reduction
// order gives the reference's winner.
__device__ __forceinline__ bool better(double c, int i, double bc, int bi) {
  return (c < bc) || (c == bc && i > bi);
}
__device__ __forceinline__ void pick_scan(const double* __restrict__ pc, const uint32_t* __restrict__ pf, int C,
                                          int first, int stride, double& bc, int& bi) {
  bc = CUDART_INF;
  bi = -1;
  // eight independent loads in flight per thread: the scan is a chain of L2 latencies otherwise (it is on the
  // latency path of every plan cycle)
  constexpr int kU = 8;
  for (int i0 = first; i0 < C; i0 += kU * stride) {
    double c[kU];
    uint32_t f[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int i = i0 + u * stride;
      c[u] = i < C ? pc[i] : CUDART_NAN;
      f[u] = i < C ? pf[i] : FISS_FLAG_INFEASIBLE_MASK;
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int i = i0 + u * stride;
      const bool feasible = (f[u] & FISS_FLAG_INFEASIBLE_MASK) == 0 && c[u] <= CUDART_INF;  // NaN -> false
      if (feasible && (bi < 0 || better(c[u], i, bc, bi))) {
        bc = c[u];
        bi = i;
      }
    }
  }
}
__device__ __forceinline__ void pick_warp_reduce(double& bc, int& bi) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double oc = __shfl_xor_sync(kFull, bc, o);
    const int oi = __shfl_xor_sync(kFull, bi, o);
    if (oi >= 0 && (bi < 0 || better(oc, oi, bc, bi))) {
      bc = oc;
      bi = oi;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Reference line.  `sp` points at [9][Kp]: knots, ax, bx, cx, dx, ay, by, cy, dy.
// Returns false when s is outside [knots[0], knots[K-1]) (cubic_spline.py:57-60; s == last knot is an
// IndexError in the reference, treated as outside: SURVEY A.4).  NaN compares false -> outside.
__device__ __forceinline__ void spline_eval(const double* __restrict__ sp, int Kp, int lo, double s, double& px,
                                            double& py, double& tx, double& ty) {
  const double dx = s - sp[lo];
  const double dx2 = dx * dx;
  const double dx3 = dx2 * dx;
  const double* cx = sp + Kp + lo;
  const double* cy = sp + 5 * Kp + lo;
  const double bx = cx[Kp], ccx = cx[2 * Kp], ddx = cx[3 * Kp];
  const double by = cy[Kp], ccy = cy[2 * Kp], ddy = cy[3 * Kp];
  px = cx[0] + bx * dx + ccx * dx2 + ddx * dx3;                // cubic_spline.py:64
  py = cy[0] + by * dx + ccy * dx2 + ddy * dx3;
  tx = bx + 2.0 * ccx * dx + 3.0 * ddx * dx2;                  // cubic_spline.py:86
  ty = by + 2.0 * ccy * dx + 3.0 * ddy * dx2;
}

__device__ __forceinline__ bool spline_frame(const double* __restrict__ sp, int K, int Kp, int iters, double s,
                                             double& px, double& py, double& tx, double& ty) {
  if (!(s >= sp[0] && s < sp[K - 1])) return false;
  int lo = 0, hi = K - 1;  // knots[lo] <= s < knots[hi]  == bisect_right(knots, s) - 1 (cubic_spline.py:112-116)
  for (int it = 0; it < iters; ++it) {
    int mid = (lo + hi) >> 1;
    bool le = sp[mid] <= s;
    lo = (le && hi - lo > 1) ? mid : lo;
    hi = (!le && hi - lo > 1) ? mid : hi;
  }
  spline_eval(sp, Kp, lo, s, px, py, tx, ty);
  return true;
}

// Two independent abscissae searched and evaluated in lockstep (two dependency chains per lane instead of one; the
// searches are the latency of the row stage).  Same arithmetic as spline_frame; out-of-range / NaN abscissae still walk
// the (in-bounds) search and are reported through okA / okB.
//
// `lut` (optional, lut_cells > 0): a uniform grid over [knots[0], knots[K-1]), lut[c] = the largest i with knots[i] <= start
// of cell c (built by fiss_set_spline).  The cell of s, computed in floating point, is off by at most one, so
// [lut[c-1], lut[c+2] + 1] brackets s like [0, K-1] does, and `iters` (the host's bound over all cells: 1 for evenly spaced
// knots instead of log2 K) iterations of the same bisection end on the same segment.
__device__ __forceinline__ void lut_window(const int32_t* __restrict__ lut, int lut_cells, double lut_inv_h, double k0, int K,
                                           double s, int& lo, int& hi) {
  int c = (int)((s - k0) * lut_inv_h);  // saturating; NaN -> 0 (such an abscissa is reported through ok)
  c = max(0, min(c, lut_cells - 1));
  lo = lut[max(c - 1, 0)];
  hi = min(lut[min(c + 2, lut_cells)] + 1, K - 1);
}

__device__ __forceinline__ void spline_frame2(const double* __restrict__ sp, int K, int Kp, int iters, double sA, double sB,
                                              bool& okA, bool& okB, double& pxA, double& pyA, double& txA, double& tyA,
                                              double& pxB, double& pyB, double& txB, double& tyB,
                                              const int32_t* __restrict__ lut = nullptr, int lut_cells = 0,
                                              double lut_inv_h = 0.0) {
  const double k0 = sp[0], k1 = sp[K - 1];
  okA = sA >= k0 && sA < k1;
  okB = sB >= k0 && sB < k1;
  int loA = 0, hiA = K - 1, loB = 0, hiB = K - 1;
  if (lut_cells > 0) {
    lut_window(lut, lut_cells, lut_inv_h, k0, K, sA, loA, hiA);
    lut_window(lut, lut_cells, lut_inv_h, k0, K, sB, loB, hiB);
  }
  for (int it = 0; it < iters; ++it) {
    const int midA = (loA + hiA) >> 1, midB = (loB + hiB) >> 1;
    const bool leA = sp[midA] <= sA, leB = sp[midB] <= sB;
    const bool openA = hiA - loA > 1, openB = hiB - loB > 1;
    loA = (leA && openA) ? midA : loA;
    hiA = (!leA && openA) ? midA : hiA;
    loB = (leB && openB) ? midB : loB;
    hiB = (!leB && openB) ? midB : hiB;
  }
  spline_eval(sp, Kp, loA, sA, pxA, pyA, txA, tyA);
  spline_eval(sp, Kp, loB, sB, pxB, pyB, txB, tyB);
}

// Closed-set SAT of two oriented rectangles in centre/axis form; true = they intersect
// (touching counts: GEOS `intersects`, frenet_optimal_planner.py:191).
__device__ __forceinline__ bool rect_sat(double dx, double dy, double ce, double se, double hle, double hwe,
                                         double co, double so, double hlo, double hwo) {
  const double cd = fabs(ce * co + se * so);   // |cos(delta)|
  const double sd = fabs(se * co - ce * so);   // |sin(delta)|
  if (fabs(dx * ce + dy * se) > hle + hlo * cd + hwo * sd) return false;
  if (fabs(dy * ce - dx * se) > hwe + hlo * sd + hwo * cd) return false;
  if (fabs(dx * co + dy * so) > hlo + hle * cd + hwe * sd) return false;
  if (fabs(dy * co - dx * so) > hwo + hle * sd + hwe * cd) return false;
  return true;
}

// ------------------------------------------------------------------------------------------------
// One candidate, one warp.  kMat: also write the (x, y, yaw, v, kappa) rows; kRec: write the full
// 16-row record.  Returns (cost, flags) in every lane.
template <bool kMat, bool kRec>
__device__ __forceinline__ void eval_candidate(const EvalArgs& a, const double* __restrict__ sp,
                                               const double* __restrict__ obs, const double* __restrict__ oc,
                                               double* __restrict__ scratch, int lane, const double* __restrict__ ego,
                                               const double* __restrict__ end, int64_t out_id, double& cost_out,
                                               uint32_t& flags_out) {
  constexpr bool kYaw = kMat || kRec;
  const fiss_params& p = a.p;
  const int n_pad = a.n_pad;
  double* xs = scratch;
  double* ys = scratch + n_pad;
  double* yw = scratch + 2 * n_pad;  // yaw (kYaw) ; kappa is recomputed from it
  double* ex = scratch + 3 * n_pad;  // compact per-checked-step ego pose: x, y, cos, sin
  double* ey = ex + a.e_cap;
  double* ec = ey + a.e_cap;
  double* es = ec + a.e_cap;

  const double s0 = ego[0], v0 = ego[1], a0 = ego[2], d0 = ego[3], dv0 = ego[4], da0 = ego[5];
  const double d_end = end[0], v_end = end[1], T = end[2];
  const int n = static_cast<int>(end[3]);

  // ---- P0: coefficients (SURVEY A.2 closed forms of the 2x2 / 3x3 solves)
  const double T2 = T * T, T3 = T2 * T;
  const double iT = 1.0 / T;
  const double iT2 = iT * iT, iT3 = iT2 * iT;
  // longitudinal quartic: end (v_end, 0)
  const double qa2 = 0.5 * a0;
  const double Vq = v_end - v0 - 2.0 * qa2 * T;
  const double Aq = -2.0 * qa2;
  const double qa3 = (3.0 * Vq - Aq * T) / (3.0 * T2);
  const double qa4 = (Aq * T - 2.0 * Vq) / (4.0 * T3);
  // lateral quintic: end (d_end, 0, 0)
  const double la2 = 0.5 * da0;
  const double Dl = d_end - d0 - dv0 * T - la2 * T2;
  const double Vl = -dv0 - 2.0 * la2 * T;
  const double Al = -2.0 * la2;
  const double la3 = (10.0 * Dl - 4.0 * Vl * T + 0.5 * Al * T2) * iT3;
  const double la4 = (-15.0 * Dl + 7.0 * Vl * T - Al * T2) * (iT3 * iT);
  const double la5 = (6.0 * Dl - 3.0 * Vl * T + 0.5 * Al * T2) * (iT3 * iT2);

  // ---- P1: per-step evaluation
  double acc = 0.0;
  unsigned viol = 0;
  int first_bad = n;
  double* rec = kRec ? a.records + out_id * (int64_t)(FISS_REC_ROWS * a.n_stride) : nullptr;
  double* mat_v = (kMat && a.mat) ? a.mat + ((int64_t)FISS_MAT_V * a.total + out_id) * a.n_stride : nullptr;
  const int n_loop = (kRec || (kMat && mat_v)) ? a.n_stride : n;
  for (int m = lane; m < n_loop; m += 32) {
    if (m < n) {
      const double t = m * p.tick_t;  // np.arange: start + m*step
      const double t2 = t * t, t3 = t2 * t, t4 = t3 * t, t5 = t4 * t;
      const double s = s0 + v0 * t + qa2 * t2 + qa3 * t3 + qa4 * t4;
      const double s_d = v0 + 2.0 * qa2 * t + 3.0 * qa3 * t2 + 4.0 * qa4 * t3;
      const double s_dd = 2.0 * qa2 + 6.0 * qa3 * t + 12.0 * qa4 * t2;
      const double s_ddd = 6.0 * qa3 + 24.0 * qa4 * t;
      const double d = d0 + dv0 * t + la2 * t2 + la3 * t3 + la4 * t4 + la5 * t5;
      const double d_d = dv0 + 2.0 * la2 * t + 3.0 * la3 * t2 + 4.0 * la4 * t3 + 5.0 * la5 * t4;
      const double d_dd = 2.0 * la2 + 6.0 * la3 * t + 12.0 * la4 * t2 + 20.0 * la5 * t3;
      const double d_ddd = 6.0 * la3 + 24.0 * la4 * t + 60.0 * la5 * t2;
      // cost_total terms (cost_function.py:43-47)
      const double dv = s_d - p.target_speed;
      acc += p.w_speed * (dv * dv) + p.w_accel * (s_dd * s_dd + d_dd * d_dd) +
             p.w_jerk * (s_ddd * s_ddd + d_ddd * d_ddd) + p.w_offset * (d * d);
      // check_constraints (frenet_optimal_planner.py:152,155)
      if (s_d > p.max_speed) viol |= FISS_FLAG_SPEED;
      if (fabs(s_dd) > p.max_accel) viol |= FISS_FLAG_ACCEL;
      // calc_global_paths (frenet_optimal_planner.py:110-119)
      double px, py, tx, ty;
      if (spline_frame(sp, a.K, a.Kp, a.search_iters, s, px, py, tx, ty)) {
        // cos(yaw + pi/2) = -ty/|t|, sin(yaw + pi/2) = tx/|t| with yaw = atan2(ty, tx)
        const double r = rsqrt(tx * tx + ty * ty);
        xs[m] = px - d * (ty * r);
        ys[m] = py + d * (tx * r);
      } else {
        first_bad = min(first_bad, m);
      }
      if (kRec) {
        rec[FISS_REC_T * a.n_stride + m] = t;
        rec[FISS_REC_S * a.n_stride + m] = s;
        rec[FISS_REC_S_D * a.n_stride + m] = s_d;
        rec[FISS_REC_S_DD * a.n_stride + m] = s_dd;
        rec[FISS_REC_S_DDD * a.n_stride + m] = s_ddd;
        rec[FISS_REC_D * a.n_stride + m] = d;
        rec[FISS_REC_D_D * a.n_stride + m] = d_d;
        rec[FISS_REC_D_DD * a.n_stride + m] = d_dd;
        rec[FISS_REC_D_DDD * a.n_stride + m] = d_ddd;
      }
      if (kMat && mat_v) mat_v[m] = s_d;
    } else {
      if (kRec) {
#pragma unroll
        for (int r = 0; r <= FISS_REC_D_DDD; ++r) rec[r * a.n_stride + m] = CUDART_NAN;
      }
      if (kMat && mat_v) mat_v[m] = CUDART_NAN;
    }
  }
  acc = warp_sum(acc);
  viol = warp_or(viol);
  const int n_cart = warp_min(first_bad);  // n' (frenet_optimal_planner.py:112-113)
  const double t_last = (n - 1) * p.tick_t;
  const double cost = ((p.cost_time_offset - t_last) + acc) / (double)n;  // cost_function.py:42,49
  __syncwarp();

  // ---- P2: headings.  yaw_m = atan2(dy, dx) for m < n'-1, last repeated (:127-130)
  const int horizon = min(n_cart, a.final_time_step - p.time_step_now);  // t_step_max (:173-174)
  const bool do_coll = a.M > 0 && horizon > 0 && (p.collide_all || viol == 0);
  const int E = do_coll ? (horizon + p.check_res - 1) / p.check_res : 0;  // checked steps i = e*check_res
  if (kYaw) {
    for (int m = lane; m < n_cart - 1; m += 32) yw[m] = atan2(ys[m + 1] - ys[m], xs[m + 1] - xs[m]);
    __syncwarp();
    if (lane == 0 && n_cart >= 2) yw[n_cart - 1] = yw[n_cart - 2];
    __syncwarp();
  }
  if (n_cart >= 2) {
    for (int e = lane; e < E; e += 32) {
      const int i = e * p.check_res;
      const int seg = min(i, n_cart - 2);
      const double dx = xs[seg + 1] - xs[seg];
      const double dy = ys[seg + 1] - ys[seg];
      const double h2 = dx * dx + dy * dy;
      double c, s;
      if (h2 > 0.0 && h2 < 1.0e300) {
        const double r = rsqrt(h2);  // cos/sin of atan2(dy, dx) without the round trip
        c = dx * r;
        s = dy * r;
      } else {
        sincos(atan2(dy, dx), &s, &c);
      }
      ex[e] = xs[i];
      ey[e] = ys[i];
      ec[e] = c;
      es[e] = s;
    }
  }
  __syncwarp();

  // ---- P3: collision (has_collision, frenet_optimal_planner.py:168-195)
  bool hit = false;
  if (do_coll) {
    if (n_cart < 2) {
      hit = true;  // traj.yaw[0] raises inside the try: "Failed to create Polygon" => collision (:178-182)
    } else {
      const double hle = 0.5 * p.ego_length, hwe = 0.5 * p.ego_width;
      const double re = sqrt(hle * hle + hwe * hwe);
      const int Mp = a.Mp;
      const int row0 = a.obs_in_smem ? 0 : p.time_step_now;
      const int row_step = a.obs_in_smem ? 1 : p.check_res;
      if (Mp <= 32) {
        const int j = lane & (Mp - 1);
        const int e_off = lane >> a.mp_shift;
        const int spi = 32 >> a.mp_shift;  // checked steps per iteration
        const double hlo = oc[j], hwo = oc[Mp + j];
        const double thr = re + oc[2 * Mp + j];
        const double thr2 = thr * thr;
        for (int e0 = 0; e0 < E; e0 += spi) {
          const int e = e0 + e_off;
          bool h = false;
          const int row = row0 + e * row_step;
          if (e < E && (a.obs_in_smem || row < a.T_obs)) {
            const double* slot = obs + (int64_t)row * 4 * Mp + j;  // conflict-free: lanes read consecutive doubles
            const double dx = slot[0] - ex[e], dy = slot[Mp] - ey[e];
            if (dx * dx + dy * dy <= thr2)
              h = rect_sat(dx, dy, ec[e], es[e], hle, hwe, slot[2 * Mp], slot[3 * Mp], hlo, hwo);
          }
          if (__any_sync(kFull, h)) {
            hit = true;
            break;
          }
        }
      } else {
        for (int e = 0; e < E && !hit; ++e) {
          const double x = ex[e], y = ey[e], c = ec[e], s = es[e];
          bool h = false;
          const int row = row0 + e * row_step;
          if (!a.obs_in_smem && row >= a.T_obs) break;  // past the predictions: no obstacle has a state
          for (int j = lane; j < Mp; j += 32) {
            const double* slot = obs + (int64_t)row * 4 * Mp + j;
            const double dx = slot[0] - x, dy = slot[Mp] - y;
            const double thr = re + oc[2 * Mp + j];
            if (dx * dx + dy * dy <= thr * thr)
              h = h || rect_sat(dx, dy, c, s, hle, hwe, slot[2 * Mp], slot[3 * Mp], oc[j], oc[Mp + j]);
          }
          hit = __any_sync(kFull, h);
        }
      }
    }
  }

  // ---- curvature rows / optional curvature mask, materialisation
  unsigned curv = 0;
  if (kYaw) {
    double* mx = (kMat && a.mat) ? a.mat + ((int64_t)FISS_MAT_X * a.total + out_id) * a.n_stride : nullptr;
    double* my = (kMat && a.mat) ? a.mat + ((int64_t)FISS_MAT_Y * a.total + out_id) * a.n_stride : nullptr;
    double* myaw = (kMat && a.mat) ? a.mat + ((int64_t)FISS_MAT_YAW * a.total + out_id) * a.n_stride : nullptr;
    double* mk = (kMat && a.mat) ? a.mat + ((int64_t)FISS_MAT_KAPPA * a.total + out_id) * a.n_stride : nullptr;
    const double dt = p.tick_t;
    const int m_end = (kRec || (kMat && mx)) ? a.n_stride : n_cart;
    for (int m = lane; m < m_end; m += 32) {
      const bool in_cart = m < n_cart;
      // with n' < 2 the reference leaves yaw/ds/c empty (:121)
      const bool has_c = n_cart >= 2 && m < n_cart - 1;
      double ds = CUDART_NAN, c0 = CUDART_NAN, c_d = CUDART_NAN, c_dd = CUDART_NAN;
      if (has_c) {
        ds = hypot(xs[m + 1] - xs[m], ys[m + 1] - ys[m]);            // :128
        c0 = (yw[m + 1] - yw[m]) / ds;                               // :132 (no unwrap)
        if (p.check_curvature && fabs(c0) > p.max_curvature) curv = FISS_FLAG_CURVATURE;
        if (kRec && m < n_cart - 2) {
          const double ds1 = hypot(xs[m + 2] - xs[m + 1], ys[m + 2] - ys[m + 1]);
          const double c1 = (yw[m + 2] - yw[m + 1]) / ds1;
          c_d = (c1 - c0) / dt;                                     // :133
          if (m < n_cart - 3) {
            const double ds2 = hypot(xs[m + 3] - xs[m + 2], ys[m + 3] - ys[m + 2]);
            const double c2 = (yw[m + 3] - yw[m + 2]) / ds2;
            c_dd = ((c2 - c1) / dt - c_d) / dt;                    // :134
          }
        }
      }
      const double xv = in_cart ? xs[m] : CUDART_NAN;
      const double yv = in_cart ? ys[m] : CUDART_NAN;
      const double yawv = (in_cart && n_cart >= 2) ? yw[m] : CUDART_NAN;
      if (kMat && mx) {
        mx[m] = xv;
        my[m] = yv;
        myaw[m] = yawv;
        mk[m] = c0;
      }
      if (kRec) {
        rec[FISS_REC_X * a.n_stride + m] = xv;
        rec[FISS_REC_Y * a.n_stride + m] = yv;
        rec[FISS_REC_YAW * a.n_stride + m] = yawv;
        rec[FISS_REC_DS * a.n_stride + m] = ds;
        rec[FISS_REC_C * a.n_stride + m] = c0;
        rec[FISS_REC_C_D * a.n_stride + m] = c_d;
        rec[FISS_REC_C_DD * a.n_stride + m] = c_dd;
      }
    }
    if (p.check_curvature) curv = warp_or(curv);
  }
  __syncwarp();

  cost_out = cost;
  flags_out = viol | curv | (hit ? FISS_FLAG_COLLISION : 0u) | ((uint32_t)n_cart << FISS_FLAG_NCART_SHIFT);
}

// ------------------------------------------------------------------------------------------------
// Shared-memory layout of the persistent CTA (doubles, after a 16-byte mbarrier slot):
//   spline [9][Kp] | obstacle consts [4][Mp] | obstacle rows [E_max][4][Mp] (if obs_in_smem) |
//   per-warp scratch [kWarpsPerCta][3 * n_pad + 4 * e_cap]
__host__ __device__ inline int64_t scratch_doubles_per_warp(int n_pad, int e_cap) {
  return 3 * (int64_t)n_pad + 4 * (int64_t)e_cap;
}

template <bool kMat, bool kRec>
__global__ void __launch_bounds__(kThreads) fiss_eval_kernel(const EvalArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
  double* sp = reinterpret_cast<double*>(smem_raw + 16);
  double* oc = sp + 9 * (int64_t)a.Kp;
  double* obs_s = oc + 4 * (int64_t)a.Mp;
  const int64_t obs_doubles = a.obs_in_smem ? (int64_t)a.E_max * a.Mp * 4 : 0;
  double* scratch_base = obs_s + obs_doubles;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // ---- stage the tables: one elected thread issues the bulk copies, everybody waits on the mbarrier
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  __syncthreads();
  const uint32_t spline_bytes = 9u * a.Kp * 8u;
  const uint32_t const_bytes = 4u * a.Mp * 8u;
  const uint32_t row_bytes = 4u * a.Mp * 8u;
  int rows_live = 0;  // staged rows that exist in the table (time < T_obs)
  if (a.obs_in_smem) {
    for (int e = 0; e < a.E_max; ++e)
      if (a.p.time_step_now + e * a.p.check_res < a.T_obs) rows_live = e + 1;
  }
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar, spline_bytes + (a.Mp > 0 ? const_bytes : 0u) + (uint32_t)rows_live * row_bytes);
    bulk_g2s(sp, a.spline, spline_bytes, bar);
    if (a.Mp > 0) bulk_g2s(oc, a.obs_const, const_bytes, bar);
    for (int e = 0; e < rows_live; ++e)
      bulk_g2s(obs_s + (int64_t)e * a.Mp * 4,
               a.obs_tab + (int64_t)(a.p.time_step_now + e * a.p.check_res) * a.Mp * 4, row_bytes, bar);
  }
  // rows past the end of the predictions: nobody has a state there (state_at_time -> None)
  if (a.obs_in_smem) {
    for (int64_t q = (int64_t)rows_live * a.Mp * 4 + threadIdx.x; q < obs_doubles; q += blockDim.x)
      obs_s[q] = (((q / a.Mp) & 3) < 2) ? kObsFar : 0.0;
  }
  mbar_wait(bar, 0);
  if (kRec && a.after_producer) pdl_wait_producer();
  __syncthreads();

  const double* obs = a.obs_in_smem ? obs_s : a.obs_tab;
  double* scratch = scratch_base + warp * scratch_doubles_per_warp(a.n_pad, a.e_cap);
  const int wpc = blockDim.x >> 5;  // the host launches fewer warps per CTA when there are few candidates
  const int64_t warps_total = (int64_t)gridDim.x * wpc;
  for (int64_t id = (int64_t)blockIdx.x * wpc + warp; id < a.total; id += warps_total) {
    int b, c;
    if (kRec) {
      int c_sel;
      if (a.pick_cost) {  // fused pick: this warp owns problem `id`
        const double* pc = a.pick_cost + id * (int64_t)a.C;
        const uint32_t* pf = a.pick_flags + id * (int64_t)a.C;
        double bc;
        pick_scan(pc, pf, a.C, lane, 32, bc, c_sel);
        pick_warp_reduce(bc, c_sel);
        if (lane == 0) {
          a.pick_idx[id] = c_sel;
          a.pick_best[id] = c_sel >= 0 ? bc : CUDART_INF;
          if (a.pick_meta) {  // (n, n') of the winner, for the host to cut the ragged record rows
            a.pick_meta[2 * id] = c_sel >= 0 ? (int)a.end[4 * (int64_t)c_sel + 3] : 0;
            a.pick_meta[2 * id + 1] =
                c_sel >= 0 ? (int)((pf[c_sel] >> FISS_FLAG_NCART_SHIFT) & FISS_FLAG_NCART_MASK) : 0;
          }
        }
      } else {
        c_sel = a.sel ? a.sel[id] : (int)id;
      }
      b = a.per_problem ? (int)id : 0;
      c = c_sel;
      if (c < 0) {  // no winner for this problem: an all-NaN record
        double* rec = a.records + id * (int64_t)(FISS_REC_ROWS * a.n_stride);
        for (int q = lane; q < FISS_REC_ROWS * a.n_stride; q += 32) rec[q] = CUDART_NAN;
        if (lane == 0) {
          if (a.cost) a.cost[id] = CUDART_NAN;
          if (a.flags) a.flags[id] = 0;
        }
        continue;
      }
    } else {
      b = (int)(id / a.C);
      c = (int)(id - (int64_t)b * a.C);
    }
    double cost;
    uint32_t flags;
    eval_candidate<kMat, kRec>(a, sp, obs, oc, scratch, lane, a.ego + 6 * (int64_t)b, a.end + 4 * (int64_t)c, id,
                               cost, flags);
    if (lane == 0) {
      if (a.cost) a.cost[id] = cost;
      if (a.flags) a.flags[id] = flags;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// argmin per problem (tie rule: see better()).  One CTA per problem; block-level reduction through shared memory
// after a warp-shuffle stage.  Used when only the winners are wanted; with records the pick is fused into the record
// kernel (EvalArgs::pick_cost).
constexpr int kPickThreads = 128;

__global__ void __launch_bounds__(kPickThreads) fiss_pick_kernel(const double* __restrict__ cost,
                                                                 const uint32_t* __restrict__ flags, int C,
                                                                 int32_t* __restrict__ best_idx,
                                                                 double* __restrict__ best_cost,
                                                                 const double* __restrict__ end,
                                                                 int32_t* __restrict__ meta) {
  __shared__ double s_cost[kPickThreads / 32];
  __shared__ int s_idx[kPickThreads / 32];
  const int b = blockIdx.x;
  const double* pc = cost + (int64_t)b * C;
  const uint32_t* pf = flags + (int64_t)b * C;
  double bc;
  int bi;
  pick_scan(pc, pf, C, threadIdx.x, kPickThreads, bc, bi);
  pick_warp_reduce(bc, bi);
  if ((threadIdx.x & 31) == 0) {
    s_cost[threadIdx.x >> 5] = bc;
    s_idx[threadIdx.x >> 5] = bi;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < kPickThreads / 32; ++w) {
      if (s_idx[w] >= 0 && (bi < 0 || better(s_cost[w], s_idx[w], bc, bi))) {
        bc = s_cost[w];
        bi = s_idx[w];
      }
    }
    best_idx[b] = bi;
    best_cost[b] = bi >= 0 ? bc : CUDART_INF;
    if (meta) {  // (n, n') of the winner, for the host to cut the ragged record rows
      meta[2 * b] = bi >= 0 ? (int)end[4 * (int64_t)bi + 3] : 0;
      meta[2 * b + 1] = bi >= 0 ? (int)((pf[bi] >> FISS_FLAG_NCART_SHIFT) & FISS_FLAG_NCART_MASK) : 0;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Obstacle table preparation (once per scene): AoS host layout -> time-major, component-planar rows.
//   in : xyth [M][T][3], lw [M][2], valid [M][T]
//   out: tab [T][4][Mp] = rows cx, cy, cos th, sin th, with (kObsFar, kObsFar, 0, 0) where there is no state
//        oc  [4][Mp]    = rows l/2, w/2, circumscribed radius, 0
// One time step is one contiguous 32*Mp-byte row (one bulk-TMA copy), and a warp whose lanes are
// obstacles reads consecutive doubles of a component (no shared-memory bank conflicts).
__device__ __forceinline__ void obstacle_store(double* __restrict__ tab, double* __restrict__ oc, int Mp, int T,
                                               int64_t q, bool present, double x, double y, double th, double l,
                                               double w) {
  if (q < Mp) {
    const int j = (int)q;
    const double hl = 0.5 * l, hw = 0.5 * w;
    oc[j] = hl;
    oc[Mp + j] = hw;
    oc[2 * Mp + j] = sqrt(hl * hl + hw * hw);
    oc[3 * Mp + j] = 0.0;
  }
  if (q >= (int64_t)T * Mp) return;
  const int t = (int)(q / Mp);
  const int j = (int)(q - (int64_t)t * Mp);
  double cx = kObsFar, cy = kObsFar, cs = 0.0, sn = 0.0;
  if (present) {
    sincos(th, &sn, &cs);
    // shapely.affinity.rotate snaps |cos|, |sin| < 2.5e-16 to 0
    if (fabs(cs) < 2.5e-16) cs = 0.0;
    if (fabs(sn) < 2.5e-16) sn = 0.0;
    cx = x;
    cy = y;
  }
  double* row = tab + (int64_t)t * 4 * Mp + j;
  row[0] = cx;
  row[Mp] = cy;
  row[2 * Mp] = cs;
  row[3 * Mp] = sn;
}

__global__ void fiss_obstacle_prep_kernel(const double* __restrict__ xyth, const double* __restrict__ lw,
                                          const uint8_t* __restrict__ valid, int M, int Mp, int T,
                                          double* __restrict__ tab, double* __restrict__ oc) {
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int jc = (int)(q < Mp ? q : 0);
  const double l = (q < Mp && jc < M) ? lw[2 * jc] : 0.0;
  const double w = (q < Mp && jc < M) ? lw[2 * jc + 1] : 0.0;
  bool present = false;
  double x = 0.0, y = 0.0, th = 0.0;
  if (q < (int64_t)T * Mp) {
    const int t = (int)(q / Mp);
    const int j = (int)(q - (int64_t)t * Mp);
    if (j < M && valid[(int64_t)j * T + t]) {
      const double* s = xyth + ((int64_t)j * T + t) * 3;
      present = true;
      x = s[0];
      y = s[1];
      th = s[2];
    }
  }
  obstacle_store(tab, oc, Mp, T, q, present, x, y, th, l, w);
}

// Waymo wire format (waymo_interface.py:24-76): float32 [N][T][11] = (x, y, z, l, w, h, heading, vx, vy, valid, type).
// The host has already applied the conversion's keep / cut rules (fiss_set_obstacles_waymo): table column j is agent
// keep[j], which has a state at steps 0..t_end[j] -- step 0 is the initial state, never masked (:36-40) -- and none
// behind (state_at_time -> None); length / width from step 0 (:33-34).  float32 values widen to double exactly, as they
// do when the reference hands them to shapely.
__global__ void fiss_obstacle_prep_waymo_kernel(const float* __restrict__ trajs, const int32_t* __restrict__ keep,
                                                const int32_t* __restrict__ t_end, int M, int Mp, int T,
                                                double* __restrict__ tab, double* __restrict__ oc) {
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int jc = (int)(q < Mp ? q : 0);
  const bool has_c = q < Mp && jc < M;
  const int64_t agent_c = has_c ? keep[jc] : 0;
  const double l = has_c ? (double)trajs[(agent_c * T) * 11 + 3] : 0.0;
  const double w = has_c ? (double)trajs[(agent_c * T) * 11 + 4] : 0.0;
  bool present = false;
  double x = 0.0, y = 0.0, th = 0.0;
  if (q < (int64_t)T * Mp) {
    const int t = (int)(q / Mp);
    const int j = (int)(q - (int64_t)t * Mp);
    if (j < M && t <= t_end[j]) {
      const float* s = trajs + ((int64_t)keep[j] * T + t) * 11;
      present = true;
      x = (double)s[0];
      y = (double)s[1];
      th = (double)s[6];
    }
  }
  obstacle_store(tab, oc, Mp, T, q, present, x, y, th, l, w);
}

}  // namespace fiss
