// fiss_record_kernel.cuh -- the winners' full records of a plan step (pick fused in): FrenetOptimalPlanner.plan()'s
// argmin (frenet_optimal_planner.py:263-268) and the winning FrenetTrajectory's sixteen arrays (frenet.py:131-148;
// calc_frenet_paths :79-99 + calc_global_paths :106-138 for ONE candidate per problem).
//
// The generic list kernel (fiss_eval_kernel<false, true>) does this with one warp per problem: two passes over the steps,
// three hypot + three divisions per step for c, c_d, c_dd, everything behind one warp's dependency chains -- 12 us for 512
// problems, a tenth of a plan step and a third of a one-problem plan cycle.  Here W warps share a problem (W = 2 covers
// the 50-step horizons in ONE pass, W = 4 the 100-step ones), the finite-difference chains go through small per-problem
// tables in shared memory (x, y, yaw, ds, c: each value computed once, by the lane that owns the step), and the winner's
// step counts n, n' come from the volume the lattice kernel already wrote instead of being reduced again.  The values
// are the list kernel's, expression for expression (same polynomials, same library atan2 / hypot / division, c_d and
// c_dd from the same neighbouring c values), so records do not depend on which kernel wrote them.
#pragma once

#include "fiss_kernels.cuh"

namespace fiss {

constexpr int kRecWarps = 8;  // warps per CTA: kRecWarps / W problems per CTA

__device__ __forceinline__ void group_barrier(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// Registers capped for three CTAs per SM (77-80 instead of 98, no spills): behind a chained winner-only lattice launch a record
// CTA then fits into the slot ONE retiring lattice CTA leaves (256 threads x 80 registers) instead of waiting for two
// (winner-only step train 0.0586 -> 0.0559 ms).
#ifndef FISS_REC_MIN_CTAS
#define FISS_REC_MIN_CTAS 3
#endif
#define FISS_REC_BOUNDS __launch_bounds__(kRecWarps * 32, FISS_REC_MIN_CTAS)
template <int W>
__global__ void FISS_REC_BOUNDS fiss_record_kernel(const EvalArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
  double* sp = reinterpret_cast<double*>(smem_raw + 16);
  constexpr int kGroups = kRecWarps / W;
  const int n_pad = a.n_pad;
  double* tables = sp + 9 * (int64_t)a.Kp;  // [kGroups][5][n_pad]: x, y, yaw, ds, c of the group's problem
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int group = warp / W, gl = (warp - group * W) * 32 + lane;  // lane within the group, 0 .. 32 W - 1
  const fiss_params& p = a.p;
  const int ns = a.n_stride;

  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t spline_bytes = 9u * a.Kp * 8u;
    mbar_expect_tx(bar, spline_bytes);
    bulk_g2s(sp, a.spline, spline_bytes, bar);
  }
  mbar_wait(bar, 0);
  // the next launch of the stream (the next step's lattice kernel, when it is launched as a programmatic dependent: chained
  // launches, fiss_abi.cu eval_grid) may be scheduled: its CTAs take the SMs that this step's last items leave idle, and
  // wait for this kernel before they write what it reads
  pdl_launch_dependents();
  if (a.after_producer) pdl_wait_producer();  // the lattice kernel's cost / flags volume
  __syncthreads();

  double* xs = tables + (int64_t)group * 5 * n_pad;
  double* ys = xs + n_pad;
  double* yw = ys + n_pad;
  double* dsv = yw + n_pad;
  double* cs = dsv + n_pad;
  const int bar_id = 1 + group;

  for (int64_t id = (int64_t)blockIdx.x * kGroups + group; id < a.total; id += (int64_t)gridDim.x * kGroups) {
    // ---- the argmin, by every warp of the group for itself (the scan is deterministic: no broadcast, no barrier)
    const double* pc = a.pick_cost + id * (int64_t)a.C;
    const uint32_t* pf = a.pick_flags + id * (int64_t)a.C;
    double bc;
    int c_sel;
    pick_scan(pc, pf, a.C, lane, 32, bc, c_sel);
    pick_warp_reduce(bc, c_sel);
    const int n = c_sel >= 0 ? (int)a.end[4 * (int64_t)c_sel + 3] : 0;
    const int n_cart = c_sel >= 0 ? (int)((__ldcg(pf + c_sel) >> FISS_FLAG_NCART_SHIFT) & FISS_FLAG_NCART_MASK) : 0;
    if (gl == 0) {
      a.pick_idx[id] = c_sel;
      a.pick_best[id] = c_sel >= 0 ? bc : CUDART_INF;
      if (a.pick_meta) {
        a.pick_meta[2 * id] = n;
        a.pick_meta[2 * id + 1] = n_cart;
      }
    }
    double* rec = a.records + id * (int64_t)(FISS_REC_ROWS * ns);
    if (c_sel < 0) {  // nothing feasible: an all-NaN record
      for (int q = gl; q < FISS_REC_ROWS * ns; q += 32 * W) rec[q] = CUDART_NAN;
      continue;
    }
    const double* ego = a.ego + 6 * id;
    const double* end = a.end + 4 * (int64_t)c_sel;
    const double s0 = ego[0], v0 = ego[1], a0 = ego[2], d0 = ego[3], dv0 = ego[4], da0 = ego[5];
    const double d_end = end[0], v_end = end[1], T = end[2];
    // coefficients: the closed forms of eval_candidate
    const double T2 = T * T, T3 = T2 * T;
    const double iT = 1.0 / T;
    const double iT2 = iT * iT, iT3 = iT2 * iT;
    const double qa2 = 0.5 * a0;
    const double Vq = v_end - v0 - 2.0 * qa2 * T;
    const double Aq = -2.0 * qa2;
    const double qa3 = (3.0 * Vq - Aq * T) / (3.0 * T2);
    const double qa4 = (Aq * T - 2.0 * Vq) / (4.0 * T3);
    const double la2 = 0.5 * da0;
    const double Dl = d_end - d0 - dv0 * T - la2 * T2;
    const double Vl = -dv0 - 2.0 * la2 * T;
    const double Al = -2.0 * la2;
    const double la3 = (10.0 * Dl - 4.0 * Vl * T + 0.5 * Al * T2) * iT3;
    const double la4 = (-15.0 * Dl + 7.0 * Vl * T - Al * T2) * (iT3 * iT);
    const double la5 = (6.0 * Dl - 3.0 * Vl * T + 0.5 * Al * T2) * (iT3 * iT2);

    // ---- pass 1: Frenet rows and positions
    for (int m = gl; m < ns; m += 32 * W) {
      double xv = CUDART_NAN, yv = CUDART_NAN;
      if (m < n) {
        const double t = m * p.tick_t;
        const double t2 = t * t, t3 = t2 * t, t4 = t3 * t, t5 = t4 * t;
        const double s = s0 + v0 * t + qa2 * t2 + qa3 * t3 + qa4 * t4;
        rec[FISS_REC_T * ns + m] = t;
        rec[FISS_REC_S * ns + m] = s;
        rec[FISS_REC_S_D * ns + m] = v0 + 2.0 * qa2 * t + 3.0 * qa3 * t2 + 4.0 * qa4 * t3;
        rec[FISS_REC_S_DD * ns + m] = 2.0 * qa2 + 6.0 * qa3 * t + 12.0 * qa4 * t2;
        rec[FISS_REC_S_DDD * ns + m] = 6.0 * qa3 + 24.0 * qa4 * t;
        const double d = d0 + dv0 * t + la2 * t2 + la3 * t3 + la4 * t4 + la5 * t5;
        rec[FISS_REC_D * ns + m] = d;
        rec[FISS_REC_D_D * ns + m] = dv0 + 2.0 * la2 * t + 3.0 * la3 * t2 + 4.0 * la4 * t3 + 5.0 * la5 * t4;
        rec[FISS_REC_D_DD * ns + m] = 2.0 * la2 + 6.0 * la3 * t + 12.0 * la4 * t2 + 20.0 * la5 * t3;
        rec[FISS_REC_D_DDD * ns + m] = 6.0 * la3 + 24.0 * la4 * t + 60.0 * la5 * t2;
        double px, py, tx, ty;
        if (m < n_cart && spline_frame(sp, a.K, a.Kp, a.search_iters, s, px, py, tx, ty)) {
          const double r = rsqrt(tx * tx + ty * ty);
          xv = px - d * (ty * r);
          yv = py + d * (tx * r);
        }
      } else {
#pragma unroll
        for (int r = 0; r <= FISS_REC_D_DDD; ++r) rec[r * ns + m] = CUDART_NAN;
      }
      if (m < n_pad) {
        xs[m] = xv;
        ys[m] = yv;
      }
      rec[FISS_REC_X * ns + m] = xv;
      rec[FISS_REC_Y * ns + m] = yv;
    }
    group_barrier(bar_id, 32 * W);
    // ---- pass 2: segment headings and lengths (:127-128)
    for (int m = gl; m < n_cart - 1; m += 32 * W) {
      const double dx = xs[m + 1] - xs[m], dy = ys[m + 1] - ys[m];
      yw[m] = atan2(dy, dx);
      dsv[m] = hypot(dx, dy);
    }
    group_barrier(bar_id, 32 * W);
    // ---- pass 3: yaw row (last point repeats the previous heading, :129-130), ds row, curvature (:132; no unwrap)
    const bool has_yaw = n_cart >= 2;  // with n' < 2 the reference leaves yaw / ds / c empty (:121)
    for (int m = gl; m < ns; m += 32 * W) {
      double yawv = CUDART_NAN, dsm = CUDART_NAN, c0 = CUDART_NAN;
      if (has_yaw && m < n_cart) {
        yawv = yw[min(m, n_cart - 2)];
        if (m < n_cart - 1) {
          dsm = dsv[m];
          c0 = (yw[min(m + 1, n_cart - 2)] - yw[m]) / dsm;
        }
      }
      if (m < n_pad) cs[m] = c0;
      rec[FISS_REC_YAW * ns + m] = yawv;
      rec[FISS_REC_DS * ns + m] = dsm;
      rec[FISS_REC_C * ns + m] = c0;
    }
    group_barrier(bar_id, 32 * W);
    // ---- pass 4: curvature rate rows (:133-134)
    const double dt = p.tick_t;
    for (int m = gl; m < ns; m += 32 * W) {
      double c_d = CUDART_NAN, c_dd = CUDART_NAN;
      if (has_yaw && m < n_cart - 2) {
        const double c0 = cs[m], c1 = cs[m + 1];
        c_d = (c1 - c0) / dt;
        if (m < n_cart - 3) c_dd = ((cs[m + 2] - c1) / dt - c_d) / dt;
      }
      rec[FISS_REC_C_D * ns + m] = c_d;
      rec[FISS_REC_C_DD * ns + m] = c_dd;
    }
    group_barrier(bar_id, 32 * W);  // the tables are free for the group's next problem
  }
}

__host__ __device__ inline size_t record_smem_bytes(int Kp, int n_pad, int W) {
  return 16 + (size_t)9 * Kp * 8 + (size_t)(kRecWarps / W) * 5 * n_pad * 8;
}

}  // namespace fiss
