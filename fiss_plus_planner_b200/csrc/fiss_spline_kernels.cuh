// fiss_spline_kernels.cuh -- reference-line set-up on the device.
//
// The reference fits the natural cubic spline of the route centre line on the host with a dense K x K solve
// (cubic_spline.py:19-43,118-142: A c = B via np.linalg.solve) and resamples it every 0.1 m in Python
// (frenet_optimal_planner.py:272-278).  Here one CTA fits one lane: arc-length knots (np.cumsum order), the
// tridiagonal system solved by the Thomas recurrence (one thread per coordinate -- the recurrence is sequential
// and K is tens to hundreds), then the b / d rows in parallel.  Batches of lanes (multi-lane / multi-scenario
// set-up) are independent CTAs.  Results agree with the LAPACK path to rounding (~1e-15 relative), NOT bit for
// bit, so the planners keep the host fit by default and this path is opt-in.
#pragma once

#include <cuda_runtime.h>
#include <math_constants.h>

namespace fiss {

constexpr int kFitThreads = 128;

// xy [L][K][2] -> tables [L][9][Kp]: knots, ax, bx, cx, dx, ay, by, cy, dy (b, d padded with 0 in the last slot,
// knots padded with +inf, like fiss_set_spline).  Dynamic shared memory: 6 * Kp doubles.
__global__ void __launch_bounds__(kFitThreads) fiss_spline_fit_kernel(const double* __restrict__ xy, int K, int Kp,
                                                                      double* __restrict__ tables) {
  extern __shared__ __align__(16) double fit_smem[];
  double* s = fit_smem;          // knots
  double* hh = s + Kp;           // h[k] = s[k+1] - s[k]   (np.diff of the knots, cubic_spline.py:24)
  double* cp = hh + Kp;          // Thomas c' / d' for x, then for y
  double* dp = cp + Kp;
  double* cq = dp + Kp;
  double* dq = cq + Kp;
  const double* pts = xy + (int64_t)blockIdx.x * K * 2;
  double* tab = tables + (int64_t)blockIdx.x * 9 * Kp;
  // segment lengths, then the knots in np.cumsum order (cubic_spline.py:162-168)
  for (int k = threadIdx.x; k < K - 1; k += blockDim.x)
    hh[k] = hypot(pts[2 * (k + 1)] - pts[2 * k], pts[2 * (k + 1) + 1] - pts[2 * k + 1]);
  __syncthreads();
  if (threadIdx.x == 0) {
    double acc = 0.0;
    s[0] = 0.0;
    for (int k = 0; k < K - 1; ++k) {
      acc += hh[k];
      s[k + 1] = acc;
    }
  }
  __syncthreads();
  for (int k = threadIdx.x; k < Kp; k += blockDim.x) {
    tab[k] = k < K ? s[k] : CUDART_INF;
    tab[Kp + k] = k < K ? pts[2 * k] : 0.0;          // a_x = x
    tab[5 * Kp + k] = k < K ? pts[2 * k + 1] : 0.0;  // a_y = y
  }
  __syncthreads();
  for (int k = threadIdx.x; k < K - 1; k += blockDim.x) hh[k] = s[k + 1] - s[k];
  __syncthreads();
  // natural spline: c[0] = c[K-1] = 0; h[i-1] c[i-1] + 2 (h[i-1] + h[i]) c[i] + h[i] c[i+1] = B[i]
  // with B[i] = 3 (a[i+1] - a[i]) / h[i] - 3 (a[i] - a[i-1]) / h[i-1]   (cubic_spline.py:118-142)
  if (threadIdx.x < 2) {
    const int co = threadIdx.x;  // 0: x, 1: y
    double* cpr = co ? cq : cp;
    double* dpr = co ? dq : dp;
    double* c = tab + (co ? 7 : 3) * Kp;
    const double* a = pts + co;
    double cprev = 0.0, dprev = 0.0;  // row 0: c[0] = 0
    for (int i = 1; i < K - 1; ++i) {
      const double B = 3.0 * (a[2 * (i + 1)] - a[2 * i]) / hh[i] - 3.0 * (a[2 * i] - a[2 * (i - 1)]) / hh[i - 1];
      const double m = 2.0 * (hh[i - 1] + hh[i]) - hh[i - 1] * cprev;
      cprev = hh[i] / m;
      dprev = (B - hh[i - 1] * dprev) / m;
      cpr[i] = cprev;
      dpr[i] = dprev;
    }
    double cnext = 0.0;  // c[K-1] = 0
    c[K - 1] = 0.0;
    for (int i = K - 2; i >= 1; --i) {
      cnext = dpr[i] - cpr[i] * cnext;
      c[i] = cnext;
    }
    c[0] = 0.0;
    for (int k = K; k < Kp; ++k) c[k] = 0.0;
  }
  __syncthreads();
  // d = (c[k+1] - c[k]) / (3 h);  b = 1/h (a[k+1] - a[k]) - h/3 (2 c[k] + c[k+1])   (cubic_spline.py:38-43)
  for (int q = threadIdx.x; q < 2 * Kp; q += blockDim.x) {
    const int co = q >= Kp, k = q - co * Kp;
    const double* a = tab + (co ? 5 : 1) * Kp;
    const double* c = tab + (co ? 7 : 3) * Kp;
    double b = 0.0, d = 0.0;
    if (k < K - 1) {
      const double h = hh[k];
      d = (c[k + 1] - c[k]) / (3.0 * h);
      b = 1.0 / h * (a[k + 1] - a[k]) - h / 3.0 * (2.0 * c[k] + c[k + 1]);
    }
    tab[(co ? 6 : 2) * Kp + k] = b;
    tab[(co ? 8 : 4) * Kp + k] = d;
  }
}

// ref [m][4] = (x, y, yaw, curvature) at s_i = i * step (np.arange(0, s_end, 0.1): start + i*step), the polyline
// FrenetState.from_state consumes (frenet_optimal_planner.py:274-278; cubic_spline.py:170-232).
__global__ void fiss_frame_samples_kernel(const double* __restrict__ sp, int K, int Kp, double step, int m,
                                          double* __restrict__ ref) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const double s = i * step;
  double x = CUDART_NAN, y = CUDART_NAN, yaw = CUDART_NAN, kap = CUDART_NAN;
  if (s >= sp[0] && s <= sp[K - 1]) {
    int lo = 0, hi = K - 1;  // bisect_right(knots, s) - 1, clamped to the last segment for s == s_end
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (sp[mid] <= s) lo = mid;
      else hi = mid;
    }
    const double dx = s - sp[lo];
    const double ax = sp[Kp + lo], bx = sp[2 * Kp + lo], cx = sp[3 * Kp + lo], ex = sp[4 * Kp + lo];
    const double ay = sp[5 * Kp + lo], by = sp[6 * Kp + lo], cy = sp[7 * Kp + lo], ey = sp[8 * Kp + lo];
    x = ax + bx * dx + cx * (dx * dx) + ex * (dx * dx * dx);
    y = ay + by * dx + cy * (dx * dx) + ey * (dx * dx * dx);
    const double vx = bx + 2.0 * cx * dx + 3.0 * ex * (dx * dx), vy = by + 2.0 * cy * dx + 3.0 * ey * (dx * dx);
    const double wx = 2.0 * cx + 6.0 * ex * dx, wy = 2.0 * cy + 6.0 * ey * dx;
    yaw = atan2(vy, vx);
    kap = (wy * vx - wx * vy) / pow(vx * vx + vy * vy, 1.5);
  }
  ref[4 * i + 0] = x;
  ref[4 * i + 1] = y;
  ref[4 * i + 2] = yaw;
  ref[4 * i + 3] = kap;
}

}  // namespace fiss
