// fiss_math.cuh -- FP64 heading / arc-length helpers of the materialising path.
//
// calc_global_paths (frenet_optimal_planner.py:121-134) needs, per step, yaw = arctan2(dy, dx),
// ds = hypot(dx, dy) and kappa = dyaw / ds.  The CUDA library versions cost ~125 + 46 + 21 issue slots
// per warp pass, most of it special-case handling; a trajectory segment is an ordinary finite vector, so
// the hot path uses a branch-free polynomial atan2 and one rsqrt, and falls back to the library for
// zero-length / non-finite segments (where the reference's NaN / 0 results must be reproduced exactly).
#pragma once

#include <cuda_runtime.h>

namespace fiss {

// atan(q) = q + q^3 * P(q^2) on q in [0, 1]: Chebyshev-node interpolant of atan(sqrt(u))/sqrt(u), degree 19
// in u (tools/fit_atan.py: max relative error of the float64 Horner evaluation 2.5e-16).
__device__ __forceinline__ double atan_unit(double q) {
  const double u = q * q;
  double p = -1.93423475928923e-05;
  p = fma(p, u, 0.00021423810738603946);
  p = fma(p, u, -0.0011252544302234645);
  p = fma(p, u, 0.003751138483965141);
  p = fma(p, u, -0.00899108054265826);
  p = fma(p, u, 0.01671959606350739);
  p = fma(p, u, -0.02556862364437174);
  p = fma(p, u, 0.03387126702700675);
  p = fma(p, u, -0.040811247503178855);
  p = fma(p, u, 0.04668745304848529);
  p = fma(p, u, -0.052374234719188166);
  p = fma(p, u, 0.058768281144872724);
  p = fma(p, u, -0.06665764910689723);
  p = fma(p, u, 0.07692198997458294);
  p = fma(p, u, -0.09090899793217341);
  p = fma(p, u, 0.11111110578002083);
  p = fma(p, u, -0.14285714266926733);
  p = fma(p, u, 0.1999999999964796);
  p = fma(p, u, -0.333333333333307);
  return fma(q * u, p, q);
}

// atan2 for a finite, non-degenerate vector (max(|x|, |y|) in the normal range); ~3e-16 relative.
__device__ __forceinline__ double atan2_finite(double y, double x) {
  const double ax = fabs(x), ay = fabs(y);
  const bool steep = ay > ax;
  const double mx = steep ? ay : ax, mn = steep ? ax : ay;
  double a = atan_unit(mn / mx);
  if (steep) a = 1.5707963267948966 - a;
  if (x < 0.0) a = 3.141592653589793 - a;
  return copysign(a, y);
}

// One segment (dx, dy): heading and 1/ds.  `ok` = the fast path applied (squared length in the normal range);
// otherwise the caller uses the library (atan2 / hypot / division) to reproduce the reference's edge cases.
__device__ __forceinline__ bool segment_fast(double dx, double dy, double& yaw, double& inv_ds) {
  const double h2 = fma(dx, dx, dy * dy);
  if (!(h2 > 1.0e-280 && h2 < 1.0e280)) return false;
  inv_ds = rsqrt(h2);
  yaw = atan2_finite(dy, dx);
  return true;
}

}  // namespace fiss
