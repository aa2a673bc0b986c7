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
// in u (tools/fit_atan.py: max relative error of the float64 Horner evaluation 2.5e-16).  kAtanP[k] multiplies
// q^(2k+3).  In constant memory so that every FMA takes its coefficient as a constant-bank operand (literals
// cost two uniform-register moves per coefficient).
__constant__ double kAtanP[19] = {
    -0.333333333333307,     0.1999999999964796,   -0.14285714266926733,  0.11111110578002083,
    -0.09090899793217341,   0.07692198997458294,  -0.06665764910689723,  0.058768281144872724,
    -0.052374234719188166,  0.04668745304848529,  -0.040811247503178855, 0.03387126702700675,
    -0.02556862364437174,   0.01671959606350739,  -0.00899108054265826,  0.003751138483965141,
    -0.0011252544302234645, 0.00021423810738603946, -1.93423475928923e-05};

__device__ __forceinline__ double atan_unit(double q) {
  // P(u) = E(w) + u * O(w) with w = u^2: two independent Horner chains, half the dependent-FMA latency
  const double u = q * q;
  const double w = u * u;
  double e = kAtanP[18];
  double o = kAtanP[17];
#pragma unroll
  for (int k = 16; k >= 2; k -= 2) {
    e = fma(e, w, kAtanP[k]);
    o = fma(o, w, kAtanP[k - 1]);
  }
  e = fma(e, w, kAtanP[0]);
  return fma(q * u, fma(u, o, e), q);
}

// mn / mx for 0 <= mn <= mx, mx in the normal range: reciprocal seed (MUFU.RCP64H, ~2^-20), one cubic Newton
// step r(1 + e + e^2) and one residual correction (the generic division's special cases are not needed); <= 1 ulp.
__device__ __forceinline__ double ratio_unit(double mn, double mx) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(mx));
  const double e = fma(-mx, r, 1.0);
  r = fma(r, fma(e, e, e), r);
  const double q = mn * r;
  return fma(r, fma(-mx, q, mn), q);
}

// 1 / sqrt(h2) for h2 in the normal range: seed (MUFU.RSQ64H) + the cubic step y(1 + e/2 + 3e^2/8), e = 1 - h2 y^2
// -- the refinement the CUDA library's rsqrt() applies, in the library's own operation order (so that the result is
// the library's for every normal input), without its special-case branch.
__device__ __forceinline__ double rsqrt_normal(double h2) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(h2));
  const double e = fma(h2, -(y * y), 1.0);
  return fma(fma(e, 0.375, 0.5), y * e, y);
}

// atan2 for a finite, non-degenerate vector (max(|x|, |y|) in the normal range); ~3e-16 relative.
// Octant fix-up folded into one add: result = off + (+-a), off in {0, pi/2, pi}.
__device__ __forceinline__ double atan2_finite(double y, double x) {
  const double ax = fabs(x), ay = fabs(y);
  const bool steep = ay > ax, neg = x < 0.0;
  const double mx = steep ? ay : ax, mn = steep ? ax : ay;
  const double a = atan_unit(ratio_unit(mn, mx));
  const double off = steep ? 1.5707963267948966 : (neg ? 3.141592653589793 : 0.0);
  const double sa = (steep != neg) ? -a : a;
  return copysign(off + sa, y);
}

// One segment (dx, dy): heading and 1/ds, straight-line for every input.  Returns true when the result is valid:
// squared length positive, finite and far from both ends of the exponent range (2^-767 <= h2 < 2^768, one integer
// test on the exponent field).  Otherwise (zero-length / non-finite segment) the outputs are unspecified and the
// caller takes the library (atan2 / hypot / division), whose special cases are the reference's.
__device__ __forceinline__ bool segment_fast(double dx, double dy, double& yaw, double& inv_ds) {
  const double h2 = fma(dx, dx, dy * dy);
  inv_ds = rsqrt_normal(h2);
  yaw = atan2_finite(dy, dx);
  return ((unsigned)__double2hiint(h2) - 0x10000000u) < 0x60000000u;
}

// R segments in lockstep: the same arithmetic as segment_fast, written so that the R dependency chains advance together
// (ratio, polynomial and rsqrt steps interleaved in program order).  A warp's issue rate through these fixed-latency
// FP64 chains -- 8 cycles per dependent DFMA on sm_100 -- is what bounds the materialisation stage; two chains per
// lane halve the exposed latency.
template <int R>
__device__ __forceinline__ void segment_fast_n(const double (&dx)[R], const double (&dy)[R], double (&yaw)[R],
                                               double (&inv_ds)[R], bool (&valid)[R]) {
  double mx[R], mn[R], r[R], e[R], q[R], u[R], w[R], pe[R], po[R], h2[R];
  bool steep[R], neg[R];
#pragma unroll
  for (int i = 0; i < R; ++i) {
    const double ax = fabs(dx[i]), ay = fabs(dy[i]);
    steep[i] = ay > ax;
    neg[i] = dx[i] < 0.0;
    mx[i] = steep[i] ? ay : ax;
    mn[i] = steep[i] ? ax : ay;
    h2[i] = fma(dx[i], dx[i], dy[i] * dy[i]);
  }
  // ratio_unit
#pragma unroll
  for (int i = 0; i < R; ++i) asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r[i]) : "d"(mx[i]));
#pragma unroll
  for (int i = 0; i < R; ++i) e[i] = fma(-mx[i], r[i], 1.0);
#pragma unroll
  for (int i = 0; i < R; ++i) e[i] = fma(e[i], e[i], e[i]);
#pragma unroll
  for (int i = 0; i < R; ++i) r[i] = fma(r[i], e[i], r[i]);
#pragma unroll
  for (int i = 0; i < R; ++i) q[i] = mn[i] * r[i];
#pragma unroll
  for (int i = 0; i < R; ++i) e[i] = fma(-mx[i], q[i], mn[i]);
#pragma unroll
  for (int i = 0; i < R; ++i) q[i] = fma(r[i], e[i], q[i]);
  // atan_unit
#pragma unroll
  for (int i = 0; i < R; ++i) u[i] = q[i] * q[i];
#pragma unroll
  for (int i = 0; i < R; ++i) w[i] = u[i] * u[i];
#pragma unroll
  for (int i = 0; i < R; ++i) {
    pe[i] = kAtanP[18];
    po[i] = kAtanP[17];
  }
#pragma unroll
  for (int k = 16; k >= 2; k -= 2) {
#pragma unroll
    for (int i = 0; i < R; ++i) {
      pe[i] = fma(pe[i], w[i], kAtanP[k]);
      po[i] = fma(po[i], w[i], kAtanP[k - 1]);
    }
  }
#pragma unroll
  for (int i = 0; i < R; ++i) pe[i] = fma(pe[i], w[i], kAtanP[0]);
#pragma unroll
  for (int i = 0; i < R; ++i) pe[i] = fma(u[i], po[i], pe[i]);
#pragma unroll
  for (int i = 0; i < R; ++i) po[i] = q[i] * u[i];
#pragma unroll
  for (int i = 0; i < R; ++i) {
    const double a = fma(po[i], pe[i], q[i]);
    const double off = steep[i] ? 1.5707963267948966 : (neg[i] ? 3.141592653589793 : 0.0);
    const double sa = (steep[i] != neg[i]) ? -a : a;
    yaw[i] = copysign(off + sa, dy[i]);
  }
  // rsqrt_normal
  double y[R];
#pragma unroll
  for (int i = 0; i < R; ++i) asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y[i]) : "d"(h2[i]));
#pragma unroll
  for (int i = 0; i < R; ++i) e[i] = fma(h2[i], -(y[i] * y[i]), 1.0);
#pragma unroll
  for (int i = 0; i < R; ++i) inv_ds[i] = fma(fma(e[i], 0.375, 0.5), y[i] * e[i], y[i]);
#pragma unroll
  for (int i = 0; i < R; ++i) valid[i] = ((unsigned)__double2hiint(h2[i]) - 0x10000000u) < 0x60000000u;
}

// ---- heading relative to the reference line ------------------------------------------------------------------------
// A candidate's segment (dx, dy) almost always points nearly along the reference line: with (ux, uy) the line's unit
// tangent at the segment's first step and th = atan2(uy, ux),
//     a = dx ux + dy uy,  b = dy ux - dx uy,   atan2(dy, dx) = th + atan(b / a)      (a > 0),
// and |b / a| -- lateral over longitudinal progress -- is below 0.3 on 97 % of the steps of the BASELINE lattices.  There
// atan(q) = q + q u P1(u), u = q^2, needs a degree-6 P1 (max error 3e-15; Chebyshev fit on u in [0, 0.09],
// tools/fit_atan.py) instead of the degree-19 polynomial of the full octant, no octant logic and a plain reciprocal:
// ~30 FP64 instructions per segment instead of ~52.  th is computed once per (longitudinal row, step) and shared by all
// the lateral rows.  kAtanS[k] multiplies q^(2k+3).
__constant__ double kAtanS[7] = {-0.3333333333252849, 0.19999999811604213, -0.14285697435227054, 0.11110367880783703,
                                 -0.0907298511837223,  0.07449657993652012, -0.04900323133489978};
constexpr double kNarrow = 0.3;  // |b| <= kNarrow * a: the range of the short polynomial

// R segments in lockstep (cf. segment_fast_n).  `narrow` comes back false when some segment leaves the short
// polynomial's range (or is not an ordinary finite vector): the caller then takes segment_fast_n for the task.  The sum
// th + atan(q) is wrapped into atan2's range (-pi, pi]: the reference differences yaw WITHOUT unwrapping
// (frenet_optimal_planner.py:132), so a heading that crosses +-pi must jump exactly where numpy's arctan2 does.
template <int R>
__device__ __forceinline__ void segment_frame_n(const double (&dx)[R], const double (&dy)[R], double ux, double uy, double th,
                                                double (&yaw)[R], double (&inv_ds)[R], bool (&valid)[R], bool& narrow) {
  double a[R], b[R], r[R], e[R], q[R], u[R], w[R], pe[R], po[R], h2[R];
  narrow = true;
#pragma unroll
  for (int i = 0; i < R; ++i) {
    a[i] = fma(dx[i], ux, dy[i] * uy);
    b[i] = fma(dy[i], ux, -(dx[i] * uy));
    h2[i] = fma(dx[i], dx[i], dy[i] * dy[i]);
  }
  // q = b / a: reciprocal seed (~2^-20) + one cubic Newton step (~2^-60)
#pragma unroll
  for (int i = 0; i < R; ++i) asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r[i]) : "d"(a[i]));
#pragma unroll
  for (int i = 0; i < R; ++i) e[i] = fma(-a[i], r[i], 1.0);
#pragma unroll
  for (int i = 0; i < R; ++i) e[i] = fma(e[i], e[i], e[i]);
#pragma unroll
  for (int i = 0; i < R; ++i) r[i] = fma(r[i], e[i], r[i]);
#pragma unroll
  for (int i = 0; i < R; ++i) q[i] = b[i] * r[i];
#pragma unroll
  for (int i = 0; i < R; ++i) u[i] = q[i] * q[i];
#pragma unroll
  for (int i = 0; i < R; ++i) w[i] = u[i] * u[i];
  // P1(u) = E(w) + u O(w): two short Horner chains per segment
#pragma unroll
  for (int i = 0; i < R; ++i) {
    pe[i] = fma(kAtanS[6], w[i], kAtanS[4]);
    po[i] = fma(kAtanS[5], w[i], kAtanS[3]);
  }
#pragma unroll
  for (int i = 0; i < R; ++i) {
    pe[i] = fma(pe[i], w[i], kAtanS[2]);
    po[i] = fma(po[i], w[i], kAtanS[1]);
  }
#pragma unroll
  for (int i = 0; i < R; ++i) pe[i] = fma(pe[i], w[i], kAtanS[0]);
#pragma unroll
  for (int i = 0; i < R; ++i) pe[i] = fma(u[i], po[i], pe[i]);
#pragma unroll
  for (int i = 0; i < R; ++i) po[i] = q[i] * u[i];
  // Range tests on the high words (integer pipe; the FP64 pipe is what this stage is short of).  narrow: a is a positive
  // normal number and |q| < 0.29999995 (q is NaN when a is: not narrow).  outside: |yaw| >= 3.1415926218 -- a
  // conservative "may have left (-pi, pi]" (NaN / Inf of a lane without a segment excluded); the selects below are exact.
  bool outside = false;
#pragma unroll
  for (int i = 0; i < R; ++i) {
    yaw[i] = th + fma(po[i], pe[i], q[i]);
    narrow = narrow && __double2hiint(a[i]) >= 0x00100000 && (uint32_t)(__double2hiint(q[i]) & 0x7fffffff) < 0x3fd33333u;
    outside = outside || ((uint32_t)(__double2hiint(yaw[i]) & 0x7fffffff) - 0x400921fbu) < (0x7ff00000u - 0x400921fbu);
  }
  // th is in (-pi, pi] and |atan(q)| < 0.3: the sum leaves that range only where the road itself points (almost) along
  // -x -- one warp-uniform vote keeps the two selects per segment off the common path
  if (__any_sync(0xffffffffu, outside)) {
#pragma unroll
    for (int i = 0; i < R; ++i) {
      const double y1 = yaw[i] > 3.141592653589793 ? yaw[i] - 6.283185307179586 : yaw[i];
      yaw[i] = y1 <= -3.141592653589793 ? y1 + 6.283185307179586 : y1;
    }
  }
  // rsqrt_normal
  double y[R];
#pragma unroll
  for (int i = 0; i < R; ++i) asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y[i]) : "d"(h2[i]));
#pragma unroll
  for (int i = 0; i < R; ++i) e[i] = fma(h2[i], -(y[i] * y[i]), 1.0);
#pragma unroll
  for (int i = 0; i < R; ++i) inv_ds[i] = fma(fma(e[i], 0.375, 0.5), y[i] * e[i], y[i]);
#pragma unroll
  for (int i = 0; i < R; ++i) valid[i] = ((unsigned)__double2hiint(h2[i]) - 0x10000000u) < 0x60000000u;
}

// Out-of-line library calls for the rare lanes (kept out of the hot loop's register budget).
__device__ __noinline__ double atan2_library(double y, double x) { return atan2(y, x); }
__device__ __noinline__ double div_hypot_library(double num, double dx, double dy) { return num / hypot(dx, dy); }

}  // namespace fiss
