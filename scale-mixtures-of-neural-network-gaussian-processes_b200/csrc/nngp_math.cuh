// Device math of the NNGP layer recursion (experiments/nt_kernels.py:12-31, :83-103 with neural_tangents
// semantics): diagonal maps, the branch-free ReLU arc-cosine step and the erf step, plus their partial
// derivatives (used by the gradient pass, grad.cu).
#pragma once
#include "common.cuh"

namespace smnngp {

constexpr double kPi = 3.14159265358979323846;
constexpr double kInv2Pi = 0.15915494309189533577;
constexpr double kTwoOverPi = 0.63661977236758134308;

__device__ __forceinline__ double act_diag(double u, int act) {
  if (act == ACT_RELU) return 0.5 * u;
  return kTwoOverPi * asin(2.0 * u / (1.0 + 2.0 * u));
}
// what the Gram epilogue needs per row and layer: relu -> u itself, erf -> 1/sqrt(1+2u)
__device__ __forceinline__ double encode_var(double u, int act) {
  return act == ACT_RELU ? u : 1.0 / sqrt(1.0 + 2.0 * u);
}

// ---- branch-free ReLU arc-cosine step ---------------------------------------------------------------------
// phi(k; u1, u2) = ( s + k * (pi - atan2(s, k)) ) / (2 pi),  s = sqrt(max(u1 u2 - k^2, 0))
//                = ( s + k * atan2(s, -k) ) / (2 pi)
// The library sqrt / atan2 carry special-case branches that keep the compiler from interleaving the 64
// independent evaluations a thread owns, and with 4-8 resident math warps the epilogue was latency bound
// (50 Geval/s vs 197 Geval/s of FP64-pipe capacity).  This version is straight-line code: MUFU seeds
// (rsqrt.approx / rcp.approx on the double directly) + Newton steps, one division, an 11-term minimax
// polynomial for atan on |z| <= tan(pi/8) (fitted in 60-digit arithmetic, relative error 2.2e-16).
// Accuracy of phi: <= 2.5e-16 * sqrt(u1 u2) absolute (checked against 50-digit mpmath), i.e. the same as the
// libm formulation.  Magnitudes below 1e-30 are treated as zero (inputs are O(1) standardised features).
__device__ __forceinline__ double rsqrt_seed(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  return y;
}
__device__ __forceinline__ double rcp_seed(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  return y;
}
__device__ __forceinline__ double sqrt_nobranch(double x) {          // x >= 0
  const double y = rsqrt_seed(fmax(x, 1e-30));
  double g = x * y, h = 0.5 * y;
  double r = fma(-g, h, 0.5);
  g = fma(g, r, g); h = fma(h, r, h);
  r = fma(-g, h, 0.5);
  g = fma(g, r, g); h = fma(h, r, h);
  return fma(fma(-g, g, x), h, g);
}
__device__ __forceinline__ double div_nobranch(double n, double d) {  // d >= 1e-30
  double y = rcp_seed(d);
  double e = fma(-d, y, 1.0);
  y = fma(y, e, y);
  e = fma(-d, y, 1.0);
  y = fma(y, e, y);
  const double q = n * y;
  return fma(y, fma(-q, d, n), q);
}
// returns phi; s_out = s, a_out = atan2(s, -k) = pi - theta  (d phi / d k = a / (2 pi), d phi / d u_i = s / (4 pi u_i))
__device__ __forceinline__ double phi_relu_parts(double k, double t1, double t2, double& s_out, double& a_out) {
  const double s = sqrt_nobranch(fmax(__dsub_rn(__dmul_rn(t1, t2), __dmul_rn(k, k)), 0.0));
  // a = atan2(s, -k) in [0, pi]
  const double ax = fabs(k);
  const bool swap = s > ax;
  const double num = swap ? ax : s;
  const double den = fmax(swap ? s : ax, 1e-30);
  const bool hi = num > 0.41421356237309503 * den;                   // tan(pi/8): atan(r) = pi/4 + atan((r-1)/(r+1))
  const double z = div_nobranch(hi ? num - den : num, hi ? num + den : den);
  const double w = z * z;
  double p = -0.017805397205419446;
  p = fma(p, w, 0.03796525745386593);
  p = fma(p, w, -0.05035102456601552);
  p = fma(p, w, 0.05846878297330872);
  p = fma(p, w, -0.06662951813629191);
  p = fma(p, w, 0.07692045330902225);
  p = fma(p, w, -0.09090896809064027);
  p = fma(p, w, 0.11111110744919658);
  p = fma(p, w, -0.14285714279250245);
  p = fma(p, w, 0.19999999999940893);
  p = fma(p, w, -0.3333333333333312);
  p = fma(p, w, 1.0);
  double a = fma(z, p, hi ? 0.78539816339744831 : 0.0);
  a = swap ? 1.5707963267948966 - a : a;
  a = (k > 0.0) ? kPi - a : a;                                        // x = -k < 0
  s_out = s;
  a_out = a;
  return fma(k, a, s) * kInv2Pi;
}
__device__ __forceinline__ double phi_relu_fast(double k, double t1, double t2) {
  double s, a;
  return phi_relu_parts(k, t1, t2, s, a);
}

// one nonlinearity on a cross-covariance k given the encoded marginals of its row and column
template <int ACT>
__device__ __forceinline__ double phi(double k, double t1, double t2) {
  if (ACT == ACT_RELU) {
    return phi_relu_fast(k, t1, t2);
  } else {
    double x = 2.0 * k * t1 * t2;
    x = fmin(fmax(x, -1.0), 1.0);
    return kTwoOverPi * asin(x);
  }
}

}  // namespace smnngp
