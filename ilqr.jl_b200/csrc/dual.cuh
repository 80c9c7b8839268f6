// dual.cuh — forward-mode dual numbers (value + ONE tangent) for device code: what ForwardDiff.Dual carries per
// partial (src/backward_pass.jl:32-37), one tangent direction per lane in the warp-per-trajectory kernels.
#pragma once

namespace ilqr {

// ---- value + one tangent ---------------------------------------------------------------------
struct Dual {
  double v, t;
};
__device__ __forceinline__ Dual operator+(Dual a, Dual b) { return {a.v + b.v, a.t + b.t}; }
__device__ __forceinline__ Dual operator-(Dual a, Dual b) { return {a.v - b.v, a.t - b.t}; }
__device__ __forceinline__ Dual operator-(Dual a) { return {-a.v, -a.t}; }
__device__ __forceinline__ Dual operator*(Dual a, Dual b) { return {a.v * b.v, fma(a.t, b.v, a.v * b.t)}; }
__device__ __forceinline__ Dual operator*(double s, Dual a) { return {s * a.v, s * a.t}; }
__device__ __forceinline__ Dual operator*(Dual a, double s) { return {s * a.v, s * a.t}; }
__device__ __forceinline__ Dual operator+(Dual a, double s) { return {a.v + s, a.t}; }
__device__ __forceinline__ Dual operator-(double s, Dual a) { return {s - a.v, -a.t}; }
__device__ __forceinline__ Dual operator+(double s, Dual a) { return {s + a.v, a.t}; }

template <class T> __device__ __forceinline__ T mk(double x);
template <> __device__ __forceinline__ double mk<double>(double x) { return x; }
template <> __device__ __forceinline__ Dual mk<Dual>(double x) { return {x, 0.0}; }

// the rest of the arithmetic user-supplied dynamics may need (custom_kernels.cuh)
__device__ __forceinline__ Dual operator-(Dual a, double s) { return {a.v - s, a.t}; }
__device__ __forceinline__ Dual operator/(Dual a, Dual b) { const double r = 1.0 / b.v, q = a.v * r; return {q, (a.t - q * b.t) * r}; }
__device__ __forceinline__ Dual operator/(Dual a, double s) { const double r = 1.0 / s; return {a.v * r, a.t * r}; }
__device__ __forceinline__ Dual operator/(double s, Dual b) { const double q = s / b.v; return {q, -q * b.t / b.v}; }
__device__ __forceinline__ Dual& operator+=(Dual& a, Dual b) { a = a + b; return a; }
__device__ __forceinline__ Dual& operator-=(Dual& a, Dual b) { a = a - b; return a; }
__device__ __forceinline__ Dual& operator*=(Dual& a, Dual b) { a = a * b; return a; }
__device__ __forceinline__ Dual sin(Dual a) { double s, c; sincos(a.v, &s, &c); return {s, c * a.t}; }
__device__ __forceinline__ Dual cos(Dual a) { double s, c; sincos(a.v, &s, &c); return {c, -s * a.t}; }
__device__ __forceinline__ Dual exp(Dual a) { const double e = ::exp(a.v); return {e, e * a.t}; }
__device__ __forceinline__ Dual log(Dual a) { return {::log(a.v), a.t / a.v}; }
__device__ __forceinline__ Dual sqrt(Dual a) { const double r = ::sqrt(a.v); return {r, 0.5 * a.t / r}; }

}  // namespace ilqr
