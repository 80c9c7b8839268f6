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


// ---- second order: value + two tangents + their mixed term (a hyper-dual number) ------------------------------
// f(z + ε₁ e_d + ε₂ e_e) = f + ε₁ ∂f/∂z_d + ε₂ ∂f/∂z_e + ε₁ε₂ ∂²f/∂z_d∂z_e: one evaluation of a user-supplied cost in this
// type yields one entry of its Hessian exactly — what ForwardDiff.hessian / jacobian(gradient) do to the reference's
// immediate_cost and final_cost callbacks (src/backward_pass.jl:95-99, 142-143).
struct Dual2 {
  double v, a, b, ab;
};
__device__ __forceinline__ Dual2 operator+(Dual2 x, Dual2 y) { return {x.v + y.v, x.a + y.a, x.b + y.b, x.ab + y.ab}; }
__device__ __forceinline__ Dual2 operator-(Dual2 x, Dual2 y) { return {x.v - y.v, x.a - y.a, x.b - y.b, x.ab - y.ab}; }
__device__ __forceinline__ Dual2 operator-(Dual2 x) { return {-x.v, -x.a, -x.b, -x.ab}; }
__device__ __forceinline__ Dual2 operator*(Dual2 x, Dual2 y) {
  return {x.v * y.v, fma(x.a, y.v, x.v * y.a), fma(x.b, y.v, x.v * y.b), fma(x.ab, y.v, fma(x.a, y.b, fma(x.b, y.a, x.v * y.ab)))};
}
__device__ __forceinline__ Dual2 operator*(double s, Dual2 x) { return {s * x.v, s * x.a, s * x.b, s * x.ab}; }
__device__ __forceinline__ Dual2 operator*(Dual2 x, double s) { return s * x; }
__device__ __forceinline__ Dual2 operator+(Dual2 x, double s) { return {x.v + s, x.a, x.b, x.ab}; }
__device__ __forceinline__ Dual2 operator+(double s, Dual2 x) { return x + s; }
__device__ __forceinline__ Dual2 operator-(Dual2 x, double s) { return {x.v - s, x.a, x.b, x.ab}; }
__device__ __forceinline__ Dual2 operator-(double s, Dual2 x) { return {s - x.v, -x.a, -x.b, -x.ab}; }
// g(x) for a scalar function with derivatives g1 = g'(x.v), g2 = g''(x.v)
__device__ __forceinline__ Dual2 chain2(Dual2 x, double g0, double g1, double g2) {
  return {g0, g1 * x.a, g1 * x.b, fma(g2, x.a * x.b, g1 * x.ab)};
}
__device__ __forceinline__ Dual2 operator/(Dual2 x, Dual2 y) {
  const double r = 1.0 / y.v;
  return x * chain2(y, r, -r * r, 2.0 * r * r * r);
}
__device__ __forceinline__ Dual2 operator/(Dual2 x, double s) { return (1.0 / s) * x; }
__device__ __forceinline__ Dual2 operator/(double s, Dual2 y) { const double r = 1.0 / y.v; return s * chain2(y, r, -r * r, 2.0 * r * r * r); }
__device__ __forceinline__ Dual2& operator+=(Dual2& x, Dual2 y) { x = x + y; return x; }
__device__ __forceinline__ Dual2& operator-=(Dual2& x, Dual2 y) { x = x - y; return x; }
__device__ __forceinline__ Dual2& operator*=(Dual2& x, Dual2 y) { x = x * y; return x; }
__device__ __forceinline__ Dual2 sin(Dual2 x) { double s, c; sincos(x.v, &s, &c); return chain2(x, s, c, -s); }
__device__ __forceinline__ Dual2 cos(Dual2 x) { double s, c; sincos(x.v, &s, &c); return chain2(x, c, -s, -c); }
__device__ __forceinline__ Dual2 exp(Dual2 x) { const double e = ::exp(x.v); return chain2(x, e, e, e); }
__device__ __forceinline__ Dual2 log(Dual2 x) { const double r = 1.0 / x.v; return chain2(x, ::log(x.v), r, -r * r); }
__device__ __forceinline__ Dual2 sqrt(Dual2 x) { const double q = ::sqrt(x.v), h = 0.5 / q; return chain2(x, q, h, -h / (2.0 * x.v)); }
template <> __device__ __forceinline__ Dual2 mk<Dual2>(double x) { return {x, 0.0, 0.0, 0.0}; }

}  // namespace ilqr
