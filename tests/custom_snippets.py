"""CUDA C++ dynamics snippets for ILQR_MODEL_CUSTOM used by the tests (what a user of the library would write)."""

# The reference's 2-link plugin (test/2_link_example/2_link_helper_functions.jl:29-69) in closed form:
# M = [α+2β c₂, δ+β c₂; δ+β c₂, δ],  C = −β s₂ θ̇₂ [1 ½; ½ 0] (the single-index sum of :43),  θ̈ = M⁻¹(u − C θ̇).
# p = (α, β, δ).
TWO_LINK = r"""
template <class T>
__device__ void ilqr_dynamics(const T* x, const T* u, const double* p, T* xdot) {
  using namespace ilqr;
  const double alpha = p[0], beta = p[1], delta = p[2];
  const T s2 = sin(x[1]), c2 = cos(x[1]);
  const T a = alpha + 2.0 * beta * c2, b = delta + beta * c2;
  const T tw2 = (-beta) * s2 * x[3];
  const T r1 = u[0] - tw2 * (x[2] + 0.5 * x[3]);
  const T r2 = u[1] - tw2 * (0.5 * x[2]);
  const T det = a * delta - b * b;
  xdot[0] = x[2];
  xdot[1] = x[3];
  xdot[2] = (delta * r1 - b * r2) / det;
  xdot[3] = (a * r2 - b * r1) / det;
}
"""

# damped pendulum on a cart-less pivot with a nonlinear spring: exercises exp / sqrt / division
PENDULUM = r"""
template <class T>
__device__ void ilqr_dynamics(const T* x, const T* u, const double* p, T* xdot) {
  using namespace ilqr;
  const double g = p[0], l = p[1], c = p[2];
  xdot[0] = x[1];
  xdot[1] = (u[0] - c * x[1] * sqrt(1.0 + x[1] * x[1]) - g / l * sin(x[0])) / (1.0 + 0.1 * exp(-x[0] * x[0]));
}
"""

BROKEN = r"""
template <class T>
__device__ void ilqr_dynamics(const T* x, const T* u, const double* p, T* xdot) { xdot[0] = undefined_symbol; }
"""
