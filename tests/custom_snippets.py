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

# ---- user-defined costs (ilqr_problem.custom_cost = 1): the same snippet also defines ilqr_cost / ilqr_final_cost ----
# The reference's own 2-link costs (2_link_helper_functions.jl:82-108) written as a user would: p[3], p[4] = θ*.
TWO_LINK_WITH_ITS_COST = TWO_LINK + r"""
template <class T>
__device__ T ilqr_cost(const T* x, const T* u, const double* p) {
  const T e0 = p[3] - x[0], e1 = p[4] - x[1];
  const T euclidean = e0 * e0 + e1 * e1;
  const T torque = u[0] * u[0] + u[1] * u[1];
  return euclidean * 1.0 + torque * 1.0;
}
template <class T>
__device__ T ilqr_final_cost(const T* x, const double* p) {
  const T e0 = p[3] - x[0], e1 = p[4] - x[1];
  return (e0 * e0 + e1 * e1) * 1.0;
}
"""

# The tool-point cost src/cost_functions.jl intended (weighted squared distance of the tool location to a target, through
# the arm's forward kinematics) plus a u·θ̇ term that makes the cross term 𝐏 = ∂²l/∂u∂x non-zero (oracle: TwoLinkToolCost).
# p = (α, β, δ, l1, l2, target_x, target_y, w_tool, w_final, gamma).
TWO_LINK_TOOL_COST = TWO_LINK + r"""
template <class T>
__device__ void tool_point(const T* x, const double* p, T& px, T& py) {
  using namespace ilqr;
  px = p[3] * cos(x[0]) + p[4] * cos(x[0] + x[1]);
  py = p[3] * sin(x[0]) + p[4] * sin(x[0] + x[1]);
}
template <class T>
__device__ T ilqr_cost(const T* x, const T* u, const double* p) {
  T px, py;
  tool_point<T>(x, p, px, py);
  const T ex = px - p[5], ey = py - p[6];
  const T dist = ex * ex + ey * ey;
  const T torque = u[0] * u[0] + u[1] * u[1];
  const T power = u[0] * x[2] + u[1] * x[3];
  return p[7] * dist + torque + p[9] * power;
}
template <class T>
__device__ T ilqr_final_cost(const T* x, const double* p) {
  T px, py;
  tool_point<T>(x, p, px, py);
  const T ex = px - p[5], ey = py - p[6];
  return p[8] * (ex * ex + ey * ey);
}
"""
