// ilqr_oracle.hpp — CPU ORACLE (test infrastructure, NOT product code).
//
// A line-by-line C++17 fp64 restatement of the reference algorithm
// (aabouman/iLQR.jl).  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may use anything under oracle/.
// The product path (libilqr_b200.so) never links, calls or falls back to it.
//
// PARITY PINNING.  The reference ships no golden vectors and Julia is not available in this image, but it does ship
// OUTPUTS OF ITS OWN RUNS: five animations (test/2_link_example/figures/iLQR_2_link*.gif, docs/extras) that
// animate_2_link.jl:27-41 drew from `iLQR.fit`'s result — every 10th knot point of an H = 900 solve, one target tool
// location per quadrant.  tests/golden/make_gif_angles.py reads the joint angles back from the frames
// (tests/golden/reference_gif_angles.json, ±0.01 rad: a pixel is 0.011 units) and tests/test_reference_gif_cpu.py holds
// this oracle to them: rms 0.002-0.006 rad, max 0.012-0.025 rad over 91 frames x 2 angles x 5 animations, while the same
// solver with textbook Coriolis terms instead of the reference's single-index sum misses by 0.016 / 0.053.  That pins
// the dynamics, costs, target kinematics, horizon bookkeeping and the solver's fixed point against the real package —
// to pixel accuracy.  Below that ("parity unpinned" in the strict sense: no NUMBER printed by the Julia package
// exists) the oracle rests on (i) an independently written closed-form NumPy restatement (tests/np_restatement.py,
// 1e-12), (ii) central-difference derivative checks, (iii) the discrete-LQR known-answer test and (iv) the cost traces
// recorded in SURVEY.md §6 (which came from a third restatement).
//
// Reference lines followed (paths relative to /root/reference):
//   src/backward_pass.jl:25-40    linearize_dynamics
//   src/backward_pass.jl:81-109   immediate_cost_quadratization
//   src/backward_pass.jl:134-153  final_cost_quadratization
//   src/backward_pass.jl:177-186  optimal_controller_param
//   src/backward_pass.jl:207-218  feedback_parameters
//   src/backward_pass.jl:262-273  step_back
//   src/backward_pass.jl:324-357  backward_pass
//   src/forward_pass.jl:55-93     forward_pass
//   src/forward_pass.jl:148-179   fit
//   src/forward_pass.jl:182-196   total_cost_generator
//   test/2_link_example/2_link_helper_functions.jl:4-108  the 2-link plugin
// Third-party arithmetic restated: ForwardDiff.jl 0.10.14 (forward-mode dual
// numbers; docs/Manifest.toml is the only pin) and LinearAlgebra `\` / `inv`
// (partial-pivot LU).
#pragma once
#include <array>
#include <cmath>
#include <cstdint>
#include <limits>
#include <vector>

namespace oracle {

// ---------------------------------------------------------------------------
// Forward-mode dual numbers (what ForwardDiff.Dual{Tag,T,N} computes).
// ---------------------------------------------------------------------------
template <class T, int N>
struct Dual {
  T v{};
  std::array<T, N> d{};
  Dual() = default;
  Dual(const T& value) : v(value) {
    for (auto& e : d) e = T(0.0);
  }
  template <class S, class = std::enable_if_t<std::is_arithmetic<S>::value && !std::is_same<S, T>::value>>
  Dual(S value) : v(T(value)) {
    for (auto& e : d) e = T(0.0);
  }
};

template <class T> struct is_dual : std::false_type {};
template <class T, int N> struct is_dual<Dual<T, N>> : std::true_type {};

inline double value_of(double x) { return x; }
template <class T, int N> double value_of(const Dual<T, N>& x) { return value_of(x.v); }

template <class T, int N> Dual<T, N> operator+(const Dual<T, N>& a, const Dual<T, N>& b) {
  Dual<T, N> r; r.v = a.v + b.v;
  for (int i = 0; i < N; ++i) r.d[i] = a.d[i] + b.d[i];
  return r;
}
template <class T, int N> Dual<T, N> operator-(const Dual<T, N>& a, const Dual<T, N>& b) {
  Dual<T, N> r; r.v = a.v - b.v;
  for (int i = 0; i < N; ++i) r.d[i] = a.d[i] - b.d[i];
  return r;
}
template <class T, int N> Dual<T, N> operator-(const Dual<T, N>& a) {
  Dual<T, N> r; r.v = -a.v;
  for (int i = 0; i < N; ++i) r.d[i] = -a.d[i];
  return r;
}
template <class T, int N> Dual<T, N> operator*(const Dual<T, N>& a, const Dual<T, N>& b) {
  Dual<T, N> r; r.v = a.v * b.v;
  for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * b.v + a.v * b.d[i];
  return r;
}
template <class T, int N> Dual<T, N> operator/(const Dual<T, N>& a, const Dual<T, N>& b) {
  Dual<T, N> r; r.v = a.v / b.v;
  // d(a/b) = (da - (a/b) db) / b
  for (int i = 0; i < N; ++i) r.d[i] = (a.d[i] - r.v * b.d[i]) / b.v;
  return r;
}
// mixed with plain double (constants)
template <class T, int N> Dual<T, N> operator*(double s, const Dual<T, N>& a) {
  Dual<T, N> r; r.v = s * a.v;
  for (int i = 0; i < N; ++i) r.d[i] = s * a.d[i];
  return r;
}
template <class T, int N> Dual<T, N> operator*(const Dual<T, N>& a, double s) { return s * a; }
template <class T, int N> Dual<T, N> operator/(const Dual<T, N>& a, double s) {
  Dual<T, N> r; r.v = a.v / s;
  for (int i = 0; i < N; ++i) r.d[i] = a.d[i] / s;
  return r;
}
template <class T, int N> Dual<T, N> operator+(double s, const Dual<T, N>& a) {
  Dual<T, N> r = a; r.v = s + a.v; return r;
}
template <class T, int N> Dual<T, N> operator+(const Dual<T, N>& a, double s) {
  Dual<T, N> r = a; r.v = a.v + s; return r;
}
template <class T, int N> Dual<T, N> operator-(double s, const Dual<T, N>& a) {
  Dual<T, N> r = -a; r.v = s - a.v; return r;
}
template <class T, int N> Dual<T, N> operator-(const Dual<T, N>& a, double s) {
  Dual<T, N> r = a; r.v = a.v - s; return r;
}
template <class T, int N> Dual<T, N>& operator+=(Dual<T, N>& a, const Dual<T, N>& b) { a = a + b; return a; }

using std::cos;
using std::sin;
template <class T, int N> Dual<T, N> sin(const Dual<T, N>& a) {
  Dual<T, N> r; r.v = sin(a.v); T c = cos(a.v);
  for (int i = 0; i < N; ++i) r.d[i] = c * a.d[i];
  return r;
}
template <class T, int N> Dual<T, N> cos(const Dual<T, N>& a) {
  Dual<T, N> r; r.v = cos(a.v); T ms = -sin(a.v);
  for (int i = 0; i < N; ++i) r.d[i] = ms * a.d[i];
  return r;
}

// ---------------------------------------------------------------------------
// Small dense column-major matrices (Julia Array semantics, no heap).
// ---------------------------------------------------------------------------
template <class T, int R, int C>
struct Mat {
  std::array<T, R * C> a{};
  T& operator()(int i, int j) { return a[i + R * j]; }
  const T& operator()(int i, int j) const { return a[i + R * j]; }
  T& operator[](int i) { return a[i]; }
  const T& operator[](int i) const { return a[i]; }
  static Mat zeros() { Mat m; for (auto& e : m.a) e = T(0.0); return m; }
};
template <class T, int R> using Vec = Mat<T, R, 1>;

template <class T, int R, int C> Mat<T, R, C> operator+(const Mat<T, R, C>& x, const Mat<T, R, C>& y) {
  Mat<T, R, C> r; for (int i = 0; i < R * C; ++i) r.a[i] = x.a[i] + y.a[i]; return r;
}
template <class T, int R, int C> Mat<T, R, C> operator-(const Mat<T, R, C>& x, const Mat<T, R, C>& y) {
  Mat<T, R, C> r; for (int i = 0; i < R * C; ++i) r.a[i] = x.a[i] - y.a[i]; return r;
}
template <class T, int R, int C> Mat<T, R, C> operator-(const Mat<T, R, C>& x) {
  Mat<T, R, C> r; for (int i = 0; i < R * C; ++i) r.a[i] = -x.a[i]; return r;
}
template <class T, int R, int C> Mat<T, R, C> operator*(double s, const Mat<T, R, C>& x) {
  Mat<T, R, C> r; for (int i = 0; i < R * C; ++i) r.a[i] = s * x.a[i]; return r;
}
template <class T, int R, int C> Mat<T, R, C> operator/(const Mat<T, R, C>& x, double s) {
  Mat<T, R, C> r; for (int i = 0; i < R * C; ++i) r.a[i] = x.a[i] / s; return r;
}
// generic matmul, accumulation order k = 0..K-1 starting from the first product
template <class T, int R, int K, int C> Mat<T, R, C> operator*(const Mat<T, R, K>& x, const Mat<T, K, C>& y) {
  Mat<T, R, C> r;
  for (int j = 0; j < C; ++j)
    for (int i = 0; i < R; ++i) {
      T acc = x(i, 0) * y(0, j);
      for (int k = 1; k < K; ++k) acc = acc + x(i, k) * y(k, j);
      r(i, j) = acc;
    }
  return r;
}
template <class T, int R, int C> Mat<T, C, R> transpose(const Mat<T, R, C>& x) {
  Mat<T, C, R> r; for (int i = 0; i < R; ++i) for (int j = 0; j < C; ++j) r(j, i) = x(i, j); return r;
}
template <class T, int R, int C> bool any_nan(const Mat<T, R, C>& x) {
  for (int i = 0; i < R * C; ++i) if (std::isnan(value_of(x.a[i]))) return true;
  return false;
}

// `A \ B` for a dense square non-triangular A: partial-pivot LU (getrf/getrs).
template <class T, int N, int M>
Mat<T, N, M> lu_solve(Mat<T, N, N> A, Mat<T, N, M> B) {
  for (int k = 0; k < N; ++k) {
    int p = k; double best = std::fabs(value_of(A(k, k)));
    for (int i = k + 1; i < N; ++i) {
      double c = std::fabs(value_of(A(i, k)));
      if (c > best) { best = c; p = i; }
    }
    if (p != k) {
      for (int j = 0; j < N; ++j) std::swap(A(k, j), A(p, j));
      for (int j = 0; j < M; ++j) std::swap(B(k, j), B(p, j));
    }
    for (int i = k + 1; i < N; ++i) {
      T l = A(i, k) / A(k, k);
      A(i, k) = l;
      for (int j = k + 1; j < N; ++j) A(i, j) = A(i, j) - l * A(k, j);
      for (int j = 0; j < M; ++j) B(i, j) = B(i, j) - l * B(k, j);
    }
  }
  for (int j = 0; j < M; ++j)
    for (int i = N - 1; i >= 0; --i) {
      T acc = B(i, j);
      for (int k = i + 1; k < N; ++k) acc = acc - A(i, k) * B(k, j);
      B(i, j) = acc / A(i, i);
    }
  return B;
}
template <class T, int N> Mat<T, N, N> inv(const Mat<T, N, N>& A) {
  Mat<T, N, N> I = Mat<T, N, N>::zeros();
  for (int i = 0; i < N; ++i) I(i, i) = T(1.0);
  return lu_solve<T, N, N>(A, I);
}

// ---------------------------------------------------------------------------
// ForwardDiff.{jacobian,gradient,hessian} restated for fixed sizes.
// F is a generic callable taking Vec<S,N> and returning Vec<S,R> (or S).
// ---------------------------------------------------------------------------
template <class T, int N> Vec<Dual<T, N>, N> seed(const Vec<T, N>& x) {
  Vec<Dual<T, N>, N> xd;
  for (int i = 0; i < N; ++i) { xd[i] = Dual<T, N>(x[i]); xd[i].d[i] = T(1.0); }
  return xd;
}
template <class T, int N, int M> Vec<Dual<T, M>, N> lift(const Vec<T, N>& x) {
  Vec<Dual<T, M>, N> xd;
  for (int i = 0; i < N; ++i) xd[i] = Dual<T, M>(x[i]);
  return xd;
}
template <int R, class T, int N, class F> Mat<T, R, N> jacobian(F&& f, const Vec<T, N>& x) {
  auto y = f(seed<T, N>(x));
  Mat<T, R, N> J;
  for (int i = 0; i < R; ++i) for (int j = 0; j < N; ++j) J(i, j) = y[i].d[j];
  return J;
}
template <class T, int N, class F> Vec<T, N> gradient(F&& f, const Vec<T, N>& x) {
  auto y = f(seed<T, N>(x));
  Vec<T, N> g; for (int j = 0; j < N; ++j) g[j] = y.d[j];
  return g;
}
template <class T, int N, class F> Mat<T, N, N> hessian(F&& f, const Vec<T, N>& x) {
  // hessian(f, x) = jacobian(y -> gradient(f, y), x)
  return jacobian<N>([&](const auto& y) {
    using S = std::decay_t<decltype(y[0])>;
    return gradient<S, N>(f, y);
  }, x);
}

// ---------------------------------------------------------------------------
// The 2-link plugin: test/2_link_example/2_link_helper_functions.jl
// ---------------------------------------------------------------------------
struct TwoLink {
  static constexpr int NX = 4, NU = 2;
  double l1, l2, r1, r2, m1, m2, Iz1, Iz2, alpha, beta, delta, dt;
  double target_tool_loc[2];
  double target_joint[2];

  TwoLink() {
    // :4-16
    l1 = std::sqrt(2.) / 2.; l2 = std::sqrt(2.) / 2.;
    r1 = 0.5 * l1; r2 = 0.5 * l2;
    m1 = 1.0; m2 = 1.0;
    Iz1 = 1.0 / 12.0 * m1 * (l1 * l1); Iz2 = 1.0 / 12.0 * m2 * (l2 * l2);
    alpha = Iz1 + Iz2 + m1 * (r1 * r1) + m2 * (l1 * l1 + r2 * r2);
    beta = m2 * l1 * r2;
    delta = Iz2 + m2 * (r2 * r2);
    dt = 0.01;
    target_tool_loc[0] = 0.6; target_tool_loc[1] = -0.5;
    inverse_kinematics(target_tool_loc, target_joint);
  }
  // :19-26 (only ever called on the Float64 constant target)
  void inverse_kinematics(const double w[2], double q[2]) const {
    double x = w[0], y = w[1];
    double q2 = std::acos((x * x + y * y - l1 * l1 - l2 * l2) / (2 * l1 * l2));
    double q1 = std::atan2(y, x) - std::atan2(l2 * std::sin(q2), l1 + l2 * std::cos(q2));
    q[0] = q1; q[1] = q2;
  }
  // :29-33
  template <class T> Mat<T, 2, 2> inertia_matrix(const Vec<T, 2>& th) const {
    Mat<T, 2, 2> M;
    M(0, 0) = alpha + 2 * beta * cos(th[1]); M(0, 1) = delta + beta * cos(th[1]);
    M(1, 0) = delta + beta * cos(th[1]);     M(1, 1) = T(delta);
    return M;
  }
  // :36-47 — including the `for k in length(θ)` single-index sum (k = 2 only)
  template <class T> Mat<T, 2, 2> coriolis_matrix(const Vec<T, 2>& th, const Vec<T, 2>& thd) const {
    // ∇M = reshape(jacobian(InertiaMatrix, θ), (2,2,2)): ∇M[a,b,c] = ∂M[a,b]/∂θ_c
    Mat<T, 4, 2> J = jacobian<4>([&](const auto& y) {
      using S = std::decay_t<decltype(y[0])>;
      Mat<S, 2, 2> M = this->template inertia_matrix<S>(y);
      Vec<S, 4> flat; for (int i = 0; i < 4; ++i) flat[i] = M.a[i];
      return flat;
    }, th);
    auto dM = [&](int a, int b, int c) -> const T& { return J(a + 2 * b, c); };
    Mat<T, 2, 2> C = Mat<T, 2, 2>::zeros();
    const int k = 1;  // Julia: k in length(θ) → k = 2 (1-based)
    for (int i = 0; i < 2; ++i)
      for (int j = 0; j < 2; ++j)
        C(i, j) = (1.0 / 2.0) * (dM(k, i, j) + dM(j, i, k) - dM(i, k, j)) * thd[k];
    return C;
  }
  // :51-69
  template <class T> Vec<T, 4> continuous_dynamics(const Vec<T, 4>& state, const Vec<T, 2>& w) const {
    Vec<T, 2> th, thd; th[0] = state[0]; th[1] = state[1]; thd[0] = state[2]; thd[1] = state[3];
    Mat<T, 2, 2> M = inertia_matrix<T>(th);
    Mat<T, 2, 2> C = coriolis_matrix<T>(th, thd);
    Mat<T, 2, 2> MC = lu_solve<T, 2, 2>(-M, C);   // -M_mat\C_mat parses as (-M_mat)\C_mat
    Mat<T, 2, 2> Mi = inv<T, 2>(M);
    Mat<T, 4, 4> mat1 = Mat<T, 4, 4>::zeros();
    mat1(0, 2) = T(1.0); mat1(1, 3) = T(1.0);
    for (int i = 0; i < 2; ++i) for (int j = 0; j < 2; ++j) mat1(2 + i, 2 + j) = MC(i, j);
    Mat<T, 4, 2> mat2 = Mat<T, 4, 2>::zeros();
    for (int i = 0; i < 2; ++i) for (int j = 0; j < 2; ++j) mat2(2 + i, j) = Mi(i, j);
    return mat1 * state + mat2 * w;
  }
  // :49-79 RK4
  template <class T> Vec<T, 4> dynamicsf(const Vec<T, 4>& x, const Vec<T, 2>& u) const {
    Vec<T, 4> k1 = dt * continuous_dynamics<T>(x, u);
    Vec<T, 4> k2 = dt * continuous_dynamics<T>(x + k1 / 2.0, u);
    Vec<T, 4> k3 = dt * continuous_dynamics<T>(x + k2 / 2.0, u);
    Vec<T, 4> k4 = dt * continuous_dynamics<T>(x + k3, u);
    return x + (1.0 / 6.0) * (k1 + 2.0 * k2 + 2.0 * k3 + k4);
  }
  // :82-97 (velocity_penalty is computed in the reference but never added)
  template <class T> T immediate_cost(const Vec<T, 4>& x, const Vec<T, 2>& u) const {
    T e0 = target_joint[0] - x[0], e1 = target_joint[1] - x[1];
    T euclidean = e0 * e0 + e1 * e1;
    T torque = u[0] * u[0] + u[1] * u[1];
    return euclidean * 1.0 + torque * 1.0;
  }
  // :100-108
  template <class T> T final_cost(const Vec<T, 4>& x) const {
    T e0 = target_joint[0] - x[0], e1 = target_joint[1] - x[1];
    T euclidean = e0 * e0 + e1 * e1;
    return euclidean * 1.0;
  }
};

// The 2-link plugin with the costs src/cost_functions.jl intended (that file is dead code in the reference: its include is
// commented out at src/iLQR.jl:9, and as written it needs RigidBodyDynamics' MechanismState).  final_cost =
// weight · ‖tool(θ) − target‖² — "the weighted euclidean distance from the specified tool location to a target location"
// (src/cost_functions.jl:1-27), with the planar forward kinematics of the 2-link arm
// (2_link_helper_functions.jl:19-26 is its inverse); immediate_cost = w_tool · the same distance + Σu² (:30-54 has Σu²)
// + gamma · u·θ̇ (mechanical power), which is NOT in the reference: it is there so that the cross term 𝐏 = ∂²l/∂u∂x of
// immediate_cost_quadratization (src/backward_pass.jl:98) is exercised — every cost the reference ships has 𝐏 = 0.
// The solver above differentiates these with the same nested dual numbers as any other callback.
struct TwoLinkToolCost : TwoLink {
  double w_tool = 1.0, w_final = 50.0, gamma = 0.3;
  template <class T> void tool(const Vec<T, 4>& x, T& px, T& py) const {
    px = l1 * cos(x[0]) + l2 * cos(x[0] + x[1]);
    py = l1 * sin(x[0]) + l2 * sin(x[0] + x[1]);
  }
  template <class T> T immediate_cost(const Vec<T, 4>& x, const Vec<T, 2>& u) const {
    T px, py; tool<T>(x, px, py);
    T ex = px - target_tool_loc[0], ey = py - target_tool_loc[1];
    T dist = ex * ex + ey * ey;
    T torque = u[0] * u[0] + u[1] * u[1];
    T power = u[0] * x[2] + u[1] * x[3];
    return w_tool * dist + torque + gamma * power;
  }
  template <class T> T final_cost(const Vec<T, 4>& x) const {
    T px, py; tool<T>(x, px, py);
    T ex = px - target_tool_loc[0], ey = py - target_tool_loc[1];
    return w_final * (ex * ex + ey * ey);
  }
};

// A linear-dynamics / quadratic-cost plugin used only for the known-answer
// (discrete LQR) test of the solver core.  Not in the reference.
template <int NXv, int NUv>
struct LinearQuadratic {
  static constexpr int NX = NXv, NU = NUv;
  Mat<double, NXv, NXv> A; Mat<double, NXv, NUv> B;
  Mat<double, NXv, NXv> Q; Mat<double, NUv, NUv> R; Mat<double, NXv, NXv> Qf;
  template <class T> Vec<T, NXv> dynamicsf(const Vec<T, NXv>& x, const Vec<T, NUv>& u) const {
    Vec<T, NXv> y;
    for (int i = 0; i < NXv; ++i) {
      T acc = T(0.0);
      for (int j = 0; j < NXv; ++j) acc = acc + A(i, j) * x[j];
      for (int j = 0; j < NUv; ++j) acc = acc + B(i, j) * u[j];
      y[i] = acc;
    }
    return y;
  }
  template <class T> T immediate_cost(const Vec<T, NXv>& x, const Vec<T, NUv>& u) const {
    T acc = T(0.0);
    for (int i = 0; i < NXv; ++i) for (int j = 0; j < NXv; ++j) acc = acc + 0.5 * Q(i, j) * (x[i] * x[j]);
    for (int i = 0; i < NUv; ++i) for (int j = 0; j < NUv; ++j) acc = acc + 0.5 * R(i, j) * (u[i] * u[j]);
    return acc;
  }
  template <class T> T final_cost(const Vec<T, NXv>& x) const {
    T acc = T(0.0);
    for (int i = 0; i < NXv; ++i) for (int j = 0; j < NXv; ++j) acc = acc + 0.5 * Qf(i, j) * (x[i] * x[j]);
    return acc;
  }
};

// ---------------------------------------------------------------------------
// Solver core.  Trajectories use the Julia column-major layout:
//   x[N×n]: element (k,c) at k + N*c;  u[H×m]: (k,i) at k + H*i;
//   K[H×m×n]: (k,i,j) at k + H*(i + m*j);  δuff[H×m]: (k,i) at k + H*i.
// ---------------------------------------------------------------------------
enum Status : int32_t {
  OK = 0,
  NAN_GAINS = 1,      // src/backward_pass.jl:353-354 assert
  NAN_ROLLOUT = 2,    // src/forward_pass.jl:89-90 assert
  LS_EXHAUSTED = 4,   // bound J_max on the reference's unbounded `while true` (:70)
  NOT_DECREASED = 8,  // src/forward_pass.jl:168 assert
};

template <class P> struct Solver {
  static constexpr int n = P::NX, m = P::NU;
  using VX = Vec<double, n>; using VU = Vec<double, m>;
  using MXX = Mat<double, n, n>; using MXU = Mat<double, n, m>;
  using MUX = Mat<double, m, n>; using MUU = Mat<double, m, m>;
  const P& p;
  double reg = 0.01;   // src/backward_pass.jl:214
  int jmax = 32;       // line-search bound (reference: unbounded)
  explicit Solver(const P& plugin) : p(plugin) {}

  static VX row_x(const double* x, int N, int k) { VX v; for (int c = 0; c < n; ++c) v[c] = x[k + N * c]; return v; }
  static VU row_u(const double* u, int H, int k) { VU v; for (int c = 0; c < m; ++c) v[c] = u[k + H * c]; return v; }

  // src/backward_pass.jl:25-40
  void linearize_dynamics(const VX& x, const VU& u, MXX& A, MXU& B) const {
    A = jacobian<n>([&](const auto& xd) { return p.dynamicsf(xd, lift<double, m, n>(u)); }, x);
    B = jacobian<n>([&](const auto& ud) { return p.dynamicsf(lift<double, n, m>(x), ud); }, u);
  }
  // src/backward_pass.jl:81-109
  void immediate_cost_quadratization(const VX& x, const VU& u, double& q, VX& qv, VU& rv, MXX& Q, MUX& Pm, MUU& R) const {
    q = p.immediate_cost(x, u);
    qv = gradient<double, n>([&](const auto& xd) {
      using S = std::decay_t<decltype(xd[0])>;
      Vec<S, m> uu; for (int i = 0; i < m; ++i) uu[i] = S(u[i]);
      return p.immediate_cost(xd, uu); }, x);
    rv = gradient<double, m>([&](const auto& ud) {
      using S = std::decay_t<decltype(ud[0])>;
      Vec<S, n> xx; for (int i = 0; i < n; ++i) xx[i] = S(x[i]);
      return p.immediate_cost(xx, ud); }, u);
    Q = hessian<double, n>([&](const auto& xd) {
      using S = std::decay_t<decltype(xd[0])>;
      Vec<S, m> uu; for (int i = 0; i < m; ++i) uu[i] = S(u[i]);
      return p.immediate_cost(xd, uu); }, x);
    // ∂²L∂u∂x = jacobian(x -> gradient(u -> l(x,u), u), x)   [m×n]
    Pm = jacobian<m>([&](const auto& xd) {
      using SX = std::decay_t<decltype(xd[0])>;
      Vec<SX, m> u0; for (int i = 0; i < m; ++i) u0[i] = SX(u[i]);
      return gradient<SX, m>([&](const auto& ud) {
        using SU = std::decay_t<decltype(ud[0])>;
        Vec<SU, n> xx; for (int i = 0; i < n; ++i) xx[i] = SU(xd[i]);
        return p.immediate_cost(xx, ud); }, u0);
    }, x);
    R = hessian<double, m>([&](const auto& ud) {
      using S = std::decay_t<decltype(ud[0])>;
      Vec<S, n> xx; for (int i = 0; i < n; ++i) xx[i] = S(x[i]);
      return p.immediate_cost(xx, ud); }, u);
  }
  // src/backward_pass.jl:134-153
  void final_cost_quadratization(const VX& x, double& q, VX& qv, MXX& Q) const {
    q = p.final_cost(x);
    qv = gradient<double, n>([&](const auto& xd) { return p.final_cost(xd); }, x);
    Q = hessian<double, n>([&](const auto& xd) { return p.final_cost(xd); }, x);
  }
  // src/backward_pass.jl:177-186 — Bᵀ*S*A is left-associated: (BᵀS)A
  static void optimal_controller_param(const MXX& A, const MXU& B, const VU& rv, const MUX& Pm, const MUU& R,
                                       const VX& sv, const MXX& S, VU& g, MUX& G, MUU& Hm) {
    auto Bt = transpose(B);
    g = rv + Bt * sv;
    G = Pm + (Bt * S) * A;
    Hm = R + (Bt * S) * B;
  }
  // src/backward_pass.jl:207-218
  void feedback_parameters(const VU& g, const MUX& G, const MUU& Hm, VU& du, MUX& K) const {
    MUU Hreg = Hm;
    for (int i = 0; i < m; ++i) Hreg(i, i) = Hm(i, i) + reg * 1.0;
    du = lu_solve<double, m, 1>(-Hreg, g);   // - H_reg \ g  ≡ (-H_reg) \ g
    K = lu_solve<double, m, n>(-Hreg, G);
  }
  // src/backward_pass.jl:262-273 — unregularised H; S not symmetrised
  static void step_back(const MXX& A, double q, const VX& qv, const MXX& Q, const VU& g, const MUX& G, const MUU& Hm,
                        const VU& du, const MUX& K, double& s, VX& sv, MXX& S) {
    auto Kt = transpose(K); auto At = transpose(A); auto Gt = transpose(G);
    auto dut = transpose(du);
    double s_new = q + s + ((0.5 * dut) * Hm * du)[0] + (dut * g)[0];
    VX sv_new = qv + At * sv + (Kt * Hm) * du + Kt * g + Gt * du;
    MXX S_new = Q + (At * S) * A + (Kt * Hm) * K + Kt * G + Gt * K;
    s = s_new; sv = sv_new; S = S_new;
  }
  // src/backward_pass.jl:324-357.  Returns status bits.
  int32_t backward_pass(int H, const double* x, const double* u, double* duff, double* Ks) const {
    const int N = H + 1;
    double s; VX sv; MXX S;
    final_cost_quadratization(row_x(x, N, N - 1), s, sv, S);
    bool nan = false;
    for (int i = N - 2; i >= 0; --i) {
      VX xi = row_x(x, N, i); VU ui = row_u(u, H, i);
      MXX A; MXU B; linearize_dynamics(xi, ui, A, B);
      double q; VX qv; VU rv; MXX Q; MUX Pm; MUU R;
      immediate_cost_quadratization(xi, ui, q, qv, rv, Q, Pm, R);
      VU g; MUX G; MUU Hm; optimal_controller_param(A, B, rv, Pm, R, sv, S, g, G, Hm);
      VU du; MUX K; feedback_parameters(g, G, Hm, du, K);
      for (int a = 0; a < m; ++a) {
        duff[i + H * a] = du[a];
        for (int b = 0; b < n; ++b) Ks[i + H * (a + m * b)] = K(a, b);
      }
      nan = nan || any_nan(du) || any_nan(K);
      step_back(A, q, qv, Q, g, G, Hm, du, K, s, sv, S);
    }
    return nan ? NAN_GAINS : OK;
  }
  // src/forward_pass.jl:182-196.  x_traj may be null (= zeros).
  double total_cost(int H, const double* xb, const double* ub, const double* x_traj) const {
    const int N = H + 1;
    double sum = 0.;
    for (int i = 0; i < H; ++i) {
      VX xi = row_x(xb, N, i);
      if (x_traj) xi = xi - row_x(x_traj, N, i);
      sum += p.immediate_cost(xi, row_u(ub, H, i));
    }
    sum += p.final_cost(row_x(xb, N, N - 1));
    return sum;
  }
  // One candidate of src/forward_pass.jl:71-76
  double rollout(int H, const double* x, const double* u, const double* x_traj, const double* duff, const double* Ks,
                 double alpha, double* xb, double* ub) const {
    const int N = H + 1;
    for (int c = 0; c < n; ++c) xb[0 + N * c] = x[0 + N * c];
    for (int k = 0; k < H; ++k) {
      VX xk = row_x(xb, N, k);
      VX dx = xk - row_x(x, N, k);
      VU uk;
      for (int a = 0; a < m; ++a) {
        // u[k,:] + α*δuff[k,:] + K[k,:,:]*δx   (left to right)
        double Kdx = Ks[k + H * (a + m * 0)] * dx[0];
        for (int b = 1; b < n; ++b) Kdx = Kdx + Ks[k + H * (a + m * b)] * dx[b];
        uk[a] = (u[k + H * a] + alpha * duff[k + H * a]) + Kdx;
        ub[k + H * a] = uk[a];
      }
      VX xn = p.dynamicsf(xk, uk);
      for (int c = 0; c < n; ++c) xb[(k + 1) + N * c] = xn[c];
    }
    return total_cost(H, xb, ub, x_traj);
  }
  // src/forward_pass.jl:55-93 with the line search bounded at jmax candidates
  // α = 2^-j, j = 0..jmax-1.  Accept the smallest j with prev_cost - J_j > 0.
  int32_t forward_pass(int H, const double* x, const double* u, const double* x_traj, const double* duff,
                       const double* Ks, double prev_cost, double* xb, double* ub, double* new_cost,
                       double* alpha_out) const {
    const int N = H + 1;
    double alpha = 1.0;
    for (int j = 0; j < jmax; ++j) {
      double c = rollout(H, x, u, x_traj, duff, Ks, alpha, xb, ub);
      double dcost = prev_cost - c;
      if (dcost > 0) {
        *new_cost = c; *alpha_out = alpha;
        bool nan = false;
        for (int i = 0; i < N * n; ++i) nan = nan || std::isnan(xb[i]);
        for (int i = 0; i < H * m; ++i) nan = nan || std::isnan(ub[i]);
        return nan ? NAN_ROLLOUT : OK;
      }
      alpha /= 2;
    }
    *new_cost = std::numeric_limits<double>::quiet_NaN(); *alpha_out = 0.0;
    return LS_EXHAUSTED;
  }

  struct FitTrace {
    int iters = 0;            // number of backward+forward iterations executed
    int32_t status = OK;
    bool converged = false;
    std::vector<double> cost, alpha, du2;   // one entry per executed iteration
  };
  // Optional per-iteration observer: (iter, duff, K, xbar, ubar)
  using Observer = void (*)(void* ctx, int iter, const double* duff, const double* K, const double* xb, const double* ub);

  // src/forward_pass.jl:148-179.  x,u are updated in place to the returned iterate.
  FitTrace fit(int H, double* x, double* u, const double* x_traj, int max_iter, double tol,
               Observer obs = nullptr, void* obs_ctx = nullptr) const {
    const int N = H + 1;
    std::vector<double> duff(H * m), Ks(H * m * n), xb(N * n), ub(H * m);
    FitTrace tr;
    double prev_cost = std::numeric_limits<double>::infinity();
    for (int iter = 1; iter <= max_iter; ++iter) {
      tr.status |= backward_pass(H, x, u, duff.data(), Ks.data());
      double new_cost, alpha;
      int32_t fs = forward_pass(H, x, u, x_traj, duff.data(), Ks.data(), prev_cost, xb.data(), ub.data(), &new_cost, &alpha);
      tr.status |= fs;
      tr.iters = iter;
      tr.cost.push_back(new_cost); tr.alpha.push_back(alpha);
      if (obs) obs(obs_ctx, iter, duff.data(), Ks.data(), xb.data(), ub.data());
      if (fs & LS_EXHAUSTED) { tr.du2.push_back(std::numeric_limits<double>::quiet_NaN()); break; }
      if (!(prev_cost > new_cost)) tr.status |= NOT_DECREASED;
      prev_cost = new_cost;
      double du2 = 0.0;
      for (int i = 0; i < H * m; ++i) { double dd = ub[i] - u[i]; du2 += dd * dd; }
      tr.du2.push_back(du2);
      if (du2 <= tol) { tr.converged = true; break; }   // break BEFORE the update (:171)
      for (int i = 0; i < N * n; ++i) x[i] = xb[i];
      for (int i = 0; i < H * m; ++i) u[i] = ub[i];
    }
    return tr;
  }
};

}  // namespace oracle
