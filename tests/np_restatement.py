"""Independent NumPy restatement of the reference algorithm (closed-form 2-link).

Written separately from oracle/ilqr_oracle.hpp: the oracle follows the Julia
sources literally with dual numbers and LU solves; this file uses the derived
closed forms and numpy.linalg.  Two independent restatements
agreeing to ~1e-12 is the substitute for the Julia oracle that cannot run here.

Reference lines: src/backward_pass.jl:324-357, src/forward_pass.jl:55-93,148-196,
test/2_link_example/2_link_helper_functions.jl:4-108.
"""
import math

import numpy as np

L1 = L2 = math.sqrt(2.0) / 2.0
R1 = R2 = 0.5 * L1
M1 = M2 = 1.0
IZ1 = 1.0 / 12.0 * M1 * L1 ** 2
IZ2 = 1.0 / 12.0 * M2 * L2 ** 2
ALPHA = IZ1 + IZ2 + M1 * R1 ** 2 + M2 * (L1 ** 2 + R2 ** 2)
BETA = M2 * L1 * R2
DELTA = IZ2 + M2 * R2 ** 2
DT = 0.01
TARGET_TOOL = (0.6, -0.5)


def inverse_kinematics(w):
    x, y = w
    q2 = math.acos((x * x + y * y - L1 ** 2 - L2 ** 2) / (2 * L1 * L2))
    q1 = math.atan2(y, x) - math.atan2(L2 * math.sin(q2), L1 + L2 * math.cos(q2))
    return np.array([q1, q2])


THETA_STAR = inverse_kinematics(TARGET_TOOL)


def fc(s, u):
    """Continuous dynamics, closed form: acc = M^-1 (u - C w)."""
    c2, s2 = math.cos(s[1]), math.sin(s[1])
    M = np.array([[ALPHA + 2 * BETA * c2, DELTA + BETA * c2], [DELTA + BETA * c2, DELTA]])
    w = s[2:4]
    C = -BETA * s2 * w[1] * np.array([[1.0, 0.5], [0.5, 0.0]])
    acc = np.linalg.solve(M, u - C @ w)
    return np.concatenate([w, acc])


def fc_jac(s, u):
    """Phi = d fc/d s (4x4), Psi = d fc/d u (4x2), analytic."""
    c2, s2 = math.cos(s[1]), math.sin(s[1])
    M = np.array([[ALPHA + 2 * BETA * c2, DELTA + BETA * c2], [DELTA + BETA * c2, DELTA]])
    Mi = np.linalg.inv(M)
    w1, w2 = s[2], s[3]
    h = -BETA * s2 * w2 * np.array([w1 + 0.5 * w2, 0.5 * w1])
    acc = Mi @ (u - h)
    dM = -BETA * s2 * np.array([[2.0, 1.0], [1.0, 0.0]])
    dh_dth2 = -BETA * c2 * w2 * np.array([w1 + 0.5 * w2, 0.5 * w1])
    dh_dw1 = -BETA * s2 * w2 * np.array([1.0, 0.5])
    dh_dw2 = -BETA * s2 * np.array([w1 + w2, 0.5 * w1])
    Phi = np.zeros((4, 4))
    Phi[0, 2] = 1.0
    Phi[1, 3] = 1.0
    Phi[2:, 1] = -Mi @ (dM @ acc + dh_dth2)
    Phi[2:, 2] = -Mi @ dh_dw1
    Phi[2:, 3] = -Mi @ dh_dw2
    Psi = np.zeros((4, 2))
    Psi[2:, :] = Mi
    return Phi, Psi


def dynamicsf(x, u):
    k1 = DT * fc(x, u)
    k2 = DT * fc(x + k1 / 2, u)
    k3 = DT * fc(x + k2 / 2, u)
    k4 = DT * fc(x + k3, u)
    return x + (1 / 6) * (k1 + 2 * k2 + 2 * k3 + k4)


def linearize(x, u):
    """A = d f/d x, B = d f/d u of the discrete RK4 map via the stage chain."""
    k1 = DT * fc(x, u)
    s2 = x + k1 / 2
    k2 = DT * fc(s2, u)
    s3 = x + k2 / 2
    k3 = DT * fc(s3, u)
    s4 = x + k3
    I = np.eye(4)
    P1, Y1 = fc_jac(x, u)
    P2, Y2 = fc_jac(s2, u)
    P3, Y3 = fc_jac(s3, u)
    P4, Y4 = fc_jac(s4, u)
    D1 = DT * P1
    D2 = DT * P2 @ (I + D1 / 2)
    D3 = DT * P3 @ (I + D2 / 2)
    D4 = DT * P4 @ (I + D3)
    A = I + (D1 + 2 * D2 + 2 * D3 + D4) / 6
    E1 = DT * Y1
    E2 = DT * (P2 @ E1 / 2 + Y2)
    E3 = DT * (P3 @ E2 / 2 + Y3)
    E4 = DT * (P4 @ E3 + Y4)
    B = (E1 + 2 * E2 + 2 * E3 + E4) / 6
    return A, B


def immediate_cost(x, u):
    return float(np.sum((THETA_STAR - x[:2]) ** 2) + np.sum(u ** 2))


def final_cost(x):
    return float(np.sum((THETA_STAR - x[:2]) ** 2))


def cost_quad(x, u):
    qv = np.zeros(4)
    qv[:2] = -2 * (THETA_STAR - x[:2])
    Q = np.diag([2.0, 2.0, 0.0, 0.0])
    rv = 2 * u
    R = 2 * np.eye(2)
    P = np.zeros((2, 4))
    return qv, rv, Q, P, R


def backward_pass(x, u, reg=0.01):
    """x[N,4], u[H,2] (row = time).  Returns d[H,2], K[H,2,4]."""
    H = u.shape[0]
    d = np.zeros((H, 2))
    K = np.zeros((H, 2, 4))
    sv = np.zeros(4)
    sv[:2] = -2 * (THETA_STAR - x[H, :2])
    S = np.diag([2.0, 2.0, 0.0, 0.0])
    for k in range(H - 1, -1, -1):
        A, B = linearize(x[k], u[k])
        qv, rv, Q, P, R = cost_quad(x[k], u[k])
        g = rv + B.T @ sv
        G = P + B.T @ S @ A
        Hm = R + B.T @ S @ B
        Hr = Hm + reg * np.eye(2)
        dk = -np.linalg.solve(Hr, g)
        Kk = -np.linalg.solve(Hr, G)
        d[k] = dk
        K[k] = Kk
        sv = qv + A.T @ sv + Kk.T @ Hm @ dk + Kk.T @ g + G.T @ dk
        S = Q + A.T @ S @ A + Kk.T @ Hm @ Kk + Kk.T @ G + G.T @ Kk
    return d, K


def total_cost(xb, ub, x_traj=None):
    H = ub.shape[0]
    s = 0.0
    for i in range(H):
        xi = xb[i] if x_traj is None else xb[i] - x_traj[i]
        s += immediate_cost(xi, ub[i])
    s += final_cost(xb[H])
    return s


def rollout_candidate(x, u, d, K, alpha, x_traj=None):
    H = u.shape[0]
    xb = np.zeros_like(x)
    ub = np.zeros_like(u)
    xb[0] = x[0]
    for k in range(H):
        ub[k] = u[k] + alpha * d[k] + K[k] @ (xb[k] - x[k])
        xb[k + 1] = dynamicsf(xb[k], ub[k])
    return xb, ub, total_cost(xb, ub, x_traj)


def forward_pass(x, u, d, K, prev_cost, jmax=32, x_traj=None):
    alpha = 1.0
    for _ in range(jmax):
        xb, ub, c = rollout_candidate(x, u, d, K, alpha, x_traj)
        if prev_cost - c > 0:
            return xb, ub, c, alpha
        alpha /= 2
    return None, None, float("nan"), 0.0


def open_loop_rollout(x0, u):
    H = u.shape[0]
    x = np.zeros((H + 1, 4))
    x[0] = x0
    for k in range(H):
        x[k + 1] = dynamicsf(x[k], u[k])
    return x


def fit(x, u, max_iter=100, tol=1e-6, reg=0.01, jmax=32, x_traj=None):
    """Returns (x, u, trace) with the reference's return-previous-iterate semantics."""
    prev = float("inf")
    trace = {"cost": [], "alpha": [], "du2": [], "converged": False}
    for _ in range(max_iter):
        d, K = backward_pass(x, u, reg)
        xb, ub, c, a = forward_pass(x, u, d, K, prev, jmax, x_traj)
        trace["cost"].append(c)
        trace["alpha"].append(a)
        if xb is None:
            break
        prev = c
        du2 = float(np.sum((ub - u) ** 2))
        trace["du2"].append(du2)
        if du2 <= tol:
            trace["converged"] = True
            break
        x, u = xb, ub
    return x, u, trace
