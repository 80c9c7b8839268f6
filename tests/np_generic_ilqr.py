"""Generic NumPy restatement of the reference solver (src/backward_pass.jl, src/forward_pass.jl) for ANY continuous
dynamics fc(x, u) written with NumPy ufuncs (complex-safe), RK4-discretised, with the diagonal quadratic costs.
A, B come from complex-step differentiation of the RK4 step (exact to rounding, like ForwardDiff).  Test infrastructure:
the checker for ILQR_MODEL_CUSTOM problems that have no C++ oracle plugin."""
import numpy as np


class Problem:
    def __init__(self, fc, n, m, dt, x_target, w_x, w_u, w_xf, reg=0.01):
        self.fc, self.n, self.m, self.dt, self.reg = fc, n, m, dt, reg
        self.xt, self.wx, self.wu, self.wxf = (np.asarray(a, dtype=np.float64) for a in (x_target, w_x, w_u, w_xf))

    def step(self, x, u):                       # RBD_helper_functions.jl:72-79 / 2_link_helper_functions.jl:72-78
        dt, f = self.dt, self.fc
        k1 = dt * f(x, u); k2 = dt * f(x + k1 / 2, u); k3 = dt * f(x + k2 / 2, u); k4 = dt * f(x + k3, u)
        return x + (1.0 / 6.0) * (k1 + 2 * k2 + 2 * k3 + k4)

    def linearize(self, x, u):                  # src/backward_pass.jl:25-40
        h = 1e-30
        A = np.zeros((self.n, self.n)); B = np.zeros((self.n, self.m))
        for j in range(self.n):
            xc = x.astype(np.complex128); xc[j] += 1j * h
            A[:, j] = self.step(xc, u.astype(np.complex128)).imag / h
        for j in range(self.m):
            uc = u.astype(np.complex128); uc[j] += 1j * h
            B[:, j] = self.step(x.astype(np.complex128), uc).imag / h
        return A, B

    def l(self, x, u):
        e = self.xt - x
        return float(np.sum(self.wx * e * e) + np.sum(self.wu * u * u))

    def lf(self, x):
        e = self.xt - x
        return float(np.sum(self.wxf * e * e))

    def rollout(self, x0, u):
        x = np.zeros((u.shape[0] + 1, self.n)); x[0] = x0
        for k in range(u.shape[0]):
            x[k + 1] = self.step(x[k], u[k])
        return x

    def backward_pass(self, x, u):              # src/backward_pass.jl:324-357
        H = u.shape[0]
        S = np.diag(2 * self.wxf); s = -2 * self.wxf * (self.xt - x[H])
        d = np.zeros((H, self.m)); K = np.zeros((H, self.m, self.n))
        for k in range(H - 1, -1, -1):
            A, B = self.linearize(x[k], u[k])
            q = -2 * self.wx * (self.xt - x[k]); Q = np.diag(2 * self.wx); r = 2 * self.wu * u[k]; R = np.diag(2 * self.wu)
            g = r + B.T @ s; G = (B.T @ S) @ A; Hm = R + (B.T @ S) @ B
            Hreg = Hm + self.reg * np.eye(self.m)
            du = np.linalg.solve(-Hreg, g); Kk = np.linalg.solve(-Hreg, G)
            d[k] = du; K[k] = Kk
            s = q + A.T @ s + (Kk.T @ Hm) @ du + Kk.T @ g + G.T @ du
            S = Q + (A.T @ S) @ A + (Kk.T @ Hm) @ Kk + Kk.T @ G + G.T @ Kk
        return d, K

    def candidate(self, x, u, d, K, alpha):     # src/forward_pass.jl:71-76
        H = u.shape[0]
        xb = np.zeros_like(x); ub = np.zeros_like(u); xb[0] = x[0]
        cost = 0.0
        for k in range(H):
            ub[k] = u[k] + alpha * d[k] + K[k] @ (xb[k] - x[k])
            cost += self.l(xb[k], ub[k])
            xb[k + 1] = self.step(xb[k], ub[k])
        return xb, ub, cost + self.lf(xb[H])

    def fit(self, x, u, max_iter=100, tol=1e-6, jmax=32):   # src/forward_pass.jl:148-179
        x = x.copy(); u = u.copy(); prev = np.inf
        costs, alphas = [], []
        for it in range(1, max_iter + 1):
            d, K = self.backward_pass(x, u)
            alpha = 1.0
            for _ in range(jmax):
                xb, ub, c = self.candidate(x, u, d, K, alpha)
                if prev - c > 0:
                    break
                alpha /= 2
            else:
                raise RuntimeError("line search exhausted")
            costs.append(c); alphas.append(alpha); prev = c
            if np.sum((ub - u) ** 2) <= tol:
                return x, u, np.array(costs), np.array(alphas), it
            x, u = xb, ub
        return x, u, np.array(costs), np.array(alphas), max_iter
