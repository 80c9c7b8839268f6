"""ctypes binding of oracle/liboracle.so — TEST INFRASTRUCTURE ONLY.

Importable only from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Never import this from the product
package (ilqr.jl_b200/).

Arrays follow the Julia column-major layouts: x[N,n] is passed as a Fortran-
ordered (N,n) NumPy array, K[H,m,n] as Fortran (H,m,n); batched arrays carry
the batch as the trailing dimension ((N,n,B) Fortran order).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int32)


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("oracle_capi.cpp", "ilqr_oracle.hpp", "serial_chain.hpp")]
    def stale():
        return not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs)
    if force or stale():
        import fcntl
        with open(os.path.join(_HERE, ".build.lock"), "w") as lock:   # several test workers / ranks may get here together
            fcntl.flock(lock, fcntl.LOCK_EX)
            try:
                if force or stale():
                    subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s"], stdout=subprocess.DEVNULL)
            finally:
                fcntl.flock(lock, fcntl.LOCK_UN)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        _LIB.oracle_two_link_total_cost.restype = ctypes.c_double
        _LIB.oracle_two_link_rollout_candidate.restype = ctypes.c_double
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _f(a, shape=None):
    a = np.asfortranarray(np.asarray(a, dtype=np.float64))
    if shape is not None:
        assert a.shape == tuple(shape), (a.shape, shape)
    return a


def constants():
    out = np.zeros(6)
    lib().oracle_two_link_constants(_p(out))
    return dict(alpha=out[0], beta=out[1], delta=out[2], dt=out[3], theta_star=out[4:6].copy())


def dynamics(x, u):
    x = _f(x, (4,)); u = _f(u, (2,)); y = np.zeros(4)
    lib().oracle_two_link_dynamics(_p(x), _p(u), _p(y))
    return y


def continuous_dynamics(x, u):
    x = _f(x, (4,)); u = _f(u, (2,)); y = np.zeros(4)
    lib().oracle_two_link_continuous_dynamics(_p(x), _p(u), _p(y))
    return y


def linearize(x, u):
    x = _f(x, (4,)); u = _f(u, (2,))
    A = np.zeros((4, 4), order="F"); B = np.zeros((4, 2), order="F")
    lib().oracle_two_link_linearize(_p(x), _p(u), _p(A), _p(B))
    return A, B


def cost_quad(x, u):
    x = _f(x, (4,)); u = _f(u, (2,))
    q = ctypes.c_double()
    qv = np.zeros(4); rv = np.zeros(2)
    Q = np.zeros((4, 4), order="F"); P = np.zeros((2, 4), order="F"); R = np.zeros((2, 2), order="F")
    lib().oracle_two_link_cost_quad(_p(x), _p(u), ctypes.byref(q), _p(qv), _p(rv), _p(Q), _p(P), _p(R))
    return q.value, qv, rv, Q, P, R


def final_cost_quad(x):
    x = _f(x, (4,))
    q = ctypes.c_double(); qv = np.zeros(4); Q = np.zeros((4, 4), order="F")
    lib().oracle_two_link_final_cost_quad(_p(x), ctypes.byref(q), _p(qv), _p(Q))
    return q.value, qv, Q


def rollout(x0, u):
    u = _f(u); H = u.shape[0]
    x = np.zeros((H + 1, 4), order="F")
    lib().oracle_two_link_rollout(H, _p(_f(x0, (4,))), _p(u), _p(x))
    return x


def backward_pass(x, u, reg=0.01):
    u = _f(u); H = u.shape[0]; x = _f(x, (H + 1, 4))
    d = np.zeros((H, 2), order="F"); K = np.zeros((H, 2, 4), order="F")
    st = lib().oracle_two_link_backward_pass(H, _p(x), _p(u), ctypes.c_double(reg), _p(d), _p(K))
    return d, K, st


def total_cost(x, u, x_traj=None):
    u = _f(u); H = u.shape[0]; x = _f(x, (H + 1, 4))
    xt = None if x_traj is None else _f(x_traj, (H + 1, 4))
    return lib().oracle_two_link_total_cost(H, _p(x), _p(u), _p(xt))


def forward_pass(x, u, d, K, prev_cost, jmax=32, x_traj=None):
    u = _f(u); H = u.shape[0]; x = _f(x, (H + 1, 4)); d = _f(d, (H, 2)); K = _f(K, (H, 2, 4))
    xt = None if x_traj is None else _f(x_traj, (H + 1, 4))
    xb = np.zeros((H + 1, 4), order="F"); ub = np.zeros((H, 2), order="F")
    c = ctypes.c_double(); a = ctypes.c_double()
    st = lib().oracle_two_link_forward_pass(H, _p(x), _p(u), _p(xt), _p(d), _p(K), ctypes.c_double(prev_cost), jmax,
                                            _p(xb), _p(ub), ctypes.byref(c), ctypes.byref(a))
    return xb, ub, c.value, a.value, st


def rollout_candidate(x, u, d, K, alpha, x_traj=None):
    u = _f(u); H = u.shape[0]; x = _f(x, (H + 1, 4)); d = _f(d, (H, 2)); K = _f(K, (H, 2, 4))
    xt = None if x_traj is None else _f(x_traj, (H + 1, 4))
    xb = np.zeros((H + 1, 4), order="F"); ub = np.zeros((H, 2), order="F")
    c = lib().oracle_two_link_rollout_candidate(H, _p(x), _p(u), _p(xt), _p(d), _p(K), ctypes.c_double(alpha),
                                                _p(xb), _p(ub))
    return xb, ub, c


def fit(x_init, u_init, x_traj=None, max_iter=100, tol=1e-6, reg=0.01, jmax=32, max_dump=0):
    """Returns dict(x,u,cost,alpha,du2,iters,converged,status[,dump_*])."""
    u = _f(u_init).copy(order="F"); H = u.shape[0]; x = _f(x_init, (H + 1, 4)).copy(order="F")
    xt = None if x_traj is None else _f(x_traj, (H + 1, 4))
    cost = np.full(max_iter, np.nan); alpha = np.full(max_iter, np.nan); du2 = np.full(max_iter, np.nan)
    it = ctypes.c_int32(); cv = ctypes.c_int32()
    dd = dK = dx = du = None
    if max_dump > 0:
        dd = np.zeros((H, 2, max_dump), order="F"); dK = np.zeros((H, 2, 4, max_dump), order="F")
        dx = np.zeros((H + 1, 4, max_dump), order="F"); du = np.zeros((H, 2, max_dump), order="F")
    st = lib().oracle_two_link_fit(H, _p(x), _p(u), _p(xt), max_iter, ctypes.c_double(tol), ctypes.c_double(reg), jmax,
                                   _p(cost), _p(alpha), _p(du2), ctypes.byref(it), ctypes.byref(cv), max_dump,
                                   _p(dd), _p(dK), _p(dx), _p(du))
    out = dict(x=x, u=u, cost=cost[: it.value], alpha=alpha[: it.value], du2=du2[: it.value], iters=it.value,
               converged=bool(cv.value), status=st)
    if max_dump > 0:
        out.update(dump_duff=dd, dump_K=dK, dump_xbar=dx, dump_ubar=du)
    return out


def fit_batch(x_init, u_init, x_traj=None, max_iter=100, tol=1e-6, reg=0.01, jmax=32, nthreads=1, traces=True):
    """x_init (N,4,B), u_init (H,2,B) Fortran-ordered.  Returns dict."""
    u = _f(u_init).copy(order="F"); H, _, B = u.shape
    x = _f(x_init, (H + 1, 4, B)).copy(order="F")
    xt = None if x_traj is None else _f(x_traj, (H + 1, 4, B))
    cost = alpha = du2 = None
    if traces:
        cost = np.full((max_iter, B), np.nan, order="F"); alpha = np.full((max_iter, B), np.nan, order="F")
        du2 = np.full((max_iter, B), np.nan, order="F")
    iters = np.zeros(B, dtype=np.int32); conv = np.zeros(B, dtype=np.int32); status = np.zeros(B, dtype=np.int32)
    lib().oracle_two_link_fit_batch(B, H, _p(x), _p(u), _p(xt), max_iter, ctypes.c_double(tol), ctypes.c_double(reg),
                                    jmax, nthreads, _p(cost), _p(alpha), _p(du2),
                                    iters.ctypes.data_as(_ip), conv.ctypes.data_as(_ip), status.ctypes.data_as(_ip))
    return dict(x=x, u=u, cost=cost, alpha=alpha, du2=du2, iters=iters, converged=conv.astype(bool), status=status)


def fit_target(x_init, u_init, target, max_iter=100, tol=1e-6, reg=0.01, jmax=32):
    """One trajectory of the 2-link plugin with target tool location `target` = (x, y) (2_link_helper_functions.jl:16).
    x_init (N,4), u_init (H,2).  Returns (x, u, iterations, status)."""
    u = _f(u_init).copy(order="F"); H = u.shape[0]
    x = _f(x_init, (H + 1, 4)).copy(order="F")
    st = ctypes.c_int32(0)
    fn = lib().oracle_two_link_fit_target
    fn.restype = ctypes.c_int
    it = fn(ctypes.c_double(target[0]), ctypes.c_double(target[1]), H, _p(x), _p(u), max_iter, ctypes.c_double(tol),
            ctypes.c_double(reg), jmax, ctypes.byref(st))
    return x, u, int(it), int(st.value)


# ---- the 2-link arm with the FK tool-point cost + a cross term (oracle TwoLinkToolCost): generic quadratisation incl. 𝐏
_cd = ctypes.c_double


def tool_cost_quad(x, u, w_tool=1.0, w_final=50.0, gamma=0.3):
    """immediate_cost_quadratization (src/backward_pass.jl:81-109) of the tool-point cost → (q, 𝐪[4], 𝐫[2], 𝐐[4,4], 𝐏[2,4], 𝐑[2,2])."""
    out = np.zeros(35)
    lib().oracle_tool_cost_quad(_cd(w_tool), _cd(w_final), _cd(gamma), _p(_f(x)), _p(_f(u)), _p(out))
    return (out[0], out[1:5].copy(), out[5:7].copy(), out[7:23].reshape(4, 4, order="F"), out[23:31].reshape(2, 4, order="F"),
            out[31:35].reshape(2, 2, order="F"))


def tool_backward_pass(x, u, w_tool=1.0, w_final=50.0, gamma=0.3, reg=0.01):
    u = _f(u); H = u.shape[0]; x = _f(x, (H + 1, 4))
    d = np.zeros((H, 2), order="F"); K = np.zeros((H, 2, 4), order="F")
    st = lib().oracle_tool_backward_pass(_cd(w_tool), _cd(w_final), _cd(gamma), H, _p(x), _p(u), _cd(reg), _p(d), _p(K))
    return d, K, st


def tool_total_cost(x, u, x_traj=None, w_tool=1.0, w_final=50.0, gamma=0.3):
    u = _f(u); H = u.shape[0]; x = _f(x, (H + 1, 4))
    lib().oracle_tool_total_cost.restype = ctypes.c_double
    return lib().oracle_tool_total_cost(_cd(w_tool), _cd(w_final), _cd(gamma), H, _p(x), _p(u), _p(None if x_traj is None else _f(x_traj)))


def tool_fit_batch(x_init, u_init, x_traj=None, w_tool=1.0, w_final=50.0, gamma=0.3, max_iter=100, tol=1e-6, reg=0.01, jmax=32,
                   nthreads=1):
    u = _f(u_init).copy(order="F"); H, _, B = u.shape
    x = _f(x_init, (H + 1, 4, B)).copy(order="F")
    xt = None if x_traj is None else _f(x_traj, (H + 1, 4, B))
    cost = np.full((max_iter, B), np.nan, order="F"); alpha = np.full((max_iter, B), np.nan, order="F")
    du2 = np.full((max_iter, B), np.nan, order="F")
    iters = np.zeros(B, dtype=np.int32); conv = np.zeros(B, dtype=np.int32); status = np.zeros(B, dtype=np.int32)
    lib().oracle_tool_fit_batch(_cd(w_tool), _cd(w_final), _cd(gamma), B, H, _p(x), _p(u), _p(xt), max_iter, _cd(tol), _cd(reg),
                                jmax, nthreads, _p(cost), _p(alpha), _p(du2), iters.ctypes.data_as(_ip), conv.ctypes.data_as(_ip),
                                status.ctypes.data_as(_ip))
    return dict(x=x, u=u, cost=cost, alpha=alpha, du2=du2, iters=iters, converged=conv.astype(bool), status=status)


def lq32_backward_pass(A, B, Q, R, Qf, x, u, reg=0.01):
    u = _f(u); H = u.shape[0]; x = _f(x, (H + 1, 3))
    d = np.zeros((H, 2), order="F"); K = np.zeros((H, 2, 3), order="F")
    st = lib().oracle_lq32_backward_pass(H, _p(_f(A)), _p(_f(B)), _p(_f(Q)), _p(_f(R)), _p(_f(Qf)), _p(x), _p(u),
                                         ctypes.c_double(reg), _p(d), _p(K))
    return d, K, st


def lq32_fit(A, B, Q, R, Qf, x, u, max_iter=10, tol=1e-6, reg=0.01):
    u = _f(u).copy(order="F"); H = u.shape[0]; x = _f(x, (H + 1, 3)).copy(order="F")
    cost = np.full(max_iter, np.nan); it = ctypes.c_int32()
    st = lib().oracle_lq32_fit(H, _p(_f(A)), _p(_f(B)), _p(_f(Q)), _p(_f(R)), _p(_f(Qf)), _p(x), _p(u), max_iter,
                               ctypes.c_double(tol), ctypes.c_double(reg), _p(cost), ctypes.byref(it))
    return x, u, cost[: it.value], it.value, st


def hardware_threads():
    return int(lib().oracle_hardware_threads())


# ---------------------------------------------------------------------------------------------
# Serial-chain rigid-body plugin (oracle/serial_chain.hpp)
# ---------------------------------------------------------------------------------------------
CHAIN_STRIDE = 20   # xyz(3) rpy(3) axis(3) mass(1) com(3) ixx ixy ixz iyy iyz izz (6) pad(1)


class ChainSpec(ctypes.Structure):
    _fields_ = [("nq", ctypes.c_int32), ("floating", ctypes.c_int32), ("dt", ctypes.c_double),
                ("gravity", ctypes.c_double * 3), ("joints", ctypes.c_double * (9 * CHAIN_STRIDE)),
                ("x_target", ctypes.c_double * 16), ("w_x", ctypes.c_double * 16), ("w_u", ctypes.c_double * 8),
                ("w_xf", ctypes.c_double * 16)]

    @property
    def nv(self):      # velocity (= control) dimension
        return self.nq + (6 if self.floating else 0)

    @property
    def nx(self):
        return 2 * self.nv


def chain_spec(joints, gravity=(0.0, 0.0, 0.0), dt=0.01, x_target=None, w_x=None, w_u=None, w_xf=None, base=None):
    """joints: (nq, 20) array, one row per joint+child link (layout CHAIN_STRIDE above).
    base (optional, 20 doubles; only mass / com / inertia are used): makes the mechanism floating-base
    (RBD_helper_functions.jl:7, floating = true): x = [p(3) MRP; r(3); θ; ω(3); v(3); θ̇]."""
    joints = np.asarray(joints, dtype=np.float64)
    nq = joints.shape[0]
    assert joints.shape == (nq, CHAIN_STRIDE), joints.shape
    s = ChainSpec()
    s.nq = nq; s.dt = dt; s.floating = 0 if base is None else 1
    assert (nq in (1, 2)) if s.floating else (nq in (2, 3, 6, 7)), nq
    for k in range(3):
        s.gravity[k] = float(gravity[k])
    if s.floating:
        assert not any(gravity), "floating base: gravity must be zero (as in the reference)"
        joints = np.concatenate([joints, np.asarray(base, dtype=np.float64).reshape(1, CHAIN_STRIDE)])
    flat = joints.reshape(-1)
    for i, v in enumerate(flat):
        s.joints[i] = float(v)
    nv = s.nv
    for name, arr, cnt in (("x_target", x_target, 2 * nv), ("w_x", w_x, 2 * nv), ("w_u", w_u, nv), ("w_xf", w_xf, 2 * nv)):
        if arr is not None:
            arr = np.asarray(arr, dtype=np.float64)
            assert arr.shape == (cnt,), (name, arr.shape)
            for i in range(cnt):
                getattr(s, name)[i] = float(arr[i])
    return s


def _chk(rc):
    if rc < 0:
        raise ValueError("oracle: unsupported chain size")
    return rc


def chain_mass_bias(spec, q, vel):
    nq, nv = spec.nq, spec.nv
    M = np.zeros((nv, nv), order="F"); b = np.zeros(nv)
    _chk(lib().oracle_chain_mass_bias(ctypes.byref(spec), _p(_f(q, (nq,))), _p(_f(vel, (nv,))), _p(M), _p(b)))
    return M, b


def chain_continuous_dynamics(spec, x, u):
    nx, nv = spec.nx, spec.nv; y = np.zeros(nx)
    _chk(lib().oracle_chain_continuous_dynamics(ctypes.byref(spec), _p(_f(x, (nx,))), _p(_f(u, (nv,))), _p(y)))
    return y


def chain_dynamics(spec, x, u):
    nx, nv = spec.nx, spec.nv; y = np.zeros(nx)
    _chk(lib().oracle_chain_dynamics(ctypes.byref(spec), _p(_f(x, (nx,))), _p(_f(u, (nv,))), _p(y)))
    return y


def chain_linearize(spec, x, u):
    nx, nv = spec.nx, spec.nv
    A = np.zeros((nx, nx), order="F"); B = np.zeros((nx, nv), order="F")
    _chk(lib().oracle_chain_linearize(ctypes.byref(spec), _p(_f(x, (nx,))), _p(_f(u, (nv,))), _p(A), _p(B)))
    return A, B


def chain_rollout(spec, x0, u):
    nx, nv = spec.nx, spec.nv; u = _f(u); H = u.shape[0]
    x = np.zeros((H + 1, nx), order="F")
    _chk(lib().oracle_chain_rollout(ctypes.byref(spec), H, _p(_f(x0, (nx,))), _p(u), _p(x)))
    return x


def chain_backward_pass(spec, x, u, reg=0.01):
    nx, nv = spec.nx, spec.nv; u = _f(u); H = u.shape[0]; x = _f(x, (H + 1, nx))
    d = np.zeros((H, nv), order="F"); K = np.zeros((H, nv, nx), order="F")
    st = _chk(lib().oracle_chain_backward_pass(ctypes.byref(spec), H, _p(x), _p(u), ctypes.c_double(reg), _p(d), _p(K)))
    return d, K, st


def chain_total_cost(spec, x, u, x_traj=None):
    nx, nv = spec.nx, spec.nv; u = _f(u); H = u.shape[0]; x = _f(x, (H + 1, nx))
    xt = None if x_traj is None else _f(x_traj, (H + 1, nx))
    c = ctypes.c_double()
    _chk(lib().oracle_chain_total_cost(ctypes.byref(spec), H, _p(x), _p(u), _p(xt), ctypes.byref(c)))
    return c.value


def chain_forward_pass(spec, x, u, d, K, prev_cost, jmax=32, x_traj=None):
    nx, nv = spec.nx, spec.nv; u = _f(u); H = u.shape[0]; x = _f(x, (H + 1, nx)); d = _f(d, (H, nv)); K = _f(K, (H, nv, nx))
    xt = None if x_traj is None else _f(x_traj, (H + 1, nx))
    xb = np.zeros((H + 1, nx), order="F"); ub = np.zeros((H, nv), order="F")
    c = ctypes.c_double(); a = ctypes.c_double()
    st = _chk(lib().oracle_chain_forward_pass(ctypes.byref(spec), H, _p(x), _p(u), _p(xt), _p(d), _p(K),
                                              ctypes.c_double(prev_cost), jmax, _p(xb), _p(ub), ctypes.byref(c),
                                              ctypes.byref(a)))
    return xb, ub, c.value, a.value, st


def chain_fit_batch(spec, x_init, u_init, x_traj=None, max_iter=100, tol=1e-6, reg=0.01, jmax=32, nthreads=1, traces=True):
    """x_init (N,n,B), u_init (H,m,B) Fortran-ordered.  Returns dict like fit_batch."""
    nx = spec.nx
    u = _f(u_init).copy(order="F"); H, _, B = u.shape
    x = _f(x_init, (H + 1, nx, B)).copy(order="F")
    xt = None if x_traj is None else _f(x_traj, (H + 1, nx, B))
    cost = alpha = du2 = None
    if traces:
        cost = np.full((max_iter, B), np.nan, order="F"); alpha = np.full((max_iter, B), np.nan, order="F")
        du2 = np.full((max_iter, B), np.nan, order="F")
    iters = np.zeros(B, dtype=np.int32); conv = np.zeros(B, dtype=np.int32); status = np.zeros(B, dtype=np.int32)
    _chk(lib().oracle_chain_fit_batch(ctypes.byref(spec), B, H, _p(x), _p(u), _p(xt), max_iter, ctypes.c_double(tol),
                                      ctypes.c_double(reg), jmax, nthreads, _p(cost), _p(alpha), _p(du2),
                                      iters.ctypes.data_as(_ip), conv.ctypes.data_as(_ip), status.ctypes.data_as(_ip)))
    return dict(x=x, u=u, cost=cost, alpha=alpha, du2=du2, iters=iters, converged=conv.astype(bool), status=status)
