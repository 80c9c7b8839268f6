"""Host-side mirror of the reference's interface for the hot path.

The reference (aabouman/iLQR.jl, paths relative to /root/reference) exposes
  fit(x_init, u_init, dynamicsf, immediate_cost, final_cost; x_traj, max_iter, tol)   src/forward_pass.jl:148
  backward_pass(x, u, dynamicsf, immediate_cost, final_cost) -> (δuff, K)              src/backward_pass.jl:324
  forward_pass(x, u, x_traj, δuff, K, prev_cost, dynamicsf, immediate_cost, final_cost) src/forward_pass.jl:55
The three callbacks are replaced by one `problem` (model id + parameters, see
include/ilqr_b200.h); everything else keeps its name, argument meaning and
error behaviour (the reference's @assert → AssertionError).  Arrays are the
reference's shapes x[N,n], u[H,m], K[H,m,n] or their batched forms with a
trailing batch axis.  Every call goes through the C ABI of libilqr_b200.so —
the same entry points a Julia host binds with ccall (julia/iLQRB200.jl).
"""
import ctypes

import numpy as np

from . import _abi
from ._abi import Problem


class IlqrError(RuntimeError):
    pass


def two_link_problem(H, B=1, n_alpha=32, trace_iters=0, device=0, reg=None, variant=_abi.VARIANT_AUTO):
    """The reference's 2-link plugin (test/2_link_example/2_link_helper_functions.jl) as an ilqr_problem."""
    lib = _abi.load_library()
    p = Problem()
    rc = lib.ilqr_problem_two_link(ctypes.byref(p), int(H), int(B))
    if rc != 0:
        raise IlqrError("ilqr_problem_two_link failed")
    p.n_alpha = n_alpha
    p.trace_iters = trace_iters
    p.device = device
    p.variant = variant
    if reg is not None:
        p.reg = reg
    return p


def serial_chain_problem(joints, H, B=1, gravity=(0.0, 0.0, 0.0), x_target=None, w_x=None, w_u=None, w_xf=None, dt=0.01,
                         n_alpha=32, trace_iters=0, device=0, reg=None, base=None):
    """The reference's rigid-body plugin (test/RBD_2_link_example/RBD_helper_functions.jl:48-116) for a fixed-base
    serial chain: `joints` is an (nq, 20) array, one row per joint + child link — origin xyz(3), rpy(3), unit
    axis(3), mass, COM(3), ixx ixy ixz iyy iyz izz, pad — e.g. from load_urdf().  Costs are the diagonal
    quadratics l = Σ w_x (x* − x)² + Σ w_u u², lf = Σ w_xf (x* − x)².

    base: (mass, com[3], inertia[6]) of the root link (load_urdf's second result) makes the mechanism floating-base
    as in the reference (RBD_helper_functions.jl:7): x = [p(3) MRP; r(3); θ; ω(3); v(3); θ̇], u = base wrench
    [torque; force] then joint torques; zero gravity only."""
    lib = _abi.load_library()
    joints = np.ascontiguousarray(np.asarray(joints, dtype=np.float64))
    nq = joints.shape[0]
    if joints.shape != (nq, _abi.CHAIN_STRIDE):
        raise ValueError("joints must be (nq, %d)" % _abi.CHAIN_STRIDE)
    g = np.ascontiguousarray(np.asarray(gravity, dtype=np.float64))
    p = Problem()
    if base is None:
        rc = lib.ilqr_problem_serial_chain(ctypes.byref(p), nq, joints.ctypes.data, g.ctypes.data, int(H), int(B))
        nv = nq
    else:
        if any(g):
            raise ValueError("floating base: gravity must be zero")
        row = np.zeros(_abi.CHAIN_STRIDE)
        if len(base) == _abi.CHAIN_STRIDE:
            row[:] = np.asarray(base, dtype=np.float64)
        else:
            row[9] = base[0]; row[10:13] = base[1]; row[13:19] = base[2]
        rc = lib.ilqr_problem_floating_chain(ctypes.byref(p), nq, joints.ctypes.data, row.ctypes.data, int(H), int(B))
        nv = 6 + nq
    if rc != 0:
        raise IlqrError("ilqr_problem_serial_chain failed")
    p.dt = dt
    p.n_alpha = n_alpha
    p.trace_iters = trace_iters
    p.device = device
    if reg is not None:
        p.reg = reg
    for name, arr, cnt in (("x_target", x_target, 2 * nv), ("w_x", w_x, 2 * nv), ("w_u", w_u, nv), ("w_xf", w_xf, 2 * nv)):
        if arr is not None:
            arr = np.asarray(arr, dtype=np.float64)
            if arr.shape != (cnt,):
                raise ValueError("%s must have %d entries" % (name, cnt))
            for i in range(cnt):
                getattr(p, name)[i] = float(arr[i])
    return p


_custom_sources = []   # keeps the source bytes of custom problems alive (ilqr_problem holds a raw pointer)


def custom_problem(dynamics_src, n, m, H, B=1, dt=0.01, params=(), x_target=None, w_x=None, w_u=None, w_xf=None, n_alpha=32,
                   trace_iters=0, device=0, reg=None, user_cost=False):
    """Any dynamics (the reference accepts any Julia function as `dynamicsf`, src/forward_pass.jl:148-153): CUDA C++ source
    defining  template <class T> __device__ void ilqr_dynamics(const T* x, const T* u, const double* p, T* xdot);
    compiled at run time (NVRTC) into the library's kernels, RK4-discretised with step dt and differentiated with dual
    numbers.  Costs are the diagonal quadratics l = Σ w_x (x* − x)² + Σ w_u u², lf = Σ w_xf (x* − x)² — or, with
    user_cost=True, the functions  template <class T> __device__ T ilqr_cost(const T* x, const T* u, const double* p)  and
    template <class T> __device__ T ilqr_final_cost(const T* x, const double* p)  defined in the same snippet (the reference
    takes any Julia function as immediate_cost / final_cost), expanded to second order incl. the cross term ∂²l/∂u∂x."""
    lib = _abi.load_library()
    src = dynamics_src.encode() if isinstance(dynamics_src, str) else bytes(dynamics_src)
    _custom_sources.append(src)
    pr = np.ascontiguousarray(np.asarray(params, dtype=np.float64))
    p = Problem()
    rc = lib.ilqr_problem_custom(ctypes.byref(p), int(n), int(m), int(H), int(B), float(dt), src,
                                 pr.ctypes.data if pr.size else None, int(pr.size))
    if rc != 0:
        raise IlqrError("ilqr_problem_custom failed (n <= 16, m <= 8, <= 32 params)")
    p.custom_cost = 1 if user_cost else 0
    p.n_alpha = n_alpha
    p.trace_iters = trace_iters
    p.device = device
    if reg is not None:
        p.reg = reg
    for name, arr, cnt in (("x_target", x_target, n), ("w_x", w_x, n), ("w_u", w_u, m), ("w_xf", w_xf, n)):
        if arr is not None:
            arr = np.asarray(arr, dtype=np.float64)
            if arr.shape != (cnt,):
                raise ValueError("%s must have %d entries" % (name, cnt))
            for i in range(cnt):
                getattr(p, name)[i] = float(arr[i])
    return p


def custom_compile_check(dynamics_src, n, m, user_cost=False):
    """(ok, compiler log) for a snippet (dynamics; with user_cost also ilqr_cost / ilqr_final_cost); needs libnvrtc, no GPU."""
    lib = _abi.load_library()
    buf = ctypes.create_string_buffer(1 << 16)
    rc = lib.ilqr_custom_compile_check(dynamics_src.encode(), int(n), int(m), 1 if user_cost else 0, buf, len(buf))
    return rc == 0, buf.value.decode(errors="replace")


def load_urdf(path):
    """Mini URDF loader for what the reference feeds parse_urdf (test/urdf/*.urdf, RBD_helper_functions.jl:6-7):
    a serial chain of revolute/continuous joints.  Returns (joints[(nq, 20)], base_inertial) where base_inertial =
    (mass, com[3], inertia[6]) of the root link (irrelevant for a fixed base)."""
    import xml.etree.ElementTree as ET
    root = ET.parse(path).getroot()

    def vec(el, attr, default):
        return [float(t) for t in el.get(attr).split()] if el is not None and el.get(attr) else list(default)

    def inertial(link):
        ine = link.find("inertial")
        if ine is None:
            return 0.0, [0.0] * 3, [0.0] * 6
        mass = float(ine.find("mass").get("value"))
        com = vec(ine.find("origin"), "xyz", (0, 0, 0))
        if ine.find("origin") is not None and any(abs(v) > 0 for v in vec(ine.find("origin"), "rpy", (0, 0, 0))):
            raise ValueError("rotated <inertial> frames are not supported")
        I = ine.find("inertia")
        return mass, com, [float(I.get(k)) for k in ("ixx", "ixy", "ixz", "iyy", "iyz", "izz")]

    links = {l.get("name"): l for l in root.findall("link")}
    joints = [j for j in root.findall("joint")]
    children = {j.find("child").get("link") for j in joints}
    roots = [n for n in links if n not in children]
    if len(roots) != 1:
        raise ValueError("expected exactly one root link")
    rows, parent = [], roots[0]
    by_parent = {}
    for j in joints:
        by_parent.setdefault(j.find("parent").get("link"), []).append(j)
    while parent in by_parent:
        js = by_parent[parent]
        if len(js) != 1:
            raise ValueError("not a serial chain: link %s has %d children" % (parent, len(js)))
        j = js[0]
        if j.get("type") == "fixed":
            # massless frames welded to the chain (6Dof_arm.urdf's tool_frame) carry no dynamics
            if inertial(links[j.find("child").get("link")])[0] != 0.0 or j.find("child").get("link") in by_parent:
                raise ValueError("fixed joint %s: only massless leaf frames are supported" % j.get("name"))
            break
        if j.get("type") not in ("revolute", "continuous"):
            raise ValueError("joint %s: only revolute/continuous joints are supported" % j.get("name"))
        axis = np.array(vec(j.find("axis"), "xyz", (1, 0, 0)), dtype=np.float64)
        axis = axis / np.linalg.norm(axis)
        child = j.find("child").get("link")
        mass, com, I = inertial(links[child])
        rows.append(np.concatenate([vec(j.find("origin"), "xyz", (0, 0, 0)), vec(j.find("origin"), "rpy", (0, 0, 0)),
                                    axis, [mass], com, I, [0.0]]))
        parent = child
    return np.stack(rows), inertial(links[roots[0]])


def _f64(a):
    return np.asfortranarray(np.asarray(a, dtype=np.float64))


class BatchSolver:
    """One ilqr_handle: device-resident batch + the passes of the hot path."""

    def __init__(self, problem):
        self._lib = _abi.load_library()
        self.problem = problem
        self._h = ctypes.c_void_p()
        rc = self._lib.ilqr_create(ctypes.byref(problem), ctypes.byref(self._h))
        if rc != 0:
            raise IlqrError("ilqr_create: %s" % self._lib.ilqr_last_error(None).decode())
        self.n, self.m, self.H, self.B = problem.n, problem.m, problem.H, problem.B
        self.N = self.H + 1

    # -- plumbing ---------------------------------------------------------
    def _ck(self, rc, what):
        if rc != 0:
            raise IlqrError("%s failed (%d): %s" % (what, rc, self._lib.ilqr_last_error(self._h).decode()))

    def close(self):
        if self._h:
            self._lib.ilqr_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _shape(self, a, lead):
        a = _f64(a)
        if a.ndim == len(lead):
            a = a.reshape(tuple(lead) + (1,), order="F")
        assert a.shape == tuple(lead) + (self.B,), "expected %s, got %s" % (tuple(lead) + (self.B,), a.shape)
        return a

    # -- data movement ----------------------------------------------------
    def upload(self, x_init, u_init, x_traj=None):
        x = self._shape(x_init, (self.N, self.n))
        u = self._shape(u_init, (self.H, self.m))
        xt = None if x_traj is None else self._shape(x_traj, (self.N, self.n))
        self._ck(self._lib.ilqr_upload(self._h, x.ctypes.data, u.ctypes.data, None if xt is None else xt.ctypes.data),
                 "ilqr_upload")

    def upload_x0(self, x0, u_init, x_traj=None):
        x0 = self._shape(x0, (self.n,))
        u = self._shape(u_init, (self.H, self.m))
        xt = None if x_traj is None else self._shape(x_traj, (self.N, self.n))
        self._ck(self._lib.ilqr_upload_x0(self._h, x0.ctypes.data, u.ctypes.data,
                                          None if xt is None else xt.ctypes.data), "ilqr_upload_x0")

    def upload_device(self, d_x, d_u, d_xtraj=None):
        """d_* are raw device addresses (int) of boundary-layout fp64 arrays."""
        self._ck(self._lib.ilqr_upload_device(self._h, d_x, d_u, d_xtraj), "ilqr_upload_device")

    def upload_gains(self, duff, K):
        d = self._shape(duff, (self.H, self.m))
        k = self._shape(K, (self.H, self.m, self.n))
        self._ck(self._lib.ilqr_upload_gains(self._h, d.ctypes.data, k.ctypes.data), "ilqr_upload_gains")

    _SHAPES = {
        _abi.X: lambda s: (s.N, s.n, s.B), _abi.XBAR: lambda s: (s.N, s.n, s.B),
        _abi.U: lambda s: (s.H, s.m, s.B), _abi.UBAR: lambda s: (s.H, s.m, s.B),
        _abi.DUFF: lambda s: (s.H, s.m, s.B), _abi.K: lambda s: (s.H, s.m, s.n, s.B),
        _abi.NEW_COST: lambda s: (s.B,), _abi.PREV_COST: lambda s: (s.B,), _abi.ALPHA: lambda s: (s.B,),
        _abi.DU2: lambda s: (s.B,),
        _abi.COST_TRACE: lambda s: (s.problem.trace_iters, s.B), _abi.ALPHA_TRACE: lambda s: (s.problem.trace_iters, s.B),
        _abi.DU2_TRACE: lambda s: (s.problem.trace_iters, s.B),
        _abi.STATUS: lambda s: (s.B,), _abi.ITERS: lambda s: (s.B,), _abi.ACTIVE: lambda s: (s.B,),
    }

    def download(self, which):
        shape = self._SHAPES[which](self)
        dtype = np.int32 if which in (_abi.STATUS, _abi.ITERS, _abi.ACTIVE) else np.float64
        out = np.empty(shape, dtype=dtype, order="F")
        self._ck(self._lib.ilqr_download(self._h, which, out.ctypes.data), "ilqr_download")
        return out

    def download_device(self, which, d_dst):
        self._ck(self._lib.ilqr_download_device(self._h, which, d_dst), "ilqr_download_device")

    # -- the hot path -----------------------------------------------------
    def backward_pass(self):
        self._ck(self._lib.ilqr_backward_pass(self._h), "ilqr_backward_pass")

    def forward_pass(self, prev_cost=None):
        pc = None
        if prev_cost is not None:
            pc = np.ascontiguousarray(np.broadcast_to(np.asarray(prev_cost, dtype=np.float64), (self.B,)))
        self._ck(self._lib.ilqr_forward_pass(self._h, None if pc is None else pc.ctypes.data), "ilqr_forward_pass")

    def commit(self, tol):
        na = ctypes.c_int32()
        self._ck(self._lib.ilqr_commit(self._h, float(tol), ctypes.byref(na)), "ilqr_commit")
        return na.value

    def set_reg(self, reg):
        self._ck(self._lib.ilqr_set_reg(self._h, float(reg)), "ilqr_set_reg")

    def set_active(self, mask):
        m = np.ascontiguousarray(np.asarray(mask).astype(np.int32))
        assert m.shape == (self.B,)
        self._ck(self._lib.ilqr_set_active(self._h, m.ctypes.data), "ilqr_set_active")

    def iterate(self, tol):
        na = ctypes.c_int32()
        self._ck(self._lib.ilqr_iterate(self._h, float(tol), ctypes.byref(na)), "ilqr_iterate")
        return na.value

    def fit(self, max_iter=100, tol=1e-6):
        it = ctypes.c_int32()
        self._ck(self._lib.ilqr_fit(self._h, int(max_iter), float(tol), ctypes.byref(it)), "ilqr_fit")
        return it.value

    def solve_ptrs(self, x, u, max_iter, tol, xo, uo, cost=None, iters=None, status=None, x_traj=None):
        """ilqr_solve on raw HOST addresses (ints) of boundary-layout arrays (pinned for full-speed copies)."""
        self._ck(self._lib.ilqr_solve(self._h, x, u, x_traj, int(max_iter), float(tol), xo, uo, cost, iters, status), "ilqr_solve")

    def solve(self, x_init, u_init, x_traj=None, max_iter=100, tol=1e-6, out=None):
        """Host in → host out (ilqr_solve).  Returns dict(x,u,cost,iters,status)."""
        x = self._shape(x_init, (self.N, self.n))
        u = self._shape(u_init, (self.H, self.m))
        xt = None if x_traj is None else self._shape(x_traj, (self.N, self.n))
        if out is None:
            out = dict(x=np.empty_like(x), u=np.empty_like(u), cost=np.empty(self.B),
                       iters=np.empty(self.B, dtype=np.int32), status=np.empty(self.B, dtype=np.int32))
        self._ck(self._lib.ilqr_solve(self._h, x.ctypes.data, u.ctypes.data, None if xt is None else xt.ctypes.data,
                                      int(max_iter), float(tol), out["x"].ctypes.data, out["u"].ctypes.data,
                                      out["cost"].ctypes.data, out["iters"].ctypes.data, out["status"].ctypes.data),
                 "ilqr_solve")
        return out

    def stream_solve_device(self, n_total, d_x, d_u, d_xo, d_uo, d_cost=None, d_iters=None, d_status=None, max_iter=100, tol=1e-6):
        """ilqr_stream_solve_device: raw device addresses of boundary-layout arrays holding n_total trajectories; the
        handle's B slots are kept full by admitting pending trajectories as others finish.  Returns batch iterations."""
        it = ctypes.c_int64()
        self._ck(self._lib.ilqr_stream_solve_device(self._h, int(n_total), d_x, d_u, int(max_iter), float(tol), d_xo, d_uo,
                                                    d_cost, d_iters, d_status, ctypes.byref(it)), "ilqr_stream_solve_device")
        return it.value

    # -- receding-horizon MPC ----------------------------------------------
    def mpc_start(self, x0, u_init=None):
        x0 = self._shape(x0, (self.n,))
        u = None if u_init is None else self._shape(u_init, (self.H, self.m))
        self._ck(self._lib.ilqr_mpc_start(self._h, x0.ctypes.data, None if u is None else u.ctypes.data), "ilqr_mpc_start")

    def mpc_step(self, max_iter=5, tol=1e-6):
        """One closed-loop step: solve (≤ max_iter iterations, warm-started), apply u[0] to the plant, shift.
        Returns (u_applied[m,B], x_plant[n,B])."""
        ua = np.empty((self.m, self.B), order="F"); xp = np.empty((self.n, self.B), order="F")
        self._ck(self._lib.ilqr_mpc_step(self._h, int(max_iter), float(tol), ua.ctypes.data, xp.ctypes.data), "ilqr_mpc_step")
        return ua, xp

    # -- introspection ----------------------------------------------------
    def launch_count(self):
        return int(self._lib.ilqr_launch_count(self._h))

    def last_kernel_ms(self):
        b, f = ctypes.c_float(), ctypes.c_float()
        self._ck(self._lib.ilqr_last_kernel_ms(self._h, ctypes.byref(b), ctypes.byref(f)), "ilqr_last_kernel_ms")
        return b.value, f.value

    def profile(self):
        """dict of the cumulative device-time profile since the last upload (ilqr_profile)."""
        out = (ctypes.c_double * 8)()
        self._ck(self._lib.ilqr_profile(self._h, out), "ilqr_profile")
        return dict(bwd_ms=out[0], fwd_ms=out[1], bwd_launches=int(out[2]), fwd_launches=int(out[3]),
                    traj_iters=out[4], first_bwd_ms=out[5], first_fwd_ms=out[6])

    def stream_profile(self):
        """Last streaming solve (ilqr_stream_profile): device ms, rounds launched, rounds at completion, trajectories."""
        out = (ctypes.c_double * 4)()
        self._ck(self._lib.ilqr_stream_profile(self._h, out), "ilqr_stream_profile")
        return dict(ms=out[0], rounds_launched=int(out[1]), rounds=int(out[2]), trajectories=int(out[3]))

    def stream_ptr(self):
        return int(self._lib.ilqr_stream(self._h) or 0)

    def set_tuning(self, split_below=-1, coop_below=-1, fwd_split_above=-1, compaction=-1):
        """ilqr_set_tuning: kernel-selection thresholds of the batch path (negative = unchanged)."""
        self._ck(self._lib.ilqr_set_tuning(self._h, int(split_below), int(coop_below), int(fwd_split_above), int(compaction)),
                 "ilqr_set_tuning")

    def set_variant(self, v):
        self._ck(self._lib.ilqr_set_variant(self._h, v), "ilqr_set_variant")

    def sync(self):
        self._ck(self._lib.ilqr_sync(self._h), "ilqr_sync")


class SolverPool:
    """ilqr_pool: several batches in flight on one GPU (one handle + worker thread each).
    submit*() return tickets without blocking; the caller keeps the buffers alive until wait()."""

    def __init__(self, problem, n_handles=4):
        self._lib = _abi.load_library()
        self.problem = problem
        self._p = ctypes.c_void_p()
        rc = self._lib.ilqr_pool_create(ctypes.byref(problem), int(n_handles), ctypes.byref(self._p))
        if rc != 0:
            raise IlqrError("ilqr_pool_create: %s" % self._lib.ilqr_pool_last_error(None).decode())
        self.n_handles = n_handles

    def close(self):
        if self._p:
            self._lib.ilqr_pool_destroy(self._p)
            self._p = ctypes.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def submit_ptrs(self, x, u, xt, max_iter, tol, xo, uo, cost=None, iters=None, status=None, device=False):
        """Raw addresses (ints): host pointers, or device pointers with device=True."""
        fn = self._lib.ilqr_pool_submit_device if device else self._lib.ilqr_pool_submit
        t = fn(self._p, x, u, xt, int(max_iter), float(tol), xo, uo, cost, iters, status)
        if t < 0:
            raise IlqrError("ilqr_pool_submit failed (%d)" % t)
        return t

    def submit(self, x_init, u_init, out, x_traj=None, max_iter=100, tol=1e-6):
        """NumPy (Fortran-ordered, boundary layout) in, preallocated `out` dict (x,u,cost,iters,status) filled on wait."""
        xt = None if x_traj is None else x_traj.ctypes.data
        return self.submit_ptrs(x_init.ctypes.data, u_init.ctypes.data, xt, max_iter, tol, out["x"].ctypes.data,
                                out["u"].ctypes.data, out["cost"].ctypes.data, out["iters"].ctypes.data,
                                out["status"].ctypes.data)

    def wait(self, ticket):
        rc = self._lib.ilqr_pool_wait(self._p, ticket)
        if rc != 0:
            raise IlqrError("pooled solve failed (%d): %s" % (rc, self._lib.ilqr_pool_last_error(self._p).decode()))

    def wait_all(self):
        rc = self._lib.ilqr_pool_wait_all(self._p)
        if rc != 0:
            raise IlqrError("pooled solve failed (%d): %s" % (rc, self._lib.ilqr_pool_last_error(self._p).decode()))

    def launch_count(self):
        return int(self._lib.ilqr_pool_launch_count(self._p))


class Streamer:
    """ilqr_streamer: continuous batching — over the fused rounds for the 2-link model, over the streaming-admission loop
    (one submitted batch at a time) for the rigid-body and NVRTC models.  `problem.B` is the number of SLOTS
    (size it to the machine: 148 SMs x 12 warps x 32 = 56,832 on B200), `batch_size` the number of trajectories in
    every submitted batch; up to `ring` batches are in flight (submit blocks while the ring is full)."""

    def __init__(self, problem, batch_size, ring=8, max_iter=100, tol=1e-6):
        self._lib = _abi.load_library()
        self.problem, self.batch_size, self.ring = problem, int(batch_size), int(ring)
        self._p = ctypes.c_void_p()
        rc = self._lib.ilqr_streamer_create(ctypes.byref(problem), int(batch_size), int(ring), int(max_iter), float(tol),
                                            ctypes.byref(self._p))
        if rc != 0:
            raise IlqrError("ilqr_streamer_create (%d): %s" % (rc, self._lib.ilqr_streamer_last_error(None).decode()))

    def close(self):
        if self._p:
            self._lib.ilqr_streamer_destroy(self._p)
            self._p = ctypes.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def submit_ptrs(self, x, u, xo=None, uo=None, cost=None, iters=None, status=None, device=False, x0=False, x_traj=None):
        """Raw addresses (ints) of boundary-layout arrays of batch_size trajectories: host pointers, or device pointers
        with device=True.  x0=True: `x` is x0[n,Bb] and x_init is rolled out on the device (`u` may be None = zeros).
        Every output is optional (None = not produced, not copied back).  Returns a ticket."""
        if x_traj is not None:     # fit's keyword argument (src/forward_pass.jl:151)
            if x0:
                raise ValueError("x_traj goes with x_init submissions")
            fn = self._lib.ilqr_streamer_submit_traj_device if device else self._lib.ilqr_streamer_submit_traj
            t = fn(self._p, x, u, x_traj, xo, uo, cost, iters, status)
        else:
            fn = {(False, False): self._lib.ilqr_streamer_submit, (True, False): self._lib.ilqr_streamer_submit_device,
                  (False, True): self._lib.ilqr_streamer_submit_x0, (True, True): self._lib.ilqr_streamer_submit_x0_device}[(bool(device), bool(x0))]
            t = fn(self._p, x, u, xo, uo, cost, iters, status)
        if t < 0:
            raise IlqrError("ilqr_streamer_submit failed (%d): %s" % (t, self._lib.ilqr_streamer_last_error(self._p).decode()))
        return t

    def _check(self, a, lead, dtype, name):
        """The C ABI takes raw pointers: a wrong dtype, a C-ordered slice or a wrong shape would be misread silently."""
        if a is None:
            return None
        if not isinstance(a, np.ndarray) or a.dtype != dtype:
            raise TypeError("%s: expected a NumPy array of %s" % (name, np.dtype(dtype)))
        shape = tuple(lead) + (self.batch_size,)
        if a.shape != shape:
            raise ValueError("%s: expected shape %s, got %s" % (name, shape, a.shape))
        if not a.flags.f_contiguous:
            raise ValueError("%s: expected a Fortran-ordered (Julia column-major) contiguous array; use np.asfortranarray" % name)
        return a.ctypes.data

    def _outs(self, out):
        p = self.problem
        return (self._check(out.get("x"), (p.H + 1, p.n), np.float64, "out['x']"), self._check(out.get("u"), (p.H, p.m), np.float64, "out['u']"),
                self._check(out.get("cost"), (), np.float64, "out['cost']"), self._check(out.get("iters"), (), np.int32, "out['iters']"),
                self._check(out.get("status"), (), np.int32, "out['status']"))

    def submit(self, x_init, u_init, out, x_traj=None):
        """NumPy (Fortran-ordered, boundary layout) in, preallocated `out` dict (any of x, u, cost, iters, status) filled on wait.
        x_traj: fit's keyword argument (src/forward_pass.jl:151), same shape as x_init."""
        p = self.problem
        return self.submit_ptrs(self._check(x_init, (p.H + 1, p.n), np.float64, "x_init"), self._check(u_init, (p.H, p.m), np.float64, "u_init"),
                                *self._outs(out), x_traj=self._check(x_traj, (p.H + 1, p.n), np.float64, "x_traj"))

    def submit_x0(self, x0, u_init, out):
        """x0[n,Bb] (+ u_init[H,m,Bb] or None = zeros): x_init is the open-loop rollout, computed on the device
        (animate_2_link.jl:11-16).  `out` as in submit()."""
        p = self.problem
        return self.submit_ptrs(self._check(x0, (p.n,), np.float64, "x0"), self._check(u_init, (p.H, p.m), np.float64, "u_init"),
                                *self._outs(out), x0=True)

    def wait(self, ticket):
        rc = self._lib.ilqr_streamer_wait(self._p, ticket)
        if rc != 0:
            raise IlqrError("streamed solve failed (%d): %s" % (rc, self._lib.ilqr_streamer_last_error(self._p).decode()))

    def wait_all(self):
        rc = self._lib.ilqr_streamer_wait_all(self._p)
        if rc != 0:
            raise IlqrError("streamed solve failed (%d): %s" % (rc, self._lib.ilqr_streamer_last_error(self._p).decode()))

    def launch_count(self):
        return int(self._lib.ilqr_streamer_launch_count(self._p))

    def rounds(self):
        return int(self._lib.ilqr_streamer_rounds(self._p))

    def profile(self):
        """Cumulative: device ms over the timed rounds, rounds covered, batches completed, rounds launched."""
        out = (ctypes.c_double * 4)()
        self._lib.ilqr_streamer_profile(self._p, out)
        return dict(round_ms=out[0], rounds_timed=int(out[1]), batches_completed=int(out[2]), rounds_launched=int(out[3]))


def _batch_of(x):
    x = _f64(x)
    return (x.shape[2] if x.ndim == 3 else 1), x.ndim == 2


def _problem_for(problem, H, B):
    p = Problem.from_buffer_copy(problem)
    p.H, p.B = H, B
    return p


def backward_pass(x, u, problem):
    """backward_pass(x, u, dynamicsf, immediate_cost, final_cost) → (δuff, K)   src/backward_pass.jl:324-357"""
    x = _f64(x); u = _f64(u)
    N, M = x.shape[0], u.shape[0]
    assert N == M + 1                                   # src/backward_pass.jl:329
    B, single = _batch_of(x)
    with BatchSolver(_problem_for(problem, M, B)) as s:
        s.upload(x, u)
        s.backward_pass()
        duff, K, st = s.download(_abi.DUFF), s.download(_abi.K), s.download(_abi.STATUS)
    assert not np.any(st & _abi.STATUS_NAN_GAINS)       # :353-354
    return (duff[..., 0], K[..., 0]) if single else (duff, K)


def forward_pass(x, u, x_traj, duff, K, prev_cost, problem):
    """forward_pass(x, u, x_traj, δuff, K, prev_cost, …) → (x̄, ū, new_cost)   src/forward_pass.jl:55-93

    The reference's `while true` halving is bounded at problem.n_alpha candidates;
    exhausting them raises (the reference would loop forever)."""
    x = _f64(x); u = _f64(u)
    N, M = x.shape[0], u.shape[0]
    assert N == M + 1                                   # src/forward_pass.jl:62
    B, single = _batch_of(x)
    with BatchSolver(_problem_for(problem, M, B)) as s:
        s.upload(x, u, x_traj)
        s.upload_gains(duff, K)
        s.forward_pass(prev_cost)
        xb, ub = s.download(_abi.XBAR), s.download(_abi.UBAR)
        cost, alpha, st = s.download(_abi.NEW_COST), s.download(_abi.ALPHA), s.download(_abi.STATUS)
    if np.any(alpha == 0.0):
        raise IlqrError("line search exhausted n_alpha candidates (reference: infinite loop, src/forward_pass.jl:70)")
    assert not np.any(st & _abi.STATUS_NAN_ROLLOUT)     # :89-90
    return (xb[..., 0], ub[..., 0], float(cost[0])) if single else (xb, ub, cost)


def fit(x_init, u_init, problem, x_traj=None, max_iter=100, tol=1e-6, info=None):
    """fit(x_init, u_init, dynamicsf, immediate_cost, final_cost; x_traj, max_iter, tol) → (x̄, ū)
    src/forward_pass.jl:148-179.  `info` (optional dict) receives cost/iters/status per trajectory."""
    x = _f64(x_init); u = _f64(u_init)
    N, M = x.shape[0], u.shape[0]
    assert N == M + 1, "size(x_init)[2] == size(u_init)[1], (# of states is 1 more than # of inputs in trajectory)"
    B, single = _batch_of(x)
    with BatchSolver(_problem_for(problem, M, B)) as s:
        out = s.solve(x, u, x_traj, max_iter=max_iter, tol=tol)
    st = out["status"]
    assert not np.any(st & _abi.STATUS_NOT_DECREASED)   # src/forward_pass.jl:168
    assert not np.any(st & _abi.STATUS_NAN_GAINS)       # src/backward_pass.jl:353-354
    assert not np.any(st & _abi.STATUS_NAN_ROLLOUT)     # src/forward_pass.jl:89-90
    if np.any(st & _abi.STATUS_LS_EXHAUSTED):
        raise IlqrError("line search exhausted n_alpha candidates on %d trajectories (reference: infinite loop, "
                        "src/forward_pass.jl:70)" % int(np.count_nonzero(st & _abi.STATUS_LS_EXHAUSTED)))
    if info is not None:
        info.update(cost=out["cost"], iters=out["iters"], status=out["status"])
    return (out["x"][..., 0], out["u"][..., 0]) if single else (out["x"], out["u"])
