"""Independent NumPy restatement of the serial-chain rigid-body plugin (test infrastructure).

Deliberately NOT the algorithm of oracle/serial_chain.hpp (spatial-vector RNEA + CRBA) nor of the
CUDA kernels (3-vector Newton–Euler in link frames): this one is the textbook Lagrangian form
  M(q) = Σ_i m_i J_vᵢᵀ J_vᵢ + J_ωᵢᵀ (R_i I_i R_iᵀ) J_ωᵢ            (link COM Jacobians, world frame)
  bias_k = Σ_ij Γ_kij q̇_i q̇_j + ∂V/∂q_k,  Γ_kij = ½(∂M_kj/∂q_i + ∂M_ki/∂q_j − ∂M_ij/∂q_k)
with ∂M/∂q and ∂V/∂q from complex-step differentiation (exact to rounding).  Agreement of the three
formulations pins the dynamics that RigidBodyDynamics.jl's mass_matrix / dynamics_bias compute
(test/RBD_2_link_example/RBD_helper_functions.jl:57-66).
"""
import numpy as np

STRIDE = 20


def joint_row(xyz=(0, 0, 0), rpy=(0, 0, 0), axis=(0, 0, 1), mass=1.0, com=(0, 0, 0), inertia=(1, 0, 0, 1, 0, 1), prismatic=False):
    """One row of the flat chain description: joint origin/axis + child link inertial.  The last slot (a pad in
    the product ABI) marks a prismatic joint — used only by this file, to model a floating base as a virtual
    chain of three prismatic and three revolute joints."""
    axis = np.asarray(axis, dtype=np.float64)
    axis = axis / np.linalg.norm(axis)
    return np.concatenate([xyz, rpy, axis, [mass], com, inertia, [1.0 if prismatic else 0.0]]).astype(np.float64)


def seven_dof_chain():
    """BASELINE config 4: 7 revolute joints in the test/urdf/6Dof_arm.urdf pattern —
    axes z,y,z,y,z,y,z, origin (1,0,0), mass 3, inertia 0.5·I, COM at the link frame, zero gravity."""
    rows = []
    for i in range(7):
        rows.append(joint_row(xyz=(1, 0, 0), axis=(0, 0, 1) if i % 2 == 0 else (0, 1, 0), mass=3.0,
                              inertia=(0.5, 0, 0, 0.5, 0, 0.5)))
    return np.stack(rows)


def random_chain(nq, rng, general=True):
    """A chain that exercises every term: rpy offsets, skew axes, off-origin COMs, full inertia tensors."""
    rows = []
    for i in range(nq):
        A = rng.normal(size=(3, 3))
        I = A @ A.T + 0.5 * np.eye(3)            # SPD inertia about the COM
        inertia = (I[0, 0], I[0, 1], I[0, 2], I[1, 1], I[1, 2], I[2, 2])
        if general:
            axis = rng.normal(size=3)
            rpy = rng.uniform(-1, 1, 3)
        else:
            axis = np.eye(3)[rng.integers(0, 3)]
            rpy = np.zeros(3)
        rows.append(joint_row(xyz=rng.uniform(-1, 1, 3), rpy=rpy, axis=axis, mass=rng.uniform(0.5, 3.0),
                              com=rng.uniform(-0.5, 0.5, 3), inertia=inertia))
    return np.stack(rows)


def _rot_rpy(rpy):
    r, p, y = rpy
    Rx = np.array([[1, 0, 0], [0, np.cos(r), -np.sin(r)], [0, np.sin(r), np.cos(r)]])
    Ry = np.array([[np.cos(p), 0, np.sin(p)], [0, 1, 0], [-np.sin(p), 0, np.cos(p)]])
    Rz = np.array([[np.cos(y), -np.sin(y), 0], [np.sin(y), np.cos(y), 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def _skew(a):
    return np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])


def _rot_axis(a, q):
    K = _skew(a)
    return np.eye(3) + np.sin(q) * K + (1 - np.cos(q)) * (K @ K)


def _kinematics(joints, q):
    """World pose of every link (complex-safe: no conj, no abs)."""
    nq = joints.shape[0]
    R = np.eye(3, dtype=q.dtype); o = np.zeros(3, dtype=q.dtype)
    Rs, os_, zs = [], [], []
    for i in range(nq):
        xyz, rpy, axis = joints[i, 0:3], joints[i, 3:6], joints[i, 6:9]
        if joints[i, 19] == 1.0:      # prismatic: slide along the axis of the (fixed-rotation) joint frame
            R = R @ _rot_rpy(rpy)
            o = o + R @ (np.linalg.solve(_rot_rpy(rpy), xyz) + axis * q[i])
        else:
            o = o + R @ xyz
            R = R @ _rot_rpy(rpy) @ _rot_axis(axis, q[i])
        Rs.append(R); os_.append(o); zs.append(R @ axis)
    return Rs, os_, zs


def mass_matrix(joints, q):
    q = np.asarray(q)
    nq = joints.shape[0]
    Rs, os_, zs = _kinematics(joints, q)
    M = np.zeros((nq, nq), dtype=q.dtype)
    for i in range(nq):
        m, com, I6 = joints[i, 9], joints[i, 10:13], joints[i, 13:19]
        Ic = np.array([[I6[0], I6[1], I6[2]], [I6[1], I6[3], I6[4]], [I6[2], I6[4], I6[5]]])
        pc = os_[i] + Rs[i] @ com
        Jv = np.zeros((3, nq), dtype=q.dtype); Jw = np.zeros((3, nq), dtype=q.dtype)
        for j in range(i + 1):
            if joints[j, 19] == 1.0:
                Jv[:, j] = zs[j]
            else:
                Jv[:, j] = np.cross(zs[j], pc - os_[j]); Jw[:, j] = zs[j]
        M = M + m * (Jv.T @ Jv) + Jw.T @ (Rs[i] @ Ic @ Rs[i].T) @ Jw
    return M


def potential(joints, q, gravity):
    q = np.asarray(q)
    Rs, os_, _ = _kinematics(joints, q)
    V = 0.0
    for i in range(joints.shape[0]):
        pc = os_[i] + Rs[i] @ joints[i, 10:13]
        V = V - joints[i, 9] * (np.asarray(gravity) @ pc)
    return V


def bias(joints, q, qd, gravity=(0, 0, 0)):
    nq = joints.shape[0]
    h = 1e-30
    dM = np.zeros((nq, nq, nq)); dV = np.zeros(nq)       # dM[:,:,k] = ∂M/∂q_k
    for k in range(nq):
        qc = np.asarray(q, dtype=np.complex128).copy(); qc[k] += 1j * h
        dM[:, :, k] = mass_matrix(joints, qc).imag / h
        dV[k] = np.imag(potential(joints, qc, gravity)) / h
    c = np.zeros(nq)
    for k in range(nq):
        acc = 0.0
        for i in range(nq):
            for j in range(nq):
                acc += 0.5 * (dM[k, j, i] + dM[k, i, j] - dM[i, j, k]) * qd[i] * qd[j]
        c[k] = acc + dV[k]
    return c


def continuous_dynamics(joints, x, u, gravity=(0, 0, 0)):
    nq = joints.shape[0]
    q, qd = x[:nq], x[nq:]
    M = mass_matrix(joints, np.asarray(q, dtype=np.float64))
    vdot = np.linalg.solve(M, u - bias(joints, q, qd, gravity))
    return np.concatenate([qd, vdot])


def dynamics(joints, x, u, gravity=(0, 0, 0), dt=0.01):
    f = lambda xx: continuous_dynamics(joints, xx, u, gravity)
    k1 = dt * f(x); k2 = dt * f(x + k1 / 2); k3 = dt * f(x + k2 / 2); k4 = dt * f(x + k3)
    return x + (1.0 / 6.0) * (k1 + 2 * k2 + 2 * k3 + k4)


# ---------------------------------------------------------------------------------------------
# Floating base as a virtual chain: three prismatic joints (world x, y, z) then three revolute joints
# (x, y, z; R = Rx·Ry·Rz), the last one carrying the base link; the arm follows.  At zero base angles the
# body twist equals the virtual joint rates, so the body-frame equations the reference integrates
# (RBD_helper_functions.jl:57-66: v̇ = M \ (u − bias) with 𝑣 = [ω; v; θ̇] in the base frame) follow from the
# Lagrangian equations of the virtual chain by  v̇_body = tw(q, q̈) + ∂tw/∂q·q̇  (tw = body twist as a function of
# the virtual coordinates and rates).  Zero gravity ⇒ the body-frame dynamics do not depend on the base pose,
# so checking at zero base angles loses no generality.
# ---------------------------------------------------------------------------------------------
def virtual_floating_chain(base_row, joints):
    rows = []
    for k in range(3):
        rows.append(joint_row(axis=np.eye(3)[k], mass=0.0, inertia=(0,) * 6, prismatic=True))
    for k in range(2):
        rows.append(joint_row(axis=np.eye(3)[k], mass=0.0, inertia=(0,) * 6))
    last = np.array(base_row, dtype=np.float64).copy()
    last[0:9] = joint_row(axis=(0, 0, 1))[0:9]; last[19] = 0.0
    rows.append(last)
    return np.concatenate([np.stack(rows), np.asarray(joints, dtype=np.float64)])


def _body_twist(qv, rates):
    """[ω_body; v_body] of the base from the six virtual coordinates' values (only the angles matter) and rates."""
    a = qv[3:6]
    Rx = np.array([[1, 0, 0], [0, np.cos(a[0]), -np.sin(a[0])], [0, np.sin(a[0]), np.cos(a[0])]])
    Ry = np.array([[np.cos(a[1]), 0, np.sin(a[1])], [0, 1, 0], [-np.sin(a[1]), 0, np.cos(a[1])]])
    Rz = np.array([[np.cos(a[2]), -np.sin(a[2]), 0], [np.sin(a[2]), np.cos(a[2]), 0], [0, 0, 1]])
    R = Rx @ Ry @ Rz
    w_world = np.array([1, 0, 0]) * rates[3] + Rx @ np.array([0, 1, 0]) * rates[4] + Rx @ Ry @ np.array([0, 0, 1]) * rates[5]
    return np.concatenate([R.T @ w_world, R.T @ rates[0:3]])


def floating_body_acceleration(base_row, joints, theta, vel, u):
    """v̇ = [ω̇; v̇; θ̈] (coordinate derivatives of the body-frame velocity vector) for 𝑣 = vel, generalised force u."""
    nq = len(theta)
    chain = virtual_floating_chain(base_row, joints)
    q = np.concatenate([np.zeros(6), theta])
    qd = np.concatenate([vel[3:6], vel[0:3], vel[6:]])        # zero base angles: ṙ = v_body, angle rates = ω_body
    tau = np.concatenate([u[3:6], u[0:3], u[6:]])
    M = mass_matrix(chain, q)
    qdd = np.linalg.solve(M, tau - bias(chain, q, qd))
    h = 1e-30
    dtw = np.imag(_body_twist(q[:6].astype(np.complex128) + 1j * h * qd[:6], qd[:6])) / h      # ∂tw/∂q · q̇
    acc6 = _body_twist(q[:6], qdd[:6]) + dtw
    return np.concatenate([acc6, qdd[6:]]), M
