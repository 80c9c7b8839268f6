# iLQRB200.jl — the reference-side binding a maintainer of aabouman/iLQR.jl would add.
# Thin `ccall` stubs over include/ilqr_b200.h (libilqr_b200.so).  Julia keeps problem setup, the
# outer loop, the convergence test and the asserts of `fit` (src/forward_pass.jl:148-179); the GPU
# runs `backward_pass` (src/backward_pass.jl:324-357) and `forward_pass` (src/forward_pass.jl:55-93)
# on a whole batch of problems.  NOTE: Julia is not available in the build image, so this file is
# kept deliberately thin and has not been executed; tests/ drive the identical ABI through ctypes.
module iLQRB200

const lib = get(ENV, "ILQR_B200_LIB", "libilqr_b200.so")

# mirrors `struct ilqr_problem` field for field
struct Problem
    abi_version::Int32
    model_id::Int32
    n::Int32
    m::Int32
    H::Int32
    B::Int32
    n_alpha::Int32
    trace_iters::Int32
    device::Int32
    variant::Int32
    dt::Float64
    reg::Float64
    model_params::NTuple{32,Float64}
    x_target::NTuple{16,Float64}
    w_x::NTuple{16,Float64}
    w_u::NTuple{8,Float64}
    w_xf::NTuple{16,Float64}
    nq::Int32
    custom_cost::Int32             # ILQR_MODEL_CUSTOM: 1 = the snippet also defines ilqr_cost / ilqr_final_cost
    gravity::NTuple{3,Float64}
    chain::NTuple{180,Float64}     # (ILQR_MAX_JOINTS + 1) × ILQR_CHAIN_STRIDE
    custom_src::Cstring            # ILQR_MODEL_CUSTOM: CUDA C++ source of ilqr_dynamics
end

const X, U, XBAR, UBAR, DUFF, K, NEW_COST, PREV_COST, ALPHA, DU2 = Int32.(0:9)
const STATUS, ITERS, ACTIVE = Int32(13), Int32(14), Int32(15)

check(rc, h) = rc == 0 || error(unsafe_string(ccall((:ilqr_last_error, lib), Cstring, (Ptr{Cvoid},), h)))

"The 2-link plugin of test/2_link_example/2_link_helper_functions.jl as a device-side problem."
function two_link_problem(H::Integer, B::Integer)
    p = Ref{Problem}()
    ccall((:ilqr_problem_two_link, lib), Int32, (Ptr{Problem}, Int32, Int32), p, H, B) == 0 || error("problem")
    return p[]
end

"""
The rigid-body plugin of test/RBD_2_link_example/RBD_helper_functions.jl for a fixed-base serial chain.
`joints` is 20 × nq (column per joint: origin xyz, rpy, unit axis, mass, COM, ixx ixy ixz iyy iyz izz, pad).
Cost weights start at zero: rebuild the immutable struct with `Setfield`/`@set` or fill them before `Solver`.
"""
function serial_chain_problem(joints::Matrix{Float64}, gravity::Vector{Float64}, H::Integer, B::Integer)
    p = Ref{Problem}()
    ccall((:ilqr_problem_serial_chain, lib), Int32, (Ptr{Problem}, Int32, Ptr{Float64}, Ptr{Float64}, Int32, Int32),
          p, size(joints, 2), joints, gravity, H, B) == 0 || error("problem")
    return p[]
end

"""
Any `dynamicsf` (what passing a Julia function does in the reference, src/forward_pass.jl:148-153): `src` is CUDA C++
defining `template <class T> __device__ void ilqr_dynamics(const T* x, const T* u, const double* p, T* xdot)`; it is
compiled by NVRTC inside `Solver(p)`.  Keep `src` alive until the Solver exists (the struct holds a raw pointer).
"""
function custom_problem(src::String, n::Integer, m::Integer, H::Integer, B::Integer, dt::Float64, params::Vector{Float64})
    p = Ref{Problem}()
    ccall((:ilqr_problem_custom, lib), Int32,
          (Ptr{Problem}, Int32, Int32, Int32, Int32, Float64, Cstring, Ptr{Float64}, Int32),
          p, n, m, H, B, dt, src, params, length(params)) == 0 || error("problem")
    return p[]
end

mutable struct Solver
    h::Ptr{Cvoid}
    p::Problem
    function Solver(p::Problem)
        h = Ref{Ptr{Cvoid}}(C_NULL)
        rc = ccall((:ilqr_create, lib), Int32, (Ptr{Problem}, Ptr{Ptr{Cvoid}}), Ref(p), h)
        rc == 0 || error(unsafe_string(ccall((:ilqr_last_error, lib), Cstring, (Ptr{Cvoid},), C_NULL)))
        s = new(h[], p)
        finalizer(s -> ccall((:ilqr_destroy, lib), Int32, (Ptr{Cvoid},), s.h), s)
        return s
    end
end

upload!(s::Solver, x::Array{Float64,3}, u::Array{Float64,3}, x_traj = nothing) =
    check(ccall((:ilqr_upload, lib), Int32, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                s.h, x, u, x_traj === nothing ? C_NULL : x_traj), s.h)

"""
Problem-setup helper (animate_2_link.jl:11-16): `x_init` = open-loop rollout of `u_init` from `x0[n,B]`, on the device.
"""
upload_x0!(s::Solver, x0::Matrix{Float64}, u::Array{Float64,3}, x_traj = nothing) =
    check(ccall((:ilqr_upload_x0, lib), Int32, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                s.h, x0, u, x_traj === nothing ? C_NULL : x_traj), s.h)

"Gains computed elsewhere (δuff[H,m,B], K[H,m,n,B]) for `forward_pass` with the reference's own argument list."
upload_gains!(s::Solver, duff::Array{Float64,3}, K::Array{Float64,4}) =
    check(ccall((:ilqr_upload_gains, lib), Int32, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), s.h, duff, K), s.h)

# ---- host-owned regularisation and convergence control (north_star: Julia owns both) ----------------------------
"The constant added to diag(H) before the gain solve (src/backward_pass.jl:214 hard-codes 0.01); from the next backward pass on."
set_reg!(s::Solver, reg::Float64) = check(ccall((:ilqr_set_reg, lib), Int32, (Ptr{Cvoid}, Float64), s.h, reg), s.h)

"Trajectories whose mask entry is 0 stop iterating (they keep their current iterate): the host's own convergence test."
set_active!(s::Solver, active::Vector{Int32}) =
    check(ccall((:ilqr_set_active, lib), Int32, (Ptr{Cvoid}, Ptr{Int32}), s.h, active), s.h)

"The tail of fit's loop body on the device (src/forward_pass.jl:168-175) → number of trajectories still iterating."
function commit!(s::Solver, tol::Float64)
    n_active = Ref{Int32}(0)
    check(ccall((:ilqr_commit, lib), Int32, (Ptr{Cvoid}, Float64, Ptr{Int32}), s.h, tol, n_active), s.h)
    return n_active[]
end

"backward + forward + commit in one call → number of trajectories still iterating."
function iterate!(s::Solver, tol::Float64)
    n_active = Ref{Int32}(0)
    check(ccall((:ilqr_iterate, lib), Int32, (Ptr{Cvoid}, Float64, Ptr{Int32}), s.h, tol, n_active), s.h)
    return n_active[]
end

"`ilqr_fit`: the whole loop on the library's side (per-trajectory convergence) → batch iterations executed."
function fit!(s::Solver; max_iter::Integer = 100, tol::Float64 = 1e-6)
    iters = Ref{Int32}(0)
    check(ccall((:ilqr_fit, lib), Int32, (Ptr{Cvoid}, Int32, Float64, Ptr{Int32}), s.h, max_iter, tol, iters), s.h)
    return iters[]
end

sync(s::Solver) = check(ccall((:ilqr_sync, lib), Int32, (Ptr{Cvoid},), s.h), s.h)

function download(s::Solver, which::Int32, dims...; T = Float64)
    out = Array{T}(undef, dims...)
    check(ccall((:ilqr_download, lib), Int32, (Ptr{Cvoid}, Int32, Ptr{Cvoid}), s.h, which, out), s.h)
    return out
end

"backward_pass(x, u, …) → (δuff[H,m,B], K[H,m,n,B])   (src/backward_pass.jl:324)"
function backward_pass(s::Solver)
    check(ccall((:ilqr_backward_pass, lib), Int32, (Ptr{Cvoid},), s.h), s.h)
    p = s.p
    st = download(s, STATUS, p.B; T = Int32)
    @assert !any(st .& 1 .!= 0)          # src/backward_pass.jl:353-354
    return (download(s, DUFF, p.H, p.m, p.B), download(s, K, p.H, p.m, p.n, p.B))
end

"forward_pass(x, u, x_traj, δuff, K, prev_cost, …) → (x̄, ū, new_cost)   (src/forward_pass.jl:55)"
function forward_pass(s::Solver, prev_cost::Vector{Float64})
    check(ccall((:ilqr_forward_pass, lib), Int32, (Ptr{Cvoid}, Ptr{Float64}), s.h, prev_cost), s.h)
    p = s.p
    st = download(s, STATUS, p.B; T = Int32)
    @assert !any(st .& 2 .!= 0)          # src/forward_pass.jl:89-90
    return (download(s, XBAR, p.H + 1, p.n, p.B), download(s, UBAR, p.H, p.m, p.B), download(s, NEW_COST, p.B))
end

"""
    fit(x_init[N,n,B], u_init[H,m,B], problem; x_traj, max_iter = 100, tol = 1e-6) → (x̄, ū)

Batched `iLQR.fit` (src/forward_pass.jl:148-179).  The loop, the convergence test
`sum((ū⁺ - ū).^2) <= tol` (per trajectory) and the monotone-cost assert stay here on the host;
`ilqr_commit` applies the decision on the device (converged trajectories keep the iterate from
BEFORE the last forward pass, exactly as the reference's `break` before the update does).
"""
function fit(x_init::Array{Float64,3}, u_init::Array{Float64,3}, p::Problem;
             x_traj = nothing, max_iter::Int64 = 100, tol::Float64 = 1e-6)
    N, n, B = size(x_init); M, m, _ = size(u_init)
    @assert(N == M + 1, "size(x_init)[2] == size(u_init)[1], (# of states is 1 more than # of inputs in trajectory)")
    s = Solver(p)
    upload!(s, x_init, u_init, x_traj)
    n_active = Ref{Int32}(B)
    for iter = 1:max_iter
        check(ccall((:ilqr_backward_pass, lib), Int32, (Ptr{Cvoid},), s.h), s.h)
        check(ccall((:ilqr_forward_pass, lib), Int32, (Ptr{Cvoid}, Ptr{Float64}), s.h, C_NULL), s.h)
        check(ccall((:ilqr_commit, lib), Int32, (Ptr{Cvoid}, Float64, Ptr{Int32}), s.h, tol, n_active), s.h)
        n_active[] == 0 && break
    end
    st = download(s, STATUS, B; T = Int32)
    @assert !any(st .& 8 .!= 0)          # src/forward_pass.jl:168  prev_cost > new_cost
    return (download(s, X, N, n, B), download(s, U, M, m, B))
end

"""
One call, host in → host out (upload, batched fit with per-trajectory convergence, download): `ilqr_solve`.
Returns (x̄, ū, cost[B], iters[B], status[B]).
"""
function solve(s::Solver, x_init::Array{Float64,3}, u_init::Array{Float64,3}; max_iter::Int64 = 100, tol::Float64 = 1e-6)
    B = s.p.B
    x = similar(x_init); u = similar(u_init)
    cost = Vector{Float64}(undef, B); iters = Vector{Int32}(undef, B); status = Vector{Int32}(undef, B)
    check(ccall((:ilqr_solve, lib), Int32,
                (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int32, Float64, Ptr{Float64}, Ptr{Float64},
                 Ptr{Float64}, Ptr{Int32}, Ptr{Int32}),
                s.h, x_init, u_init, C_NULL, max_iter, tol, x, u, cost, iters, status), s.h)
    return (x, u, cost, iters, status)
end

"Receding-horizon MPC (BASELINE config 5): `mpc_start!` once, then `mpc_step!` per plant step → (u_applied[m,B], x_plant[n,B])."
mpc_start!(s::Solver, x0::Matrix{Float64}) =
    check(ccall((:ilqr_mpc_start, lib), Int32, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), s.h, x0, C_NULL), s.h)
function mpc_step!(s::Solver; max_iter::Int64 = 3, tol::Float64 = 1e-6)
    ua = Matrix{Float64}(undef, s.p.m, s.p.B); xp = Matrix{Float64}(undef, s.p.n, s.p.B)
    check(ccall((:ilqr_mpc_step, lib), Int32, (Ptr{Cvoid}, Int32, Float64, Ptr{Float64}, Ptr{Float64}), s.h, max_iter, tol, ua, xp), s.h)
    return (ua, xp)
end

"""
Continuous batching (`ilqr_streamer_*`, 2-link model): batches of `batch_size` trajectories are solved as ONE stream through the
`p.B` slots of a handle — a slot that finishes a trajectory takes the next pending one in the same launch — with per-trajectory
`fit` semantics (src/forward_pass.jl:148-179).  `submit!` returns a ticket at once; the arrays must stay alive (and pinned, for
full-speed copies) until `wait(streamer, ticket)` returns.  Size `p.B` to the machine (56,832 on a B200), not to the batch.
"""
mutable struct Streamer
    h::Ptr{Cvoid}
    batch_size::Int
    function Streamer(p::Problem, batch_size::Integer; ring::Integer = 12, max_iter::Integer = 100, tol::Float64 = 1e-6)
        h = Ref{Ptr{Cvoid}}(C_NULL)
        rc = ccall((:ilqr_streamer_create, lib), Int32, (Ref{Problem}, Int32, Int32, Int32, Float64, Ptr{Ptr{Cvoid}}),
                   p, batch_size, ring, max_iter, tol, h)
        rc == 0 || error("ilqr_streamer_create: " * unsafe_string(ccall((:ilqr_streamer_last_error, lib), Cstring, (Ptr{Cvoid},), C_NULL)))
        s = new(h[], batch_size)
        finalizer(x -> ccall((:ilqr_streamer_destroy, lib), Int32, (Ptr{Cvoid},), x.h), s)
        s
    end
end
function submit!(s::Streamer, x_init::Array{Float64,3}, u_init::Array{Float64,3}, x_out::Array{Float64,3}, u_out::Array{Float64,3},
                 cost::Vector{Float64}, iters::Vector{Int32}, status::Vector{Int32})
    @assert size(x_init, 3) == s.batch_size == size(u_init, 3) "every submitted batch has batch_size trajectories"
    t = ccall((:ilqr_streamer_submit, lib), Int64,
              (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Int32}),
              s.h, x_init, u_init, x_out, u_out, cost, iters, status)
    t >= 0 || error("ilqr_streamer_submit failed ($t)")
    t
end
"""
`x0[n,batch_size]` (+ `u_init` or `nothing` = zeros): `x_init` is rolled out on the device (animate_2_link.jl:11-16), so only
x0 crosses the bus.  Any output may be `nothing` (not produced, not copied back).
"""
function submit_x0!(s::Streamer, x0::Matrix{Float64}, u_init, x_out, u_out, cost, iters, status)
    @assert size(x0, 2) == s.batch_size "every submitted batch has batch_size trajectories"
    nul(a) = a === nothing ? C_NULL : pointer(a)
    t = GC.@preserve x0 u_init x_out u_out cost iters status ccall((:ilqr_streamer_submit_x0, lib), Int64,
              (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Int32}),
              s.h, x0, nul(u_init), nul(x_out), nul(u_out), nul(cost), nul(iters), nul(status))
    t >= 0 || error("ilqr_streamer_submit_x0 failed ($t)")
    t
end
"Device-pointer forms (CUDA.jl `CuPtr`s converted to `Ptr{Float64}` by the caller): read and written by the kernels directly."
submit_device!(s::Streamer, d_x::Ptr{Float64}, d_u::Ptr{Float64}, d_xo::Ptr{Float64}, d_uo::Ptr{Float64},
               d_cost::Ptr{Float64}, d_iters::Ptr{Int32}, d_status::Ptr{Int32}) =
    ccall((:ilqr_streamer_submit_device, lib), Int64,
          (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Int32}),
          s.h, d_x, d_u, d_xo, d_uo, d_cost, d_iters, d_status)
submit_x0_device!(s::Streamer, d_x0::Ptr{Float64}, d_u::Ptr{Float64}, d_xo::Ptr{Float64}, d_uo::Ptr{Float64},
                  d_cost::Ptr{Float64}, d_iters::Ptr{Int32}, d_status::Ptr{Int32}) =
    ccall((:ilqr_streamer_submit_x0_device, lib), Int64,
          (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Int32}),
          s.h, d_x0, d_u, d_xo, d_uo, d_cost, d_iters, d_status)
wait_all(s::Streamer) = ccall((:ilqr_streamer_wait_all, lib), Int32, (Ptr{Cvoid},), s.h) == 0 ||
    error(unsafe_string(ccall((:ilqr_streamer_last_error, lib), Cstring, (Ptr{Cvoid},), s.h)))
rounds(s::Streamer) = ccall((:ilqr_streamer_rounds, lib), Int64, (Ptr{Cvoid},), s.h)

"""
`ilqr_stream_solve_device`: `n_total` trajectories resident in HBM (device pointers, boundary layout) streamed through the
solver's `p.B` slots; per-trajectory `fit` semantics.  Returns the number of batch iterations (launches of a whole iteration).
"""
function stream_solve_device(s::Solver, n_total::Integer, d_x::Ptr{Float64}, d_u::Ptr{Float64}, d_xo::Ptr{Float64}, d_uo::Ptr{Float64},
                             d_cost::Ptr{Float64}, d_iters::Ptr{Int32}, d_status::Ptr{Int32}; max_iter::Integer = 100, tol::Float64 = 1e-6)
    its = Ref{Int64}(0)
    check(ccall((:ilqr_stream_solve_device, lib), Int32,
                (Ptr{Cvoid}, Int64, Ptr{Float64}, Ptr{Float64}, Int32, Float64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32},
                 Ptr{Int32}, Ptr{Int64}),
                s.h, n_total, d_x, d_u, max_iter, tol, d_xo, d_uo, d_cost, d_iters, d_status, its), s.h)
    return its[]
end

"""
Batch scheduler (`ilqr_pool_*`): `n_handles` batches in flight on one GPU, each solved exactly as `solve` would.
`submit!` returns a ticket; keep the arrays alive until `wait(pool, ticket)`.
"""
mutable struct Pool
    h::Ptr{Cvoid}
    function Pool(p::Problem, n_handles::Integer)
        h = Ref{Ptr{Cvoid}}(C_NULL)
        rc = ccall((:ilqr_pool_create, lib), Int32, (Ref{Problem}, Int32, Ptr{Ptr{Cvoid}}), p, n_handles, h)
        rc == 0 || error("ilqr_pool_create: " * unsafe_string(ccall((:ilqr_pool_last_error, lib), Cstring, (Ptr{Cvoid},), C_NULL)))
        s = new(h[])
        finalizer(x -> ccall((:ilqr_pool_destroy, lib), Int32, (Ptr{Cvoid},), x.h), s)
        s
    end
end
function submit!(pool::Pool, x_init::Array{Float64,3}, u_init::Array{Float64,3}, x_out::Array{Float64,3}, u_out::Array{Float64,3},
                 cost::Vector{Float64}, iters::Vector{Int32}, status::Vector{Int32}; x_traj = nothing, max_iter::Integer = 100,
                 tol::Float64 = 1e-6)
    t = ccall((:ilqr_pool_submit, lib), Int64,
              (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int32, Float64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
               Ptr{Int32}, Ptr{Int32}),
              pool.h, x_init, u_init, x_traj === nothing ? C_NULL : x_traj, max_iter, tol, x_out, u_out, cost, iters, status)
    t >= 0 || error("ilqr_pool_submit failed ($t)")
    t
end
submit_device!(pool::Pool, d_x::Ptr{Float64}, d_u::Ptr{Float64}, d_xt::Ptr{Float64}, d_xo::Ptr{Float64}, d_uo::Ptr{Float64},
               d_cost::Ptr{Float64}, d_iters::Ptr{Int32}, d_status::Ptr{Int32}; max_iter::Integer = 100, tol::Float64 = 1e-6) =
    ccall((:ilqr_pool_submit_device, lib), Int64,
          (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int32, Float64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
           Ptr{Int32}, Ptr{Int32}),
          pool.h, d_x, d_u, d_xt, max_iter, tol, d_xo, d_uo, d_cost, d_iters, d_status)
Base.wait(pool::Pool, ticket::Int64) =
    ccall((:ilqr_pool_wait, lib), Int32, (Ptr{Cvoid}, Int64), pool.h, ticket) == 0 ||
    error(unsafe_string(ccall((:ilqr_pool_last_error, lib), Cstring, (Ptr{Cvoid},), pool.h)))
wait_all(pool::Pool) = ccall((:ilqr_pool_wait_all, lib), Int32, (Ptr{Cvoid},), pool.h) == 0 ||
    error(unsafe_string(ccall((:ilqr_pool_last_error, lib), Cstring, (Ptr{Cvoid},), pool.h)))

"Page-locked host arrays for full-speed PCIe copies (`ilqr_host_alloc`); release with `host_free`."
function host_alloc(::Type{T}, dims...) where {T}
    p = Ref{Ptr{Cvoid}}(C_NULL)
    ccall((:ilqr_host_alloc, lib), Int32, (Ptr{Ptr{Cvoid}}, UInt64), p, sizeof(T) * prod(dims)) == 0 || error("ilqr_host_alloc")
    return unsafe_wrap(Array, Ptr{T}(p[]), dims)
end
host_free(a::Array) = ccall((:ilqr_host_free, lib), Int32, (Ptr{Cvoid},), a)

"With fit's keyword argument `x_traj` (src/forward_pass.jl:151); same shape as `x_init`."
function submit_traj!(s::Streamer, x_init::Array{Float64,3}, u_init::Array{Float64,3}, x_traj::Array{Float64,3}, x_out::Array{Float64,3},
                      u_out::Array{Float64,3}, cost::Vector{Float64}, iters::Vector{Int32}, status::Vector{Int32})
    t = ccall((:ilqr_streamer_submit_traj, lib), Int64,
              (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Int32}),
              s.h, x_init, u_init, x_traj, x_out, u_out, cost, iters, status)
    t >= 0 || error("ilqr_streamer_submit_traj failed ($t)")
    t
end
Base.wait(s::Streamer, ticket::Int64) =
    ccall((:ilqr_streamer_wait, lib), Int32, (Ptr{Cvoid}, Int64), s.h, ticket) == 0 ||
    error(unsafe_string(ccall((:ilqr_streamer_last_error, lib), Cstring, (Ptr{Cvoid},), s.h)))

# single-problem convenience with the reference's exact shapes x[N×n], u[H×m]
fit(x_init::Matrix{Float64}, u_init::Matrix{Float64}, p::Problem; kw...) = begin
    (x, u) = fit(reshape(x_init, size(x_init)..., 1), reshape(u_init, size(u_init)..., 1), p; kw...)
    (x[:, :, 1], u[:, :, 1])
end

end # module
