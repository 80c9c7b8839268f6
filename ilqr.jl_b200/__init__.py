"""ilqr.jl_b200 — B200-native batched iLQR hot path behind iLQR.jl's solver interface.

Only what the path needs lives here: csrc/ (sm_100a CUDA kernels + the C ABI of
include/ilqr_b200.h) and host.py (the host-side mirror of the reference's
fit / backward_pass / forward_pass).  There is no CPU implementation in this
package: importing works anywhere, but every compute call requires the CUDA
library and a B200.
"""
from . import _abi, _build
from ._abi import Problem, load_library
from ._build import build
from .host import BatchSolver, IlqrError, SolverPool, Streamer, backward_pass, custom_compile_check, custom_problem, fit, forward_pass, load_urdf, serial_chain_problem, two_link_problem
from .sharding import fit_sharded, shard_range

__all__ = ["BatchSolver", "IlqrError", "Problem", "SolverPool", "Streamer", "backward_pass", "build", "custom_compile_check", "custom_problem", "fit", "fit_sharded", "forward_pass", "load_library", "load_urdf", "serial_chain_problem", "shard_range",
           "two_link_problem"]
