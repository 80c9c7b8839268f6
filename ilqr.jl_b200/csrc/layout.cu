// layout.cu — boundary layout <-> device layout.
//
// Boundary ("TF", time-fastest) = the reference's Julia column-major arrays with the batch
// trailing: x[N,n,B] ⇒ trajectory t is one contiguous slab of ncomp·T doubles, element (k,c) at
// c·T + k (single-trajectory shapes: src/backward_pass.jl:332-333).
// Device ("KSC") = [k][slot][component]: element (k,c) of slot s at (k·S + s)·ncomp + c.
//   * the 32 lanes of a warp own 32 consecutive slots ⇒ the slab a warp needs for time step k
//     is ONE contiguous run of 32·ncomp doubles (1 KB for x) — a single TMA bulk copy;
//   * a lane's own components are one 16/32/64-byte vector ⇒ LDG/STG.128.
// Both directions go through a padded shared tile so global reads and writes are both
// contiguous runs (128 B along k on the boundary side, 32·ncomp·8 B on the device side).
#include "internal.cuh"

namespace ilqr {
namespace {

constexpr int kMaxTS = 32;   // slots per tile (fewer when ncomp is large, so that a tile stays ≤ 64 KB)
constexpr int TK = 16;   // time steps per tile
constexpr int kThreads = 256;

// dynamic smem: TK rows of (TS*ncomp + 1) doubles
__global__ void __launch_bounds__(kThreads)
tf_to_ksc_kernel(const double* __restrict__ tf, double* __restrict__ ksc, const int32_t* __restrict__ slot_traj,
                 int nslots, int T, int ncomp, int64_t S, int shift, int TS) {
  extern __shared__ double tile[];
  const int row = TS * ncomp + 1;
  const int k0 = blockIdx.y * TK, s0 = blockIdx.x * TS;
  const int L = ncomp * T, E = TS * TK * ncomp;
  for (int idx = threadIdx.x; idx < E; idx += kThreads) {
    const int kk = idx % TK, rest = idx / TK, c = rest % ncomp, sl = rest / ncomp;
    const int k = k0 + kk, s = s0 + sl;
    if (k < T && s < nslots) {
      const int64_t t = slot_traj ? slot_traj[s] : s;
      // shift > 0: read time index k + shift, zero past the end (receding-horizon warm start)
      tile[kk * row + sl * ncomp + c] = (k + shift < T) ? tf[t * L + (int64_t)c * T + k + shift] : 0.0;
    }
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < E; idx += kThreads) {
    const int within = idx % (TS * ncomp), kk = idx / (TS * ncomp);
    const int sl = within / ncomp;
    const int k = k0 + kk, s = s0 + sl;
    if (k < T && s < nslots) ksc[((int64_t)k * S + s0) * ncomp + within] = tile[kk * row + within];
  }
}

__global__ void __launch_bounds__(kThreads)
ksc_to_tf_kernel(const double* __restrict__ b0, const double* __restrict__ b1, const int32_t* __restrict__ sel,
                 double* __restrict__ tf, const int32_t* __restrict__ slot_traj, int nslots, int T, int ncomp, int64_t S,
                 int TS) {
  extern __shared__ double tile[];
  const int row = TS * ncomp + 1;
  const int k0 = blockIdx.y * TK, s0 = blockIdx.x * TS;
  const int L = ncomp * T, E = TS * TK * ncomp;
  for (int idx = threadIdx.x; idx < E; idx += kThreads) {
    const int within = idx % (TS * ncomp), kk = idx / (TS * ncomp);
    const int sl = within / ncomp;
    const int k = k0 + kk, s = s0 + sl;
    if (k < T && s < nslots) {
      const double* src = (sel && sel[s]) ? b1 : b0;
      tile[kk * row + within] = src[((int64_t)k * S + s0) * ncomp + within];
    }
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < E; idx += kThreads) {
    const int kk = idx % TK, rest = idx / TK, c = rest % ncomp, sl = rest / ncomp;
    const int k = k0 + kk, s = s0 + sl;
    if (k < T && s < nslots) {
      const int64_t t = slot_traj ? slot_traj[s] : s;
      tf[t * L + (int64_t)c * T + k] = tile[kk * row + sl * ncomp + c];
    }
  }
}

// slots per tile: 32 for the small models; large component counts (K of a 7-DoF chain: 98) shrink the tile
inline int tile_slots(int ncomp) { int ts = 512 / ncomp; return ts < 1 ? 1 : (ts > kMaxTS ? kMaxTS : ts); }

}  // namespace

void init_layout_attributes() {
  cudaFuncSetAttribute(tf_to_ksc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
  cudaFuncSetAttribute(ksc_to_tf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
}

void launch_tf_to_bf(const double* tf, double* bf, const int32_t* slot_traj, int nslots, int T, int ncomp, int64_t S,
                     cudaStream_t s, int shift) {
  if (nslots <= 0) return;
  const int TS = tile_slots(ncomp);
  dim3 grid((nslots + TS - 1) / TS, (T + TK - 1) / TK);
  const size_t smem = sizeof(double) * TK * (TS * ncomp + 1);
  tf_to_ksc_kernel<<<grid, kThreads, smem, s>>>(tf, bf, slot_traj, nslots, T, ncomp, S, shift, TS);
}

void launch_bf_to_tf(const double* bf0, const double* bf1, const int32_t* sel, double* tf, const int32_t* slot_traj,
                     int nslots, int T, int ncomp, int64_t S, cudaStream_t s) {
  if (nslots <= 0) return;
  const int TS = tile_slots(ncomp);
  dim3 grid((nslots + TS - 1) / TS, (T + TK - 1) / TK);
  const size_t smem = sizeof(double) * TK * (TS * ncomp + 1);
  ksc_to_tf_kernel<<<grid, kThreads, smem, s>>>(bf0, bf1, sel, tf, slot_traj, nslots, T, ncomp, S, TS);
}

}  // namespace ilqr
