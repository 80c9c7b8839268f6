// layout.cu — boundary layout <-> device layout.
//
// Boundary ("TF", time-fastest) = the reference's Julia column-major arrays
// with the batch trailing: x[N,n,B] ⇒ trajectory b is one contiguous slab of
// ncomp·T doubles, element (k,c) at c·T + k (src/backward_pass.jl:332-333 for
// the single-trajectory shapes).  Device ("BF", batch-fastest): element (k,c)
// of slot b at (k·ncomp + c)·S + b.  Both directions go through a 32×33 shared
// tile so that global reads and writes are both full 256 B lines.
#include "internal.cuh"

namespace ilqr {
namespace {

constexpr int TILE = 32, ROWS = 8;

// grid: (ceil(L/32), ceil(B/32)), block (32, 8); L = ncomp*T
__global__ void tf_to_bf_kernel(const double* __restrict__ tf, double* __restrict__ bf,
                                const int32_t* __restrict__ slot_traj, int B, int T, int ncomp, int64_t S) {
  __shared__ double tile[TILE][TILE + 1];
  const int L = ncomp * T;
  const int j0 = blockIdx.x * TILE, b0 = blockIdx.y * TILE;
  for (int r = threadIdx.y; r < TILE; r += ROWS) {
    const int b = b0 + r, j = j0 + threadIdx.x;
    if (b < B && j < L) tile[r][threadIdx.x] = tf[(int64_t)(slot_traj ? slot_traj[b] : b) * L + j];
  }
  __syncthreads();
  for (int r = threadIdx.y; r < TILE; r += ROWS) {
    const int j = j0 + r, b = b0 + threadIdx.x;
    if (b < B && j < L) {
      const int c = j / T, k = j - c * T;
      bf[(int64_t)(k * ncomp + c) * S + b] = tile[threadIdx.x][r];
    }
  }
}

__global__ void bf_to_tf_kernel(const double* __restrict__ bf0, const double* __restrict__ bf1,
                                const int32_t* __restrict__ sel, double* __restrict__ tf,
                                const int32_t* __restrict__ slot_traj, int B, int T, int ncomp, int64_t S) {
  __shared__ double tile[TILE][TILE + 1];
  const int L = ncomp * T;
  const int j0 = blockIdx.x * TILE, b0 = blockIdx.y * TILE;
  {
    const int b = b0 + threadIdx.x;
    const double* src = bf0;
    if (sel && b < B && sel[b]) src = bf1;
    for (int r = threadIdx.y; r < TILE; r += ROWS) {
      const int j = j0 + r;
      if (b < B && j < L) {
        const int c = j / T, k = j - c * T;
        tile[r][threadIdx.x] = src[(int64_t)(k * ncomp + c) * S + b];
      }
    }
  }
  __syncthreads();
  for (int r = threadIdx.y; r < TILE; r += ROWS) {
    const int b = b0 + r, j = j0 + threadIdx.x;
    if (b < B && j < L) tf[(int64_t)(slot_traj ? slot_traj[b] : b) * L + j] = tile[threadIdx.x][r];
  }
}

}  // namespace

void launch_tf_to_bf(const double* tf, double* bf, const int32_t* slot_traj, int nslots, int T, int ncomp, int64_t S,
                     cudaStream_t s) {
  if (nslots <= 0) return;
  const int L = ncomp * T;
  dim3 grid((L + TILE - 1) / TILE, (nslots + TILE - 1) / TILE), block(TILE, ROWS);
  tf_to_bf_kernel<<<grid, block, 0, s>>>(tf, bf, slot_traj, nslots, T, ncomp, S);
}

void launch_bf_to_tf(const double* bf0, const double* bf1, const int32_t* sel, double* tf, const int32_t* slot_traj,
                     int nslots, int T, int ncomp, int64_t S, cudaStream_t s) {
  if (nslots <= 0) return;
  const int L = ncomp * T;
  dim3 grid((L + TILE - 1) / TILE, (nslots + TILE - 1) / TILE), block(TILE, ROWS);
  bf_to_tf_kernel<<<grid, block, 0, s>>>(bf0, bf1, sel, tf, slot_traj, nslots, T, ncomp, S);
}

}  // namespace ilqr
