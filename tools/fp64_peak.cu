// fp64_peak.cu — DFMA micro-benchmark: measured FP64-pipe peak (the roofline denominator that
// MEASURED_PEAKS.json does not carry) and the dependent-issue latency of DFMA / sincos / division.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_peak tools/fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void dfma_throughput(double* out, int iters, double a, double b, int active_lanes) {
  if ((threadIdx.x & 31) >= active_lanes) return;
  double acc[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = fma(acc[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += acc[i];
  if (s == 12345.678) out[0] = s;
}

// Operand side of the FP64 pipe: the same DFMA stream with 2 or 3 distinct VECTOR-register sources per instruction.  The
// operands are loop invariant and loaded from memory (nothing for the compiler to fold), the loop body is 32 DFMAs and a
// branch — check with `cuobjdump -sass tools/fp64_peak`.  (The first version of this probe perturbed its operands inside the
// loop; the compiler turned that into a stream that was half MOV / ISETP / predicated DADD, and the rates it reported —
// 26.5 and 18.2 TFLOP/s — measured that, not the operand path.)
//   MODE 2: fma(acc, x_j, b_uniform)        two vector sources
//   MODE 3: fma(acc, x_j, y_k)              three distinct vector sources
//   MODE 5: fma(x_r, y_k, acc), x_r shared by eight consecutive instructions (what a small matrix product issues: .reuse)
template <int MODE>
__global__ void __launch_bounds__(512) dfma_operands(double* out, const double* __restrict__ in, int iters) {
  double acc[8], x[8], y[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { acc[i] = in[threadIdx.x + 32 * i]; x[i] = in[1024 + threadIdx.x + 32 * i]; y[i] = in[2048 + threadIdx.x + 32 * i]; }
  const double ub = in[4097];
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (MODE == 2) acc[i] = fma(acc[i], x[(i + 3 + r) % 8], ub);
        if (MODE == 3) acc[i] = fma(acc[i], x[(i + 3 + r) % 8], y[(i + 5 + 2 * r) % 8]);
        if (MODE == 5) acc[i] = fma(x[r], y[(i + 5 + 2 * r) % 8], acc[i]);
      }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void latency_kernel(double* out, long long* cycles, int iters, double a, double b, int mode) {
  double x = a;
  long long t0 = clock64();
  if (mode == 0) for (int i = 0; i < iters; ++i) x = fma(x, a, b);
  if (mode == 1) for (int i = 0; i < iters; ++i) { double s, c; sincos(x, &s, &c); x = s + c; }
  if (mode == 2) for (int i = 0; i < iters; ++i) x = 1.0 / (x + b);
  if (mode == 3) for (int i = 0; i < iters; ++i) x = x + b;
  if (mode == 4) for (int i = 0; i < iters; ++i) x = x * a;
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) cycles[0] = t1 - t0;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  double* out; cudaMalloc(&out, 1 << 20);
  long long* cyc; cudaMalloc(&cyc, 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 1 << 14;
  float best = 1e30f;
  const int blocks = p.multiProcessorCount * 4, threads = 512;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0);
    dfma_throughput<8><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9, 32);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best) best = ms;
  }
  const double flops = 2.0 * 8 * iters * (double)blocks * threads;
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("{\"device\": \"%s\", \"sms\": %d, \"fp64_dfma_tflops\": %.3f, \"ms\": %.4f, \"clock_khz_attr\": %d",
         p.name, p.multiProcessorCount, flops / (best * 1e-3) / 1e12, best, clk);
  // does a half-empty warp issue DFMA faster?  (per-warp-instruction time with 16 / 8 active lanes)
  for (int lanes : {32, 16, 8}) {
    float b2 = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(e0);
      dfma_throughput<8><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9, lanes);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (ms < b2) b2 = ms;
    }
    printf(", \"ms_lanes%d\": %.4f", lanes, b2);
  }
  // operand count: full occupancy, and the round kernel's residency (one block of 12 warps per SM) for the three-source case
  {
    double* in; cudaMalloc(&in, 1 << 16);
    double* h = new double[8192];
    for (int i = 0; i < 8192; ++i) h[i] = 1.0 + 1e-9 * i;
    h[4096] = 1.0000001; h[4097] = 1e-9;
    cudaMemcpy(in, h, 8192 * sizeof(double), cudaMemcpyHostToDevice);
    delete[] h;
    const int it2 = 1 << 12;
    auto time_mode = [&](int mode, int bl, int th) {
      float bm = 1e30f;
      for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        if (mode == 2) dfma_operands<2><<<bl, th>>>(out, in, it2);
        if (mode == 3) dfma_operands<3><<<bl, th>>>(out, in, it2);
        if (mode == 5) dfma_operands<5><<<bl, th>>>(out, in, it2);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < bm) bm = ms;
      }
      return 2.0 * 32 * it2 * (double)bl * th / (bm * 1e-3) / 1e12;
    };
    printf(", \"fp64_dfma_2reg_tflops\": %.3f", time_mode(2, blocks, threads));
    printf(", \"fp64_dfma_3reg_tflops\": %.3f", time_mode(3, blocks, threads));
    printf(", \"fp64_dfma_3reg_12warps_tflops\": %.3f", time_mode(3, p.multiProcessorCount, 384));
    printf(", \"fp64_dfma_3reg_shared_operand_tflops\": %.3f", time_mode(5, blocks, threads));
    cudaFree(in);
  }
  const char* names[5] = {"dfma", "sincos", "div", "dadd", "dmul"};
  for (int mode = 0; mode < 5; ++mode) {
    latency_kernel<<<1, 32>>>(out, cyc, 4096, 0.999, 0.001, mode);
    cudaDeviceSynchronize();
    latency_kernel<<<1, 32>>>(out, cyc, 4096, 0.999, 0.001, mode);
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf(", \"lat_%s_cycles\": %.1f", names[mode], c / 4096.0);
  }
  printf("}\n");
  return 0;
}
