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

// The same with three distinct VECTOR register operands per DFMA (what real code issues: the loop above multiplies by two
// uniform values, i.e. one register operand per instruction) — the register-file side of the FP64 pipe's peak.
template <int ILP>
__global__ void dfma3_throughput(double* out, int iters, double a, double b) {
  double acc[ILP], x[ILP], y[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) { acc[i] = threadIdx.x * 1e-3 + i; x[i] = a + 1e-12 * (threadIdx.x + i); y[i] = b + 1e-13 * (threadIdx.x * 3 + i); }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = fma(acc[i], x[(i + 3) % ILP], y[(i + 5) % ILP]);
    if (it == iters - 7) { x[0] += 1e-15; y[1] += 1e-15; }   // keeps x, y in registers as live, non-constant values
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += acc[i];
  if (s == 12345.678) out[0] = s;
}

// Operand mix of a DFMA stream, MODE: 0 = fma(acc, x_vec, b_uniform) (two vector register operands),
// 1 = fma(x_i, y_j, acc) with x_i shared by two consecutive instructions (what a small matrix product issues: the
// compiler can mark it .reuse), 2 = DMUL + DADD pairs on vector registers.
template <int MODE>
__global__ void dfma_mix_throughput(double* out, int iters, double a, double b) {
  double acc[8], x[8], y[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { acc[i] = threadIdx.x * 1e-3 + i; x[i] = a + 1e-12 * (threadIdx.x + i); y[i] = b + 1e-13 * (threadIdx.x * 3 + i); }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) acc[i] = fma(acc[i], x[(i + 3) % 8], b);
      if (MODE == 1) acc[i] = fma(x[i / 2], y[(i + 5) % 8], acc[i]);
      if (MODE == 2) acc[i] = (i & 1) ? acc[i] * x[(i + 3) % 8] : acc[i] + y[(i + 5) % 8];
    }
    if (it == iters - 7) { x[0] += 1e-15; y[1] += 1e-15; }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += acc[i];
  if (s == 12345.678) out[0] = s;
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
  // three register operands; full occupancy and the round kernel's residency (one block of 12 warps per SM)
  for (int cfg = 0; cfg < 2; ++cfg) {
    const int bl = cfg == 0 ? blocks : p.multiProcessorCount, th = cfg == 0 ? threads : 384;
    float b3 = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
      cudaEventRecord(e0);
      dfma3_throughput<8><<<bl, th>>>(out, iters, 1.0000001, 1e-9);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (rep > 0 && ms < b3) b3 = ms;
    }
    printf(", \"%s\": %.3f", cfg == 0 ? "fp64_dfma_3reg_tflops" : "fp64_dfma_3reg_12warps_tflops", 2.0 * 8 * iters * (double)bl * th / (b3 * 1e-3) / 1e12);
  }
  {
    const char* mix[3] = {"fp64_dfma_2reg_tflops", "fp64_dfma_3reg_shared_operand_tflops", "fp64_dmul_dadd_2reg_slot_tflops"};
    for (int mode = 0; mode < 3; ++mode) {
      float bm = 1e30f;
      for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        if (mode == 0) dfma_mix_throughput<0><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
        if (mode == 1) dfma_mix_throughput<1><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
        if (mode == 2) dfma_mix_throughput<2><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < bm) bm = ms;
      }
      printf(", \"%s\": %.3f", mix[mode], 2.0 * 8 * iters * (double)blocks * threads / (bm * 1e-3) / 1e12);
    }
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
