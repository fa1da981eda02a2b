// fp64_peak.cu — measures the B200 FP64 pipe: DFMA/DADD/DMUL throughput (TFLOP/s, instr/clk/SM) and dependent-issue latency.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_peak tools/fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int OP, int ILP>
__global__ void k(double* out, int iters, double a, double b) {
  double x[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) x[i] = threadIdx.x * 1e-9 + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) {
      if (OP == 0) x[i] = fma(x[i], a, b);
      if (OP == 1) x[i] = x[i] + a;
      if (OP == 2) x[i] = x[i] * a;
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += x[i];
  if (s == 123.456) out[0] = s;
}
template <int OP, int ILP>
void run(const char* name, int blocks, int threads) {
  double* d; cudaMalloc(&d, 8);
  int iters = 20000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<OP, ILP><<<blocks, threads>>>(d, 100, 1.0000001, 1e-9);
  cudaEventRecord(e0);
  k<OP, ILP><<<blocks, threads>>>(d, iters, 1.0000001, 1e-9);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double inst = (double)blocks * threads * iters * ILP;
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("%-6s ILP=%2d blocks=%4d thr=%4d : %8.3f ms  %7.2f Ginst/s  %6.2f TFLOP/s(fma=2)  %6.1f inst/clk/SM @%d MHz nominal\n", name, ILP, blocks, threads, ms,
         inst / ms / 1e6, inst * (OP == 0 ? 2 : 1) / ms / 1e9, inst / (ms * 1e-3) / 148 / (clk * 1e3), clk / 1000);
  cudaFree(d);
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  printf("%s SMs=%d\n", p.name, p.multiProcessorCount);
  run<0, 8>("DFMA", 148 * 4, 256); run<0, 8>("DFMA", 148 * 2, 1024); run<0, 4>("DFMA", 148, 128); run<0, 1>("DFMA", 148, 128); run<0, 2>("DFMA", 148, 128);
  run<0, 1>("DFMA", 148, 32); run<0, 2>("DFMA", 148, 32); run<0, 4>("DFMA", 148, 32); run<0, 8>("DFMA", 148, 32);
  run<1, 8>("DADD", 148 * 4, 256); run<2, 8>("DMUL", 148 * 4, 256);
  run<0, 8>("DFMA", 148 * 1, 256); run<0, 8>("DFMA", 148 * 1, 512);
  return 0;
}
