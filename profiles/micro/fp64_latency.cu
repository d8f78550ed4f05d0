// Micro-benchmark: dependent-issue latency and throughput of FP64 FMA on B200 (one warp per SMSP vs many).
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_latency fp64_latency.cu && ./fp64_latency
#include <cstdio>
#include <cuda_runtime.h>
template <int CH>
__global__ void chain(int iters, double a, double b, double* out, long long* cyc) {
  double x[CH];
#pragma unroll
  for (int j = 0; j < CH; ++j) x[j] = threadIdx.x * 1e-9 + j;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
#pragma unroll
      for (int j = 0; j < CH; ++j) x[j] = fma(x[j], a, b);
    }
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int j = 0; j < CH; ++j) s += x[j];
  if (s == 1234.5) out[0] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int CH>
void run(int warps) {
  double* out; long long* cyc; cudaMalloc(&out, 8); cudaMalloc(&cyc, 8);
  const int iters = 4096;
  chain<CH><<<1, 32 * warps>>>(iters, 0.999, 1e-9, out, cyc);
  chain<CH><<<1, 32 * warps>>>(iters, 0.999, 1e-9, out, cyc);
  long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  double per = (double)h / (iters * 8.0 * CH);
  printf("chains/thread=%d warps/SM=%2d : %.2f cycles per DFMA per warp  (=> %.2f DFMA warp-instr/cycle/SM)\n", CH, warps, per * 1.0, warps / per);
  cudaFree(out); cudaFree(cyc);
}
int main() {
  run<1>(1); run<2>(1); run<4>(1); run<8>(1);
  run<1>(4); run<2>(4); run<4>(4); run<1>(8); run<2>(8); run<4>(8); run<1>(12); run<2>(12); run<3>(12); run<1>(16); run<2>(16); run<1>(32);
  return 0;
}
