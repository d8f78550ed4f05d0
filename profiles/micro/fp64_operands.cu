// Micro-benchmark: FP64 FMA throughput on B200 when the three source operands are distinct, changing registers
// (as in the adjoint kernel) instead of one accumulator and two loop constants, and the effect of interleaved
// non-FP64 instructions on the dispatch port.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_operands fp64_operands.cu && ./fp64_operands
#include <cstdio>
#include <cuda_runtime.h>

// MODE 0: x[j] = fma(x[j], a, b)                 accumulator + loop constants (what a peak probe measures)
// MODE 1: x[j] = fma(x[j+1], x[j+3], -x[j+5])      three distinct, changing register operands
// MODE 2: MODE 1 + one integer instruction per FMA
// MODE 3: MODE 1 + two integer instructions per FMA
// MODE 4: x[j] = fma(x[j+1], x[j+3], x[j])         three registers, accumulating in place
// MODE 5: x[j] = x[j+1] * x[j+3]                   DMUL, two registers
// MODE 6: x[j] = fma(x[j+1], a, -x[j+5])           two registers + a kernel parameter (constant-bank operand)
// explicit versions (one FP64 instruction per step)
template <int MODE>
__global__ void k2(int iters, double* out, long long* cyc, int seed, double a, double b) {
  constexpr int N = 12;
  double x[N];
  int q[N];
#pragma unroll
  for (int j = 0; j < N; ++j) { x[j] = 0.5 + (threadIdx.x + j) * 1e-3; q[j] = seed + j; }
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int j = 0; j < N; ++j) {
        if (MODE == 0) x[j] = fma(x[j], a, b);
        else if (MODE == 4) x[j] = fma(x[(j + 1) % N], x[(j + 3) % N], x[j]);
        else if (MODE == 5) x[j] = x[(j + 1) % N] * x[(j + 3) % N];
        else if (MODE == 6) x[j] = fma(x[(j + 1) % N], a, -x[(j + 5) % N]);
        else x[j] = fma(x[(j + 1) % N], x[(j + 3) % N], -x[(j + 5) % N]);  // values may over/underflow: speed only
        if (MODE == 2 || MODE == 3) q[j] = q[j] * 3 + q[(j + 1) % N];
        if (MODE == 3) q[(j + 2) % N] ^= q[j] >> 3;
      }
    }
  }
  long long t1 = clock64();
  double s = 0; int r = 0;
#pragma unroll
  for (int j = 0; j < N; ++j) { s += x[j]; r += q[j]; }
  if (s == 1234.5 || r == 17) out[0] = s + r;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <int MODE>
void run(int warps) {
  double* out; long long* cyc; cudaMalloc(&out, 8); cudaMalloc(&cyc, 8);
  const int iters = 2048;
  k2<MODE><<<1, 32 * warps>>>(iters, out, cyc, 3, 0.999, 1e-9);
  k2<MODE><<<1, 32 * warps>>>(iters, out, cyc, 3, 0.999, 1e-9);
  long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  double per = (double)h / (iters * 4.0 * 12);
  printf("mode %d warps/SM=%2d : %.2f cycles per FP64 step per warp  (=> %.2f DFMA warp-instr/cycle/SM)\n", MODE, warps, per, warps / per);
  cudaFree(out); cudaFree(cyc);
}
int main() {
  for (int w : {8, 12, 16}) { run<0>(w); run<1>(w); run<2>(w); run<3>(w); run<4>(w); run<5>(w); run<6>(w); }
  return 0;
}
