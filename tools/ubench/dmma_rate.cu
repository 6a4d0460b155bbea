// Microbenchmark: peak issue rate of mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4) from registers only.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_rate dmma_rate.cu && ./dmma_rate
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int NACC, int NA, int NB>
__global__ void k(double *out, int iters, double a0, double b0) {
  double acc[NACC][2];
  double a[NA], b[NB];
  for (int i = 0; i < NACC; i++) { acc[i][0] = 0; acc[i][1] = 0; }
  for (int i = 0; i < NA; i++) a[i] = a0 + i + threadIdx.x;
  for (int i = 0; i < NB; i++) b[i] = b0 + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) dmma884(acc[i][0], acc[i][1], a[i % NA], b[(i / NA) % NB]);
  }
  double s = 0;
  for (int i = 0; i < NACC; i++) s += acc[i][0] + acc[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NACC, int NA, int NB>
void run(int warps, const char *name) {
  double *out; cudaMalloc(&out, 148 * 1024 * 8);
  int iters = 20000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<NACC, NA, NB><<<148, warps * 32>>>(out, 100, 1.0, 2.0);
  cudaEventRecord(e0);
  k<NACC, NA, NB><<<148, warps * 32>>>(out, iters, 1.0, 2.0);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double flop = 2.0 * 256 * NACC * (double)iters * warps * 148;
  printf("%-28s warps/SM=%2d  %.2f TFLOP/s\n", name, warps, flop / ms / 1e9);
  cudaFree(out);
}
int main() {
  for (int w : {4, 8, 16, 32}) {
    if (w == 4) { run<32, 8, 4>(4, "acc32 a8 b4"); run<16, 4, 4>(4, "acc16 a4 b4"); run<8, 1, 1>(4, "acc8 a1 b1"); }
    if (w == 8) { run<32, 8, 4>(8, "acc32 a8 b4"); run<16, 4, 4>(8, "acc16 a4 b4"); run<8, 1, 1>(8, "acc8 a1 b1"); }
    if (w == 16) { run<32, 8, 4>(16, "acc32 a8 b4"); run<16, 4, 4>(16, "acc16 a4 b4"); run<8, 1, 1>(16, "acc8 a1 b1"); }
    if (w == 32) { run<16, 4, 4>(32, "acc16 a4 b4"); run<8, 1, 1>(32, "acc8 a1 b1"); }
  }
  return 0;
}
