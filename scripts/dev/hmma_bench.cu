// micro-benchmark: legacy mma.sync m16n8k16 f16 -> f32 issue rate on sm_100a (development aid)
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
__global__ void k(float* out, int iters, int ilp) {
  uint32_t a0 = threadIdx.x, a1 = 2, a2 = 3, a3 = 4, b0 = 5, b1 = 6;
  float c[8][4];
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                   : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
  float s = 0;
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  float* d; cudaMalloc(&d, 148 * 1024 * 4);
  for (int warps = 4; warps <= 32; warps *= 2) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    k<<<148, warps * 32>>>(d, 100, 8); cudaDeviceSynchronize();
    cudaEventRecord(e0); k<<<148, warps * 32>>>(d, iters, 8); cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double mmas = (double)148 * warps * iters * 8;
    printf("warps/SM %2d: %.3f ms, %.2f HMMA.16816 per ns per SM, %.1f TFLOP/s total\n", warps, ms,
           mmas / 148 / (ms * 1e6), mmas * 4096 / (ms * 1e-3) / 1e12);
  }
  return 0;
}
