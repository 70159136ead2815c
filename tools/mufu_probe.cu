// MUFU throughput probe: f32 vs packed f16x2 / bf16x2 transcendental ops on sm_100a
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)
template <int MODE>
__global__ void k(float* out, int iters) {
  float a[8]; unsigned h[8];
  for (int i = 0; i < 8; ++i) { a[i] = 0.001f * (threadIdx.x + i); h[i] = 0x2C002C00u + threadIdx.x + i; }
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (MODE == 1) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a[i]));
      if (MODE == 2) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h[i]));
      if (MODE == 3) asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(h[i]));
      if (MODE == 4) asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(h[i]));
      if (MODE == 5) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(h[i]));
      if (MODE == 6) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (MODE == 7) asm volatile("sin.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
    }
  }
  long long t1 = clock64();
  float s = 0; for (int i = 0; i < 8; ++i) s += a[i] + __uint_as_float(h[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + (float)(t1 - t0);
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (float)(t1 - t0);
}
int main() {
  float* d; CK(cudaMalloc(&d, 148 * 1024 * 4));
  const char* names[] = {"ex2.f32", "tanh.f32", "ex2.f16x2", "tanh.f16x2", "tanh.bf16x2", "ex2.bf16x2", "rcp.f32", "sin.f32"};
  const int iters = 2000;
  for (int mode = 0; mode < 8; ++mode) {
    switch (mode) {
      case 0: k<0><<<148, 1024>>>(d, iters); break; case 1: k<1><<<148, 1024>>>(d, iters); break;
      case 2: k<2><<<148, 1024>>>(d, iters); break; case 3: k<3><<<148, 1024>>>(d, iters); break;
      case 4: k<4><<<148, 1024>>>(d, iters); break; case 5: k<5><<<148, 1024>>>(d, iters); break;
      case 6: k<6><<<148, 1024>>>(d, iters); break; case 7: k<7><<<148, 1024>>>(d, iters); break;
    }
    CK(cudaDeviceSynchronize());
    float cyc; CK(cudaMemcpy(&cyc, d, 4, cudaMemcpyDeviceToHost));
    double instr_per_sm = 32.0 * 8 * iters;   // warp-instructions per SM (32 warps)
    printf("%-12s %8.0f cycles  -> %.2f cycles per warp-instruction per SM (%.1f lane-results/clk/SM, x2 elements if packed)\n",
           names[mode], cyc, cyc / instr_per_sm, 32.0 * instr_per_sm / cyc);
  }
  return 0;
}
