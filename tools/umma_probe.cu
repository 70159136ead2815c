// Bring-up probe for the tcgen05 path (run on a B200 via gpurun):
//  1. descriptor conventions: K-major no-swizzle A/B, arbitrary 16-byte start offsets, arbitrary LBO
//  2. MMA issue rate for the narrow-N shapes of the enhancer (is the SS-mode A read exposed?)
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include "../fs_uae_image_enhancer_project_b200/csrc/tc_ptx.cuh"

using namespace fsuae::tc;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

// ---------------- test 1: correctness ----------------
// A: planes layout [chunk][row][8] bf16 (row = M index incl. extra rows), B: [2 chunks][N][8]
// D[m][n] = sum_k A[m + row_off][k] * B[n][k],  K = 16 taken from chunk c0 and the chunk at +lbo
__global__ void probe_correct(const __nv_bfloat16* A, const __nv_bfloat16* B, float* D, int a_rows, int N,
                              int row_off, int a_lbo_bytes) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  uint8_t* sA = smem;                       // 2 planes x a_rows x 16 B
  uint8_t* sB = smem + 2 * a_rows * 16;     // 2 x N x 16 B
  for (int i = threadIdx.x; i < 2 * a_rows * 8; i += blockDim.x) ((__nv_bfloat16*)sA)[i] = A[i];
  for (int i = threadIdx.x; i < 2 * N * 8; i += blockDim.x) ((__nv_bfloat16*)sB)[i] = B[i];
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (threadIdx.x < 32) tmem_alloc(&tmem_base, 256);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tb = tmem_base;
  if (threadIdx.x == 0) {
    uint64_t ad = umma_desc(smem_u32(sA) + row_off * 16, a_lbo_bytes, 128);
    uint64_t bd = umma_desc(smem_u32(sB), N * 16, 128);
    umma_bf16(tb, ad, bd, umma_idesc_bf16(128, N), 0);
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  for (int c0 = 0; c0 < N; c0 += 8) {
    uint32_t v[8];
    tmem_ld_x8(tb + ((uint32_t)(warp * 32) << 16) + c0, v);
    tmem_ld_wait();
    for (int j = 0; j < 8; ++j) D[(warp * 32 + lane) * N + c0 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tb, 256);
}

// ---------------- test 2: issue rate ----------------
// one CTA per SM; thread 0 issues `iters` x 64 MMAs; descriptors advance by adds only (lean issue loop)
__global__ void probe_rate(int N, int iters, uint32_t a_step16, uint32_t b_step16, long long* cycles, uint32_t d_off = 0, uint32_t d_alt = 256) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3c003c00u;  // finite bf16
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (threadIdx.x < 32) tmem_alloc(&tmem_base, 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tb = tmem_base;
  const int warp_idx = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  if (warp_idx == 0 && elect_one()) {
    const uint32_t idesc = umma_idesc_bf16(128, N);
    const uint64_t ad0 = umma_desc(smem_u32(smem), 16 * 1024, 128);
    const uint64_t bd0 = umma_desc(smem_u32(smem) + 128 * 1024, N * 16, 128);
    const uint32_t a_hi = (uint32_t)(ad0 >> 32), b_hi = (uint32_t)(bd0 >> 32);
    const uint32_t a_lo0 = (uint32_t)ad0, b_lo0 = (uint32_t)bd0;
    long long t0 = clock64();
    uint32_t phase = 0;
    for (int it = 0; it < iters; ++it) {
      uint32_t a_lo = a_lo0, b_lo = b_lo0;
#pragma unroll
      for (int j = 0; j < 64; ++j) {
        umma_bf16(tb + d_off + (it & 1) * d_alt, ((uint64_t)a_hi << 32) | a_lo, ((uint64_t)b_hi << 32) | b_lo, idesc, 1);
        a_lo += a_step16;   // 64 x step must stay inside the 96 KB A window
        b_lo += b_step16;
      }
      umma_commit(&bar);
      mbar_wait(&bar, phase);
      phase ^= 1;
    }
    long long t1 = clock64();
    cycles[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tb, 512);
}

// ---------------- test 3: A-operand collector reuse ----------------
// Triples of MMAs share one A tile (the three kernel rows of a 3x3 convolution: same input row, three accumulators, three
// weight blocks): .collector::a::fill / ::use / ::lastuse keep A in the collector buffer, so only the first reads it from smem.
__device__ __forceinline__ void umma_bf16_coll(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, int which, uint32_t acc) {
  if (which == 0)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16.collector::a::fill [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
  else if (which == 1)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16.collector::a::use [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16.collector::a::lastuse [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}
__global__ void probe_reuse(int N, int iters, int reuse, long long* cycles, float* dout) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3c003c00u + (uint32_t)((i * 2654435761u) >> 28) * 0x00100010u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (threadIdx.x < 32) tmem_alloc(&tmem_base, 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tb = tmem_base;
  const int warp_idx = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  if (warp_idx == 0 && elect_one()) {
    const uint32_t idesc = umma_idesc_bf16(128, N);
    const uint64_t ad0 = umma_desc(smem_u32(smem), 16 * 1024, 128);
    const uint64_t bd0 = umma_desc(smem_u32(smem) + 128 * 1024, N * 16, 128);
    const uint32_t a_hi = (uint32_t)(ad0 >> 32), b_hi = (uint32_t)(bd0 >> 32);
    const uint32_t a_lo0 = (uint32_t)ad0, b_lo0 = (uint32_t)bd0;
    const uint32_t bstep = (uint32_t)(N * 32) >> 4;
    long long t0 = clock64();
    uint32_t phase = 0;
    for (int it = 0; it < iters; ++it) {
      uint32_t a_lo = a_lo0;
#pragma unroll
      for (int j = 0; j < 21; ++j) {          // 21 triples = 63 MMAs
        const uint64_t a = ((uint64_t)a_hi << 32) | a_lo;
        const uint32_t acc = j > 0 ? 1u : 0u;     // every iteration starts its three accumulators afresh
        if (reuse) {
          umma_bf16_coll(tb + 0 * 128, a, ((uint64_t)b_hi << 32) | (b_lo0 + 0 * bstep), idesc, 0, acc);
          umma_bf16_coll(tb + 1 * 128, a, ((uint64_t)b_hi << 32) | (b_lo0 + 1 * bstep), idesc, 1, acc);
          umma_bf16_coll(tb + 2 * 128, a, ((uint64_t)b_hi << 32) | (b_lo0 + 2 * bstep), idesc, 2, acc);
        } else {
          umma_bf16(tb + 0 * 128, a, ((uint64_t)b_hi << 32) | (b_lo0 + 0 * bstep), idesc, acc);
          umma_bf16(tb + 1 * 128, a, ((uint64_t)b_hi << 32) | (b_lo0 + 1 * bstep), idesc, acc);
          umma_bf16(tb + 2 * 128, a, ((uint64_t)b_hi << 32) | (b_lo0 + 2 * bstep), idesc, acc);
        }
        a_lo += 65;
      }
      umma_commit(&bar);
      mbar_wait(&bar, phase);
      phase ^= 1;
    }
    long long t1 = clock64();
    cycles[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  // read one accumulator back so both variants can be compared for equality (lane = row, first 8 columns of accumulator 1)
  if (threadIdx.x < 128 && dout) {
    uint32_t v[8];
    tc_fence_after();
    tmem_ld_x8(tb + ((uint32_t)(threadIdx.x & ~31u) << 16) + 128, v);
    tmem_ld_wait();
    for (int i = 0; i < 8; ++i) dout[(blockIdx.x * 128 + threadIdx.x) * 8 + i] = __uint_as_float(v[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tb, 512);
}

static float bf(float x) { return __bfloat162float(__float2bfloat16(x)); }

static bool run_correct(int N, int row_off, int lbo_rows_extra, const char* name) {
  // A: 2 "planes" each a_rows x 8; the second K half lives at +a_lbo bytes from the first
  const int a_rows = 192;
  std::vector<__nv_bfloat16> hA(2 * a_rows * 8), hB(2 * N * 8);
  std::vector<float> fA(hA.size()), fB(hB.size());
  for (size_t i = 0; i < hA.size(); ++i) { float v = (float)((rand() % 17) - 8) / 8.f; hA[i] = __float2bfloat16(v); fA[i] = bf(v); }
  for (size_t i = 0; i < hB.size(); ++i) { float v = (float)((rand() % 13) - 6) / 4.f; hB[i] = __float2bfloat16(v); fB[i] = bf(v); }
  int a_lbo = a_rows * 16 + lbo_rows_extra * 16;   // plane stride (+ extra row shift for the 2nd half)
  __nv_bfloat16 *dA, *dB; float* dD;
  CK(cudaMalloc(&dA, hA.size() * 2)); CK(cudaMalloc(&dB, hB.size() * 2)); CK(cudaMalloc(&dD, 128 * N * 4));
  CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
  size_t smem = 2 * a_rows * 16 + 2 * N * 16 + 1024;
  CK(cudaFuncSetAttribute(probe_correct, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  probe_correct<<<1, 128, smem>>>(dA, dB, dD, a_rows, N, row_off, a_lbo);
  CK(cudaDeviceSynchronize());
  std::vector<float> hD(128 * N);
  CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost));
  double maxerr = 0;
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < N; ++n) {
      double ref = 0;
      for (int k = 0; k < 16; ++k) {
        int half = k / 8, kk = k % 8;
        int arow = m + row_off + (half ? lbo_rows_extra : 0);
        float a = fA[(size_t)half * a_rows * 8 + (size_t)arow * 8 + kk];
        float b = fB[(size_t)half * N * 8 + (size_t)n * 8 + kk];
        ref += (double)a * b;
      }
      maxerr = fmax(maxerr, fabs(ref - hD[m * N + n]));
    }
  printf("correct[%s] N=%d row_off=%d lbo_extra=%d: max err %.3g %s\n", name, N, row_off, lbo_rows_extra, maxerr, maxerr < 1e-3 ? "OK" : "FAIL");
  cudaFree(dA); cudaFree(dB); cudaFree(dD);
  return maxerr < 1e-3;
}

int main() {
  bool ok = true;
  ok &= run_correct(48, 0, 0, "base");
  ok &= run_correct(80, 0, 0, "n80");
  ok &= run_correct(16, 0, 0, "n16");
  ok &= run_correct(48, 3, 0, "shift3");
  ok &= run_correct(80, 47, 0, "shift47");
  ok &= run_correct(48, 5, 11, "lbo+11rows");
  ok &= run_correct(256, 1, 0, "n256");
  printf("CORRECTNESS %s\n", ok ? "PASS" : "FAIL");

  int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  long long* dcy; CK(cudaMalloc(&dcy, sms * 8));
  size_t smem = 200 * 1024;
  CK(cudaFuncSetAttribute(probe_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int Ns[] = {16, 48, 80, 128, 256};
  for (int grid : {1, sms}) {
    for (int N : Ns) {
      for (int mode = 0; mode < 3; ++mode) {
        // mode 0: same A, same B ; mode 1: A walks (+1040 B per MMA), B fixed ; mode 2: A walks, B walks (+N*32 B)
        uint32_t a_step = mode == 0 ? 0 : 65;                       // 16-byte units
        uint32_t b_step = mode == 2 ? (uint32_t)(N * 2) : 0;        // N*32 bytes; 64 steps x 8 KB max = 512 KB?? cap below
        if (b_step * 64 * 16 > 60 * 1024) b_step = 60 * 1024 / 64 / 16;
        const int iters = 50;
        probe_rate<<<grid, 128, smem>>>(N, iters, a_step, b_step, dcy);
        CK(cudaDeviceSynchronize());
        std::vector<long long> h(grid);
        CK(cudaMemcpy(h.data(), dcy, grid * 8, cudaMemcpyDeviceToHost));
        long long mx = 0; for (auto v : h) mx = v > mx ? v : mx;
        printf("rate grid=%3d N=%3d mode=%d: %.1f cycles/MMA (ideal N/2 = %d; A-read bound 32)\n", grid, N, mode,
               (double)mx / (iters * 64), N / 2);
      }
    }
  }
  // A-operand collector reuse: triples of MMAs on one A tile
  {
    CK(cudaFuncSetAttribute(probe_reuse, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    float* dd; CK(cudaMalloc(&dd, (size_t)sms * 128 * 8 * 4));
    for (int N : {16, 48, 80, 128}) {
      std::vector<float> ref;
      for (int reuse = 0; reuse < 2; ++reuse) {
        CK(cudaMemset(dd, 0, (size_t)sms * 128 * 8 * 4));
        probe_reuse<<<sms, 128, smem>>>(N, 50, reuse, dcy, dd);
        CK(cudaDeviceSynchronize());
        std::vector<long long> h(sms);
        CK(cudaMemcpy(h.data(), dcy, sms * 8, cudaMemcpyDeviceToHost));
        std::vector<float> out((size_t)sms * 128 * 8);
        CK(cudaMemcpy(out.data(), dd, out.size() * 4, cudaMemcpyDeviceToHost));
        long long mx = 0; for (auto v : h) mx = v > mx ? v : mx;
        bool same = true;
        if (reuse == 0) ref = out; else for (size_t i = 0; i < out.size(); ++i) same &= (out[i] == ref[i]);
        printf("collector N=%3d %s: %.1f cycles/MMA%s (sample acc %.1f)\n", N, reuse ? "fill/use/lastuse" : "plain           ",
               (double)mx / (50 * 63), reuse ? (same ? "  results identical" : "  RESULTS DIFFER") : "", out[5]);
      }
    }
  }
  // wide N (three output rows per instruction) and accumulators that do not start at column 0
  for (int N : {96, 144, 192, 240}) {
    for (uint32_t d_off : {0u, 16u, 48u}) {
      const uint32_t d_alt = (N + d_off <= 256) ? 256u : 0u;
      probe_rate<<<sms, 128, smem>>>(N, 50, 65, 0, dcy, d_off, d_alt);
      CK(cudaDeviceSynchronize());
      std::vector<long long> h(sms);
      CK(cudaMemcpy(h.data(), dcy, sms * 8, cudaMemcpyDeviceToHost));
      long long mx = 0; for (auto v : h) mx = v > mx ? v : mx;
      printf("wide rate N=%3d d_col=%2u: %.1f cycles/MMA (N/2 = %d, (4096+32N)/128 = %d)\n", N, d_off, (double)mx / (50 * 64), N / 2,
             (4096 + 32 * N) / 128);
    }
  }
  return ok ? 0 : 1;
}
