// Device-side vocabulary shared by the tensor-core translation units (bf16_tc.cu: the layer kernels; mega.cu: the single
// fused pass): build-flavour macros, strip geometry, activation math, the compile-time epilogue chains, operand packing
// and the block-major MMA issue loop of a CTA pair.  Included inside each unit once; everything sits in an anonymous
// namespace, so every unit gets its own copy.
#pragma once

#include "engine.h"
#include "tc_ptx.cuh"

// This translation unit is compiled twice: as the bf16 build and, with -DFSUAE_OPERAND_FP16, as the fp16 build
// (same kernels, fp16 A/B operands: the precision the reference deploys, convertion_tools/torch2onnx.py:58, 358-412).
#ifdef FSUAE_OPERAND_FP16
#define TC_FN(name) fp16_##name
#define TC_FIELD fp16
#define TcPlan Fp16Plan
#define TC_VARIANT_NAME "fp16_tcgen05"
#define TC_BUILD_TAG "fp16 build: "
#else
#define TC_FN(name) bf16_##name
#define TC_FIELD bf16
#define TcPlan Bf16Plan
#define TC_VARIANT_NAME "bf16_tcgen05"
#define TC_BUILD_TAG "bf16 build: "
#endif

// Experiment switches that make a launch produce garbage (bound analysis of DESIGN.md section 5) exist only in builds
// with -DFSUAE_DEBUG_SWITCHES; the production library has no such code path.
#ifdef FSUAE_DEBUG_SWITCHES
#define FSUAE_DBG_BIT(P, bit) (((P).dbg & (bit)) != 0)
#else
#define FSUAE_DBG_BIT(P, bit) false
#endif

namespace fsuae {
namespace {

using namespace tc;

constexpr int MAXC = 128;           // widest N of one launch
constexpr int STRIP = 126;            // valid output columns per strip row
constexpr int MROWS = 128;            // MMA M = slots per strip row
constexpr int PLANE_ROW = MROWS * 16; // bytes of one plane of one strip row
constexpr int BORDER = 3;             // zero pixels baked in on every side of a plane (1 for a 3x3 layer, 2 for a fused pair / 5x5, 3 for 7x7)
__host__ __device__ constexpr int plane_width(int S) { return STRIP * S + 2 * BORDER; }   // slots per padded row (PW)
constexpr int SMEM_LIMIT = 232448 - 2048;
constexpr int EPI_WG = 4;                       // epilogue warpgroups; group g owns accumulator stage g
constexpr int NTHREADS = 64 + 128 * EPI_WG;

enum { EPI_STORE = 0, EPI_TAIL_SHUFFLE = 1, EPI_TAIL_PLAIN = 2 };

// ---- fast activation math for the bf16 build (error well below bf16 resolution) ----------------
__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_fast(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float sigmoid_fast(float x) { return fmaf(tanh_fast(0.5f * x), 0.5f, 0.5f); }

// Products that end an activation are __fmul_rn: a plain `*` may be contracted with the residual add or the next slot's
// subtraction into one FFMA, or not, depending on the surrounding code -- the layer kernels and the fused pass must agree bit
// for bit, and an unfused product is also what the reference computes.
__device__ __forceinline__ float act_rt(int op, float x, float p0, float p1) {
  switch (op) {
    case FSUAE_ACT_IDENTITY: return x;
    case FSUAE_ACT_RELU: return fmaxf(x, 0.f);
    case FSUAE_ACT_RELU6: return fminf(fmaxf(x, 0.f), 6.f);
    case FSUAE_ACT_TANH: return tanh_fast(x);
    case FSUAE_ACT_SIGMOID: return sigmoid_fast(x);
    case FSUAE_ACT_SILU: return __fmul_rn(x, sigmoid_fast(x));
    case FSUAE_ACT_MISH: {   // x * tanh(softplus(x)) = x * w / (w + 2) = x - 2x / (w + 2), w = e^x (e^x + 2)
      float n = __expf(x);                     // overflow is benign: d = inf -> 1/d = 0 -> x
      float d = fmaf(n, n + 2.f, 2.f);         // w + 2
      return fmaf(x * rcp_fast(d), -2.f, x);
    }
    case FSUAE_ACT_GELU: return __fmul_rn(0.5f * x, 1.f + erff(x * 0.70710678118654752440f));
    case FSUAE_ACT_ELU: return x > 0.f ? x : __fmul_rn(p0, __expf(x) - 1.f);
    case FSUAE_ACT_SOFTPLUS: {
      float bx = x * p0;
      return bx > p1 ? x : __fdividef(__logf(1.f + __expf(bx)), p0);
    }
    case FSUAE_ACT_LEAKY_RELU: return x >= 0.f ? x : __fmul_rn(p0, x);
    case FSUAE_ACT_PRELU: return fmaf(p0 - 1.f, fminf(x, 0.f), x);     // x + (slope - 1) * min(x, 0)
    case FSUAE_ACT_SCALED_TANH: return fmaf(tanh_fast(x), 0.5f, 0.5f);
    case FSUAE_ACT_TELU: return __fmul_rn(x, tanh_fast(__expf(x)));
    case FSUAE_ACT_SINLU: return __fmul_rn(sigmoid_fast(x), fmaf(p0, __sinf(p1 * x), x));
    case FSUAE_ACT_BIASED_RELU: return fmaxf(x - p0, 0.f);
    case FSUAE_ACT_BIASED_PRELU: {   // p1 holds (slope - 1), prepared on the host: y + (slope - 1) * min(y, 0)
      float y = x - p0;
      return fmaf(p1, fminf(y, 0.f), y);
    }
    default: return x;
  }
}

// run-time op-code applied to the 8 channels of one chunk: ONE (warp-uniform) switch per slot and chunk, the
// per-element work inside each case stays branch-free
__device__ __forceinline__ void act_rt8(int op, float (&t)[8], const float (&p0)[8], const float (&p1)[8]) {
#define FSUAE_CASE(OP) case OP: _Pragma("unroll") for (int i = 0; i < 8; ++i) t[i] = act_rt(OP, t[i], p0[i], p1[i]); break;
  switch (op) {
    FSUAE_CASE(FSUAE_ACT_RELU) FSUAE_CASE(FSUAE_ACT_RELU6) FSUAE_CASE(FSUAE_ACT_TANH) FSUAE_CASE(FSUAE_ACT_SIGMOID)
    FSUAE_CASE(FSUAE_ACT_SILU) FSUAE_CASE(FSUAE_ACT_MISH) FSUAE_CASE(FSUAE_ACT_GELU) FSUAE_CASE(FSUAE_ACT_ELU)
    FSUAE_CASE(FSUAE_ACT_SOFTPLUS) FSUAE_CASE(FSUAE_ACT_LEAKY_RELU) FSUAE_CASE(FSUAE_ACT_PRELU) FSUAE_CASE(FSUAE_ACT_SCALED_TANH)
    FSUAE_CASE(FSUAE_ACT_TELU) FSUAE_CASE(FSUAE_ACT_SINLU) FSUAE_CASE(FSUAE_ACT_BIASED_RELU) FSUAE_CASE(FSUAE_ACT_BIASED_PRELU)
    default: break;   // identity
  }
#undef FSUAE_CASE
}

// op-codes that take per-channel parameters (p0 / p1)
constexpr uint32_t ACT_PARAM_MASK = (1u << FSUAE_ACT_ELU) | (1u << FSUAE_ACT_SOFTPLUS) | (1u << FSUAE_ACT_LEAKY_RELU) |
                                    (1u << FSUAE_ACT_PRELU) | (1u << FSUAE_ACT_SINLU) | (1u << FSUAE_ACT_BIASED_RELU) |
                                    (1u << FSUAE_ACT_BIASED_PRELU);

// OP >= 0: compile-time op (the switch folds away); OP < 0: op-code read from the layer parameters
template <int OP, class PK>
__device__ __forceinline__ float act_slot(const PK& P, int slot, int ch, float x) {
  if constexpr (OP == FSUAE_ACT_IDENTITY) return x;
  else if constexpr (OP >= 0) return act_rt(OP, x, P.p0[slot][ch], P.p1[slot][ch]);
  else return act_rt(P.op[slot], x, P.p0[slot][ch], P.p1[slot][ch]);
}

// SKIP: 0 = no residual add, 1 = the layer's own input (src0) is added -> read from the centre row of
// the shared-memory ring (no second trip to global memory)
template <int PRE0, int PRE1, int POST0, int POST1, bool SKIP>
struct Epi {
  static constexpr bool kSkip = SKIP;
  static constexpr bool kRuntime = PRE0 < 0;      // op-codes come from the layer parameters
  static constexpr int kOp0 = PRE0, kOp1 = PRE1, kOp2 = POST0, kOp3 = POST1;
  template <class PK>      // PK: LayerK, or the fused pass's slimmer per-layer parameter block
  __device__ static __forceinline__ float pre(const PK& P, int ch, float v) {
    v = act_slot<PRE0>(P, 0, ch, v);
    return act_slot<PRE1>(P, 1, ch, v);
  }
  template <class PK>
  __device__ static __forceinline__ float post(const PK& P, int ch, float v) {
    v = act_slot<POST0>(P, 2, ch, v);
    return act_slot<POST1>(P, 3, ch, v);
  }
};

// two activations -> one 32-bit word of the operand type (bf16, or fp16 with -DFSUAE_OPERAND_FP16) and back
#ifdef FSUAE_OPERAND_FP16
__device__ __forceinline__ uint32_t pack_op2(float lo, float hi) {
  __half2 h = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float op_lo(uint32_t w) { return __half2float(__ushort_as_half((unsigned short)(w & 0xFFFFu))); }
__device__ __forceinline__ float op_hi(uint32_t w) { return __half2float(__ushort_as_half((unsigned short)(w >> 16))); }
#else
__device__ __forceinline__ uint32_t pack_op2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float op_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float op_hi(uint32_t w) { return __uint_as_float(w & 0xFFFF0000u); }
#endif

// ReLU (SIX: ReLU6) on a packed pair of operand-type values
template <bool SIX>
__device__ __forceinline__ uint32_t op2_relu(uint32_t w) {
#ifdef FSUAE_OPERAND_FP16
  __half2 h = *reinterpret_cast<__half2*>(&w);
  h = __hmax2(h, __float2half2_rn(0.f));
  if constexpr (SIX) h = __hmin2(h, __float2half2_rn(6.f));
  return *reinterpret_cast<uint32_t*>(&h);
#else
  __nv_bfloat162 h = *reinterpret_cast<__nv_bfloat162*>(&w);
  h = __hmax2(h, __float2bfloat162_rn(0.f));
  if constexpr (SIX) h = __hmin2(h, __float2bfloat162_rn(6.f));
  return *reinterpret_cast<uint32_t*>(&h);
#endif
}

__device__ __forceinline__ uint8_t to_u8_fast(float v, int gamma_out) {
  if (gamma_out) v = __powf(fmaxf(v, 0.f), 1.0f / 2.2f);
  v = fminf(fmaxf(v, 0.f), 1.f) * 255.0f;
  return (uint8_t)v;
}

// one strip row = 3 input rows x STEPS_ROW instructions; see conv3x3_tc_kernel for the unit/LBO pattern
template <int PT, int PR16, int NBROWS>
__device__ __forceinline__ void issue_block_2cta(uint32_t d_tmem, uint32_t ring_lo, uint32_t row16, uint32_t rs, uint32_t ring_n,
                                                 uint32_t w_lo, uint32_t idesc) {
  constexpr uint32_t HI = (uint32_t)((128u >> 4)) | (1u << 14);
  constexpr uint32_t BSTEP = (NBROWS * 32) >> 4;
  constexpr uint32_t L16 = 1u << 16, L2K = (uint32_t)(PR16 - 2) << 16;
  constexpr int STEPS_ROW = (3 * PT + 1) / 2, G3 = STEPS_ROW / 3, REM = STEPS_ROW % 3;
  static_assert(REM == 0 || REM == 2, "unexpected instruction count per row");
  uint32_t acc = 0, b_lo = w_lo;
#pragma unroll
  for (int dy = 0; dy < 3; ++dy) {
    uint32_t a_lo = ring_lo + rs * row16;
#pragma unroll 1
    for (int g3 = 0; g3 < G3; ++g3) {
      umma_bf16_2cta(d_tmem, ((uint64_t)HI << 32) | (a_lo | L16), ((uint64_t)HI << 32) | b_lo, idesc, acc);
      umma_bf16_2cta(d_tmem, ((uint64_t)HI << 32) | ((a_lo + 2) | L2K), ((uint64_t)HI << 32) | (b_lo + BSTEP), idesc, 1);
      umma_bf16_2cta(d_tmem, ((uint64_t)HI << 32) | ((a_lo + PR16 + 1) | L16), ((uint64_t)HI << 32) | (b_lo + 2 * BSTEP), idesc, 1);
      acc = 1;
      a_lo += 2 * PR16;
      b_lo += 3 * BSTEP;
    }
    if constexpr (REM == 2) {
      umma_bf16_2cta(d_tmem, ((uint64_t)HI << 32) | (a_lo | L16), ((uint64_t)HI << 32) | b_lo, idesc, acc);
      acc = 1;
      umma_bf16_2cta(d_tmem, ((uint64_t)HI << 32) | ((a_lo + 1) | L16), ((uint64_t)HI << 32) | (b_lo + BSTEP), idesc, 1);
      b_lo += 2 * BSTEP;
    }
    if (++rs == ring_n) rs = 0;
  }
}

}  // namespace
}  // namespace fsuae
