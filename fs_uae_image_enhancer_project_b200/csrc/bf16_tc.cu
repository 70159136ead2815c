// bf16 build: every 3x3 convolution runs as an implicit GEMM on the 5th-gen tensor cores
// (tcgen05.mma, kind::f16, bf16 operands, fp32 accumulators in TMEM) with the layer's whole
// epilogue fused behind the accumulator read-back (tcgen05.ld).
//
// Data layout ("chunk-planar"): an activation map with C channels is ceil(C/8) planes; a plane is
// (Hw+4) rows x PW slots of 16 bytes (8 bf16 channels of one pixel), with a zero border of two
// pixels baked in on every side so the per-layer zero padding of
// nn.Conv2d(padding=1) needs no special case anywhere: borders are never written.
//
// Implicit GEMM: M = 128 consecutive pixels of one image row (a "strip row": 126 valid outputs +
// 2 discarded), N = Cout padded to 16, K = 16 per instruction.  A strip row of every input plane
// is one contiguous 2 KB run in global memory, brought into a shared-memory ring by TMA bulk
// copies; in the no-swizzle K-major canonical layout one plane row IS a valid A operand
// (core matrix = 8 pixels x 16 B), and a 3x3 tap is nothing but a different start address
// (row slot +/- 1, +/- 16 bytes) -- no im2col, no data movement.  The two 8-channel halves of one
// K=16 instruction are any two (tap, channel-chunk) units; their distance is the descriptor's LBO.
// Weights for the whole layer stay resident in shared memory, pre-packed per instruction.
//
// Warp roles (576 threads, one persistent CTA per SM): warp 0 = TMA producer, warp 1 = MMA issuer
// (one elected lane), warps 2-17 = four epilogue warpgroups (thread = pixel, TMEM lane = pixel);
// warpgroup g drains every fourth strip row; there are up to 8 accumulator stages in TMEM, so the
// MMAs run several rows ahead of the activation math.
//
// Reference semantics: model/model_pix_shuffle.py:227-298 (and activations.py for the slots).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <climits>

#include "mega_params.h"
#include "tc_common.cuh"

namespace fsuae {

namespace {

// The layer kernel has a second MMA-issuing warp behind the epilogue warps (two issuers take the input rows in turn and
// hand the MMA stream to each other: the barrier polling / bookkeeping of one row runs under the other's issue phase).
#ifndef FSUAE_TWO_ISSUERS
#define FSUAE_TWO_ISSUERS 1
#endif
constexpr int ISSUER2_WARP = NTHREADS / 32;                     // warp 18
constexpr int NTHREADS_L = NTHREADS + (FSUAE_TWO_ISSUERS ? 32 : 0);


#ifndef FSUAE_ROW_MAJOR
#define FSUAE_ROW_MAJOR 1      // 1: input-row-major MMA order with A-collector reuse; 0: the older block-major order (A/B builds)
#endif

#ifdef FSUAE_EPI_TIMING
__device__ unsigned long long g_epi_timing[32];
__device__ __forceinline__ long long clk() { long long t; asm volatile("mov.u64 %0, %%clock64;" : "=l"(t) :: "memory"); return t; }
#define EPI_T(var) const long long var = clk()
#define EPI_ACC(i, v) do { if (blockIdx.x == 0 && warp == 2 && lane == 0) g_epi_timing[i] += (unsigned long long)(v); } while (0)
#define ISS_ACC(i, v) do { iss_t[i] += (unsigned long long)(v); } while (0)
#else
#define EPI_T(var)
#define EPI_ACC(i, v)
#define ISS_ACC(i, v)
#endif

struct LayerK {
  int Hw, Ww, PW, S, n_frames, n_blocks;   // n_blocks = n_frames * S * Hw strip rows, ordered (frame, strip, row)
  int P0, P1, cout;        // cout: channels this launch produces (a wide layer may be split over launches)
  int out_planes, dst_plane0, skip_plane0, tail;   // planes written / first plane in dst / first residual plane in the ring / FSUAE_TAIL_*
  unsigned long long fs0, fs1, fs_skip, fs_dst;  // frame strides in bytes
  const unsigned char* src0;
  const unsigned char* src1;
  const unsigned char* skip;
  unsigned char* dst;
  const unsigned char* wpack;
  const float* dparams;    // device copy of {bias[MAXC], p0[4][MAXC], p1[4][MAXC]} for the run-time epilogue (dynamic channel index)
  // tail (EPI_TAIL_SHUFFLE)
  const void* frame_in;
  void* frame_out;
  int in_fmt, out_fmt, H, W, xoff, gamma_in, gamma_out;
  // epilogue parameters, expanded per channel on the host
  int n_pre, n_post;
  int softmax_slot, softmax_log;   // run-time epilogue: slot (0..3) of a channel softmax / log_softmax, -1: none
  int dbg;                 // debugging aid (FSUAE_DBG): 1 = epilogue skips its global stores, 2 = producer re-reads the segment's first row, 4 = one TMA per row, 8 = no MMAs, 16 = all frames alias frame 0's activation buffers
  int op[4];               // pre0, pre1, post0, post1 (identity-padded)
  float bias[MAXC];
  float p0[4][MAXC];
  float p1[4][MAXC];
};

// CTAS = 2: CTA pair (cta_group::2), each CTA holds half of the weight rows.
// R3: "three output rows per instruction" mode (single CTA): one MMA multiplies an input row by the weights of all
// three kernel rows at once (N = 3 * NPAD) and accumulates into the three neighbouring output rows, whose
// accumulators sit side by side in a circular TMEM buffer.  The A operand (4 KB per instruction, the bottleneck of
// narrow-N UMMA: max(N/2, (4096 + 32 N)/128) cycles) is then read once per input row instead of three times.
template <int PT, int NPAD, int CTAS = 1, bool R3 = false>
struct Cfg {
  static constexpr int UNITS_ROW = 3 * PT;                 // (chunk, dx) units of one input row
  static constexpr int STEPS_ROW = (UNITS_ROW + 1) / 2;    // K=16 instructions per input row
  static constexpr int STEPS = 3 * STEPS_ROW;
  static constexpr int NB = NPAD / CTAS;                   // weight rows (output channels) resident in this CTA
  static constexpr int WBYTES = STEPS * NB * 32;
  static constexpr int ROWBYTES = PT * PLANE_ROW;
  static constexpr int RING_FIT = (SMEM_LIMIT - WBYTES - 64 - 512) / ROWBYTES;
  static constexpr int RING = RING_FIT > 10 ? 10 : RING_FIT;
  static constexpr int STAGES_CAP = R3 ? 12 : 10;     // the row-major order keeps three accumulators open: 10 stages leave 7 for the epilogues (N <= 48)
  static constexpr int STAGES = (512 / NPAD) > STAGES_CAP ? STAGES_CAP : (512 / NPAD);   // accumulator stages: more than warpgroups, so the MMAs run ahead of the (MUFU-bound) epilogues
  static constexpr int BAR_OFF = WBYTES + RING * ROWBYTES + 64;
  static constexpr int SMEM = BAR_OFF + 512;
  static_assert(RING >= 4, "layer does not fit: weights + 4 ring rows exceed shared memory");
  static_assert(STAGES * NPAD <= 512, "accumulator stages exceed TMEM");
  static_assert(NB % 8 == 0, "each CTA of a pair needs whole 8-row core matrices of B");
  static_assert(!R3 || (CTAS == 1 && 3 * NPAD <= 256 && STAGES >= 6), "R3 mode: single CTA, N = 3 NPAD <= 256, >= 6 accumulators");
};


// Run-time activation chain of one 8-channel chunk: o[] holds accumulator + bias on entry.  The four slots are walked
// by a ROLLED loop around one switch: unrolled, the compiler duplicates the 16-way switch along every path (a 49 KB
// loop body whose indirect branches miss the instruction cache: 1400 cycles per chunk measured); rolled it is 2 KB.
// Each slot fetches only the parameters its op-code needs (uniform addresses: broadcast loads).
//   ops = op0 | op1 << 8 | op2 << 16 | op3 << 24;  prm = dparams + first channel of the chunk;  stride = floats per row
__device__ __forceinline__ void act_chain_rt(uint32_t ops, const float* __restrict__ prm, int stride, bool has_skip,
                                             const uint4& skc, float (&o)[8], int s_begin = 0, int s_end = 4) {
#pragma unroll 1
  for (int s = s_begin; s < s_end; ++s) {
    if (s == 2 && has_skip) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint32_t w = (&skc.x)[i >> 1];
        o[i] += (i & 1) ? op_hi(w) : op_lo(w);
      }
    }
    const int op = (int)((ops >> (8 * s)) & 0xFFu);
    if (op == FSUAE_ACT_IDENTITY) continue;
    float p0[8], p1[8];
    if ((ACT_PARAM_MASK >> op) & 1u) {
      const float4 a0 = __ldg(reinterpret_cast<const float4*>(prm + (size_t)(1 + s) * stride));
      const float4 a1 = __ldg(reinterpret_cast<const float4*>(prm + (size_t)(1 + s) * stride + 4));
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(prm + (size_t)(5 + s) * stride));
      const float4 b1 = __ldg(reinterpret_cast<const float4*>(prm + (size_t)(5 + s) * stride + 4));
      p0[0] = a0.x; p0[1] = a0.y; p0[2] = a0.z; p0[3] = a0.w; p0[4] = a1.x; p0[5] = a1.y; p0[6] = a1.z; p0[7] = a1.w;
      p1[0] = b0.x; p1[1] = b0.y; p1[2] = b0.z; p1[3] = b0.w; p1[4] = b1.x; p1[5] = b1.y; p1[6] = b1.z; p1[7] = b1.w;
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) { p0[i] = 0.f; p1[i] = 0.f; }
    }
    act_rt8(op, o, p0, p1);
  }
}

// compile-time op-code of one slot (run-time channel count): parameters are fetched only if the op has any
template <int OP>
__device__ __forceinline__ void act_slot8(int s, const float* __restrict__ prm, int stride, float (&o)[8]) {
  if constexpr (OP != FSUAE_ACT_IDENTITY) {
    float p0[8], p1[8];
    if constexpr (((ACT_PARAM_MASK >> OP) & 1u) != 0) {
      const float4 a0 = __ldg(reinterpret_cast<const float4*>(prm + (size_t)(1 + s) * stride));
      const float4 a1 = __ldg(reinterpret_cast<const float4*>(prm + (size_t)(1 + s) * stride + 4));
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(prm + (size_t)(5 + s) * stride));
      const float4 b1 = __ldg(reinterpret_cast<const float4*>(prm + (size_t)(5 + s) * stride + 4));
      p0[0] = a0.x; p0[1] = a0.y; p0[2] = a0.z; p0[3] = a0.w; p0[4] = a1.x; p0[5] = a1.y; p0[6] = a1.z; p0[7] = a1.w;
      p1[0] = b0.x; p1[1] = b0.y; p1[2] = b0.z; p1[3] = b0.w; p1[4] = b1.x; p1[5] = b1.y; p1[6] = b1.z; p1[7] = b1.w;
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) { p0[i] = 0.f; p1[i] = 0.f; }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = act_rt(OP, o[i], p0[i], p1[i]);
  }
}
// activation chain + residual of one chunk for a run-time channel count: compile-time ops when the variant has them
template <class EPI>
__device__ __forceinline__ void epi_chain8(uint32_t ops, const float* __restrict__ prm, int stride, bool has_skip, const uint4& skc,
                                           float (&o)[8]) {
  if constexpr (EPI::kRuntime) {
    act_chain_rt(ops, prm, stride, has_skip, skc, o);
  } else {
    act_slot8<EPI::kOp0>(0, prm, stride, o);
    act_slot8<EPI::kOp1>(1, prm, stride, o);
    if (has_skip) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint32_t w = (&skc.x)[i >> 1];
        o[i] += (i & 1) ? op_hi(w) : op_lo(w);
      }
    }
    act_slot8<EPI::kOp2>(2, prm, stride, o);
    act_slot8<EPI::kOp3>(3, prm, stride, o);
  }
}

// Channel softmax / log_softmax slot (activations.py:91-92, dim=1) in a run-time chain.  Thread = pixel, so the
// reduction over channels is thread-local, but it spans every 8-channel chunk of the accumulator row: the row is read
// back three times (running max of the values entering the slot; sum of exponentials; normalise + rest of the chain +
// store), recomputing the cheap part of the chain each time.  `emit(c, o)` stores chunk c.
struct SoftmaxCfg { int slot, log_form; };
template <class LoadSkip, class Emit>
__device__ __forceinline__ void softmax_row_rt(uint32_t taddr, int nplanes, int cout, uint32_t ops, const float* __restrict__ prm0,
                                               int stride, bool has_skip, SoftmaxCfg sm, LoadSkip load_skip, Emit emit) {
  float mx = -3.0e38f, sum = 0.f;
  for (int pass = 0; pass < 3; ++pass) {
    for (int c = 0; c < nplanes; ++c) {
      const float* prm = prm0 + c * 8;
      const uint4 skc = load_skip(c);
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(prm)), b1 = __ldg(reinterpret_cast<const float4*>(prm + 4));
      uint32_t v[8];
      tmem_ld_x8(taddr + c * 8, v);
      tmem_ld_wait();
      float o[8];
      o[0] = __uint_as_float(v[0]) + b0.x; o[1] = __uint_as_float(v[1]) + b0.y; o[2] = __uint_as_float(v[2]) + b0.z;
      o[3] = __uint_as_float(v[3]) + b0.w; o[4] = __uint_as_float(v[4]) + b1.x; o[5] = __uint_as_float(v[5]) + b1.y;
      o[6] = __uint_as_float(v[6]) + b1.z; o[7] = __uint_as_float(v[7]) + b1.w;
      act_chain_rt(ops, prm, stride, has_skip, skc, o, 0, sm.slot);      // the residual is added on the way past slot 2 ...
      if (sm.slot == 2 && has_skip) {                                    // ... or here, when the softmax is the first slot behind it
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const uint32_t w = (&skc.x)[i >> 1];
          o[i] += (i & 1) ? op_hi(w) : op_lo(w);
        }
      }
      if (pass == 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) if (c * 8 + i < cout) mx = fmaxf(mx, o[i]);
      } else if (pass == 1) {
#pragma unroll
        for (int i = 0; i < 8; ++i) if (c * 8 + i < cout) sum += __expf(o[i] - mx);
      } else {
        const float inv = 1.0f / sum, lse = __logf(sum);
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = sm.log_form ? (o[i] - mx) - lse : __expf(o[i] - mx) * inv;
        act_chain_rt(ops, prm, stride, has_skip && sm.slot < 2, skc, o, sm.slot + 1, 4);
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = (c * 8 + i < cout) ? o[i] : 0.f;
        emit(c, o);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// work partition: the n_blocks strip rows are split into gridDim.x contiguous, equally long ranges
// (perfect balance for the persistent CTAs); a range is walked as segments that stay inside one
// strip of one frame.  A segment of `rows` output rows streams rows+2 input rows through the ring.
// Every warp role iterates the identical segment sequence.
// ------------------------------------------------------------------------------------------------
struct Seg { int f, s, y0, rows; };

// With CTA pairs the unit of work is a pair of frames: both CTAs of a pair walk the same segment
// sequence in lockstep (same ring slot, same TMEM stage), CTA `rank` on frame 2*fp + rank.
struct SegIter {
  int cur, end, ctas, rank;
  __device__ SegIter(const LayerK& P, int ctas_, int rank_) : ctas(ctas_), rank(rank_) {
    const int unit = blockIdx.x / ctas, units = gridDim.x / ctas;
    cur = (int)((long long)P.n_blocks * unit / units);
    end = (int)((long long)P.n_blocks * (unit + 1) / units);
  }
  __device__ bool next(const LayerK& P, Seg& g) {
    if (cur >= end) return false;
    const int per_frame = P.S * P.Hw;
    const int fp = cur / per_frame;
    g.f = min(fp * ctas + rank, P.n_frames - 1);   // odd frame count: the last pair computes its frame twice (same bits)
    const int r = cur - fp * per_frame;
    g.s = r / P.Hw;
    g.y0 = r - g.s * P.Hw;
    g.rows = min(P.Hw - g.y0, end - cur);
    cur += g.rows;
    return true;
  }
};

// ------------------------------------------------------------------------------------------------
// the layer kernel
// ------------------------------------------------------------------------------------------------
template <int PT, int NPAD, int COUT, int KIND, class EPI, int CTAS = 1, bool R3 = false>
__global__ void __launch_bounds__(NTHREADS_L, 1) conv3x3_tc_kernel(const __grid_constant__ LayerK P) {
  using C = Cfg<PT, NPAD, CTAS, R3>;
  // The input-row-major order keeps three accumulators open at a time: it needs TMEM stages to spare for the epilogues
  // (N <= 80: 6-8 stages).  Wider layers (4 stages) keep the block-major order, one accumulator per block
  // (measured: conv5 heavyweight's 64 -> 128 layer 49 us/frame block-major, 76 row-major).
  constexpr bool kRowMajor = FSUAE_ROW_MAJOR != 0 && C::STAGES >= 6;
  static_assert(!R3 || !EPI::kSkip, "R3 mode releases a ring row as soon as its MMAs are done: no residual from the ring");
  const uint32_t rank = CTAS == 2 ? cluster_ctarank() : 0u;   // rank 0 of a pair issues the MMAs
  constexpr int OUT_PLANES = COUT > 0 ? (COUT + 7) / 8 : 1;   // COUT <= 0: channel count is a run-time parameter
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* s_w = smem;
  uint8_t* s_ring = smem + C::WBYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::BAR_OFF);
  uint64_t* full = bars;                       // [RING]  TMA -> MMA
  uint64_t* empty = bars + C::RING;            // [RING]  MMA -> TMA
  uint64_t* tfull = bars + 2 * C::RING;        // [STAGES] MMA -> epilogue
  uint64_t* tempty = tfull + C::STAGES;        // [STAGES] epilogue -> MMA
  uint64_t* wbar = tempty + C::STAGES;         // weights landed
  uint64_t* pfull = wbar + 1;                  // [RING]  pair only, leader: the peer CTA's ring row has landed
  uint64_t* tok = pfull + C::RING;             // [2]     the MMA stream is handed from one issuing warp to the other
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tok + 2);
  __shared__ float s_lut[256];

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    // a ring row is released by the MMA commit and, when the residual is read from it, by the 4 epilogue warps
    for (int i = 0; i < C::RING; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], EPI::kSkip ? 5 : 1); mbar_init(&pfull[i], 1); }
    for (int i = 0; i < C::STAGES; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 4 * CTAS); }   // pair: both CTAs' epilogues
    mbar_init(wbar, 1);
    mbar_init(&tok[0], 1); mbar_init(&tok[1], 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    if constexpr (CTAS == 2) tmem_alloc_2cta(tmem_slot, 512); else tmem_alloc(tmem_slot, 512);
  }
  if (threadIdx.x >= 64 && threadIdx.x < 64 + 16) {   // zero the 64-byte overrun pad behind the ring
    reinterpret_cast<uint32_t*>(s_ring + C::RING * C::ROWBYTES)[threadIdx.x - 64] = 0u;
  }
  if constexpr (KIND == EPI_TAIL_SHUFFLE) {
    for (int i = threadIdx.x; i < 256; i += NTHREADS_L) {
      float t = (float)i * (1.0f / 255.0f);
      s_lut[i] = P.gamma_in ? powf(t, 2.2f) : t;
    }
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  if constexpr (CTAS == 2) cluster_sync_all();     // the peer's barriers exist before anyone signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // Programmatic dependent launch: let the next layer's grid start its prologue (barrier init, TMEM
  // allocation, weight fetch) on SMs as they drain; everything that touches activations written by the
  // previous layer sits behind griddepcontrol.wait.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  const size_t row_pitch = (size_t)P.PW * 16;                  // bytes of one padded image row of one plane
  const size_t plane_pitch = (size_t)(P.Hw + 2 * BORDER) * row_pitch;

  if (warp == 0) {
    // ======================= TMA producer =======================
    if (elect_one()) {
      mbar_arrive_expect_tx(wbar, C::WBYTES);
      tma_load_1d(s_w, P.wpack + (size_t)rank * C::WBYTES, C::WBYTES, wbar);   // constants: fetch before the dependency wait
      asm volatile("griddepcontrol.wait;" ::: "memory");
      uint32_t slot = 0, par = 1;   // waiting on parity 1 of a fresh barrier passes immediately
#ifdef FSUAE_EPI_TIMING
      unsigned long long prod_wait = 0, prod_rows = 0;
      const long long prod_t0 = clk();
#endif
      SegIter it(P, CTAS, (int)rank);
      Seg sg;
      while (it.next(P, sg)) {
        const int f = sg.f, s = sg.s, y0 = sg.y0, rows = sg.rows;
        const size_t org = (size_t)(y0 - 1 + BORDER) * row_pitch + (size_t)(s * STRIP - 1 + BORDER) * 16;   // pixel (y0-1, x0-1)
        const unsigned char* g0 = P.src0 + (size_t)f * P.fs0 + org;
        const unsigned char* g1 = P.P1 ? P.src1 + (size_t)f * P.fs1 + org : nullptr;
        for (int k = 0; k < rows + 2; ++k) {          // padded rows y0 .. y0+rows+1
#ifdef FSUAE_EPI_TIMING
          const long long tp0 = clk();
          mbar_wait(&empty[slot], par);
          prod_wait += (unsigned long long)(clk() - tp0); prod_rows += 1;
#else
          mbar_wait(&empty[slot], par);
#endif
          uint8_t* d = s_ring + slot * C::ROWBYTES;
          const int kk = FSUAE_DBG_BIT(P, 2) ? 0 : k;
          if (FSUAE_DBG_BIT(P, 4)) {      // timing experiment: one plane only (results are garbage)
            mbar_arrive_expect_tx(&full[slot], PLANE_ROW);
            tma_load_1d(d, g0 + (size_t)kk * row_pitch, PLANE_ROW, &full[slot]);
            if (++slot == C::RING) { slot = 0; par ^= 1; }
            continue;
          }
          mbar_arrive_expect_tx(&full[slot], C::ROWBYTES);
          for (int j = 0; j < P.P0; ++j)
            tma_load_1d(d + j * PLANE_ROW, g0 + (size_t)j * plane_pitch + (size_t)kk * row_pitch, PLANE_ROW, &full[slot]);
          for (int j = 0; j < P.P1; ++j)
            tma_load_1d(d + (P.P0 + j) * PLANE_ROW, g1 + (size_t)j * plane_pitch + (size_t)kk * row_pitch, PLANE_ROW, &full[slot]);
          if (++slot == C::RING) { slot = 0; par ^= 1; }
        }
      }
#ifdef FSUAE_EPI_TIMING
      if (blockIdx.x == 0) { g_epi_timing[8] += prod_wait; g_epi_timing[9] += prod_rows; g_epi_timing[10] += (unsigned long long)(clk() - prod_t0); }
#endif
    }
  } else if (warp == 1 && rank != 0) {
    // ======================= peer CTA of a pair: relay "my ring row has landed" to the leader =======================
    // One lane per ring slot, all in flight at once: a remote mbarrier arrive is a round trip over the cluster network,
    // and relayed one row after the other by a single thread it was what the leader's MMA issuers waited for most
    // (measured: ~1 800 of their ~4 000 cycles per row pair).
    {
      uint32_t total = 0;                 // input rows this CTA streams through its ring
      SegIter it(P, CTAS, (int)rank);
      Seg sg;
      while (it.next(P, sg)) total += (uint32_t)sg.rows + 2u;
      if (lane == 0) mbar_wait(wbar, 0);  // my half of the weights is part of every MMA the leader issues
      __syncwarp();
      const bool active = lane < C::RING;
      const uint32_t fills = active && total > (uint32_t)lane ? (total - (uint32_t)lane + C::RING - 1) / C::RING : 0u;   // fills of slot `lane`
      uint32_t done = 0, par = 0;
      const long long t0 = clock64();
      while (__any_sync(0xffffffffu, done < fills)) {
        if (done < fills && mbar_try_wait(&full[lane], par)) {
          mbar_arrive_cluster(&pfull[lane], 0);
          par ^= 1u;
          ++done;
        }
        if (clock64() - t0 > (1ll << 33)) __trap();     // bounded like every other wait
      }
    }
  } else if (warp == ISSUER2_WARP && (rank != 0 || !((kRowMajor || R3) && FSUAE_TWO_ISSUERS != 0))) {
    // the second issuing warp has no work in the peer CTA of a pair, nor in the single-issuer orders
  } else if (warp == 1 || warp == ISSUER2_WARP) {
    // ======================= MMA issuers (leader CTA of a pair, or the only CTA) =======================
    if (elect_one()) {
      const uint32_t me = warp == 1 ? 0u : 1u;       // issuer 0 takes the even input rows of the CTA's row sequence, issuer 1 the odd ones
      constexpr uint32_t IDESC = umma_idesc_op(MROWS * CTAS, NPAD);
      auto mma = [](uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
        if constexpr (CTAS == 2) umma_bf16_2cta(d, a, b, idesc, acc); else umma_bf16(d, a, b, idesc, acc);
      };
      auto commit = [](uint64_t* bar) {
        if constexpr (CTAS == 2) umma_commit_2cta(bar); else umma_commit(bar);
      };
      auto wait_row = [&](uint32_t slot, uint32_t par) {
        mbar_wait(&full[slot], par);
        if constexpr (CTAS == 2) mbar_wait(&pfull[slot], par);   // signal only: the data is consumed through the async proxy
      };
      const uint32_t ring_lo = (smem_u32(s_ring) & 0x3FFFFu) >> 4;
      const uint32_t w_lo = ((smem_u32(s_w) & 0x3FFFFu) >> 4) | ((uint32_t)((C::NB * 16) >> 4) << 16);   // LBO = one K half of this CTA's rows
      constexpr uint32_t HI = (uint32_t)((128u >> 4)) | (1u << 14);   // SBO = 128 B, descriptor version 1
      mbar_wait(wbar, 0);
      uint32_t wslot = 0, wpar = 0;       // next ring slot to wait for
      uint32_t stage = 0, spar = 1;       // accumulator stage / parity for tempty
      SegIter it(P, CTAS, (int)rank);
      Seg sg;
      if constexpr (R3) {
        // ---- three output rows per instruction ----
        // Input row k of a segment (image row y0-1+k) feeds output rows j = k-2 (kernel row 2), k-1 (row 1), k (row 0);
        // the B operand of a step is [W2 | W1 | W0] (3 NPAD rows), a sub-range of it is just a start-address offset.
        // Output row j (block number blk0 + j) accumulates in TMEM slot (blk0 + j) % STAGES; the slots of j-2..j are
        // adjacent except where the circular buffer wraps, there the instruction is split in two.  The newest row's
        // accumulator must be overwritten, not accumulated into, by its first instruction: that one is issued apart.
        constexpr uint32_t N3 = 3 * NPAD;
        constexpr uint32_t BSTEP3 = (N3 * 32) >> 4;
        const uint32_t w3_lo = ((smem_u32(s_w) & 0x3FFFFu) >> 4) | ((uint32_t)((N3 * 16) >> 4) << 16);
        uint32_t blk0 = 0, grow3 = 0, tokpar3 = 0;
        constexpr bool TWO3 = FSUAE_TWO_ISSUERS != 0;       // the two issuing warps take the input rows in turn, as in the row-major order below
        constexpr uint32_t IDESC0 = umma_idesc_op(MROWS, 0);             // N field (bits 17..22, N >> 3) added per run
        constexpr uint32_t IDN = (uint32_t)(NPAD >> 3) << 17;
        while (it.next(P, sg)) {
          const int rows = sg.rows;
          for (int k = 0; k < rows + 2; ++k) {
            const uint32_t r3 = grow3 + (uint32_t)k;         // position in the CTA's input-row sequence = ring fill number
            if (TWO3 && (r3 & 1u) != me) continue;
            const uint32_t rs = r3 % C::RING;
            wait_row(rs, (r3 / C::RING) & 1u);
            const uint32_t nb = blk0 + (uint32_t)k;
            const uint32_t sk = nb % C::STAGES, pk = ((nb / C::STAGES) & 1u) ^ 1u;      // stage / tempty parity of block blk0 + k
            const int jlo = max(k - 2, 0), jhi = min(k, rows - 1);
            const bool has_new = k <= rows - 1;
            if (has_new) mbar_wait(&tempty[sk], pk);
            if (TWO3 && r3 > 0) {
              mbar_wait(&tok[me], tokpar3);
              tokpar3 ^= 1u;
            }
            tc_fence_after();
            const int back = k - jlo;                                           // 0..2 older rows in flight
            const uint32_t slot_lo = sk >= (uint32_t)back ? sk - (uint32_t)back : sk + C::STAGES - (uint32_t)back;
            const uint32_t boff0 = (uint32_t)((jlo - (k - 2)) * NPAD);     // first B row of the run, in 16-byte units
            // steps >= 1 (and step 0 of the last two input rows): output rows jlo..jhi, split where the accumulator ring wraps
            const int cnt = jhi - jlo + 1;
            const int c0 = min(cnt, (int)(C::STAGES - slot_lo)), c1 = cnt - c0;
            const uint32_t d0 = tmem_base + slot_lo * NPAD, d1 = tmem_base;
            const uint32_t id0 = IDESC0 + (uint32_t)c0 * IDN, id1 = IDESC0 + (uint32_t)c1 * IDN;
            const uint32_t bo1 = boff0 + (uint32_t)(c0 * NPAD);
            // step 0 of a row that opens a new accumulator: rows jlo..k-1 accumulate, row k is overwritten
            const int co0 = min(back, (int)(C::STAGES - slot_lo)), co1 = back - co0;
            const uint32_t ido0 = IDESC0 + (uint32_t)co0 * IDN, ido1 = IDESC0 + (uint32_t)co1 * IDN;
            const uint32_t boo1 = boff0 + (uint32_t)(co0 * NPAD);
            const uint32_t dn = tmem_base + sk * NPAD, bon = 2u * NPAD;
            const uint32_t s_done = sk >= 2u ? sk - 2u : sk + C::STAGES - 2u;   // slot of output row k-2
            const uint32_t a_row = ring_lo + rs * (C::ROWBYTES >> 4);
#pragma unroll
            for (int st = 0; st < C::STEPS_ROW; ++st) {
              // units 2 st, 2 st + 1 of the row (unit u: plane u / 3, tap u % 3); odd tail: units 3 PT - 2 (zero weights), 3 PT - 1
              const int u0 = (2 * st + 1 < 3 * PT) ? 2 * st : 3 * PT - 2;
              const int pl = u0 / 3, dx = u0 - 3 * pl;
              const uint32_t lbo = dx == 2 ? (uint32_t)((PLANE_ROW >> 4) - 2) : 1u;
              const uint64_t a_desc = ((uint64_t)HI << 32) | ((a_row + (uint32_t)(pl * (PLANE_ROW >> 4) + dx)) | (lbo << 16));
              const uint64_t b_hi = (uint64_t)HI << 32;
              const uint32_t b_st = w3_lo + (uint32_t)st * BSTEP3;
              if (st == 0 && has_new) {
                if (co0 > 0) umma_bf16(d0, a_desc, b_hi | (b_st + boff0), ido0, 1u);
                if (co1 > 0) umma_bf16(d1, a_desc, b_hi | (b_st + boo1), ido1, 1u);
                umma_bf16(dn, a_desc, b_hi | (b_st + bon), IDESC0 + IDN, 0u);
              } else {
                umma_bf16(d0, a_desc, b_hi | (b_st + boff0), id0, 1u);
                if (c1 > 0) umma_bf16(d1, a_desc, b_hi | (b_st + bo1), id1, 1u);
              }
            }
            commit(&empty[rs]);                                              // this input row is not needed again
            if (k >= 2) commit(&tfull[s_done]);                               // output row k-2 is complete
            if (TWO3) {
              tc_fence_before();
              mbar_arrive(&tok[me ^ 1u]);
            }
          }
          blk0 += (uint32_t)rows;
          grow3 += (uint32_t)rows + 2u;
        }
      } else if constexpr (kRowMajor) {
        // ---- input-row-major issue order with A-collector reuse ----
        // Input row k of a segment feeds output rows j = k-2 (kernel row 2), k-1 (row 1), k (row 0).  For every step the
        // three MMAs share the A tile (fill / use / lastuse), so it is read from shared memory once instead of three
        // times -- the A read is what bounds narrow-N UMMA.  Each output row keeps its own accumulator stage
        // (block n = blk0 + j in stage n % STAGES): no contiguity constraint, accumulate flags per instruction.
        constexpr uint32_t BSTEP = (C::NB * 32) >> 4;
        constexpr bool TWO = FSUAE_TWO_ISSUERS != 0;
#ifdef FSUAE_EPI_TIMING
        unsigned long long iss_t[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#endif
        uint32_t blk0 = 0, grow = 0, tokpar = 0;       // blocks / input rows of the CTA before this segment; parity of my token waits
        while (it.next(P, sg)) {
          const int rows = sg.rows;
          for (int k = 0; k < rows + 2; ++k) {
            const uint32_t r = grow + (uint32_t)k;           // position in the CTA's input-row sequence = ring fill number
            if (TWO && (r & 1u) != me) continue;
            const uint32_t rs = r % C::RING;
            EPI_T(ti0);
#ifdef FSUAE_EPI_TIMING
            mbar_wait(&full[rs], (r / C::RING) & 1u);
            EPI_T(ti0b);
            if constexpr (CTAS == 2) mbar_wait(&pfull[rs], (r / C::RING) & 1u);
            ISS_ACC(5, ti0b - ti0);
#else
            wait_row(rs, (r / C::RING) & 1u);
#endif
            EPI_T(ti1);
            const uint32_t n = blk0 + (uint32_t)k;           // block (output row) this input row opens
            const uint32_t sk = n % C::STAGES, pk = ((n / C::STAGES) & 1u) ^ 1u;
            const bool v0 = k <= rows - 1, v1 = k >= 1 && k <= rows, v2 = k >= 2;
            if (v0) mbar_wait(&tempty[sk], pk);
            EPI_T(ti2);
            if (TWO && r > 0) {                               // my turn: the other issuer has issued the previous row
              mbar_wait(&tok[me], tokpar);
              tokpar ^= 1u;
            }
            tc_fence_after();
            EPI_T(ti3);
            ISS_ACC(0, ti1 - ti0); ISS_ACC(1, ti2 - ti1); ISS_ACC(2, ti3 - ti2); ISS_ACC(7, 1);
            const uint32_t s1 = sk >= 1u ? sk - 1u : sk + C::STAGES - 1u, s2 = sk >= 2u ? sk - 2u : sk + C::STAGES - 2u;
            const uint32_t d0 = tmem_base + sk * NPAD, d1 = tmem_base + s1 * NPAD, d2 = tmem_base + s2 * NPAD;
            const uint32_t a_row = ring_lo + rs * (C::ROWBYTES >> 4);
            const uint64_t hi = (uint64_t)HI << 32;
            if (FSUAE_DBG_BIT(P, 8)) {
              // timing experiment: no MMAs at all (garbage results): what the barrier / commit chain alone costs
            } else if (v0 && v1 && v2) {
#pragma unroll
              for (int st = 0; st < C::STEPS_ROW; ++st) {
                const int u0 = (2 * st + 1 < 3 * PT) ? 2 * st : 3 * PT - 2;      // odd tail: units 3 PT - 2 (zero weights), 3 PT - 1
                const int pl = u0 / 3, dx = u0 - 3 * pl;
                const uint32_t lbo = dx == 2 ? (uint32_t)((PLANE_ROW >> 4) - 2) : 1u;
                const uint64_t a_desc = hi | ((a_row + (uint32_t)(pl * (PLANE_ROW >> 4) + dx)) | (lbo << 16));
                const uint32_t b_st = w_lo + (uint32_t)st * BSTEP;
                umma_bf16_coll<CTAS, 1>(d2, a_desc, hi | (b_st + 2u * C::STEPS_ROW * BSTEP), IDESC, 1u);
                umma_bf16_coll<CTAS, 2>(d1, a_desc, hi | (b_st + 1u * C::STEPS_ROW * BSTEP), IDESC, 1u);
                umma_bf16_coll<CTAS, 3>(d0, a_desc, hi | b_st, IDESC, st == 0 ? 0u : 1u);
              }
            } else {
              // the first and last two input rows of a segment feed fewer than three output rows
#pragma unroll
              for (int st = 0; st < C::STEPS_ROW; ++st) {
                const int u0 = (2 * st + 1 < 3 * PT) ? 2 * st : 3 * PT - 2;
                const int pl = u0 / 3, dx = u0 - 3 * pl;
                const uint32_t lbo = dx == 2 ? (uint32_t)((PLANE_ROW >> 4) - 2) : 1u;
                const uint64_t a_desc = hi | ((a_row + (uint32_t)(pl * (PLANE_ROW >> 4) + dx)) | (lbo << 16));
                const uint32_t b_st = w_lo + (uint32_t)st * BSTEP;
                if (v2) mma(d2, a_desc, hi | (b_st + 2u * C::STEPS_ROW * BSTEP), IDESC, 1u);
                if (v1) mma(d1, a_desc, hi | (b_st + 1u * C::STEPS_ROW * BSTEP), IDESC, 1u);
                if (v0) mma(d0, a_desc, hi | b_st, IDESC, st == 0 ? 0u : 1u);
              }
            }
            EPI_T(ti4);
            commit(&empty[rs]);                    // the MMAs are done with this input row (the pipe completes in issue order)
            if (v2) commit(&tfull[s2]);            // output row k-2 is complete
            if (TWO) {                             // hand the MMA stream over: tcgen05 ops of two threads are ordered by fence + sync
              tc_fence_before();
              mbar_arrive(&tok[me ^ 1u]);
            }
            EPI_T(ti5);
            ISS_ACC(3, ti4 - ti3); ISS_ACC(4, ti5 - ti4);
          }
          blk0 += (uint32_t)rows;
          grow += (uint32_t)rows + 2u;
        }
#ifdef FSUAE_EPI_TIMING
        if (blockIdx.x == 0 && me == 0)
          for (int i = 0; i < 8; ++i) g_epi_timing[16 + i] += iss_t[i];      // counters 16..23: the issuer's (0..7 belong to the epilogue probe)
#endif
      } else
      while (it.next(P, sg)) {
        const int rows = sg.rows;
        uint32_t s0 = wslot;              // slot of the block's first input row
        // first two rows of the item
        for (int k = 0; k < 2; ++k) {
          wait_row(wslot, wpar);
          if (++wslot == C::RING) { wslot = 0; wpar ^= 1; }
        }
        for (int b = 0; b < rows; ++b) {
          wait_row(wslot, wpar);
          if (++wslot == C::RING) { wslot = 0; wpar ^= 1; }
          mbar_wait(&tempty[stage], spar);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + stage * NPAD;
          // One input row = 3*PT (chunk, dx) units, two units per K=16 instruction.  Unit u sits at
          // (u/3) * 2 KB + (u%3) * 16 B inside the ring row, so every 3 instructions (= 2 chunks) the
          // pattern {+0 | lbo 16 B, +32 B | lbo 2 KB - 32 B, +2 KB + 16 B | lbo 16 B} repeats.  When 3*PT
          // is odd the last instruction re-reads unit 3*PT-2 against zero weights in its first half and
          // carries unit 3*PT-1 in its second, so both halves always point at landed, finite data --
          // anything else could turn never-written shared memory into 0 * NaN.
          constexpr uint32_t BSTEP = (C::NB * 32) >> 4;
          constexpr uint32_t L16 = 1u << 16, L2K = (uint32_t)((PLANE_ROW >> 4) - 2) << 16;
          constexpr int G3 = C::STEPS_ROW / 3, REM = C::STEPS_ROW % 3;
          static_assert(REM == 0 || REM == 2, "unexpected instruction count per row");
          uint32_t rs = s0, acc = 0, b_lo = w_lo;
#pragma unroll
          for (int dy = 0; dy < 3; ++dy) {
            uint32_t a_lo = ring_lo + rs * (C::ROWBYTES >> 4);
#pragma unroll 1
            for (int g3 = 0; g3 < G3; ++g3) {
              mma(d_tmem, ((uint64_t)HI << 32) | (a_lo | L16), ((uint64_t)HI << 32) | b_lo, IDESC, acc);
              mma(d_tmem, ((uint64_t)HI << 32) | ((a_lo + 2) | L2K), ((uint64_t)HI << 32) | (b_lo + BSTEP), IDESC, 1);
              mma(d_tmem, ((uint64_t)HI << 32) | ((a_lo + (PLANE_ROW >> 4) + 1) | L16), ((uint64_t)HI << 32) | (b_lo + 2 * BSTEP), IDESC, 1);
              acc = 1;
              a_lo += 2 * (PLANE_ROW >> 4);
              b_lo += 3 * BSTEP;
            }
            if constexpr (REM == 2) {
              mma(d_tmem, ((uint64_t)HI << 32) | (a_lo | L16), ((uint64_t)HI << 32) | b_lo, IDESC, acc);
              acc = 1;
              mma(d_tmem, ((uint64_t)HI << 32) | ((a_lo + 1) | L16), ((uint64_t)HI << 32) | (b_lo + BSTEP), IDESC, 1);
              b_lo += 2 * BSTEP;
            }
            if (++rs == C::RING) rs = 0;
          }
          commit(&tfull[stage]);
          commit(&empty[s0]);            // the block's first row is not needed again
          if (b == rows - 1) {                // item done: release its last two rows as well
            uint32_t s1 = s0 + 1 == C::RING ? 0 : s0 + 1;
            uint32_t s2 = s1 + 1 == C::RING ? 0 : s1 + 1;
            commit(&empty[s1]);
            commit(&empty[s2]);
          }
          if (++s0 == C::RING) s0 = 0;
          if (++stage == C::STAGES) { stage = 0; spar ^= 1; }
        }
      }
    }
  } else {
    // ======================= epilogue (warps 2..17) =======================
    const int q = warp & 3;                    // TMEM lane quadrant this warp may access
    const int m = q * 32 + lane;               // pixel index inside the strip row
    const uint32_t group = (uint32_t)(warp - 2) >> 2;   // warpgroup g drains blocks g, g+4, g+8, ...
    asm volatile("griddepcontrol.wait;" ::: "memory");   // epilogues read the frame / write buffers earlier layers still read
    // Block n lives in accumulator stage n % STAGES (use number n / STAGES).  A group's previous block is n-4 and
    // the MMAs commit in order, so when it waits for block n the stage's previous use (block n-STAGES) has long
    // been committed: the parity wait can never be a whole phase ahead.
    uint32_t blk = 0, qrow = 0;                // qrow: ring-row counter at the start of the item
    SegIter it(P, CTAS, (int)rank);
    Seg sg;
    while (it.next(P, sg)) {
      const int f = sg.f, s = sg.s, y0 = sg.y0, rows = sg.rows;
      const int x = s * STRIP + m;
      const bool valid = m < STRIP && x < P.Ww && !FSUAE_DBG_BIT(P, 1);
      for (int b = 0; b < rows; ++b, ++blk) {
        if ((blk & (EPI_WG - 1)) != group) continue;
        const uint32_t stage = blk % C::STAGES, spar = (blk / C::STAGES) & 1u;
        const int y = y0 + b;
        const size_t pix = (size_t)(y + BORDER) * row_pitch + (size_t)(x + BORDER) * 16;
        // tail: issue the loads of the input pixels (global residual) now, consume them only after the accumulator
        // is ready, so their latency hides behind the MMAs of this row
        uint32_t raw[KIND == EPI_TAIL_SHUFFLE ? 2 : 1][6];     // per output row dy: 2 pixels x 3 channels (f32 bits or u8)
        if constexpr (KIND == EPI_TAIL_SHUFFLE) {
          if (valid) {
            const size_t fpl = (size_t)P.H * P.W;
#pragma unroll
            for (int dy = 0; dy < 2; ++dy) {
              const size_t p0 = (size_t)(2 * y + dy) * P.W + 2 * x + P.xoff;
              if (P.in_fmt == FSUAE_FMT_F32_NCHW3) {
                const float* ip = (const float*)P.frame_in + (size_t)f * 3 * fpl + p0;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                  float2 t2 = __ldg(reinterpret_cast<const float2*>(ip + c * fpl));
                  raw[dy][2 * c] = __float_as_uint(t2.x); raw[dy][2 * c + 1] = __float_as_uint(t2.y);
                }
              } else if (P.in_fmt == FSUAE_FMT_U8_NHWC4) {
                uint2 t2 = __ldg(reinterpret_cast<const uint2*>((const unsigned char*)P.frame_in + ((size_t)f * fpl + p0) * 4));
                raw[dy][0] = t2.x; raw[dy][1] = t2.y;
              } else {
                const unsigned char* ip = (const unsigned char*)P.frame_in + (size_t)f * 4 * fpl + p0;
#pragma unroll
                for (int c = 0; c < 3; ++c) { raw[dy][2 * c] = __ldg(ip + c * fpl); raw[dy][2 * c + 1] = __ldg(ip + c * fpl + 1); }
              }
            }
          }
        }
        EPI_T(t_a);
        mbar_wait(&tfull[stage], spar);
        tc_fence_after();
        EPI_T(t_b);
        EPI_ACC(0, t_b - t_a); EPI_ACC(1, 1);
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + stage * NPAD;
        // residual = this layer's input at the same pixel = centre row of the block, still in the ring.
        // Only touch the ring barriers AFTER tfull: the block's MMAs have consumed rows kc-1..kc+1, so their
        // fills are complete and cannot be overtaken before we arrive on `empty` -- a waiter that ran a whole
        // phase ahead of an mbarrier would see its parity test pass on the wrong fill.
        const uint32_t kc = qrow + (uint32_t)b + 1;
        const uint32_t cslot = kc % C::RING;
        const uint8_t* sp = s_ring + cslot * C::ROWBYTES + (size_t)P.skip_plane0 * PLANE_ROW + (m + 1) * 16;
        if constexpr (EPI::kSkip) mbar_wait(&full[cslot], (kc / C::RING) & 1);   // already complete: acquires the TMA bytes
        auto release_rows = [&]() {
          __syncwarp();
          if (lane == 0) {
            mbar_arrive(&empty[cslot]);
            if (b == 0) mbar_arrive(&empty[(kc - 1) % C::RING]);          // first row of the segment: no block centres on it
            if (b == rows - 1) mbar_arrive(&empty[(kc + 1) % C::RING]);   // nor on the last one
          }
        };

        if constexpr (KIND == EPI_STORE && COUT > 0) {
          // ---- compile-time channel count (flagship preset): fully unrolled, parameters from the constant bank ----
          uint4 sk[EPI::kSkip ? OUT_PLANES : 1];
          if constexpr (EPI::kSkip) {
#pragma unroll
            for (int c = 0; c < OUT_PLANES; ++c)
              sk[c] = valid ? *reinterpret_cast<const uint4*>(sp + c * PLANE_ROW) : make_uint4(0, 0, 0, 0);
          }
          unsigned char* dp = P.dst + (size_t)f * P.fs_dst + pix + (size_t)P.dst_plane0 * plane_pitch;
#pragma unroll
          for (int c = 0; c < OUT_PLANES; ++c) {
            uint32_t v[8];
            tmem_ld_x8(taddr + c * 8, v);
            tmem_ld_wait();
            float o[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int ch = c * 8 + i;
              if (ch >= COUT) { o[i] = 0.f; continue; }     // padding channels stay exactly zero (and cost nothing)
              float t = EPI::pre(P, ch, __uint_as_float(v[i]) + P.bias[ch]);
              if constexpr (EPI::kSkip) {
                const uint32_t w = (&sk[c].x)[i >> 1];
                t += (i & 1) ? op_hi(w) : op_lo(w);
              }
              t = EPI::post(P, ch, t);
              o[i] = t;
            }
            if (valid) {
              uint4 pk = make_uint4(pack_op2(o[0], o[1]), pack_op2(o[2], o[3]), pack_op2(o[4], o[5]),
                                    pack_op2(o[6], o[7]));
              *reinterpret_cast<uint4*>(dp + (size_t)c * plane_pitch) = pk;
            }
          }
          // Release the ring rows only now: every residual value has been consumed, so no shared-memory load of
          // this row can still be in flight when the producer's TMA refills the slot (an arrive issued right
          // behind the LDS instructions does not wait for them).
          if constexpr (EPI::kSkip) release_rows();
        } else if constexpr (KIND == EPI_STORE) {
          // ---- run-time channel count / op-codes (every other network) ----
          // Per-channel parameters come from global memory here: indexing the kernel-parameter arrays with a
          // run-time channel would make the compiler spill the whole 4.7 KB parameter block to local memory.
          unsigned char* dp = P.dst + (size_t)f * P.fs_dst + pix + (size_t)P.dst_plane0 * plane_pitch;
          const uint32_t ops_packed = (uint32_t)P.op[0] | ((uint32_t)P.op[1] << 8) | ((uint32_t)P.op[2] << 16) | ((uint32_t)P.op[3] << 24);
          if (P.softmax_slot >= 0) {
            auto load_skip = [&](int c) {
              uint4 t = make_uint4(0, 0, 0, 0);
              if constexpr (EPI::kSkip) {
                if (valid) t = *reinterpret_cast<const uint4*>(sp + c * PLANE_ROW);
              } else if (P.skip != nullptr) {
                if (valid) t = __ldg(reinterpret_cast<const uint4*>(P.skip + (size_t)f * P.fs_skip + pix + (size_t)(P.skip_plane0 + c) * plane_pitch));
              }
              return t;
            };
            auto emit = [&](int c, const float (&o)[8]) {
              if (valid)
                *reinterpret_cast<uint4*>(dp + (size_t)c * plane_pitch) =
                    make_uint4(pack_op2(o[0], o[1]), pack_op2(o[2], o[3]), pack_op2(o[4], o[5]), pack_op2(o[6], o[7]));
            };
            softmax_row_rt(taddr, P.out_planes, P.cout, ops_packed, P.dparams, MAXC, EPI::kSkip || P.skip != nullptr,
                           SoftmaxCfg{P.softmax_slot, P.softmax_log}, load_skip, emit);
          } else
          for (int c = 0; c < P.out_planes; ++c) {
            uint4 skc = make_uint4(0, 0, 0, 0);
            if constexpr (EPI::kSkip) {
              if (valid) skc = *reinterpret_cast<const uint4*>(sp + c * PLANE_ROW);
            } else if (P.skip != nullptr) {      // residual from another buffer (1x1 skip projection): read from global memory
              if (valid) skc = __ldg(reinterpret_cast<const uint4*>(P.skip + (size_t)f * P.fs_skip + pix + (size_t)(P.skip_plane0 + c) * plane_pitch));
            }
            EPI_T(t_c);
            const float* prm = P.dparams + c * 8;
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(prm)), b1 = __ldg(reinterpret_cast<const float4*>(prm + 4));
            uint32_t v[8];
            EPI_T(t_d);
            tmem_ld_x8(taddr + c * 8, v);
            tmem_ld_wait();
            EPI_T(t_e);
            float o[8];
            o[0] = __uint_as_float(v[0]) + b0.x; o[1] = __uint_as_float(v[1]) + b0.y; o[2] = __uint_as_float(v[2]) + b0.z;
            o[3] = __uint_as_float(v[3]) + b0.w; o[4] = __uint_as_float(v[4]) + b1.x; o[5] = __uint_as_float(v[5]) + b1.y;
            o[6] = __uint_as_float(v[6]) + b1.z; o[7] = __uint_as_float(v[7]) + b1.w;
            EPI_T(t_f);
            EPI_ACC(2, t_d - t_c); EPI_ACC(3, t_e - t_d); EPI_ACC(4, t_f - t_e); EPI_ACC(7, 1);
            epi_chain8<EPI>(ops_packed, prm, MAXC, EPI::kSkip || P.skip != nullptr, skc, o);
            if (c * 8 + 8 > P.cout) {      // padding channels of the last plane stay exactly zero (uniform branch)
#pragma unroll
              for (int i = 0; i < 8; ++i) o[i] = (c * 8 + i < P.cout) ? o[i] : 0.f;
            }
            if (valid) {
              uint4 pk = make_uint4(pack_op2(o[0], o[1]), pack_op2(o[2], o[3]), pack_op2(o[4], o[5]),
                                    pack_op2(o[6], o[7]));
              *reinterpret_cast<uint4*>(dp + (size_t)c * plane_pitch) = pk;
            }
            EPI_T(t_g);
            EPI_ACC(5, t_g - t_f);
          }
          if constexpr (EPI::kSkip) release_rows();
          EPI_T(t_h);
          EPI_ACC(6, t_h - t_a);
        } else if constexpr (KIND == EPI_TAIL_PLAIN) {
          // ---- 3-channel full-resolution tail (conv5: as is; conv3: x255 + alpha) straight to the frame ----
          uint32_t v[8];
          tmem_ld_x8(taddr, v);
          tmem_ld_wait();
          float o[3];
#pragma unroll
          for (int ch = 0; ch < 3; ++ch) o[ch] = EPI::post(P, ch, EPI::pre(P, ch, __uint_as_float(v[ch]) + P.bias[ch]));
          if (valid) {
            const size_t fpl = (size_t)P.H * P.W;
            const size_t p0 = (size_t)y * P.W + x + P.xoff;
            if (P.out_fmt == FSUAE_FMT_F32_NCHW3) {
              float* op = (float*)P.frame_out + (size_t)f * 3 * fpl + p0;
              op[0] = o[0]; op[fpl] = o[1]; op[2 * fpl] = o[2];
            } else if (P.out_fmt == FSUAE_FMT_F32_NCHW4) {     // model_conv3.py:145-153
              float* op = (float*)P.frame_out + (size_t)f * 4 * fpl + p0;
              op[0] = o[0] * 255.0f; op[fpl] = o[1] * 255.0f; op[2 * fpl] = o[2] * 255.0f; op[3 * fpl] = 255.0f;
            } else {
              const uint32_t px = (uint32_t)to_u8_fast(o[0], P.gamma_out) | ((uint32_t)to_u8_fast(o[1], P.gamma_out) << 8) |
                                  ((uint32_t)to_u8_fast(o[2], P.gamma_out) << 16) | 0xFF000000u;
              *reinterpret_cast<uint32_t*>((unsigned char*)P.frame_out + ((size_t)f * fpl + p0) * 4) = px;
            }
          }
        } else {
          // PixelShuffle(2) + input residual + ReLU, straight to the output frame
          uint32_t v[16];
          tmem_ld_x16(taddr, v);
          tmem_ld_wait();
          float o[12];
          if constexpr (EPI::kRuntime) {
            // run-time op-codes: the chunk-wise chain (one switch per slot and 8 channels), not one switch per element
            const uint32_t ops_packed = (uint32_t)P.op[0] | ((uint32_t)P.op[1] << 8) | ((uint32_t)P.op[2] << 16) | ((uint32_t)P.op[3] << 24);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              float t8[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) t8[i] = __uint_as_float(v[h * 8 + i]) + P.bias[h * 8 + i];
              epi_chain8<EPI>(ops_packed, P.dparams + h * 8, MAXC, false, make_uint4(0, 0, 0, 0), t8);
#pragma unroll
              for (int i = 0; i < 8; ++i)
                if (h * 8 + i < 12) o[h * 8 + i] = t8[i];
            }
          } else {
#pragma unroll
            for (int ch = 0; ch < 12; ++ch) {
              float t = EPI::pre(P, ch, __uint_as_float(v[ch]) + P.bias[ch]);
              o[ch] = EPI::post(P, ch, t);
            }
          }
          if (valid) {
            const size_t fpl = (size_t)P.H * P.W;
            float idv[2][3][2];
#pragma unroll
            for (int dy = 0; dy < 2; ++dy)
#pragma unroll
              for (int c = 0; c < 3; ++c) {
                if (P.in_fmt == FSUAE_FMT_F32_NCHW3) {
                  idv[dy][c][0] = __uint_as_float(raw[dy][2 * c]); idv[dy][c][1] = __uint_as_float(raw[dy][2 * c + 1]);
                } else if (P.in_fmt == FSUAE_FMT_U8_NHWC4) {
                  idv[dy][c][0] = s_lut[(raw[dy][0] >> (8 * c)) & 0xFF]; idv[dy][c][1] = s_lut[(raw[dy][1] >> (8 * c)) & 0xFF];
                } else {
                  idv[dy][c][0] = s_lut[raw[dy][2 * c] & 0xFF]; idv[dy][c][1] = s_lut[raw[dy][2 * c + 1] & 0xFF];
                }
              }
#pragma unroll
            for (int dy = 0; dy < 2; ++dy) {
              const int Y = 2 * y + dy, X = 2 * x + P.xoff;
              const size_t p0 = (size_t)Y * P.W + X;
              float res[3][2];
#pragma unroll
              for (int c = 0; c < 3; ++c)
#pragma unroll
                for (int dx = 0; dx < 2; ++dx) res[c][dx] = fmaxf(o[c * 4 + dy * 2 + dx] + idv[dy][c][dx], 0.f);
              if (P.out_fmt == FSUAE_FMT_F32_NCHW3) {
                float* op = (float*)P.frame_out + (size_t)f * 3 * fpl + p0;
#pragma unroll
                for (int c = 0; c < 3; ++c) *reinterpret_cast<float2*>(op + c * fpl) = make_float2(res[c][0], res[c][1]);
              } else {
                uint32_t px[2];
#pragma unroll
                for (int dx = 0; dx < 2; ++dx)
                  px[dx] = (uint32_t)to_u8_fast(res[0][dx], P.gamma_out) | ((uint32_t)to_u8_fast(res[1][dx], P.gamma_out) << 8) |
                           ((uint32_t)to_u8_fast(res[2][dx], P.gamma_out) << 16) | 0xFF000000u;
                *reinterpret_cast<uint2*>((unsigned char*)P.frame_out + ((size_t)f * fpl + p0) * 4) = make_uint2(px[0], px[1]);
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if constexpr (CTAS == 2) mbar_arrive_cluster(&tempty[stage], 0); else mbar_arrive(&tempty[stage]);
        }
      }
      qrow += (uint32_t)rows + 2;
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (CTAS == 2) cluster_sync_all();     // nobody leaves while the peer may still signal or read
  if (warp == 1) {
    if constexpr (CTAS == 2) tmem_dealloc_2cta(tmem_base, 512); else tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------
// fused pair of layers A -> B in one kernel (CTA pairs only)
//
// Layer A's output never goes to global memory: its epilogue writes bf16 rows straight into the
// shared-memory ring layer B's MMAs read.  A computes 128 valid columns per strip row (x0-1 .. x0+126,
// from a 130-slot input row) and rows y0-1 .. y0+rows of a segment, zeroing everything outside the
// frame (B's conv must see zero padding, not A applied to padding).  Used for conv3 -> conv4 of the
// flagship: conv3 alone is bound by its 9-plane HBM write, conv4 by its activation epilogue; fused,
// the write disappears and A's MMAs run underneath B's epilogue.
//
// Block schedule of a segment with R rows:  A(0) A(1) A(2) B(0) A(3) B(1) ... A(R+1) B(R-1);
// every block takes the next TMEM stage / epilogue warpgroup in turn.
// ------------------------------------------------------------------------------------------------
template <int PA, int NA, int CA, int NBP, int CB>
struct FusedCfg {
  static constexpr int PB = (CA + 7) / 8;                   // B's input planes = A's output planes
  static constexpr int PLANE_ROW_A = (MROWS + 2) * 16;      // 130 slots per plane row of A's input
  static constexpr int STEPS_ROW_A = (3 * PA + 1) / 2, STEPS_A = 3 * STEPS_ROW_A;
  static constexpr int STEPS_ROW_B = (3 * PB + 1) / 2, STEPS_B = 3 * STEPS_ROW_B;
  static constexpr int NBA = NA / 2, NBB = NBP / 2;         // weight rows per CTA of the pair
  static constexpr int WBYTES_A = STEPS_A * NBA * 32, WBYTES_B = STEPS_B * NBB * 32;
  static constexpr int ROWBYTES_A = PA * PLANE_ROW_A, ROWBYTES_B = PB * PLANE_ROW;
  static constexpr int RA = 4, RB = 5;                      // ring depths
  static constexpr int LAG = 3;                             // B(j) is issued after A(j+LAG): A's epilogue for row j+2 has a whole block of slack
  static_assert(RB >= LAG + 2, "ringB must hold the rows B reads plus the rows A runs ahead");
  static constexpr int OFF_WB = WBYTES_A, OFF_RA = OFF_WB + WBYTES_B, OFF_RB = OFF_RA + RA * ROWBYTES_A;
  static constexpr int BAR_OFF = OFF_RB + RB * ROWBYTES_B + 64;
  static constexpr int SMEM = BAR_OFF + 1024;
  static constexpr int NSTAGE = NA > NBP ? NA : NBP;        // TMEM columns per accumulator stage
  static constexpr int STAGES = (512 / NSTAGE) > 8 ? 8 : (512 / NSTAGE);
  static_assert(SMEM <= SMEM_LIMIT + 1024, "fused pair does not fit in shared memory");
  static_assert(STAGES >= 6, "the warpgroup rotation below needs at least 6 accumulator stages");
  static_assert(OFF_RA % 16 == 0 && OFF_RB % 16 == 0 && NBA % 8 == 0 && NBB % 8 == 0, "operand alignment");
};


template <int PA, int NA, int CA, class EPIA, int NBP, int CB, class EPIB>
__global__ void __launch_bounds__(NTHREADS, 1)
conv3x3_tc_fused_pair_kernel(const __grid_constant__ LayerK A, const __grid_constant__ LayerK P) {
  using C = FusedCfg<PA, NA, CA, NBP, CB>;
  constexpr int OUT_PLANES_B = (CB + 7) / 8;
  const uint32_t rank = cluster_ctarank();
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* s_wa = smem;
  uint8_t* s_wb = smem + C::OFF_WB;
  uint8_t* s_ra = smem + C::OFF_RA;
  uint8_t* s_rb = smem + C::OFF_RB;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::BAR_OFF);
  uint64_t* fullA = bars;                      // [RA] TMA -> MMA
  uint64_t* emptyA = fullA + C::RA;            // [RA] MMA -> TMA
  uint64_t* pfullA = emptyA + C::RA;           // [RA] leader: peer's row landed
  uint64_t* fullB = pfullA + C::RA;            // [RB] A-epilogue (4 warps) -> MMA / residual readers
  uint64_t* emptyB = fullB + C::RB;            // [RB] MMA commit (+ 4 residual readers) -> A-epilogue
  uint64_t* pfullB = emptyB + C::RB;           // [RB] leader: peer's row written
  uint64_t* tfull = pfullB + C::RB;            // [STAGES]
  uint64_t* tempty = tfull + C::STAGES;        // [STAGES]
  uint64_t* wbar = tempty + C::STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wbar + 1);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < C::RA; ++i) { mbar_init(&fullA[i], 1); mbar_init(&emptyA[i], 1); mbar_init(&pfullA[i], 1); }
    for (int i = 0; i < C::RB; ++i) { mbar_init(&fullB[i], 4); mbar_init(&emptyB[i], EPIB::kSkip ? 5 : 1); mbar_init(&pfullB[i], 1); }
    for (int i = 0; i < C::STAGES; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 8); }
    mbar_init(wbar, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc_2cta(tmem_slot, 512);
  if (threadIdx.x >= 64 && threadIdx.x < 64 + 16) reinterpret_cast<uint32_t*>(s_rb + C::RB * C::ROWBYTES_B)[threadIdx.x - 64] = 0u;
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  const size_t row_pitch = (size_t)P.PW * 16;
  const size_t plane_pitch = (size_t)(P.Hw + 2 * BORDER) * row_pitch;

  if (warp == 0) {
    // ======================= TMA producer: A's input rows y0-2 .. y0+rows+1, columns x0-2 .. x0+127 =======================
    if (elect_one()) {
      mbar_arrive_expect_tx(wbar, C::WBYTES_A + C::WBYTES_B);
      tma_load_1d(s_wa, A.wpack + (size_t)rank * C::WBYTES_A, C::WBYTES_A, wbar);
      tma_load_1d(s_wb, P.wpack + (size_t)rank * C::WBYTES_B, C::WBYTES_B, wbar);
      asm volatile("griddepcontrol.wait;" ::: "memory");
      uint32_t slot = 0, par = 1;
      SegIter it(P, 2, (int)rank);
      Seg sg;
      while (it.next(P, sg)) {
        const unsigned char* g0 = A.src0 + (size_t)sg.f * A.fs0 + (size_t)(sg.y0 - 2 + BORDER) * row_pitch +
                                  (size_t)(sg.s * STRIP - 2 + BORDER) * 16;
        for (int k = 0; k < sg.rows + 4; ++k) {
          mbar_wait(&emptyA[slot], par);
          mbar_arrive_expect_tx(&fullA[slot], C::ROWBYTES_A);
          uint8_t* d = s_ra + slot * C::ROWBYTES_A;
          for (int j = 0; j < PA; ++j)
            tma_load_1d(d + j * C::PLANE_ROW_A, g0 + (size_t)j * plane_pitch + (size_t)k * row_pitch, C::PLANE_ROW_A, &fullA[slot]);
          if (++slot == C::RA) { slot = 0; par ^= 1; }
        }
      }
    }
  } else if (warp == 1 && rank != 0) {
    // ======================= peer: relay landed / written rows to the leader, in the leader's wait order =======================
    // One lane per ring slot (lanes 0.. ringA, lanes 16.. ringB), all in flight at once: no head-of-line blocking between
    // the two rings, and the remote arrives (a round trip over the cluster network each) overlap.
    {
      uint32_t totalA = 0, totalB = 0;    // rows through the rings per segment: rows + 4 input rows of A, rows + 2 output rows of A
      SegIter it(P, 2, (int)rank);
      Seg sg;
      while (it.next(P, sg)) { totalA += (uint32_t)sg.rows + 4u; totalB += (uint32_t)sg.rows + 2u; }
      if (lane == 0) mbar_wait(wbar, 0);
      __syncwarp();
      static_assert(C::RA <= 16 && C::RB <= 16, "one lane per ring slot");
      const bool isA = lane < 16;
      const uint32_t slot = isA ? (uint32_t)lane : (uint32_t)lane - 16u, R = isA ? (uint32_t)C::RA : (uint32_t)C::RB;
      const uint32_t total = isA ? totalA : totalB;
      const uint32_t fills = slot < R && total > slot ? (total - slot + R - 1) / R : 0u;
      uint64_t* src = isA ? &fullA[slot < R ? slot : 0] : &fullB[slot < R ? slot : 0];
      uint64_t* dst = isA ? &pfullA[slot < R ? slot : 0] : &pfullB[slot < R ? slot : 0];
      uint32_t done = 0, par = 0;
      const long long t0 = clock64();
      while (__any_sync(0xffffffffu, done < fills)) {
        if (done < fills && mbar_try_wait(src, par)) {
          mbar_arrive_cluster(dst, 0);
          par ^= 1u;
          ++done;
        }
        if (clock64() - t0 > (1ll << 33)) __trap();
      }
    }
  } else if (warp == 1) {
    // ======================= leader: MMA issue for both layers =======================
    if (elect_one()) {
      constexpr uint32_t IDA = umma_idesc_op(2 * MROWS, NA), IDB = umma_idesc_op(2 * MROWS, NBP);
      const uint32_t ra_lo = (smem_u32(s_ra) & 0x3FFFFu) >> 4, rb_lo = (smem_u32(s_rb) & 0x3FFFFu) >> 4;
      const uint32_t wa_lo = ((smem_u32(s_wa) & 0x3FFFFu) >> 4) | ((uint32_t)((C::NBA * 16) >> 4) << 16);
      const uint32_t wb_lo = ((smem_u32(s_wb) & 0x3FFFFu) >> 4) | ((uint32_t)((C::NBB * 16) >> 4) << 16);
      mbar_wait(wbar, 0);
      uint32_t wa = 0, wpa = 0;           // next ringA slot to wait for
      uint32_t wb = 0, wpb = 0;           // next ringB slot to wait for
      uint32_t nblk = 0, stage = 0, spar = 1;      // block counter; stage = n % STAGES, use number n / STAGES
      auto waitA = [&]() { mbar_wait(&fullA[wa], wpa); mbar_wait(&pfullA[wa], wpa); if (++wa == C::RA) { wa = 0; wpa ^= 1; } };
      auto waitB = [&]() { mbar_wait(&fullB[wb], wpb); mbar_wait(&pfullB[wb], wpb); if (++wb == C::RB) { wb = 0; wpb ^= 1; } };
      auto next_stage = [&]() { ++nblk; stage = nblk % C::STAGES; spar = ((nblk / C::STAGES) & 1u) ^ 1u; };
      SegIter it(P, 2, (int)rank);
      Seg sg;
      while (it.next(P, sg)) {
        const int rows = sg.rows;
        uint32_t a0 = wa, b0 = wb;        // slots of the first input row of the next A / B block
        for (int i = 0; i < rows + C::LAG; ++i) {
          // ---- A(i): image row y0-1+i from ringA rows i, i+1, i+2 ----
          if (i < rows + 2) {
          if (i == 0) { waitA(); waitA(); }
          waitA();
          mbar_wait(&tempty[stage], spar);
          tc_fence_after();
          issue_block_2cta<PA, C::PLANE_ROW_A / 16, C::NBA>(tmem_base + stage * C::NSTAGE, ra_lo, C::ROWBYTES_A >> 4, a0, C::RA, wa_lo, IDA);
          umma_commit_2cta(&tfull[stage]);
          umma_commit_2cta(&emptyA[a0]);
          if (i == rows + 1) {
            uint32_t s1 = a0 + 1 == C::RA ? 0 : a0 + 1, s2 = s1 + 1 == C::RA ? 0 : s1 + 1;
            umma_commit_2cta(&emptyA[s1]);
            umma_commit_2cta(&emptyA[s2]);
          }
          if (++a0 == C::RA) a0 = 0;
          next_stage();
          }
          // ---- B(i-LAG): image row y0+i-LAG from ringB rows i-LAG .. i-LAG+2 ----
          if (i >= C::LAG) {
            if (i == C::LAG) { waitB(); waitB(); }
            waitB();
            mbar_wait(&tempty[stage], spar);
            tc_fence_after();
            issue_block_2cta<C::PB, PLANE_ROW / 16, C::NBB>(tmem_base + stage * C::NSTAGE, rb_lo, C::ROWBYTES_B >> 4, b0, C::RB, wb_lo, IDB);
            umma_commit_2cta(&tfull[stage]);
            umma_commit_2cta(&emptyB[b0]);
            if (i == rows + C::LAG - 1) {
              uint32_t s1 = b0 + 1 == C::RB ? 0 : b0 + 1, s2 = s1 + 1 == C::RB ? 0 : s1 + 1;
              umma_commit_2cta(&emptyB[s1]);
              umma_commit_2cta(&emptyB[s2]);
            }
            if (++b0 == C::RB) b0 = 0;
            next_stage();
          }
        }
      }
    }
  } else {
    // ======================= epilogue warpgroups =======================
    const int q = warp & 3;
    const int m = q * 32 + lane;
    const uint32_t group = (uint32_t)(warp - 2) >> 2;
    asm volatile("griddepcontrol.wait;" ::: "memory");
    // Block n lives in accumulator stage n % STAGES (use number n / STAGES) and is drained by warpgroup
    // (n + n/2) & 3.  That rotation gives every group both A blocks (light) and B blocks (heavy epilogue) -- with
    // n & 3 the strict A B A B order would leave all of B's activation math to two groups -- and consecutive
    // blocks of one group are at most 5 apart (< STAGES), so when a group waits for block n the stage's previous
    // use (block n-STAGES) was committed before the group's own previous block: no parity wait can run a phase ahead.
    uint32_t blk = 0, qB = 0;                  // qB: ringB row counter at the start of the segment
    auto mine = [&](uint32_t n) { return ((n + (n >> 1)) & 3u) == group; };
    SegIter it(P, 2, (int)rank);
    Seg sg;
    while (it.next(P, sg)) {
      const int f = sg.f, s = sg.s, y0 = sg.y0, rows = sg.rows;
      for (int i = 0; i < rows + C::LAG; ++i) {
        // ---------------- A(i): write image row ya of A's output into ringB ----------------
        if (i < rows + 2)
        if (const uint32_t n = blk++; mine(n)) {
          const uint32_t stage = n % C::STAGES, spar = (n / C::STAGES) & 1u;
          const int ya = y0 - 1 + i, xa = s * STRIP - 1 + m;
          const bool inframe = ya >= 0 && ya < P.Hw && xa >= 0 && xa < P.Ww;
          const uint32_t kb = qB + (uint32_t)i, slot = kb % C::RB;
          mbar_wait(&tfull[stage], spar);
          tc_fence_after();
          mbar_wait(&emptyB[slot], ((kb / C::RB) & 1) ^ 1);      // the row that lived here has been consumed
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + stage * C::NSTAGE;
          uint8_t* dp = s_rb + slot * C::ROWBYTES_B + m * 16;
#pragma unroll
          for (int c = 0; c < C::PB; ++c) {
            uint32_t v[8];
            tmem_ld_x8(taddr + c * 8, v);
            tmem_ld_wait();
            float o[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const int ch = c * 8 + e;
              if (ch >= CA) { o[e] = 0.f; continue; }
              float t = EPIA::post(A, ch, EPIA::pre(A, ch, __uint_as_float(v[e]) + A.bias[ch]));
              o[e] = inframe ? t : 0.f;         // B's conv sees zero padding outside the frame
            }
            *reinterpret_cast<uint4*>(dp + c * PLANE_ROW) =
                make_uint4(pack_op2(o[0], o[1]), pack_op2(o[2], o[3]), pack_op2(o[4], o[5]), pack_op2(o[6], o[7]));
          }
          fence_proxy_async_smem();             // generic-proxy stores -> visible to the tensor core's reads
          tc_fence_before();
          __syncwarp();
          if (lane == 0) { mbar_arrive(&fullB[slot]); mbar_arrive_cluster(&tempty[stage], 0); }
        }
        // ---------------- B(i-LAG): image row y of B's output to global memory ----------------
        if (i >= C::LAG) {
          if (const uint32_t n = blk++; mine(n)) {
            const uint32_t stage = n % C::STAGES, spar = (n / C::STAGES) & 1u;
            const int b = i - C::LAG, y = y0 + b, x = s * STRIP + m;
            const bool valid = m < STRIP && x < P.Ww;
            const size_t pix = (size_t)(y + BORDER) * row_pitch + (size_t)(x + BORDER) * 16;
            mbar_wait(&tfull[stage], spar);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + stage * C::NSTAGE;
            const uint32_t kc = qB + (uint32_t)b + 1, cslot = kc % C::RB;     // centre row of the block
            uint4 sk[EPIB::kSkip ? OUT_PLANES_B : 1];
            if constexpr (EPIB::kSkip) {
              mbar_wait(&fullB[cslot], (kc / C::RB) & 1);         // acquire the other warps' row stores
              const uint8_t* sp = s_rb + cslot * C::ROWBYTES_B + (m + 1) * 16;
#pragma unroll
              for (int c = 0; c < OUT_PLANES_B; ++c)
                sk[c] = valid ? *reinterpret_cast<const uint4*>(sp + c * PLANE_ROW) : make_uint4(0, 0, 0, 0);
            }
            unsigned char* dp = P.dst + (size_t)f * P.fs_dst + pix;
#pragma unroll
            for (int c = 0; c < OUT_PLANES_B; ++c) {
              uint32_t v[8];
              tmem_ld_x8(taddr + c * 8, v);
              tmem_ld_wait();
              float o[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const int ch = c * 8 + e;
                if (ch >= CB) { o[e] = 0.f; continue; }
                float t = EPIB::pre(P, ch, __uint_as_float(v[e]) + P.bias[ch]);
                if constexpr (EPIB::kSkip) {
                  const uint32_t w = (&sk[c].x)[e >> 1];
                  t += (e & 1) ? op_hi(w) : op_lo(w);
                }
                o[e] = EPIB::post(P, ch, t);
              }
              if (valid)
                *reinterpret_cast<uint4*>(dp + (size_t)c * plane_pitch) =
                    make_uint4(pack_op2(o[0], o[1]), pack_op2(o[2], o[3]), pack_op2(o[4], o[5]), pack_op2(o[6], o[7]));
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if constexpr (EPIB::kSkip) {       // residual values consumed: release the ring rows (see conv3x3_tc_kernel)
                mbar_arrive(&emptyB[cslot]);
                if (b == 0) mbar_arrive(&emptyB[(kc - 1) % C::RB]);
                if (b == rows - 1) mbar_arrive(&emptyB[(kc + 1) % C::RB]);
              }
              mbar_arrive_cluster(&tempty[stage], 0);
            }
          }
        }
      }
      qB += (uint32_t)rows + 2;
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) tmem_dealloc_2cta(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------
// head: frame -> chunk-planar bf16 (PixelUnshuffle(2): 12 channels -> 2 planes, 4 zero channels)
// ------------------------------------------------------------------------------------------------
__global__ void head_unshuffle_bf16_kernel(const void* __restrict__ in, unsigned char* __restrict__ dst, int n_frames,
                                           int in_fmt, int H, int W, int xoff, int Hw, int Ww, int PW,
                                           unsigned long long fs_dst, int gamma_in) {
  __shared__ float lut[256];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) {
    float t = (float)i * (1.0f / 255.0f);
    lut[i] = gamma_in ? powf(t, 2.2f) : t;
  }
  __syncthreads();
  const size_t plane = (size_t)Hw * Ww, total = (size_t)n_frames * plane;
  const size_t fpl = (size_t)H * W;
  const size_t row_pitch = (size_t)PW * 16, plane_pitch = (size_t)(Hw + 2 * BORDER) * row_pitch;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int f = (int)(idx / plane);
    const int r = (int)(idx - (size_t)f * plane);
    const int h = r / Ww, w = r - h * Ww;
    float v[12];
#pragma unroll
    for (int dy = 0; dy < 2; ++dy) {
      const size_t p0 = (size_t)(2 * h + dy) * W + 2 * w + xoff;
      if (in_fmt == FSUAE_FMT_F32_NCHW3) {
        const float* ip = (const float*)in + (size_t)f * 3 * fpl + p0;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          float2 t2 = *reinterpret_cast<const float2*>(ip + c * fpl);
          v[c * 4 + dy * 2] = t2.x; v[c * 4 + dy * 2 + 1] = t2.y;
        }
      } else if (in_fmt == FSUAE_FMT_U8_NHWC4) {
        uint2 t2 = *reinterpret_cast<const uint2*>((const unsigned char*)in + ((size_t)f * fpl + p0) * 4);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          v[c * 4 + dy * 2] = lut[(t2.x >> (8 * c)) & 0xFF];
          v[c * 4 + dy * 2 + 1] = lut[(t2.y >> (8 * c)) & 0xFF];
        }
      } else {
        const unsigned char* ip = (const unsigned char*)in + (size_t)f * 4 * fpl + p0;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          v[c * 4 + dy * 2] = lut[ip[c * fpl]];
          v[c * 4 + dy * 2 + 1] = lut[ip[c * fpl + 1]];
        }
      }
    }
    unsigned char* dp = dst + (size_t)f * fs_dst + (size_t)(h + BORDER) * row_pitch + (size_t)(w + BORDER) * 16;
    *reinterpret_cast<uint4*>(dp) = make_uint4(pack_op2(v[0], v[1]), pack_op2(v[2], v[3]), pack_op2(v[4], v[5]),
                                               pack_op2(v[6], v[7]));
    *reinterpret_cast<uint4*>(dp + plane_pitch) = make_uint4(pack_op2(v[8], v[9]), pack_op2(v[10], v[11]), 0u, 0u);
  }
}

// head for the full-resolution families (conv3 / conv5): 3 channels -> one plane (5 zero channels)
__global__ void head_plain_bf16_kernel(const void* __restrict__ in, unsigned char* __restrict__ dst, int n_frames, int in_fmt,
                                       int H, int W, int xoff, int Hw, int Ww, int PW, unsigned long long fs_dst, int gamma_in) {
  __shared__ float lut[256];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) {
    float t = (float)i * (1.0f / 255.0f);
    lut[i] = gamma_in ? powf(t, 2.2f) : t;
  }
  __syncthreads();
  const size_t plane = (size_t)Hw * Ww, total = (size_t)n_frames * plane;
  const size_t fpl = (size_t)H * W;
  const size_t row_pitch = (size_t)PW * 16;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int f = (int)(idx / plane);
    const int r = (int)(idx - (size_t)f * plane);
    const int h = r / Ww, w = r - h * Ww;
    const size_t p0 = (size_t)h * W + w + xoff;
    float v[3];
    if (in_fmt == FSUAE_FMT_F32_NCHW3) {
      const float* ip = (const float*)in + (size_t)f * 3 * fpl + p0;
      v[0] = ip[0]; v[1] = ip[fpl]; v[2] = ip[2 * fpl];
    } else if (in_fmt == FSUAE_FMT_U8_NHWC4) {
      const uint32_t t = *reinterpret_cast<const uint32_t*>((const unsigned char*)in + ((size_t)f * fpl + p0) * 4);
      v[0] = lut[t & 0xFF]; v[1] = lut[(t >> 8) & 0xFF]; v[2] = lut[(t >> 16) & 0xFF];
    } else {
      const unsigned char* ip = (const unsigned char*)in + (size_t)f * 4 * fpl + p0;
      v[0] = lut[ip[0]]; v[1] = lut[ip[fpl]]; v[2] = lut[ip[2 * fpl]];
    }
    unsigned char* dp = dst + (size_t)f * fs_dst + (size_t)(h + BORDER) * row_pitch + (size_t)(w + BORDER) * 16;
    *reinterpret_cast<uint4*>(dp) = make_uint4(pack_op2(v[0], v[1]), pack_op2(v[2], 0.f), 0u, 0u);
  }
}

// feature-map networks (FSUAE_HEAD_FEATURES / FSUAE_TAIL_FEATURES): float [B,C,H,W] <-> chunk-planar operands; thread = (pixel, plane)
__global__ void head_features_kernel(const float* __restrict__ in, unsigned char* __restrict__ dst, int n_frames, int C, int Hw, int Ww,
                                     int PW, unsigned long long fs_dst) {
  const int planes = (C + 7) / 8;
  const size_t plane = (size_t)Hw * Ww, total = (size_t)n_frames * planes * plane;
  const size_t row_pitch = (size_t)PW * 16, plane_pitch = (size_t)(Hw + 2 * BORDER) * row_pitch;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const size_t r = idx % plane;
    const size_t t = idx / plane;
    const int j = (int)(t % planes), f = (int)(t / planes);
    const int h = (int)(r / Ww), w = (int)(r - (size_t)h * Ww);
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = j * 8 + k < C ? in[((size_t)f * C + j * 8 + k) * plane + r] : 0.f;
    unsigned char* dp = dst + (size_t)f * fs_dst + (size_t)j * plane_pitch + (size_t)(h + BORDER) * row_pitch + (size_t)(w + BORDER) * 16;
    *reinterpret_cast<uint4*>(dp) = make_uint4(pack_op2(v[0], v[1]), pack_op2(v[2], v[3]), pack_op2(v[4], v[5]), pack_op2(v[6], v[7]));
  }
}
__global__ void tail_features_kernel(const unsigned char* __restrict__ src, float* __restrict__ out, int n_frames, int C, int Hw, int Ww,
                                     int PW, unsigned long long fs_src) {
  const int planes = (C + 7) / 8;
  const size_t plane = (size_t)Hw * Ww, total = (size_t)n_frames * planes * plane;
  const size_t row_pitch = (size_t)PW * 16, plane_pitch = (size_t)(Hw + 2 * BORDER) * row_pitch;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const size_t r = idx % plane;
    const size_t t = idx / plane;
    const int j = (int)(t % planes), f = (int)(t / planes);
    const int h = (int)(r / Ww), w = (int)(r - (size_t)h * Ww);
    const uint4 q = *reinterpret_cast<const uint4*>(src + (size_t)f * fs_src + (size_t)j * plane_pitch + (size_t)(h + BORDER) * row_pitch +
                                                    (size_t)(w + BORDER) * 16);
#pragma unroll
    for (int k = 0; k < 8; ++k)
      if (j * 8 + k < C) {
        const uint32_t wd = (&q.x)[k >> 1];
        out[((size_t)f * C + j * 8 + k) * plane + r] = (k & 1) ? op_hi(wd) : op_lo(wd);
      }
  }
}

__global__ void black_columns_bf16_kernel(void* __restrict__ out, int n_frames, int out_fmt, int H, int W, int ncols) {
  const size_t total = (size_t)n_frames * H * ncols;
  const size_t fplane = (size_t)H * W;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    int x = (int)(i % ncols);
    size_t t = i / ncols;
    int y = (int)(t % H);
    int f = (int)(t / H);
    size_t pix = (size_t)y * W + x;
    if (out_fmt == FSUAE_FMT_U8_NHWC4) {
      ((uchar4*)out)[(size_t)f * fplane + pix] = make_uchar4(0, 0, 0, 255);
    } else {
      const int C = out_fmt == FSUAE_FMT_F32_NCHW4 ? 4 : 3;
      float* o = (float*)out + (size_t)f * C * fplane + pix;
      for (int c = 0; c < C; ++c) o[(size_t)c * fplane] = c == 3 ? 255.0f : 0.f;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// host: weight packing, kernel table, plan
// ------------------------------------------------------------------------------------------------

#ifdef FSUAE_OPERAND_FP16
uint16_t f2op(float f) {   // float -> IEEE half, round to nearest even (like __float2half_rn), subnormals kept, overflow -> inf
  uint32_t u;
  std::memcpy(&u, &f, 4);
  const uint32_t sign = (u >> 16) & 0x8000u;
  u &= 0x7FFFFFFFu;
  if (u >= 0x7F800000u) return (uint16_t)(sign | 0x7C00u | (u > 0x7F800000u ? 0x200u : 0u));
  if (u >= 0x477FF000u) return (uint16_t)(sign | 0x7C00u);                  // rounds to >= 65520 -> inf
  if (u < 0x33000001u) return (uint16_t)sign;                               // <= 2^-25 -> 0
  const int e = (int)(u >> 23) - 127;
  uint32_t man = (u & 0x7FFFFFu) | 0x800000u;
  int shift = e < -14 ? 13 + (-14 - e) : 13;                                // subnormal halves lose more bits
  uint32_t half_man = man >> shift;
  const uint32_t rem = man & ((1u << shift) - 1u), halfway = 1u << (shift - 1);
  if (rem > halfway || (rem == halfway && (half_man & 1u))) ++half_man;
  uint32_t out = e < -14 ? half_man : (((uint32_t)(e + 15) << 10) + (half_man - 0x400u));   // mantissa carry bumps the exponent
  return (uint16_t)(sign | out);
}
#else
uint16_t f2op(float f) {   // round to nearest even, like __float2bfloat16_rn
  uint32_t u;
  std::memcpy(&u, &f, 4);
  if ((u & 0x7F800000u) == 0x7F800000u) return (uint16_t)(u >> 16);
  u += 0x7FFFu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}
#endif

// B operand of instruction (dy, st): [2 halves][NPAD rows][8 k]; half h <-> unit u = 2 st + h (see the kernel for the odd tail),
// chunk j = u / 3 (plane of the concatenated sources), tap dx = u % 3, channels 8j .. 8j+7.
std::vector<uint16_t> pack_weights(const float* w, int cout, int cin0, int cin1, int P0, int P1, int NPAD, int ctas = 1) {
  const int PT = P0 + P1, cin = cin0 + cin1;
  const int steps_row = (3 * PT + 1) / 2;
  const int NB = NPAD / ctas;   // rows per CTA; CTA r of a pair holds output channels [r*NB, (r+1)*NB), stored [rank][step][half][NB][8]
  std::vector<uint16_t> out((size_t)3 * steps_row * NPAD * 16, 0);
  for (int dy = 0; dy < 3; ++dy)
    for (int st = 0; st < steps_row; ++st)
      for (int h = 0; h < 2; ++h) {
        int u = 2 * st + h;
        if (2 * st + 1 >= 3 * PT) {       // odd tail: real unit in the second half, zeros in the first
          if (h == 0) continue;
          u = 3 * PT - 1;
        }
        const int j = u / 3, dx = u % 3;
        for (int n = 0; n < cout; ++n)
          for (int k = 0; k < 8; ++k) {
            int ci;   // channel inside the concatenated input
            if (j < P0) { ci = j * 8 + k; if (ci >= cin0) continue; }
            else { ci = (j - P0) * 8 + k; if (ci >= cin1) continue; ci += cin0; }
            const float v = w[((size_t)n * cin + ci) * 9 + dy * 3 + dx];
            out[(size_t)(n / NB) * 3 * steps_row * NB * 16 + (((size_t)(dy * steps_row + st) * 2 + h) * NB + n % NB) * 8 + k] = f2op(v);
          }
      }
  return out;
}

#include "bf16_wide.cuh"

// R3 mode: B operand of step st of an input row = [2 halves][3 NPAD rows][8 k]; row block b holds kernel row dy = 2 - b
// (the input row is the bottom / middle / top row of output rows k-2 / k-1 / k).
std::vector<uint16_t> pack_weights_r3(const float* w, int cout, int cin0, int cin1, int P0, int P1, int NPAD) {
  const int PT = P0 + P1, cin = cin0 + cin1;
  const int steps_row = (3 * PT + 1) / 2, N3 = 3 * NPAD;
  std::vector<uint16_t> out((size_t)steps_row * 2 * N3 * 8, 0);
  for (int st = 0; st < steps_row; ++st)
    for (int h = 0; h < 2; ++h) {
      int u = 2 * st + h;
      if (2 * st + 1 >= 3 * PT) {
        if (h == 0) continue;
        u = 3 * PT - 1;
      }
      const int j = u / 3, dx = u % 3;
      for (int b = 0; b < 3; ++b)
        for (int n = 0; n < cout; ++n)
          for (int k = 0; k < 8; ++k) {
            int ci;
            if (j < P0) { ci = j * 8 + k; if (ci >= cin0) continue; }
            else { ci = (j - P0) * 8 + k; if (ci >= cin1) continue; ci += cin0; }
            const float v = w[((size_t)n * cin + ci) * 9 + (2 - b) * 3 + dx];
            out[(((size_t)st * 2 + h) * N3 + b * NPAD + n) * 8 + k] = f2op(v);
          }
    }
  return out;
}

typedef void (*KernelFn)(const LayerK);

struct Variant {
  int R3;                               // 1: three-output-rows-per-instruction kernel (single CTA), preferred when instantiated
  int CTAS;                             // 2: CTA-pair kernel (cta_group::2), used when a launch has >= 2 frames
  int PT, NPAD, COUT, KIND;             // COUT <= 0: run-time channel count
  int pre0, pre1, post0, post1, skip;   // -1 = run-time op-codes
  KernelFn fn;
  int smem;
};

template <int PT, int NPAD, int COUT, int KIND, class EPI, int CTAS = 1, bool R3 = false>
Variant make_variant(int a, int b, int c, int d, int skip) {
  Variant v{R3 ? 1 : 0, CTAS, PT, NPAD, COUT, KIND, a, b, c, d, skip, conv3x3_tc_kernel<PT, NPAD, COUT, KIND, EPI, CTAS, R3>,
            Cfg<PT, NPAD, CTAS, R3>::SMEM};
  return v;
}
template <int PT, int NPAD, int KIND, bool SKIP>
Variant generic_variant() {
  return make_variant<PT, NPAD, -1, KIND, Epi<-1, -1, -1, -1, SKIP>>(-1, -1, -1, -1, SKIP ? 1 : 0);
}

#define A(x) FSUAE_ACT_##x
const std::vector<Variant>& variants() {
  static const std::vector<Variant> v = {
      // ---- pix_shuffle lightweight, compile-time epilogues (model_pix_shuffle.py:306-311) ----
      make_variant<2, 48, 36, EPI_STORE, Epi<A(SINLU), A(RELU6), 0, 0, false>>(A(SINLU), A(RELU6), 0, 0, 0),
      make_variant<5, 48, 36, EPI_STORE, Epi<A(TELU), 0, A(SINLU), A(BIASED_PRELU), true>>(A(TELU), 0, A(SINLU), A(BIASED_PRELU), 1),
      make_variant<5, 80, 72, EPI_STORE, Epi<0, 0, 0, 0, false>>(0, 0, 0, 0, 0),
      make_variant<9, 80, 72, EPI_STORE, Epi<A(MISH), A(BIASED_PRELU), A(TANH), A(RELU), true>>(A(MISH), A(BIASED_PRELU), A(TANH), A(RELU), 1),
      make_variant<9, 48, 36, EPI_STORE, Epi<0, 0, 0, 0, false>>(0, 0, 0, 0, 0),
      make_variant<10, 48, 36, EPI_STORE, Epi<A(MISH), A(RELU6), 0, 0, false>>(A(MISH), A(RELU6), 0, 0, 0),
      make_variant<5, 16, 12, EPI_TAIL_SHUFFLE, Epi<A(BIASED_PRELU), 0, 0, 0, false>>(A(BIASED_PRELU), 0, 0, 0, 0),
      // ---- the same seven layers as CTA-pair kernels (two frames in lockstep, B operand split across the pair) ----
      make_variant<2, 48, 36, EPI_STORE, Epi<A(SINLU), A(RELU6), 0, 0, false>, 2>(A(SINLU), A(RELU6), 0, 0, 0),
      make_variant<5, 48, 36, EPI_STORE, Epi<A(TELU), 0, A(SINLU), A(BIASED_PRELU), true>, 2>(A(TELU), 0, A(SINLU), A(BIASED_PRELU), 1),
      make_variant<5, 80, 72, EPI_STORE, Epi<0, 0, 0, 0, false>, 2>(0, 0, 0, 0, 0),
      make_variant<9, 80, 72, EPI_STORE, Epi<A(MISH), A(BIASED_PRELU), A(TANH), A(RELU), true>, 2>(A(MISH), A(BIASED_PRELU), A(TANH), A(RELU), 1),
      make_variant<9, 48, 36, EPI_STORE, Epi<0, 0, 0, 0, false>, 2>(0, 0, 0, 0, 0),
      make_variant<10, 48, 36, EPI_STORE, Epi<A(MISH), A(RELU6), 0, 0, false>, 2>(A(MISH), A(RELU6), 0, 0, 0),
      make_variant<5, 16, 12, EPI_TAIL_SHUFFLE, Epi<A(BIASED_PRELU), 0, 0, 0, false>, 2>(A(BIASED_PRELU), 0, 0, 0, 0),
      // ---- pix_shuffle heavyweight = the constructor's default activations (model_pix_shuffle.py:20-68, 312-314), 108-channel middle;
      //      its conv4 (108 -> 108) runs on the wide kernel ----
      make_variant<2, 48, 36, EPI_STORE, Epi<A(RELU), 0, 0, 0, false>>(A(RELU), 0, 0, 0, 0),
      make_variant<5, 48, 36, EPI_STORE, Epi<A(MISH), A(BIASED_RELU), A(TANH), A(RELU6), true>>(A(MISH), A(BIASED_RELU), A(TANH), A(RELU6), 1),
      make_variant<5, 112, 108, EPI_STORE, Epi<0, 0, 0, 0, false>>(0, 0, 0, 0, 0),
      make_variant<14, 48, 36, EPI_STORE, Epi<0, 0, 0, 0, false>>(0, 0, 0, 0, 0),
      make_variant<10, 48, 36, EPI_STORE, Epi<A(MISH), A(PRELU), 0, 0, false>>(A(MISH), A(PRELU), 0, 0, 0),
      make_variant<5, 16, 12, EPI_TAIL_SHUFFLE, Epi<A(SINLU), A(PRELU), 0, 0, false>>(A(SINLU), A(PRELU), 0, 0, 0),
      make_variant<2, 48, 36, EPI_STORE, Epi<A(RELU), 0, 0, 0, false>, 2>(A(RELU), 0, 0, 0, 0),
      make_variant<5, 48, 36, EPI_STORE, Epi<A(MISH), A(BIASED_RELU), A(TANH), A(RELU6), true>, 2>(A(MISH), A(BIASED_RELU), A(TANH), A(RELU6), 1),
      make_variant<5, 112, 108, EPI_STORE, Epi<0, 0, 0, 0, false>, 2>(0, 0, 0, 0, 0),
      make_variant<14, 48, 36, EPI_STORE, Epi<0, 0, 0, 0, false>, 2>(0, 0, 0, 0, 0),
      make_variant<10, 48, 36, EPI_STORE, Epi<A(MISH), A(PRELU), 0, 0, false>, 2>(A(MISH), A(PRELU), 0, 0, 0),
      make_variant<5, 16, 12, EPI_TAIL_SHUFFLE, Epi<A(SINLU), A(PRELU), 0, 0, false>, 2>(A(SINLU), A(PRELU), 0, 0, 0),
      // ---- the layers without a residual as three-output-rows-per-instruction kernels (A operand read once per input row) ----
      make_variant<2, 48, 36, EPI_STORE, Epi<A(SINLU), A(RELU6), 0, 0, false>, 1, true>(A(SINLU), A(RELU6), 0, 0, 0),
      make_variant<9, 48, 36, EPI_STORE, Epi<0, 0, 0, 0, false>, 1, true>(0, 0, 0, 0, 0),
      make_variant<10, 48, 36, EPI_STORE, Epi<A(MISH), A(RELU6), 0, 0, false>, 1, true>(A(MISH), A(RELU6), 0, 0, 0),
      make_variant<5, 16, 12, EPI_TAIL_SHUFFLE, Epi<A(BIASED_PRELU), 0, 0, 0, false>, 1, true>(A(BIASED_PRELU), 0, 0, 0, 0),
      // ---- run-time epilogues: any activation chain of the registry (except channel softmax), any Cout <= NPAD ----
      // pix_shuffle channel plans (36/72 and the heavyweight 108)
      generic_variant<2, 48, EPI_STORE, false>(), generic_variant<5, 48, EPI_STORE, false>(), generic_variant<5, 48, EPI_STORE, true>(),
      generic_variant<5, 80, EPI_STORE, false>(), generic_variant<9, 80, EPI_STORE, true>(), generic_variant<9, 80, EPI_STORE, false>(),
      generic_variant<9, 48, EPI_STORE, false>(), generic_variant<10, 48, EPI_STORE, false>(),
      generic_variant<5, 112, EPI_STORE, false>(), generic_variant<14, 48, EPI_STORE, true>(), generic_variant<14, 48, EPI_STORE, false>(),
      generic_variant<5, 16, EPI_TAIL_SHUFFLE, false>(),
      // conv3 / conv5 lightweight (32/64) and conv5 heavyweight (64/128)
      generic_variant<1, 32, EPI_STORE, false>(), generic_variant<4, 64, EPI_STORE, false>(), generic_variant<4, 32, EPI_STORE, true>(),
      generic_variant<8, 64, EPI_STORE, true>(), generic_variant<1, 64, EPI_STORE, false>(), generic_variant<8, 128, EPI_STORE, false>(),
      generic_variant<16, 32, EPI_STORE, true>(),
      generic_variant<8, 16, EPI_TAIL_PLAIN, false>(), generic_variant<16, 16, EPI_TAIL_PLAIN, false>(),
      // the same BatchNorm-folded families with compile-time op-codes (ReLU; residual + ReLU; identity / sigmoid tails):
      // the run-time chain costs ~4x the instructions per chunk and these layers are issue-bound at full resolution
      make_variant<1, 32, -1, EPI_STORE, Epi<A(RELU), 0, 0, 0, false>>(A(RELU), 0, 0, 0, 0),
      make_variant<1, 64, -1, EPI_STORE, Epi<A(RELU), 0, 0, 0, false>>(A(RELU), 0, 0, 0, 0),
      make_variant<4, 64, -1, EPI_STORE, Epi<A(RELU), 0, 0, 0, false>>(A(RELU), 0, 0, 0, 0),
      make_variant<8, 128, -1, EPI_STORE, Epi<A(RELU), 0, 0, 0, false>>(A(RELU), 0, 0, 0, 0),
      make_variant<4, 32, -1, EPI_STORE, Epi<0, 0, A(RELU), 0, true>>(0, 0, A(RELU), 0, 1),
      make_variant<8, 64, -1, EPI_STORE, Epi<0, 0, A(RELU), 0, true>>(0, 0, A(RELU), 0, 1),
      make_variant<8, 16, -1, EPI_TAIL_PLAIN, Epi<0, 0, 0, 0, false>>(0, 0, 0, 0, 0),
      make_variant<8, 16, -1, EPI_TAIL_PLAIN, Epi<A(SIGMOID), 0, 0, 0, false>>(A(SIGMOID), 0, 0, 0, 0),
      make_variant<16, 16, -1, EPI_TAIL_PLAIN, Epi<A(SIGMOID), 0, 0, 0, false>>(A(SIGMOID), 0, 0, 0, 0),
      // ... the residual layers as CTA pairs (measured: 8-13 % faster; the plain ReLU layers are bound by their full-resolution
      //     stores, ~4.4 TB/s of HBM traffic, and lose 5-10 % as pairs, so they stay single-CTA) ...
      make_variant<4, 32, -1, EPI_STORE, Epi<0, 0, A(RELU), 0, true>, 2>(0, 0, A(RELU), 0, 1),
      make_variant<8, 64, -1, EPI_STORE, Epi<0, 0, A(RELU), 0, true>, 2>(0, 0, A(RELU), 0, 1),
      // ... and the 3-channel full-resolution tails with three output rows per instruction: at N = 16 the instruction is
      // bound by its 4 KB A read, so tripling N (48) is free and the tail needs a third of the instructions
      make_variant<8, 16, -1, EPI_TAIL_PLAIN, Epi<0, 0, 0, 0, false>, 1, true>(0, 0, 0, 0, 0),
      make_variant<8, 16, -1, EPI_TAIL_PLAIN, Epi<A(SIGMOID), 0, 0, 0, false>, 1, true>(A(SIGMOID), 0, 0, 0, 0),
      make_variant<16, 16, -1, EPI_TAIL_PLAIN, Epi<A(SIGMOID), 0, 0, 0, false>, 1, true>(A(SIGMOID), 0, 0, 0, 0),
  };
  return v;
}
#undef A

struct Launch {
  const Variant* var = nullptr;
  unsigned char* d_w = nullptr;
  float* d_params = nullptr;       // device copy of the per-channel epilogue parameters
  const Variant* var2 = nullptr;   // CTA-pair kernel for the same layer (weights packed per CTA), if instantiated
  unsigned char* d_w2 = nullptr;
  const Variant* var3 = nullptr;   // three-output-rows-per-instruction kernel for the same layer, if instantiated
  unsigned char* d_w3 = nullptr;
  LayerK k{};   // geometry-independent fields prefilled
  const WideVariant* wide = nullptr;   // K-streamed tile kernel instead (layers too wide for resident weights / full-depth rows)
  WideK wk{};
};

struct LayerPlan {
  std::vector<Launch> launches;   // a wide layer is split over output-channel groups (A is re-read per group)
};

}  // namespace

struct TcPlan {
  std::vector<LayerPlan> layers;
  std::vector<unsigned char*> buf;     // chunk-planar activation buffers, id 0..n_layers-1 (the last layer writes the frame)
  std::vector<int> planes;
  std::vector<size_t> buf_bytes;
  int zero_Hw = -1, zero_Ww = -1;      // geometry the borders are currently valid for
  // fused layer pair (0-based index of layer A, -1: none): A's output never leaves the SM
  int fused_at = -1;
  void (*fused_fn)(const LayerK, const LayerK) = nullptr;
  int fused_smem = 0;
  // the single fused pass (mega.cuh): the whole flagship network as one persistent kernel, feature maps in L2-resident rings
  bool mega_ok = false;
  int mega_groups = 0;                 // groups of 4 stage pairs the GPU holds
  unsigned char* mega_scratch = nullptr;
  size_t mega_scratch_bytes = 0;
  unsigned int* mega_flags = nullptr;
  size_t mega_flag_words = 0;
  unsigned char* mega_zero = nullptr;
  int mega_zero_Ww = -1;               // geometry the scratch rows' zero borders are valid for
  MegaK mega_k{};                      // per-layer parameters prefilled
};

// bytes of one (team, rank) channel block for rows of PW slots
static size_t mega_block_bytes(int PW, unsigned long long* ch_off = nullptr) {
  const int planes[MG_NCH] = {5, 5, 9, 9, 5, 5, 2};
  size_t off = 0;
  for (int c = 0; c < MG_NCH; ++c) {
    if (ch_off) ch_off[c] = off;
    off += (size_t)planes[c] * mg_depth(c) * PW * 16;
  }
  return (off + 255) / 256 * 256;
}

static int planes_of(int c) { return (c + 7) / 8; }

// debugging aid (not part of the public header): raw chunk-planar bytes of activation buffer `id`
#ifndef FSUAE_OPERAND_FP16
extern "C" __attribute__((visibility("default"))) long long fsuae_debug_read_bf16_buffer(fsuae_engine* e, int id, void* dst,
                                                                                        long long max_bytes) {
  if (!e || !e->bf16 || id < 0 || id >= (int)e->bf16->buf.size()) return -1;
  cudaDeviceSynchronize();
  long long n = std::min<long long>(max_bytes, (long long)e->bf16->buf_bytes[id]);
  if (cudaMemcpy(dst, e->bf16->buf[id], n, cudaMemcpyDeviceToHost) != cudaSuccess) return -2;
  return n;
}
#endif

#if defined(FSUAE_EPI_TIMING) && !defined(FSUAE_OPERAND_FP16)
extern "C" __attribute__((visibility("default"))) int fsuae_debug_epi_timing(unsigned long long* out8, int reset) {   // 16 counters
  cudaDeviceSynchronize();
  if (cudaMemcpyFromSymbol(out8, g_epi_timing, sizeof(g_epi_timing)) != cudaSuccess) return -1;
  if (reset) { unsigned long long z[32] = {0}; cudaMemcpyToSymbol(g_epi_timing, z, sizeof(z)); }
  return 0;
}
#endif

int TC_FN(create)(fsuae_engine* e) {
  const fsuae_net_desc& d = e->desc;
  const bool unshuffle = d.head == FSUAE_HEAD_UNSHUFFLE2;
  TcPlan* plan = new TcPlan();
  e->TC_FIELD = plan;
  const bool features = d.head == FSUAE_HEAD_FEATURES;      // float feature maps in and out (residual-UNet building blocks)
  std::vector<int> ch(d.n_layers + 1);
  ch[0] = unshuffle ? 12 : (features ? d.in_channels : 3);
  for (int i = 0; i < d.n_layers; ++i) ch[i + 1] = d.layers[i].cout;

  plan->layers.resize(d.n_layers);
  for (int i = 0; i < d.n_layers; ++i) {
    const fsuae_layer_desc& L = d.layers[i];
    LayerPlan& lp = plan->layers[i];
    const bool last = i == d.n_layers - 1;
    const std::string tag = std::string(TC_BUILD_TAG) + "layer " + std::to_string(i + 1) + ": ";
    const int P0 = planes_of(L.cin0), P1 = L.cin1 > 0 ? planes_of(L.cin1) : 0, PT = P0 + P1;
    int ops[4] = {0, 0, 0, 0};
    if (L.n_pre > 2 || L.n_post > 2) return set_error(e, FSUAE_ERR_UNSUPPORTED, tag + "at most 2 activation slots before and after the skip add");
    for (int k = 0; k < L.n_pre; ++k) ops[k] = L.pre[k].op;
    for (int k = 0; k < L.n_post; ++k) ops[2 + k] = L.post[k].op;
    // a channel softmax / log_softmax slot: handled by the run-time store epilogue (three passes over the accumulator row)
    int softmax_slot = -1, softmax_log = 0;
    for (int k = 0; k < 4; ++k)
      if (act_is_softmax(ops[k])) {
        if (softmax_slot >= 0 || last)
          return set_error(e, FSUAE_ERR_UNSUPPORTED, tag + "more than one channel softmax slot, or one in the last layer, is not implemented on the bf16 build (use the fp32 build)");
        softmax_slot = k;
        softmax_log = ops[k] == FSUAE_ACT_LOG_SOFTMAX;
        ops[k] = FSUAE_ACT_IDENTITY;
      }
    const int kind = (!last || features) ? EPI_STORE : (d.tail == FSUAE_TAIL_SHUFFLE2_RESIDUAL_RELU ? EPI_TAIL_SHUFFLE : EPI_TAIL_PLAIN);
    // The residual is normally the layer's own input and is then read from the centre row of the shared-memory ring;
    // a residual from another buffer (the 1x1 skip projections of model_pix_shuffle.py:126-128, 143-145) is read from
    // global memory by the run-time epilogues.
    const int any_skip = L.skip_src >= 0 ? 1 : 0;
    const bool global_skip = any_skip && (L.skip_src != L.src0 || L.cin1 > 0);
    const int skip = any_skip && !global_skip ? 1 : 0;

    // pick the kernel: exact compile-time epilogue if there is one, else the run-time one; a layer wider than the
    // widest instantiated N for its input plane count is split into output-channel groups
    const int need = (L.cout + 15) / 16 * 16;
    const Variant* exact = nullptr;
    const Variant* fit = nullptr;        // smallest run-time-op NPAD >= need
    const Variant* widest = nullptr;     // widest run-time-op NPAD
    const Variant* fit_ops = nullptr;    // the same with this layer's op-codes compiled in (run-time channel count)
    const Variant* widest_ops = nullptr;
    for (const Variant& v : variants()) {
      if (v.CTAS != 1 || v.R3 || v.PT != PT || v.KIND != kind || v.skip != skip) continue;
      const bool ops_eq = v.pre0 == ops[0] && v.pre1 == ops[1] && v.post0 == ops[2] && v.post1 == ops[3];
      if (v.COUT == L.cout && ops_eq) exact = &v;
      if (v.COUT > 0) continue;
      if (ops_eq) {
        if (v.NPAD >= need && (!fit_ops || v.NPAD < fit_ops->NPAD)) fit_ops = &v;
        if (!widest_ops || v.NPAD > widest_ops->NPAD) widest_ops = &v;
      } else if (v.pre0 == -1) {
        if (v.NPAD >= need && (!fit || v.NPAD < fit->NPAD)) fit = &v;
        if (!widest || v.NPAD > widest->NPAD) widest = &v;
      }
    }
    if (global_skip || softmax_slot >= 0) exact = fit_ops = widest_ops = nullptr;   // only the run-time epilogue knows about a residual in global memory / a softmax slot
    const bool big_kernel = L.ksize > 3;       // 5x5 / 7x7: window decomposition on the K-streamed wide kernel
    if (big_kernel) exact = fit_ops = widest_ops = fit = widest = nullptr;
    if (softmax_slot >= 0 && fit == nullptr && widest != nullptr) widest = nullptr;      // the softmax needs every channel in one launch: wide kernel, one group
    const Variant* var = exact ? exact : fit_ops ? fit_ops : fit ? fit : widest_ops ? widest_ops : widest;
    // Layers whose weights / full-depth input rows do not fit beside each other in shared memory stream K through the
    // wide tile kernel (thin-input layers are cheap to split over output-channel groups instead).
    const bool too_wide = !var || var->NPAD < need;
    const bool use_wide = big_kernel || (!exact && !e->tuning.no_wide && ((too_wide && (PT >= 4 || !var)) || e->tuning.force_wide));
    if (use_wide) {
      const int ngroups = kind == EPI_STORE ? (need + 127) / 128 : 1;
      if (softmax_slot >= 0 && ngroups > 1)
        return set_error(e, FSUAE_ERR_UNSUPPORTED, tag + "channel softmax over more than 128 channels is not implemented on the bf16 build");
      const int per = ((need + ngroups - 1) / ngroups + 15) / 16 * 16;
      const WideVariant* wv = nullptr;
      for (int pass = softmax_slot >= 0 ? 1 : 0; pass < 2 && !wv; ++pass)      // first a variant with this layer's op-codes compiled in, then the run-time one
        for (const WideVariant& v : wide_variants()) {
          const bool ops_eq = v.pre0 == ops[0] && v.pre1 == ops[1] && v.post0 == ops[2] && v.post1 == ops[3] && v.skip == any_skip;
          if (v.KIND != kind || v.NT < per || (pass == 0 ? !ops_eq : v.pre0 != -1)) continue;
          if (!wv || v.NT < wv->NT) wv = &v;
        }
      if (!wv) return set_error(e, FSUAE_ERR_UNSUPPORTED, tag + "no wide tensor-core kernel for " + std::to_string(L.cout) + " output channels");
      Launch ln;
      ln.wide = wv;
      const KWindows kw = k_windows(L.ksize);
      const int nwin = kw.n1 * kw.n1;
      std::vector<uint16_t> wp;
      if (nwin == 1) {
        wp = pack_weights_wide(e->h_blob.data() + L.w_off, L.cout, L.cin0, L.cin1, P0, P1, wv->NT, ngroups);
      } else {
        const std::vector<float> wexp = expand_weights_windows(e->h_blob.data() + L.w_off, L.cout, L.cin0, L.cin1, P0, P1, L.ksize);
        wp = pack_weights_wide(wexp.data(), L.cout, nwin * PT * 8, 0, nwin * PT, 0, wv->NT, ngroups);
      }
      FSUAE_CUDA_CHECK(e, cudaMalloc(&ln.d_w, wp.size() * 2));
      FSUAE_CUDA_CHECK(e, cudaMemcpy(ln.d_w, wp.data(), wp.size() * 2, cudaMemcpyHostToDevice));
      e->device_bytes += wp.size() * 2;
      WideK& k = ln.wk;
      std::memset(&k, 0, sizeof(k));
      k.P0 = P0; k.pin_real = PT; k.nwin = nwin; k.pin = nwin * PT; k.kchunks = (nwin * PT + 1) / 2;
      for (int wy = 0; wy < kw.n1; ++wy)
        for (int wx = 0; wx < kw.n1; ++wx) { k.win_dy[wy * kw.n1 + wx] = (signed char)kw.centre[wy]; k.win_dx[wy * kw.n1 + wx] = (signed char)kw.centre[wx]; }
      k.ngroups = ngroups; k.cout = L.cout; k.cpad = ngroups * wv->NT;
      k.has_skip = any_skip;
      k.softmax_slot = softmax_slot; k.softmax_log = softmax_log;
      k.wpack = ln.d_w;
      std::vector<float> hp((size_t)9 * k.cpad, 0.f);
      for (int c = 0; c < k.cpad; ++c) {
        const int cc = std::min(c, L.cout - 1);
        hp[c] = (c < L.cout && L.b_off >= 0) ? e->h_blob[L.b_off + c] : 0.f;
        for (int s = 0; s < 4; ++s) {
          const fsuae_act_desc* a = nullptr;
          if (s < 2 && s < L.n_pre) a = &L.pre[s];
          if (s >= 2 && s - 2 < L.n_post) a = &L.post[s - 2];
          float v0 = (a && a->n0 > 0) ? e->h_blob[a->p0_off + (a->n0 == 1 ? 0 : cc)] : 0.f;
          float v1 = (a && a->n1 > 0) ? e->h_blob[a->p1_off + (a->n1 == 1 ? 0 : cc)] : 0.f;
          if (a && a->op == FSUAE_ACT_BIASED_PRELU) v1 -= 1.f;
          hp[(size_t)(1 + s) * k.cpad + c] = v0;
          hp[(size_t)(5 + s) * k.cpad + c] = v1;
        }
      }
      for (int s = 0; s < 4; ++s) k.op[s] = ops[s];
      FSUAE_CUDA_CHECK(e, cudaMalloc(&ln.d_params, hp.size() * sizeof(float)));
      FSUAE_CUDA_CHECK(e, cudaMemcpy(ln.d_params, hp.data(), hp.size() * sizeof(float), cudaMemcpyHostToDevice));
      k.dparams = ln.d_params;
      FSUAE_CUDA_CHECK(e, cudaFuncSetAttribute((const void*)wv->fn, cudaFuncAttributeMaxDynamicSharedMemorySize, wv->smem));
      lp.launches.push_back(ln);
      continue;
    }
    if (!var || (kind != EPI_STORE && var->NPAD < need))
      return set_error(e, FSUAE_ERR_UNSUPPORTED,
                       tag + "no tensor-core kernel instantiated for " + std::to_string(PT) + " input planes / " +
                           std::to_string(L.cout) + " output channels; use the fp32 build");
    for (int c0 = 0; c0 < L.cout; c0 += var->NPAD) {
      const int cg = std::min(var->NPAD, L.cout - c0);
      Launch ln;
      ln.var = var;
      std::vector<uint16_t> wp = pack_weights(e->h_blob.data() + L.w_off + (size_t)c0 * (L.cin0 + L.cin1) * 9, cg, L.cin0,
                                              L.cin1, P0, P1, var->NPAD);
      FSUAE_CUDA_CHECK(e, cudaMalloc(&ln.d_w, wp.size() * 2));
      FSUAE_CUDA_CHECK(e, cudaMemcpy(ln.d_w, wp.data(), wp.size() * 2, cudaMemcpyHostToDevice));
      e->device_bytes += wp.size() * 2;
      LayerK& k = ln.k;
      std::memset(&k, 0, sizeof(k));
      k.P0 = P0; k.P1 = P1; k.cout = cg;
      k.out_planes = planes_of(cg);
      k.dst_plane0 = c0 / 8;
      k.skip_plane0 = c0 / 8;
      k.tail = d.tail;
      k.wpack = ln.d_w;
      k.n_pre = L.n_pre; k.n_post = L.n_post;
      k.softmax_slot = softmax_slot; k.softmax_log = softmax_log;
      for (int c = 0; c < MAXC; ++c) k.bias[c] = (c < cg && L.b_off >= 0) ? e->h_blob[L.b_off + c0 + c] : 0.f;
      for (int s = 0; s < 4; ++s) {
        k.op[s] = ops[s];
        const fsuae_act_desc* a = nullptr;
        if (s < 2 && s < L.n_pre) a = &L.pre[s];
        if (s >= 2 && s - 2 < L.n_post) a = &L.post[s - 2];
        for (int c = 0; c < MAXC; ++c) {
          const int cc = std::min(c0 + c, L.cout - 1);
          k.p0[s][c] = (a && a->n0 > 0) ? e->h_blob[a->p0_off + (a->n0 == 1 ? 0 : cc)] : 0.f;
          k.p1[s][c] = (a && a->n1 > 0) ? e->h_blob[a->p1_off + (a->n1 == 1 ? 0 : cc)] : 0.f;
          if (a && a->op == FSUAE_ACT_BIASED_PRELU) k.p1[s][c] -= 1.f;   // the kernel evaluates y + (slope - 1) * min(y, 0)
        }
      }
      {
        std::vector<float> hp((size_t)9 * MAXC);
        for (int c = 0; c < MAXC; ++c) {
          hp[c] = k.bias[c];
          for (int sl = 0; sl < 4; ++sl) { hp[(1 + sl) * MAXC + c] = k.p0[sl][c]; hp[(5 + sl) * MAXC + c] = k.p1[sl][c]; }
        }
        FSUAE_CUDA_CHECK(e, cudaMalloc(&ln.d_params, hp.size() * sizeof(float)));
        FSUAE_CUDA_CHECK(e, cudaMemcpy(ln.d_params, hp.data(), hp.size() * sizeof(float), cudaMemcpyHostToDevice));
        k.dparams = ln.d_params;
      }
      // Opt-in (FSUAE_R3=1): measured on B200 it only ties the CTA-pair kernels (conv5 6.4 vs 6.0 us/frame, conv7 3.6 vs
      // 3.6) -- the N = 144 instruction costs 85 cycles against 3 x 43 for the pair kernel's three N = 48 ones, but the
      // accumulator-ring wrap splits, the separate first instruction of every new row and the longer per-row scalar
      // prologue of the single issuing thread eat the difference.
      const bool matched = var == exact || var == fit_ops;      // op-codes compiled in: sibling kernels exist for the same signature
      const bool r3_default = var->KIND == EPI_TAIL_PLAIN;      // N = 16 tails: clear win (conv3 lightweight tail 21.7 -> see profiles)
      if (matched && (e->tuning.r3 >= 0 ? e->tuning.r3 != 0 : r3_default)) {
        for (const Variant& v : variants())
          if (v.R3 && v.PT == var->PT && v.NPAD == var->NPAD && v.COUT == var->COUT && v.KIND == var->KIND && v.skip == var->skip &&
              v.pre0 == var->pre0 && v.pre1 == var->pre1 && v.post0 == var->post0 && v.post1 == var->post1)
            ln.var3 = &v;
        if (ln.var3) {
          std::vector<uint16_t> wp3 = pack_weights_r3(e->h_blob.data() + L.w_off, cg, L.cin0, L.cin1, P0, P1, var->NPAD);
          FSUAE_CUDA_CHECK(e, cudaMalloc(&ln.d_w3, wp3.size() * 2));
          FSUAE_CUDA_CHECK(e, cudaMemcpy(ln.d_w3, wp3.data(), wp3.size() * 2, cudaMemcpyHostToDevice));
          e->device_bytes += wp3.size() * 2;
          FSUAE_CUDA_CHECK(e, cudaFuncSetAttribute((const void*)ln.var3->fn, cudaFuncAttributeMaxDynamicSharedMemorySize, ln.var3->smem));
        }
      }
      if (matched && !e->tuning.no_pairs) {
        for (const Variant& v : variants())
          if (v.CTAS == 2 && !v.R3 && v.PT == var->PT && v.NPAD == var->NPAD && v.COUT == var->COUT && v.KIND == var->KIND &&
              v.skip == var->skip && v.pre0 == var->pre0 && v.pre1 == var->pre1 && v.post0 == var->post0 && v.post1 == var->post1)
            ln.var2 = &v;
        if (ln.var2) {
          std::vector<uint16_t> wp2 = pack_weights(e->h_blob.data() + L.w_off, cg, L.cin0, L.cin1, P0, P1, var->NPAD, 2);
          FSUAE_CUDA_CHECK(e, cudaMalloc(&ln.d_w2, wp2.size() * 2));
          FSUAE_CUDA_CHECK(e, cudaMemcpy(ln.d_w2, wp2.data(), wp2.size() * 2, cudaMemcpyHostToDevice));
          e->device_bytes += wp2.size() * 2;
          FSUAE_CUDA_CHECK(e, cudaFuncSetAttribute((const void*)ln.var2->fn, cudaFuncAttributeMaxDynamicSharedMemorySize, ln.var2->smem));
        }
      }
      lp.launches.push_back(ln);
    }
    FSUAE_CUDA_CHECK(e, cudaFuncSetAttribute((const void*)var->fn, cudaFuncAttributeMaxDynamicSharedMemorySize, var->smem));
  }
  // conv3 -> conv4 of the flagship preset can run as one fused kernel on CTA pairs
  if (!e->tuning.no_fusion) {
    using FusedEpiA = Epi<0, 0, 0, 0, false>;
    using FusedEpiB = Epi<FSUAE_ACT_MISH, FSUAE_ACT_BIASED_PRELU, FSUAE_ACT_TANH, FSUAE_ACT_RELU, true>;
    for (int i = 0; i + 1 < d.n_layers - 1; ++i) {
      const fsuae_layer_desc& LA = d.layers[i];
      const fsuae_layer_desc& LB = d.layers[i + 1];
      const LayerPlan& pa = plan->layers[i];
      const LayerPlan& pb = plan->layers[i + 1];
      if (pa.launches.size() != 1 || pb.launches.size() != 1 || !pa.launches[0].var2 || !pb.launches[0].var2) continue;
      const Variant* va = pa.launches[0].var;
      const Variant* vb = pb.launches[0].var;
      const bool a_ok = va->PT == 5 && va->NPAD == 80 && va->COUT == 72 && va->skip == 0 && va->pre0 == 0 && va->pre1 == 0 &&
                        va->post0 == 0 && va->post1 == 0 && LA.cin1 == 0;
      const bool b_ok = vb->PT == 9 && vb->NPAD == 80 && vb->COUT == 72 && vb->skip == 1 && vb->pre0 == FSUAE_ACT_MISH &&
                        vb->pre1 == FSUAE_ACT_BIASED_PRELU && vb->post0 == FSUAE_ACT_TANH && vb->post1 == FSUAE_ACT_RELU &&
                        LB.src0 == i + 1 && LB.cin1 == 0 && LB.skip_src == i + 1;
      bool a_private = true;     // nobody else reads A's output
      for (int j = 0; j < d.n_layers; ++j)
        if (j != i + 1 && (d.layers[j].src0 == i + 1 || (d.layers[j].cin1 > 0 && d.layers[j].src1 == i + 1) || d.layers[j].skip_src == i + 1))
          a_private = false;
      if (a_ok && b_ok && a_private) {
        plan->fused_at = i;
        plan->fused_fn = conv3x3_tc_fused_pair_kernel<5, 80, 72, FusedEpiA, 80, 72, FusedEpiB>;
        plan->fused_smem = FusedCfg<5, 80, 72, 80, 72>::SMEM;
        FSUAE_CUDA_CHECK(e, cudaFuncSetAttribute((const void*)plan->fused_fn, cudaFuncAttributeMaxDynamicSharedMemorySize, plan->fused_smem));
        break;
      }
    }
  }
  // The whole flagship network as ONE persistent kernel (mega.cuh).  Needs exactly the lightweight preset's seven layers
  // (its engines are compile-time specialised), wired conv1 -> ... -> conv7 with the residuals on conv2 / conv4 and the
  // long skip into conv6 (model_pix_shuffle.py:227-298).
  if (!e->tuning.no_mega && d.n_layers == 7 && unshuffle && d.tail == FSUAE_TAIL_SHUFFLE2_RESIDUAL_RELU) {
    struct Sig { int PT, NPAD, COUT, pre0, pre1, post0, post1, skip, src0, src1, skip_src; };
    const Sig want[7] = {
        {2, 48, 36, FSUAE_ACT_SINLU, FSUAE_ACT_RELU6, 0, 0, 0, 0, -1, -1},
        {5, 48, 36, FSUAE_ACT_TELU, 0, FSUAE_ACT_SINLU, FSUAE_ACT_BIASED_PRELU, 1, 1, -1, 1},
        {5, 80, 72, 0, 0, 0, 0, 0, 2, -1, -1},
        {9, 80, 72, FSUAE_ACT_MISH, FSUAE_ACT_BIASED_PRELU, FSUAE_ACT_TANH, FSUAE_ACT_RELU, 1, 3, -1, 3},
        {9, 48, 36, 0, 0, 0, 0, 0, 4, -1, -1},
        {10, 48, 36, FSUAE_ACT_MISH, FSUAE_ACT_RELU6, 0, 0, 0, 1, 5, -1},
        {5, 16, 12, FSUAE_ACT_BIASED_PRELU, 0, 0, 0, 0, 6, -1, -1}};
    bool ok = true;
    for (int i = 0; i < 7 && ok; ++i) {
      const LayerPlan& lp = plan->layers[i];
      const fsuae_layer_desc& L = d.layers[i];
      ok = lp.launches.size() == 1 && lp.launches[0].var2 != nullptr;
      if (!ok) break;
      const Variant* v = lp.launches[0].var2;
      const Sig& w = want[i];
      ok = v->PT == w.PT && v->NPAD == w.NPAD && v->COUT == w.COUT && v->pre0 == w.pre0 && v->pre1 == w.pre1 && v->post0 == w.post0 &&
           v->post1 == w.post1 && v->skip == w.skip && L.src0 == w.src0 && (w.src1 < 0 ? L.cin1 == 0 : (L.cin1 > 0 && L.src1 == w.src1)) &&
           L.skip_src == w.skip_src;
    }
    plan->mega_groups = e->sm_count / 8;
    const int Hw0 = e->H / 2;
    (void)Hw0;
    size_t need = 0, flag_words = 0;
    for (int crop = 0; crop < 2 && ok; ++crop) {            // both geometries an enqueue may ask for (FSUAE_FLAG_CROP16)
      const int We = e->W - (crop ? 16 : 0);
      if (We < 2) continue;
      const int Ww = We / 2, S = (Ww + STRIP - 1) / STRIP, teams = plan->mega_groups / std::max(S, 1);
      if (S > MG_SMAX || teams < 1) { if (!crop) ok = false; continue; }
      need = std::max(need, (size_t)teams * 2 * mega_block_bytes(plane_width(S)));
      flag_words = std::max(flag_words, (size_t)teams * 2 * MG_FLAG_WORDS);
    }
    if (ok) {
      FSUAE_CUDA_CHECK(e, cudaMalloc(&plan->mega_scratch, need));
      FSUAE_CUDA_CHECK(e, cudaMalloc(&plan->mega_flags, flag_words * sizeof(unsigned int)));
      FSUAE_CUDA_CHECK(e, cudaMalloc(&plan->mega_zero, PLANE_ROW + 256));
      FSUAE_CUDA_CHECK(e, cudaMemset(plan->mega_zero, 0, PLANE_ROW + 256));
      plan->mega_scratch_bytes = need;
      plan->mega_flag_words = flag_words;
      e->device_bytes += need + flag_words * sizeof(unsigned int) + PLANE_ROW + 256;
      for (int i = 0; i < 7; ++i) {
        const Launch& ln = plan->layers[i].launches[0];
        MegaLayerP& lp = plan->mega_k.L[i];
        for (int c = 0; c < MG_MAXC; ++c) {
          lp.bias[c] = ln.k.bias[c];
          for (int sl = 0; sl < 4; ++sl) { lp.p0[sl][c] = ln.k.p0[sl][c]; lp.p1[sl][c] = ln.k.p1[sl][c]; }
        }
        lp.wpack = ln.d_w2;
      }
      FSUAE_CUDA_CHECK(e, (cudaError_t)TC_FN(mega_prepare)());
      plan->mega_ok = true;
    }
  }
  // activation buffers at the largest geometry (no crop)
  const int Hw = unshuffle ? e->H / 2 : e->H, Ww = unshuffle ? e->W / 2 : e->W;
  const int S = (Ww + STRIP - 1) / STRIP, PW = plane_width(S);
  const int n_bufs = d.n_layers + (features ? 1 : 0);       // a feature-map network also keeps its last layer's output chunk-planar
  plan->buf.assign(n_bufs, nullptr);
  plan->planes.assign(n_bufs, 0);
  plan->buf_bytes.assign(n_bufs, 0);
  for (int i = 0; i < n_bufs; ++i) {
    plan->planes[i] = planes_of(ch[i]);
    size_t bytes = (size_t)e->chunk * plan->planes[i] * (Hw + 2 * BORDER) * PW * 16 + 256;
    FSUAE_CUDA_CHECK(e, cudaMalloc(&plan->buf[i], bytes));
    plan->buf_bytes[i] = bytes;
    e->device_bytes += bytes;
  }
  e->variant = TC_VARIANT_NAME;
  return FSUAE_OK;
}

void TC_FN(destroy)(fsuae_engine* e) {
  if (!e->TC_FIELD) return;
  for (auto& lp : e->TC_FIELD->layers)
    for (auto& ln : lp.launches)
      { if (ln.d_w) cudaFree(ln.d_w); if (ln.d_w2) cudaFree(ln.d_w2); if (ln.d_w3) cudaFree(ln.d_w3); if (ln.d_params) cudaFree(ln.d_params); }
  for (unsigned char* p : e->TC_FIELD->buf)
    if (p) cudaFree(p);
  if (e->TC_FIELD->mega_scratch) cudaFree(e->TC_FIELD->mega_scratch);
  if (e->TC_FIELD->mega_flags) cudaFree(e->TC_FIELD->mega_flags);
  if (e->TC_FIELD->mega_zero) cudaFree(e->TC_FIELD->mega_zero);
  delete e->TC_FIELD;
  e->TC_FIELD = nullptr;
}

int TC_FN(enqueue_chunk)(fsuae_engine* e, const void* in, void* out, int n, int in_fmt, int out_fmt, uint32_t flags,
                          cudaStream_t st) {
  TcPlan* plan = e->TC_FIELD;
  const fsuae_net_desc& d = e->desc;
  const Geom g = make_geom(e, flags);
  const int S = (g.Ww + STRIP - 1) / STRIP, PW = plane_width(S);
  // ---- the single fused pass: one persistent kernel, feature maps in L2-resident rings (mega.cuh) ----
  {
    const int teams = plan->mega_ok ? plan->mega_groups / S : 0;
    // Default: framebuffer (uint8 RGBA) passes of 8 frames or more.  Float / planar frames run correctly through the fused pass
    // (FSUAE_MEGA_MIN_FRAMES=n sends them there) but its head reads them through registers, one row ahead, instead of the
    // framebuffer path's cp.async staging two rows ahead, and the layer kernels are faster for those formats at every batch
    // size (float in -> float out, 64 frames: 42.9 vs 45.5 us/frame).
    const int min_frames = e->tuning.mega_min_frames >= 0 ? e->tuning.mega_min_frames : (in_fmt == FSUAE_FMT_U8_NHWC4 ? 8 : INT_MAX);
    if (teams >= 1 && S <= MG_SMAX && n >= std::max(2, min_frames) &&
        (size_t)teams * 2 * mega_block_bytes(PW) <= plan->mega_scratch_bytes) {
      if (plan->mega_zero_Ww != g.Ww) {      // zero borders / unused tail slots of the ring rows for this geometry
        FSUAE_CUDA_CHECK(e, cudaMemsetAsync(plan->mega_scratch, 0, plan->mega_scratch_bytes, st));
        plan->mega_zero_Ww = g.Ww;
      }
      FSUAE_CUDA_CHECK(e, cudaMemsetAsync(plan->mega_flags, 0, (size_t)teams * 2 * MG_FLAG_WORDS * sizeof(unsigned int), st));
      MegaK& k = plan->mega_k;
      k.Hw = g.Hw; k.Ww = g.Ww; k.PW = PW; k.S = S; k.n_frames = n; k.n_fp = (n + 1) / 2; k.teams = teams;
      k.H = g.H; k.W = g.W; k.xoff = g.xoff; k.in_fmt = in_fmt; k.out_fmt = out_fmt;
      k.gamma_in = (flags & FSUAE_FLAG_GAMMA_IN) ? 1 : 0;
      k.gamma_out = (flags & FSUAE_FLAG_GAMMA_OUT) ? 1 : 0;
      k.frame_in = in; k.frame_out = out;
      k.scratch = plan->mega_scratch;
      k.rank_stride = mega_block_bytes(PW, k.ch_off);
      k.flags = plan->mega_flags;
      k.zero_row = plan->mega_zero;
      { ProfScope ps(e, st, "fused_pass");
        FSUAE_CUDA_CHECK(e, (cudaError_t)TC_FN(mega_launch)(k, 8 * teams * S, st)); }
      e->launches++;
      if (g.xoff > 0) {
        black_columns_bf16_kernel<<<64, 256, 0, st>>>(out, n, out_fmt, g.H, g.W, g.xoff);
        e->launches++;
      }
      FSUAE_CUDA_CHECK(e, cudaGetLastError());
      return FSUAE_OK;
    }
  }
  if (plan->zero_Hw != g.Hw || plan->zero_Ww != g.Ww) {   // (re)establish the zero borders for this geometry
    for (size_t i = 0; i < plan->buf.size(); ++i) FSUAE_CUDA_CHECK(e, cudaMemsetAsync(plan->buf[i], 0, plan->buf_bytes[i], st));
    plan->zero_Hw = g.Hw;
    plan->zero_Ww = g.Ww;
  }
  // FSUAE_DBG & 16 (timing experiment, garbage results): every frame aliases frame 0's activation buffers, so all
  // inter-layer traffic stays in L2 -- what the pass would cost if the layers did not stream through HBM
#ifdef FSUAE_DEBUG_SWITCHES
  const bool alias_frames = getenv("FSUAE_DBG") && (atoi(getenv("FSUAE_DBG")) & 16);
#else
  const bool alias_frames = false;
#endif
  auto fstride = [&](int id) { return alias_frames ? 0ull : (unsigned long long)plan->planes[id] * (g.Hw + 2 * BORDER) * PW * 16; };

  const int gin = (flags & FSUAE_FLAG_GAMMA_IN) ? 1 : 0;
  { ProfScope ps(e, st, "head");
  if (d.head == FSUAE_HEAD_UNSHUFFLE2)
    head_unshuffle_bf16_kernel<<<e->sm_count * 8, 256, 0, st>>>(in, plan->buf[0], n, in_fmt, g.H, g.W, g.xoff, g.Hw, g.Ww, PW,
                                                                fstride(0), gin);
  else if (d.head == FSUAE_HEAD_FEATURES)
    head_features_kernel<<<e->sm_count * 8, 256, 0, st>>>((const float*)in, plan->buf[0], n, d.in_channels, g.Hw, g.Ww, PW, fstride(0));
  else
    head_plain_bf16_kernel<<<e->sm_count * 8, 256, 0, st>>>(in, plan->buf[0], n, in_fmt, g.H, g.W, g.xoff, g.Hw, g.Ww, PW,
                                                            fstride(0), gin);
  }
  e->launches++;

  auto fill = [&](int i, const Launch& ln, bool pair) {
    const fsuae_layer_desc& L = d.layers[i];
    LayerK k = ln.k;
    k.Hw = g.Hw; k.Ww = g.Ww; k.PW = PW; k.S = S; k.n_frames = n;
    if (pair) k.wpack = ln.d_w2;
    k.n_blocks = (pair ? (n + 1) / 2 : n) * S * g.Hw;
    k.src0 = plan->buf[L.src0]; k.fs0 = fstride(L.src0);
    if (L.cin1 > 0) { k.src1 = plan->buf[L.src1]; k.fs1 = fstride(L.src1); }
    if (L.skip_src >= 0) { k.skip = plan->buf[L.skip_src]; k.fs_skip = fstride(L.skip_src); }
    if (i + 1 < (int)plan->buf.size()) { k.dst = plan->buf[i + 1]; k.fs_dst = fstride(i + 1); }
    k.frame_in = in; k.frame_out = out; k.in_fmt = in_fmt; k.out_fmt = out_fmt;
    k.H = g.H; k.W = g.W; k.xoff = g.xoff;
    k.gamma_in = gin;
    k.gamma_out = (flags & FSUAE_FLAG_GAMMA_OUT) ? 1 : 0;
#ifdef FSUAE_DEBUG_SWITCHES
    if (const char* dbg = getenv("FSUAE_DBG")) k.dbg = atoi(dbg);
#endif
    return k;
  };
  auto launch_cfg = [&](cudaLaunchConfig_t& cfg, cudaLaunchAttribute* attr, int n_blocks, int ctas, int smem, int threads = NTHREADS) {
    int grid = std::min(n_blocks * ctas, e->sm_count / ctas * ctas);
    if (e->tuning.grid > 0) grid = std::max(ctas, std::min(n_blocks * ctas, e->tuning.grid / ctas * ctas));   // test aid (FSUAE_DEBUG_GRID at creation): force the CTA count
    cfg = cudaLaunchConfig_t{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.numAttrs = 1;
    if (ctas == 2) {
      attr[1].id = cudaLaunchAttributeClusterDimension;
      attr[1].val.clusterDim.x = 2; attr[1].val.clusterDim.y = 1; attr[1].val.clusterDim.z = 1;
      cfg.numAttrs = 2;
    }
    cfg.attrs = attr;
  };
  for (int i = 0; i < d.n_layers; ++i) {
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attr[2];
    if (i == plan->fused_at && n >= 2) {          // layers i and i+1 as one kernel
      const LayerK ka = fill(i, plan->layers[i].launches[0], true);
      const LayerK kb = fill(i + 1, plan->layers[i + 1].launches[0], true);
      launch_cfg(cfg, attr, kb.n_blocks, 2, plan->fused_smem);
      { ProfScope ps(e, st, ("conv" + std::to_string(i + 1) + "+conv" + std::to_string(i + 2) + "_fused_pair").c_str());
        FSUAE_CUDA_CHECK(e, cudaLaunchKernelEx(&cfg, plan->fused_fn, ka, kb)); }
      e->launches++;
      ++i;
      continue;
    }
    for (Launch& ln : plan->layers[i].launches) {
      if (ln.wide) {
        const fsuae_layer_desc& L = d.layers[i];
        WideK k = ln.wk;
        k.Hw = g.Hw; k.Ww = g.Ww; k.PW = PW; k.S = S; k.n_frames = n;
        k.rowblocks = (g.Hw + 7) / 8;
        k.n_units = n * S * k.rowblocks * k.ngroups;
        k.src0 = plan->buf[L.src0]; k.fs0 = fstride(L.src0);
        if (L.cin1 > 0) { k.src1 = plan->buf[L.src1]; k.fs1 = fstride(L.src1); }
        if (L.skip_src >= 0) { k.skip = plan->buf[L.skip_src]; k.fs_skip = fstride(L.skip_src); }
        if (i + 1 < (int)plan->buf.size()) { k.dst = plan->buf[i + 1]; k.fs_dst = fstride(i + 1); }
        k.frame_in = in; k.frame_out = out; k.in_fmt = in_fmt; k.out_fmt = out_fmt; k.H = g.H; k.W = g.W; k.xoff = g.xoff;
        k.gamma_in = gin;
        k.gamma_out = (flags & FSUAE_FLAG_GAMMA_OUT) ? 1 : 0;
        launch_cfg(cfg, attr, k.n_units, 2, ln.wide->smem);
        { ProfScope ps(e, st, ("conv" + std::to_string(i + 1) + "_wide").c_str());
          FSUAE_CUDA_CHECK(e, cudaLaunchKernelEx(&cfg, ln.wide->fn, k)); }
        e->launches++;
        continue;
      }
      const bool r3 = ln.var3 != nullptr;
      const bool pair = !r3 && ln.var2 != nullptr && n >= 2;      // CTA pairs process two frames in lockstep
      const Variant* var = r3 ? ln.var3 : (pair ? ln.var2 : ln.var);
      LayerK k = fill(i, ln, pair);
      if (r3) k.wpack = ln.d_w3;
      launch_cfg(cfg, attr, k.n_blocks, pair ? 2 : 1, var->smem, NTHREADS_L);
      { ProfScope ps(e, st, ("conv" + std::to_string(i + 1) + (r3 ? "_r3" : (pair ? "_pair" : ""))).c_str());
        FSUAE_CUDA_CHECK(e, cudaLaunchKernelEx(&cfg, var->fn, k)); }
      e->launches++;
    }
  }
  if (d.tail == FSUAE_TAIL_FEATURES) {
    tail_features_kernel<<<e->sm_count * 8, 256, 0, st>>>(plan->buf[d.n_layers], (float*)out, n, d.layers[d.n_layers - 1].cout, g.Hw, g.Ww, PW,
                                                          fstride(d.n_layers));
    e->launches++;
  }
  if (g.xoff > 0) {
    black_columns_bf16_kernel<<<64, 256, 0, st>>>(out, n, out_fmt, g.H, g.W, g.xoff);
    e->launches++;
  }
  FSUAE_CUDA_CHECK(e, cudaGetLastError());
  return FSUAE_OK;
}

}  // namespace fsuae
