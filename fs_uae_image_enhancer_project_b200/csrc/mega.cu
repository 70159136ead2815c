// The single fused pass of the flagship network (its own translation unit; compiled once per operand type).
//
// Reference semantics: model/model_pix_shuffle.py:227-298 executed as ONE persistent kernel -- input normalisation (gamma LUT,
// PixelUnshuffle), conv1..conv7 with their activation chains / residual adds / the long-skip concat, PixelShuffle, global
// residual, ReLU, gamma, uint8 pack.  HBM is touched for the input frame and the output frame only: every inter-layer feature
// map lives in a small ring of image rows that is overwritten in place every few microseconds and therefore never leaves L2.
//
// Organisation ("stage pipeline through L2 rings"):
//   * The layer kernels of bf16_tc.cu stay what they are -- a strip-marching implicit GEMM on a CTA pair (cta_group::2,
//     two frames in lockstep) with a shared-memory ring of input rows, resident weights and TMEM accumulator stages -- but
//     here each one is an ENGINE: a set of warps (producer, MMA issuer / peer relay, two or four epilogue warpgroups) with
//     its own barriers, ring, weights and TMEM columns.  A CTA hosts one or two engines side by side; the hardware warp
//     scheduler interleaves them, so an MMA-bound layer runs under an activation-bound one:
//         stage A = conv5 (no activation: tensor-bound)          + conv2 (TeLU, SinLU, BiasedPReLU: SFU-bound)
//         stage B = conv4 (Mish, BiasedPReLU, Tanh: both)        + the network head on the two service warps a second engine would use
//         stage C = conv3 (no activation)                        + conv7 (PixelShuffle tail: store / pow-bound)
//         stage D = conv6 (Mish)                                 + conv1 (SinLU)
//     MMA instructions per strip row: A 42+24, B 42, C 24+24, D 45+9 (see the budget table at the engine definitions).
//   * 4 stage pairs = 1 group = one 126-column strip of a frame pair; S groups (S strips) = 1 team = whole rows of a frame
//     pair; the rows of all frame pairs are cut into equal contiguous ranges, one per team.  Where a range starts or ends
//     inside a frame, layer i recomputes 7-i halo rows (1.3 % extra MMAs at 64 frames) -- teams never talk to each other.
//   * Engines talk through CHANNELS in global memory (7 of them: the outputs of conv1..conv6 and the head's unshuffled input
//     frame): the producing epilogue stores row q of its output into slot q % D of a D-row ring (chunk-planar, full image
//     width, zero borders baked in exactly as in the per-layer buffers); the last warp to finish a group of MG_PUB_ROWS rows
//     bumps their per-slot counters behind ONE gpu-scope release (cumulative over a CTA-scope counter, so no other warp ever
//     waits for its stores to reach L2).  The consuming layer's producer warp polls those counters with relaxed loads (an
//     acquire LOAD invalidates the SM's L1 every time), fences once per successful poll, issues its TMA bulk copies, and
//     hands out credits ("rows below b have landed in my shared memory") so the ring slot can be overwritten.  All S strips
//     of a row must be complete before any strip of the next layer reads it (the 3x3 taps reach one column into the
//     neighbouring strips).
//   * Rings: MG_D1 = 13 rows per channel, MG_D0 = 56 for conv1 -> conv2 / conv6 (the long skip spans the whole pipeline):
//     4.5 MB per (team, frame of the pair), 54 MB for the 6 teams a B200 holds.  The depth is what hides the credit loop
//     (store -> release -> poll -> copy -> credit -> poll); measured at 64 frames, D0/D1 -> DRAM bytes per pass / time per
//     frame: 48/12 -> 0.31 GB / 40.1 us, 56/13 -> 0.45 / 38.5, 64/16 -> 1.15 / 37.9, 64/20 -> 2.33 / 37.6 (the layer kernels:
//     7.7 GB / 37.2-38.2 us; algorithmic: 0.22 GB).  Past ~55 MB of rings the dirty lines of consumed rows start to be
//     evicted before they are overwritten.

#include <algorithm>

#include "mega_params.h"
#include "tc_common.cuh"

namespace fsuae {
namespace {


// ---- flags ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t ld_acquire_gpu(const unsigned int* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// Polls use relaxed loads (served by L2, no side effects); the one acquire fence after a successful poll orders what follows.
// An acquire LOAD invalidates the SM's whole L1 (CCTL.IVALL) every time it is executed -- in a polling loop that stalls
// every other memory instruction of the SM.
__device__ __forceinline__ uint32_t ld_relaxed_gpu(const unsigned int* p) {
  uint32_t v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
__device__ __forceinline__ void st_relaxed_gpu(unsigned int* p, uint32_t v) {
  asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void red_relaxed_gpu_add(unsigned int* p, uint32_t v) {
  asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void red_release_gpu_add(unsigned int* p, uint32_t v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_cta_smem(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t atom_acq_rel_cta_smem_add(uint32_t* p, uint32_t v) {
  uint32_t old;
  asm volatile("atom.acq_rel.cta.shared::cta.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(smem_u32(p)), "r"(v) : "memory");
  return old;
}
__device__ __forceinline__ unsigned int* mg_prod(unsigned int* fl, int ch) { return fl + ch * MG_DMAX; }

// Cycle accounting of one probe CTA per stage (team 0, middle strip, leader): builds with -DFSUAE_EPI_TIMING only.
// Per layer: [0] issuer total  [1] issuer waits for input rows  [2] ... for a drained accumulator  [3] issue + commits
// [4] producer waits for a free ring slot  [5] ... for the upstream layer's rows (flags)  [6] producer total
// [7] epilogue warp 0 waits for the downstream ring (back-pressure)  [8] ... for the accumulator  [9] math + stores + signals
// [10] epilogue total  [11] rows of that warp  [12] producer rows  [13] issuer rows  [14] writer-side proxy fence  [15] release of the row counter
// Row 7 (the head and extras): [0..4] see g_mega_t, [5..11] credit step of conv1..conv7's producers, [12] conv4 relay wait, [13] head worker 0 total
#ifdef FSUAE_EPI_TIMING
__device__ unsigned long long g_mega_t[8][16];     // row 7: the head producer in detail ([0] next row + loads  [1] slot wait  [2] LUT + stores  [3] fence + arrive  [4] rows)
#define MG_T(var) const long long var = clock64()
#define MG_ACC(cond, L, i, v) do { if (cond) atomicAdd(&g_mega_t[(L) - 1][i], (unsigned long long)(v)); } while (0)   // RED: fire and forget, a load-add-store would stall the probed warp for a memory round trip
#else
#define MG_T(var)
#define MG_ACC(cond, L, i, v)
#endif
#ifdef MG_DBG_FINISH      // timing experiment: cycles every CTA took from entry to its last engine's last row (per-stage run time),
__device__ long long g_mega_finish[160];            // and its entry / exit on the global timer for the last four launches
__device__ unsigned long long g_mega_gt[4][2][160];
__device__ unsigned long long g_mega_rows[2][8];    // probe CTA: global timer when the head / conv1..7 finished their first and last row
__device__ unsigned int g_mega_launch;
__device__ __forceinline__ unsigned long long mg_globaltimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#endif
__device__ __forceinline__ unsigned int* mg_cons(unsigned int* fl, int ch, int consumer) {
  return fl + MG_NCH * MG_DMAX + (ch * 2 + consumer) * MG_SMAX;
}

// ---- work partition ---------------------------------------------------------------------------------------------------
// The n_fp * Hw image rows (frame pair major) are cut into `teams` equal contiguous ranges; a range is walked as segments
// that stay inside one frame pair.  Layer i (halo h = 7 - i) computes rows [max(0, y0 - h), min(Hw, y0 + rows + h)).
struct MSeg { int fp, y0, rows; };
struct MSegIter {
  int cur, end, Hw;
  __device__ MSegIter(const MegaK& M, int team) : Hw(M.Hw) {
    const long long tot = (long long)M.n_fp * M.Hw;
    cur = (int)(tot * team / M.teams);
    end = (int)(tot * (team + 1) / M.teams);
  }
  __device__ bool next(MSeg& g) {
    if (cur >= end) return false;
    g.fp = cur / Hw;
    g.y0 = cur - g.fp * Hw;
    g.rows = min(Hw - g.y0, end - cur);
    cur += g.rows;
    return true;
  }
};
__device__ __forceinline__ int mg_lo(const MSeg& g, int halo) { return max(0, g.y0 - halo); }
__device__ __forceinline__ int mg_hi(const MSeg& g, int halo, int Hw) { return min(Hw, g.y0 + g.rows + halo); }

// ---- one engine = one layer ---------------------------------------------------------------------------------------------
// LAYER 1..7.  Sources: conv1 <- channel 6 (the head's output), conv{2,3,4,5} <- channel LAYER-2, conv6 <- channels 0 (conv1, long skip)
// and 4 (conv5), conv7 <- channel 5.  Consumer index inside a channel: conv6 is consumer 1 of channel 0, everything else 0.
// Per-channel epilogue parameters live in shared memory as [9][MG_MAXC] floats (bias, p0[4 slots], p1[4 slots]): the epilogue
// walks the 8-channel chunks in a ROLLED loop.  Unrolled over 36-72 compile-time channels the code of one stage is ~70 KB of
// straight-line instructions that a handful of warps run through once per row -- they then wait for the instruction cache
// more than for anything else (ncu: stall_no_inst).
constexpr int MG_PRM_BYTES = (9 * MG_MAXC * 4 + 127) / 128 * 128;
template <int LAYER, int PT_, int NPAD_, int COUT_, int KIND_, class EPI_, int RING_, int STAGES_, int NWG_, bool RM_ = true, bool SPLIT_ = false>
struct MEng {
  static constexpr bool RM = RM_;      // input-row-major MMA order (three accumulators open at a time) / block-major
  // SPLIT: two warpgroups drain one output row together (the channel chunks are cut in two), so an accumulator stage is held
  // for half as long.  With the row-major order only STAGES - 3 stages are being drained at a time: the row time of a layer
  // with a long activation chain is (drain + hand-over) / (STAGES - 3), whatever the number of warpgroups.
  static constexpr bool SPLIT = SPLIT_;
  static constexpr int NG = SPLIT_ ? NWG_ / 2 : NWG_;      // row groups: output row n is drained by group n % NG
  static constexpr int WPR = SPLIT_ ? 8 : 4;               // epilogue warps per output row
  using EPI = EPI_;
  static constexpr int L = LAYER, PT = PT_, NPAD = NPAD_, COUT = COUT_, KIND = KIND_, RING = RING_, STAGES = STAGES_, NWG = NWG_;
  static constexpr int HALO = 7 - LAYER;
  static constexpr int STEPS_ROW = (3 * PT + 1) / 2, STEPS = 3 * STEPS_ROW, NB = NPAD / 2;
  static constexpr int WBYTES = STEPS * NB * 32, ROWBYTES = PT * PLANE_ROW;
  static constexpr int PRM_OFF = WBYTES + RING * ROWBYTES + 128;    // behind the 64-byte overrun pad of the ring (descriptors of the last units reach past it)
  static constexpr int SMEM = PRM_OFF + MG_PRM_BYTES;               // + this layer's per-channel epilogue parameters
  static constexpr int TCOLS = STAGES * NPAD;
  static constexpr int NBARS = 3 * RING + 2 * STAGES + 1;           // full, empty, pfull; tfull, tempty; wbar
  static constexpr int IN0 = LAYER == 1 ? 6 : (LAYER == 6 ? 0 : LAYER - 2);
  static constexpr int IN1 = LAYER == 6 ? 4 : -1;
  static constexpr int P0 = LAYER == 6 ? 5 : PT, P1 = LAYER == 6 ? 5 : 0;
  static constexpr int HALO0 = LAYER == 6 ? 6 : HALO + 1;           // halo of the layer that produced source 0 / 1
  static constexpr int HALO1 = HALO + 1;
  static constexpr int CONS0 = LAYER == 6 ? 1 : 0;                  // my consumer index in source channel 0 (source 1: always 0)
  static constexpr int OUT = LAYER <= 6 ? LAYER - 1 : -1;
  static constexpr int OUT_PLANES = (COUT + 7) / 8;
  static constexpr int OUT_NCONS = LAYER == 1 ? 2 : 1;              // conv1's output is read by conv2 and conv6
  static_assert(STAGES >= NWG, "a warpgroup's previous block must be at least one use of the stage back (mbarrier parity waits)");
  static_assert(!SPLIT_ || (NWG_ % 2 == 0 && KIND_ == EPI_STORE), "split drain: pairs of warpgroups, store epilogue");
  static_assert(!RM || STAGES >= 4, "row-major order: three open accumulators and at least one being drained");
  static_assert(RING >= 4 && RING <= 16, "ring depth (one relay lane per slot)");
  static_assert(NB % 8 == 0, "each CTA of the pair holds whole core matrices of B");
};

// Phase of a layer's publish groups.  Output rows are released MG_PUB_ROWS at a time, and layer L + 1 needs layer L's rows up
// to y + 1 to finish its own row y: with every layer's groups ending on the same rows ([0..3], [4..7], ...) layer L + 1 can
// finish row 3 -- and release ITS first group -- only after layer L's SECOND group, so every hop of the pipeline lags a whole
// group (4 rows = 5 us) behind the one before, at the start of a pass and again at its end.  Layer L's groups therefore end
// `step` rows earlier than layer L - 1's ([0..3], [0..2] [3..6], [0..1] [2..5], ...): step = 1 when the team's range starts at
// the top of a frame (sequence number q = image row y in every layer), 2 when it starts inside one (layer L - 1 has one halo
// row more in front, so its q runs one ahead).  Any phase is correct; this one takes ~35 us out of every pass.
#ifndef MG_PUB_ROWS
#define MG_PUB_ROWS 4      // rows per release of an engine's output (one fence for all of them)
#endif
#ifndef MG_PUB_PHASE
#define MG_PUB_PHASE 1
#endif
__device__ __forceinline__ uint32_t mg_pub_phase(const MegaK& M, int team, int layer) {
  if (!MG_PUB_PHASE) return 0u;
  const long long first = (long long)M.n_fp * M.Hw * team / M.teams;      // first row of the team's range (see MSegIter)
  const uint32_t step = (first % M.Hw) != 0 ? 2u : 1u;
  return ((uint32_t)layer * step) % (uint32_t)MG_PUB_ROWS;
}
// The two issuers of a stage take the tensor pipe in turn, one input row at a time: interleaved instruction by instruction,
// each engine's MMA evicts the other's A tile from the collector (fill / use / lastuse triples, see mg_issuer_rm) and every
// instruction pays the full shared-memory operand read again.
#ifndef MG_MMA_LOCK
#define MG_MMA_LOCK 0xD     // bit i: stage i (A = 0 .. D = 3)
#endif
struct MEngSmem {      // shared-memory carve-up of one engine
  uint8_t* w;
  uint8_t* ring;
  const float* prm;      // [9][MG_MAXC]
  uint64_t *full, *empty, *pfull, *tfull, *tempty, *wbar;
  uint32_t* rowcnt;      // [MG_NR] epilogue warps that have stored their part of output row q (slot q % MG_NR, never reset)
  uint32_t* mmalock;     // MG_MMA_LOCK: the stage's MMA-stream lock (nullptr in a stage with one engine)
  uint32_t ph;           // phase of this engine's publish groups (see mg_pub_phase)
};
__device__ __forceinline__ void mg_lock(uint32_t* l) {
  if (!MG_MMA_LOCK || l == nullptr) return;
  uint32_t old;
  do {
    asm volatile("atom.acquire.cta.shared::cta.cas.b32 %0, [%1], 0, 1;" : "=r"(old) : "r"(smem_u32(l)) : "memory");
  } while (old != 0u);
}
__device__ __forceinline__ void mg_unlock(uint32_t* l) {
  if (!MG_MMA_LOCK || l == nullptr) return;
  asm volatile("st.release.cta.shared::cta.u32 [%0], %1;" ::"r"(smem_u32(l)), "r"(0u) : "memory");
}
constexpr int MG_NR = 8;   // counters of MG_NR row groups in flight: group g + MG_NR cannot reach its epilogue before group g has released its stages
template <class E>
__device__ __forceinline__ MEngSmem mg_carve(uint8_t* data, uint64_t* bars, uint32_t* rowcnt) {
  static_assert(E::STAGES <= MG_NR, "row counters vs accumulator stages");
  MEngSmem s;
  s.rowcnt = rowcnt;
  s.mmalock = nullptr;
  s.ph = 0u;
  s.w = data;
  s.ring = data + E::WBYTES;
  s.prm = reinterpret_cast<const float*>(data + E::PRM_OFF);
  s.full = bars;
  s.empty = bars + E::RING;
  s.pfull = bars + 2 * E::RING;
  s.tfull = bars + 3 * E::RING;
  s.tempty = s.tfull + E::STAGES;
  s.wbar = s.tempty + E::STAGES;
  return s;
}
template <class E>
__device__ __forceinline__ void mg_init_bars(const MEngSmem& s) {     // one thread
  // a ring row is released by the MMA commit and, when the residual is read from it, by the 4 epilogue warps of its block
  for (int i = 0; i < E::RING; ++i) { mbar_init(&s.full[i], 1); mbar_init(&s.empty[i], E::EPI::kSkip ? 1 + E::WPR : 1); mbar_init(&s.pfull[i], 1); }
  for (int i = 0; i < E::STAGES; ++i) { mbar_init(&s.tfull[i], 1); mbar_init(&s.tempty[i], 2 * E::WPR); }      // the row's epilogue warps in each CTA of the pair
  mbar_init(s.wbar, 1);
}

struct MCtx {          // where this CTA sits
  int team, strip, rank, f;       // f: the frame this CTA works on inside the current segment is 2 * fp + rank (clamped)
  bool probe;                     // this CTA feeds the cycle counters (FSUAE_EPI_TIMING builds)
  unsigned char* scratch;         // this (team, rank)'s channel block
  unsigned int* flags;            // this (team, rank)'s flag block
};

// Wait until rows [q, q + 1) of channel `ch` are complete in every strip.  Whole warp; caches how far the ring is known to be
// complete (`upto`, exclusive) so that one coalesced poll releases several rows.
// A successful poll costs two fences (~1500 cycles) on the one warp that feeds the engine, so it asks for MG_POLL_BATCH rows
// at a time (fewer at the end of the channel: `qtotal` rows in all).
// Timing experiment (garbage results): bit ch set = nobody waits on channel ch, neither for its rows nor for its credits, so
// the engines on either side run at their own pace (all bits: every engine free-running -> its stand-alone row time).
#ifndef MG_DBG_CUT
#define MG_DBG_CUT 0
#endif
// Acquire fence behind a successful back-pressure poll of a WRITER (epilogue warps, head).  What the poll guards is a
// write-after-read: the slot about to be overwritten was read by the consumer's bulk copy, whose completion the consumer
// observed on its mbarrier BEFORE it stored the credit this poll has just seen.  The overwriting stores depend on the polled
// value through the loop exit and are not performed speculatively, so nothing is left for a fence to order -- and a
// fence.acq_rel.gpu here first waits for all the warp's stores of the previous row to reach L2 (1000+ cycles per row, on the
// epilogue's critical path: 38.9 -> 37.8 us/frame without it).  The READER side (mg_fence_after_poll) keeps its fences: there
// later loads -- the bulk copies -- must not pass the poll.
#ifndef MG_BP_FENCE
#define MG_BP_FENCE 0
#endif
// 36-channel plain layer (conv5: ONE warpgroup drains every row, one accumulator stage beyond the three open ones): the whole
// accumulator row goes to registers and the stage is released before bias / pack / store (38.9 -> 38.0 us/frame).
#ifndef MG_EARLY_REL
#define MG_EARLY_REL 1
#endif
#ifndef MG_POLL_ACQ
#define MG_POLL_ACQ 1
#endif
#ifndef MG_POLL_PROXY
#define MG_POLL_PROXY 1
#endif
#ifndef MG_POLL_BATCH
#define MG_POLL_BATCH 3
#endif
// Returns true when it polled (the caller then fences once for all its sources).
__device__ __forceinline__ bool mg_wait_rows(const MegaK& M, const MCtx& c, int ch, uint32_t q, uint32_t qtotal, uint32_t& upto, int lane) {
  if (q < upto) return false;
  if ((MG_DBG_CUT >> ch) & 1) { upto = qtotal; return false; }      // timing experiment: this link is cut
  const uint32_t want = min(q + (uint32_t)MG_POLL_BATCH, qtotal) - q;
  const uint32_t D = (uint32_t)mg_depth(ch), need_per_use = (uint32_t)M.S;          // every strip's publisher bumps the counter once per row
  const unsigned int* prod = mg_prod(c.flags, ch);
  const long long t0 = clock64();
  for (;;) {
    bool ok = false;
    if ((uint32_t)lane < D) {
      const uint32_t qq = q + (uint32_t)lane;
      ok = ld_relaxed_gpu(prod + qq % D) >= need_per_use * (qq / D + 1u);
    }
    const uint32_t mask = __ballot_sync(0xffffffffu, ok);
    const uint32_t cnt = (uint32_t)__ffs((int)~mask) - 1u;      // consecutive complete rows from q on (all ones: ffs(0) = 0 -> handled below)
    const uint32_t n = mask == 0xffffffffu ? 32u : cnt;
    if (n >= want) { upto = q + n; break; }
    if (clock64() - t0 > (1ll << 31)) __trap();
    __nanosleep(64);
  }
  return true;
}
// Once per successful poll (it usually reveals several rows, of both sources of conv6): order the bulk copies (async proxy)
// behind the counters just read (generic proxy).  The two fences cost ~1500 cycles, so they must not sit between "ring slot
// free" and the copy.
__device__ __forceinline__ void mg_fence_after_poll(int lane) {
#if MG_POLL_ACQ
  fence_acq_rel_gpu();
#endif
#if MG_POLL_PROXY
  if (lane == 0) fence_proxy_async_all();
#endif
  __syncwarp();
}

// ---- producer warp: channel rows -> shared-memory ring (TMA) -----------------------------------------------------------
template <class E>
__device__ void mg_producer(const MegaK& M, const MCtx& c, const MEngSmem& s, int lane) {
  const MegaLayerP& LP = M.L[E::L - 1];
  if (lane == 0) {
    mbar_arrive_expect_tx(s.wbar, E::WBYTES);
    tma_load_1d(s.w, LP.wpack + (size_t)c.rank * E::WBYTES, E::WBYTES, s.wbar);
  }
  const size_t row_pitch = (size_t)M.PW * 16;
  const size_t xoff_b = (size_t)(c.strip * STRIP - 1 + BORDER) * 16;       // first slot of this strip's 128-slot window
  const unsigned char* ch0 = c.scratch + M.ch_off[E::IN0];
  const size_t pp0 = (size_t)mg_depth(E::IN0) * row_pitch;
  const unsigned char* ch1 = E::IN1 >= 0 ? c.scratch + M.ch_off[E::IN1 >= 0 ? E::IN1 : 0] : nullptr;
  const size_t pp1 = (size_t)mg_depth(E::IN1 >= 0 ? E::IN1 : 0) * row_pitch;
  uint32_t fill = 0;                 // ring fills so far
  uint32_t qb0 = 0, qb1 = 0;         // sequence number of the first row the source layers produce in this segment
  uint32_t upto0 = 0, upto1 = 0;
  uint32_t tot0 = 0, tot1 = 0;       // rows the source layers produce in all
  {
    MSegIter it0(M, c.team);
    MSeg g0;
    while (it0.next(g0)) {
      tot0 += (uint32_t)(mg_hi(g0, E::HALO0, M.Hw) - mg_lo(g0, E::HALO0));
      tot1 += (uint32_t)(mg_hi(g0, E::HALO1, M.Hw) - mg_lo(g0, E::HALO1));
    }
  }
  uint32_t land0[2] = {0, 0}, land1[2] = {0, 0};     // rows of source 0 / 1 below this have been requested by fill - 1 / fill - 2 (lane 0)
  uint32_t pub0 = 0, pub1 = 0;                       // ... and this is what the producers have been told
  const bool tprobe = c.probe && lane == 0;
  MG_T(tp_begin);
  MSegIter it(M, c.team);
  MSeg g;
  while (it.next(g)) {
    const int lo = mg_lo(g, E::HALO), hi = mg_hi(g, E::HALO, M.Hw);
    const int lo0 = mg_lo(g, E::HALO0), lo1 = mg_lo(g, E::HALO1);
    for (int y = lo - 1; y <= hi; ++y) {       // input rows of output rows lo .. hi-1
      const uint32_t slot = fill % E::RING, par = ((fill / E::RING) & 1u) ^ 1u;
      const bool inframe = y >= 0 && y < M.Hw;
      const uint32_t q0 = qb0 + (uint32_t)(y - lo0), q1 = qb1 + (uint32_t)(y - lo1);
      MG_T(tp0);
      if (lane == 0) {
        // Credits for the producing layer(s): a channel row may be overwritten once my bulk copy of it has landed.  The copy
        // of fill - 2 has normally landed long ago (the wait below is a formality), so the producers may run D - 3 rows
        // ahead of my load pointer.
        if (fill >= 2u) {
          const uint32_t f2 = fill - 2u;
          mbar_wait(&s.full[f2 % E::RING], (f2 / E::RING) & 1u);
          // relaxed: the store cannot be performed before the wait above has returned, and that wait is what it reports
          if (land0[1] > pub0) { pub0 = land0[1]; st_relaxed_gpu(mg_cons(c.flags, E::IN0, E::CONS0) + c.strip, pub0); }
          if constexpr (E::IN1 >= 0)
            if (land1[1] > pub1) { pub1 = land1[1]; st_relaxed_gpu(mg_cons(c.flags, E::IN1 >= 0 ? E::IN1 : 0, 0) + c.strip, pub1); }
        }
        land0[1] = land0[0]; land1[1] = land1[0];
        if (inframe) { land0[0] = q0 + 1u; land1[0] = q1 + 1u; }
      }
      __syncwarp();
      MG_T(tp0b);
      MG_ACC(tprobe, 8, 4 + E::L, tp0b - tp0);      // row 7, [5 .. 11]: the credit step (waits for the copy of fill - 2)
      if (inframe) {       // usually known from an earlier poll: the upstream layer runs ahead while my ring is full
        bool polled = mg_wait_rows(M, c, E::IN0, q0, tot0, upto0, lane);
        if constexpr (E::IN1 >= 0) polled = mg_wait_rows(M, c, E::IN1 >= 0 ? E::IN1 : 0, q1, tot1, upto1, lane) || polled;
        if (polled) mg_fence_after_poll(lane);
      }
      MG_T(tp1);
      if (lane == 0) mbar_wait(&s.empty[slot], par);
      MG_T(tp2);
      MG_ACC(tprobe, E::L, 5, tp1 - tp0b); MG_ACC(tprobe, E::L, 4, tp2 - tp1); MG_ACC(tprobe, E::L, 12, 1);
      // one bulk copy per plane, one lane each: a single issue slot for the whole row instead of PT in a row
      if (lane == 0) mbar_arrive_expect_tx(&s.full[slot], E::ROWBYTES);
      __syncwarp();
      if (lane < E::PT) {
        uint8_t* d = s.ring + slot * E::ROWBYTES + lane * PLANE_ROW;
        const unsigned char* src = M.zero_row;
        if (inframe) {
          if (lane < E::P0) src = ch0 + (size_t)(q0 % (uint32_t)mg_depth(E::IN0)) * row_pitch + xoff_b + (size_t)lane * pp0;
          else src = ch1 + (size_t)(q1 % (uint32_t)mg_depth(E::IN1 >= 0 ? E::IN1 : 0)) * row_pitch + xoff_b + (size_t)(lane - E::P0) * pp1;
        }
        tma_load_1d(d, src, PLANE_ROW, &s.full[slot]);
      }
      ++fill;
    }
    qb0 += (uint32_t)(mg_hi(g, E::HALO0, M.Hw) - lo0);
    qb1 += (uint32_t)(mg_hi(g, E::HALO1, M.Hw) - lo1);
  }
  MG_T(tp_end);
  MG_ACC(tprobe, E::L, 6, tp_end - tp_begin);
}

// ---- the network head: frame -> gamma LUT -> PixelUnshuffle(2) -> channel 6 ------------------------------------------------
// Two warps (the service warps stage B has no second engine for), each converts one half of every strip row: a slot is one
// half-resolution pixel = 2x2 full-resolution pixels x RGB = 12 channels = plane 0 (8 channels) + plane 1 (4 channels, 4 zeros).
// One warp cannot keep a stage fed: ~500 dependent instructions per row are ~3000 cycles however they are arranged.
// Framebuffer format (uint8 RGBA, the streaming case): the raw pixels travel global -> shared memory by cp.async two rows
// ahead (no registers in between), issued behind the row's stores; out-of-frame columns are zero filled -> LUT[0] = 0.
__device__ void mg_head_worker(const MegaK& M, const MCtx& c, const float* s_lut, uint4* s_raw, int w, int lane) {
  constexpr int CH = 6, HALO = 7;
  constexpr uint32_t D = (uint32_t)mg_depth(CH);
  const size_t row_pitch = (size_t)M.PW * 16, plane_pitch = (size_t)D * row_pitch, fpl = (size_t)M.H * M.W;
  unsigned char* och = c.scratch + M.ch_off[CH];
  unsigned int* prod = mg_prod(c.flags, CH);
  const bool fb = M.in_fmt == FSUAE_FMT_U8_NHWC4;
  const bool tprobe = c.probe && w == 0 && lane == 0;
  // my two slots of a strip row: m = 63 w + lane, 63 w + 32 + lane (the second one only for lane < 31)
  int mm[2], xx[2];
  bool okx[2];
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    mm[k] = 63 * w + 32 * k + lane;
    xx[k] = c.strip * STRIP + mm[k];
    okx[k] = (k == 0 || lane < 31) && xx[k] < M.Ww;
  }
  uint4* stage = s_raw + 64 * w;                 // my half of the three staging rows
  // the CTA's row sequence: every segment's rows lo .. hi-1 (in-frame rows only; conv1's producer supplies the zero rows)
  MSegIter it(M, c.team);
  MSeg g;
  int ycur = 0, yend = 0;
  auto next_row = [&](int& rf, int& ry) -> bool {
    if (ycur >= yend) {
      if (!it.next(g)) return false;
      ycur = mg_lo(g, HALO);
      yend = mg_hi(g, HALO, M.Hw);
    }
    rf = min(2 * g.fp + c.rank, M.n_frames - 1);
    ry = ycur++;
    return true;
  };
  auto prefetch = [&](int rf, int ry, uint4* buf) {
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const unsigned char* ip = (const unsigned char*)M.frame_in + (okx[k] ? ((size_t)rf * fpl + (size_t)(2 * ry) * M.W + 2 * xx[k] + M.xoff) * 4 : 0);
      const uint32_t dst = smem_u32(buf + 32 * k + lane), nb = okx[k] ? 8u : 0u;
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(ip), "r"(nb) : "memory");
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst + 8u), "l"(ip + (size_t)M.W * 4), "r"(nb) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  const bool f32 = M.in_fmt == FSUAE_FMT_F32_NCHW3;
  float cv[2][12];                               // float frames: the current row's values of my two slots
  auto ldrow_f32 = [&](int rf, int ry, float (&v)[2][12]) {
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      if (!okx[k]) continue;
#pragma unroll
      for (int dy = 0; dy < 2; ++dy) {
        const float* ip = (const float*)M.frame_in + (size_t)rf * 3 * fpl + (size_t)(2 * ry + dy) * M.W + 2 * xx[k] + M.xoff;
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
          const float2 t2 = __ldg(reinterpret_cast<const float2*>(ip + ch * fpl));
          v[k][ch * 4 + dy * 2] = t2.x; v[k][ch * 4 + dy * 2 + 1] = t2.y;
        }
      }
    }
  };
  int fC, yC, fN = 0, yN = 0, fP = 0, yP = 0;
  bool haveC = next_row(fC, yC), haveN = false;
  uint32_t q = 0, cons_seen = 0, qtotal = 0;
  {
    MSegIter itq(M, c.team);
    MSeg gq;
    while (itq.next(gq)) qtotal += (uint32_t)(mg_hi(gq, HALO, M.Hw) - mg_lo(gq, HALO));
  }
  if (haveC && fb) {
    prefetch(fC, yC, stage);
    haveN = next_row(fN, yN);
    if (haveN) prefetch(fN, yN, stage + MROWS);
  } else if (haveC) {
    haveN = next_row(fN, yN);
    if (f32) ldrow_f32(fC, yC, cv);
  }
  MG_T(th_begin);
  while (haveC) {
    const bool haveP = haveN && next_row(fP, yP);
    MG_T(th0);
    // back-pressure: slot q % D still holds row q - D until conv1's producers of strips s-1, s, s+1 have loaded it
    if (!((MG_DBG_CUT >> CH) & 1) && q >= D && cons_seen < q - D + 1u) {
      const uint32_t need = q - D + 1u;
      const long long t0 = clock64();
      for (;;) {
        uint32_t v = 0xFFFFFFFFu;
        if (lane < 3) {
          const int sn = c.strip - 1 + lane;
          if (sn >= 0 && sn < M.S) v = ld_relaxed_gpu(mg_cons(c.flags, CH, 0) + sn);
        }
        v = __reduce_min_sync(0xffffffffu, v);
        if (v >= need) { cons_seen = v; if (MG_BP_FENCE) fence_acq_rel_gpu(); break; }
        if (clock64() - t0 > (1ll << 31)) __trap();
        __nanosleep(64);
      }
    }
    MG_T(th1);
    unsigned char* dp = och + (size_t)(q % D) * row_pitch;
    if (fb) {
      if (haveN) asm volatile("cp.async.wait_group 1;" ::: "memory"); else asm volatile("cp.async.wait_group 0;" ::: "memory");
      const uint4* src = stage + (q % 3u) * MROWS;
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const uint4 cur = src[32 * k + lane];          // written by this very lane: {dy 0: px 0, px 1; dy 1: px 0, px 1}
        float v[12];
#ifdef MG_DBG_NOHEAD      // timing experiment (garbage results): the head without its table look-ups
#pragma unroll
        for (int ch = 0; ch < 12; ++ch) v[ch] = __uint_as_float(cur.x);
#else
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
          v[ch * 4 + 0] = s_lut[(cur.x >> (8 * ch)) & 0xFF];
          v[ch * 4 + 1] = s_lut[(cur.y >> (8 * ch)) & 0xFF];
          v[ch * 4 + 2] = s_lut[(cur.z >> (8 * ch)) & 0xFF];
          v[ch * 4 + 3] = s_lut[(cur.w >> (8 * ch)) & 0xFF];
        }
#endif
        if (okx[k]) {
          unsigned char* d = dp + (size_t)(xx[k] + BORDER) * 16;
          *reinterpret_cast<uint4*>(d) = make_uint4(pack_op2(v[0], v[1]), pack_op2(v[2], v[3]), pack_op2(v[4], v[5]), pack_op2(v[6], v[7]));
          *reinterpret_cast<uint4*>(d + plane_pitch) = make_uint4(pack_op2(v[8], v[9]), pack_op2(v[10], v[11]), 0u, 0u);
        }
      }
    } else if (f32) {
      // float frames: the NEXT row's 24 values per lane are requested before this row is stored (registers, one row = ~1.4 us
      // ahead: about one DRAM round trip), so the loads are in flight under the stores, the barrier and the release
      float nv[2][12];
      if (haveN) ldrow_f32(fN, yN, nv);
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        if (okx[k]) {
          unsigned char* d = dp + (size_t)(xx[k] + BORDER) * 16;
          *reinterpret_cast<uint4*>(d) = make_uint4(pack_op2(cv[k][0], cv[k][1]), pack_op2(cv[k][2], cv[k][3]), pack_op2(cv[k][4], cv[k][5]), pack_op2(cv[k][6], cv[k][7]));
          *reinterpret_cast<uint4*>(d + plane_pitch) = make_uint4(pack_op2(cv[k][8], cv[k][9]), pack_op2(cv[k][10], cv[k][11]), 0u, 0u);
        }
      }
#pragma unroll
      for (int k = 0; k < 2; ++k)
#pragma unroll
        for (int i = 0; i < 12; ++i) cv[k][i] = nv[k][i];
    } else {
      // planar uint8 frames (not a streaming format): this path only has to be correct
#pragma unroll 1
      for (int k = 0; k < 2; ++k) {
        if (!okx[k]) continue;
        float v[12];
#pragma unroll
        for (int dy = 0; dy < 2; ++dy) {
          const size_t p0 = (size_t)(2 * yC + dy) * M.W + 2 * xx[k] + M.xoff;
          const unsigned char* ip = (const unsigned char*)M.frame_in + (size_t)fC * 4 * fpl + p0;
#pragma unroll
          for (int ch = 0; ch < 3; ++ch) { v[ch * 4 + dy * 2] = s_lut[__ldg(ip + ch * fpl)]; v[ch * 4 + dy * 2 + 1] = s_lut[__ldg(ip + ch * fpl + 1)]; }
        }
        unsigned char* d = dp + (size_t)(xx[k] + BORDER) * 16;
        *reinterpret_cast<uint4*>(d) = make_uint4(pack_op2(v[0], v[1]), pack_op2(v[2], v[3]), pack_op2(v[4], v[5]), pack_op2(v[6], v[7]));
        *reinterpret_cast<uint4*>(d + plane_pitch) = make_uint4(pack_op2(v[8], v[9]), pack_op2(v[10], v[11]), 0u, 0u);
      }
    }
    MG_T(th2);
    // both halves stored -> the row counters, four rows per release: the fence behind it costs ~1000 cycles (every store of
    // the SM must have reached L2) and the head runs rows ahead of conv1 anyway.  The two warps take it in turn; its gpu scope
    // is cumulative over the CTA-scope barrier, so it covers the other warp's stores as well.
    asm volatile("bar.sync 1, 64;" ::: "memory");
    if (((q & 3u) == 3u || q + 1u == qtotal) && ((q >> 2) & 1u) == (uint32_t)w && lane == 0) {
      fence_acq_rel_gpu();
      for (uint32_t qq = q & ~3u; qq <= q; ++qq) red_relaxed_gpu_add(prod + qq % D, 1u);
    }
    MG_T(th3);
#ifdef MG_DBG_FINISH
    if (c.probe && w == 0 && lane == 0) {
      if (q == 0) g_mega_rows[0][0] = mg_globaltimer();
      if (q + 1u == qtotal) g_mega_rows[1][0] = mg_globaltimer();
    }
#endif
    if (haveP && fb) prefetch(fP, yP, stage + ((q + 2u) % 3u) * MROWS);
    MG_T(th4);
    MG_ACC(tprobe, 8, 0, th4 - th3); MG_ACC(tprobe, 8, 1, th1 - th0); MG_ACC(tprobe, 8, 2, th2 - th1); MG_ACC(tprobe, 8, 3, th3 - th2); MG_ACC(tprobe, 8, 4, 1);
    fC = fN; yC = yN; haveC = haveN;
    fN = fP; yN = yP; haveN = haveP;
    ++q;
  }
  MG_T(th_end);
  MG_ACC(tprobe, 8, 13, th_end - th_begin);
}

// ---- peer CTA of the pair: relay "my ring row has landed" to the leader (one lane per ring slot) -------------------------
// test_wait, not try_wait: try_wait suspends a lane for a while when its barrier is not complete, and the warp's instruction
// only retires when the slowest lane does -- most lanes watch slots that fill much later.
template <class E>
__device__ void mg_relay(const MegaK& M, const MCtx& c, const MEngSmem& s, int lane) {
  uint32_t total = 0;
  MSegIter it(M, c.team);
  MSeg g;
  while (it.next(g)) total += (uint32_t)(mg_hi(g, E::HALO, M.Hw) - mg_lo(g, E::HALO)) + 2u;
  if (lane == 0) mbar_wait(s.wbar, 0);          // my half of the weights is part of every MMA the leader issues
  __syncwarp();
  const bool active = lane < E::RING;
  const uint32_t fills = active && total > (uint32_t)lane ? (total - (uint32_t)lane + E::RING - 1) / E::RING : 0u;
  uint32_t done = 0, par = 0;
  long long t0 = clock64();
  while (__any_sync(0xffffffffu, done < fills)) {
    if (done < fills && mbar_test_wait(&s.full[active ? lane : 0], par)) {
      mbar_arrive_cluster(&s.pfull[lane], 0);
      par ^= 1u;
      ++done;
      t0 = clock64();
    }
    if (clock64() - t0 > (1ll << 33)) __trap();
  }
}

// ---- MMA issuer (leader CTA), input-row-major order with A-collector reuse ---------------------------------------------------
// Input row k of a segment feeds output rows k-2 (kernel row 2), k-1 (row 1) and k (row 0).  The three MMAs of a step share
// their A tile (collector fill / use / lastuse): it is read from shared memory once instead of three times, and the A read
// is what bounds narrow-N UMMA (tools/umma_probe.cu: 19 / 33 / 49 cycles per instruction in such triples at N = 16 / 48 / 80
// against 38 / 46 / 59).  Every output row keeps its own accumulator stage (block n in stage n % STAGES), three are open at a
// time, so the order needs STAGES >= 5.  One input row = STEPS_ROW steps; the unit / LBO pattern is that of issue_block_2cta.
template <class E>
__device__ void mg_issuer_rm(const MegaK& M, const MCtx& c, const MEngSmem& s, uint32_t tmem_cols) {
  constexpr uint32_t IDESC = umma_idesc_op(2 * MROWS, E::NPAD);
  constexpr uint32_t HI = (uint32_t)((128u >> 4)) | (1u << 14);
  constexpr uint32_t BSTEP = (E::NB * 32) >> 4, BROW = E::STEPS_ROW * BSTEP, PR16 = PLANE_ROW / 16;
  constexpr uint32_t L16 = 1u << 16, L2K = (uint32_t)(PR16 - 2) << 16;
  constexpr int G3 = E::STEPS_ROW / 3, REM = E::STEPS_ROW % 3;
  static_assert(REM == 0 || REM == 2, "unexpected instruction count per row");
  const uint32_t ring_lo = (smem_u32(s.ring) & 0x3FFFFu) >> 4;
  const uint32_t w_lo = ((smem_u32(s.w) & 0x3FFFFu) >> 4) | ((uint32_t)((E::NB * 16) >> 4) << 16);
  const uint64_t hi = (uint64_t)HI << 32;
  mbar_wait(s.wbar, 0);
  uint32_t blk0 = 0, grow = 0;       // blocks / input rows of the CTA before this segment
  MG_T(ti_begin);
  MSegIter it(M, c.team);
  MSeg g;
  while (it.next(g)) {
    const int rows = mg_hi(g, E::HALO, M.Hw) - mg_lo(g, E::HALO);
    for (int k = 0; k < rows + 2; ++k) {
      const uint32_t r = grow + (uint32_t)k, rs = r % E::RING, rpar = (r / E::RING) & 1u;
      MG_T(ti0);
      mbar_wait(&s.full[rs], rpar);
      mbar_wait(&s.pfull[rs], rpar);
      MG_T(ti1);
      const uint32_t n = blk0 + (uint32_t)k;           // block (output row) this input row opens
      const uint32_t sk = n % E::STAGES, pk = ((n / E::STAGES) & 1u) ^ 1u;
      const bool v0 = k <= rows - 1, v1 = k >= 1 && k <= rows, v2 = k >= 2;
      if (v0) mbar_wait(&s.tempty[sk], pk);
      tc_fence_after();
      MG_T(ti2);
      const uint32_t s1 = sk >= 1u ? sk - 1u : sk + E::STAGES - 1u, s2 = sk >= 2u ? sk - 2u : sk + E::STAGES - 2u;
      const uint32_t d0 = tmem_cols + sk * E::NPAD, d1 = tmem_cols + s1 * E::NPAD, d2 = tmem_cols + s2 * E::NPAD;
      uint32_t a_lo = ring_lo + rs * (E::ROWBYTES >> 4), b_lo = w_lo, acc0 = 0;
      mg_lock(s.mmalock);
      if (v0 && v1 && v2) {
        auto step3 = [&](uint32_t a, uint32_t b) {
          umma_bf16_coll<2, 1>(d2, hi | a, hi | (b + 2u * BROW), IDESC, 1u);
          umma_bf16_coll<2, 2>(d1, hi | a, hi | (b + BROW), IDESC, 1u);
          umma_bf16_coll<2, 3>(d0, hi | a, hi | b, IDESC, acc0);
          acc0 = 1u;
        };
#pragma unroll 1
        for (int g3 = 0; g3 < G3; ++g3) {
          step3(a_lo | L16, b_lo);
          step3((a_lo + 2) | L2K, b_lo + BSTEP);
          step3((a_lo + PR16 + 1) | L16, b_lo + 2 * BSTEP);
          a_lo += 2 * PR16;
          b_lo += 3 * BSTEP;
        }
        if constexpr (REM == 2) {
          step3(a_lo | L16, b_lo);
          step3((a_lo + 1) | L16, b_lo + BSTEP);
        }
      } else {
        // the first and last two input rows of a segment feed fewer than three output rows
        auto step = [&](uint32_t a, uint32_t b) {
          if (v2) umma_bf16_2cta(d2, hi | a, hi | (b + 2u * BROW), IDESC, 1u);
          if (v1) umma_bf16_2cta(d1, hi | a, hi | (b + BROW), IDESC, 1u);
          if (v0) umma_bf16_2cta(d0, hi | a, hi | b, IDESC, acc0);
          acc0 = 1u;
        };
#pragma unroll 1
        for (int g3 = 0; g3 < G3; ++g3) {
          step(a_lo | L16, b_lo);
          step((a_lo + 2) | L2K, b_lo + BSTEP);
          step((a_lo + PR16 + 1) | L16, b_lo + 2 * BSTEP);
          a_lo += 2 * PR16;
          b_lo += 3 * BSTEP;
        }
        if constexpr (REM == 2) {
          step(a_lo | L16, b_lo);
          step((a_lo + 1) | L16, b_lo + BSTEP);
        }
      }
      umma_commit_2cta(&s.empty[rs]);                  // the MMAs are done with this input row (the pipe completes in issue order)
      if (v2) umma_commit_2cta(&s.tfull[s2]);          // output row k-2 is complete
      mg_unlock(s.mmalock);
      MG_T(ti3);
      MG_ACC(c.probe, E::L, 1, ti1 - ti0); MG_ACC(c.probe, E::L, 2, ti2 - ti1); MG_ACC(c.probe, E::L, 3, ti3 - ti2); MG_ACC(c.probe && v2, E::L, 13, 1);
    }
    blk0 += (uint32_t)rows;
    grow += (uint32_t)rows + 2u;
  }
  MG_T(ti_end);
  MG_ACC(c.probe, E::L, 0, ti_end - ti_begin);
}

// ---- MMA issuer (leader CTA): block-major, one accumulator per strip row (engines with fewer than 5 accumulator stages) ------
template <class E>
__device__ void mg_issuer(const MegaK& M, const MCtx& c, const MEngSmem& s, uint32_t tmem_cols) {
  constexpr uint32_t IDESC = umma_idesc_op(2 * MROWS, E::NPAD);
  const uint32_t ring_lo = (smem_u32(s.ring) & 0x3FFFFu) >> 4;
  const uint32_t w_lo = ((smem_u32(s.w) & 0x3FFFFu) >> 4) | ((uint32_t)((E::NB * 16) >> 4) << 16);
  mbar_wait(s.wbar, 0);
  uint32_t wslot = 0, wpar = 0, stage = 0, spar = 1;
  auto wait_row = [&]() {
    mbar_wait(&s.full[wslot], wpar);
    MG_T(tw_a);
    mbar_wait(&s.pfull[wslot], wpar);
    MG_T(tw_b);
    MG_ACC(c.probe && E::L == 4, 8, 12, tw_b - tw_a);   // row 7, [12]: conv4's input-row waits spent on the peer's relay
    if (++wslot == E::RING) { wslot = 0; wpar ^= 1u; }
  };
  MG_T(ti_begin);
  MSegIter it(M, c.team);
  MSeg g;
  while (it.next(g)) {
    const int rows = mg_hi(g, E::HALO, M.Hw) - mg_lo(g, E::HALO);
    uint32_t s0 = wslot;
    MG_T(ti_a);
    wait_row();
    wait_row();
    MG_T(ti_b);
    MG_ACC(c.probe, E::L, 1, ti_b - ti_a);
    for (int b = 0; b < rows; ++b) {
      MG_T(ti0);
      wait_row();
      MG_T(ti1);
      mbar_wait(&s.tempty[stage], spar);
      tc_fence_after();
      MG_T(ti2);
      mg_lock(s.mmalock);
      issue_block_2cta<E::PT, PLANE_ROW / 16, E::NB>(tmem_cols + stage * E::NPAD, ring_lo, E::ROWBYTES >> 4, s0, E::RING, w_lo, IDESC);
      umma_commit_2cta(&s.tfull[stage]);
      umma_commit_2cta(&s.empty[s0]);
      mg_unlock(s.mmalock);
      if (b == rows - 1) {
        const uint32_t s1 = s0 + 1 == E::RING ? 0 : s0 + 1, s2 = s1 + 1 == E::RING ? 0 : s1 + 1;
        umma_commit_2cta(&s.empty[s1]);
        umma_commit_2cta(&s.empty[s2]);
      }
      MG_T(ti3);
      MG_ACC(c.probe, E::L, 1, ti1 - ti0); MG_ACC(c.probe, E::L, 2, ti2 - ti1); MG_ACC(c.probe, E::L, 3, ti3 - ti2); MG_ACC(c.probe, E::L, 13, 1);
      if (++s0 == E::RING) s0 = 0;
      if (++stage == E::STAGES) { stage = 0; spar ^= 1u; }
    }
  }
  MG_T(ti_end);
  MG_ACC(c.probe, E::L, 0, ti_end - ti_begin);
}

// one activation slot (compile-time op-code) on the 8 channels of a chunk; prm -> this chunk's column of the parameter table
template <int OP>
__device__ __forceinline__ void mg_slot8(int slot, const float* prm, float (&o)[8]) {
  if constexpr (OP != FSUAE_ACT_IDENTITY) {
    float p0[8], p1[8];
    if constexpr (((ACT_PARAM_MASK >> OP) & 1u) != 0) {
      const float4 a0 = *reinterpret_cast<const float4*>(prm + (1 + slot) * MG_MAXC), a1 = *reinterpret_cast<const float4*>(prm + (1 + slot) * MG_MAXC + 4);
      const float4 b0 = *reinterpret_cast<const float4*>(prm + (5 + slot) * MG_MAXC), b1 = *reinterpret_cast<const float4*>(prm + (5 + slot) * MG_MAXC + 4);
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

// Layers without activations or residual (conv3, conv5): bias, round, store -- NCHUNK 8-channel chunks per accumulator read-back
// (one tcgen05.ld + one wait for up to 32 columns instead of a wait per chunk: the read-back latency is all this epilogue costs).
template <int NCHUNK>
__device__ __forceinline__ void mg_ident_group(uint32_t taddr, const float* prm, unsigned char* dp, size_t plane_pitch, bool valid, int nlast) {
  uint32_t v[NCHUNK * 8];
  tmem_ld_cols<NCHUNK * 8>(taddr, v);
  tmem_ld_wait();
#pragma unroll
  for (int j = 0; j < NCHUNK; ++j) {
    const float4 b0 = *reinterpret_cast<const float4*>(prm + j * 8), b1 = *reinterpret_cast<const float4*>(prm + j * 8 + 4);
    float o[8];
    o[0] = __uint_as_float(v[j * 8 + 0]) + b0.x; o[1] = __uint_as_float(v[j * 8 + 1]) + b0.y; o[2] = __uint_as_float(v[j * 8 + 2]) + b0.z;
    o[3] = __uint_as_float(v[j * 8 + 3]) + b0.w; o[4] = __uint_as_float(v[j * 8 + 4]) + b1.x; o[5] = __uint_as_float(v[j * 8 + 5]) + b1.y;
    o[6] = __uint_as_float(v[j * 8 + 6]) + b1.z; o[7] = __uint_as_float(v[j * 8 + 7]) + b1.w;
    if (j == NCHUNK - 1) {           // padding channels stay exactly zero
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = i < nlast ? o[i] : 0.f;
    }
    if (valid)
      *reinterpret_cast<uint4*>(dp + (size_t)j * plane_pitch) =
          make_uint4(pack_op2(o[0], o[1]), pack_op2(o[2], o[3]), pack_op2(o[4], o[5]), pack_op2(o[6], o[7]));
  }
}

// ---- epilogue warpgroup `wg` of the engine's NWG --------------------------------------------------------------------------
template <class E>
__device__ void mg_epilogue(const MegaK& M, const MCtx& c, const MEngSmem& s, uint32_t tmem_cols, const float* s_lut, int wg, int warp, int lane) {
  using EPI = typename E::EPI;
  const MegaLayerP& P = M.L[E::L - 1];
  const int q4 = warp & 3;                       // TMEM lane quadrant
  const int m = q4 * 32 + lane;                  // pixel inside the strip row
  const int x = c.strip * STRIP + m;
  const bool valid = m < STRIP && x < M.Ww;
  const size_t row_pitch = (size_t)M.PW * 16;
  constexpr int D = E::OUT >= 0 ? mg_depth(E::OUT >= 0 ? E::OUT : 0) : 1;
  const size_t plane_pitch = (size_t)D * row_pitch;
  unsigned char* och = E::OUT >= 0 ? c.scratch + M.ch_off[E::OUT >= 0 ? E::OUT : 0] : nullptr;
  unsigned int* prod = mg_prod(c.flags, E::OUT >= 0 ? E::OUT : 0);
  uint32_t blk = 0, qrow = 0, qout = 0;          // blocks / ring rows / output rows before this segment
  uint32_t qtotal = 0;                           // output rows of this engine in all
  {
    MSegIter itq(M, c.team);
    MSeg gq;
    while (itq.next(gq)) qtotal += (uint32_t)(mg_hi(gq, E::HALO, M.Hw) - mg_lo(gq, E::HALO));
  }
  uint32_t cons_seen = 0;                        // every row below this has been read by all my consumers
  const bool tprobe = c.probe && wg == 0 && q4 == 0 && lane == 0;
  MG_T(te_begin);
  MSegIter it(M, c.team);
  MSeg g;
  while (it.next(g)) {
    const int f = min(2 * g.fp + c.rank, M.n_frames - 1);
    const int lo = mg_lo(g, E::HALO), rows = mg_hi(g, E::HALO, M.Hw) - lo;
    for (int b = 0; b < rows; ++b, ++blk) {
      if ((int)(blk % E::NG) != (E::SPLIT ? wg >> 1 : wg)) continue;
      const uint32_t stage = blk % E::STAGES, spar = (blk / E::STAGES) & 1u;
      const int y = lo + b;
      const uint32_t q = qout + (uint32_t)b;     // sequence number of this output row in my channel
      MG_T(te0);

      uint32_t raw[E::KIND == EPI_TAIL_SHUFFLE ? 2 : 1][6];
      if constexpr (E::KIND == EPI_TAIL_SHUFFLE) {
        if (valid) {
          const size_t fpl = (size_t)M.H * M.W;
#pragma unroll
          for (int dy = 0; dy < 2; ++dy) {
            const size_t p0 = (size_t)(2 * y + dy) * M.W + 2 * x + M.xoff;
            if (M.in_fmt == FSUAE_FMT_F32_NCHW3) {
              const float* ip = (const float*)M.frame_in + (size_t)f * 3 * fpl + p0;
#pragma unroll
              for (int ch = 0; ch < 3; ++ch) {
                const float2 t2 = __ldg(reinterpret_cast<const float2*>(ip + ch * fpl));
                raw[dy][2 * ch] = __float_as_uint(t2.x); raw[dy][2 * ch + 1] = __float_as_uint(t2.y);
              }
            } else if (M.in_fmt == FSUAE_FMT_U8_NHWC4) {
              const uint2 t2 = __ldg(reinterpret_cast<const uint2*>((const unsigned char*)M.frame_in + ((size_t)f * fpl + p0) * 4));
              raw[dy][0] = t2.x; raw[dy][1] = t2.y;
            } else {
              const unsigned char* ip = (const unsigned char*)M.frame_in + (size_t)f * 4 * fpl + p0;
#pragma unroll
              for (int ch = 0; ch < 3; ++ch) { raw[dy][2 * ch] = __ldg(ip + ch * fpl); raw[dy][2 * ch + 1] = __ldg(ip + ch * fpl + 1); }
            }
          }
        }
      } else {
        // back-pressure: slot q % D still holds row q - D until every consumer strip that reads my columns (s-1, s, s+1)
        // has loaded it.  Checked before the accumulator wait, so the poll hides behind the MMAs.
        if (!((MG_DBG_CUT >> (E::OUT >= 0 ? E::OUT : 0)) & 1) && q >= (uint32_t)D && cons_seen < q - (uint32_t)D + 1u) {
          const uint32_t need = q - (uint32_t)D + 1u;
          const long long t0 = clock64();
          for (;;) {
            uint32_t v = 0xFFFFFFFFu;
            if (lane < 3 * E::OUT_NCONS) {
              const int k = lane / 3, sn = c.strip - 1 + lane % 3;
              if (sn >= 0 && sn < M.S) v = ld_relaxed_gpu(mg_cons(c.flags, E::OUT >= 0 ? E::OUT : 0, k) + sn);
            }
            v = __reduce_min_sync(0xffffffffu, v);
            if (v >= need) { cons_seen = v; if (MG_BP_FENCE) fence_acq_rel_gpu(); break; }
            if (clock64() - t0 > (1ll << 31)) __trap();
            __nanosleep(64);
          }
        }
      }

      MG_T(te1);
      mbar_wait(&s.tfull[stage], spar);
      tc_fence_after();
      MG_T(te2);
      const uint32_t taddr = tmem_cols + ((uint32_t)(q4 * 32) << 16) + stage * E::NPAD;
      // residual = this layer's input at the same pixel = centre row of the block, still in the ring (see conv3x3_tc_kernel)
      const uint32_t kc = qrow + (uint32_t)b + 1;
      const uint32_t cslot = kc % E::RING;
      const uint8_t* sp = s.ring + cslot * E::ROWBYTES + (m + 1) * 16;
      if constexpr (EPI::kSkip) mbar_wait(&s.full[cslot], (kc / E::RING) & 1);

      constexpr bool kPlain = !EPI::kSkip && EPI::kOp0 == 0 && EPI::kOp1 == 0 && EPI::kOp2 == 0 && EPI::kOp3 == 0;
      if constexpr (E::KIND == EPI_STORE) {
        unsigned char* dp = och + (size_t)(q % (uint32_t)D) * row_pitch + (size_t)(x + BORDER) * 16;
        if constexpr (kPlain) {
          constexpr int NC = E::OUT_PLANES, NLAST = E::COUT - 8 * (NC - 1);
          static_assert(NC == 5 || NC == 9, "plain layers of the flagship: 36 or 72 channels");
          if constexpr (NC == 5 && MG_EARLY_REL) {
            // one WG drains every row of this engine and only one accumulator stage is not open: hold it for the read-back only
            uint32_t v[40];
            tmem_ld_cols<40>(taddr, v);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(&s.tempty[stage], 0);
#pragma unroll
            for (int j = 0; j < 5; ++j) {
              const float4 b0 = *reinterpret_cast<const float4*>(s.prm + j * 8), b1 = *reinterpret_cast<const float4*>(s.prm + j * 8 + 4);
              float o[8];
              o[0] = __uint_as_float(v[j * 8 + 0]) + b0.x; o[1] = __uint_as_float(v[j * 8 + 1]) + b0.y; o[2] = __uint_as_float(v[j * 8 + 2]) + b0.z;
              o[3] = __uint_as_float(v[j * 8 + 3]) + b0.w; o[4] = __uint_as_float(v[j * 8 + 4]) + b1.x; o[5] = __uint_as_float(v[j * 8 + 5]) + b1.y;
              o[6] = __uint_as_float(v[j * 8 + 6]) + b1.z; o[7] = __uint_as_float(v[j * 8 + 7]) + b1.w;
              if (j == 4) {
#pragma unroll
                for (int i = 0; i < 8; ++i) o[i] = i < NLAST ? o[i] : 0.f;
              }
              if (valid)
                *reinterpret_cast<uint4*>(dp + (size_t)j * plane_pitch) =
                    make_uint4(pack_op2(o[0], o[1]), pack_op2(o[2], o[3]), pack_op2(o[4], o[5]), pack_op2(o[6], o[7]));
            }
          } else {
          mg_ident_group<4>(taddr, s.prm, dp, plane_pitch, valid, 8);
          if constexpr (NC == 9) mg_ident_group<4>(taddr + 32, s.prm + 32, dp + 4 * plane_pitch, plane_pitch, valid, 8);
          mg_ident_group<1>(taddr + (NC - 1) * 8, s.prm + (NC - 1) * 8, dp + (size_t)(NC - 1) * plane_pitch, plane_pitch, valid, NLAST);
          }
        } else {
        // split drain: the two warpgroups of a row take 4 and 5 (of 9) chunks, in turn
        int c0 = 0, c1 = E::OUT_PLANES;
        if constexpr (E::SPLIT) {
          const int cut = E::OUT_PLANES / 2 + (int)((blk / E::NG) & 1u);
          if (wg & 1) c0 = cut; else c1 = cut;
        }
        // Two chunks per trip, their accumulators in two register sets (the next read-back is issued before this chunk's math:
        // a tcgen05.ld takes several hundred cycles while the MMAs keep the tensor memory busy); pointers advance by
        // increments.  The one-chunk loop spent 27 % of its 171 instructions on register moves and 64-bit address products.
        const float* prm = s.prm + c0 * 8;
        const uint8_t* spc = sp + c0 * PLANE_ROW;
        unsigned char* dpc = dp + (size_t)c0 * plane_pitch;
        uint32_t tc = taddr + c0 * 8;
        auto chunk = [&](const uint32_t (&v)[8], int cc) {
          const float4 b0 = *reinterpret_cast<const float4*>(prm), b1 = *reinterpret_cast<const float4*>(prm + 4);
          uint4 skc = make_uint4(0, 0, 0, 0);
          if constexpr (EPI::kSkip) {
            if (valid) skc = *reinterpret_cast<const uint4*>(spc);
          }
          float o[8];
          o[0] = __uint_as_float(v[0]) + b0.x; o[1] = __uint_as_float(v[1]) + b0.y; o[2] = __uint_as_float(v[2]) + b0.z;
          o[3] = __uint_as_float(v[3]) + b0.w; o[4] = __uint_as_float(v[4]) + b1.x; o[5] = __uint_as_float(v[5]) + b1.y;
          o[6] = __uint_as_float(v[6]) + b1.z; o[7] = __uint_as_float(v[7]) + b1.w;
          // A chain that ends in ReLU / ReLU6 gets it on the packed words behind the rounding (one HMNMX2 per two channels instead
          // of an FMNMX each): rounding is monotone and 0 and 6 are exact in both operand types, so the bits are the same.
          constexpr bool kTailRelu = EPI::kOp3 == FSUAE_ACT_RELU;
          constexpr bool kTailRelu6 = !EPI::kSkip && EPI::kOp2 == 0 && EPI::kOp3 == 0 && EPI::kOp1 == FSUAE_ACT_RELU6;
#ifndef MG_DBG_NOACT      // timing experiment (garbage results): what the pass costs without any activation math
          mg_slot8<EPI::kOp0>(0, prm, o);
          mg_slot8<kTailRelu6 ? 0 : EPI::kOp1>(1, prm, o);
#endif
          if constexpr (EPI::kSkip) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const uint32_t w = (&skc.x)[i >> 1];
              o[i] += (i & 1) ? op_hi(w) : op_lo(w);
            }
          }
#ifndef MG_DBG_NOACT
          mg_slot8<EPI::kOp2>(2, prm, o);
          mg_slot8<kTailRelu ? 0 : EPI::kOp3>(3, prm, o);
#endif
          if constexpr (E::COUT % 8 != 0) {
            if (cc == E::OUT_PLANES - 1) {       // padding channels stay exactly zero
#pragma unroll
              for (int i = E::COUT % 8; i < 8; ++i) o[i] = 0.f;
            }
          }
          uint4 pk = make_uint4(pack_op2(o[0], o[1]), pack_op2(o[2], o[3]), pack_op2(o[4], o[5]), pack_op2(o[6], o[7]));
#ifndef MG_DBG_NOACT
          if constexpr (kTailRelu || kTailRelu6) { pk.x = op2_relu<kTailRelu6>(pk.x); pk.y = op2_relu<kTailRelu6>(pk.y); pk.z = op2_relu<kTailRelu6>(pk.z); pk.w = op2_relu<kTailRelu6>(pk.w); }
#endif
          if (valid) *reinterpret_cast<uint4*>(dpc) = pk;
          prm += 8;
          spc += PLANE_ROW;
          dpc += plane_pitch;
        };
        uint32_t va[8], vb[8];
        tmem_ld_x8(tc, va);
#pragma unroll 1
        for (int cc = c0; cc < c1; cc += 2) {
          tmem_ld_wait();
          if (cc + 1 < c1) tmem_ld_x8(tc + 8, vb);
          chunk(va, cc);
          if (cc + 1 < c1) {
            tmem_ld_wait();
            if (cc + 2 < c1) tmem_ld_x8(tc + 16, va);
            chunk(vb, cc + 1);
          }
          tc += 16;
        }
        }
        // Rows are published in groups of MG_PUB_ROWS: the last warp to finish the group (4 warps per row, whichever
        // warpgroups they belong to) bumps the rows' counters behind ONE fence.  Its gpu scope is cumulative over the
        // CTA-scope acquire / release on the shared-memory counter, so it covers the other warps' stores as well, and only
        // that warp waits (1000-1700 cycles) for the SM's stores to reach L2.
        MG_T(tf0);
        __syncwarp();
        if (lane == 0) {
          // group g = rows [g * MG_PUB_ROWS - ph, (g + 1) * MG_PUB_ROWS - ph): the first ph rows of group 0 do not exist (their
          // arrivals are preloaded into the counter), the channel's last group may be short
          const uint32_t grp = (q + s.ph) / (uint32_t)MG_PUB_ROWS;
          const int q_lo_s = (int)(grp * (uint32_t)MG_PUB_ROWS) - (int)s.ph;
          const uint32_t q_lo = (uint32_t)max(q_lo_s, 0), q_hi = min((uint32_t)(q_lo_s + MG_PUB_ROWS), qtotal);
          const uint32_t nrows = (uint32_t)((int)q_hi - q_lo_s);
          const uint32_t before = atom_acq_rel_cta_smem_add(s.rowcnt + grp % (uint32_t)MG_NR, 1u);
          MG_T(tf1);
          if ((before + 1u) % ((uint32_t)E::WPR * (uint32_t)MG_PUB_ROWS) == ((uint32_t)E::WPR * nrows) % ((uint32_t)E::WPR * (uint32_t)MG_PUB_ROWS)) {
            fence_acq_rel_gpu();
            for (uint32_t qq = q_lo; qq < q_hi; ++qq) red_relaxed_gpu_add(prod + qq % (uint32_t)D, 1u);
          }
          MG_T(tf2);
          MG_ACC(tprobe, E::L, 14, tf1 - tf0); MG_ACC(tprobe, E::L, 15, tf2 - tf1);
          if constexpr (EPI::kSkip) {            // residual values consumed: release the ring rows
            mbar_arrive(&s.empty[cslot]);
            if (b == 0) mbar_arrive(&s.empty[(kc - 1) % E::RING]);
            if (b == rows - 1) mbar_arrive(&s.empty[(kc + 1) % E::RING]);
          }
        }
      } else {
        // PixelShuffle(2) + input residual + ReLU (+ gamma, uint8 pack), straight to the output frame
        uint32_t v[16];
        tmem_ld_x16(taddr, v);
        tmem_ld_wait();
        float o[12];
#pragma unroll
        for (int ch = 0; ch < 12; ++ch) o[ch] = EPI::post(P, ch, EPI::pre(P, ch, __uint_as_float(v[ch]) + P.bias[ch]));
        if (valid) {
          const size_t fpl = (size_t)M.H * M.W;
          float idv[2][3][2];
#pragma unroll
          for (int dy = 0; dy < 2; ++dy)
#pragma unroll
            for (int cc = 0; cc < 3; ++cc) {
              if (M.in_fmt == FSUAE_FMT_F32_NCHW3) {
                idv[dy][cc][0] = __uint_as_float(raw[dy][2 * cc]); idv[dy][cc][1] = __uint_as_float(raw[dy][2 * cc + 1]);
              } else if (M.in_fmt == FSUAE_FMT_U8_NHWC4) {
                idv[dy][cc][0] = s_lut[(raw[dy][0] >> (8 * cc)) & 0xFF]; idv[dy][cc][1] = s_lut[(raw[dy][1] >> (8 * cc)) & 0xFF];
              } else {
                idv[dy][cc][0] = s_lut[raw[dy][2 * cc] & 0xFF]; idv[dy][cc][1] = s_lut[raw[dy][2 * cc + 1] & 0xFF];
              }
            }
#pragma unroll
          for (int dy = 0; dy < 2; ++dy) {
            const size_t p0 = (size_t)(2 * y + dy) * M.W + 2 * x + M.xoff;
            float res[3][2];
#pragma unroll
            for (int cc = 0; cc < 3; ++cc)
#pragma unroll
              for (int dx = 0; dx < 2; ++dx) res[cc][dx] = fmaxf(o[cc * 4 + dy * 2 + dx] + idv[dy][cc][dx], 0.f);
            if (M.out_fmt == FSUAE_FMT_F32_NCHW3) {
              float* op = (float*)M.frame_out + (size_t)f * 3 * fpl + p0;
#pragma unroll
              for (int cc = 0; cc < 3; ++cc) *reinterpret_cast<float2*>(op + cc * fpl) = make_float2(res[cc][0], res[cc][1]);
            } else {
              uint32_t px[2];
#pragma unroll
              for (int dx = 0; dx < 2; ++dx)
                px[dx] = (uint32_t)to_u8_fast(res[0][dx], M.gamma_out) | ((uint32_t)to_u8_fast(res[1][dx], M.gamma_out) << 8) |
                         ((uint32_t)to_u8_fast(res[2][dx], M.gamma_out) << 16) | 0xFF000000u;
              *reinterpret_cast<uint2*>((unsigned char*)M.frame_out + ((size_t)f * fpl + p0) * 4) = make_uint2(px[0], px[1]);
            }
          }
        }
      }
      constexpr bool kReleased = E::KIND == EPI_STORE && kPlain && E::OUT_PLANES == 5 && MG_EARLY_REL;
      if constexpr (!kReleased) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(&s.tempty[stage], 0);
      }
#ifdef MG_DBG_FINISH
      if (c.probe && q4 == 0 && lane == 0) {
        if (q == 0) g_mega_rows[0][E::L] = mg_globaltimer();
        if (q + (uint32_t)E::NG >= qtotal) g_mega_rows[1][E::L] = mg_globaltimer();       // this warpgroup's last row
      }
#endif
      MG_T(te3);
      MG_ACC(tprobe, E::L, 7, te1 - te0); MG_ACC(tprobe, E::L, 8, te2 - te1); MG_ACC(tprobe, E::L, 9, te3 - te2); MG_ACC(tprobe, E::L, 11, 1);
    }
    qrow += (uint32_t)rows + 2u;
    qout += (uint32_t)rows;
  }
  MG_T(te_end);
  MG_ACC(tprobe, E::L, 10, te_end - te_begin);
}

// ---- the engines of the flagship preset (model_pix_shuffle.py:306-311) ---------------------------------------------------
#define MG_A(x) FSUAE_ACT_##x
// <layer, input planes, N, Cout, kind, epilogue, ring rows, accumulator stages, epilogue warpgroups, row-major MMA order>
// Accumulator stages: the row-major order keeps three open; on top of that one per row being drained at a time -- a plain
// layer drains a row in ~1300 cycles (one stage), an activation chain takes ~4500 (one stage per warpgroup).
// Who shares an SM is decided by three budgets -- tensor time (MMA instructions per row x their cost), issue slots /
// SFU time of the activation chains, and TMEM columns (stages x N <= 512 per SM):
//   conv5 (42 MMAs, no activation: ONE warpgroup)     + conv2 (24 MMAs, the longest chain: THREE warpgroups)
//   conv4 (42 MMAs at N = 80, long chain: all four)   + the head
//   conv3 (24 MMAs at N = 80, no activation)          + conv7 (24 MMAs at N = 16, PixelShuffle tail)
//   conv6 (45 MMAs, Mish)                             + conv1 (9 MMAs, SinLU)
using MgConv1 = MEng<1, 2, 48, 36, EPI_STORE, Epi<MG_A(SINLU), MG_A(RELU6), 0, 0, false>, 8, 4, 2, false>;
using MgConv2 = MEng<2, 5, 48, 36, EPI_STORE, Epi<MG_A(TELU), 0, MG_A(SINLU), MG_A(BIASED_PRELU), true>, 6, 6, 3>;
// Which stage hosts the head (1 = B, on the service warps conv4 has no second engine for; 2 = C, on two warps of a warpgroup
// conv3 gives up).  Stage B is the slowest stage on its own (tools/mega_finish.py, free-running engines: A 2190, B 2640,
// C 1820, D 2460 cycles per row) -- conv4's activation chain keeps its 16 epilogue warps' issue slots and the SFU busy, and
// the head's ~1000 warp instructions per row come on top -- while stage C has the most slack: conv3 is a plain layer that one
// warpgroup drains in ~1500 cycles per row.
#ifndef MG_HEAD_STAGE
#define MG_HEAD_STAGE 1
#endif
using MgConv3 = MEng<3, 5, 80, 72, EPI_STORE, Epi<0, 0, 0, 0, false>, 8, 5, MG_HEAD_STAGE == 2 ? 1 : 2>;
#ifndef MG_SPLIT4
#define MG_SPLIT4 0      // measured: 38.2 vs 37.4 us/frame -- conv4 is bound by its 16 warps' issue slots and the SFU, not by how long a stage is held
#endif
using MgConv4 = MEng<4, 9, 80, 72, EPI_STORE, Epi<MG_A(MISH), MG_A(BIASED_PRELU), MG_A(TANH), MG_A(RELU), true>, 9, 6, 4, true, MG_SPLIT4 != 0>;
using MgConv5 = MEng<5, 9, 48, 36, EPI_STORE, Epi<0, 0, 0, 0, false>, 5, 4, 1>;
using MgConv6 = MEng<6, 10, 48, 36, EPI_STORE, Epi<MG_A(MISH), MG_A(RELU6), 0, 0, false>, 6, 6, 2>;
using MgConv7 = MEng<7, 5, 16, 12, EPI_TAIL_SHUFFLE, Epi<MG_A(BIASED_PRELU), 0, 0, 0, false>, 6, 6, 2>;
#undef MG_A
struct MgNone { static constexpr int SMEM = 0, TCOLS = 0, NBARS = 0, NWG = 0, L = 1, WPR = 0; };

template <class E0, class E1, int IDX_>
struct MStage {
  static constexpr int IDX = IDX_;
  static constexpr bool kTwo = E1::NWG > 0;
  static constexpr bool kHead = IDX_ == MG_HEAD_STAGE;       // this stage hosts the head
  static constexpr int SMEM = E0::SMEM + E1::SMEM;
  static_assert(E0::NWG + E1::NWG == (kHead && kTwo ? 3 : 4), "16 epilogue warps per CTA (a stage with two engines and the head: 12 + the head's two)");
  static_assert(E0::TCOLS + E1::TCOLS <= 512, "TMEM columns");
  static_assert((E0::SMEM % 128) == 0, "operand alignment of the second engine");
};
using MgStageA = MStage<MgConv5, MgConv2, 0>;
using MgStageB = MStage<MgConv4, MgNone, 1>;
using MgStageC = MStage<MgConv3, MgConv7, 2>;
using MgStageD = MStage<MgConv6, MgConv1, 3>;
constexpr int mg_max(int a, int b) { return a > b ? a : b; }
constexpr int MG_BAR_BYTES = 1152;       // 120 barriers, the TMEM slot, 2 x MG_NR row counters
constexpr int MG_SMEM = mg_max(mg_max(MgStageA::SMEM, MgStageB::SMEM), mg_max(MgStageC::SMEM, MgStageD::SMEM)) + MG_BAR_BYTES;
static_assert(MG_SMEM <= SMEM_LIMIT, "fused pass does not fit in shared memory");

template <class E>
__device__ __forceinline__ void mg_run_service(const MegaK& M, const MCtx& c, const MEngSmem& s, uint32_t tmem_cols, int role, int lane) {
  if (role == 0) {           // producer warp
    mg_producer<E>(M, c, s, lane);
  } else if (c.rank != 0) {  // peer CTA: relay
    mg_relay<E>(M, c, s, lane);
  } else if (elect_one()) {  // leader CTA: MMA issue
    if constexpr (E::RM) mg_issuer_rm<E>(M, c, s, tmem_cols);
    else mg_issuer<E>(M, c, s, tmem_cols);
  }
}

template <class ST, class E0, class E1>
__device__ __forceinline__ void mg_run_stage(const MegaK& M, MCtx& c, uint8_t* smem, float* s_lut, uint4* s_raw, int warp, int lane) {
#ifdef MG_DBG_FINISH
  const long long t_start = clock64();
  const unsigned int dbg_slot = g_mega_launch & 3u;            // bumped by the host-side launcher between launches
  if (threadIdx.x == 0) g_mega_gt[dbg_slot][0][blockIdx.x] = mg_globaltimer();
#endif
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + MG_SMEM - MG_BAR_BYTES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 120);
  uint32_t* rowcnt = reinterpret_cast<uint32_t*>(bars + 128);
  const MEngSmem s0c = mg_carve<E0>(smem, bars, rowcnt);
  MEngSmem s1 = s0c;
  MEngSmem s0m = s0c;
  if constexpr (ST::kTwo) s1 = mg_carve<E1>(smem + E0::SMEM, bars + E0::NBARS, rowcnt + MG_NR);
  if constexpr (ST::kTwo && ((MG_MMA_LOCK >> ST::IDX) & 1) != 0) {
    uint32_t* lock = reinterpret_cast<uint32_t*>(bars + 140);
    if (threadIdx.x == 0) *lock = 0u;
    s0m.mmalock = lock;
    s1.mmalock = lock;
  }
  s0m.ph = mg_pub_phase(M, c.team, E0::L);
  if constexpr (ST::kTwo) s1.ph = mg_pub_phase(M, c.team, E1::L);
  const MEngSmem s0 = s0m;
  if (threadIdx.x < 2 * MG_NR)      // slot 0 starts with the arrivals of group 0's rows that do not exist
    rowcnt[threadIdx.x] = threadIdx.x == 0 ? (uint32_t)E0::WPR * s0.ph : (ST::kTwo && threadIdx.x == MG_NR ? (uint32_t)E1::WPR * s1.ph : 0u);
  for (int i = threadIdx.x; i < 9 * MG_MAXC; i += MG_THREADS) {     // MegaLayerP starts with bias[MG_MAXC], p0[4][MG_MAXC], p1[4][MG_MAXC]
    const_cast<float*>(s0.prm)[i] = reinterpret_cast<const float*>(&M.L[E0::L - 1])[i];
    if constexpr (ST::kTwo) const_cast<float*>(s1.prm)[i] = reinterpret_cast<const float*>(&M.L[E1::L - 1])[i];
  }
  static_assert(E0::NBARS + E1::NBARS <= 120, "barrier block");
  if (threadIdx.x == 0) {
    mg_init_bars<E0>(s0);
    if constexpr (ST::kTwo) mg_init_bars<E1>(s1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc_2cta(tmem_slot, 512);
  if (threadIdx.x >= 64 && threadIdx.x < 64 + 32) {   // zero the pads behind the rings
    reinterpret_cast<uint32_t*>(s0.ring + E0::RING * E0::ROWBYTES)[threadIdx.x - 64] = 0u;
    if constexpr (ST::kTwo) reinterpret_cast<uint32_t*>(s1.ring + E1::RING * E1::ROWBYTES)[threadIdx.x - 64] = 0u;
  }
  for (int i = threadIdx.x; i < 256; i += MG_THREADS) {
    const float t = (float)i * (1.0f / 255.0f);
    s_lut[i] = M.gamma_in ? powf(t, 2.2f) : t;
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 2) {
    mg_run_service<E0>(M, c, s0, tmem_base, warp, lane);
  } else if (warp < 4) {
    if constexpr (ST::kTwo) mg_run_service<E1>(M, c, s1, tmem_base + E0::TCOLS, warp - 2, lane);
    else if constexpr (ST::kHead) mg_head_worker(M, c, s_lut, s_raw, warp - 2, lane);      // no second engine: the head takes its service warps
  } else {
    const int wg = (warp - 4) >> 2;
    if (wg < E0::NWG) {
      mg_epilogue<E0>(M, c, s0, tmem_base, s_lut, wg, warp, lane);
    } else if constexpr (ST::kTwo) {
      if (wg < E0::NWG + E1::NWG) mg_epilogue<E1>(M, c, s1, tmem_base + E0::TCOLS, s_lut, wg - E0::NWG, warp, lane);
      else if constexpr (ST::kHead) {           // two warps of the warpgroup the engines leave free
        const int hw = warp - 4 - 4 * (E0::NWG + E1::NWG);
        if (hw < 2) mg_head_worker(M, c, s_lut, s_raw, hw, lane);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
#ifdef MG_DBG_FINISH
  if (threadIdx.x == 0) { g_mega_finish[blockIdx.x] = clock64() - t_start; g_mega_gt[dbg_slot][1][blockIdx.x] = mg_globaltimer(); }
#endif
  cluster_sync_all();
  if (warp == 1) tmem_dealloc_2cta(tmem_base, 512);
}

__global__ void __launch_bounds__(MG_THREADS, 1) fused_pass_kernel(const __grid_constant__ MegaK M) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ float s_lut[256];
  __shared__ uint4 s_raw[3 * MROWS];             // head: raw framebuffer pixels of three strip rows (cp.async staging)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int pair = blockIdx.x >> 1, group = pair >> 2, stage = pair & 3;
  MCtx c;
  c.team = group / M.S;
  c.strip = group - c.team * M.S;
  c.rank = (int)cluster_ctarank();
  c.f = 0;
  c.probe = c.team == 0 && c.strip == (M.S > 1 ? 1 : 0) && c.rank == 0;
  if (c.team >= M.teams) return;                 // both CTAs of a pair leave together
  c.scratch = M.scratch + ((size_t)c.team * 2 + c.rank) * M.rank_stride;
  c.flags = M.flags + ((size_t)c.team * 2 + c.rank) * MG_FLAG_WORDS;
  switch (stage) {
    case 0: mg_run_stage<MgStageA, MgConv5, MgConv2>(M, c, smem, s_lut, s_raw, warp, lane); break;
    case 1: mg_run_stage<MgStageB, MgConv4, MgNone>(M, c, smem, s_lut, s_raw, warp, lane); break;
    case 2: mg_run_stage<MgStageC, MgConv3, MgConv7>(M, c, smem, s_lut, s_raw, warp, lane); break;
    default: mg_run_stage<MgStageD, MgConv6, MgConv1>(M, c, smem, s_lut, s_raw, warp, lane); break;
  }
}


}  // namespace

#if defined(FSUAE_EPI_TIMING) && !defined(FSUAE_OPERAND_FP16)
extern "C" __attribute__((visibility("default"))) int fsuae_debug_mega_timing(unsigned long long* out, int reset) {   // [7][16] counters
  cudaDeviceSynchronize();
  if (cudaMemcpyFromSymbol(out, g_mega_t, sizeof(g_mega_t)) != cudaSuccess) return -1;
  if (reset) { static unsigned long long z[8 * 16]; cudaMemcpyToSymbol(g_mega_t, z, sizeof(z)); }
  return 0;
}
#endif

#if defined(MG_DBG_FINISH) && !defined(FSUAE_OPERAND_FP16)
extern "C" __attribute__((visibility("default"))) int fsuae_debug_mega_finish(long long* out) {   // [160] cycles per CTA of the last pass
  cudaDeviceSynchronize();
  return cudaMemcpyFromSymbol(out, g_mega_finish, sizeof(g_mega_finish)) == cudaSuccess ? 0 : -1;
}
extern "C" __attribute__((visibility("default"))) int fsuae_debug_mega_globaltimer(unsigned long long* out, unsigned int* launches) {   // [4][2][160] ns
  cudaDeviceSynchronize();
  if (cudaMemcpyFromSymbol(launches, g_mega_launch, sizeof(unsigned int)) != cudaSuccess) return -1;
  return cudaMemcpyFromSymbol(out, g_mega_gt, sizeof(g_mega_gt)) == cudaSuccess ? 0 : -1;
}
extern "C" __attribute__((visibility("default"))) int fsuae_debug_mega_rows(unsigned long long* out) {   // [2][8] ns: first / last row of head, conv1..7
  cudaDeviceSynchronize();
  return cudaMemcpyFromSymbol(out, g_mega_rows, sizeof(g_mega_rows)) == cudaSuccess ? 0 : -1;
}
__global__ void mg_dbg_next_launch() { g_mega_launch++; }
#endif

int TC_FN(mega_prepare)() {
  return (int)cudaFuncSetAttribute((const void*)fused_pass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MG_SMEM);
}

int TC_FN(mega_launch)(const MegaK& k, int grid, cudaStream_t st) {
  cudaLaunchConfig_t cfg = cudaLaunchConfig_t{};
  cudaLaunchAttribute attr[1];
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(MG_THREADS);
  cfg.dynamicSmemBytes = MG_SMEM;
  cfg.stream = st;
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
#if defined(MG_DBG_FINISH) && !defined(FSUAE_OPERAND_FP16)
  const cudaError_t ce = cudaLaunchKernelEx(&cfg, fused_pass_kernel, k);
  mg_dbg_next_launch<<<1, 1, 0, st>>>();
  return (int)ce;
#else
  return (int)cudaLaunchKernelEx(&cfg, fused_pass_kernel, k);
#endif
}

}  // namespace fsuae
