// K-streamed tile kernel for WIDE layers of the bf16 build (included by bf16_tc.cu inside its anonymous namespace).
//
// The layer kernel keeps the whole layer's weights and a ring of full-depth input rows in shared memory; that stops
// fitting around 100 input or output channels (conv3 heavyweight: 192 -> 256 channels = 884 KB of weights, 48 KB per
// input row; model_conv3.py:36-52).  Here neither is resident.  A CTA pair (cta_group::2) owns an 8-row x 126-column
// x NT-channel output tile -- rank r the rows y0+4r .. y0+4r+3 -- and accumulates it in TMEM (4 rows x NT fp32 columns)
// while the K dimension streams through a shared-memory stage ring, 16 input channels per stage:
//   A: 6 input rows x 2 planes x 2 KB  (rows y-1 .. y+4 of this CTA; one TMA bulk copy per plane row)
//   B: [tap 9][k half 2][NT/2 rows][8] bf16 = this CTA's half of the weight rows of that K slice (one TMA bulk copy)
// 36 MMAs per stage (4 rows x 9 taps, M = 256 across the pair, N = NT, K = 16); a tap is a descriptor start-address
// shift exactly as in the layer kernel.  Weight re-use: one stage of B feeds 4 rows x 2 CTAs; L2 -> shared traffic is
// 42 KB per 2304 MMA cycles at NT = 128 (18 B/clk/SM, ~45 % of the L2 throughput cap).
// Per-row accumulator barriers: in a tile's last K slice the MMAs go row by row and commit each row, so the epilogue
// of row 0 (warpgroup 0) starts while rows 1-3 still compute, and the next tile's first slice re-uses row 0 as soon
// as it has been drained.
// Work unit = (frame, strip, 8-row block, output-channel group), group innermost so the second pass over the same
// input comes from L2.  Units are cut into equal contiguous ranges over the pairs.
#pragma once

struct WideK {
  int Hw, Ww, PW, S, n_frames;
  int P0, pin, kchunks;          // planes of the first source, total input planes, K slices of 16 channels
  // Kernels larger than 3x3 (model_pix_shuffle.py:108-115 `layer{i}_kernel_size`; residual_feature_block.py:6): a k x k
  // convolution is the sum of nwin 3x3 convolutions of the same input shifted by (win_dy, win_dx) pixels -- 4 windows for
  // 5x5, 9 for 7x7 -- i.e. one 3x3 convolution over nwin * pin_real "virtual" planes whose weights hold each real tap once
  // and zeros elsewhere.  A shifted window is only a different TMA source address: the planes' baked-in zero border of
  // BORDER = 3 pixels covers the widest reach (7x7: 2 + 1).
  int pin_real, nwin;            // planes of the real input (pin = nwin * pin_real); 3x3: nwin = 1
  signed char win_dy[9], win_dx[9];
  int ngroups, cout, cpad;       // output-channel groups of NT, channels of the layer, stride of the dparams rows
  int rowblocks, n_units;        // 8-row blocks per strip; n_frames * S * rowblocks * ngroups
  int has_skip;
  int softmax_slot, softmax_log;   // channel softmax slot of the run-time chain (-1: none); needs ngroups == 1
  unsigned long long fs0, fs1, fs_skip, fs_dst;
  const unsigned char* src0;
  const unsigned char* src1;
  const unsigned char* skip;
  unsigned char* dst;
  const unsigned char* wpack;    // [group][slice][rank][tap][half][NT/2][8] bf16
  const float* dparams;          // {bias, p0[4], p1[4]} x cpad
  int op[4];
  const void* frame_in;          // EPI_TAIL_SHUFFLE: the network input is added back behind PixelShuffle(2)
  void* frame_out;
  int in_fmt, out_fmt, H, W, xoff, gamma_in, gamma_out;
};

template <int NT>
struct WideCfg {
  static constexpr int ROWS = 4;                                  // output rows per CTA and tile
  static constexpr int A_BYTES = (ROWS + 2) * 2 * PLANE_ROW;      // 24 KB
  static constexpr int B_BYTES = 9 * 2 * (NT / 2) * 16;           // this CTA's half of one K slice
  static constexpr int STAGE = A_BYTES + B_BYTES;
  static constexpr int NSTAGE_FIT = (SMEM_LIMIT - 1024) / STAGE;
  static constexpr int NSTAGE = NSTAGE_FIT > 6 ? 6 : NSTAGE_FIT;
  static constexpr int BAR_OFF = NSTAGE * STAGE + 64;
  static constexpr int SMEM = BAR_OFF + 512;
  static_assert(NT % 16 == 0 && ROWS * NT <= 512, "accumulators of one tile exceed TMEM");
  static_assert(NSTAGE >= 3, "stage ring too short");
};

struct WideUnit { int f, s, y0, ng; };
__device__ __forceinline__ WideUnit wide_unit(const WideK& P, int u, int rank) {
  WideUnit w;
  w.ng = u % P.ngroups; u /= P.ngroups;
  const int rb = u % P.rowblocks; u /= P.rowblocks;
  w.s = u % P.S;
  w.f = u / P.S;
  w.y0 = rb * 8 + rank * 4;
  return w;
}

template <int NT, int KIND, class EPI>
__global__ void __launch_bounds__(NTHREADS, 1) conv3x3_tc_wide_kernel(const __grid_constant__ WideK P) {
  using C = WideCfg<NT>;
  const uint32_t rank = cluster_ctarank();
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::BAR_OFF);
  uint64_t* full = bars;                        // [NSTAGE] TMA -> MMA
  uint64_t* empty = full + C::NSTAGE;           // [NSTAGE] MMA (multicast commit) -> TMA of both CTAs
  uint64_t* pfull = empty + C::NSTAGE;          // [NSTAGE] leader only: the peer's stage has landed
  uint64_t* tfull = pfull + C::NSTAGE;          // [ROWS]   MMA -> epilogue warpgroup of that row (both CTAs)
  uint64_t* tempty = tfull + C::ROWS;           // [ROWS]   leader only: both CTAs' warpgroups have drained the row
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + C::ROWS);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < C::NSTAGE; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); mbar_init(&pfull[i], 1); }
    for (int i = 0; i < C::ROWS; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 8); }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc_2cta(tmem_slot, 512);
  __shared__ float s_lut[KIND == EPI_TAIL_SHUFFLE ? 256 : 1];
  if constexpr (KIND == EPI_TAIL_SHUFFLE) {
    for (int i = threadIdx.x; i < 256; i += NTHREADS) {
      const float t = (float)i * (1.0f / 255.0f);
      s_lut[i] = P.gamma_in ? powf(t, 2.2f) : t;
    }
  }
  if (threadIdx.x >= 64 && threadIdx.x < 64 + 16)      // pad behind the last stage (the dx = 2 tap of the last row reads 32 B past it)
    reinterpret_cast<uint32_t*>(smem + C::NSTAGE * C::STAGE)[threadIdx.x - 64] = 0u;
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  const size_t row_pitch = (size_t)P.PW * 16;
  const size_t plane_pitch = (size_t)(P.Hw + 2 * BORDER) * row_pitch;
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int u_begin = (int)((long long)P.n_units * pair / npairs);
  const int u_end = (int)((long long)P.n_units * (pair + 1) / npairs);

  if (warp == 0) {
    // ======================= TMA producer =======================
    if (elect_one()) {
      asm volatile("griddepcontrol.wait;" ::: "memory");
      uint32_t stage = 0, par = 1;
      for (int u = u_begin; u < u_end; ++u) {
        const WideUnit w = wide_unit(P, u, (int)rank);
        const size_t col = (size_t)(w.s * STRIP - 1 + BORDER) * 16;
        const unsigned char* wsrc = P.wpack + ((size_t)w.ng * P.kchunks * 2 + rank) * C::B_BYTES;
        for (int kc = 0; kc < P.kchunks; ++kc) {
          mbar_wait(&empty[stage], par);
          uint8_t* sa = smem + stage * C::STAGE;
          mbar_arrive_expect_tx(&full[stage], C::STAGE);
          tma_load_1d(sa + C::A_BYTES, wsrc + (size_t)kc * 2 * C::B_BYTES, C::B_BYTES, &full[stage]);
          const unsigned char* gp[2];
          int wdy[2];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int pv = min(2 * kc + h, P.pin - 1);     // odd plane count: the last slice re-reads a landed plane against zero weights
            const int win = pv / P.pin_real, p = pv - win * P.pin_real;
            wdy[h] = P.win_dy[win];
            gp[h] = (p < P.P0 ? P.src0 + (size_t)w.f * P.fs0 + (size_t)p * plane_pitch
                              : P.src1 + (size_t)w.f * P.fs1 + (size_t)(p - P.P0) * plane_pitch) + col + (ptrdiff_t)P.win_dx[win] * 16;
          }
#pragma unroll
          for (int k = 0; k < C::ROWS + 2; ++k) {
            // rows below the frame (results discarded) stay inside the buffer
            const int py0 = max(0, min(w.y0 + k - 1 + BORDER + wdy[0], P.Hw + 2 * BORDER - 1));
            const int py1 = max(0, min(w.y0 + k - 1 + BORDER + wdy[1], P.Hw + 2 * BORDER - 1));
            tma_load_1d(sa + (k * 2 + 0) * PLANE_ROW, gp[0] + (size_t)py0 * row_pitch, PLANE_ROW, &full[stage]);
            tma_load_1d(sa + (k * 2 + 1) * PLANE_ROW, gp[1] + (size_t)py1 * row_pitch, PLANE_ROW, &full[stage]);
          }
          if (++stage == C::NSTAGE) { stage = 0; par ^= 1; }
        }
      }
    }
  } else if (warp == 1 && rank != 0) {
    // ======================= peer: relay "my stage has landed" to the leader =======================
    if (elect_one()) {
      uint32_t stage = 0, par = 0;
      for (int u = u_begin; u < u_end; ++u)
        for (int kc = 0; kc < P.kchunks; ++kc) {
          mbar_wait(&full[stage], par);
          mbar_arrive_cluster(&pfull[stage], 0);
          if (++stage == C::NSTAGE) { stage = 0; par ^= 1; }
        }
    }
  } else if (warp == 1) {
    // ======================= MMA issuer (leader) =======================
    if (elect_one()) {
      constexpr uint32_t IDESC = umma_idesc_op(2 * MROWS, NT);
      constexpr uint32_t HI = (uint32_t)((128u >> 4)) | (1u << 14);                  // SBO = 128 B, descriptor version 1
      constexpr uint32_t A_LBO = (uint32_t)(PLANE_ROW >> 4) << 16;                   // second K half = the other plane of the row
      constexpr uint32_t B_LBO = (uint32_t)(((NT / 2) * 16) >> 4) << 16;
      constexpr uint32_t B_TAP = ((NT / 2) * 32) >> 4;
      const uint32_t base_lo = (smem_u32(smem) & 0x3FFFFu) >> 4;
      uint32_t stage = 0, par = 0, tpar = 1;
      for (int u = u_begin; u < u_end; ++u) {
        for (int kc = 0; kc < P.kchunks; ++kc) {
          mbar_wait(&full[stage], par);
          mbar_wait(&pfull[stage], par);
          tc_fence_after();
          const uint32_t a0 = base_lo + stage * (C::STAGE >> 4);
          const uint32_t b0 = a0 + (C::A_BYTES >> 4);
          const bool first = kc == 0, last = kc == P.kchunks - 1;
          // Input-row-major order: input row i of the stage (image row y0 - 1 + i) feeds output rows r = i - dy for the three
          // kernel rows dy; the MMAs of one (i, dx) share the A tile -- collector fill / use / lastuse, so it is read from
          // shared memory once.  Output row r is first touched at (i = r, dx = 0, dy = 0) and complete after i = r + 2.
#pragma unroll
          for (int i = 0; i < C::ROWS + 2; ++i) {
            if (first && i < C::ROWS) { mbar_wait(&tempty[i], tpar); tc_fence_after(); }     // row i's accumulator opens here
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
              const uint64_t a_desc = ((uint64_t)HI << 32) | ((a0 + (uint32_t)(i * 2 * (PLANE_ROW >> 4) + dx)) | A_LBO);
              const int dy_lo = i - (C::ROWS - 1) > 0 ? i - (C::ROWS - 1) : 0, dy_hi = i < 2 ? i : 2;      // kernel rows with 0 <= i - dy < ROWS
#pragma unroll
              for (int dy = 0; dy < 3; ++dy) {
                if (dy < dy_lo || dy > dy_hi) continue;
                const int r = i - dy;
                const uint64_t b_desc = ((uint64_t)HI << 32) | ((b0 + (uint32_t)(dy * 3 + dx) * B_TAP) | B_LBO);
                const uint32_t acc = (first && dy == 0 && dx == 0) ? 0u : 1u;
                const uint32_t d_tmem = tmem_base + r * NT;
                if (dy_lo == dy_hi) umma_bf16_2cta(d_tmem, a_desc, b_desc, IDESC, acc);
                else if (dy == dy_lo) umma_bf16_coll<2, 1>(d_tmem, a_desc, b_desc, IDESC, acc);
                else if (dy == dy_hi) umma_bf16_coll<2, 3>(d_tmem, a_desc, b_desc, IDESC, acc);
                else umma_bf16_coll<2, 2>(d_tmem, a_desc, b_desc, IDESC, acc);
              }
            }
            if (last && i >= 2) umma_commit_2cta(&tfull[i - 2]);
          }
          umma_commit_2cta(&empty[stage]);
          if (++stage == C::NSTAGE) { stage = 0; par ^= 1; }
        }
        tpar ^= 1;
      }
    }
  } else {
    // ======================= epilogue: warpgroup g owns output row g of the tile =======================
    const int q = warp & 3;
    const int m = q * 32 + lane;
    const int g = (warp - 2) >> 2;
    asm volatile("griddepcontrol.wait;" ::: "memory");
    uint32_t upar = 0;
    for (int u = u_begin; u < u_end; ++u, upar ^= 1) {
      const WideUnit w = wide_unit(P, u, (int)rank);
      const int x = w.s * STRIP + m, y = w.y0 + g;
      const bool valid = m < STRIP && x < P.Ww && y < P.Hw;
      const size_t pix = (size_t)(y + BORDER) * row_pitch + (size_t)(x + BORDER) * 16;
      mbar_wait(&tfull[g], upar);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + g * NT;
      const int ch0 = w.ng * NT;
      if constexpr (KIND == EPI_STORE) {
        const int plane0 = ch0 >> 3;
        const int nplanes = min(NT / 8, ((P.cout + 7) >> 3) - plane0);
        unsigned char* dp = P.dst + (size_t)w.f * P.fs_dst + pix + (size_t)plane0 * plane_pitch;
        const bool has_skip = EPI::kRuntime ? P.has_skip != 0 : EPI::kSkip;
        const unsigned char* sp = has_skip ? P.skip + (size_t)w.f * P.fs_skip + pix + (size_t)plane0 * plane_pitch : nullptr;
        const uint32_t ops_packed = (uint32_t)P.op[0] | ((uint32_t)P.op[1] << 8) | ((uint32_t)P.op[2] << 16) | ((uint32_t)P.op[3] << 24);
        if (EPI::kRuntime && P.softmax_slot >= 0) {
          auto load_skip = [&](int c) {
            return (has_skip && valid) ? __ldg(reinterpret_cast<const uint4*>(sp + (size_t)c * plane_pitch)) : make_uint4(0, 0, 0, 0);
          };
          auto emit = [&](int c, const float (&o)[8]) {
            if (valid)
              *reinterpret_cast<uint4*>(dp + (size_t)c * plane_pitch) =
                  make_uint4(pack_op2(o[0], o[1]), pack_op2(o[2], o[3]), pack_op2(o[4], o[5]), pack_op2(o[6], o[7]));
          };
          softmax_row_rt(taddr, nplanes, P.cout, ops_packed, P.dparams + ch0, P.cpad, has_skip, SoftmaxCfg{P.softmax_slot, P.softmax_log},
                         load_skip, emit);
        } else
        for (int c = 0; c < nplanes; ++c) {
          uint4 skc = make_uint4(0, 0, 0, 0);
          if (has_skip && valid) skc = __ldg(reinterpret_cast<const uint4*>(sp + (size_t)c * plane_pitch));
          const float* prm = P.dparams + ch0 + c * 8;
          const float4 b0 = __ldg(reinterpret_cast<const float4*>(prm)), b1 = __ldg(reinterpret_cast<const float4*>(prm + 4));
          uint32_t v[8];
          tmem_ld_x8(taddr + c * 8, v);
          tmem_ld_wait();
          float o[8];
          o[0] = __uint_as_float(v[0]) + b0.x; o[1] = __uint_as_float(v[1]) + b0.y; o[2] = __uint_as_float(v[2]) + b0.z;
          o[3] = __uint_as_float(v[3]) + b0.w; o[4] = __uint_as_float(v[4]) + b1.x; o[5] = __uint_as_float(v[5]) + b1.y;
          o[6] = __uint_as_float(v[6]) + b1.z; o[7] = __uint_as_float(v[7]) + b1.w;
          epi_chain8<EPI>(ops_packed, prm, P.cpad, has_skip, skc, o);
          if (ch0 + c * 8 + 8 > P.cout) {
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] = (ch0 + c * 8 + i < P.cout) ? o[i] : 0.f;
          }
          if (valid)
            *reinterpret_cast<uint4*>(dp + (size_t)c * plane_pitch) =
                make_uint4(pack_op2(o[0], o[1]), pack_op2(o[2], o[3]), pack_op2(o[4], o[5]), pack_op2(o[6], o[7]));
        }
      } else if constexpr (KIND == EPI_TAIL_SHUFFLE) {
        // ---- 12 channels -> PixelShuffle(2) + network input + ReLU straight to the frame (model_pix_shuffle.py:293-296) ----
        uint32_t v[16];
        tmem_ld_x16(taddr, v);
        tmem_ld_wait();
        float o[12];
#pragma unroll
        for (int h = 0; h < 2; ++h) {          // channels 0..7 and 8..11 through the chunk-wise chain (4 padding channels ride along)
          float t8[8];
          const float* prm = P.dparams + h * 8;
          const float4 b0 = __ldg(reinterpret_cast<const float4*>(prm)), b1 = __ldg(reinterpret_cast<const float4*>(prm + 4));
          t8[0] = __uint_as_float(v[h * 8 + 0]) + b0.x; t8[1] = __uint_as_float(v[h * 8 + 1]) + b0.y;
          t8[2] = __uint_as_float(v[h * 8 + 2]) + b0.z; t8[3] = __uint_as_float(v[h * 8 + 3]) + b0.w;
          t8[4] = __uint_as_float(v[h * 8 + 4]) + b1.x; t8[5] = __uint_as_float(v[h * 8 + 5]) + b1.y;
          t8[6] = __uint_as_float(v[h * 8 + 6]) + b1.z; t8[7] = __uint_as_float(v[h * 8 + 7]) + b1.w;
          const uint32_t ops_packed = (uint32_t)P.op[0] | ((uint32_t)P.op[1] << 8) | ((uint32_t)P.op[2] << 16) | ((uint32_t)P.op[3] << 24);
          epi_chain8<EPI>(ops_packed, prm, P.cpad, false, make_uint4(0, 0, 0, 0), t8);
#pragma unroll
          for (int i = 0; i < 8; ++i)
            if (h * 8 + i < 12) o[h * 8 + i] = t8[i];
        }
        if (valid) {
          const size_t fpl = (size_t)P.H * P.W;
#pragma unroll
          for (int dy = 0; dy < 2; ++dy) {
            const size_t p0 = (size_t)(2 * y + dy) * P.W + 2 * x + P.xoff;
            float idv[3][2];
            if (P.in_fmt == FSUAE_FMT_F32_NCHW3) {
              const float* ip = (const float*)P.frame_in + (size_t)w.f * 3 * fpl + p0;
#pragma unroll
              for (int c = 0; c < 3; ++c) { const float2 t2 = __ldg(reinterpret_cast<const float2*>(ip + c * fpl)); idv[c][0] = t2.x; idv[c][1] = t2.y; }
            } else if (P.in_fmt == FSUAE_FMT_U8_NHWC4) {
              const uint2 t2 = __ldg(reinterpret_cast<const uint2*>((const unsigned char*)P.frame_in + ((size_t)w.f * fpl + p0) * 4));
#pragma unroll
              for (int c = 0; c < 3; ++c) { idv[c][0] = s_lut[(t2.x >> (8 * c)) & 0xFF]; idv[c][1] = s_lut[(t2.y >> (8 * c)) & 0xFF]; }
            } else {
              const unsigned char* ip = (const unsigned char*)P.frame_in + (size_t)w.f * 4 * fpl + p0;
#pragma unroll
              for (int c = 0; c < 3; ++c) { idv[c][0] = s_lut[__ldg(ip + c * fpl)]; idv[c][1] = s_lut[__ldg(ip + c * fpl + 1)]; }
            }
            float res[3][2];
#pragma unroll
            for (int c = 0; c < 3; ++c)
#pragma unroll
              for (int dx = 0; dx < 2; ++dx) res[c][dx] = fmaxf(o[c * 4 + dy * 2 + dx] + idv[c][dx], 0.f);
            if (P.out_fmt == FSUAE_FMT_F32_NCHW3) {
              float* op = (float*)P.frame_out + (size_t)w.f * 3 * fpl + p0;
#pragma unroll
              for (int c = 0; c < 3; ++c) *reinterpret_cast<float2*>(op + c * fpl) = make_float2(res[c][0], res[c][1]);
            } else {
              uint32_t px[2];
#pragma unroll
              for (int dx = 0; dx < 2; ++dx)
                px[dx] = (uint32_t)to_u8_fast(res[0][dx], P.gamma_out) | ((uint32_t)to_u8_fast(res[1][dx], P.gamma_out) << 8) |
                         ((uint32_t)to_u8_fast(res[2][dx], P.gamma_out) << 16) | 0xFF000000u;
              *reinterpret_cast<uint2*>((unsigned char*)P.frame_out + ((size_t)w.f * fpl + p0) * 4) = make_uint2(px[0], px[1]);
            }
          }
        }
      } else {
        // ---- 3-channel full-resolution tail straight to the frame (model_conv3.py:145-153, model_conv5.py:149) ----
        float prm[9][4];
#pragma unroll
        for (int a = 0; a < 9; ++a) {
          const float4 t4 = __ldg(reinterpret_cast<const float4*>(P.dparams + (size_t)a * P.cpad));
          prm[a][0] = t4.x; prm[a][1] = t4.y; prm[a][2] = t4.z; prm[a][3] = t4.w;
        }
        uint32_t v[8];
        tmem_ld_x8(taddr, v);
        tmem_ld_wait();
        float o[3];
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
          float t = __uint_as_float(v[ch]) + prm[0][ch];
          t = act_rt(EPI::kRuntime ? P.op[0] : EPI::kOp0, t, prm[1][ch], prm[5][ch]);
          t = act_rt(EPI::kRuntime ? P.op[1] : EPI::kOp1, t, prm[2][ch], prm[6][ch]);
          t = act_rt(EPI::kRuntime ? P.op[2] : EPI::kOp2, t, prm[3][ch], prm[7][ch]);
          o[ch] = act_rt(EPI::kRuntime ? P.op[3] : EPI::kOp3, t, prm[4][ch], prm[8][ch]);
        }
        if (valid) {
          const size_t fpl = (size_t)P.H * P.W;
          const size_t p0 = (size_t)y * P.W + x + P.xoff;
          if (P.out_fmt == FSUAE_FMT_F32_NCHW3) {
            float* op = (float*)P.frame_out + (size_t)w.f * 3 * fpl + p0;
            op[0] = o[0]; op[fpl] = o[1]; op[2 * fpl] = o[2];
          } else if (P.out_fmt == FSUAE_FMT_F32_NCHW4) {
            float* op = (float*)P.frame_out + (size_t)w.f * 4 * fpl + p0;
            op[0] = o[0] * 255.0f; op[fpl] = o[1] * 255.0f; op[2 * fpl] = o[2] * 255.0f; op[3 * fpl] = 255.0f;
          } else {
            const uint32_t px = (uint32_t)to_u8_fast(o[0], P.gamma_out) | ((uint32_t)to_u8_fast(o[1], P.gamma_out) << 8) |
                                ((uint32_t)to_u8_fast(o[2], P.gamma_out) << 16) | 0xFF000000u;
            *reinterpret_cast<uint32_t*>((unsigned char*)P.frame_out + ((size_t)w.f * fpl + p0) * 4) = px;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(&tempty[g], 0);
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) tmem_dealloc_2cta(tmem_base, 512);
}

// B operand stream of the wide kernel: [group][slice][rank][tap][half][NT/2][8]
std::vector<uint16_t> pack_weights_wide(const float* w, int cout, int cin0, int cin1, int P0, int P1, int NT, int ngroups) {
  const int pin = P0 + P1, cin = cin0 + cin1, kchunks = (pin + 1) / 2, NB = NT / 2;
  std::vector<uint16_t> out((size_t)ngroups * kchunks * 9 * 2 * NT * 8, 0);
  for (int ng = 0; ng < ngroups; ++ng)
    for (int kc = 0; kc < kchunks; ++kc)
      for (int r = 0; r < 2; ++r)
        for (int tap = 0; tap < 9; ++tap)
          for (int h = 0; h < 2; ++h) {
            const int j = 2 * kc + h;
            if (j >= pin) continue;                       // duplicated plane of an odd tail: zero weights
            for (int nn = 0; nn < NB; ++nn) {
              const int n = ng * NT + r * NB + nn;
              if (n >= cout) continue;
              for (int k = 0; k < 8; ++k) {
                int ci;
                if (j < P0) { ci = j * 8 + k; if (ci >= cin0) continue; }
                else { ci = (j - P0) * 8 + k; if (ci >= cin1) continue; ci += cin0; }
                const float v = w[((size_t)n * cin + ci) * 9 + tap];
                out[(((((size_t)(ng * kchunks + kc) * 2 + r) * 9 + tap) * 2 + h) * NB + nn) * 8 + k] = f2op(v);
              }
            }
          }
  return out;
}

// Window decomposition of a k x k kernel (k = 5, 7) into 3x3 windows: window centres along one axis and the window that owns
// each tap offset (the first one that reaches it).
struct KWindows {
  int n1;                 // windows per axis
  int centre[3];
  int owner[7];           // tap offset o + R -> window index along the axis
};
inline KWindows k_windows(int ksize) {
  KWindows kw{};
  const int R = ksize / 2;
  if (ksize == 3) { kw.n1 = 1; kw.centre[0] = 0; }
  else if (ksize == 5) { kw.n1 = 2; kw.centre[0] = -1; kw.centre[1] = 1; }
  else { kw.n1 = 3; kw.centre[0] = -2; kw.centre[1] = 0; kw.centre[2] = 2; }
  for (int o = -R; o <= R; ++o)
    for (int w = kw.n1 - 1; w >= 0; --w)
      if (std::abs(o - kw.centre[w]) <= 1) kw.owner[o + R] = w;
  return kw;
}
// [cout][nwin * pin_real * 8][3][3]: the k x k weights spread over the windows' virtual channels (plane-padded channel space)
inline std::vector<float> expand_weights_windows(const float* w, int cout, int cin0, int cin1, int P0, int P1, int ksize) {
  const KWindows kw = k_windows(ksize);
  const int R = ksize / 2, pin = P0 + P1, cin = cin0 + cin1, nwin = kw.n1 * kw.n1, cv = nwin * pin * 8;
  std::vector<float> out((size_t)cout * cv * 9, 0.f);
  for (int n = 0; n < cout; ++n)
    for (int pr = 0; pr < pin; ++pr)
      for (int k = 0; k < 8; ++k) {
        int ci;
        if (pr < P0) { ci = pr * 8 + k; if (ci >= cin0) continue; }
        else { ci = (pr - P0) * 8 + k; if (ci >= cin1) continue; ci += cin0; }
        for (int oy = -R; oy <= R; ++oy)
          for (int ox = -R; ox <= R; ++ox) {
            const int wy = kw.owner[oy + R], wx = kw.owner[ox + R], win = wy * kw.n1 + wx;
            const int dy = oy - kw.centre[wy] + 1, dx = ox - kw.centre[wx] + 1;      // tap inside the window's 3x3
            out[((size_t)n * cv + (size_t)(win * pin + pr) * 8 + k) * 9 + dy * 3 + dx] =
                w[(((size_t)n * cin + ci) * ksize + (oy + R)) * ksize + (ox + R)];
          }
      }
  return out;
}

typedef void (*WideFn)(const WideK);
struct WideVariant { int NT, KIND, pre0, pre1, post0, post1, skip; WideFn fn; int smem; };   // pre0 = -1: run-time op-codes (skip flag then run-time too)
template <int NT, int KIND, class EPI>
WideVariant make_wide() {
  return WideVariant{NT, KIND, EPI::kOp0, EPI::kOp1, EPI::kOp2, EPI::kOp3, EPI::kSkip ? 1 : 0, conv3x3_tc_wide_kernel<NT, KIND, EPI>, WideCfg<NT>::SMEM};
}
const std::vector<WideVariant>& wide_variants() {
  using RT = Epi<-1, -1, -1, -1, false>;
  static const std::vector<WideVariant> v = {
      make_wide<64, EPI_STORE, RT>(), make_wide<96, EPI_STORE, RT>(), make_wide<112, EPI_STORE, RT>(), make_wide<128, EPI_STORE, RT>(),
      make_wide<16, EPI_TAIL_PLAIN, RT>(), make_wide<16, EPI_TAIL_SHUFFLE, RT>(),
      // BatchNorm-folded families (model_conv3.py:127-145, model_conv5.py:123-149): ReLU, residual + ReLU, identity / sigmoid tails
      make_wide<128, EPI_STORE, Epi<FSUAE_ACT_RELU, 0, 0, 0, false>>(), make_wide<128, EPI_STORE, Epi<0, 0, FSUAE_ACT_RELU, 0, true>>(),
      make_wide<112, EPI_STORE, Epi<FSUAE_ACT_TELU, FSUAE_ACT_LEAKY_RELU, FSUAE_ACT_TANH, 0, true>>(),      // pix_shuffle heavyweight conv4
      make_wide<16, EPI_TAIL_PLAIN, Epi<0, 0, 0, 0, false>>(), make_wide<16, EPI_TAIL_PLAIN, Epi<FSUAE_ACT_SIGMOID, 0, 0, 0, false>>(),
  };
  return v;
}
