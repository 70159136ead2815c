// Parameters of the single fused pass (mega.cu), shared with the plan builder in bf16_tc.cu.  Plain data only.
#pragma once

#include <cuda_runtime.h>

namespace fsuae {

constexpr int MG_MAXC = 80;          // widest layer of the flagship (72 -> N = 80)
constexpr int MG_NCH = 7;            // channels: output of conv1..conv6, and the unshuffled input frame (the head's output)
#ifndef MG_DMAX_ROWS
#define MG_DMAX_ROWS 64
#endif
constexpr int MG_DMAX = MG_DMAX_ROWS;  // deepest ring
constexpr int MG_SMAX = 8;           // strips per row the flag block is laid out for
constexpr int MG_THREADS = 640;      // 4 service warps + 16 epilogue warps
constexpr int MG_FLAG_WORDS = MG_NCH * MG_DMAX + MG_NCH * 2 * MG_SMAX;      // per (team, rank): prod[ch][slot], cons[ch][consumer][strip]
#ifndef MG_D0
#define MG_D0 56                     // rows of the conv1 -> conv2 / conv6 ring: the long skip spans the whole pipeline
#endif
#ifndef MG_D1
#define MG_D1 13                     // rows of every other ring
#endif
__host__ __device__ constexpr int mg_depth(int ch) { return ch == 0 ? MG_D0 : MG_D1; }
static_assert(MG_D0 <= MG_DMAX && MG_D1 <= MG_DMAX, "flag block layout");

struct MegaLayerP {
  float bias[MG_MAXC];
  float p0[4][MG_MAXC];
  float p1[4][MG_MAXC];
  const unsigned char* wpack;        // CTA-pair packing of pack_weights(..., ctas = 2)
};

struct MegaK {
  int Hw, Ww, PW, S, n_frames, n_fp, teams;
  int H, W, xoff, in_fmt, out_fmt, gamma_in, gamma_out;
  const void* frame_in;
  void* frame_out;
  unsigned char* scratch;                       // [team][rank]{channel rings}
  unsigned long long rank_stride;               // bytes of one (team, rank) block; team stride = 2 * rank_stride
  unsigned long long ch_off[MG_NCH];            // channel c inside the block; plane stride = depth * PW * 16
  unsigned int* flags;                          // [team][rank][MG_FLAG_WORDS]
  const unsigned char* zero_row;                // one plane row of zeros (rows above / below the frame)
  MegaLayerP L[7];
};

// mega.cu, compiled once per operand type like bf16_tc.cu
int bf16_mega_prepare();                                               // raises the kernel's shared-memory limit; cudaError_t
int bf16_mega_launch(const MegaK& k, int grid, cudaStream_t st);      // grid = CTAs (8 per group of 4 stage pairs); cudaError_t
int fp16_mega_prepare();
int fp16_mega_launch(const MegaK& k, int grid, cudaStream_t st);

}  // namespace fsuae
