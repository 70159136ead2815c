// Internal engine state shared by the ABI layer and the two arithmetic builds.
#pragma once

#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "common.cuh"
#include "fsuae_enhancer.h"

namespace fsuae {

struct Bf16Plan;  // bf16_tc.cu
struct Fp16Plan;  // bf16_tc.cu built with -DFSUAE_OPERAND_FP16

// Experiment / test switches.  They are read from the environment ONCE, when an engine is created, and every one that is
// set is appended to fsuae_engine_variant(); no launch reads the environment.  None is needed in production.
struct Tuning {
  int grid = 0;                 // FSUAE_DEBUG_GRID=n: CTA count of the persistent tensor-core kernels (partition tests)
  int r3 = -1;                  // FSUAE_R3=0|1: three-output-rows-per-instruction kernels for no / every layer that has one
  bool no_fusion = false;       // FSUAE_NO_FUSION: conv3 -> conv4 as two kernels
  bool no_pairs = false;        // FSUAE_NO_PAIRS: single-CTA kernels only
  bool no_wide = false;         // FSUAE_NO_WIDE / FSUAE_FORCE_WIDE: K-streamed wide kernel off / for every layer without a
  bool force_wide = false;      //   compile-time variant
  bool no_mega = false;         // FSUAE_NO_MEGA: never use the single fused pass (layer-by-layer kernels only)
  int mega_min_frames = -1;     // FSUAE_MEGA_MIN_FRAMES=n: smallest pass the fused kernel takes
  std::vector<int> host_stages; // FSUAE_HOST_STAGES=a,b,...: stage sizes of the host-buffer pipeline
  int host_chunk = 0;           // FSUAE_HOST_CHUNK=n: largest stage of the host-buffer pipeline (default 32)
  std::string tag;              // " [FSUAE_R3=1 ...]" for the variant string
};

}  // namespace fsuae

#define FSUAE_STAGE_BUFS 3   // staging buffers of the host pipeline (upload / compute / download of three stages in flight)

struct fsuae_engine {
  fsuae_net_desc desc;
  int device = 0;
  int precision = 0;
  int H = 0, W = 0;          // full-resolution frame
  int chunk = 1;             // frames per internal pass
  int host_chunk = 1;        // frames per stage of the host-buffer pipeline
  int sm_count = 148;
  std::string variant;
  std::string last_error;
  int64_t launches = 0;
  size_t device_bytes = 0;

  // parameters
  float* d_blob = nullptr;
  size_t blob_floats = 0;
  std::vector<float> h_blob;

  // fp32 build: planar fp32 activations, one buffer per layer output (+ head buffer 0)
  std::vector<float*> f32_buf;      // [n_layers + 1]
  std::vector<int> buf_channels;    // channels of each buffer

  // tensor-core builds (bf16 / fp16 operands)
  fsuae::Bf16Plan* bf16 = nullptr;
  fsuae::Fp16Plan* fp16 = nullptr;
  fsuae::Tuning tuning;

  // optional per-launch timing (fsuae_engine_set_profiling)
  bool profiling = false;
  std::vector<cudaEvent_t> prof_ev;        // 2 events per launch slot
  std::vector<std::string> prof_label;
  int prof_n = 0;

  // run_host staging
  void* d_stage_in[FSUAE_STAGE_BUFS] = {};
  void* d_stage_out[FSUAE_STAGE_BUFS] = {};
  cudaStream_t s_in = nullptr, s_comp = nullptr, s_out = nullptr;
  unsigned long long stage_seq = 0;   // stages ever submitted: staging set = stage_seq % FSUAE_STAGE_BUFS, carries across submissions
  cudaEvent_t ev_in[FSUAE_STAGE_BUFS] = {}, ev_comp[FSUAE_STAGE_BUFS] = {}, ev_out[FSUAE_STAGE_BUFS] = {};
};

namespace fsuae {

// geometry of one pass
struct Geom {
  int H, W;      // full-res frame
  int xoff;      // first column the network sees (0 or 16)
  int We;        // effective full-res width = W - xoff
  int Hw, Ww;    // working resolution (half-res for the unshuffle head)
};

inline Geom make_geom(const fsuae_engine* e, uint32_t flags) {
  Geom g;
  g.H = e->H;
  g.W = e->W;
  g.xoff = (flags & FSUAE_FLAG_CROP16) ? 16 : 0;
  g.We = e->W - g.xoff;
  if (e->desc.head == FSUAE_HEAD_FEATURES) g.xoff = 0, g.We = e->W;
  if (e->desc.head == FSUAE_HEAD_UNSHUFFLE2) {
    g.Hw = e->H / 2;
    g.Ww = g.We / 2;
  } else {
    g.Hw = e->H;
    g.Ww = g.We;
  }
  return g;
}

inline size_t fmt_frame_bytes(int fmt, int H, int W) {
  switch (fmt) {
    case FSUAE_FMT_F32_NCHW3: return (size_t)3 * H * W * 4;
    case FSUAE_FMT_U8_NHWC4: return (size_t)4 * H * W;
    case FSUAE_FMT_U8_NCHW4: return (size_t)4 * H * W;
    case FSUAE_FMT_F32_NCHW4: return (size_t)4 * H * W * 4;
  }
  return 0;
}

// fp32 build (fp32_path.cu)
int fp32_create(fsuae_engine* e);
void fp32_destroy(fsuae_engine* e);
int fp32_enqueue_chunk(fsuae_engine* e, const void* in, void* out, int n, int in_fmt, int out_fmt,
                       uint32_t flags, cudaStream_t st);

// bf16 build (bf16_tc.cu)
int bf16_create(fsuae_engine* e);
void bf16_destroy(fsuae_engine* e);
int bf16_enqueue_chunk(fsuae_engine* e, const void* in, void* out, int n, int in_fmt, int out_fmt,
                       uint32_t flags, cudaStream_t st);

// fp16 build (bf16_tc.cu with -DFSUAE_OPERAND_FP16)
int fp16_create(fsuae_engine* e);
void fp16_destroy(fsuae_engine* e);
int fp16_enqueue_chunk(fsuae_engine* e, const void* in, void* out, int n, int in_fmt, int out_fmt,
                       uint32_t flags, cudaStream_t st);

int set_error(fsuae_engine* e, int code, const std::string& msg);

inline int enqueue_chunk(fsuae_engine* e, const void* in, void* out, int n, int in_fmt, int out_fmt, uint32_t flags,
                         cudaStream_t st) {
  switch (e->precision) {
    case FSUAE_PREC_FP32: return fp32_enqueue_chunk(e, in, out, n, in_fmt, out_fmt, flags, st);
    case FSUAE_PREC_BF16: return bf16_enqueue_chunk(e, in, out, n, in_fmt, out_fmt, flags, st);
    default: return fp16_enqueue_chunk(e, in, out, n, in_fmt, out_fmt, flags, st);
  }
}

// bracket one kernel launch with events when profiling is on
struct ProfScope {
  fsuae_engine* e; cudaStream_t st; int slot;
  ProfScope(fsuae_engine* e_, cudaStream_t st_, const char* label) : e(e_), st(st_), slot(-1) {
    if (!e->profiling) return;
    slot = e->prof_n++;
    while ((int)e->prof_ev.size() < 2 * (slot + 1)) { cudaEvent_t ev; cudaEventCreate(&ev); e->prof_ev.push_back(ev); }
    if ((int)e->prof_label.size() <= slot) e->prof_label.resize(slot + 1);
    e->prof_label[slot] = label;
    cudaEventRecord(e->prof_ev[2 * slot], st);
  }
  ~ProfScope() { if (slot >= 0) cudaEventRecord(e->prof_ev[2 * slot + 1], st); }
};

#define FSUAE_CUDA_CHECK(e, call)                                                              \
  do {                                                                                         \
    cudaError_t _err = (call);                                                                 \
    if (_err != cudaSuccess)                                                                   \
      return fsuae::set_error((e), FSUAE_ERR_CUDA,                                             \
                              std::string(#call) + ": " + cudaGetErrorString(_err));           \
  } while (0)

}  // namespace fsuae
