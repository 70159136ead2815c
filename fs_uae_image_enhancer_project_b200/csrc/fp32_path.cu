// fp32 build: every layer is a direct 3x3 convolution on the FMA pipes with the layer's whole
// epilogue (bias, activation chain, skip add, activation chain) fused in registers; activations
// use accurate libm math so the result tracks the PyTorch fp32 eval forward to ~1e-6.
// This is the correctness anchor of the engine (gate: max-abs <= 1e-5 vs the oracle) and the
// path that runs arbitrary activation configurations (incl. channel softmax).
//
// Reference semantics: model/model_pix_shuffle.py:227-298, model_conv3.py:102-155,
// model_conv5.py:114-151 (zero padding 1 at the frame border on EVERY layer's input).
#include <algorithm>

#include "engine.h"

namespace fsuae {

// ------------------------------------------------------------------------------------------
// head: frame -> buffer 0 (planar fp32 at working resolution)
// ------------------------------------------------------------------------------------------

__device__ __forceinline__ float u8_to_lin(const float* lut, uint8_t v) { return lut[v]; }

// One thread per working-resolution pixel.  UNSHUFFLE: channel c*4+dy*2+dx of (h,w) = in[c][2h+dy][2w+dx].
template <bool UNSHUFFLE>
__global__ void head_kernel(const void* __restrict__ in, float* __restrict__ buf0, int n_frames, int in_fmt,
                            int H, int W, int xoff, int Hw, int Ww, int gamma_in) {
  __shared__ float lut[256];
  if (in_fmt != FSUAE_FMT_F32_NCHW3) {
    for (int i = threadIdx.x; i < 256; i += blockDim.x) {
      float t = (float)i / 255.0f;
      lut[i] = gamma_in ? powf(t, 2.2f) : t;
    }
    __syncthreads();
  }
  const size_t plane = (size_t)Hw * Ww;
  const size_t total = (size_t)n_frames * plane;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    int f = (int)(idx / plane);
    int r = (int)(idx - (size_t)f * plane);
    int h = r / Ww, w = r - h * Ww;
    constexpr int S = UNSHUFFLE ? 2 : 1;
    float* o = buf0 + (size_t)f * (UNSHUFFLE ? 12 : 3) * plane + r;
#pragma unroll
    for (int dy = 0; dy < S; ++dy) {
#pragma unroll
      for (int dx = 0; dx < S; ++dx) {
        int y = h * S + dy, x = w * S + dx + xoff;
        float v[3];
        if (in_fmt == FSUAE_FMT_F32_NCHW3) {
          const float* p = (const float*)in + (size_t)f * 3 * H * W + (size_t)y * W + x;
          v[0] = p[0]; v[1] = p[(size_t)H * W]; v[2] = p[2 * (size_t)H * W];
        } else if (in_fmt == FSUAE_FMT_U8_NHWC4) {
          uchar4 q = ((const uchar4*)in)[(size_t)f * H * W + (size_t)y * W + x];
          v[0] = lut[q.x]; v[1] = lut[q.y]; v[2] = lut[q.z];
        } else {
          const uint8_t* p = (const uint8_t*)in + (size_t)f * 4 * H * W + (size_t)y * W + x;
          v[0] = lut[p[0]]; v[1] = lut[p[(size_t)H * W]]; v[2] = lut[p[2 * (size_t)H * W]];
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) o[(size_t)(UNSHUFFLE ? c * 4 + dy * 2 + dx : c) * plane] = v[c];
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// direct conv 3x3 + fused epilogue
// ------------------------------------------------------------------------------------------

struct ConvArgs {
  const float* src0; const float* src1; const float* skip;
  float* out;
  const float* w; const float* bias;
  int cin0, cin1, cout;
  int Hw, Ww;
  int co_groups;       // ceil(cout / CO_T)
  int epilogue;        // 0: bias only (chain runs in separate kernels)
  EpiDev epi;
};

constexpr int TX = 32, TYT = 8, PY = 2;  // tile 32 x 16 pixels, 256 threads

// KS x KS taps (3, 5, 7: model_pix_shuffle.py:108-115 padding = (k - 1) / 2); CK input channels per stage (8; 4 for 7x7 so the
// patch + weight stage stays inside the 48 KB of static shared memory)
template <int CO_T, int KS>
__global__ void __launch_bounds__(TX * TYT) conv3x3_fp32_kernel(ConvArgs a) {
  constexpr int R = KS / 2, CK = KS == 7 ? 4 : 8, KK = KS * KS;
  __shared__ float patch[CK][TYT * PY + 2 * R][TX + 2 * R];
  __shared__ __align__(16) float wsm[CK][KK][CO_T];

  const int tx = threadIdx.x % TX, ty = threadIdx.x / TX;
  const int x0 = blockIdx.x * TX, y0 = blockIdx.y * (TYT * PY);
  const int f = blockIdx.z / a.co_groups;
  const int co0 = (blockIdx.z % a.co_groups) * CO_T;
  const int cin = a.cin0 + a.cin1;
  const size_t plane = (size_t)a.Hw * a.Ww;
  const float* s0 = a.src0 + (size_t)f * a.cin0 * plane;
  const float* s1 = a.src1 ? a.src1 + (size_t)f * a.cin1 * plane : nullptr;

  float acc[PY][CO_T];
#pragma unroll
  for (int j = 0; j < PY; ++j)
#pragma unroll
    for (int c = 0; c < CO_T; ++c) acc[j][c] = 0.f;

  for (int c0 = 0; c0 < cin; c0 += CK) {
    // stage the input patch with zero padding outside the frame
    for (int i = threadIdx.x; i < CK * (TYT * PY + 2 * R) * (TX + 2 * R); i += TX * TYT) {
      int ci = i / ((TYT * PY + 2 * R) * (TX + 2 * R));
      int r = i - ci * ((TYT * PY + 2 * R) * (TX + 2 * R));
      int py = r / (TX + 2 * R), px = r - py * (TX + 2 * R);
      int y = y0 + py - R, x = x0 + px - R, c = c0 + ci;
      float v = 0.f;
      if (c < cin && y >= 0 && y < a.Hw && x >= 0 && x < a.Ww) {
        const float* p = c < a.cin0 ? s0 + (size_t)c * plane : s1 + (size_t)(c - a.cin0) * plane;
        v = __ldg(p + (size_t)y * a.Ww + x);
      }
      patch[ci][py][px] = v;
    }
    // stage weights [ci][tap][co]
    for (int i = threadIdx.x; i < CO_T * CK * KK; i += TX * TYT) {
      int co = i / (CK * KK);
      int r = i - co * (CK * KK);
      int ci = r / KK, tap = r - ci * KK;
      float v = 0.f;
      if (co0 + co < a.cout && c0 + ci < cin) v = __ldg(a.w + ((size_t)(co0 + co) * cin + (c0 + ci)) * KK + tap);
      wsm[ci][tap][co] = v;
    }
    __syncthreads();
#pragma unroll
    for (int ci = 0; ci < CK; ++ci) {
#pragma unroll
      for (int dy = 0; dy < KS; ++dy) {
#pragma unroll
        for (int dx = 0; dx < KS; ++dx) {
          float v[PY];
#pragma unroll
          for (int j = 0; j < PY; ++j) v[j] = patch[ci][ty + j * TYT + dy][tx + dx];
#pragma unroll
          for (int c4 = 0; c4 < CO_T / 4; ++c4) {
            float4 w4 = *reinterpret_cast<const float4*>(&wsm[ci][dy * KS + dx][c4 * 4]);
#pragma unroll
            for (int j = 0; j < PY; ++j) {
              acc[j][c4 * 4 + 0] = fmaf(v[j], w4.x, acc[j][c4 * 4 + 0]);
              acc[j][c4 * 4 + 1] = fmaf(v[j], w4.y, acc[j][c4 * 4 + 1]);
              acc[j][c4 * 4 + 2] = fmaf(v[j], w4.z, acc[j][c4 * 4 + 2]);
              acc[j][c4 * 4 + 3] = fmaf(v[j], w4.w, acc[j][c4 * 4 + 3]);
            }
          }
        }
      }
    }
    __syncthreads();
  }

  const int x = x0 + tx;
  if (x >= a.Ww) return;
#pragma unroll
  for (int j = 0; j < PY; ++j) {
    const int y = y0 + ty + j * TYT;
    if (y >= a.Hw) continue;
    const size_t pix = (size_t)y * a.Ww + x;
#pragma unroll
    for (int c = 0; c < CO_T; ++c) {
      const int co = co0 + c;
      if (co >= a.cout) break;
      float v = acc[j][c] + (a.bias ? __ldg(a.bias + co) : 0.f);
      if (a.epilogue) {
        for (int k = 0; k < a.epi.n_pre; ++k) v = act_apply_accurate(a.epi.pre[k], v, co);
        if (a.skip) v += __ldg(a.skip + ((size_t)f * a.cout + co) * plane + pix);
        for (int k = 0; k < a.epi.n_post; ++k) v = act_apply_accurate(a.epi.post[k], v, co);
      }
      a.out[((size_t)f * a.cout + co) * plane + pix] = v;
    }
  }
}

// ------------------------------------------------------------------------------------------
// out-of-line activation chain (only for layers whose chain contains a channel softmax)
// ------------------------------------------------------------------------------------------

__global__ void chain_segment_kernel(float* __restrict__ buf, const float* __restrict__ skip, int n_frames,
                                     int C, size_t plane, int n_acts, ActDev a0, ActDev a1, ActDev a2, ActDev a3) {
  const ActDev acts[4] = {a0, a1, a2, a3};
  const size_t total = (size_t)n_frames * C * plane;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    int ch = (int)((i / plane) % C);
    float v = buf[i];
    if (skip) v += skip[i];
    for (int k = 0; k < n_acts; ++k) v = act_apply_accurate(acts[k], v, ch);
    buf[i] = v;
  }
}

// softmax / log_softmax over the channel dimension (activations.py:147-148: dim=1); thread = pixel
__global__ void channel_softmax_kernel(float* __restrict__ buf, int n_frames, int C, size_t plane, int log_form) {
  const size_t total = (size_t)n_frames * plane;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    size_t f = i / plane, p = i - f * plane;
    float* b = buf + f * C * plane + p;
    float m = -INFINITY;
    for (int c = 0; c < C; ++c) m = fmaxf(m, b[(size_t)c * plane]);
    float s = 0.f;
    for (int c = 0; c < C; ++c) s += expf(b[(size_t)c * plane] - m);
    if (log_form) {
      float ls = logf(s);
      for (int c = 0; c < C; ++c) b[(size_t)c * plane] = b[(size_t)c * plane] - m - ls;
    } else {
      float inv = 1.f / s;
      for (int c = 0; c < C; ++c) b[(size_t)c * plane] = expf(b[(size_t)c * plane] - m) * inv;
    }
  }
}

// ------------------------------------------------------------------------------------------
// tail: last buffer (+ buffer 0 for the residual) -> output frame
// ------------------------------------------------------------------------------------------

__device__ __forceinline__ uint8_t to_u8(float v, int gamma_out) {
  if (gamma_out) v = powf(v, 1.0f / 2.2f);
  v = fminf(fmaxf(v, 0.f), 1.f) * 255.0f;   // NaN -> 0 via fmaxf
  return (uint8_t)v;                       // truncation (train.py:70-73, torch2onnx.py:585-632)
}

__global__ void tail_kernel(const float* __restrict__ last, const float* __restrict__ buf0, void* __restrict__ out,
                            int n_frames, int tail, int out_fmt, int H, int W, int xoff, int Hw, int Ww,
                            int gamma_out) {
  const size_t plane = (size_t)Hw * Ww;
  const size_t total = (size_t)n_frames * plane;
  const int S = tail == FSUAE_TAIL_SHUFFLE2_RESIDUAL_RELU ? 2 : 1;
  const size_t fplane = (size_t)H * W;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    int f = (int)(idx / plane);
    int r = (int)(idx - (size_t)f * plane);
    int h = r / Ww, w = r - h * Ww;
    for (int dy = 0; dy < S; ++dy)
      for (int dx = 0; dx < S; ++dx) {
        float v[3];
        for (int c = 0; c < 3; ++c) {
          if (S == 2) {
            size_t o = ((size_t)f * 12 + c * 4 + dy * 2 + dx) * plane + r;
            v[c] = fmaxf(last[o] + buf0[o], 0.f);   // model_pix_shuffle.py:293-296
          } else {
            v[c] = last[((size_t)f * 3 + c) * plane + r];
          }
        }
        int y = h * S + dy, x = w * S + dx + xoff;
        size_t pix = (size_t)y * W + x;
        if (out_fmt == FSUAE_FMT_F32_NCHW3) {
          float* o = (float*)out + (size_t)f * 3 * fplane + pix;
          o[0] = v[0]; o[fplane] = v[1]; o[2 * fplane] = v[2];
        } else if (out_fmt == FSUAE_FMT_F32_NCHW4) {
          float* o = (float*)out + (size_t)f * 4 * fplane + pix;   // model_conv3.py:145-153
          o[0] = v[0] * 255.0f; o[fplane] = v[1] * 255.0f; o[2 * fplane] = v[2] * 255.0f; o[3 * fplane] = 255.0f;
        } else {
          uchar4 q;
          q.x = to_u8(v[0], gamma_out); q.y = to_u8(v[1], gamma_out); q.z = to_u8(v[2], gamma_out); q.w = 255;
          ((uchar4*)out)[(size_t)f * fplane + pix] = q;
        }
      }
  }
}

// the 16 black columns of the CROP16 contract (torch2onnx.py:634-674)
__global__ void black_columns_kernel(void* __restrict__ out, int n_frames, int out_fmt, int H, int W, int ncols) {
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
      int C = out_fmt == FSUAE_FMT_F32_NCHW4 ? 4 : 3;
      float* o = (float*)out + (size_t)f * C * fplane + pix;
      for (int c = 0; c < C; ++c) o[(size_t)c * fplane] = (c == 3) ? 255.0f : 0.f;
    }
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------

static ActDev make_act_dev(const fsuae_act_desc& a, const float* d_blob) {
  ActDev r;
  r.op = a.op;
  r.n0 = a.n0;
  r.n1 = a.n1;
  r.p0 = a.n0 > 0 ? d_blob + a.p0_off : nullptr;
  r.p1 = a.n1 > 0 ? d_blob + a.p1_off : nullptr;
  return r;
}

int fp32_create(fsuae_engine* e) {
  const fsuae_net_desc& d = e->desc;
  const bool unshuffle = d.head == FSUAE_HEAD_UNSHUFFLE2;
  const size_t plane = unshuffle ? (size_t)(e->H / 2) * (e->W / 2) : (size_t)e->H * e->W;
  e->buf_channels.assign(d.n_layers + 1, 0);
  e->f32_buf.assign(d.n_layers + 1, nullptr);
  e->buf_channels[0] = unshuffle ? 12 : (d.head == FSUAE_HEAD_FEATURES ? 1 : 3);   // a feature-map input is used in place
  for (int i = 0; i < d.n_layers; ++i) e->buf_channels[i + 1] = d.layers[i].cout;
  for (int i = 0; i <= d.n_layers; ++i) {
    size_t bytes = (size_t)e->chunk * e->buf_channels[i] * plane * sizeof(float);
    FSUAE_CUDA_CHECK(e, cudaMalloc(&e->f32_buf[i], bytes));
    e->device_bytes += bytes;
  }
  e->variant = "fp32_fma";
  return FSUAE_OK;
}

void fp32_destroy(fsuae_engine* e) {
  for (float* p : e->f32_buf)
    if (p) cudaFree(p);
  e->f32_buf.clear();
}

template <int CO_T>
static void launch_conv(const ConvArgs& a, int ksize, int n, cudaStream_t st) {
  ConvArgs b = a;
  b.co_groups = (a.cout + CO_T - 1) / CO_T;
  dim3 grid((a.Ww + TX - 1) / TX, (a.Hw + TYT * PY - 1) / (TYT * PY), n * b.co_groups);
  if (ksize == 3) conv3x3_fp32_kernel<CO_T, 3><<<grid, TX * TYT, 0, st>>>(b);
  else if (ksize == 5) conv3x3_fp32_kernel<CO_T, 5><<<grid, TX * TYT, 0, st>>>(b);
  else conv3x3_fp32_kernel<CO_T, 7><<<grid, TX * TYT, 0, st>>>(b);
}

int fp32_enqueue_chunk(fsuae_engine* e, const void* in, void* out, int n, int in_fmt, int out_fmt,
                       uint32_t flags, cudaStream_t st) {
  const fsuae_net_desc& d = e->desc;
  const Geom g = make_geom(e, flags);
  const size_t plane = (size_t)g.Hw * g.Ww;
  const int ew_blocks = e->sm_count * 8;

  // head
  const float* buf0 = e->f32_buf[0];
  if ((d.head == FSUAE_HEAD_PLAIN && in_fmt == FSUAE_FMT_F32_NCHW3 && g.xoff == 0) || d.head == FSUAE_HEAD_FEATURES) {
    buf0 = (const float*)in;  // already planar fp32 at working resolution
  } else if (d.head == FSUAE_HEAD_UNSHUFFLE2) {
    head_kernel<true><<<ew_blocks, 256, 0, st>>>(in, e->f32_buf[0], n, in_fmt, g.H, g.W, g.xoff, g.Hw, g.Ww,
                                                 (flags & FSUAE_FLAG_GAMMA_IN) ? 1 : 0);
    e->launches++;
  } else {
    head_kernel<false><<<ew_blocks, 256, 0, st>>>(in, e->f32_buf[0], n, in_fmt, g.H, g.W, g.xoff, g.Hw, g.Ww,
                                                  (flags & FSUAE_FLAG_GAMMA_IN) ? 1 : 0);
    e->launches++;
  }

  auto buf = [&](int id) -> const float* { return id == 0 ? buf0 : e->f32_buf[id]; };

  for (int i = 0; i < d.n_layers; ++i) {
    const fsuae_layer_desc& L = d.layers[i];
    ConvArgs a{};
    a.src0 = buf(L.src0);
    a.src1 = L.cin1 > 0 ? buf(L.src1) : nullptr;
    a.skip = L.skip_src >= 0 ? buf(L.skip_src) : nullptr;
    // a feature-map network's last layer writes the caller's output (same planar fp32 layout)
    a.out = (i == d.n_layers - 1 && d.tail == FSUAE_TAIL_FEATURES) ? (float*)out : e->f32_buf[i + 1];
    a.w = e->d_blob + L.w_off;
    a.bias = L.b_off >= 0 ? e->d_blob + L.b_off : nullptr;
    a.cin0 = L.cin0; a.cin1 = L.cin1; a.cout = L.cout;
    a.Hw = g.Hw; a.Ww = g.Ww;
    bool has_softmax = false;
    for (int k = 0; k < L.n_pre; ++k) has_softmax |= act_is_softmax(L.pre[k].op);
    for (int k = 0; k < L.n_post; ++k) has_softmax |= act_is_softmax(L.post[k].op);
    a.epilogue = has_softmax ? 0 : 1;
    a.epi.n_pre = L.n_pre; a.epi.n_post = L.n_post;
    for (int k = 0; k < L.n_pre; ++k) a.epi.pre[k] = make_act_dev(L.pre[k], e->d_blob);
    for (int k = 0; k < L.n_post; ++k) a.epi.post[k] = make_act_dev(L.post[k], e->d_blob);
    if (L.cout % 16 == 0) launch_conv<16>(a, L.ksize, n, st);
    else if (L.cout % 12 == 0) launch_conv<12>(a, L.ksize, n, st);
    else if (L.cout <= 4) launch_conv<4>(a, L.ksize, n, st);
    else launch_conv<16>(a, L.ksize, n, st);
    e->launches++;

    if (has_softmax) {
      // run the chain as segments split at the softmax slots; the skip add sits between pre and post
      float* b = a.out;
      auto run_ops = [&](const fsuae_act_desc* acts, int cnt, const float* skip_first) {
        ActDev seg[4] = {};
        int ns = 0;
        const float* skip = skip_first;
        auto flush = [&]() {
          if (ns > 0 || skip) {
            chain_segment_kernel<<<ew_blocks, 256, 0, st>>>(b, skip, n, L.cout, plane, ns, seg[0], seg[1], seg[2], seg[3]);
            e->launches++;
          }
          ns = 0;
          skip = nullptr;
        };
        for (int k = 0; k < cnt; ++k) {
          if (act_is_softmax(acts[k].op)) {
            flush();
            channel_softmax_kernel<<<ew_blocks, 256, 0, st>>>(b, n, L.cout, plane, acts[k].op == FSUAE_ACT_LOG_SOFTMAX);
            e->launches++;
          } else {
            seg[ns++] = make_act_dev(acts[k], e->d_blob);
          }
        }
        flush();
      };
      run_ops(L.pre, L.n_pre, nullptr);
      run_ops(L.post, L.n_post, a.skip);
    }
  }

  if (d.tail != FSUAE_TAIL_FEATURES) {
    tail_kernel<<<ew_blocks, 256, 0, st>>>(e->f32_buf[d.n_layers], buf0, out, n, d.tail, out_fmt, g.H, g.W, g.xoff,
                                           g.Hw, g.Ww, (flags & FSUAE_FLAG_GAMMA_OUT) ? 1 : 0);
    e->launches++;
  }
  if (g.xoff > 0) {
    black_columns_kernel<<<64, 256, 0, st>>>(out, n, out_fmt, g.H, g.W, g.xoff);
    e->launches++;
  }
  FSUAE_CUDA_CHECK(e, cudaGetLastError());
  return FSUAE_OK;
}

}  // namespace fsuae
