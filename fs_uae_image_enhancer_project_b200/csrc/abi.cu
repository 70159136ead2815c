// C ABI of the engine (include/fsuae_enhancer.h): descriptor validation, lifetime, chunk loop,
// host-buffer pipeline.  The arithmetic lives in fp32_path.cu / bf16_tc.cu.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>

#include "engine.h"

namespace fsuae {

static std::string g_create_error;
static std::mutex g_create_mu;

int set_error(fsuae_engine* e, int code, const std::string& msg) {
  if (e) {
    e->last_error = msg;
  } else {
    std::lock_guard<std::mutex> lk(g_create_mu);
    g_create_error = msg;
  }
  return code;
}

// experiment switches: environment -> Tuning, once per engine
static Tuning read_tuning() {
  Tuning t;
  auto flag = [&](const char* name, bool& dst) {
    if (const char* v = getenv(name); v && *v) { dst = true; t.tag += std::string(" ") + name; }
  };
  auto num = [&](const char* name, int& dst) {
    if (const char* v = getenv(name); v && *v) { dst = atoi(v); t.tag += std::string(" ") + name + "=" + v; }
  };
  num("FSUAE_DEBUG_GRID", t.grid);
  num("FSUAE_R3", t.r3);
  num("FSUAE_MEGA_MIN_FRAMES", t.mega_min_frames);
  num("FSUAE_HOST_CHUNK", t.host_chunk);
  flag("FSUAE_NO_FUSION", t.no_fusion);
  flag("FSUAE_NO_PAIRS", t.no_pairs);
  flag("FSUAE_NO_WIDE", t.no_wide);
  flag("FSUAE_FORCE_WIDE", t.force_wide);
  flag("FSUAE_NO_MEGA", t.no_mega);
  if (const char* env = getenv("FSUAE_HOST_STAGES"); env && *env) {
    for (const char* c = env; *c;) { t.host_stages.push_back(std::max(1, atoi(c))); while (*c && *c != ',') ++c; if (*c) ++c; }
    t.tag += std::string(" FSUAE_HOST_STAGES=") + env;
  }
  return t;
}

static int act_param_counts(int op, int* need0, int* need1) {
  // returns 0 if the op is known; need{0,1}: 0 = unused, 1 = scalar, 2 = 1-or-C
  *need0 = *need1 = 0;
  switch (op) {
    case FSUAE_ACT_IDENTITY: case FSUAE_ACT_RELU: case FSUAE_ACT_RELU6: case FSUAE_ACT_TANH:
    case FSUAE_ACT_SIGMOID: case FSUAE_ACT_SILU: case FSUAE_ACT_MISH: case FSUAE_ACT_GELU:
    case FSUAE_ACT_SCALED_TANH: case FSUAE_ACT_TELU: case FSUAE_ACT_SOFTMAX:
    case FSUAE_ACT_LOG_SOFTMAX:
      return 0;
    case FSUAE_ACT_ELU: case FSUAE_ACT_LEAKY_RELU: *need0 = 1; return 0;
    case FSUAE_ACT_SOFTPLUS: case FSUAE_ACT_SINLU: *need0 = 1; *need1 = 1; return 0;
    case FSUAE_ACT_PRELU: case FSUAE_ACT_BIASED_RELU: *need0 = 2; return 0;
    case FSUAE_ACT_BIASED_PRELU: *need0 = 2; *need1 = 2; return 0;
  }
  return -1;
}

static int validate(const fsuae_net_desc* d, size_t blob_floats, int H, int W, std::string* why) {
  auto bad = [&](const std::string& s) { *why = s; return FSUAE_ERR_INVALID; };
  if (d->abi_version != FSUAE_ABI_VERSION) return bad("descriptor abi_version mismatch");
  if (d->n_layers < 1 || d->n_layers > FSUAE_MAX_LAYERS) return bad("n_layers out of range");
  if (d->head < FSUAE_HEAD_PLAIN || d->head > FSUAE_HEAD_FEATURES) return bad("unknown head");
  if (d->tail < FSUAE_TAIL_PLAIN || d->tail > FSUAE_TAIL_FEATURES) return bad("unknown tail");
  if ((d->head == FSUAE_HEAD_FEATURES) != (d->tail == FSUAE_TAIL_FEATURES))
    return bad("the feature-map head and tail come together");
  if (d->head == FSUAE_HEAD_FEATURES && (d->in_channels < 1 || d->in_channels > 4096)) return bad("in_channels out of range");
  if (H < 2 || W < 2) return bad("frame too small");
  if (d->head == FSUAE_HEAD_UNSHUFFLE2 && ((H | W) & 1))
    return bad("PixelUnshuffle(2) needs even height and width");
  std::vector<int> ch(d->n_layers + 1);
  ch[0] = d->head == FSUAE_HEAD_UNSHUFFLE2 ? 12 : (d->head == FSUAE_HEAD_FEATURES ? d->in_channels : 3);
  for (int i = 0; i < d->n_layers; ++i) {
    const fsuae_layer_desc& L = d->layers[i];
    const std::string tag = "layer " + std::to_string(i + 1) + ": ";
    if (L.cout < 1 || L.cin0 < 1 || L.cin1 < 0) return bad(tag + "bad channel count");
    if (L.ksize != 3 && L.ksize != 5 && L.ksize != 7) return bad(tag + "kernel size must be 3, 5 or 7 (1x1: centre tap of 3x3)");
    if (L.src0 < 0 || L.src0 > i || ch[L.src0] != L.cin0) return bad(tag + "src0 mismatch");
    if (L.cin1 > 0 && (L.src1 < 0 || L.src1 > i || ch[L.src1] != L.cin1)) return bad(tag + "src1 mismatch");
    if (L.skip_src >= 0 && (L.skip_src > i || ch[L.skip_src] != L.cout))
      return bad(tag + "skip source channel mismatch (a 1x1 skip projection is described as a layer of its own)");
    size_t wn = (size_t)L.cout * (L.cin0 + L.cin1) * L.ksize * L.ksize;
    if (L.w_off < 0 || (size_t)L.w_off + wn > blob_floats) return bad(tag + "weights outside blob");
    if (L.b_off >= 0 && (size_t)L.b_off + L.cout > blob_floats) return bad(tag + "bias outside blob");
    if (L.n_pre < 0 || L.n_pre > FSUAE_MAX_ACTS || L.n_post < 0 || L.n_post > FSUAE_MAX_ACTS)
      return bad(tag + "too many activation slots");
    for (int k = 0; k < L.n_pre + L.n_post; ++k) {
      const fsuae_act_desc& a = k < L.n_pre ? L.pre[k] : L.post[k - L.n_pre];
      int n0, n1;
      if (act_param_counts(a.op, &n0, &n1) != 0)
        return bad(tag + "Unsupported activation op-code " + std::to_string(a.op));
      auto chk = [&](int need, int n, int off) {
        if (need == 0) return true;
        if (need == 1 && n != 1) return false;
        if (need == 2 && n != 1 && n != L.cout) return false;
        return off >= 0 && (size_t)off + n <= blob_floats;
      };
      if (!chk(n0, a.n0, a.p0_off) || !chk(n1, a.n1, a.p1_off))
        return bad(tag + "activation parameter count must be 1 or C and lie inside the blob");
    }
    ch[i + 1] = L.cout;
  }
  int last = ch[d->n_layers];
  if (d->tail == FSUAE_TAIL_SHUFFLE2_RESIDUAL_RELU && (last != 12 || d->head != FSUAE_HEAD_UNSHUFFLE2))
    return bad("shuffle tail needs 12 output channels and the unshuffle head");
  if ((d->tail == FSUAE_TAIL_PLAIN || d->tail == FSUAE_TAIL_SCALE255_ALPHA) && (last != 3 || d->head != FSUAE_HEAD_PLAIN))
    return bad("plain / scale255 tail needs 3 output channels and the plain head");
  return FSUAE_OK;
}

static int check_formats(fsuae_engine* e, int in_fmt, int out_fmt, uint32_t flags) {
  if (e->desc.head == FSUAE_HEAD_FEATURES) {      // feature-map networks: float [B,C,H,W] in and out, nothing else
    if (in_fmt != FSUAE_FMT_F32_NCHW || out_fmt != FSUAE_FMT_F32_NCHW || flags != 0)
      return set_error(e, FSUAE_ERR_INVALID, "a feature-map network takes FSUAE_FMT_F32_NCHW in and out, no flags");
    return FSUAE_OK;
  }
  if (in_fmt != FSUAE_FMT_F32_NCHW3 && in_fmt != FSUAE_FMT_U8_NHWC4 && in_fmt != FSUAE_FMT_U8_NCHW4)
    return set_error(e, FSUAE_ERR_INVALID, "invalid input format");
  if (out_fmt != FSUAE_FMT_F32_NCHW3 && out_fmt != FSUAE_FMT_U8_NHWC4 && out_fmt != FSUAE_FMT_F32_NCHW4)
    return set_error(e, FSUAE_ERR_INVALID, "invalid output format");
  if ((out_fmt == FSUAE_FMT_F32_NCHW4) != (e->desc.tail == FSUAE_TAIL_SCALE255_ALPHA) &&
      out_fmt != FSUAE_FMT_U8_NHWC4)
    return set_error(e, FSUAE_ERR_INVALID, "float output format does not match the network tail");
  if (flags & ~(FSUAE_FLAG_GAMMA_IN | FSUAE_FLAG_GAMMA_OUT | FSUAE_FLAG_CROP16))
    return set_error(e, FSUAE_ERR_INVALID, "unknown flag bits");
  if ((flags & FSUAE_FLAG_GAMMA_IN) && in_fmt == FSUAE_FMT_F32_NCHW3)
    return set_error(e, FSUAE_ERR_INVALID, "GAMMA_IN applies to uint8 input only");
  if ((flags & FSUAE_FLAG_GAMMA_OUT) && out_fmt != FSUAE_FMT_U8_NHWC4)
    return set_error(e, FSUAE_ERR_INVALID, "GAMMA_OUT applies to uint8 output only");
  if ((flags & FSUAE_FLAG_CROP16) && (e->W <= 16 + 2))
    return set_error(e, FSUAE_ERR_INVALID, "CROP16 needs width > 18");
  return FSUAE_OK;
}

static size_t frame_bytes(const fsuae_engine* e, int fmt, bool input) {
  if (fmt == FSUAE_FMT_F32_NCHW)
    return (size_t)(input ? e->desc.in_channels : e->desc.layers[e->desc.n_layers - 1].cout) * e->H * e->W * 4;
  return fmt_frame_bytes(fmt, e->H, e->W);
}

}  // namespace fsuae

using namespace fsuae;

extern "C" {

int fsuae_abi_version(void) { return FSUAE_ABI_VERSION; }

int fsuae_engine_create(const fsuae_net_desc* desc, const float* blob, size_t blob_floats, int device,
                        int precision, int height, int width, int max_chunk_frames, fsuae_engine** out) {
  if (!desc || !blob || !out) return set_error(nullptr, FSUAE_ERR_INVALID, "null argument");
  *out = nullptr;
  std::string why;
  int rc = validate(desc, blob_floats, height, width, &why);
  if (rc != FSUAE_OK) return set_error(nullptr, rc, why);
  if (precision != FSUAE_PREC_FP32 && precision != FSUAE_PREC_BF16 && precision != FSUAE_PREC_FP16)
    return set_error(nullptr, FSUAE_ERR_INVALID, "unknown precision");
  if (max_chunk_frames > FSUAE_MAX_CHUNK_FRAMES)
    return set_error(nullptr, FSUAE_ERR_INVALID, "max_chunk_frames above FSUAE_MAX_CHUNK_FRAMES (enqueue accepts any frame count and loops)");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return set_error(nullptr, FSUAE_ERR_NO_DEVICE, "no CUDA device visible (this engine has no CPU path)");
  if (device < 0 || device >= ndev) return set_error(nullptr, FSUAE_ERR_INVALID, "device index out of range");
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess)
    return set_error(nullptr, FSUAE_ERR_CUDA, "cudaGetDeviceProperties failed");
  if (prop.major != 10)
    return set_error(nullptr, FSUAE_ERR_NO_DEVICE,
                     std::string("kernels are built for sm_100a only; device is ") + prop.name);

  fsuae_engine* e = new (std::nothrow) fsuae_engine();
  if (!e) return set_error(nullptr, FSUAE_ERR_CUDA, "out of host memory");
  e->tuning = read_tuning();
  e->desc = *desc;
  e->device = device;
  e->precision = precision;
  e->H = height;
  e->W = width;
  e->chunk = std::max(1, max_chunk_frames);
  e->sm_count = prop.multiProcessorCount;
  e->h_blob.assign(blob, blob + blob_floats);
  e->blob_floats = blob_floats;

  auto fail = [&](int code) {
    set_error(nullptr, code, e->last_error);
    fsuae_engine_destroy(e);
    return code;
  };
  int prev = 0;
  cudaGetDevice(&prev);
  cudaError_t ce = cudaSetDevice(device);
  if (ce == cudaSuccess) ce = cudaMalloc(&e->d_blob, blob_floats * sizeof(float));
  if (ce == cudaSuccess) ce = cudaMemcpy(e->d_blob, blob, blob_floats * sizeof(float), cudaMemcpyHostToDevice);
  if (ce != cudaSuccess) {
    e->last_error = std::string("parameter upload: ") + cudaGetErrorString(ce);
    cudaSetDevice(prev);
    return fail(FSUAE_ERR_CUDA);
  }
  e->device_bytes += blob_floats * sizeof(float);

  rc = precision == FSUAE_PREC_FP32 ? fp32_create(e) : precision == FSUAE_PREC_BF16 ? bf16_create(e) : fp16_create(e);
  if (rc == FSUAE_OK) e->variant += e->tuning.tag.empty() ? "" : " [" + e->tuning.tag.substr(1) + "]";
  if (rc != FSUAE_OK) {
    cudaSetDevice(prev);
    return fail(rc);
  }

  // host-pipeline staging: sized for the widest formats
  e->host_chunk = std::min(e->chunk, e->tuning.host_chunk > 0 ? e->tuning.host_chunk : 32);  // largest stage of the host-buffer pipeline (H2D / compute / D2H overlap): the fused pass is 9 % faster per frame on 32 frames than on 16
  size_t in_b = (size_t)e->host_chunk * std::max<size_t>(12, frame_bytes(e, FSUAE_FMT_F32_NCHW, true) / ((size_t)height * width)) * height * width;
  size_t out_b = (size_t)e->host_chunk * std::max<size_t>(16, frame_bytes(e, FSUAE_FMT_F32_NCHW, false) / ((size_t)height * width)) * height * width;
  for (int i = 0; i < FSUAE_STAGE_BUFS && ce == cudaSuccess; ++i) {
    ce = cudaMalloc(&e->d_stage_in[i], in_b);
    if (ce == cudaSuccess) ce = cudaMalloc(&e->d_stage_out[i], out_b);
    if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&e->ev_in[i], cudaEventDisableTiming);
    if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&e->ev_comp[i], cudaEventDisableTiming);
    if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&e->ev_out[i], cudaEventDisableTiming);
    e->device_bytes += in_b + out_b;
  }
  if (ce == cudaSuccess) ce = cudaStreamCreateWithFlags(&e->s_in, cudaStreamNonBlocking);
  if (ce == cudaSuccess) ce = cudaStreamCreateWithFlags(&e->s_comp, cudaStreamNonBlocking);
  if (ce == cudaSuccess) ce = cudaStreamCreateWithFlags(&e->s_out, cudaStreamNonBlocking);
  cudaSetDevice(prev);
  if (ce != cudaSuccess) {
    e->last_error = std::string("staging allocation: ") + cudaGetErrorString(ce);
    return fail(FSUAE_ERR_CUDA);
  }
  *out = e;
  return FSUAE_OK;
}

int fsuae_engine_create_from_file(const char* path, int device, int precision, int height, int width,
                                  int max_chunk_frames, fsuae_engine** out) {
  if (!path || !out) return set_error(nullptr, FSUAE_ERR_INVALID, "null argument");
  *out = nullptr;
  FILE* f = fopen(path, "rb");
  if (!f) return set_error(nullptr, FSUAE_ERR_INVALID, std::string("cannot open engine file ") + path);
  char magic[8];
  uint32_t hdr[2];
  fsuae_net_desc desc;
  std::vector<float> blob;
  bool ok = fread(magic, 1, 8, f) == 8 && memcmp(magic, "FSUAEENG", 8) == 0 && fread(hdr, 4, 2, f) == 2 &&
            hdr[0] == FSUAE_ABI_VERSION && fread(&desc, sizeof(desc), 1, f) == 1;
  if (ok) {      // the header's float count must be what the file actually holds (a corrupt count must not drive an allocation)
    const long pos = ftell(f);
    ok = pos >= 0 && fseek(f, 0, SEEK_END) == 0;
    const long end = ok ? ftell(f) : -1;
    ok = ok && end >= pos && (unsigned long long)(end - pos) == 4ull * hdr[1] && fseek(f, pos, SEEK_SET) == 0;
  }
  if (ok) {
    try {
      blob.resize(hdr[1]);
    } catch (const std::exception&) {
      ok = false;
    }
    ok = ok && fread(blob.data(), 4, blob.size(), f) == blob.size();
  }
  fclose(f);
  if (!ok) return set_error(nullptr, FSUAE_ERR_INVALID, std::string("malformed engine file ") + path);
  return fsuae_engine_create(&desc, blob.data(), blob.size(), device, precision, height, width, max_chunk_frames, out);
}

int fsuae_engine_destroy(fsuae_engine* e) {
  if (!e) return FSUAE_OK;
  int prev = 0;
  cudaGetDevice(&prev);
  cudaSetDevice(e->device);
  cudaDeviceSynchronize();
  fp32_destroy(e);
  bf16_destroy(e);
  fp16_destroy(e);
  for (cudaEvent_t ev : e->prof_ev) cudaEventDestroy(ev);
  for (int i = 0; i < FSUAE_STAGE_BUFS; ++i) {
    if (e->d_stage_in[i]) cudaFree(e->d_stage_in[i]);
    if (e->d_stage_out[i]) cudaFree(e->d_stage_out[i]);
    if (e->ev_in[i]) cudaEventDestroy(e->ev_in[i]);
    if (e->ev_comp[i]) cudaEventDestroy(e->ev_comp[i]);
    if (e->ev_out[i]) cudaEventDestroy(e->ev_out[i]);
  }
  if (e->s_in) cudaStreamDestroy(e->s_in);
  if (e->s_comp) cudaStreamDestroy(e->s_comp);
  if (e->s_out) cudaStreamDestroy(e->s_out);
  if (e->d_blob) cudaFree(e->d_blob);
  cudaSetDevice(prev);
  delete e;
  return FSUAE_OK;
}

static int enqueue_impl(fsuae_engine* e, const void* in_dev, void* out_dev, int n_frames, int in_fmt,
                        int out_fmt, uint32_t flags, cudaStream_t st) {
  size_t in_fb = frame_bytes(e, in_fmt, true), out_fb = frame_bytes(e, out_fmt, false);
  for (int f0 = 0; f0 < n_frames; f0 += e->chunk) {
    int n = std::min(e->chunk, n_frames - f0);
    const char* ip = (const char*)in_dev + (size_t)f0 * in_fb;
    char* op = (char*)out_dev + (size_t)f0 * out_fb;
    int rc = enqueue_chunk(e, ip, op, n, in_fmt, out_fmt, flags, st);
    if (rc != FSUAE_OK) return rc;
  }
  return FSUAE_OK;
}

int fsuae_engine_enqueue(fsuae_engine* e, const void* in_dev, void* out_dev, int n_frames, int in_fmt,
                         int out_fmt, uint32_t flags, void* cuda_stream) {
  if (!e) return FSUAE_ERR_INVALID;
  if (!in_dev || !out_dev || n_frames < 0) return set_error(e, FSUAE_ERR_INVALID, "null buffer or negative frame count");
  int rc = check_formats(e, in_fmt, out_fmt, flags);
  if (rc != FSUAE_OK) return rc;
  e->launches = 0;
  e->prof_n = 0;
  if (n_frames == 0) return FSUAE_OK;
  int prev = 0;
  cudaGetDevice(&prev);
  if (prev != e->device) cudaSetDevice(e->device);
  rc = enqueue_impl(e, in_dev, out_dev, n_frames, in_fmt, out_fmt, flags, (cudaStream_t)cuda_stream);
  if (prev != e->device) cudaSetDevice(prev);
  return rc;
}

int fsuae_engine_submit_host(fsuae_engine* e, const void* in_host, void* out_host, int n_frames, int in_fmt,
                             int out_fmt, uint32_t flags) {
  if (!e) return FSUAE_ERR_INVALID;
  if (!in_host || !out_host || n_frames < 0) return set_error(e, FSUAE_ERR_INVALID, "null buffer or negative frame count");
  int rc = check_formats(e, in_fmt, out_fmt, flags);
  if (rc != FSUAE_OK) return rc;
  e->launches = 0;
  e->prof_n = 0;
  if (n_frames == 0) return FSUAE_OK;
  int prev = 0;
  cudaGetDevice(&prev);
  if (prev != e->device) cudaSetDevice(e->device);
  size_t in_fb = frame_bytes(e, in_fmt, true), out_fb = frame_bytes(e, out_fmt, false);
  rc = FSUAE_OK;
  // Stage sizes (measured on a B200 / PCIe 5 x16 box, 64-frame calls, tools/e2e_stage_sweep.sh): when earlier
  // submissions are still in flight the stream is full and uniform large stages are best (22.8 k frames/s); when the
  // pipeline is idle the call is most likely a blocking one that pays fill (first upload) and drain (last download),
  // so stages ramp up 4, 6, 8, 10, 12 ... and back down ... 10, 8, 6 (19.5 k frames/s against 16.9 k for uniform 16).
  std::vector<int> stages;
  if (!e->tuning.host_stages.empty()) {      // experiment aid (FSUAE_HOST_STAGES at creation): explicit schedule "2,4,8,..." (repeated to cover n_frames)
    const std::vector<int>& pat = e->tuning.host_stages;
    for (int rem = n_frames, i = 0; rem > 0; ++i) { int t = std::min(std::min(pat[i % pat.size()], e->host_chunk), rem); stages.push_back(t); rem -= t; }
  } else if (e->stage_seq > 0 && cudaStreamQuery(e->s_out) == cudaErrorNotReady) {
    for (int rem = n_frames; rem > 0; rem -= stages.back()) stages.push_back(std::min(e->host_chunk, rem));
  } else {
    const int up[4] = {4, 6, 8, 10}, down[3] = {10, 8, 6};
    std::vector<int> tail;
    int rem = n_frames;
    if (n_frames >= 40)
      for (int d : down) { tail.push_back(std::min(d, e->host_chunk)); rem -= tail.back(); }
    for (int u : up) {
      if (rem <= 0) break;
      stages.push_back(std::min(std::min(u, e->host_chunk), rem));
      rem -= stages.back();
    }
    while (rem > 0) { stages.push_back(std::min(std::min(12, e->host_chunk), rem)); rem -= stages.back(); }
    stages.insert(stages.end(), tail.begin(), tail.end());
  }
  int f0 = 0;
  for (size_t it = 0; it < stages.size() && rc == FSUAE_OK; ++it) {
    const int n = stages[it];
    const int b = (int)(e->stage_seq % FSUAE_STAGE_BUFS);
    const bool reused = e->stage_seq >= FSUAE_STAGE_BUFS;     // staging pair b has been used before (possibly by an earlier submission)
    cudaError_t ce = cudaSuccess;
    // Input staging b is free once the pass that read it (two stages ago) has finished; output staging b once its
    // download has.  Keeping the two dependencies apart lets upload(i), compute(i-1) and download(i-2) overlap with
    // two buffers each (tied together, the period was upload + download instead of max(upload, compute, download)).
    if (reused) ce = cudaStreamWaitEvent(e->s_in, e->ev_comp[b], 0);
    if (ce == cudaSuccess)
      ce = cudaMemcpyAsync(e->d_stage_in[b], (const char*)in_host + (size_t)f0 * in_fb, (size_t)n * in_fb,
                           cudaMemcpyHostToDevice, e->s_in);
    if (ce == cudaSuccess) ce = cudaEventRecord(e->ev_in[b], e->s_in);
    if (ce == cudaSuccess) ce = cudaStreamWaitEvent(e->s_comp, e->ev_in[b], 0);
    if (ce == cudaSuccess && reused) ce = cudaStreamWaitEvent(e->s_comp, e->ev_out[b], 0);
    if (ce != cudaSuccess) { rc = set_error(e, FSUAE_ERR_CUDA, cudaGetErrorString(ce)); break; }
    const int64_t launches = e->launches;
    rc = enqueue_chunk(e, e->d_stage_in[b], e->d_stage_out[b], n, in_fmt, out_fmt, flags, e->s_comp);
    (void)launches;
    if (rc != FSUAE_OK) break;
    ce = cudaEventRecord(e->ev_comp[b], e->s_comp);
    if (ce == cudaSuccess) ce = cudaStreamWaitEvent(e->s_out, e->ev_comp[b], 0);
    if (ce == cudaSuccess)
      ce = cudaMemcpyAsync((char*)out_host + (size_t)f0 * out_fb, e->d_stage_out[b], (size_t)n * out_fb,
                           cudaMemcpyDeviceToHost, e->s_out);
    if (ce == cudaSuccess) ce = cudaEventRecord(e->ev_out[b], e->s_out);
    if (ce != cudaSuccess) rc = set_error(e, FSUAE_ERR_CUDA, cudaGetErrorString(ce));
    e->stage_seq++;
    f0 += n;
  }
  if (prev != e->device) cudaSetDevice(prev);
  return rc;
}

int fsuae_engine_wait_host(fsuae_engine* e) {
  if (!e) return FSUAE_ERR_INVALID;
  int prev = 0;
  cudaGetDevice(&prev);
  if (prev != e->device) cudaSetDevice(e->device);
  cudaError_t ce = cudaStreamSynchronize(e->s_out);
  cudaError_t ce2 = cudaStreamSynchronize(e->s_comp);
  cudaError_t ce3 = cudaStreamSynchronize(e->s_in);
  if (prev != e->device) cudaSetDevice(prev);
  if (ce != cudaSuccess || ce2 != cudaSuccess || ce3 != cudaSuccess)
    return set_error(e, FSUAE_ERR_CUDA, cudaGetErrorString(ce != cudaSuccess ? ce : (ce2 != cudaSuccess ? ce2 : ce3)));
  return FSUAE_OK;
}

int fsuae_engine_run_host(fsuae_engine* e, const void* in_host, void* out_host, int n_frames, int in_fmt,
                          int out_fmt, uint32_t flags) {
  int rc = fsuae_engine_submit_host(e, in_host, out_host, n_frames, in_fmt, out_fmt, flags);
  if (!e) return rc;
  const int rc2 = fsuae_engine_wait_host(e);      // drain even after a failed submission: nothing stays in flight
  return rc != FSUAE_OK ? rc : rc2;
}

int fsuae_engine_set_profiling(fsuae_engine* e, int enabled) {
  if (!e) return FSUAE_ERR_INVALID;
  e->profiling = enabled != 0;
  e->prof_n = 0;
  return FSUAE_OK;
}

float fsuae_engine_kernel_time(fsuae_engine* e, int index, char* label, int label_bytes) {
  if (!e || index < 0 || index >= e->prof_n) return -1.f;
  float ms = 0.f;
  if (cudaEventElapsedTime(&ms, e->prof_ev[2 * index], e->prof_ev[2 * index + 1]) != cudaSuccess) return -1.f;
  if (label && label_bytes > 0) {
    strncpy(label, e->prof_label[index].c_str(), label_bytes - 1);
    label[label_bytes - 1] = 0;
  }
  return ms;
}

size_t fsuae_engine_device_bytes(const fsuae_engine* e) { return e ? e->device_bytes : 0; }
int64_t fsuae_engine_last_launch_count(const fsuae_engine* e) { return e ? e->launches : 0; }
const char* fsuae_engine_variant(const fsuae_engine* e) { return e ? e->variant.c_str() : ""; }

const char* fsuae_last_error(const fsuae_engine* e) {
  if (e) return e->last_error.c_str();
  static thread_local std::string copy;
  std::lock_guard<std::mutex> lk(g_create_mu);
  copy = g_create_error;
  return copy.c_str();
}

}  // extern "C"
