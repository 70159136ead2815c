/*
 * fsuae_enhancer.h -- C ABI of the B200-native FS-UAE image-enhancer inference engine.
 *
 * This is the drop-in boundary for the ONE hot path of cminnoy/fs_uae_image_enhancer_project:
 * the per-frame forward of the upscaling network plus its uint8/gamma framebuffer glue.
 * Plain pointers and sizes only; no torch / C++ types cross it.  Every entry point returns an
 * int status (FSUAE_OK == 0), never throws, and -- except fsuae_engine_run_host and
 * fsuae_engine_create/destroy -- never synchronises the device or allocates memory.
 *
 * Reference interfaces each entry point replaces (paths relative to the reference repo):
 *
 *   fsuae_engine_create      model/model_pix_shuffle.py:20-182 Model.__init__ + :304-314 get_model
 *                            (and model_conv3.py:20-55/:206-211, model_conv5.py:22-68/:157-162)
 *                            followed by load_state_dict / .to(device) / .half()
 *                            -- and, at deploy time, OrtCreateSession on the exported graph
 *                            (convertion_tools/convert_raw_to_png_using_final_model.py:66)
 *   fsuae_engine_enqueue     Model.forward(x): model_pix_shuffle.py:227-298,
 *                            model_conv3.py:102-155, model_conv5.py:114-151; with the
 *                            FSUAE_FMT_U8_NHWC4 formats it is the whole exported graph
 *                            input_rgba_chunky -> output_rgba_uint8_chunky
 *                            (convertion_tools/torch2onnx.py:184-768), i.e. OrtRun
 *                            (convert_raw_to_png_using_final_model.py:82)
 *   fsuae_engine_run_host    the same call as made from the emulator side with HOST framebuffers
 *   fsuae_engine_submit_host / fsuae_engine_wait_host   streaming (asynchronous) form of run_host
 *                            (README.md:21-24: upload, upscale, copy back)
 *   fsuae_engine_destroy     Python GC of the module / OrtReleaseSession
 *   fsuae_quantize_frames    dataset_generator/quantize.py:464-473, 512-521 (grid quantisation, no dithering) +
 *                            dataset_generator/util.py:318-350 (pixel-mode replication)
 *   fsuae_dither_frames      dataset_generator/quantize.py:137-331 (checkerboard and ordered/Bayer palette dithers, numba) and
 *                            :529-537 (nearest palette colour)
 *   fsuae_synth_rgb444_frames  the synthetic benchmark stream of SURVEY 8d (README.md:7-10 pixel modes,
 *                            rgb444_flat_image_generator.py:28-30 expansion), generated on the device
 *   fsuae_last_error         Python exception text (ValueError at model_conv3.py:109-110,
 *                            model_pix_shuffle.py:81-83, activations.py:123-127)
 *
 * The network is handed over as a flat descriptor + one float32 parameter blob, produced on the
 * host from a reference state_dict (fs_uae_image_enhancer_project_b200/descriptor.py): conv
 * weights in PyTorch order [Cout][Cin][3][3] with BatchNorm already folded
 * (model_conv3.py:41-52 eval semantics), biases, and activation parameters.
 */
#ifndef FSUAE_ENHANCER_H
#define FSUAE_ENHANCER_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define FSUAE_API __attribute__((visibility("default")))
#else
#define FSUAE_API
#endif

#define FSUAE_ABI_VERSION 2
#define FSUAE_MAX_LAYERS 16
#define FSUAE_MAX_ACTS 4 /* activation slots before / after the skip add (reference uses <= 2) */
#define FSUAE_MAX_CHUNK_FRAMES 1024 /* upper bound of fsuae_engine_create's max_chunk_frames */

/* status codes */
enum {
  FSUAE_OK = 0,
  FSUAE_ERR_INVALID = 1,     /* bad argument / descriptor (reference: ValueError) */
  FSUAE_ERR_UNSUPPORTED = 2, /* valid network the selected precision build cannot run; never a CPU fallback */
  FSUAE_ERR_CUDA = 3,        /* CUDA runtime error; text in fsuae_last_error */
  FSUAE_ERR_NO_DEVICE = 4
};

/* activation op-codes: the registry of model/activations.py:69-95 */
enum {
  FSUAE_ACT_IDENTITY = 0,
  FSUAE_ACT_RELU = 1,
  FSUAE_ACT_RELU6 = 2,
  FSUAE_ACT_TANH = 3,
  FSUAE_ACT_SIGMOID = 4,
  FSUAE_ACT_SILU = 5,        /* 'silu' and 'swish' */
  FSUAE_ACT_MISH = 6,
  FSUAE_ACT_GELU = 7,        /* exact erf form */
  FSUAE_ACT_ELU = 8,         /* p0 = alpha (scalar) */
  FSUAE_ACT_SOFTPLUS = 9,    /* p0 = beta, p1 = threshold (scalars) */
  FSUAE_ACT_LEAKY_RELU = 10, /* p0 = negative_slope (scalar) */
  FSUAE_ACT_PRELU = 11,      /* p0 = slope[n0], n0 in {1, C} */
  FSUAE_ACT_SCALED_TANH = 12,
  FSUAE_ACT_TELU = 13,
  FSUAE_ACT_SINLU = 14,      /* p0 = a, p1 = b (scalars) */
  FSUAE_ACT_BIASED_RELU = 15,  /* p0 = bias[n0] */
  FSUAE_ACT_BIASED_PRELU = 16, /* p0 = bias[n0], p1 = slope[n1] */
  FSUAE_ACT_SOFTMAX = 17,      /* over channels (dim=1) */
  FSUAE_ACT_LOG_SOFTMAX = 18,
  FSUAE_ACT_COUNT = 19
};

typedef struct fsuae_act_desc {
  int32_t op;
  int32_t n0, p0_off; /* count and float offset into the blob of parameter 0 */
  int32_t n1, p1_off; /* count and float offset of parameter 1 */
} fsuae_act_desc;

/* One k x k / stride 1 / zero-pad (k-1)/2 convolution (k = 3, 5 or 7: the `layer{i}_kernel_size` constructor arguments of
 * model/model_pix_shuffle.py:21-64, 108-115 and the `kernel_size` of residual_feature_block.py:6; a 1x1 convolution is
 * handed over as the centre tap of a 3x3 one) with its fused epilogue:
 *   y = post( skip + pre( conv(cat[src0, src1]) + bias ) )
 * Buffer ids: 0 = the network input after the head stage, i = output of layers[i-1]. */
typedef struct fsuae_layer_desc {
  int32_t cin0, cin1; /* channels taken from src0 / src1 (cin1 == 0: no concat) */
  int32_t cout;
  int32_t src0, src1;
  int32_t skip_src;   /* -1: no skip add */
  int32_t w_off;      /* float offset: weights [cout][cin0+cin1][ksize][ksize] */
  int32_t b_off;      /* float offset: bias [cout]; -1: none */
  int32_t n_pre, n_post;
  fsuae_act_desc pre[FSUAE_MAX_ACTS];
  fsuae_act_desc post[FSUAE_MAX_ACTS];
  int32_t ksize;      /* 3, 5 or 7 */
  int32_t reserved;   /* 0 */
} fsuae_layer_desc;

/* head: how buffer 0 is derived from the frame */
enum {
  FSUAE_HEAD_PLAIN = 0,      /* 3 channels at full resolution (conv3, conv5) */
  FSUAE_HEAD_UNSHUFFLE2 = 1, /* PixelUnshuffle(2): 12 channels at half resolution, channel c*4+dy*2+dx */
  FSUAE_HEAD_FEATURES = 2    /* buffer 0 is the input as it is: a float feature map [B,in_channels,H,W] (the building blocks
                                of model/model_residual_unet.py, e.g. residual_feature_block.py:44-55) */
};
/* tail: how the last layer's output becomes the frame */
enum {
  FSUAE_TAIL_PLAIN = 0,                 /* 3 channels as they are (conv5) */
  FSUAE_TAIL_SHUFFLE2_RESIDUAL_RELU = 1,/* PixelShuffle(2), + input, ReLU (model_pix_shuffle.py:293-296) */
  FSUAE_TAIL_SCALE255_ALPHA = 2,        /* x255 and alpha=255.0 appended (model_conv3.py:145-153) */
  FSUAE_TAIL_FEATURES = 3               /* the last layer's output as it is: a float feature map [B,cout,H,W] */
};

typedef struct fsuae_net_desc {
  int32_t abi_version; /* FSUAE_ABI_VERSION */
  int32_t n_layers;
  int32_t head, tail;
  int32_t in_channels; /* FSUAE_HEAD_FEATURES: channels of the input feature map; otherwise 0 */
  int32_t reserved;    /* 0 */
  fsuae_layer_desc layers[FSUAE_MAX_LAYERS];
} fsuae_net_desc;

/* arithmetic builds */
enum {
  FSUAE_PREC_FP32 = 0, /* fp32 FMA kernels, accurate libm activations */
  FSUAE_PREC_BF16 = 1, /* bf16 operands on tcgen05 tensor cores, fp32 accumulate + epilogue */
  FSUAE_PREC_FP16 = 2  /* fp16 operands on the same tensor-core kernels (tcgen05 kind::f16, same rate, 8x finer mantissa):
                          the precision the reference deploys (convertion_tools/torch2onnx.py:58 .half(), :358-412 Cast) */
};

/* frame formats at the boundary (all row-major, contiguous) */
enum {
  FSUAE_FMT_F32_NCHW3 = 0, /* float [B,3,H,W] in [0,1]; pix_shuffle: linear light (train.py:61) */
  FSUAE_FMT_U8_NHWC4 = 1,  /* uint8 [B,H,W,4] RGBA framebuffer (torch2onnx.py:225-232, 717-756) */
  FSUAE_FMT_U8_NCHW4 = 2,  /* uint8 [B,4,H,W] planar RGBA, input only (model_conv3.py:109-113) */
  FSUAE_FMT_F32_NCHW4 = 3, /* float [B,4,H,W], output only, with FSUAE_TAIL_SCALE255_ALPHA */
  FSUAE_FMT_F32_NCHW = 4   /* float [B,C,H,W] feature map: C = in_channels on the input side (FSUAE_HEAD_FEATURES), the last
                              layer's cout on the output side (FSUAE_TAIL_FEATURES) */
};

/* flags */
#define FSUAE_FLAG_GAMMA_IN  1u /* uint8 input: (u8/255)**2.2 (gamma.py:13-15; torch2onnx.py:391-412) */
#define FSUAE_FLAG_GAMMA_OUT 2u /* uint8 output: clamp(y**(1/2.2),0,1)*255, truncate (gamma.py:31-33; train.py:70) */
#define FSUAE_FLAG_CROP16    4u /* run the net on columns [16,W) and emit 16 black columns (torch2onnx.py:299-355, 634-674) */

typedef struct fsuae_engine fsuae_engine;

FSUAE_API int fsuae_abi_version(void);

/* Build an engine for frames of height x width on CUDA device `device`.
 * `blob` holds `blob_floats` float32 parameters addressed by the descriptor's offsets.
 * `max_chunk_frames` bounds the frames processed per internal pass (workspace is sized for it);
 * enqueue accepts any n_frames and loops.  All device memory is allocated here. */
FSUAE_API int fsuae_engine_create(const fsuae_net_desc* desc, const float* blob, size_t blob_floats,
                        int device, int precision, int height, int width, int max_chunk_frames,
                        fsuae_engine** out);
/* Same, from an engine file written by fs_uae_image_enhancer_project_b200/export.py
 * (bytes: "FSUAEENG" | uint32 abi | uint32 blob_floats | fsuae_net_desc | float32 blob): the
 * counterpart of loading the exported .onnx at deploy time
 * (convertion_tools/convert_raw_to_png_using_final_model.py:66). */
FSUAE_API int fsuae_engine_create_from_file(const char* path, int device, int precision, int height, int width,
                                  int max_chunk_frames, fsuae_engine** out);
FSUAE_API int fsuae_engine_destroy(fsuae_engine* e);

/* Asynchronous: enqueue the forward of n_frames frames on `cuda_stream` (a cudaStream_t).
 * `in_dev` / `out_dev` are device pointers in the given formats.
 * Ownership / threading (reference: caller owns the tensors, forward is stateless): the caller owns
 * in_dev / out_dev; the engine owns its workspace, which every call re-uses -- so calls on ONE engine must be ordered
 * (one host thread at a time, and either one stream or streams the caller orders with events); the host-buffer calls use
 * the engine's internal streams and must not run concurrently with fsuae_engine_enqueue on the same engine.  Use one
 * engine per GPU and per concurrent stream; engines are independent. */
FSUAE_API int fsuae_engine_enqueue(fsuae_engine* e, const void* in_dev, void* out_dev, int n_frames,
                         int in_fmt, int out_fmt, uint32_t flags, void* cuda_stream);

/* Synchronous end-to-end call with HOST buffers (pinned memory recommended): chunks the frames,
 * overlaps H2D copy, compute and D2H copy on internal streams, returns when out_host is complete. */
FSUAE_API int fsuae_engine_run_host(fsuae_engine* e, const void* in_host, void* out_host, int n_frames,
                          int in_fmt, int out_fmt, uint32_t flags);

/* Streaming form of the same call (no reference counterpart: ONNX Runtime's Run is synchronous; this is what removes
 * the per-frame copy-back stall README.md:23-24 describes).  submit queues upload, forward and download of n_frames
 * frames on the engine's internal streams and returns at once; consecutive submissions overlap each other.  The
 * buffers must stay valid -- and must be pinned for the copies to be asynchronous -- until fsuae_engine_wait_host
 * returns, which blocks until every submitted frame has arrived in its out_host.
 * fsuae_engine_run_host == submit + wait. */
FSUAE_API int fsuae_engine_submit_host(fsuae_engine* e, const void* in_host, void* out_host, int n_frames,
                             int in_fmt, int out_fmt, uint32_t flags);
FSUAE_API int fsuae_engine_wait_host(fsuae_engine* e);

/* ---- input side: what turns an image into the framebuffer the enhancer sees (byte work, bit-exact) ---------------
 * Colour depths of dataset_generator/quantize.py:464-473, 512-521 (dithering_method='none': floor onto the grid). */
enum {
  FSUAE_CS_RGB888 = 0,
  FSUAE_CS_RGB444 = 1,
  FSUAE_CS_RGB555 = 2,
  FSUAE_CS_RGB565 = 3,
  FSUAE_CS_RGB666 = 4
};
/* out[f][y][x] = quantise(in[f][y / sy][x / sx]), alpha = 255: grid quantisation followed by the pixel-mode
 * replication of dataset_generator/util.py:318-350 (post_apply_resolution_style: lores sy=sx=2, lores_laced sy=1 sx=2,
 * hires sy=2 sx=1, hires_laced 1x1; PIL NEAREST by an integer factor).  in: uint8 [n][h][w][in_channels] (3 or 4),
 * out: uint8 RGBA [n][h*sy][w*sx][4], both on the current device.  expand17: RGB444 values become (v >> 4) * 17, the
 * 4 -> 8 bit expansion a real framebuffer holds (rgb444_flat_image_generator.py:28-30), instead of floor(v/16)*16. */
FSUAE_API int fsuae_quantize_frames(const void* in_dev, void* out_dev_rgba, int n_frames, int in_height, int in_width,
                          int in_channels, int color_space, int sy, int sx, int expand17, void* cuda_stream);
/* Palette dithers of the dataset generator (dataset_generator/quantize.py:137-331, the numba kernels
 * _apply_checkerboard_dithering_numba_optimized / _apply_ordered_dithering_numba_optimized, and the palette mapping of
 * :529-537): every pixel is replaced by a colour of `palette_dev` (uint8 [n_colors][3], device memory, n_colors <= 4096;
 * how the palette is chosen -- k-means / median cut / octree -- stays on the host, it is offline data preparation).
 * in: uint8 [n][h][w][in_channels] (3 or 4), out: uint8 RGBA [n][h][w][4], alpha 255.  Bit-exact against the reference. */
enum {
  FSUAE_DITHER_NONE = 0,         /* nearest palette colour (first index of the minimum squared distance) */
  FSUAE_DITHER_CHECKERBOARD = 1, /* closest / second closest colour alternate on (x + y) parity; exact matches stay */
  FSUAE_DITHER_BAYER2 = 2,       /* ordered dither between the two closest colours by Rec.709 luminance, 2x2 Bayer matrix */
  FSUAE_DITHER_BAYER4 = 3,
  FSUAE_DITHER_BAYER8 = 4
};
FSUAE_API int fsuae_dither_frames(const void* in_dev, void* out_dev_rgba, int n_frames, int height, int width, int in_channels,
                        const void* palette_dev, int n_colors, int method, void* cuda_stream);
/* Synthetic RGB444 framebuffers generated on the device (SURVEY 8d workload): frame g = first_frame + i uses pixel mode
 * g & 3 (lores, lores_laced, hires, hires_laced); every sy x sx cell holds one counter-hashed 12-bit colour
 * (splitmix64 of seed, g, cell), stored as q * 17 (expand17) or q * 16; alpha = 255.  Deterministic in (seed, g). */
FSUAE_API int fsuae_synth_rgb444_frames(void* out_dev_rgba, int n_frames, int height, int width, uint64_t seed,
                              int64_t first_frame, int expand17, void* cuda_stream);

/* Device bytes held by the engine (parameters + workspace + staging). */
FSUAE_API size_t fsuae_engine_device_bytes(const fsuae_engine* e);
/* Number of kernel launches the last enqueue / run_host issued (bench.py's gpu_launches). */
FSUAE_API int64_t fsuae_engine_last_launch_count(const fsuae_engine* e);
/* Name of the kernel variant the engine selected, e.g. "fp32_fma" / "bf16_tcgen05" / "fp16_tcgen05"; experiment
 * switches found in the environment when the engine was created (FSUAE_R3, FSUAE_NO_PAIRS, ... -- read once, never per
 * launch) are listed behind it in brackets. */
FSUAE_API const char* fsuae_engine_variant(const fsuae_engine* e);

/* Optional per-kernel timing (measurement aid for bench.py's roofline line, no reference counterpart): when
 * enabled, every kernel launch of an enqueue is bracketed by CUDA events on the launching stream.
 * fsuae_engine_kernel_time returns the device time in ms of launch `index` of the last enqueue (after the caller
 * synchronised the stream) and copies a short kernel label; returns < 0 past the end. */
FSUAE_API int fsuae_engine_set_profiling(fsuae_engine* e, int enabled);
FSUAE_API float fsuae_engine_kernel_time(fsuae_engine* e, int index, char* label, int label_bytes);

/* Text of the last error on this engine (or of the last failed create when e == NULL). */
FSUAE_API const char* fsuae_last_error(const fsuae_engine* e);

#ifdef __cplusplus
}
#endif
#endif /* FSUAE_ENHANCER_H */
