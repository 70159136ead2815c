// Input-side kernels next to the hot path: what turns an image into the framebuffer the enhancer sees.
//   * grid quantisation to the Amiga colour depths + pixel-mode replication
//     (dataset_generator/quantize.py:464-473, 512-521 with dithering_method='none';
//      dataset_generator/util.py:318-350 post_apply_resolution_style)
//   * a counter-based generator of synthetic RGB444 framebuffers in the four pixel modes (README.md:7-10,
//     rgb444_flat_image_generator.py:28-30: 4 -> 8 bit expansion r*16 + r), so benchmark streams are born on the device.
//   * the palette dithers of dataset_generator/quantize.py:137-331: nearest palette colour, checkerboard, ordered (Bayer 2/4/8)
// Byte / integer work: results are bit-exact against oracle/enhancer_oracle.py (and the reference's own functions,
// tests/golden/quantize.npz).
#include <cuda_runtime.h>
#include <stdint.h>

#include "fsuae_enhancer.h"

namespace {

__device__ __forceinline__ uint32_t grid_floor(uint32_t v, int cs, int ch) {
  switch (cs) {
    case FSUAE_CS_RGB444: return v & 0xF0u;                              // floor(v / 16) * 16
    case FSUAE_CS_RGB555: return v & 0xF8u;
    case FSUAE_CS_RGB565: return ch == 1 ? (v & 0xFCu) : (v & 0xF8u);
    case FSUAE_CS_RGB666: return v & 0xFCu;
    default: return v;                                                   // RGB888
  }
}

// out[f][y][x] = quantise(in[f][y / sy][x / sx]); PIL's NEAREST up-scaling by an integer factor is exactly x / sx
__global__ void quantize_upscale_kernel(const unsigned char* __restrict__ in, unsigned char* __restrict__ out, int n_frames, int h,
                                        int w, int in_channels, int cs, int sy, int sx, int expand17) {
  const int H = h * sy, W = w * sx;
  const size_t total = (size_t)n_frames * H * W;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % W);
    const size_t t = i / W;
    const int y = (int)(t % H), f = (int)(t / H);
    const unsigned char* p = in + (((size_t)f * h + y / sy) * w + x / sx) * in_channels;
    uint32_t px = 0xFF000000u;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      uint32_t v = grid_floor(p[c], cs, c);
      if (expand17 && cs == FSUAE_CS_RGB444) v |= v >> 4;                // (v >> 4) * 17: what the emulator's framebuffer holds
      px |= v << (8 * c);
    }
    reinterpret_cast<uint32_t*>(out)[i] = px;
  }
}

__device__ __forceinline__ uint64_t mix64(uint64_t z) {                 // splitmix64 finaliser
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

// frame g = first_frame + f uses pixel mode g & 3: lores (2x2 blocks), lores_laced (1 row x 2 columns), hires (2 x 1), hires_laced (1 x 1)
__global__ void synth_rgb444_kernel(unsigned char* __restrict__ out, int n_frames, int H, int W, unsigned long long seed,
                                    long long first_frame, int expand17) {
  const size_t total = (size_t)n_frames * H * W;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % W);
    const size_t t = i / W;
    const int y = (int)(t % H), f = (int)(t / H);
    const unsigned long long g = (unsigned long long)(first_frame + f);
    const int mode = (int)(g & 3ull);
    const int sy = (mode == 0 || mode == 2) ? 2 : 1, sx = (mode == 0 || mode == 1) ? 2 : 1;
    const unsigned long long cell = (unsigned long long)(y / sy) * 65536ull + (unsigned long long)(x / sx);
    const uint64_t r = mix64(mix64(seed + 0x9E3779B97F4A7C15ull * (g + 1ull)) ^ cell);
    uint32_t px = 0xFF000000u;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const uint32_t q = (uint32_t)(r >> (20 * c + 4)) & 15u;
      px |= (expand17 ? q * 17u : q * 16u) << (8 * c);
    }
    reinterpret_cast<uint32_t*>(out)[i] = px;
  }
}

// ---- palette dithers of the dataset generator (dataset_generator/quantize.py:137-331), thread = pixel ----------------------
// The reference works in float64 on integer-valued data: squared distances are exact integers here.  Closest colour =
// first index of the minimum (strict '<', quantize.py:185-190); second closest = first index of the minimum over the rest
// (:206-215).  Checkerboard (:137-229): exact match -> closest, else closest on (x + y) even, second closest on odd.
// Ordered / Bayer (:232-331): exact match -> closest; else order the two by Rec.709 luminance, interpolation fraction of the
// pixel's luminance between them (clamped, 0 if they are closer than 1e-6), second one iff fraction > matrix[y % m][x % m] / m^2.
// The luminance arithmetic is done in double with explicitly unfused multiplies and adds, in the reference's order, so the
// comparison against the threshold gives the same bit as numba's float64 code.
__constant__ int c_bayer2[4] = {0, 2, 3, 1};
__constant__ int c_bayer4[16] = {0, 8, 2, 10, 12, 4, 14, 6, 3, 11, 1, 9, 15, 7, 13, 5};
__constant__ int c_bayer8[64] = {0, 32, 8, 40, 2, 34, 10, 42, 48, 16, 56, 24, 50, 18, 58, 26, 12, 44, 4, 36, 14, 46, 6, 38,
                                 60, 28, 52, 20, 62, 30, 54, 22, 3, 35, 11, 43, 1, 33, 9, 41, 51, 19, 59, 27, 49, 17, 57, 25,
                                 15, 47, 7, 39, 13, 45, 5, 37, 63, 31, 55, 23, 61, 29, 53, 21};

__device__ __forceinline__ double lum709(double r, double g, double b) {
  return __dadd_rn(__dadd_rn(__dmul_rn(r, 0.2126), __dmul_rn(g, 0.7152)), __dmul_rn(b, 0.0722));
}

__global__ void dither_palette_kernel(const unsigned char* __restrict__ in, unsigned char* __restrict__ out, int n_frames, int H, int W,
                                      int in_channels, const unsigned char* __restrict__ palette, int n_colors, int method) {
  extern __shared__ uint32_t s_pal[];        // r | g << 8 | b << 16
  for (int i = threadIdx.x; i < n_colors; i += blockDim.x)
    s_pal[i] = (uint32_t)palette[3 * i] | ((uint32_t)palette[3 * i + 1] << 8) | ((uint32_t)palette[3 * i + 2] << 16);
  __syncthreads();
  const size_t total = (size_t)n_frames * H * W;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % W);
    const int y = (int)((i / W) % H);
    const unsigned char* p = in + i * in_channels;
    const int r = p[0], g = p[1], b = p[2];
    uint32_t px;
    if (n_colors == 0) {
      px = 0u;
    } else if (n_colors == 1) {
      px = s_pal[0];
    } else {
      int d1 = 0x7FFFFFFF, i1 = 0;
      for (int k = 0; k < n_colors; ++k) {
        const uint32_t c = s_pal[k];
        const int dr = r - (int)(c & 0xFF), dg = g - (int)((c >> 8) & 0xFF), db = b - (int)((c >> 16) & 0xFF);
        const int d = dr * dr + dg * dg + db * db;
        if (d < d1) { d1 = d; i1 = k; }
      }
      int chosen = i1;
      if (method != FSUAE_DITHER_NONE && d1 != 0) {
        int d2 = 0x7FFFFFFF, i2 = i1;
        for (int k = 0; k < n_colors; ++k) {
          if (k == i1) continue;
          const uint32_t c = s_pal[k];
          const int dr = r - (int)(c & 0xFF), dg = g - (int)((c >> 8) & 0xFF), db = b - (int)((c >> 16) & 0xFF);
          const int d = dr * dr + dg * dg + db * db;
          if (d < d2) { d2 = d; i2 = k; }
        }
        if (method == FSUAE_DITHER_CHECKERBOARD) {
          chosen = ((x + y) & 1) == 0 ? i1 : i2;
        } else {
          const uint32_t ca = s_pal[i1], cb = s_pal[i2];
          const double lp = lum709((double)r, (double)g, (double)b);
          double l1 = lum709((double)(ca & 0xFF), (double)((ca >> 8) & 0xFF), (double)((ca >> 16) & 0xFF));
          double l2 = lum709((double)(cb & 0xFF), (double)((cb >> 8) & 0xFF), (double)((cb >> 16) & 0xFF));
          int dark = i1, light = i2;
          if (l1 > l2) { const double t = l1; l1 = l2; l2 = t; dark = i2; light = i1; }
          double frac = 0.0;
          if (!(fabs(__dsub_rn(l2, l1)) < 1e-6)) frac = __ddiv_rn(__dsub_rn(lp, l1), __dsub_rn(l2, l1));
          frac = fmax(0.0, fmin(1.0, frac));
          const int msz = method == FSUAE_DITHER_BAYER2 ? 2 : (method == FSUAE_DITHER_BAYER4 ? 4 : 8);
          const int mv = method == FSUAE_DITHER_BAYER2 ? c_bayer2[(y % 2) * 2 + x % 2]
                         : (method == FSUAE_DITHER_BAYER4 ? c_bayer4[(y % 4) * 4 + x % 4] : c_bayer8[(y % 8) * 8 + x % 8]);
          const double thr = __ddiv_rn((double)mv, (double)(msz * msz));
          chosen = frac > thr ? light : dark;
        }
      }
      px = s_pal[chosen];
    }
    reinterpret_cast<uint32_t*>(out)[i] = px | 0xFF000000u;
  }
}

int launch_dims(size_t total) { return (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16); }

}  // namespace

extern "C" {

int fsuae_quantize_frames(const void* in_dev, void* out_dev_rgba, int n_frames, int in_height, int in_width, int in_channels,
                          int color_space, int sy, int sx, int expand17, void* cuda_stream) {
  if (!in_dev || !out_dev_rgba || n_frames < 0 || in_height < 1 || in_width < 1 || (in_channels != 3 && in_channels != 4) ||
      color_space < FSUAE_CS_RGB888 || color_space > FSUAE_CS_RGB666 || sy < 1 || sx < 1)
    return FSUAE_ERR_INVALID;
  const size_t total = (size_t)n_frames * in_height * sy * in_width * sx;
  if (total == 0) return FSUAE_OK;
  quantize_upscale_kernel<<<launch_dims(total), 256, 0, (cudaStream_t)cuda_stream>>>(
      (const unsigned char*)in_dev, (unsigned char*)out_dev_rgba, n_frames, in_height, in_width, in_channels, color_space, sy, sx, expand17);
  return cudaGetLastError() == cudaSuccess ? FSUAE_OK : FSUAE_ERR_CUDA;
}

int fsuae_dither_frames(const void* in_dev, void* out_dev_rgba, int n_frames, int height, int width, int in_channels,
                        const void* palette_dev, int n_colors, int method, void* cuda_stream) {
  if (!in_dev || !out_dev_rgba || n_frames < 0 || height < 1 || width < 1 || (in_channels != 3 && in_channels != 4) || n_colors < 0 ||
      n_colors > 4096 || (n_colors > 0 && !palette_dev) || method < FSUAE_DITHER_NONE || method > FSUAE_DITHER_BAYER8)
    return FSUAE_ERR_INVALID;
  const size_t total = (size_t)n_frames * height * width;
  if (total == 0) return FSUAE_OK;
  dither_palette_kernel<<<launch_dims(total), 256, (size_t)(n_colors > 0 ? n_colors : 1) * 4, (cudaStream_t)cuda_stream>>>(
      (const unsigned char*)in_dev, (unsigned char*)out_dev_rgba, n_frames, height, width, in_channels, (const unsigned char*)palette_dev,
      n_colors, method);
  return cudaGetLastError() == cudaSuccess ? FSUAE_OK : FSUAE_ERR_CUDA;
}

int fsuae_synth_rgb444_frames(void* out_dev_rgba, int n_frames, int height, int width, uint64_t seed, int64_t first_frame,
                              int expand17, void* cuda_stream) {
  if (!out_dev_rgba || n_frames < 0 || height < 1 || width < 1 || width > 65535 || first_frame < 0) return FSUAE_ERR_INVALID;
  const size_t total = (size_t)n_frames * height * width;
  if (total == 0) return FSUAE_OK;
  synth_rgb444_kernel<<<launch_dims(total), 256, 0, (cudaStream_t)cuda_stream>>>((unsigned char*)out_dev_rgba, n_frames, height, width,
                                                                                 seed, first_frame, expand17);
  return cudaGetLastError() == cudaSuccess ? FSUAE_OK : FSUAE_ERR_CUDA;
}

}  // extern "C"
