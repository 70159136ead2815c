// Input-side kernels next to the hot path: what turns an image into the framebuffer the enhancer sees.
//   * grid quantisation to the Amiga colour depths + pixel-mode replication
//     (dataset_generator/quantize.py:464-473, 512-521 with dithering_method='none';
//      dataset_generator/util.py:318-350 post_apply_resolution_style)
//   * a counter-based generator of synthetic RGB444 framebuffers in the four pixel modes (README.md:7-10,
//     rgb444_flat_image_generator.py:28-30: 4 -> 8 bit expansion r*16 + r), so benchmark streams are born on the device.
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
