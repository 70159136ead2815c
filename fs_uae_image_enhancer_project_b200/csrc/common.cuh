// Shared device/host helpers for the FS-UAE enhancer engine (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "fsuae_enhancer.h"

namespace fsuae {

// ---- device-side view of one activation slot (pointers into the device parameter blob) ----
struct ActDev {
  int op;
  int n0, n1;
  const float* p0;
  const float* p1;
};

struct EpiDev {
  int n_pre, n_post;
  ActDev pre[FSUAE_MAX_ACTS];
  ActDev post[FSUAE_MAX_ACTS];
};

// Accurate (libm-grade) activation math for the fp32 build: matches PyTorch's CPU kernels to a
// few ulp (definitions: reference model/activations.py:6-65 and the torch.nn modules its
// registry names, :69-95).
__device__ __forceinline__ float act_apply_accurate(const ActDev& a, float x, int ch) {
  switch (a.op) {
    case FSUAE_ACT_IDENTITY: return x;
    case FSUAE_ACT_RELU: return fmaxf(x, 0.f);
    case FSUAE_ACT_RELU6: return fminf(fmaxf(x, 0.f), 6.f);
    case FSUAE_ACT_TANH: return tanhf(x);
    case FSUAE_ACT_SIGMOID: return 1.f / (1.f + expf(-x));
    case FSUAE_ACT_SILU: return x / (1.f + expf(-x));
    case FSUAE_ACT_MISH: return x * tanhf(log1pf(expf(x)));
    case FSUAE_ACT_GELU: return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f));
    case FSUAE_ACT_ELU: return x > 0.f ? x : a.p0[0] * expm1f(x);
    case FSUAE_ACT_SOFTPLUS: {
      float beta = a.p0[0], thr = a.p1[0];
      float bx = x * beta;
      return bx > thr ? x : log1pf(expf(bx)) / beta;
    }
    case FSUAE_ACT_LEAKY_RELU: return x >= 0.f ? x : a.p0[0] * x;
    case FSUAE_ACT_PRELU: return x >= 0.f ? x : a.p0[a.n0 == 1 ? 0 : ch] * x;
    case FSUAE_ACT_SCALED_TANH: return (tanhf(x) + 1.f) * 0.5f;
    case FSUAE_ACT_TELU: return x * tanhf(expf(x));
    case FSUAE_ACT_SINLU: return (1.f / (1.f + expf(-x))) * (x + a.p0[0] * sinf(a.p1[0] * x));
    case FSUAE_ACT_BIASED_RELU: return fmaxf(x - a.p0[a.n0 == 1 ? 0 : ch], 0.f);
    case FSUAE_ACT_BIASED_PRELU: {
      float y = x - a.p0[a.n0 == 1 ? 0 : ch];
      return y >= 0.f ? y : a.p1[a.n1 == 1 ? 0 : ch] * y;
    }
    default: return x;  // softmax handled by dedicated kernels
  }
}

inline bool act_is_softmax(int op) { return op == FSUAE_ACT_SOFTMAX || op == FSUAE_ACT_LOG_SOFTMAX; }

}  // namespace fsuae
