// bf16 tcgen05 build (placeholder until the tensor-core path lands).
#include "engine.h"

namespace fsuae {
int bf16_create(fsuae_engine* e) { return set_error(e, FSUAE_ERR_UNSUPPORTED, "bf16 build not implemented yet"); }
void bf16_destroy(fsuae_engine*) {}
int bf16_enqueue_chunk(fsuae_engine* e, const void*, void*, int, int, int, uint32_t, cudaStream_t) {
  return set_error(e, FSUAE_ERR_UNSUPPORTED, "bf16 build not implemented yet");
}
}  // namespace fsuae
