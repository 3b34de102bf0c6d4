// tcgen05 encoder self-attention (placeholder until the fused kernel lands; fails loudly).
#include "kernels.cuh"

namespace wlk {

void enc_attention_tcgen05(const void*, int, int, int, void*, cudaStream_t, int) {
    WLK_CHECK(false, "enc_attention_tcgen05 is not built in this revision; use WLK_BACKEND_SIMT for attention");
}

}  // namespace wlk
