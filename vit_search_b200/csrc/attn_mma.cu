// Tensor-core attention (bf16 mma.sync, flash-style) -- placeholder until the kernel lands: reports "unsupported"
// so the dispatcher in attn.cu uses the fp32-math kernel.
#include "common.cuh"

namespace vsx {
bool attn_mma_supported(int, int) { return false; }
int attn_fwd_mma(const void*, void*, float*, int, int, int, int, int, float, cudaStream_t) {
  set_error("attn_fwd_mma: not built");
  return VSX_ERR_ARG;
}
int attn_bwd_mma(const void*, const void*, const void*, const float*, void*, int, int, int, int, int, float, cudaStream_t) {
  set_error("attn_bwd_mma: not built");
  return VSX_ERR_ARG;
}
}  // namespace vsx
