// Tensor-core attention for bf16 (placeholder dispatch until the kernel lands in this file).
#include "common.cuh"

bool avj_attention_mma_supported(int dtype, int hd) { (void)dtype; (void)hd; return false; }
int avj_attention_fwd_mma(const void*, void*, float*, int, int, int, int, float, cudaStream_t) {
  avj_set_error("attention tensor-core kernel not built"); return 1;
}
int avj_attention_bwd_mma(const void*, const void*, const void*, const float*, void*, float*, int, int, int, int, float, cudaStream_t) {
  avj_set_error("attention tensor-core kernel not built"); return 1;
}
