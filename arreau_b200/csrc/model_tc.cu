// tcgen05 (bf16 operands, fp32 accumulate in TMEM) variants of the two GEMM-shaped kernels.
#include "common.cuh"

extern "C" int arreau_edge_kernels_bf16(const double*, const double*, const double*, const int32_t*, const int32_t*,
                                        const int32_t*, int64_t, const float*, const void*, const void*, const float*,
                                        const void*, double, void*, void*) {
  return ARREAU_ERR_UNSUPPORTED;
}
extern "C" int arreau_convnext_mlp_bf16(const void*, const void*, const float*, const void*, const float*,
                                        const float*, int64_t, float*, void*) {
  return ARREAU_ERR_UNSUPPORTED;
}
