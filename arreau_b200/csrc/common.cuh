// Shared device helpers for the arreau_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/arreau_b200.h"

// number of kernels launched through the C ABI (bench.py reports it as gpu_launches)
extern long long g_arreau_launches;

// Model dimensions the kernels are specialised on (reference defaults, main_diffusion.py:88-120).
constexpr int kO = 16;     // orientations
constexpr int kC = 128;    // hidden channels
constexpr int kD = 256;    // kernel-basis width
constexpr int kW = 512;    // ConvNext widening (4 * C)
constexpr int kL = 5;      // layers
constexpr int kMono = 83;  // distinct monomials of degree <= 3 in 6 variables (258 as written)
constexpr int kMonoPad = 96;

#define CUDA_LAUNCH_CHECK()                               \
  do {                                                    \
    cudaError_t e__ = cudaGetLastError();                 \
    if (e__ != cudaSuccess) return (int)e__;              \
    ++g_arreau_launches;                                  \
  } while (0)

__device__ __forceinline__ float gelu_erf(float x) {
  // torch.nn.GELU() exact form (ponita/models/ponita.py:61)
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---- cp.async (LDGSTS) helpers -------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// 6-variable monomials of degree <= 3 in the canonical order used by the folded first basis layer
// (arreau_b200/weights.py builds the same table): all i, then all i<=j, then all i<=j<=k.
template <typename T>
__device__ __forceinline__ void monomials83(const T (&v)[6], T* out, int stride) {
  int n = 0;
#pragma unroll
  for (int i = 0; i < 6; ++i) out[(n++) * stride] = v[i];
#pragma unroll
  for (int i = 0; i < 6; ++i)
#pragma unroll
    for (int j = i; j < 6; ++j) out[(n++) * stride] = v[i] * v[j];
#pragma unroll
  for (int i = 0; i < 6; ++i)
#pragma unroll
    for (int j = i; j < 6; ++j)
#pragma unroll
      for (int k = j; k < 6; ++k) out[(n++) * stride] = v[i] * v[j] * v[k];
}

// Edge invariants of one (edge, orientation) pair, evaluated in fp64 like the reference and rounded
// once: [dir.ori, |dir - (dir.ori) ori|, dist, cos(dir,a), cos(dir,b), cos(dir,c)]
// (ponita/geometry/invariants.py:17-22, ponita/transforms/invariants.py:81-87).
__device__ __forceinline__ void edge_invariants(const double* __restrict__ dir3, double dist,
                                                const double* __restrict__ lat9, const float* __restrict__ ori3,
                                                float (&attr)[6]) {
  const double dx = dir3[0], dy = dir3[1], dz = dir3[2];
  const double ox = (double)ori3[0], oy = (double)ori3[1], oz = (double)ori3[2];
  const double i1 = dx * ox + dy * oy + dz * oz;
  const double px = dx - i1 * ox, py = dy - i1 * oy, pz = dz - i1 * oz;
  attr[0] = (float)i1;
  attr[1] = (float)sqrt(px * px + py * py + pz * pz);
  attr[2] = (float)dist;
  const double dd = dx * dx + dy * dy + dz * dz;
#pragma unroll
  for (int m = 0; m < 3; ++m) {
    const double ax = lat9[3 * m], ay = lat9[3 * m + 1], az = lat9[3 * m + 2];
    const double w12 = dx * ax + dy * ay + dz * az;
    const double w2 = ax * ax + ay * ay + az * az;
    // torch.nn.CosineSimilarity: x.y / sqrt(clamp(|x|^2 |y|^2, eps^2)), eps = 1e-8
    attr[3 + m] = (float)(w12 / sqrt(fmax(dd * w2, 1e-16)));
  }
}

// PolynomialCutoff(p = 6) * [x < r_max]   (ponita/utils/windowing.py:21-29)
__device__ __forceinline__ float cutoff_window(double x, double r_max) {
  if (!(x < r_max)) return 0.0f;
  const double u = x / r_max;
  const double u2 = u * u, u6 = u2 * u2 * u2;
  return (float)(1.0 - 28.0 * u6 + 48.0 * u6 * u - 21.0 * u6 * u2);
}
