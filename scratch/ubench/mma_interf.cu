// microbenchmark: tcgen05.mma rate (N=128, SS) of one issuing thread under interference from 16 other warps:
//   mode 0 none, 1 spinning on an mbarrier (try_wait), 2 st.shared.v4 stream, 3 tcgen05.ld stream, 4 FMA/MUFU math,
//   5 = per-slab mbarrier wait+commit pattern on the issuer (ring emulation, barriers pre-armed by helper warp)
#include <cstdio>
#include <cuda_runtime.h>
#include "../../arreau_b200/csrc/tc_common.cuh"
using namespace tc;

__global__ void __launch_bounds__(640, 1) k(long long* out, int rounds, int mode) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  __shared__ uint64_t bar, never;
  __shared__ uint32_t tmem_s;
  __shared__ volatile int stop;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&never, 1); stop = 0; fence_barrier_init(); }
  if (threadIdx.x >= 64 && threadIdx.x < 96) tmem_alloc(&tmem_s, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 1 && lane == 0) {
    constexpr uint32_t idesc = umma_idesc_f16(128, 128);
    const long long t0 = clock64();
    for (int r = 0; r < rounds; ++r) {
      const uint32_t d = tmem + (r & 1) * 384;
      for (int s = 0; s < 4; ++s) {
        const uint64_t ad = umma_desc_sw128(base + s * 16384), bd = umma_desc_sw128(base + 65536 + s * 16384);
        for (int kk = 0; kk < 4; ++kk) umma_f16(d, ad + kk * 2, bd + kk * 2, idesc, (s | kk) ? 1u : 0u);
      }
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    const long long t1 = clock64();
    stop = 1;
    if (blockIdx.x == 0) out[0] = t1 - t0;
  } else if (warp >= 4) {
    const int q = warp & 3;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    float acc = 0.f;
    if (mode == 1) {
      while (!stop) {
        uint32_t done;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(&never)), "r"(0) : "memory");
        acc += done;
      }
    } else if (mode == 2) {
      uint8_t* p = smem + 131072 + (threadIdx.x - 128) * 16;
      while (!stop) {
#pragma unroll
        for (int i = 0; i < 8; ++i) *reinterpret_cast<uint4*>(p + i * 8192) = make_uint4(1, 2, 3, 4);
      }
    } else if (mode == 3) {
      while (!stop) {
        float v[32];
        tmem_ld32(tmem + lane_addr + 128 + ((warp - 4) >> 2) * 32, v);
        acc += v[0] + v[31];
      }
    } else if (mode == 4) {
      float x = lane * 0.01f;
      while (!stop) {
#pragma unroll
        for (int i = 0; i < 16; ++i) { x = fmaf(x, 1.0001f, 0.5f); x = tanhf(x) + 0.1f; }
        acc += x;
      }
    }
    if (acc == 12345.678f) out[1] = 1;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x >= 64 && threadIdx.x < 96) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

int main() {
  long long* d; cudaMalloc(&d, 16);
  const int smem = 220 * 1024;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const char* names[] = {"none", "mbarrier spin", "st.shared stream", "tcgen05.ld stream", "fma/mufu math"};
  for (int mode = 0; mode < 5; ++mode) {
    const int rounds = 200;
    long long h;
    k<<<148, 640, smem>>>(d, rounds, mode); cudaDeviceSynchronize();
    k<<<148, 640, smem>>>(d, rounds, mode); cudaDeviceSynchronize();
    cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("mode %d (%s): %.1f cycles/mma (%s)\n", mode, names[mode], (double)h / (rounds * 16), cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
