// microbenchmark: cycles per tcgen05.mma (kind::f16, SS mode, cta_group::1) for N = 128 / 256, M = 128, K = 16
#include <cstdio>
#include <cuda_runtime.h>
#include "../../arreau_b200/csrc/tc_common.cuh"
using namespace tc;

template <int N>
__global__ void __launch_bounds__(128, 1) mma_rate(long long* out, int rounds, int distinct, int extra) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ uint64_t bar;
  __shared__ uint64_t bar2[4];
  __shared__ uint32_t tmem_s;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); for (int i = 0; i < 4; ++i) mbar_init(&bar2[i], 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc(&tmem_s, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_s;
  if (threadIdx.x == 0) {
    constexpr uint32_t idesc = umma_idesc_f16(128, N);
    uint32_t ph[4] = {0, 0, 0, 0};
    if (extra & 4) for (int i = 0; i < 4; ++i) { mbar_arrive(&bar2[i]); }
    const long long t0 = clock64();
    for (int r = 0; r < rounds; ++r) {
      const uint32_t d = tmem + (r & 1) * 256;
      for (int s = 0; s < 4; ++s) {          // 4 slabs of 4 k-steps = K 256
        const uint32_t a = base + (distinct ? s * 16384 : 0), b = base + 65536 + (distinct ? s * 32768 : 0);
        if (extra & 1) tc_fence_after();
        if (extra & 4) { mbar_wait(&bar2[s], ph[s]); }
        const uint64_t ad = umma_desc_sw128(a), bd = umma_desc_sw128(b);
        for (int k = 0; k < 4; ++k) umma_f16(d, ad + k * 2, bd + k * 2, idesc, (s | k) ? 1u : 0u);
        if (extra & 2) umma_commit(&bar2[s]);
        if (extra & 4) { umma_commit(&bar2[s]); ph[s] ^= 1; }
      }
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    const long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

int main() {
  long long* d; cudaMalloc(&d, 8);
  const int smem = 200 * 1024;
  cudaFuncSetAttribute(mma_rate<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const char* names[] = {"plain", "tcgen05.fence::after per slab", "commit per slab", "fence + commit per slab", "mbar wait + commit per slab (ring emulation: wait for own previous commit)", "ring emu + fence"};
  int extras[] = {0, 1, 2, 3, 4, 5};
  for (int e = 0; e < 6; ++e) {
    const int rounds = 200;
    long long h;
    mma_rate<128><<<148, 128, smem>>>(d, rounds, 1, extras[e]); cudaDeviceSynchronize();
    mma_rate<128><<<148, 128, smem>>>(d, rounds, 1, extras[e]); cudaDeviceSynchronize();
    cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("%-70s: %.1f cycles/mma  (%s)\n", names[e], (double)h / (rounds * 16), cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
