// microbenchmark: as mma_ring.cu, but a cluster of 2 CTAs shares every weight chunk through TMA multicast
// (CTA r issues the chunks with c % 2 == r to both CTAs), tcgen05.commit multicast releases the stage in both.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../arreau_b200/csrc/tc_common.cuh"
using namespace tc;

__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void bulk_g2s_mc(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar, uint16_t mask) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(
                   smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"(mask)
               : "memory");
}

template <int NS, int CHUNK>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) k(long long* out, const uint8_t* w, int chunks_total, int nchunks_src) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  __shared__ uint64_t full[NS], empty[NS], done;
  __shared__ uint32_t tmem_s;
  const uint32_t rank = cluster_rank();
  if (threadIdx.x == 0) { for (int i = 0; i < NS; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 2); } mbar_init(&done, 1); fence_barrier_init(); }
  if (threadIdx.x >= 64 && threadIdx.x < 96) tmem_alloc(&tmem_s, 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync();
  tc_fence_after();
  const uint32_t tmem = tmem_s;
  uint8_t* W = smem + 65536;
  constexpr int SL = CHUNK / 16384;
  if (threadIdx.x == 0) {
    uint32_t st = 0, par = 1;
    for (int c = 0; c < chunks_total; ++c) {
      mbar_wait(&empty[st], par);
      mbar_expect_tx(&full[st], CHUNK);
      if ((uint32_t)(c & 1) == rank) bulk_g2s_mc(W + st * CHUNK, w + (size_t)(c % nchunks_src) * CHUNK, CHUNK, &full[st], 3);
      if (++st == NS) { st = 0; par ^= 1; }
    }
  } else if (threadIdx.x == 32) {
    constexpr uint32_t idesc = umma_idesc_f16(128, 128);
    const uint32_t a_lo = umma_desc_lo(base), w_lo = umma_desc_lo(smem_u32(W));
    uint32_t st = 0, par = 0;
    const long long t0 = clock64();
    for (int c = 0; c < chunks_total; ++c) {
      mbar_wait(&full[st], par);
      for (int s = 0; s < SL; ++s) {
        const uint32_t b_lo = w_lo + (st * CHUNK + s * 16384) / 16, al = a_lo + ((c * SL + s) & 3) * 1024;
        const uint32_t d = tmem + (((c * SL + s) >> 2) & 1) * 256;
        umma_f16_lo_p(d, al, b_lo, idesc, ((c * SL + s) & 3) ? 1u : 0u);
        umma_f16_lo<true>(d, al + 2, b_lo + 2, idesc);
        umma_f16_lo<true>(d, al + 4, b_lo + 4, idesc);
        umma_f16_lo<true>(d, al + 6, b_lo + 6, idesc);
      }
      umma_commit_mc(&empty[st], 3);
      if (++st == NS) { st = 0; par ^= 1; }
    }
    umma_commit(&done);
    mbar_wait(&done, 0);
    const long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync();
  if (threadIdx.x >= 64 && threadIdx.x < 96) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

template <int NS, int CHUNK>
void run(long long* d, const uint8_t* w) {
  const int smem = 65536 + NS * CHUNK + 2048;
  cudaFuncSetAttribute(k<NS, CHUNK>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int total = 24 * 40 * 16384 / CHUNK, nsrc = 384 * 1024 / CHUNK;
  long long h;
  for (int grid : {2, 148}) {
    k<NS, CHUNK><<<grid, 128, smem>>>(d, w, total, nsrc); cudaDeviceSynchronize();
    k<NS, CHUNK><<<grid, 128, smem>>>(d, w, total, nsrc); cudaDeviceSynchronize();
    cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("multicast x2: stages %2d x %2d KB grid %3d: %.1f cycles/mma  (%s)\n", NS, CHUNK / 1024, grid,
           (double)h / (total * (CHUNK / 16384) * 4), cudaGetErrorString(cudaGetLastError()));
  }
}

int main() {
  long long* d; cudaMalloc(&d, 8);
  uint8_t* w; cudaMalloc(&w, 384 * 1024); cudaMemset(w, 0, 384 * 1024);
  run<4, 16384>(d, w); run<6, 16384>(d, w); run<8, 16384>(d, w);
  run<3, 32768>(d, w); run<4, 32768>(d, w);
  return 0;
}
