// microbenchmark: MMA rate when every B slab (16 KB) comes through a TMA bulk-copy ring of NS stages from an
// L2-resident weight buffer (as in the edge kernel); A fixed in smem.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../arreau_b200/csrc/tc_common.cuh"
using namespace tc;

template <int NS, int CHUNK>
__global__ void __launch_bounds__(128, 1) k(long long* out, const uint8_t* w, int chunks_total, int nchunks_src) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  __shared__ uint64_t full[NS], empty[NS], done;
  __shared__ uint32_t tmem_s;
  if (threadIdx.x == 0) { for (int i = 0; i < NS; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); } mbar_init(&done, 1); fence_barrier_init(); }
  if (threadIdx.x >= 64 && threadIdx.x < 96) tmem_alloc(&tmem_s, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_s;
  uint8_t* W = smem + 65536;
  constexpr int SL = CHUNK / 16384;     // slabs per chunk
  if (threadIdx.x == 0) {
    uint32_t st = 0, par = 1;
    for (int c = 0; c < chunks_total; ++c) {
      mbar_wait(&empty[st], par);
      mbar_expect_tx(&full[st], CHUNK);
      bulk_g2s(W + st * CHUNK, w + (size_t)(c % nchunks_src) * CHUNK, CHUNK, &full[st]);
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
      umma_commit(&empty[st]);
      if (++st == NS) { st = 0; par ^= 1; }
    }
    umma_commit(&done);
    mbar_wait(&done, 0);
    const long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x >= 64 && threadIdx.x < 96) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred;
}
template <bool kAcc>
__device__ __forceinline__ void umma_elect(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t elected) {
  asm volatile(
      "{\n\t.reg .pred p, e;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %4};\n\tmov.b64 db, {%2, %4};\n\t"
      "setp.ne.u32 e, %5, 0;\n\t"
      "setp.ne.u32 p, %6, 0;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(kDescHiSw128), "r"(elected), "r"(kAcc ? 1u : 0u)
      : "memory");
}
// warp-uniform issue: the whole warp runs the loop, one elected lane issues
template <int NS, int CHUNK>
__global__ void __launch_bounds__(128, 1) k2(long long* out, const uint8_t* w, int chunks_total, int nchunks_src) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  __shared__ uint64_t full[NS], empty[NS], done;
  __shared__ uint32_t tmem_s;
  if (threadIdx.x == 0) { for (int i = 0; i < NS; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); } mbar_init(&done, 1); fence_barrier_init(); }
  if (threadIdx.x >= 64 && threadIdx.x < 96) tmem_alloc(&tmem_s, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_s;
  uint8_t* W = smem + 65536;
  constexpr int SL = CHUNK / 16384;
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  if (threadIdx.x == 0) {
    uint32_t st = 0, par = 1;
    for (int c = 0; c < chunks_total; ++c) {
      mbar_wait(&empty[st], par);
      mbar_expect_tx(&full[st], CHUNK);
      bulk_g2s(W + st * CHUNK, w + (size_t)(c % nchunks_src) * CHUNK, CHUNK, &full[st]);
      if (++st == NS) { st = 0; par ^= 1; }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = umma_idesc_f16(128, 128);
    const uint32_t a_lo = umma_desc_lo(base), w_lo = umma_desc_lo(smem_u32(W));
    const uint32_t fulla = smem_u32(&full[0]), emptya = smem_u32(&empty[0]);
    uint32_t st = 0, par = 0;
    const uint32_t el = elect_one();
    const long long t0 = clock64();
    for (int c = 0; c < chunks_total; ++c) {
      mbar_wait_addr(fulla + 8 * st, par);
      for (int s = 0; s < SL; ++s) {
        const uint32_t b_lo = w_lo + (st * CHUNK + s * 16384) / 16, al = a_lo + ((c * SL + s) & 3) * 1024;
        const uint32_t d = tmem + (((c * SL + s) >> 2) & 1) * 256;
        if (((c * SL + s) & 3) == 0) umma_elect<false>(d, al, b_lo, idesc, el); else umma_elect<true>(d, al, b_lo, idesc, el);
        umma_elect<true>(d, al + 2, b_lo + 2, idesc, el);
        umma_elect<true>(d, al + 4, b_lo + 4, idesc, el);
        umma_elect<true>(d, al + 6, b_lo + 6, idesc, el);
      }
      if (el) umma_commit_addr(emptya + 8 * st);
      if (++st == NS) { st = 0; par ^= 1; }
    }
    if (el) umma_commit(&done);
    mbar_wait(&done, 0);
    const long long t1 = clock64();
    if (blockIdx.x == 0 && el) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x >= 64 && threadIdx.x < 96) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}
template <int NS, int CHUNK>
void run2(long long* d, const uint8_t* w) {
  const int smem = 65536 + NS * CHUNK + 2048;
  cudaFuncSetAttribute(k2<NS, CHUNK>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int total = 24 * 40 * 16384 / CHUNK, nsrc = 384 * 1024 / CHUNK;
  long long h;
  k2<NS, CHUNK><<<148, 128, smem>>>(d, w, total, nsrc); cudaDeviceSynchronize();
  k2<NS, CHUNK><<<148, 128, smem>>>(d, w, total, nsrc); cudaDeviceSynchronize();
  cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  printf("warp-uniform+elect: stages %2d x %2d KB: %.1f cycles/mma  (%s)\n", NS, CHUNK / 1024,
         (double)h / (total * (CHUNK / 16384) * 4), cudaGetErrorString(cudaGetLastError()));
}

template <int NS, int CHUNK>
void run(long long* d, const uint8_t* w) {
  const int smem = 65536 + NS * CHUNK + 2048;
  cudaFuncSetAttribute(k<NS, CHUNK>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int total = 24 * 40 * 16384 / CHUNK, nsrc = 384 * 1024 / CHUNK;
  long long h;
  for (int grid : {1, 148}) {
    k<NS, CHUNK><<<grid, 128, smem>>>(d, w, total, nsrc); cudaDeviceSynchronize();
    k<NS, CHUNK><<<grid, 128, smem>>>(d, w, total, nsrc); cudaDeviceSynchronize();
    cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("stages %2d x %2d KB (%3d KB in flight) grid %3d: %.1f cycles/mma  (%s)\n", NS, CHUNK / 1024, NS * CHUNK / 1024, grid,
           (double)h / (total * (CHUNK / 16384) * 4), cudaGetErrorString(cudaGetLastError()));
  }
}

int main() {
  long long* d; cudaMalloc(&d, 8);
  uint8_t* w; cudaMalloc(&w, 384 * 1024); cudaMemset(w, 0, 384 * 1024);
  run<6, 16384>(d, w); run<3, 32768>(d, w);
  run2<4, 16384>(d, w); run2<6, 16384>(d, w); run2<3, 32768>(d, w);
  return 0;
}
