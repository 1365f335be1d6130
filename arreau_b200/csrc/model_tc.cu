// tcgen05 (fp16 operands, fp32 accumulation in TMEM) variants of the two GEMM-shaped kernels of the step.
//
//   convnext_mlp_tc_kernel   K6   h += layer_scale * (W2 gelu(W1 y + b1) + b2)          convnext.py:26-32
//   edge_kernels_tc_kernel   K3+K4a invariants -> monomials -> basis MLP -> window -> 5 kernel projections
//
// Both are persistent (one CTA per SM, 128-row tiles) and warp specialised (640 threads):
//   warp 0   producer: cp.async.bulk (TMA engine) of pre-swizzled weight / activation tiles into an mbarrier ring
//   warp 1   MMA issuer: one thread issues tcgen05.mma (M = 128, N = 128, K = 16) and tcgen05.commit
//   warp 2   TMEM allocation / release
//   warps 4..19  epilogue (16 warps: 4 per TMEM lane quarter, each a quarter of the columns): tcgen05.ld ->
//                bias / GELU / window in registers -> fp16 operand tile of the next GEMM written back to shared
//                memory in the UMMA layout (the intermediate never leaves the SM), or the final result to HBM.
#include "common.cuh"
#include "tc_common.cuh"

namespace {

using namespace tc;

constexpr int kThreads = 640;
constexpr int kEpiWarp0 = 4;
constexpr int kEpiWarps = 16;
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kTileM = 128;
constexpr uint32_t kIdesc128 = umma_idesc_f16(128, 128);

__device__ long long* g_tc_prof = nullptr;   // debug: per-phase clock64 stamps of CTA 0 (scratch/prof_tc.py)
#ifdef ARREAU_TC_PROFILE
#define TC_STAMP(slot)                                                                     \
  do {                                                                                      \
    if (prof && it < 6) prof[(it * 2 + prof_role) * 16 + (slot)] = clock64();               \
  } while (0)
#else
#define TC_STAMP(slot) do { (void)prof; } while (0)
#endif

// issue the UMMA_K = 16 steps of one 64-wide K slab: D[128 x 128] (+)= A_slab[128 x 64] * B_slab[128 x 64]^T
__device__ __forceinline__ void mma_slab(uint32_t tmem_d, uint32_t a_slab, uint32_t b_slab, int ksteps, bool accumulate_first) {
  const uint64_t ad = umma_desc_sw128(a_slab), bd = umma_desc_sw128(b_slab);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (k < ksteps) umma_f16(tmem_d, ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), kIdesc128, (accumulate_first || k > 0) ? 1u : 0u);
  }
}

// 32 accumulator columns -> (+ bias) -> GELU (* scale) -> fp16 -> the four 16-byte chunks chunk0..chunk0+3 of
// operand row m (row base pointer `row`, chunk positions XOR-swizzled by the row index)
template <bool kBias, bool kScale>
__device__ __forceinline__ void gelu_store32(const float (&v)[32], const float* __restrict__ bias_s, float scale,
                                             uint8_t* row, int chunk0, int m) {
  const __half2 sc = __float2half2_rn(scale);
#pragma unroll
  for (int cc = 0; cc < 4; ++cc) {
    float x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = v[cc * 8 + i];
    if constexpr (kBias) {
      const float4 b0 = *reinterpret_cast<const float4*>(bias_s + cc * 8);       // warp-uniform: smem broadcast
      const float4 b1 = *reinterpret_cast<const float4*>(bias_s + cc * 8 + 4);
      x[0] += b0.x; x[1] += b0.y; x[2] += b0.z; x[3] += b0.w;
      x[4] += b1.x; x[5] += b1.y; x[6] += b1.z; x[7] += b1.w;
    }
    uint4 pk;
    if constexpr (kScale) {
      pk.x = gelu2_scaled_f16(x[0], x[1], sc);
      pk.y = gelu2_scaled_f16(x[2], x[3], sc);
      pk.z = gelu2_scaled_f16(x[4], x[5], sc);
      pk.w = gelu2_scaled_f16(x[6], x[7], sc);
    } else {
      pk.x = gelu2_f16(x[0], x[1]);
      pk.y = gelu2_f16(x[2], x[3]);
      pk.z = gelu2_f16(x[4], x[5]);
      pk.w = gelu2_f16(x[6], x[7]);
    }
    *reinterpret_cast<uint4*>(row + (((chunk0 + cc) ^ (m & 7)) << 4)) = pk;
  }
}

// 32 accumulator columns -> + bias -> GELU * scale -> 16 packed fp16 registers
__device__ __forceinline__ void gelu_scaled_pack32(const float (&v)[32], const float* __restrict__ bias_s, float scale,
                                                   uint32_t (&pk)[16]) {
  const __half2 sc = __float2half2_rn(scale);
#pragma unroll
  for (int cc = 0; cc < 4; ++cc) {
    const float4 b0 = *reinterpret_cast<const float4*>(bias_s + cc * 8);
    const float4 b1 = *reinterpret_cast<const float4*>(bias_s + cc * 8 + 4);
    pk[cc * 4 + 0] = gelu2_scaled_f16(v[cc * 8 + 0] + b0.x, v[cc * 8 + 1] + b0.y, sc);
    pk[cc * 4 + 1] = gelu2_scaled_f16(v[cc * 8 + 2] + b0.z, v[cc * 8 + 3] + b0.w, sc);
    pk[cc * 4 + 2] = gelu2_scaled_f16(v[cc * 8 + 4] + b1.x, v[cc * 8 + 5] + b1.y, sc);
    pk[cc * 4 + 3] = gelu2_scaled_f16(v[cc * 8 + 6] + b1.z, v[cc * 8 + 7] + b1.w, sc);
  }
}

// 32 accumulator columns -> + bias -> GELU -> four packed fp16 chunks (registers only)
__device__ __forceinline__ void gelu_pack32(const float (&v)[32], const float* __restrict__ bias_s, uint4 (&pk)[4]) {
#pragma unroll
  for (int cc = 0; cc < 4; ++cc) {
    const float4 b0 = *reinterpret_cast<const float4*>(bias_s + cc * 8);
    const float4 b1 = *reinterpret_cast<const float4*>(bias_s + cc * 8 + 4);
    pk[cc].x = gelu2_f16(v[cc * 8 + 0] + b0.x, v[cc * 8 + 1] + b0.y);
    pk[cc].y = gelu2_f16(v[cc * 8 + 2] + b0.z, v[cc * 8 + 3] + b0.w);
    pk[cc].z = gelu2_f16(v[cc * 8 + 4] + b1.x, v[cc * 8 + 5] + b1.y);
    pk[cc].w = gelu2_f16(v[cc * 8 + 6] + b1.z, v[cc * 8 + 7] + b1.w);
  }
}

// =================================================================================================
// K6  ConvNext channel MLP
// =================================================================================================
namespace mlp {
constexpr int kTileBytes = 32768;     // [128 rows x 128 K] fp16 = 2 slabs
constexpr int kWStages = 3;
constexpr int kTilesBytes = (2 + 2 + kWStages) * kTileBytes;
constexpr int kSmemBytes = 232448;    // everything (tiles, bias, barriers) is carved from the dynamic window
constexpr int kChunksPerTile = 8;     // W1_0, W1_1, W2_0, W1_2, W2_1, W1_3, W2_2, W2_3
constexpr uint32_t kHCol = 384;       // TMEM: D1[2] = [0,256), D2 = [256,384), hidden-slice operand H[2] = [384,512) (64 columns each)

struct Bars {
  uint64_t a_full[2], a_empty[2], w_full[kWStages], w_empty[kWStages];
  uint64_t d1_full[2], d1_empty[2], h_full[2], h_empty[2], d2_full, d2_empty;
};
}  // namespace mlp

__global__ void __launch_bounds__(kThreads, 1)
convnext_mlp_tc_kernel(const uint8_t* __restrict__ y_img, const uint8_t* __restrict__ w_img, const float* __restrict__ b1,
                       const float* __restrict__ b2, const float* __restrict__ layer_scale, long long rows,
                       float* __restrict__ h, const float* __restrict__ ori, float* __restrict__ pool_out,
                       const float* __restrict__ pool_wz, int pool_cols) {
  using namespace mlp;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  uint8_t* const A0 = smem;                       // A[ab] = A0 + ab * kTileBytes
  uint8_t* const H0 = smem + 2 * kTileBytes;      // H[b]  = H0 + b * kTileBytes
  uint8_t* const W = smem + 4 * kTileBytes;
  float* const s_b1 = reinterpret_cast<float*>(smem + kTilesBytes);                 // [kW]
  Bars& bars = *reinterpret_cast<Bars*>(smem + kTilesBytes + kW * sizeof(float));
  uint32_t& tmem_base_s = *reinterpret_cast<uint32_t*>(smem + kTilesBytes + kW * sizeof(float) + sizeof(Bars));
  float* const s_ori = reinterpret_cast<float*>(smem + kTilesBytes + kW * sizeof(float) + sizeof(Bars) + 16);   // [kO][3]
  if ((base - smem_u32(smem_raw)) + kTilesBytes + kW * sizeof(float) + sizeof(Bars) + 16 + kO * 3 * sizeof(float) >
      (uint32_t)kSmemBytes)
    __trap();
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // warp-uniform
  const long long tiles = (rows + kTileM - 1) / kTileM;

  for (int i = threadIdx.x; i < kW; i += kThreads) s_b1[i] = b1[i];
  if (pool_out && threadIdx.x < kO * 3) s_ori[threadIdx.x] = ori[threadIdx.x];
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars.a_full[i], 1); mbar_init(&bars.a_empty[i], 1);
      mbar_init(&bars.d1_full[i], 1); mbar_init(&bars.d1_empty[i], kEpiWarps);
      mbar_init(&bars.h_full[i], kEpiWarps); mbar_init(&bars.h_empty[i], 1);
    }
    for (int i = 0; i < kWStages; ++i) { mbar_init(&bars.w_full[i], 1); mbar_init(&bars.w_empty[i], 1); }
    mbar_init(&bars.d2_full, 1); mbar_init(&bars.d2_empty, kEpiWarps);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(&tmem_base_s, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;

  if (warp == 0 && lane == 0) {
    // ---------------- producer: y tiles and 32 KB weight chunks in the order the MMA warp consumes them.  The first
    // two GEMM1 slices of a tile are issued under the previous tile's tail (see the MMA warp), so per tile the chunk
    // order is W2_0 W1_2 W2_1 W1_3 W2_2 [W1_0 of the next tile] W2_3 [W1_1 of the next tile] ----------------
    uint32_t chunk = 0;
    auto push = [&](int c) {                       // c = chunk index in w_img (W1_0 W1_1 W2_0 W1_2 W2_1 W1_3 W2_2 W2_3)
      const int ws = chunk % kWStages;
      mbar_wait(&bars.w_empty[ws], ((chunk / kWStages) & 1) ^ 1);
      mbar_expect_tx(&bars.w_full[ws], kTileBytes);
      bulk_g2s(W + ws * kTileBytes, w_img + (size_t)c * kTileBytes, kTileBytes, &bars.w_full[ws]);
      ++chunk;
    };
    auto load_y = [&](long long tile, int ti) {
      const int ab = ti & 1;
      mbar_wait(&bars.a_empty[ab], ((ti >> 1) & 1) ^ 1);
      mbar_expect_tx(&bars.a_full[ab], kTileBytes);
      bulk_g2s(A0 + ab * kTileBytes, y_img + (size_t)tile * kTileBytes, kTileBytes, &bars.a_full[ab]);
    };
    if ((long long)blockIdx.x < tiles) {
      load_y(blockIdx.x, 0);
      push(0); push(1);
    }
    int it = 0;
    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
      const bool has_next = tile + gridDim.x < tiles;
      push(2); push(3); push(4); push(5); push(6);
      if (has_next) {
        load_y(tile + gridDim.x, it + 1);
        push(0);
      }
      push(7);
      if (has_next) push(1);
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer: the whole warp runs the loop uniformly, one elected lane issues
    // (see tc_common.cuh: a divergent single thread pays ~50 cycles per MMA for operand moves) ----------------
    const uint32_t el = elect_one();
    const uint32_t a_lo0 = umma_desc_lo(smem_u32(A0)), w_lo0 = umma_desc_lo(smem_u32(W));
    const uint32_t wfull = smem_u32(&bars.w_full[0]), wempty = smem_u32(&bars.w_empty[0]);
    constexpr uint32_t kSlabLo = 16384 >> 4, kTileLo = kTileBytes >> 4;
    uint32_t ws = 0, wpar = 0;
    int it = 0;
    long long* prof = (blockIdx.x == 0 && el) ? g_tc_prof : nullptr;
    constexpr int prof_role = 0;
    // one 32 KB weight tile (two K slabs): D (+)= X[:, 0:64] . W[0]^T + X[:, 64:128] . W[1]^T, then release it
    auto mma_tile = [&](uint32_t d, uint32_t x_lo, uint32_t accumulate) {
      const uint32_t w_lo = w_lo0 + ws * kTileLo;
      umma_slab_e<4>(d, x_lo, w_lo, kIdesc128, el, accumulate);
      umma_slab_e<4>(d, x_lo + kSlabLo, w_lo + kSlabLo, kIdesc128, el, 1u);
      umma_commit_e(wempty + 8 * ws, el);
      if (++ws == (uint32_t)kWStages) { ws = 0; wpar ^= 1; }
    };
    auto gemm1 = [&](int ti, int j) {        // tile number ti of this CTA: D1[j&1] = y_tile . W1_j^T
      const int b = j & 1, ab = ti & 1;
      const uint32_t use = (uint32_t)ti * 2 + (j >> 1);
      if (j == 0) mbar_wait(&bars.a_full[ab], (ti >> 1) & 1);
      mbar_wait_addr(wfull + 8 * ws, wpar);
      mbar_wait(&bars.d1_empty[b], (use & 1) ^ 1);
      tc_fence_after();
      mma_tile(tmem + b * 128, a_lo0 + ab * kTileLo, 0u);
      umma_commit_e(smem_u32(&bars.d1_full[b]), el);
      if (j == 3) umma_commit_e(smem_u32(&bars.a_empty[ab]), el);
      TC_STAMP(2 + j);
    };
    auto gemm2 = [&](int ti, int j) {        // D2 (+)= H[j&1] . W2_j^T
      const int b = j & 1;
      const uint32_t use = (uint32_t)ti * 2 + (j >> 1);
      mbar_wait_addr(wfull + 8 * ws, wpar);
      mbar_wait(&bars.h_full[b], use & 1);
      if (j == 0) mbar_wait(&bars.d2_empty, (ti & 1) ^ 1);
      tc_fence_after();
      {
        // the hidden slice is the A operand straight out of tensor memory (64 columns of packed fp16 per buffer)
        const uint32_t w_lo = w_lo0 + ws * kTileLo, a_t = tmem + kHCol + b * 64;
        umma_slab_ts_e<4>(tmem + 256, a_t, w_lo, kIdesc128, el, j > 0 ? 1u : 0u);
        umma_slab_ts_e<4>(tmem + 256, a_t + 32, w_lo + kSlabLo, kIdesc128, el, 1u);
        umma_commit_e(wempty + 8 * ws, el);
        if (++ws == (uint32_t)kWStages) { ws = 0; wpar ^= 1; }
      }
      umma_commit_e(smem_u32(&bars.h_empty[b]), el);
      if (j == 3) umma_commit_e(smem_u32(&bars.d2_full), el);
      TC_STAMP(6 + j);
    };
    // Issue order: the first two GEMM1 slices of tile i+1 go in around the last GEMM2 of tile i, so the epilogue warps
    // find slice 0 of the next tile ready while that last GEMM2 runs instead of waiting for it
    if ((long long)blockIdx.x < tiles) { gemm1(0, 0); gemm1(0, 1); }
    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
      const bool has_next = tile + gridDim.x < tiles;
      TC_STAMP(0);
      gemm2(it, 0); gemm1(it, 2); gemm2(it, 1); gemm1(it, 3); gemm2(it, 2);
      if (has_next) gemm1(it + 1, 0);
      gemm2(it, 3);
      if (has_next) gemm1(it + 1, 1);
    }
  } else if (warp >= kEpiWarp0) {
    // ---------------- epilogue: warp = (lane quarter q, column group cgi of 32 columns) ----------------
    const int q = warp & 3, cgi = (warp - kEpiWarp0) >> 2;
    const int m = q * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    int it = 0;
    const bool is_issuer = threadIdx.x == kEpiWarp0 * 32;
    long long* prof = (blockIdx.x == 0 && is_issuer) ? g_tc_prof : nullptr;
    constexpr int prof_role = 1;
    // GELU epilogue of hidden slice j of the CTA's tile number `ti`: D1[j & 1] -> + b1 -> GELU -> packed fp16 -> TMEM operand H[j & 1]
    auto gelu_slice = [&](int ti, int j) {
      const int b = j & 1;
      const uint32_t use = (uint32_t)ti * 2 + (j >> 1);
      mbar_wait(&bars.d1_full[b], use & 1);
      mbar_wait(&bars.h_empty[b], (use & 1) ^ 1);
      tc_fence_after();
      float v[32];
      tmem_ld32(tmem + lane_addr + b * 128 + cgi * 32, v);
      uint4 pk[4];
      gelu_pack32(v, s_b1 + j * 128 + cgi * 32, pk);
      {
        // hidden unit k = cgi*32 .. +31 of this 128-slice -> 16 columns of packed pairs in the TMEM operand H[b]
        const uint32_t* pr = reinterpret_cast<const uint32_t*>(pk);
        uint32_t regs[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) regs[i] = pr[i];
        tmem_st16(tmem + lane_addr + kHCol + b * 64 + cgi * 16, regs);
        tmem_wait_st();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&bars.d1_empty[b]);
        mbar_arrive(&bars.h_full[b]);
      }
    };
    // this warp's 32 output channels: lane i keeps b2 and layer_scale of channel cgi*32 + i (broadcast by shuffles in the
    // final epilogue: with 227 KB of shared memory there is no L1 left, every per-tile reload would be an L2 round trip)
    const float my_b2 = __ldg(b2 + cgi * 32 + lane), my_ls = __ldg(layer_scale + cgi * 32 + lane);
    if (blockIdx.x < tiles) gelu_slice(0, 0);
    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
      TC_STAMP(0);
      // slice 0 of this tile was done ahead (below / above): it fills the wait for the previous tile's last GEMM
      for (int j = 1; j < 4; ++j) {
        TC_STAMP(1 + 2 * j);
        gelu_slice(it, j);
        TC_STAMP(2 + 2 * j);
      }
      if (tile + gridDim.x < tiles) gelu_slice(it + 1, 0);
      TC_STAMP(9);
      mbar_wait(&bars.d2_full, it & 1);
      tc_fence_after();
      TC_STAMP(10);
      {
        // h += layer_scale * (D2 + b2): the update tile [128 rows x 128 ch] fp32 is 64 KB contiguous in HBM.  It is
        // staged in the two (now idle) H buffers and added into h by the TMA engine (cp.reduce.async.bulk .add.f32:
        // every element is added exactly once, so the result is deterministic) -- no uncoalesced read-modify-write
        // from the SM.  A thread owns 8 16-byte chunks of its row; it writes them in a lane-rotated order
        // (chunk (s + lane) & 7 at step s), which spreads a warp over all banks (4-way = optimal for 512 B).
        float v[32];
        tmem_ld32(tmem + lane_addr + 256 + cgi * 32, v);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars.d2_empty);       // accumulators are in registers: release D2 early
        if (it > 0) {
          // the staging area still holds the previous tile's update: its bulk reduce must have finished reading shared
          // memory, and with pooling the pool warps (barriers 3 / 4, one per half) must be done with it too
          if (is_issuer) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          if (!pool_out) {
            asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
          } else {
            asm volatile("bar.sync 3, %0;" ::"n"(kEpiThreads + 64) : "memory");
            asm volatile("bar.sync 4, %0;" ::"n"(kEpiThreads + 64) : "memory");
          }
        }
        float4 d[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float o4[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float bb = __shfl_sync(0xffffffffu, my_b2, 4 * i + k), ls = __shfl_sync(0xffffffffu, my_ls, 4 * i + k);
            o4[k] = ls * (v[4 * i + k] + bb);
          }
          d[i] = make_float4(o4[0], o4[1], o4[2], o4[3]);
        }
        // barrel-rotate the 8 chunks by r = lane & 7 so that register slot s holds chunk (s + r) & 7
        const int r = lane & 7;
#pragma unroll
        for (int sh = 4; sh >= 1; sh >>= 1) {
          if (r & sh) {
            float4 t[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) t[i] = d[(i + sh) & 7];
#pragma unroll
            for (int i = 0; i < 8; ++i) d[i] = t[i];
          }
        }
        uint8_t* srow = H0 + m * 512 + cgi * 128;
#pragma unroll
        for (int st = 0; st < 8; ++st) *reinterpret_cast<float4*>(srow + (((st + r) & 7) << 4)) = d[st];
        fence_proxy_async();
        if (pool_out) {                                  // staged tile complete -> pool warps
          __threadfence_block();
          asm volatile("bar.arrive 2, %0;" ::"n"(kEpiThreads + 64) : "memory");
        }
        asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
        if (is_issuer) {
          const long long row0 = tile * kTileM;
          const long long left = rows - row0;
          const int valid = (int)(left < kTileM ? left : kTileM);
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            const int n = valid - half * 64 < 64 ? valid - half * 64 : 64;
            if (n > 0)
              asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(
                               h + (size_t)(row0 + half * 64) * kC),
                           "r"(smem_u32(H0 + half * kTileBytes)), "r"((uint32_t)n * 512u)
                           : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");   // (possibly empty) group per half
          }
        }
      }
      TC_STAMP(11);
    }
    if (is_issuer) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // every update has been added into h
  } else if ((warp == 2 || warp == 3) && pool_out) {
    // ---------------- pooled read-out (see readout_pooled_kernel): the two otherwise idle warps pool the staged residual
    // update over the 16 orientation rows of each of the tile's 8 atoms.  Two phases, one per staging half (rows 0..63 =
    // atoms 0..3, then atoms 4..7), each released to the epilogue warps separately (barriers 3 and 4) so the next
    // tile's H writes are not held up; per phase warp 2 takes the first atom pair, warp 3 the second, lane = 2 channels.
    // Pool entry (include/arreau_b200.h): the orientation mean goes to [group of 16 atoms][C][16 atoms] (a tile is one
    // half of a group); the three vector-pooled parts only meet the score row of the read-out, so this warp contracts
    // them with its 64 channels of that row and leaves 3 partial sums per atom ([atom][2 d + half]) ----------------
    const float* stage = reinterpret_cast<const float*>(H0);
    constexpr float inv = 1.0f / kO;
    const int half = warp - 2;
    const int c2 = half * 64 + lane * 2;                     // this thread's channel pair
    const float wz0 = pool_wz[(size_t)c2 * pool_cols] * inv, wz1 = pool_wz[(size_t)(c2 + 1) * pool_cols] * inv;
    const long long groups = (rows / kO + 15) / 16;
    float* const partials = pool_out + (size_t)groups * kC * 16;
    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
      asm volatile("bar.sync 2, %0;" ::"n"(kEpiThreads + 64) : "memory");
      const bool more = tile + gridDim.x < tiles;            // the last tile has no successor waiting on barriers 3 / 4
      float2 mean[2][4];                                     // [phase][atom of the phase]: (channel c2, c2 + 1)
      float sv[32];                                          // [(phase * 4 + atom) * 3 + d], padded to 32 for the reduction
#pragma unroll
      for (int i = 24; i < 32; ++i) sv[i] = 0.f;
#pragma unroll
      for (int ph = 0; ph < 2; ++ph) {
        float2 acc[3][4];
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          mean[ph][a] = make_float2(0.f, 0.f);
#pragma unroll
          for (int d = 0; d < 3; ++d) acc[d][a] = make_float2(0.f, 0.f);
        }
#pragma unroll 2
        for (int o = 0; o < kO; ++o) {
          const float ox = s_ori[3 * o], oy = s_ori[3 * o + 1], oz = s_ori[3 * o + 2];
          const float2 dx = make_float2(ox, ox), dy = make_float2(oy, oy), dz = make_float2(oz, oz);
#pragma unroll
          for (int a = 0; a < 4; ++a) {
            const float2 v = *reinterpret_cast<const float2*>(stage + ((ph * 4 + a) * kO + o) * kC + c2);
            mean[ph][a].x += v.x; mean[ph][a].y += v.y;
            acc[0][a] = __ffma2_rn(dx, v, acc[0][a]);
            acc[1][a] = __ffma2_rn(dy, v, acc[1][a]);
            acc[2][a] = __ffma2_rn(dz, v, acc[2][a]);
          }
        }
        // done reading this staging half
        if (more) {
          if (ph == 0) asm volatile("bar.arrive 3, %0;" ::"n"(kEpiThreads + 64) : "memory");
          else asm volatile("bar.arrive 4, %0;" ::"n"(kEpiThreads + 64) : "memory");
        }
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int d = 0; d < 3; ++d) sv[(ph * 4 + a) * 3 + d] = fmaf(wz0, acc[d][a].x, wz1 * acc[d][a].y);
      }
      // sum the 24 partial contractions over the warp's lanes: after the five halving exchanges lane i holds value i
#pragma unroll
      for (int bit = 16, n = 32; bit > 0; bit >>= 1, n >>= 1) {
        const bool up = lane & bit;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          if (j < n / 2) {
            const float send = up ? sv[j] : sv[j + n / 2], keep = up ? sv[j + n / 2] : sv[j];
            sv[j] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
          }
        }
      }
      if (lane < 24) partials[(size_t)(tile * 8 + lane / 3) * 8 + 2 * (lane % 3) + half] = sv[0];
      // the tile's 8 atoms of one channel are one full 32-byte sector of the group's block
      float* const pg = pool_out + (size_t)(tile >> 1) * kC * 16 + (size_t)(tile & 1) * 8;
      float4* p0 = reinterpret_cast<float4*>(pg + (size_t)c2 * 16);
      float4* p1 = reinterpret_cast<float4*>(pg + (size_t)(c2 + 1) * 16);
      p0[0] = make_float4(mean[0][0].x * inv, mean[0][1].x * inv, mean[0][2].x * inv, mean[0][3].x * inv);
      p0[1] = make_float4(mean[1][0].x * inv, mean[1][1].x * inv, mean[1][2].x * inv, mean[1][3].x * inv);
      p1[0] = make_float4(mean[0][0].y * inv, mean[0][1].y * inv, mean[0][2].y * inv, mean[0][3].y * inv);
      p1[1] = make_float4(mean[1][0].y * inv, mean[1][1].y * inv, mean[1][2].y * inv, mean[1][3].y * inv);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// =================================================================================================
// K3 + K4a  edge pipeline
// =================================================================================================
namespace edge {
constexpr int kChunkBytes = 32768;      // ring unit: two K slabs [128 rows x 64 K] fp16 of one weight matrix
constexpr int kStages = 3;
constexpr int kStageBytes = 32768;      // output staging: two 16 KB halves
constexpr int kA2Bytes = 32768;         // doubles as the monomial tile A1
constexpr int kA3Bytes = 65536;
constexpr int kTilesBytes = kA2Bytes + kA3Bytes + kStageBytes + kStages * kChunkBytes;
constexpr int kSmemBytes = 232448;      // everything (tiles, bias, barriers) is carved from the dynamic window
constexpr int kChunksPerTile = 3 + 2 * kL;   // W1, W2 (n-half) x2, then two per Wk_l
constexpr int kEdgesPerTile = kTileM / kO;
constexpr int kGeo = 12;                // floats per edge: dir (3), dist, cos(dir, a/b/c) (3), window, valid, pad

struct Bars {
  uint64_t w_full[kStages], w_empty[kStages];
  uint64_t a1_full, a2_full, a3_full, d1_full, d2_full;
  uint64_t g_full[2], g_empty[2];     // per-edge geometry of a tile (warp 3 -> epilogue warps), double buffered
  uint64_t x_full[2], x_empty[2];     // TMEM buffers X0 (D1, D3 even layers) / X1 (D3 odd layers)
};

// monomial n of csrc/common.cuh monomials83 as compile-time index triples; 83 = constant 1 (bias), > 83 = 0
struct MonoIdx { int deg, a, b, c; };
__host__ __device__ constexpr MonoIdx mono_idx(int n) {
  if (n < 6) return {1, n, 0, 0};
  int m = 6;
  for (int i = 0; i < 6; ++i)
    for (int j = i; j < 6; ++j) {
      if (m == n) return {2, i, j, 0};
      ++m;
    }
  for (int i = 0; i < 6; ++i)
    for (int j = i; j < 6; ++j)
      for (int k = j; k < 6; ++k) {
        if (m == n) return {3, i, j, k};
        ++m;
      }
  return {n == kMono ? 0 : -1, 0, 0, 0};
}
template <int N>
__device__ __forceinline__ float mono_val(const float (&v)[6], float one) {
  constexpr MonoIdx mi = mono_idx(N);
  if constexpr (mi.deg == 1) return v[mi.a];
  else if constexpr (mi.deg == 2) return v[mi.a] * v[mi.b];
  else if constexpr (mi.deg == 3) return v[mi.a] * v[mi.b] * v[mi.c];
  else if constexpr (mi.deg == 0) return one;
  else return 0.f;
}
template <int C>
__device__ __forceinline__ uint4 mono_chunk(const float (&v)[6], float one) {
  uint4 pk;
  pk.x = pack_f16(mono_val<C * 8 + 0>(v, one), mono_val<C * 8 + 1>(v, one));
  pk.y = pack_f16(mono_val<C * 8 + 2>(v, one), mono_val<C * 8 + 3>(v, one));
  pk.z = pack_f16(mono_val<C * 8 + 4>(v, one), mono_val<C * 8 + 5>(v, one));
  pk.w = pack_f16(mono_val<C * 8 + 6>(v, one), mono_val<C * 8 + 7>(v, one));
  return pk;
}
// chunks 3*PART .. 3*PART+2 of monomial row m (12 chunks of 8 = 96 columns; slab = chunk / 8)
template <int PART>
__device__ __forceinline__ void store_mono_part(const float (&v)[6], float one, uint8_t* a1, int m) {
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    constexpr int kBase = PART * 3;
    const int c = kBase + i;
    uint4 pk = i == 0 ? mono_chunk<kBase>(v, one) : (i == 1 ? mono_chunk<kBase + 1>(v, one) : mono_chunk<kBase + 2>(v, one));
    *reinterpret_cast<uint4*>(a1 + (c >> 3) * 16384 + m * kRowBytes + (((c & 7) ^ (m & 7)) << 4)) = pk;
  }
}

// fp32 edge invariants for the fp16 path (the monomials are rounded to fp16 right after; the fp32 path and the
// graph keep fp64): [dir.ori, |dir - (dir.ori) ori|, dist, cos(dir,a), cos(dir,b), cos(dir,c)]
__device__ __forceinline__ void edge_invariants_f32(const double* __restrict__ dir3, double dist, const double* __restrict__ lat9,
                                                    const float* __restrict__ ori3, float (&attr)[6]) {
  const float dx = (float)dir3[0], dy = (float)dir3[1], dz = (float)dir3[2];
  const float ox = ori3[0], oy = ori3[1], oz = ori3[2];
  const float i1 = dx * ox + dy * oy + dz * oz;
  const float px = dx - i1 * ox, py = dy - i1 * oy, pz = dz - i1 * oz;
  attr[0] = i1;
  attr[1] = sqrtf(px * px + py * py + pz * pz);
  attr[2] = (float)dist;
  const float dd = dx * dx + dy * dy + dz * dz;
#pragma unroll
  for (int mm = 0; mm < 3; ++mm) {
    const float ax = (float)lat9[3 * mm], ay = (float)lat9[3 * mm + 1], az = (float)lat9[3 * mm + 2];
    const float w12 = dx * ax + dy * ay + dz * az, w2 = ax * ax + ay * ay + az * az;
    attr[3 + mm] = w12 * rsqrtf(fmaxf(dd * w2, 1e-16f));     // CosineSimilarity, eps = 1e-8
  }
}
}  // namespace edge

__global__ void __launch_bounds__(kThreads, 1)
edge_kernels_tc_kernel(const double* __restrict__ dir, const double* __restrict__ dist, const double* __restrict__ lattice,
                       const int32_t* __restrict__ crystal_of_atom, const int32_t* __restrict__ src,
                       const int32_t* __restrict__ num_edges_ptr, long long edge_capacity, const float* __restrict__ ori,
                       const uint8_t* __restrict__ w1_img, const uint8_t* __restrict__ w_img, const float* __restrict__ b2,
                       double radius, __half* __restrict__ kernels) {
  // Software pipeline over the CTA's tiles: while the tensor pipe runs the five kernel projections (GEMM3) of tile i
  // out of A3, the head of tile i+1 -- monomials, GEMM1, GELU, GEMM2 -- runs in the shadow, so that only the second
  // GELU epilogue (which has to overwrite A3) sits between two GEMM3 phases.
  //   shared:  A2: monomial tile A1, then the hidden layer | A3: kernel basis | S: output staging | weight ring
  //   TMEM:    X0 = [0,128) and X1 = [384,512): GEMM3 double buffer;  D2 = [128,384): GEMM1 (first half) then GEMM2
  //   MMA issue order per tile i:   L0 L1 G1(i+1) L2 L3 G2(i+1) L4   (ring order Wk0 Wk1 W1 Wk2 Wk3 W2 Wk4, 26 chunks)
  //   epilogue order per tile i:    gen(i+1) E0 E1 Q1(i+1) E2 E3 E4 Q2(i+1)
  using namespace edge;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  float* const s_b2 = reinterpret_cast<float*>(smem + kTilesBytes);                                   // [kD]
  float* const s_geo = s_b2 + kD;                                                                      // [2][8 edges][kGeo]
  Bars& bars = *reinterpret_cast<Bars*>(smem + kTilesBytes + (kD + 2 * kEdgesPerTile * kGeo) * sizeof(float));
  uint32_t& tmem_base_s =
      *reinterpret_cast<uint32_t*>(smem + kTilesBytes + (kD + 2 * kEdgesPerTile * kGeo) * sizeof(float) + sizeof(Bars));
  if ((base - smem_u32(smem_raw)) + kTilesBytes + (kD + 2 * kEdgesPerTile * kGeo) * sizeof(float) + sizeof(Bars) + 16 > (uint32_t)kSmemBytes)
    __trap();
  uint8_t* const A2 = smem;                   // A1 (monomials) aliases A2
  uint8_t* const A3 = A2 + kA2Bytes;
  uint8_t* const S = A3 + kA3Bytes;           // two 16 KB halves (rows 0..63 / 64..127) of one layer's output tile
  uint8_t* const W = S + kStageBytes;
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // warp-uniform
  long long E = *num_edges_ptr;
  if (E > edge_capacity) E = edge_capacity;
  const long long tiles = (E + kEdgesPerTile - 1) / kEdgesPerTile;
  const long long first = blockIdx.x, stride = gridDim.x;

  for (int i = threadIdx.x; i < kD; i += kThreads) s_b2[i] = b2[i];
  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) { mbar_init(&bars.w_full[i], 1); mbar_init(&bars.w_empty[i], 1); }
    mbar_init(&bars.a1_full, kEpiWarps); mbar_init(&bars.a2_full, kEpiWarps); mbar_init(&bars.a3_full, kEpiWarps);
    mbar_init(&bars.d1_full, 1); mbar_init(&bars.d2_full, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&bars.x_full[i], 1); mbar_init(&bars.x_empty[i], kEpiWarps); }
    for (int i = 0; i < 2; ++i) { mbar_init(&bars.g_full[i], 1); mbar_init(&bars.g_empty[i], kEpiWarps); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(&tmem_base_s, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;

  if (warp == 3) {
    // ---------------- geometry: the per-edge part of the invariants, up to two tiles ahead ----------------
    // (src -> crystal -> lattice, dir, dist are dependent global loads of ~3000 cycles; here they are off every
    // critical path).  fp32 like the rest of the fp16 path: the monomials are rounded to fp16 right after.
    uint32_t k = 0;
    for (long long tile = first; tile < tiles; tile += stride, ++k) {
      const int buf = k & 1;
      mbar_wait(&bars.g_empty[buf], ((k >> 1) & 1) ^ 1);
      if (lane < kEdgesPerTile) {
        const long long e = tile * kEdgesPerTile + lane;
        float* gp = s_geo + (buf * kEdgesPerTile + lane) * kGeo;
        float vals[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (e < E) {
          const double* lat9 = lattice + 9 * (size_t)crystal_of_atom[src[e]];
          const float dx = (float)dir[3 * e], dy = (float)dir[3 * e + 1], dz = (float)dir[3 * e + 2];
          const float dd = dx * dx + dy * dy + dz * dz;
          vals[0] = dx; vals[1] = dy; vals[2] = dz;
          vals[3] = (float)dist[e];
#pragma unroll
          for (int mm = 0; mm < 3; ++mm) {
            const float ax = (float)lat9[3 * mm], ay = (float)lat9[3 * mm + 1], az = (float)lat9[3 * mm + 2];
            const float w12 = dx * ax + dy * ay + dz * az, w2 = ax * ax + ay * ay + az * az;
            vals[4 + mm] = w12 * rsqrtf(fmaxf(dd * w2, 1e-16f));     // CosineSimilarity, eps = 1e-8
          }
          vals[7] = cutoff_window(dist[e], radius);
          vals[8] = 1.0f;
        }
#pragma unroll
        for (int i = 0; i < 9; ++i) gp[i] = vals[i];
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars.g_full[buf]);
    }
  } else if (warp == 0 && lane == 0) {
    // ---------------- producer: 32 KB weight chunks (two K slabs) in the order the MMA warp consumes them ----------------
    if (first < tiles) {
      uint32_t st = 0, par = 1;          // waits on w_empty start at parity 1 (a fresh barrier passes)
      auto push = [&](const uint8_t* src_chunk) {
        mbar_wait(&bars.w_empty[st], par);
        mbar_expect_tx(&bars.w_full[st], kChunkBytes);
        bulk_g2s(W + st * kChunkBytes, src_chunk, kChunkBytes, &bars.w_full[st]);
        if (++st == kStages) { st = 0; par ^= 1; }
      };
      auto push_w = [&](int c0, int n) { for (int c = c0; c < c0 + n; ++c) push(w_img + (size_t)c * kChunkBytes); };
      push(w1_img);
      push_w(0, 2);                                   // first tile: W1, W2
      for (long long tile = first; tile < tiles; tile += stride) {
        const bool has_next = tile + stride < tiles;
        push_w(2, 4);                                 // Wk_0, Wk_1
        if (has_next) push(w1_img);
        push_w(6, 4);                                 // Wk_2, Wk_3
        if (has_next) push_w(0, 2);                   // W2 of the next tile
        push_w(10, 2);                                // Wk_4
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer ----------------
    // The tensor pipe's issue queue is shallow: every cycle the issuer spends between two tcgen05.mma is an idle
    // tensor cycle.  The whole warp runs this loop uniformly (descriptors and barrier addresses stay in uniform
    // registers) and one elected lane issues; ring chunks are waited for in pairs, so that a barrier round covers
    // 8 MMAs = 512 tensor cycles (measured: the pipe's floor of 64 cycles per MMA, scratch/ubench/mma_ring.cu).
    if (first < tiles) {
      const uint32_t el = elect_one();
      const uint32_t a2_lo = umma_desc_lo(smem_u32(A2)), a3_lo = umma_desc_lo(smem_u32(A3)), w_lo = umma_desc_lo(smem_u32(W));
      const uint32_t wfull = smem_u32(&bars.w_full[0]), wempty = smem_u32(&bars.w_empty[0]);
      constexpr uint32_t kSlabLo = 16384 >> 4;          // descriptor step of one 16 KB slab / ring stage
      uint32_t st = 0, par = 0;
      // one ring chunk = two K slabs: D (+)= A[:, slab a] . chunk[0]^T + A[:, slab a+1] . chunk[1]^T, then hand it back
      auto pair = [&](uint32_t d, uint32_t a_lo, int ksteps1, uint32_t accumulate) {
        mbar_wait_addr(wfull + 8 * st, par);
        const uint32_t b_lo = w_lo + st * (2 * kSlabLo);
        umma_slab_e<4>(d, a_lo, b_lo, kIdesc128, el, accumulate);
        if (ksteps1 == 4) umma_slab_e<4>(d, a_lo + kSlabLo, b_lo + kSlabLo, kIdesc128, el, 1u);
        else umma_slab_e<2>(d, a_lo + kSlabLo, b_lo + kSlabLo, kIdesc128, el, 1u);
        umma_commit_e(wempty + 8 * st, el);
        if (++st == kStages) { st = 0; par ^= 1; }
      };
      auto gemm1 = [&](uint32_t k) {       // D2[:, 0:128] = A1[128 x 96] . W1m^T   (tile ordinal k)
        mbar_wait(&bars.a1_full, k & 1);
        tc_fence_after();
        pair(tmem + 128, a2_lo, 2, 0u);
        umma_commit_e(smem_u32(&bars.d1_full), el);
      };
      auto gemm2 = [&](uint32_t k) {       // D2[:, nh*128 ..] = A2[128 x 128] . W2[nh]^T, chunks (nh, ks)
        mbar_wait(&bars.a2_full, k & 1);
        tc_fence_after();
        pair(tmem + 128, a2_lo, 4, 0u);
        pair(tmem + 256, a2_lo, 4, 0u);
        umma_commit_e(smem_u32(&bars.d2_full), el);
      };
      gemm1(0);
      gemm2(0);
      uint32_t it = 0;
      long long* prof = (blockIdx.x == 0 && el) ? g_tc_prof : nullptr;
      constexpr int prof_role = 0;
      for (long long tile = first; tile < tiles; tile += stride, ++it) {
        const bool has_next = tile + stride < tiles;
        TC_STAMP(0);
        mbar_wait(&bars.a3_full, it & 1);      // kernel basis of tile `it` is in A3 (and D2 has been read out)
        tc_fence_after();
        TC_STAMP(1);
#pragma unroll 1
        for (int l = 0; l < kL; ++l) {
          // X0 is used by layers 0, 2, 4 (use number 3 it + l/2), X1 by layers 1, 3 (2 it + l/2)
          const int b = l & 1;
          const uint32_t use_par = b ? (uint32_t)(l >> 1) & 1u : (it + (uint32_t)(l >> 1)) & 1u;
          mbar_wait(&bars.x_empty[b], use_par ^ 1u);
          tc_fence_after();
          const uint32_t d = tmem + b * 384;
          pair(d, a3_lo, 4, 0u);
          pair(d, a3_lo + 2 * kSlabLo, 4, 1u);
          umma_commit_e(smem_u32(&bars.x_full[b]), el);
          TC_STAMP(2 + 2 * l);
          if (has_next) {
            if (l == 1) gemm1(it + 1);
            if (l == 3) gemm2(it + 1);
          }
          TC_STAMP(3 + 2 * l);
        }
      }
    }
  } else if (warp >= kEpiWarp0) {
    // ---------------- generator + epilogues: warp = (lane quarter q, column group cgi) ----------------
    const int q = warp & 3, cgi = (warp - kEpiWarp0) >> 2;
    const int m = q * 32 + lane;                     // tile row = (edge m / 16, orientation m % 16)
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const int hf = q >> 1;                           // staging half of this warp's rows
    const bool is_issuer = cgi == 0 && (q & 1) == 0 && lane == 0;     // one bulk-store issuer per half tile
    const float ox = ori[3 * (m & 15)], oy = ori[3 * (m & 15) + 1], oz = ori[3 * (m & 15) + 2];
    float win_cur = 0.f;
    // A1 of tile ordinal k: [dir.ori, |dir - (dir.ori) ori|, dist, cos(dir,a), cos(dir,b), cos(dir,c)] -> 83 monomials
    // + constant 1 (bias) -> fp16 into A2; 3 of the 12 16-byte chunks per thread
    auto gen = [&](uint32_t k) {
      const int buf = k & 1;
      mbar_wait(&bars.g_full[buf], (k >> 1) & 1);
      const float* gp = s_geo + (buf * kEdgesPerTile + (m >> 4)) * kGeo;
      const float dx = gp[0], dy = gp[1], dz = gp[2];
      float attr[6];
      const float i1 = dx * ox + dy * oy + dz * oz;
      const float px = dx - i1 * ox, py = dy - i1 * oy, pz = dz - i1 * oz;
      attr[0] = i1;
      attr[1] = sqrtf(px * px + py * py + pz * pz);
      attr[2] = gp[3]; attr[3] = gp[4]; attr[4] = gp[5]; attr[5] = gp[6];
      win_cur = gp[7];
      const float one = gp[8];
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars.g_empty[buf]);
      switch (cgi) {
        case 0: store_mono_part<0>(attr, one, A2, m); break;
        case 1: store_mono_part<1>(attr, one, A2, m); break;
        case 2: store_mono_part<2>(attr, one, A2, m); break;
        default: store_mono_part<3>(attr, one, A2, m); break;
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars.a1_full);
    };
    // epilogue 1: hidden = GELU(D2[:, 0:128]) -> A2 (hidden unit cgi*32.. -> slab cgi>>1, chunks (cgi&1)*4..)
    auto epi1 = [&](uint32_t k) {
      mbar_wait(&bars.d1_full, k & 1);
      tc_fence_after();
      float v[32];
      tmem_ld32(tmem + 128 + lane_addr + cgi * 32, v);
      gelu_store32<false, false>(v, nullptr, 1.0f, A2 + (cgi >> 1) * 16384 + m * kRowBytes, (cgi & 1) * 4, m);
      tc_fence_before();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars.a2_full);
    };
    // epilogue 2: kernel basis = GELU(D2 + b2) * window -> A3; this warp: columns cgi*64 .. +63 = slab cgi.
    // A3 is still being read by GEMM3 of the current tile when D2 of the next tile is ready, and the 64 GELUs per
    // thread are the longest stretch of the epilogue: so the math runs early (epi2_compute) and parks the packed
    // fp16 rows in the accumulator buffer X1, which is idle between layers 3 and 1; once GEMM3 has finished only
    // a TMEM -> shared copy (epi2_store) stands between two GEMM3 phases.  The first tile writes A3 directly.
    auto epi2_direct = [&](uint32_t k) {
      mbar_wait(&bars.d2_full, k & 1);
      tc_fence_after();
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        const int col0 = cgi * 64 + g * 32;
        float v[32];
        tmem_ld32(tmem + 128 + lane_addr + col0, v);
        gelu_store32<true, true>(v, s_b2 + col0, win_cur, A3 + cgi * 16384 + m * kRowBytes, g * 4, m);
      }
      tc_fence_before();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars.a3_full);
    };
    auto epi2_compute = [&](uint32_t k) {          // needs: D2 of tile k complete, X1 drained (layer 3 read out)
      mbar_wait(&bars.d2_full, k & 1);
      tc_fence_after();
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        const int col0 = cgi * 64 + g * 32;
        float v[32];
        tmem_ld32(tmem + 128 + lane_addr + col0, v);
        uint32_t pk[16];
        gelu_scaled_pack32(v, s_b2 + col0, win_cur, pk);
        tmem_st16(tmem + 384 + lane_addr + cgi * 32 + g * 16, pk);
      }
      tmem_wait_st();
    };
    auto epi2_store = [&]() {                      // needs: GEMM3 of the current tile complete (A3 free)
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        uint32_t pk[16];
        tmem_ld16(tmem + 384 + lane_addr + cgi * 32 + g * 16, pk);
        uint8_t* row = A3 + cgi * 16384 + m * kRowBytes;
#pragma unroll
        for (int cc = 0; cc < 4; ++cc)
          *reinterpret_cast<uint4*>(row + (((g * 4 + cc) ^ (m & 7)) << 4)) =
              make_uint4(pk[cc * 4], pk[cc * 4 + 1], pk[cc * 4 + 2], pk[cc * 4 + 3]);
      }
      tc_fence_before();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars.a3_full);
    };
    // epilogue 3: kernels[l][e][o][c] = X[l&1] (fp16), channels cgi*32 .. +31.  The layer's output tile is 32 KB
    // contiguous in HBM: it is staged in shared memory and written by the TMA engine with bulk stores (scattered
    // 32-byte stores from 512 threads would monopolise the LSU).  Rows are 256 B; the 16-byte chunk k of row (e, o)
    // is stored at chunk position k ^ o (the fp16 kernels layout, undone by the message kernel's loads), which
    // makes these 16-byte shared stores conflict free.  The two 16 KB halves (rows 0..63 / 64..127) are staged,
    // stored and recycled independently by the 8 warps owning those rows.
    auto epi3 = [&](int l, long long tile, uint32_t it, bool store_next_a3) {
      const int b = l & 1;
      const uint32_t use_par = b ? (uint32_t)(l >> 1) & 1u : (it + (uint32_t)(l >> 1)) & 1u;
      mbar_wait(&bars.x_full[b], use_par);
      tc_fence_after();
      float v[32];
      tmem_ld32(tmem + b * 384 + lane_addr + cgi * 32, v);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars.x_empty[b]);       // accumulators are in registers: release the buffer
      if (store_next_a3) epi2_store();                    // layer 4: GEMM3 is done with A3 -> next tile's kernel basis
      // The layer's output tile is 32 KB contiguous in HBM: it is staged in shared memory and written by the TMA
      // engine with bulk stores.  (Measured alternatives: scattered 32-byte stores straight from registers 2.94 ms,
      // coalesced 16-byte stores from the staging buffer 2.49 ms, bulk stores 2.39 ms; without any output store the
      // kernel takes 1.75 ms -- a per-SM write path of ~32 B/clk has to carry 160 KB per tile.)  Rows are 256 B; the
      // 16-byte chunk k of row (e, o) is staged at chunk position k ^ o -- conflict-free shared stores -- and that
      // is also the fp16 kernels layout in HBM (undone by the message kernel's loads).  The two 16 KB halves
      // (rows 0..63 / 64..127) are staged, stored and recycled independently by the 8 warps owning those rows.
      uint8_t* stage = S + hf * 16384;
      if (is_issuer) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // this half's previous store has left smem
      asm volatile("bar.sync %0, %1;" ::"r"(1 + hf), "n"(kEpiThreads / 2) : "memory");
      uint8_t* orow = stage + (m & 63) * 256;
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        uint4 pk;
        pk.x = pack_f16(v[cc * 8 + 0], v[cc * 8 + 1]);
        pk.y = pack_f16(v[cc * 8 + 2], v[cc * 8 + 3]);
        pk.z = pack_f16(v[cc * 8 + 4], v[cc * 8 + 5]);
        pk.w = pack_f16(v[cc * 8 + 6], v[cc * 8 + 7]);
        *reinterpret_cast<uint4*>(orow + (((cgi * 4 + cc) ^ (m & 15)) << 4)) = pk;
      }
      fence_proxy_async();
      asm volatile("bar.sync %0, %1;" ::"r"(1 + hf), "n"(kEpiThreads / 2) : "memory");
      if (is_issuer) {
        const long long left = E - tile * kEdgesPerTile - hf * (kEdgesPerTile / 2);   // edges of this half still valid
        if (left > 0) {
          const uint32_t bytes = (uint32_t)(left < kEdgesPerTile / 2 ? left : kEdgesPerTile / 2) * (kO * kC * 2);
          __half* out = kernels + ((size_t)l * edge_capacity * kO + (size_t)tile * kTileM + hf * 64) * kC;
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out), "r"(smem_u32(stage)), "r"(bytes)
                       : "memory");
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    };
    if (first < tiles) {
      gen(0);
      epi1(0);
      epi2_direct(0);
      uint32_t it = 0;
      long long* prof = (blockIdx.x == 0 && threadIdx.x == kEpiWarp0 * 32) ? g_tc_prof : nullptr;
      constexpr int prof_role = 1;
      for (long long tile = first; tile < tiles; tile += stride, ++it) {
        const bool has_next = tile + stride < tiles;
        TC_STAMP(0);
        if (has_next) gen(it + 1);                 // A2 is free: GEMM2 of this tile completed before d2_full fired
        TC_STAMP(1);
        TC_STAMP(2);
        epi3(0, tile, it, false);
        TC_STAMP(3);
        epi3(1, tile, it, false);
        TC_STAMP(4);
        if (has_next) epi1(it + 1);
        TC_STAMP(5);
        epi3(2, tile, it, false);
        TC_STAMP(6);
        epi3(3, tile, it, false);                  // X1 is drained: it can hold the next tile's packed kernel basis
        TC_STAMP(7);
        if (has_next) epi2_compute(it + 1);
        TC_STAMP(8);
        epi3(4, tile, it, has_next);               // its x_full also says GEMM3 has finished reading A3
        TC_STAMP(9);
      }
      if (is_issuer) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // all output tiles have landed
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

int num_sms_tc() {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  return sms;
}

}  // namespace

// debug only (not part of the public ABI): clock64 stamps of CTA 0 of the edge kernel, [6 tiles][2 roles][16]
extern "C" int arreau_debug_set_tc_profile(long long* buf) {
  return (int)cudaMemcpyToSymbol(g_tc_prof, &buf, sizeof(buf));
}

static int convnext_mlp_f16_launch(const void* y_img, const void* w_img, const float* b1, const float* b2,
                                   const float* layer_scale, int64_t num_rows, float* h, const float* ori,
                                   float* pool_out, const float* pool_wz, int pool_cols, void* stream) {
  if (num_rows == 0) return ARREAU_OK;
  if (!y_img || !w_img || !b1 || !b2 || !layer_scale || !h) return ARREAU_ERR_NULL;
  if (num_rows < 0 || (pool_out && num_rows % kO != 0)) return ARREAU_ERR_BAD_SHAPE;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(convnext_mlp_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, mlp::kSmemBytes);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  const long long tiles = (num_rows + kTileM - 1) / kTileM;
  const int grid = (int)(tiles < (long long)num_sms_tc() ? tiles : (long long)num_sms_tc());
  convnext_mlp_tc_kernel<<<grid, kThreads, mlp::kSmemBytes, (cudaStream_t)stream>>>(
      (const uint8_t*)y_img, (const uint8_t*)w_img, b1, b2, layer_scale, (long long)num_rows, h, ori, pool_out, pool_wz,
      pool_cols);
  CUDA_LAUNCH_CHECK();
  return ARREAU_OK;
}

extern "C" int arreau_convnext_mlp_f16(const void* y_img, const void* w_img, const float* b1, const float* b2,
                                        const float* layer_scale, int64_t num_rows, float* h, void* stream) {
  return convnext_mlp_f16_launch(y_img, w_img, b1, b2, layer_scale, num_rows, h, nullptr, nullptr, nullptr, 0, stream);
}

extern "C" int arreau_convnext_mlp_f16_pooled(const void* y_img, const void* w_img, const float* b1, const float* b2,
                                               const float* layer_scale, int64_t num_rows, float* h, const float* ori,
                                               float* pool_out, const float* readout_v_k, int32_t num_states,
                                               void* stream) {
  if (num_rows > 0 && (!ori || !pool_out || !readout_v_k)) return ARREAU_ERR_NULL;
  if (num_states <= 0) return ARREAU_ERR_BAD_SHAPE;
  // the spare warps contract the vector-pooled parts with column Z of the entry's [C][Z+6] read-out matrix
  return convnext_mlp_f16_launch(y_img, w_img, b1, b2, layer_scale, num_rows, h, ori, pool_out, readout_v_k + num_states,
                                 num_states + 6, stream);
}

extern "C" int arreau_edge_kernels_f16(const double* dir, const double* dist, const double* lattice,
                                        const int32_t* crystal_of_atom, const int32_t* src, const int32_t* num_edges_ptr,
                                        int64_t edge_capacity, const float* ori, const void* w1_img, const void* w_img,
                                        const float* b2, double radius, void* kernels_f16, void* stream) {
  if (edge_capacity == 0) return ARREAU_OK;
  if (!dir || !dist || !lattice || !crystal_of_atom || !src || !num_edges_ptr || !ori || !w1_img || !w_img || !b2 ||
      !kernels_f16)
    return ARREAU_ERR_NULL;
  if (edge_capacity < 0) return ARREAU_ERR_BAD_SHAPE;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(edge_kernels_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, edge::kSmemBytes);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  const long long tiles = (edge_capacity + edge::kEdgesPerTile - 1) / edge::kEdgesPerTile;
  const int grid = (int)(tiles < (long long)num_sms_tc() ? tiles : (long long)num_sms_tc());
  edge_kernels_tc_kernel<<<grid, kThreads, edge::kSmemBytes, (cudaStream_t)stream>>>(
      dir, dist, lattice, crystal_of_atom, src, num_edges_ptr, (long long)edge_capacity, ori, (const uint8_t*)w1_img,
      (const uint8_t*)w_img, b2, radius, (__half*)kernels_f16);
  CUDA_LAUNCH_CHECK();
  return ARREAU_OK;
}
