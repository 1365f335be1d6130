// tcgen05 / TMEM / mbarrier / bulk-copy primitives for the sm_100a tensor-core kernels (inline PTX).
//
// Operand tiles live in shared memory in the canonical K-major SWIZZLE_128B layout of the UMMA shared
// memory descriptor: a tile of R rows x 64 fp16 (one "K slab") is R rows of 128 bytes, 8-row groups 1024 B
// apart, and inside every 1024-B group the 16-byte chunk c of row r sits at chunk position c ^ (r & 7).
// Tiles wider than 64 in K are several slabs back to back.  Weight tiles are pre-arranged in this image on
// the host (arreau_b200/weights.py: umma_tile_image) so that one cp.async.bulk moves a whole tile.
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

namespace tc {

constexpr int kSlabK = 64;                 // fp16 elements per 128-byte swizzle row
constexpr int kRowBytes = 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// byte offset of element (row r, k) inside a tile of `rows` rows (k may span several slabs)
__device__ __forceinline__ uint32_t tile_offset(int rows, int r, int k) {
  const int slab = k >> 6, kk = k & 63;
  return (uint32_t)(slab * rows * kRowBytes + r * kRowBytes + ((((kk >> 3) ^ (r & 7)) << 4) | ((kk & 7) << 1)));
}

// ---- mbarrier ----------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// generic-proxy writes to shared memory -> visible to the async proxy (UMMA operand reads)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- bulk copy global -> shared (TMA engine, no tensor map), completion on an mbarrier ----------
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ---- TMEM ---------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result, uint32_t cols) {   // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {          // the same warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// 32 lanes x 32 columns of 32-bit accumulators: thread i of the warp gets row (lane base + i)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// 32 lanes x 16 columns: load / store 16 registers per thread (used to park packed fp16 operand rows in a free
// accumulator buffer)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// the same load without the wait: issue several, then tmem_wait_ld() once (their latencies overlap); the registers
// must not be read before the wait
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- UMMA descriptors and issue ----------------------------------------------------------------
// K-major SWIZZLE_128B operand starting at shared address `addr` (1024-B aligned tile base, plus a
// multiple of 32 B to step through the 64-wide slab in UMMA_K = 16 chunks).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t addr) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFFu) >> 4);   // start address            bits [0,14)
  d |= (uint64_t)1 << 16;                     // leading byte offset (unused for swizzled K-major) [16,30)
  d |= (uint64_t)(1024 >> 4) << 32;           // stride byte offset: 8-row groups 1024 B apart     [32,46)
  d |= (uint64_t)1 << 46;                     // descriptor version 1 (sm_100)                     [46,48)
  d |= (uint64_t)2 << 61;                     // SWIZZLE_128B                                      [61,64)
  return d;
}
// kind::f16 instruction descriptor: fp16 x fp16 -> fp32, both operands K-major, M x N tile
__host__ __device__ constexpr uint32_t umma_idesc_f16(int M, int N) {
  // c_format F32 = 1 at bit 4; a_format / b_format F16 = 0 at bits 7 / 10; both K-major (bits 15, 16 = 0)
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// The single MMA-issuing thread is instruction bound (a 128x128x16 MMA retires every 64 cycles): keep the
// descriptors as precomputed 32-bit halves.  The low word carries the start address (>> 4, 14 bits) and the
// leading byte offset; stepping K by 16 fp16 (32 B) adds 2 to it and never carries into the high word.
constexpr uint32_t kDescHiSw128 = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint32_t umma_desc_lo(uint32_t addr) { return ((addr & 0x3FFFFu) >> 4) | (1u << 16); }
template <bool kAccumulate>
__device__ __forceinline__ void umma_f16_lo(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc) {
  if constexpr (kAccumulate)
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %4};\n\tmov.b64 db, {%2, %4};\n\t"
        "setp.eq.u32 p, 0, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(kDescHiSw128)
        : "memory");
  else
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %4};\n\tmov.b64 db, {%2, %4};\n\t"
        "setp.ne.u32 p, 0, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(kDescHiSw128)
        : "memory");
}
// runtime accumulate flag (first K step of a GEMM)
__device__ __forceinline__ void umma_f16_lo_p(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %4};\n\tmov.b64 db, {%2, %4};\n\t"
      "setp.ne.u32 p, %5, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(kDescHiSw128), "r"(accumulate)
      : "memory");
}
// the KSTEPS (<= 4) UMMA_K = 16 steps of one 64-wide K slab; kFirst: the first step overwrites the accumulator
template <int KSTEPS, bool kFirst>
__device__ __forceinline__ void umma_slab_lo(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc) {
  if constexpr (kFirst) umma_f16_lo<false>(tmem_d, a_lo, b_lo, idesc);
  else umma_f16_lo<true>(tmem_d, a_lo, b_lo, idesc);
#pragma unroll
  for (int k = 1; k < KSTEPS; ++k) umma_f16_lo<true>(tmem_d, a_lo + 2 * k, b_lo + 2 * k, idesc);
}
__device__ __forceinline__ void mbar_wait_addr(uint32_t addr, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!done);
}
// Warp-uniform MMA issue: the whole warp runs the issue loop (so that descriptors and addresses live in uniform
// registers -- a single divergent thread pays an R2UR round trip per operand, ~50 cycles per MMA) and one elected
// lane issues.  Measured (scratch/ubench/mma_ring.cu): 64 cycles per 128x128x16 MMA = the tensor-pipe floor with one
// barrier round per 8 MMAs, against 121 for single-thread issue with a round per 4.
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred;
}
__device__ __forceinline__ void umma_f16_e(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t elected,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, e;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %4};\n\tmov.b64 db, {%2, %4};\n\t"
      "setp.ne.u32 e, %5, 0;\n\t"
      "setp.ne.u32 p, %6, 0;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(kDescHiSw128), "r"(elected), "r"(accumulate)
      : "memory");
}
// KSTEPS (2 or 4) UMMA_K = 16 steps of one 64-wide K slab
template <int KSTEPS>
__device__ __forceinline__ void umma_slab_e(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t elected,
                                            uint32_t accumulate_first) {
  umma_f16_e(tmem_d, a_lo, b_lo, idesc, elected, accumulate_first);
#pragma unroll
  for (int k = 1; k < KSTEPS; ++k) umma_f16_e(tmem_d, a_lo + 2 * k, b_lo + 2 * k, idesc, elected, 1u);
}
// A operand from tensor memory (row m in lane m, the K values packed two fp16 per 32-bit column, 8 columns per
// UMMA_K = 16 step), B from shared memory: D (+)= A[tmem] . B^T
__device__ __forceinline__ void umma_f16_ts_e(uint32_t tmem_d, uint32_t a_tmem, uint32_t b_lo, uint32_t idesc, uint32_t elected,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, e;\n\t.reg .b64 db;\n\t"
      "mov.b64 db, {%2, %4};\n\t"
      "setp.ne.u32 e, %5, 0;\n\t"
      "setp.ne.u32 p, %6, 0;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(a_tmem), "r"(b_lo), "r"(idesc), "r"(kDescHiSw128), "r"(elected), "r"(accumulate)
      : "memory");
}
template <int KSTEPS>
__device__ __forceinline__ void umma_slab_ts_e(uint32_t tmem_d, uint32_t a_tmem, uint32_t b_lo, uint32_t idesc, uint32_t elected,
                                               uint32_t accumulate_first) {
  umma_f16_ts_e(tmem_d, a_tmem, b_lo, idesc, elected, accumulate_first);
#pragma unroll
  for (int k = 1; k < KSTEPS; ++k) umma_f16_ts_e(tmem_d, a_tmem + 8 * k, b_lo + 2 * k, idesc, elected, 1u);
}
__device__ __forceinline__ void umma_commit_e(uint32_t bar_addr, uint32_t elected) {
  asm volatile(
      "{\n\t.reg .pred e;\n\tsetp.ne.u32 e, %1, 0;\n\t"
      "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(bar_addr), "r"(elected)
      : "memory");
}

// wait for two / four barriers at once: the try_waits are issued back to back so that their latencies overlap
// (the MMA-issuing thread has no slack: the tensor pipe's issue queue is shallow, any gap between two
// tcgen05.mma shows up as idle tensor cycles)
__device__ __forceinline__ void mbar_wait4_addr(uint32_t a0, uint32_t p0, uint32_t a1, uint32_t p1, uint32_t a2, uint32_t p2,
                                                uint32_t a3, uint32_t p3) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred q0, q1, q2, q3;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 q0, [%1], %2;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 q1, [%3], %4;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 q2, [%5], %6;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 q3, [%7], %8;\n\t"
        "and.pred q0, q0, q1;\n\tand.pred q2, q2, q3;\n\tand.pred q0, q0, q2;\n\t"
        "selp.u32 %0, 1, 0, q0;\n\t}"
        : "=r"(done)
        : "r"(a0), "r"(p0), "r"(a1), "r"(p1), "r"(a2), "r"(p2), "r"(a3), "r"(p3)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void mbar_wait2_addr(uint32_t a0, uint32_t p0, uint32_t a1, uint32_t p1) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred q0, q1;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 q0, [%1], %2;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 q1, [%3], %4;\n\t"
        "and.pred q0, q0, q1;\n\t"
        "selp.u32 %0, 1, 0, q0;\n\t}"
        : "=r"(done)
        : "r"(a0), "r"(p0), "r"(a1), "r"(p1)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void umma_commit_addr(uint32_t bar_addr) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_addr) : "memory");
}

// arrive on an mbarrier when every tcgen05 operation issued so far by this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// ---- epilogue math -----------------------------------------------------------------------------
// GELU on the fp16 path: x Phi(x) with Phi(x) = 0.5 (1 + tanh(x (a + b x^2))), (a, b) refitted to the exact
// erf form (max |gelu_fast - gelu_erf| = 2.7e-4 over all x).  Two values at a time in packed fp16
// (HMUL2 / HFMA2 / MUFU.TANH.F16x2): five packed FMA-pipe instructions and one MUFU per PAIR, and the result
// is already the packed fp16 operand of the next GEMM.  fp16 arithmetic (11 significant bits) adds ~5e-4
// relative, the same size as the rounding of the stored operand.  |x| > 255 overflows x^2 to +inf, which
// saturates tanh to +-1 (the correct limit).  The function returns x (1 + tanh(.)) = 2 gelu(x): every consumer is a GEMM
// whose weight image was packed with the factor 1/2 (arreau_b200/weights.py: W2 of the basis MLP, the five kernel
// projections, linear_2 of the ConvNext MLP), a power of two and therefore exact, which saves one packed multiply per pair.  The 3.7e9 GELUs of a step are otherwise the bound of both
// tensor-core kernels.  The fp32 path keeps the exact erff form.
__device__ __forceinline__ uint32_t gelu2_f16(float x0, float x1) {
  const __half2 x = __floats2half2_rn(x0, x1);
  const __half2 ca = __float2half2_rn(0.8001570785450266f), cb = __float2half2_rn(0.03470089338901844f);
  const __half2 u = __hmul2(x, __hfma2(cb, __hmul2(x, x), ca));
  uint32_t tb;
  asm("tanh.approx.f16x2 %0, %1;" : "=r"(tb) : "r"(*reinterpret_cast<const uint32_t*>(&u)));
  const __half2 t = *reinterpret_cast<const __half2*>(&tb);
  const __half2 y = __hfma2(x, t, x);       // = 2 gelu(x): the factor 1/2 lives in the consuming weight image (exact)
  return *reinterpret_cast<const uint32_t*>(&y);
}
__device__ __forceinline__ uint32_t gelu2_scaled_f16(float x0, float x1, __half2 scale) {
  const uint32_t g = gelu2_f16(x0, x1);
  const __half2 y = __hmul2(*reinterpret_cast<const __half2*>(&g), scale);
  return *reinterpret_cast<const uint32_t*>(&y);
}

// 32-byte global accesses (sm_100: LDG.256 / STG.256): one full sector per thread and instruction
__device__ __forceinline__ void ldg256(const float* p, float (&v)[8]) {
  asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
               : "l"(p));
}
__device__ __forceinline__ void stg256(float* p, const float (&v)[8]) {
  asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]),
               "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
               : "memory");
}

__device__ __forceinline__ void stg256_b32(void* p, const uint32_t* v) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]),
               "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}

__device__ __forceinline__ uint32_t pack_f16(float lo, float hi) {
  const __half2 v = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&v);
}

}  // namespace tc
