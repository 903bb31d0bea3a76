// umma.cuh -- thin inline-PTX wrappers for the sm_100a tensor-core path (tcgen05 + TMEM + mbarrier).
//
// Conventions used by the kernels in this repo:
//   * cta_group::1 only (one CTA drives the tensor core of its own SM);
//   * A operand in TMEM (tcgen05.st by the owning threads: lane = row = point, column = k), B operand in
//     shared memory as a K-major, no-swizzle "interleaved" image of 8x(16 byte) core matrices;
//   * accumulators (D) in TMEM, read back with tcgen05.ld.32x32b (thread i of warp w reads lane 32*(w%4)+i);
//   * fp32-class accuracy from TF32 tensor cores by the 3xTF32 split  x = hi + lo,
//       A.B ~= A_lo.B_hi + A_hi.B_lo + A_hi.B_hi   (fp32 accumulate in TMEM).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ifd {
namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- TMEM allocation (one full warp executes these) ------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// ---- fences ----------------------------------------------------------------------------------------
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// generic-proxy smem writes -> visible to the async proxy (the tensor core reads B through it)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- mbarrier --------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done = 0;
  int spins = 0;
  do {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (!done && ++spins > (1 << 22)) __trap();   // a lost tcgen05.commit must fail the launch, not hang the GPU
  } while (!done);
}
// ---- TMA bulk copy (cp.async.bulk, SASS UBLKCP): global -> shared, completion counted in bytes on an mbarrier -------------
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// dst, src 16-byte aligned, bytes a multiple of 16.  One thread issues; the data lands through the async proxy, which is
// also the proxy the tensor core reads shared-memory operands through (no fence.proxy.async on this path).
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// arrive on `bar` once every previously issued tcgen05.mma of this thread has completed
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- single-thread issue without the divergence tax ------------------------------------------------------------------------
// tcgen05.mma / commit are per-thread instructions.  Issued under `if (threadIdx.x == leader)` ptxas must assume any subset of
// lanes may be active and wraps EVERY instruction in an ELECT / R2UR.BROADCAST / BRA.U.ANY loop (12-20 SASS instructions per
// MMA, measured); for the N = 32 MMAs of the decode chain that loop, not the tensor pipe, set the issue rate.  The pattern
// ptxas understands: a WARP-UNIFORM branch selects the issuing warp (uniform values come from warp_bcast), elect_one() picks
// the lane -- the operands then live in uniform registers and the MMAs issue back to back.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ uint32_t warp_bcast(uint32_t v) { return __shfl_sync(0xffffffffu, v, 0); }

// ---- descriptors -----------------------------------------------------------------------------------
// K-major, SWIZZLE_NONE shared-memory matrix descriptor (sm_100 version bits = 1).  In 16-byte units the
// canonical layout is ((8,n),2):((1,SBO),LBO): 8 rows x 16 B contiguous per core matrix, `lbo` bytes between
// the two core matrices along K of one MMA, `sbo` bytes between 8-row groups along M/N.
__device__ __forceinline__ uint64_t smem_desc_kmajor(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  return d;                // base_offset 0, lbo_mode 0, layout_type 0 (no swizzle)
}
// Instruction descriptor, kind::tf32, fp32 accumulate, A and B K-major, dense.
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// D[tmem] (+)= A[tmem] . B[smem]^T ; one elected thread issues.  K = 8 per instruction for tf32.
__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}

// D[tmem] (+)= A[smem] . B[smem]^T, kind::tf32, one elected thread issues
__device__ __forceinline__ void mma_tf32_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}

// ---- TMEM <-> registers: 32 lanes x 32 columns per warp ----------------------------------------------
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]),
      "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]),
      "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- 3xTF32 split ----------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t tf32_rna(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return u;
}
__device__ __forceinline__ uint32_t tf32_hi(float x) { return tf32_rna(x); }
__device__ __forceinline__ uint32_t tf32_lo(float x) { return tf32_rna(x - __uint_as_float(tf32_rna(x))); }
// Hot-path variants: round-to-nearest (ties away) by integer arithmetic on the bit pattern -- 2 instructions
// instead of the 3 of cvt.rna (which also screens Inf/NaN; activations on this path are finite).  x - hi is exact.
__device__ __forceinline__ uint32_t tf32_hi_fast(float x) { return (__float_as_uint(x) + 0x1000u) & 0xffffe000u; }
// Round-to-nearest-even to TF32 in ONE instruction (SASS F2FP.SATFINITE.TF32.F32.PACK_B; satfinite only clamps +-Inf / NaN,
// which this path never sees).  Differs from tf32_hi_fast on exact ties only; x - hi stays exact.
__device__ __forceinline__ uint32_t tf32_hi_cvt(float x) {
  uint32_t u;
  asm("cvt.rn.satfinite.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return u;
}
__device__ __forceinline__ uint32_t tf32_lo_fast(float x, uint32_t hi) {
  return (__float_as_uint(x - __uint_as_float(hi)) + 0x1000u) & 0xffffe000u;
}

// Byte offset of element (n, k) inside a K-major no-swizzle image of an [N][K] fp32/tf32 matrix, laid out as
// [K/4][N/8] core matrices of 128 bytes: LBO = (N/8)*128, SBO = 128.
__host__ __device__ constexpr uint32_t img_offset(int n, int k, int N) {
  return (uint32_t)(((k >> 2) * (N >> 3) + (n >> 3)) * 128 + (n & 7) * 16 + (k & 3) * 4);
}

}  // namespace umma
}  // namespace ifd
