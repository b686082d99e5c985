// tc.cuh — tcgen05 / TMEM building blocks shared by the many-chain kernels (chains.cu, chains_wide.cu):
// inline-PTX wrappers, shared-memory matrix descriptors, the kind::tf32 instruction descriptor, 3xTF32 split.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace edhmc {

// ------------------------------------------------------------------------------------------------
// tcgen05 helpers (inline PTX; SASS: UTCHMMA-family / LDTM / STTM)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t mbar_s) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar_s) : "memory");
}
// D[tmem] (+)= A[smem] · B[smem]
__device__ __forceinline__ void tc_mma_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
// D[tmem] (+)= A[tmem] · B[smem]
__device__ __forceinline__ void tc_mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
// Narrower shapes of the same 32x32b access (lane = TMEM lane of the warp's quarter, N consecutive 32-bit columns per lane):
// used by the single-chain sampler, which parks rows of X in TMEM for a whole persistent launch (stream_ldg.cuh).
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld2(uint32_t taddr, uint32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];" : "=r"(v[0]), "=r"(v[1]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]),
               "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const uint32_t* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3])
               : "memory");
}
__device__ __forceinline__ void tmem_st2(uint32_t taddr, const uint32_t* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1,%2};" ::"r"(taddr), "r"(v[0]), "r"(v[1]) : "memory");
}
// N (even) consecutive columns as a ladder of x16 / x8 / x4 / x2 accesses
template <int N>
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, uint32_t (&v)[N]) {
  static_assert(N % 2 == 0, "even column counts only");
  constexpr int n16 = N / 16, r16 = N - 16 * n16;
#pragma unroll
  for (int i = 0; i < n16; ++i) tmem_ld16(taddr + 16 * i, &v[16 * i]);
  if constexpr ((r16 & 8) != 0) tmem_ld8(taddr + 16 * n16, &v[16 * n16]);
  if constexpr ((r16 & 4) != 0) tmem_ld4(taddr + 16 * n16 + (r16 & 8), &v[16 * n16 + (r16 & 8)]);
  if constexpr ((r16 & 2) != 0) tmem_ld2(taddr + 16 * n16 + (r16 & 12), &v[16 * n16 + (r16 & 12)]);
}
template <int N>
__device__ __forceinline__ void tmem_st_cols(uint32_t taddr, const uint32_t (&v)[N]) {
  static_assert(N % 2 == 0, "even column counts only");
  constexpr int n16 = N / 16, r16 = N - 16 * n16;
#pragma unroll
  for (int i = 0; i < n16; ++i) tmem_st16(taddr + 16 * i, &v[16 * i]);
  if constexpr ((r16 & 8) != 0) tmem_st8(taddr + 16 * n16, &v[16 * n16]);
  if constexpr ((r16 & 4) != 0) tmem_st4(taddr + 16 * n16 + (r16 & 8), &v[16 * n16 + (r16 & 8)]);
  if constexpr ((r16 & 2) != 0) tmem_st2(taddr + 16 * n16 + (r16 & 12), &v[16 * n16 + (r16 & 12)]);
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor, K-major, no swizzle (canonical layout ((8,n),2):((1,SBO),LBO) in 16-byte
// units): a core matrix is 8 rows x 16 bytes stored contiguously (128 B); SBO = byte step between 8-row
// groups along M/N, LBO = byte step between core matrices along K. Bits: start>>4 [0,14), LBO>>4 [16,30),
// SBO>>4 [32,46), version=1 [46,48), layout_type=0 (SWIZZLE_NONE) [61,64).
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr_s, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr_s >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= 1ull << 46;
  return d;
}
// Instruction descriptor kind::tf32: D=f32 (bits[4,6)=1), A=B=tf32 (bits[7,10)=bits[10,13)=2), K-major A and B,
// N>>3 at [17,23), M>>4 at [24,29).
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}

// One lane of a converged warp (elect.sync): the pattern under which tcgen05.mma issues at the hardware rate.
// Issued from a divergent `if (tid == 0)` region the compiler wraps every UTCHMMA in an ELECT/R2UR/BRA.U.ANY
// loop and the issue cost rises from ~35 to ~190 cycles per instruction (measured, tools/micro/mma_chain*.cu).
__device__ __forceinline__ bool elect_one() {
  uint32_t p;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(p));
  return p != 0;
}

__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
  hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);  // the 19 bits the tensor core reads
  lo = x - hi;                                               // exact in fp32
}


}  // namespace edhmc
