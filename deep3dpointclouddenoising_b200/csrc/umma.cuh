// sm_100a building blocks shared by the staged-tile kernels: mbarrier, bulk asynchronous copies (cp.async.bulk ->
// SASS UBLKCP), tcgen05 (alloc / mma / commit / ld -> SASS UTC*MMA, LDTM) and the shared-memory matrix descriptors of
// the NO-SWIZZLE canonical layouts.  Inline PTX only; nothing here needs a header outside the CUDA toolkit.
//
// Canonical no-swizzle operand layouts (core matrix = 8 rows x 16 bytes, stored as 128 contiguous bytes):
//   K-major  (row = M/N index, 16 bytes = 8 bf16 along K):  SBO = byte stride between 8-row groups along M/N,
//                                                           LBO = byte stride between 16-byte K chunks
//   MN-major (row = K index,   16 bytes = 8 bf16 along M/N): SBO = byte stride between 16-byte chunks along M/N,
//                                                           LBO = byte stride between 8-row groups along K
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace umma {

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

// ---- mbarrier -------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// try_wait suspends the thread up to the hinted time (ns) before it reports "not yet": few re-issues while waiting
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, 2000;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}\n" ::"r"(bar), "r"(parity) : "memory");
}

// the same without a suspend-time hint (the hardware's default, short time limit): for waits on a critical path
__device__ __forceinline__ void mbar_wait_short(unsigned bar, unsigned parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}\n" ::"r"(bar), "r"(parity) : "memory");
}

// ---- bulk asynchronous copy global -> shared, completion counted in bytes on an mbarrier ------------------------
// dst / src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void bulk_g2s(unsigned dst_smem, const void* src, unsigned bytes, unsigned bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
               "l"(__cvta_generic_to_global(src)), "r"(bytes), "r"(bar)
               : "memory");
}

// generic-proxy shared-memory writes -> visible to the async proxy (tensor core / bulk copies)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- tcgen05 --------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// one full warp; cols a power of two >= 32
__device__ __forceinline__ void tmem_alloc(unsigned slot_smem, unsigned cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(unsigned base, unsigned cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(cols) : "memory");
}

// shared-memory matrix descriptor, no swizzle, sm_100 version field
__device__ __forceinline__ unsigned long long smem_desc(unsigned addr, unsigned lbo, unsigned sbo) {
  return (unsigned long long)((addr >> 4) & 0x3fffu) | ((unsigned long long)((lbo >> 4) & 0x3fffu) << 16) |
         ((unsigned long long)((sbo >> 4) & 0x3fffu) << 32) | (1ull << 46);
}

// instruction descriptor for kind::f16 with bf16 operands, fp32 accumulator, M = 128
__host__ __device__ constexpr unsigned idesc_bf16(int n, bool a_mn_major, bool b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn_major ? 1u : 0u) << 15) | ((b_mn_major ? 1u : 0u) << 16) |
         ((unsigned)(n >> 3) << 17) | ((128u >> 4) << 24);
}

// D[tmem] (+)= A[smem] . B[smem]; issued by ONE thread
__device__ __forceinline__ void mma_bf16(unsigned tmem_d, unsigned long long a_desc, unsigned long long b_desc, unsigned idesc,
                                         unsigned accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives on the mbarrier once every MMA issued so far by this thread has completed
__device__ __forceinline__ void mma_commit(unsigned bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// 16 consecutive accumulator columns of this thread's TMEM lane (lane = 32 * (warp % 4) + lane id)
__device__ __forceinline__ void tmem_ld16(unsigned taddr, unsigned (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// vector reduction into global memory: four consecutive fp32 values, 16-byte aligned (SASS REDG.E.ADD.F32x4)
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// ---- exact three-way bf16 split of two fp32 values: x = h1 + h2 + h3 (8 + 8 + 8 mantissa bits) -------------------
// each result packs (x0 -> low half, x1 -> high half) like two adjacent bf16 in memory
__device__ __forceinline__ unsigned pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<unsigned*>(&v);
}
__device__ __forceinline__ void split3_bf16x2(float x0, float x1, unsigned& h1, unsigned& h2, unsigned& h3) {
  h1 = pack_bf16x2(x0, x1);
  const float r0 = x0 - __uint_as_float(h1 << 16), r1 = x1 - __uint_as_float(h1 & 0xffff0000u);
  h2 = pack_bf16x2(r0, r1);
  const float s0 = r0 - __uint_as_float(h2 << 16), s1 = r1 - __uint_as_float(h2 & 0xffff0000u);
  h3 = pack_bf16x2(s0, s1);
}

}  // namespace umma
