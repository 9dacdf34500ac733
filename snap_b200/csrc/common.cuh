// Shared device helpers for the sm_100a kernels: mbarrier, TMA (cp.async.bulk.tensor),
// tcgen05 (MMA / TMEM alloc / TMEM load) PTX wrappers, and small bf16 utilities.
// Everything here is inline PTX for sm_100a; there is no fallback path.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace snapb200 {

// ---------------------------------------------------------------------------------------------
// generic
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

// ---------------------------------------------------------------------------------------------
// mbarrier
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  // make generic-proxy smem writes visible to the async proxy (TMA / tcgen05 operand reads)
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      // suspend-time hint (ns): a failing probe parks the warp in hardware instead of spinning through the issue
      // slots the epilogue warps need (the polling loops were ~30% of all issued instructions of the conv GEMMs)
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(0x4000u)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

// for single-thread roles that wait through a whole phase of the worker warps (microseconds): sleep between probes so
// that the polling loop does not compete with the workers of the same scheduler for issue slots
__device__ __forceinline__ void mbar_wait_sleep(uint64_t* bar, uint32_t parity, unsigned ns) {
  while (!mbar_try_wait(bar, parity)) __nanosleep(ns);
}

// ---------------------------------------------------------------------------------------------
// TMA
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
// 2D tiled load: crd0 = innermost (K / channel) coordinate, crd1 = row coordinate (signed; rows
// outside the tensor are zero-filled by the hardware).
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* m, uint64_t* bar, void* smem_dst,
                                            int crd0, int crd1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(crd0),
        "r"(crd1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(const CUtensorMap* m, uint64_t* bar, void* smem_dst, int crd0, int crd1,
                                            int crd2, int crd3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(crd0), "r"(crd1),
        "r"(crd2), "r"(crd3)
      : "memory");
}
// 2D tiled store smem -> global (bulk async group); rows / columns outside the tensor are clipped by the hardware
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* smem_src, int crd0, int crd1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               :
               : "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(smem_src)), "r"(crd0), "r"(crd1)
               : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, const void* smem_src, int crd0, int crd1, int crd2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               :
               : "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(smem_src)), "r"(crd0), "r"(crd1), "r"(crd2)
               : "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all of this thread's bulk groups have finished READING their shared-memory source (it may be overwritten)
__device__ __forceinline__ void bulk_wait_group_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_group0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// 64-thread named barrier 1 + q (q in 0..3); immediates so that the kernel reserves 5 barriers, not all 16
__device__ __forceinline__ void pair_bar_sync(int q) {
  switch (q) {
    case 0: asm volatile("bar.sync 1, 64;" ::: "memory"); break;
    case 1: asm volatile("bar.sync 2, 64;" ::: "memory"); break;
    case 2: asm volatile("bar.sync 3, 64;" ::: "memory"); break;
    default: asm volatile("bar.sync 4, 64;" ::: "memory"); break;
  }
}
__device__ __forceinline__ void tma_load_3d(const CUtensorMap* m, uint64_t* bar, void* smem_dst,
                                            int crd0, int crd1, int crd2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(crd0),
        "r"(crd1), "r"(crd2)
      : "memory");
}

// ---------------------------------------------------------------------------------------------
// tcgen05
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tc_fence_before_sync() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after_sync() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// whole warp; writes the TMEM base address to *smem_dst
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]; single thread issues.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n"
      :
      : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier once all previously issued MMAs of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// 32 lanes x 16 consecutive 32-bit columns: thread i of the warp gets row (lane_base+i)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32"
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}

// Shared-memory matrix descriptor, K-major operand, hardware swizzle.
//   swizzle_bytes = 128: rows of 64 bf16 (128 B), 8-row atoms of 1024 B (SBO = 1024)
//   swizzle_bytes =  64: rows of 32 bf16 ( 64 B), 8-row atoms of  512 B (SBO =  512)
// Field layout (cute::UMMA::SmemDescriptor): [0,14) addr>>4, [16,30) LBO>>4, [32,46) SBO>>4,
// [46,48) version=1, [61,64) layout type (2 = SWIZZLE_128B, 4 = SWIZZLE_64B).
template <int SWIZZLE_BYTES>
__device__ __forceinline__ uint64_t make_kmajor_desc(uint32_t smem_addr) {
  constexpr uint64_t layout = (SWIZZLE_BYTES == 128) ? 2ull : (SWIZZLE_BYTES == 64 ? 4ull : 6ull);
  constexpr uint64_t sbo = (uint64_t)(8 * SWIZZLE_BYTES) >> 4;
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;  // LBO (ignored for swizzled K-major; canonical value 1)
  d |= sbo << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  d |= layout << 61;
  return d;
}

// Instruction descriptor for kind::f16, BF16 x BF16 -> FP32, both operands K-major, M = 128.
__host__ __device__ constexpr uint32_t make_idesc_bf16_m128(int n) {
  return (1u << 4)                      // D format  = F32
         | (1u << 7)                    // A format  = BF16
         | (1u << 10)                   // B format  = BF16
         | ((uint32_t)(n >> 3) << 17)   // N >> 3
         | ((uint32_t)(128 >> 4) << 24);  // M >> 4
}

// ---------------------------------------------------------------------------------------------
// bf16 helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
// packed bf16 add with one rounding per lane (never contracted)
__device__ __forceinline__ uint32_t hadd2_bf16_rn(uint32_t a, uint32_t b) {
  __nv_bfloat162 x = *reinterpret_cast<__nv_bfloat162*>(&a);
  const __nv_bfloat162 y = *reinterpret_cast<__nv_bfloat162*>(&b);
  x = __hadd2_rn(x, y);
  return *reinterpret_cast<uint32_t*>(&x);
}
// bf16 -> f32 is a shift / mask (ALU pipe)
__device__ __forceinline__ float bf16_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }
// round-to-nearest-even to bf16, result as f32.  Goes through the PACKED conversion (F2FP.BF16.F32.PACK_AB, ALU
// pipe) + a shift: the scalar cvt.rn.bf16.f32 compiles to F2F.BF16.F32, a quarter-rate conversion-unit instruction
// that dominated the issue slots of the rounding-heavy epilogues.
__device__ __forceinline__ float bf16_round(float x) { return bf16_lo(pack_bf16(x, 0.f)); }
// round two values at once: one conversion + shift + mask
__device__ __forceinline__ void bf16_round2(float& a, float& b) {
  const uint32_t u = pack_bf16(a, b);
  a = bf16_lo(u);
  b = bf16_hi(u);
}
__device__ __forceinline__ float2 unpack_bf16(uint32_t u) { return make_float2(bf16_lo(u), bf16_hi(u)); }
// Mixed-precision arithmetic of sm_100 (PTX ISA 8.6: {add,sub,fma}.rn.f32.bf16): one operand is a bf16 HALF of a packed
// register, the result is fp32 with a single rounding -- SASS FHADD.BF16 / FHFMA.BF16 with .H0 / .H1 operand selectors,
// i.e. no shift / mask to unpack first.  d = bf16 - c, d = bf16 + c, d = bf16 * bf16 + c.
__device__ __forceinline__ float bf16_lo_sub(uint32_t u, float c) {
  const unsigned short h = (unsigned short)(u & 0xffffu);
  float d;
  asm("sub.rn.f32.bf16 %0, %1, %2;" : "=f"(d) : "h"(h), "f"(c));
  return d;
}
__device__ __forceinline__ float bf16_hi_sub(uint32_t u, float c) {
  const unsigned short h = (unsigned short)(u >> 16);
  float d;
  asm("sub.rn.f32.bf16 %0, %1, %2;" : "=f"(d) : "h"(h), "f"(c));
  return d;
}
__device__ __forceinline__ float bf16_lo_add(uint32_t u, float c) {
  const unsigned short h = (unsigned short)(u & 0xffffu);
  float d;
  asm("add.rn.f32.bf16 %0, %1, %2;" : "=f"(d) : "h"(h), "f"(c));
  return d;
}
__device__ __forceinline__ float bf16_hi_add(uint32_t u, float c) {
  const unsigned short h = (unsigned short)(u >> 16);
  float d;
  asm("add.rn.f32.bf16 %0, %1, %2;" : "=f"(d) : "h"(h), "f"(c));
  return d;
}
__device__ __forceinline__ float bf16_lo_sqacc(uint32_t u, float c) {  // lo * lo + c
  const unsigned short h = (unsigned short)(u & 0xffffu);
  float d;
  asm("fma.rn.f32.bf16 %0, %1, %1, %2;" : "=f"(d) : "h"(h), "f"(c));
  return d;
}
__device__ __forceinline__ float bf16_hi_sqacc(uint32_t u, float c) {  // hi * hi + c
  const unsigned short h = (unsigned short)(u >> 16);
  float d;
  asm("fma.rn.f32.bf16 %0, %1, %1, %2;" : "=f"(d) : "h"(h), "f"(c));
  return d;
}
// packed bf16 max (HMNMX2.BF16)
__device__ __forceinline__ uint32_t hmax2_bf16(uint32_t a, uint32_t b) {
  __nv_bfloat162 x = *reinterpret_cast<__nv_bfloat162*>(&a);
  const __nv_bfloat162 y = *reinterpret_cast<__nv_bfloat162*>(&b);
  x = __hmax2(x, y);
  return *reinterpret_cast<uint32_t*>(&x);
}

}  // namespace snapb200
