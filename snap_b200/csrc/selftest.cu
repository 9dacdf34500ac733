// Hardware self-tests of descriptor semantics the hot kernels rely on (one small launch each; used by tests/).
//
// shifted_desc: a K-major SWIZZLE_128B operand whose START ROW is not a multiple of 8.  The halo form of the 3x3 conv
// keeps one [rows x 64] bf16 block per K chunk in shared memory and reads all nine taps from it by moving the A
// descriptor's start address by (tap offset) x 128 B.  Measured on B200 (tests/test_gemm_gpu.py): the hardware derives
// the swizzle phase from the absolute shared-memory address, so ANY start row works with the descriptor's matrix base
// offset (bits 49..51) left 0 -- mode 0 -- provided the block is 1024-byte aligned and was written with the same
// address-based pattern (TMA); setting the field to (address >> 7) & 7 -- mode 1 -- breaks non-multiples of 8.
#include "common.cuh"
#include "host_common.h"

namespace snapb200 {

__global__ void __launch_bounds__(128) shifted_desc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                           const __grid_constant__ CUtensorMap tmB, int shift, int mode,
                                                           float* __restrict__ out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* sA = smem;                 // [256 rows][128 B], two TMA boxes of 128 rows
  uint8_t* sB = smem + 32768;         // [64 rows][128 B]
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 32768 + 8192);
  uint32_t* tptr = reinterpret_cast<uint32_t*>(bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tptr, 64);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tptr;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(&bar[0], 32768 + 8192);
    tma_load_2d(&tmA, &bar[0], sA, 0, 0);
    tma_load_2d(&tmA, &bar[0], sA + 16384, 0, 128);
    tma_load_2d(&tmB, &bar[0], sB, 0, 0);
    mbar_wait(&bar[0], 0);
    tc_fence_after_sync();
    const uint32_t a_addr = smem_u32(sA) + (uint32_t)shift * 128u;
    uint64_t da = make_kmajor_desc<128>(a_addr);
    if (mode == 1) da |= (uint64_t)((a_addr >> 7) & 7u) << 49;   // matrix base offset
    const uint64_t db = make_kmajor_desc<128>(smem_u32(sB));
    constexpr uint32_t idesc = make_idesc_bf16_m128(64);
#pragma unroll
    for (int k = 0; k < 4; ++k) umma_bf16(tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, k != 0 ? 1u : 0u);
    umma_commit(&bar[1]);
  }
  mbar_wait(&bar[1], 0);
  tc_fence_after_sync();
  const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
  for (int c = 0; c < 4; ++c) {
    uint32_t v[16];
    tmem_ld16(taddr + (uint32_t)(c * 16), v);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; ++j) out[(size_t)(warp * 32 + lane) * 64 + c * 16 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 64);
}

}  // namespace snapb200

using namespace snapb200;

extern "C" int snapb200_selftest_shifted_desc(const void* a, const void* b, int shift, int mode, float* out, void* stream) {
  SNAP_REQUIRE(a && b && out, "null operand");
  SNAP_REQUIRE(shift >= 0 && shift <= 128 && (mode == 0 || mode == 1), "shift must be in 0..128, mode 0 or 1");
  CUtensorMap tmA, tmB;
  int rc = make_tmap_2d_bf16(&tmA, a, 256, 64, 64, 128, 64);
  if (rc) return rc;
  rc = make_tmap_2d_bf16(&tmB, b, 64, 64, 64, 64, 64);
  if (rc) return rc;
  static DynSmemState st;
  const size_t smem = 32768 + 8192 + 64;
  if (int rc2 = ensure_dyn_smem(reinterpret_cast<const void*>(&shifted_desc_kernel), smem, &st, "cudaFuncSetAttribute(selftest)"))
    return rc2;
  shifted_desc_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(tmA, tmB, shift, mode, out);
  return check_launch("shifted_desc_kernel");
}
