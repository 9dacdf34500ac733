// Fused camera->BEV lift, second generation: ONE persistent, WARP-SPECIALISED sm_100a kernel
// for all scenes of a batch.
//
//   producers (10 warps)   frustum-plane culling -> bit-exact projection of the surviving (voxel, view) pairs ->
//                          compaction of the visible voxels into 128-row tiles -> bilinear gather (cp.async into a
//                          shared-memory staging ring, one pair ahead) + depth score + multi-view softmax pooling ->
//                          statistics tile A[128 x 256] in 128B-swizzled shared memory (double-buffered)
//   MMA issuer (1 thread)  tcgen05 GEMM1 (A x W1^T -> TMEM, W1 streamed as 8 chunks of 16 KB), GEMM2 (H x W2^T)
//   TMA producer (1 thread) weight chunks through a 2-slot ring
//   consumers (4 warps, one per TMEM lane quadrant)  epilogue 1 (rank-1 score term, bias, ReLU -> H written over A),
//                          epilogue 2 (bias) -> staged volume rows -> running max over z per BEV column ->
//                          plane[B, X*Y, 128] + valid[B, X*Y]
//
// The first generation (lift_fused.cu) walked all 16 worker warps through fill -> gather -> epilogues -> z-max in lockstep,
// so the latency-bound phases (gather: L2 round trips) never overlapped the issue-bound ones (epilogues): 41 % issue-slot
// utilisation at 4.5 warps per scheduler (profiles/r01_v20_lift_fused_ncu_full.txt).  Here the two halves of the pipeline
// run concurrently on different tiles, the tap loads are asynchronous copies (no registers held across the round trip),
// the exact projection only runs for pairs that pass four plane tests, and one launch covers the whole batch.
//
// Numerics are those of lift_fused.cu / lift_kernels.cu (same rounding points, same operation order); visibility and tap
// indices are bit-exact with the oracle (tests/test_lift_gpu.py).  The culling is conservative: a pair is only skipped
// when it is outside the image by more than the worst-case rounding error of both this test and the reference's own
// projection (margins below), so the visible set is unchanged.
#include <cuda_bf16.h>
#include <math.h>
#include <string.h>

#include "common.cuh"
#include "host_common.h"
#include "lift_common.cuh"

#define F2_NP_VALUE 10
#define F2_NS f2_np10
#include "lift_fused2_impl.cuh"
#undef F2_NP_VALUE
#undef F2_NS
#undef F2_MARK
#define F2_NP_VALUE 12
#define F2_NS f2_np12
#include "lift_fused2_impl.cuh"
#undef F2_NP_VALUE
#undef F2_NS

#include <stdlib.h>

using namespace snapb200;

// Producer-warp count: 10 (16 warps, 128 registers per thread, no spills) unless SNAPB200_LIFT_NP=12 (18 warps, 96
// registers).  Measured on a B200 at the cfg2 shape: 0.242 vs 0.245 ms per tile (profiles/r02_notes.md) -- two more
// producer warps buy nothing once the consumers no longer compete for issue slots.
static int lift2_np() {
  static const int np = [] {
    const char* v = getenv("SNAPB200_LIFT_NP");
    return v != nullptr && atoi(v) == 12 ? 12 : 10;
  }();
  return np;
}

extern "C" size_t snapb200_lift_fused_batched_scratch_bytes(void) {
  const size_t a = f2_np10::scratch_bytes_needed(), b = f2_np12::scratch_bytes_needed();
  return a > b ? a : b;
}

extern "C" int snapb200_lift_fused_batched(const SnapLiftParams* q, int B, const SnapLiftView* views, long long views_stride,
                                           const void* fimg, long long fimg_stride, const float* xs, const float* ys,
                                           const float* zs, long long zs_stride, const void* w1t, long long ldw1,
                                           const float* w256, const float* b1, const void* w2t, const float* b2,
                                           void* plane, uint8_t* pvalid, int* col_counter, void* scratch,
                                           size_t scratch_bytes, void* stream) {
  if (lift2_np() == 10)
    return f2_np10::launch(q, B, views, views_stride, fimg, fimg_stride, xs, ys, zs, zs_stride, w1t, ldw1, w256, b1, w2t, b2,
                           plane, pvalid, col_counter, scratch, scratch_bytes, stream);
  return f2_np12::launch(q, B, views, views_stride, fimg, fimg_stride, xs, ys, zs, zs_stride, w1t, ldw1, w256, b1, w2t, b2,
                         plane, pvalid, col_counter, scratch, scratch_bytes, stream);
}
