// Backward of the sampling localizer's loss (what jax.grad computes through snap/models/bev_localizer.py:156-160,
// 183-216,244-262 and snap/models/pose_estimation.py:50-85):
//
//   nll_b = logsumexp_{k kept} s_bk - s_b0 (mean over the batch, trainer.py:221)   -> loc_nll_bwd_kernel
//   s_bp  = sum_n valid . bilinear(scale w_n sim[b,n], T_p q_n / cell)              -> loc_pose_scoring_bwd_kernel
//   sim   = relu(bf16(f_q . f_m))                                                   -> relu mask folded into the kernel
//                                                                                      above; the two products with the
//                                                                                      cotangent run on the GEMM engine /
//                                                                                      the split-K kernel
//
// The pose samples themselves (soft-max sampling, RANSAC, Kabsch) carry no gradient in the reference either: the poses
// are produced by integer draws (jax.random.choice) and enter the scores as constants.
//
// loc_pose_scoring_bwd_kernel: one CTA per (point, example).  The cotangent map of the point's similarity map (H x W
// fp32, 64 KB for 128 x 128) lives in shared memory; the CTA's threads walk over the P poses, each pose adds its four
// bilinear tap weights x dscore_p x point_scale_n with shared-memory atomics; the map is then written ONCE, as bf16 and
// masked by the ReLU of the forward (sim > 0), so there is no global atomic traffic at all.
#include <cuda_bf16.h>
#include <math.h>

#include "common.cuh"
#include "host_common.h"

namespace snapb200 {

__global__ void __launch_bounds__(256)
loc_nll_bwd_kernel(const float* __restrict__ scores, const float* __restrict__ dr, const float* __restrict__ dt, int P1,
                   int use_remove, float dr_min, float dt_min, float inv_B, float* __restrict__ dscores,
                   float* __restrict__ dtemp) {
  __shared__ float red[8];
  __shared__ float s_bc;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* sc = scores + (size_t)b * P1;
  auto removed = [&](int k) {
    return use_remove && k > 0 && dr[(size_t)b * P1 + k] < dr_min && dt[(size_t)b * P1 + k] < dt_min;
  };
  auto block_reduce = [&](float v, bool is_max) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float t = __shfl_xor_sync(0xffffffffu, v, o);
      v = is_max ? fmaxf(v, t) : v + t;
    }
    if (lane == 0) red[warp] = v;
    __syncthreads();
    if (tid == 0) {
      float r = red[0];
      for (int w = 1; w < 8; ++w) r = is_max ? fmaxf(r, red[w]) : r + red[w];
      s_bc = r;
    }
    __syncthreads();
    const float out = s_bc;
    __syncthreads();
    return out;
  };
  float mx = -INFINITY;
  for (int k = tid; k < P1; k += 256)
    if (!removed(k)) mx = fmaxf(mx, sc[k]);
  mx = block_reduce(mx, true);
  float se = 0.f;
  for (int k = tid; k < P1; k += 256)
    if (!removed(k)) se += expf(sc[k] - mx);
  se = block_reduce(se, false);
  float dT = 0.f;
  for (int k = tid; k < P1; k += 256) {
    float g = 0.f;
    if (!removed(k)) g = inv_B * (expf(sc[k] - mx) / se - (k == 0 ? 1.f : 0.f));
    dscores[(size_t)b * P1 + k] = g;
    dT += g * sc[k];  // scores are proportional to exp(temperature): d s / d T = s
  }
  dT = block_reduce(dT, false);
  if (tid == 0 && dtemp != nullptr) dtemp[b] = dT;
}

// dsim bf16 [B, N, H*W] = relu-masked cotangent of the similarities from dscores f32 [B, P]
__global__ void __launch_bounds__(256)
loc_pose_scoring_bwd_kernel(const __nv_bfloat16* __restrict__ sim, const float* __restrict__ point_scale,
                            const float* __restrict__ i_xy, int i_xy_batched, const uint8_t* __restrict__ valid_j,
                            const float* __restrict__ poses, const float* __restrict__ dscores, int N, int H, int W, int P,
                            float inv_cell, int mask_oob, int relu_mask, int dsim_rows, __nv_bfloat16* __restrict__ dsim) {
  extern __shared__ float gmap[];  // [H * W]
  const int n = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
  const int HW = H * W;
  const size_t row = ((size_t)b * N + n) * HW;            // row of sim
  const size_t drow = ((size_t)b * dsim_rows + n) * HW;   // row of dsim (dsim_rows >= N rows per example)
  const float ps = point_scale[(size_t)b * N + n];
  if (ps == 0.f) {  // invalid query point: contributes to no score
    for (int i = tid; i < HW / 8; i += 256) reinterpret_cast<uint4*>(dsim + drow)[i] = make_uint4(0u, 0u, 0u, 0u);
    return;
  }
  for (int i = tid; i < HW; i += 256) gmap[i] = 0.f;
  __syncthreads();
  const float* xy = i_xy + (i_xy_batched ? ((size_t)b * N + n) * 2 : (size_t)n * 2);
  const float qx = xy[0], qy = xy[1];
  const uint8_t* vj = valid_j ? valid_j + (size_t)b * HW : nullptr;
  for (int p = tid; p < P; p += 256) {
    const float g = dscores[(size_t)b * P + p];
    if (g == 0.f) continue;
    const float* T = poses + ((size_t)b * P + p) * 3;
    float sn, cs;
    sincosf(T[0], &sn, &cs);
    const float u = (cs * qx - sn * qy + T[1]) * inv_cell, v = (sn * qx + cs * qy + T[2]) * inv_cell;
    const float cu = u - 0.5f, cw = v - 0.5f;
    const float fu = floorf(cu), fw = floorf(cw);
    const float wu1 = cu - fu, ww1 = cw - fw;
    const int r0 = min(max((int)fu, 0), H - 1), r1 = min(max((int)fu + 1, 0), H - 1);
    const int c0 = min(max((int)fw, 0), W - 1), c1 = min(max((int)fw + 1, 0), W - 1);
    if (mask_oob) {  // pose_estimation.py:78-79: the point must fall inside the map and on valid map cells
      bool ok = u >= 0.f && u < (float)H && v >= 0.f && v < (float)W;
      if (ok && vj) ok = vj[r0 * W + c0] && vj[r0 * W + c1] && vj[r1 * W + c0] && vj[r1 * W + c1];
      if (!ok) continue;
    }
    const float gs = g * ps;
    atomicAdd(&gmap[r0 * W + c0], gs * (1.f - wu1) * (1.f - ww1));
    atomicAdd(&gmap[r0 * W + c1], gs * (1.f - wu1) * ww1);
    atomicAdd(&gmap[r1 * W + c0], gs * wu1 * (1.f - ww1));
    atomicAdd(&gmap[r1 * W + c1], gs * wu1 * ww1);
  }
  __syncthreads();
  for (int i = tid; i < HW / 8; i += 256) {
    const uint4 s = __ldg(reinterpret_cast<const uint4*>(sim + row) + i);
    const uint32_t su[4] = {s.x, s.y, s.z, s.w};
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float a = gmap[i * 8 + 2 * j], c = gmap[i * 8 + 2 * j + 1];
      if (relu_mask) {
        if (!(bf16_lo(su[j]) > 0.f)) a = 0.f;
        if (!(bf16_hi(su[j]) > 0.f)) c = 0.f;
      }
      o[j] = pack_bf16(a, c);
    }
    reinterpret_cast<uint4*>(dsim + drow)[i] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

}  // namespace snapb200

using namespace snapb200;

extern "C" {

int snapb200_loc_nll_backward(const float* scores, const float* dr_samples, const float* dt_samples, int B, int P1,
                              int use_remove, float dr_min, float dt_min, float* dscores, float* dtemperature,
                              void* stream) {
  SNAP_REQUIRE(scores && dscores && B >= 1 && P1 >= 1, "bad arguments");
  SNAP_REQUIRE(!use_remove || (dr_samples && dt_samples), "threshold_remove_accurate_poses needs the sample errors");
  loc_nll_bwd_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(scores, dr_samples, dt_samples, P1, use_remove, dr_min, dt_min,
                                                          1.f / (float)B, dscores, dtemperature);
  return check_launch("loc_nll_bwd_kernel");
}

int snapb200_loc_pose_scoring_backward(const SnapLocScoreParams* p, const void* sim, const float* point_scale,
                                       const float* i_xy, const uint8_t* valid_j, const float* poses,
                                       const float* dscores, int relu_mask, int dsim_rows, void* dsim, void* stream) {
  SNAP_REQUIRE(p && sim && point_scale && i_xy && poses && dscores && dsim, "null pointer");
  SNAP_REQUIRE(p->B >= 1 && p->N >= 1 && p->P >= 1 && p->H >= 1 && p->W >= 1 && p->cell_size > 0.f, "empty problem");
  SNAP_REQUIRE((p->H * p->W) % 8 == 0, "H * W must be a multiple of 8");
  SNAP_REQUIRE(dsim_rows >= p->N, "dsim_rows (rows per example of dsim) must be >= N");
  const size_t smem = (size_t)p->H * p->W * sizeof(float);
  SNAP_REQUIRE(smem <= 200 * 1024, "cotangent map of %d x %d does not fit in shared memory", p->H, p->W);
  static DynSmemState smem_state;
  if (smem > 48 * 1024)
    if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(&loc_pose_scoring_bwd_kernel), smem, &smem_state,
                                 "cudaFuncSetAttribute(loc_pose_scoring_bwd)"))
      return rc;
  dim3 grid((unsigned)p->N, (unsigned)p->B);
  loc_pose_scoring_bwd_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)sim, point_scale, i_xy, p->i_xy_batched, p->mask_out_of_bounds ? valid_j : nullptr, poses,
      dscores, p->N, p->H, p->W, p->P, 1.f / p->cell_size, p->mask_out_of_bounds, relu_mask, dsim_rows, (__nv_bfloat16*)dsim);
  return check_launch("loc_pose_scoring_bwd_kernel");
}

}  // extern "C"
