// VerticalPooling for the non-default modes (snap/models/bev_mapper.py:56-88): 'sum', 'mean', 'softmax',
// 'weighted' (softmax of log-sigmoid confidences) -- 'max' lives in lift_kernels.cu / the fused lift, 'mlp' is
// mask_rows + two tcgen05 GEMMs on the host side.  The same kernel pools over the modality axis
// (fuse_neural_maps, bev_mapper.py:225-252) when called with Z = number of modalities.
//
// One warp per cell: lane l owns channels [4l, 4l+4) of C = 128.  Pass 1 computes the Z confidence logits
// (confidence_head = Dense(C -> 1): dot -> bf16, + bias -> bf16, as nn.Dense materialises them in the feature
// dtype) and leaves logit z in lane z % 32; the masked softmax (jax.nn.softmax(where=valid_any_or_all,
// initial=0): shift by max(0, max_valid s)) is a warp reduction; pass 2 re-reads the column (L1/L2 hit) and
// accumulates sum_z w_z f_z in fp32.  Fully invalid columns are reduced over all z and then zeroed (double-where).
#include <cuda_bf16.h>
#include <math.h>

#include "common.cuh"
#include "host_common.h"

namespace snapb200 {

enum PoolMode { POOL_SUM = 1, POOL_MEAN = 2, POOL_SOFTMAX = 3, POOL_WEIGHTED = 4 };

__device__ __forceinline__ float log_sigmoid(float x) {  // jax.nn.log_sigmoid = -softplus(-x)
  return -(fmaxf(-x, 0.f) + log1pf(expf(-fabsf(x))));
}

template <int MODE>
__global__ void __launch_bounds__(256)
vertical_pool_kernel(const __nv_bfloat16* __restrict__ vol, const uint8_t* __restrict__ valid, long long cells,
                     int Z, const float* __restrict__ conf_w, float conf_b, __nv_bfloat16* __restrict__ plane,
                     uint8_t* __restrict__ pvalid, float* __restrict__ scores_out, float* __restrict__ weights_out) {
  constexpr int C = 128;
  const unsigned FULL = 0xffffffffu;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float w4[4] = {0.f, 0.f, 0.f, 0.f};
  if (MODE >= POOL_SOFTMAX) {
#pragma unroll
    for (int j = 0; j < 4; ++j) w4[j] = conf_w[lane * 4 + j];
  }
  for (long long cell = (long long)blockIdx.x * 8 + warp; cell < cells; cell += (long long)gridDim.x * 8) {
    const __nv_bfloat16* col = vol + (size_t)cell * Z * C;
    const uint8_t* vcol = valid + cell * Z;
    // validity of z = lane and z = lane + 32
    const bool v0 = lane < Z && vcol[lane] != 0;
    const bool v1 = lane + 32 < Z && vcol[lane + 32] != 0;
    const unsigned m0 = __ballot_sync(FULL, v0), m1 = __ballot_sync(FULL, v1);
    const bool any = (m0 | m1) != 0;
    // valid_any_or_all: where no z is valid every z takes part (the result is zeroed afterwards)
    const unsigned in0 = any ? m0 : __ballot_sync(FULL, lane < Z);
    const unsigned in1 = any ? m1 : __ballot_sync(FULL, lane + 32 < Z);
    float wz0 = 0.f, wz1 = 0.f;  // pooling weight of z = lane / lane + 32
    if (MODE >= POOL_SOFTMAX) {
      // ---- pass 1: logits ------------------------------------------------------------------------------
      float s0 = 0.f, s1 = 0.f;
      for (int z = 0; z < Z; ++z) {
        const uint2 u = __ldg(reinterpret_cast<const uint2*>(col + (size_t)z * C + lane * 4));
        const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y);
        float d = a.x * w4[0] + a.y * w4[1] + b.x * w4[2] + b.y * w4[3];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(FULL, d, o);
        float s = bf16_round(bf16_round(d) + conf_b);  // nn.Dense: dot -> dtype, + bias -> dtype; then .astype(f32)
        if (MODE == POOL_WEIGHTED) s = log_sigmoid(s);
        if ((z & 31) == lane) {
          if (z < 32) s0 = s; else s1 = s;
        }
      }
      if (scores_out != nullptr) {
        if (lane < Z) scores_out[cell * Z + lane] = s0;
        if (lane + 32 < Z) scores_out[cell * Z + lane + 32] = s1;
      }
      // ---- masked softmax over z ------------------------------------------------------------------------
      const bool i0 = (in0 >> lane) & 1u, i1 = (in1 >> lane) & 1u;
      float mx = fmaxf(i0 ? s0 : -INFINITY, i1 ? s1 : -INFINITY);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(FULL, mx, o));
      mx = fmaxf(mx, 0.f);  // initial=0
      const float e0 = i0 ? expf(s0 - mx) : 0.f, e1 = i1 ? expf(s1 - mx) : 0.f;
      float den = e0 + e1;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) den += __shfl_xor_sync(FULL, den, o);
      wz0 = v0 ? __fdiv_rn(e0, den) : 0.f;  // weights = where(valid, softmax, 0)
      wz1 = v1 ? __fdiv_rn(e1, den) : 0.f;
      if (weights_out != nullptr) {
        if (lane < Z) weights_out[cell * Z + lane] = wz0;
        if (lane + 32 < Z) weights_out[cell * Z + lane + 32] = wz1;
      }
    }
    // ---- pass 2: weighted sum over z ----------------------------------------------------------------------
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int z = 0; z < Z; ++z) {
      float wz;
      if (MODE == POOL_SUM || MODE == POOL_MEAN) {
        wz = (((z < 32 ? in0 : in1) >> (z & 31)) & 1u) ? 1.f : 0.f;
      } else {
        wz = __shfl_sync(FULL, z < 32 ? wz0 : wz1, z & 31);
      }
      if (wz == 0.f) continue;  // warp-uniform
      const uint2 u = __ldg(reinterpret_cast<const uint2*>(col + (size_t)z * C + lane * 4));
      const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y);
      acc[0] += a.x * wz;
      acc[1] += a.y * wz;
      acc[2] += b.x * wz;
      acc[3] += b.y * wz;
    }
    if (MODE == POOL_MEAN) {
      const float cnt = (float)(__popc(in0) + __popc(in1));
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[j] = __fdiv_rn(acc[j], cnt);
    }
    if (!any) {
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[j] = 0.f;
    }
    *reinterpret_cast<uint2*>(plane + cell * C + lane * 4) =
        make_uint2(pack_bf16(acc[0], acc[1]), pack_bf16(acc[2], acc[3]));
    if (lane == 0) pvalid[cell] = any ? 1 : 0;
  }
}

// bev_confidence (bev_mapper.py:292-295): where(valid, log_sigmoid(Dense(C -> 1)(plane)), 0), fp32.  One warp per cell.
__global__ void __launch_bounds__(256)
confidence_kernel(const __nv_bfloat16* __restrict__ plane, const uint8_t* __restrict__ valid, long long cells,
                  const float* __restrict__ conf_w, float conf_b, float* __restrict__ out) {
  constexpr int C = 128;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float w4[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) w4[j] = conf_w[lane * 4 + j];
  for (long long cell = (long long)blockIdx.x * 8 + warp; cell < cells; cell += (long long)gridDim.x * 8) {
    const uint2 u = __ldg(reinterpret_cast<const uint2*>(plane + cell * C + lane * 4));
    const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y);
    float d = a.x * w4[0] + a.y * w4[1] + b.x * w4[2] + b.y * w4[3];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
    const float s = bf16_round(bf16_round(d) + conf_b);
    if (lane == 0) out[cell] = valid[cell] ? log_sigmoid(s) : 0.f;
  }
}

// features = where(valid[..., None], features, 0) for the 'mlp' pooling mode (bev_mapper.py:75): rows of C bf16
__global__ void mask_rows_kernel(const __nv_bfloat16* __restrict__ x, const uint8_t* __restrict__ valid,
                                 long long rows, int C, __nv_bfloat16* __restrict__ y) {
  const int cv = C / 8;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * cv) return;
  const long long r = idx / cv;
  uint4 v = make_uint4(0, 0, 0, 0);
  if (valid[r]) v = __ldg(reinterpret_cast<const uint4*>(x) + idx);
  reinterpret_cast<uint4*>(y)[idx] = v;
}

// valid_any over the last axis: [cells, Z] -> [cells]
__global__ void valid_any_kernel(const uint8_t* __restrict__ valid, long long cells, int Z, uint8_t* __restrict__ out) {
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cells) return;
  uint8_t a = 0;
  for (int z = 0; z < Z; ++z) a |= valid[c * Z + z];
  out[c] = a ? 1 : 0;
}

}  // namespace snapb200

using namespace snapb200;

extern "C" {

int snapb200_vertical_pool(int mode, const void* volume, const uint8_t* valid, long long cells, int Z, int C,
                           const float* conf_w, float conf_b, void* plane, uint8_t* plane_valid, float* scores,
                           float* weights, void* stream) {
  SNAP_REQUIRE(volume && valid && plane && plane_valid, "null pointer");
  SNAP_REQUIRE(C == 128, "vertical_pool needs C == 128 (got %d)", C);
  SNAP_REQUIRE(Z >= 1 && Z <= 64, "1 <= Z <= 64 required (got %d)", Z);
  SNAP_REQUIRE(mode >= POOL_SUM && mode <= POOL_WEIGHTED, "mode must be 1 (sum), 2 (mean), 3 (softmax) or 4 (weighted)");
  SNAP_REQUIRE(mode < POOL_SOFTMAX || conf_w != nullptr, "softmax / weighted pooling need the confidence head");
  unsigned grid = (unsigned)((cells + 7) / 8);
  const unsigned cap = 16u * (unsigned)num_sms();
  if (grid > cap) grid = cap;
  cudaStream_t s = (cudaStream_t)stream;
  const __nv_bfloat16* v = (const __nv_bfloat16*)volume;
  __nv_bfloat16* p = (__nv_bfloat16*)plane;
  switch (mode) {
    case POOL_SUM: vertical_pool_kernel<POOL_SUM><<<grid, 256, 0, s>>>(v, valid, cells, Z, conf_w, conf_b, p, plane_valid, scores, weights); break;
    case POOL_MEAN: vertical_pool_kernel<POOL_MEAN><<<grid, 256, 0, s>>>(v, valid, cells, Z, conf_w, conf_b, p, plane_valid, scores, weights); break;
    case POOL_SOFTMAX: vertical_pool_kernel<POOL_SOFTMAX><<<grid, 256, 0, s>>>(v, valid, cells, Z, conf_w, conf_b, p, plane_valid, scores, weights); break;
    default: vertical_pool_kernel<POOL_WEIGHTED><<<grid, 256, 0, s>>>(v, valid, cells, Z, conf_w, conf_b, p, plane_valid, scores, weights); break;
  }
  return check_launch("vertical_pool_kernel");
}

int snapb200_confidence(const void* plane, const uint8_t* valid, long long cells, int C, const float* conf_w,
                        float conf_b, float* out, void* stream) {
  SNAP_REQUIRE(plane && valid && conf_w && out, "null pointer");
  SNAP_REQUIRE(C == 128, "confidence head needs C == 128 (got %d)", C);
  unsigned grid = (unsigned)((cells + 7) / 8);
  const unsigned cap = 16u * (unsigned)num_sms();
  if (grid > cap) grid = cap;
  confidence_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)plane, valid, cells, conf_w, conf_b, out);
  return check_launch("confidence_kernel");
}

int snapb200_mask_rows(const void* x, const uint8_t* valid, long long rows, int C, void* y, void* stream) {
  SNAP_REQUIRE(x && valid && y && C % 8 == 0, "bad arguments");
  const long long total = rows * (C / 8);
  mask_rows_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)x, valid, rows, C, (__nv_bfloat16*)y);
  return check_launch("mask_rows_kernel");
}

int snapb200_valid_any(const uint8_t* valid, long long cells, int Z, uint8_t* out, void* stream) {
  SNAP_REQUIRE(valid && out && Z >= 1, "bad arguments");
  valid_any_kernel<<<(unsigned)((cells + 255) / 256), 256, 0, (cudaStream_t)stream>>>(valid, cells, Z, out);
  return check_launch("valid_any_kernel");
}

}  // extern "C"
