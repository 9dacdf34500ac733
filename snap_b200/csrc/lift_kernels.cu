// Camera -> BEV lift, v1 (unfused): gather + multi-view pooling kernel that writes the 257-wide
// statistics rows consumed by the fusion MLP (two tcgen05 GEMMs, gemm_tc.cuh), then vertical /
// modality max pooling and the matching head.
//
// Reference: snap/models/streetview_encoder.py:42-65 (projection), :69-76 + snap/utils/grids.py:116-137
// (bilinear gather), :109-124 (depth score), :141-178 (weighted pooling), snap/models/bev_mapper.py:56-88
// (vertical pooling), :225-252 (modality fusion), :284-291 + snap/models/layers.py:45-52 (matching head).
//
// The projection prologue uses explicit round-to-nearest mul/add/div in the oracle's operation order
// (no FMA contraction) so that visibility masks and tap indices are bit-exact w.r.t. oracle/geometry.py.
#include <cuda_bf16.h>
#include <math.h>
#include <string.h>

#include "common.cuh"
#include "host_common.h"
#include "lift_common.cuh"

namespace snapb200 {

// One warp per voxel; one CTA (8 warps) per (x,y) column.  D must be 128 (16 lanes x 8 channels).
__global__ void __launch_bounds__(256)
lift_gather_pool_kernel(const __grid_constant__ LiftParams P, const LiftView* __restrict__ views,
                        const __nv_bfloat16* __restrict__ fimg,
                        const float* __restrict__ xs, const float* __restrict__ ys,
                        const float* __restrict__ zs, __nv_bfloat16* __restrict__ stats,
                        uint8_t* __restrict__ valid, uint8_t* __restrict__ dbg_vis,
                        int* __restrict__ dbg_taps) {
  // camera / pose table lives in device memory (not in the launch parameters) so that a captured
  // CUDA graph can be replayed for a new scene after a plain H2D copy of the table
  __shared__ LiftView sview[LIFT_MAX_VIEWS];
  for (int i = threadIdx.x; i < P.V * (int)(sizeof(LiftView) / 4); i += blockDim.x)
    reinterpret_cast<uint32_t*>(sview)[i] = reinterpret_cast<const uint32_t*>(views)[i];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int col_id = blockIdx.x;  // x * Y + y
  const int ix = col_id / P.Y, iy = col_id - ix * P.Y;
  const float px = xs[P.xy_paired ? col_id : ix], py = ys[P.xy_paired ? col_id : iy];
  const int half = lane >> 4;  // 0: lower row tap, 1: upper row tap
  const int c8 = lane & 15;    // feature channels [8*c8, 8*c8+8)
  const float score_scale = (float)(P.S - 1);

  for (int iz = warp; iz < P.Z; iz += 8) {
    const long long n = (long long)col_id * P.Z + iz;
    const float pz = zs[iz];
    float fv[LIFT_MAX_VIEWS][8];
    float score[LIFT_MAX_VIEWS];
    unsigned vis_mask = 0;
#pragma unroll
    for (int v = 0; v < LIFT_MAX_VIEWS; ++v) {
      score[v] = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) fv[v][j] = 0.f;
    }
#pragma unroll
    for (int v = 0; v < LIFT_MAX_VIEWS; ++v) {
      if (v >= P.V) break;
      const Proj pr = project_point(sview[v], px, py, pz);
      const Taps t = make_taps(pr.row, pr.col, P.Hf, P.Wf);
      if (dbg_vis != nullptr && lane == 0) {
        dbg_vis[n * P.V + v] = pr.vis ? 1 : 0;
        dbg_taps[(n * P.V + v) * 2 + 0] = t.rlo;
        dbg_taps[(n * P.V + v) * 2 + 1] = t.clo;
      }
      if (!pr.vis) continue;  // warp-uniform
      vis_mask |= 1u << v;
      const __nv_bfloat16* img = fimg + (size_t)v * P.Hf * P.Wf * P.CF;
      // --- features: each half-warp reads one tap row, both column taps -------------------------
      const int rr = half ? t.r1 : t.r0;
      const float wr = half ? t.wr1 : __fadd_rn(1.0f, -t.wr1);
      const float wc0 = __fadd_rn(1.0f, -t.wc1);
      const uint4 ua = __ldg(reinterpret_cast<const uint4*>(img + ((size_t)rr * P.Wf + t.c0) * P.CF + c8 * 8));
      const uint4 ub = __ldg(reinterpret_cast<const uint4*>(img + ((size_t)rr * P.Wf + t.c1) * P.CF + c8 * 8));
      float fa[8], fb[8];
      unpack8(ua, fa);
      unpack8(ub, fb);
      // tap weights = row weight x column weight (one rounding each); (w_r0 f_00 + w_r1 f_01) per tap row, then the
      // two tap rows are added: the same operation sequence as the fused kernel (lift_fused.cu), pinned with
      // explicit round-to-nearest intrinsics so that both produce identical bits
      const float wx0 = __fmul_rn(wr, wc0), wx1 = __fmul_rn(wr, t.wc1);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float part = __fmaf_rn(wx1, fb[j], __fmul_rn(wx0, fa[j]));
        part = __fadd_rn(part, __shfl_xor_sync(0xffffffffu, part, 16));
        fv[v][j] = bf16_round(part);  // interpolated features are materialised in the feature dtype
      }
      // --- depth score: linear interpolation over the S log-depth bins of the 4 taps ------------
      const float d = fminf(fmaxf(pr.depth, P.depth_min), P.depth_max);
      const float tt = logf(d / P.depth_min) * P.inv_log_range;
      const float bi = tt * score_scale;  // (0.5 + t*(S-1)) - 0.5
      const float bf = floorf(bi);
      const int b0 = min(max((int)bf, 0), P.S - 1), b1 = min(max((int)bf + 1, 0), P.S - 1);
      const float wb1 = bi - bf;
      // lanes 0..7 = 4 taps x 2 bins: spatial interpolation per bin (-> bf16, as the reference
      // interpolates all S logits first), then the linear interpolation across bins (-> bf16)
      float sp = 0.f;
      if (lane < 8) {
        const int tap = lane >> 1, bsel = lane & 1;
        const int r = (tap & 2) ? t.r1 : t.r0;
        const int c = (tap & 1) ? t.c1 : t.c0;
        const float wt = ((tap & 2) ? t.wr1 : 1.0f - t.wr1) * ((tap & 1) ? t.wc1 : 1.0f - t.wc1);
        const float s = __bfloat162float(img[((size_t)r * P.Wf + c) * P.CF + P.D + (bsel ? b1 : b0)]);
        sp = wt * s;
      }
      sp += __shfl_xor_sync(0xffffffffu, sp, 2);
      sp += __shfl_xor_sync(0xffffffffu, sp, 4);
      sp = bf16_round(sp) * ((lane & 1) ? wb1 : 1.0f - wb1);
      sp += __shfl_xor_sync(0xffffffffu, sp, 1);
      sp = bf16_round(sp);
      score[v] = __shfl_sync(0xffffffffu, sp, 0);
    }
    // --- weighted pooling over views (softmax with where=valid, initial=0) ----------------------
    float mean[8], var[8], smax = -INFINITY;
#pragma unroll
    for (int j = 0; j < 8; ++j) mean[j] = var[j] = 0.f;
    if (vis_mask != 0) {
      float mx = 0.f;
#pragma unroll
      for (int v = 0; v < LIFT_MAX_VIEWS; ++v)
        if (vis_mask & (1u << v)) {
          mx = fmaxf(mx, score[v]);
          smax = fmaxf(smax, score[v]);
        }
      float wv[LIFT_MAX_VIEWS], den = 0.f;
#pragma unroll
      for (int v = 0; v < LIFT_MAX_VIEWS; ++v) {
        wv[v] = (vis_mask & (1u << v)) ? expf(score[v] - mx) : 0.f;
        den += wv[v];
      }
#pragma unroll
      for (int v = 0; v < LIFT_MAX_VIEWS; ++v) {
        wv[v] = __fdiv_rn(wv[v], den);  // true division: a single visible view gets weight exactly 1
#pragma unroll
        for (int j = 0; j < 8; ++j) mean[j] += wv[v] * fv[v][j];
      }
#pragma unroll
      for (int v = 0; v < LIFT_MAX_VIEWS; ++v) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float dd = fv[v][j] - mean[j];
          var[j] += wv[v] * dd * dd;
        }
      }
    } else {
      smax = 0.f;
    }
    // --- write the statistics row: [mean(D) | var(D)? | max(D) min(D)? | score_max | 0 ...] -------
    // (pool_multiview_features :165-177: fusion_use_variance / fusion_add_minmax select the blocks)
    __nv_bfloat16* row = stats + n * P.stats_ld;
    const int use_var = P.no_variance ? 0 : 1;
    const int off_minmax = P.D * (1 + use_var), off_score = off_minmax + (P.add_minmax ? 2 * P.D : 0);
    if (half == 0 || use_var) {
      const float* src = half ? var : mean;
      *reinterpret_cast<uint4*>(row + half * P.D + c8 * 8) =
          make_uint4(pack_bf16(src[0], src[1]), pack_bf16(src[2], src[3]), pack_bf16(src[4], src[5]),
                     pack_bf16(src[6], src[7]));
    }
    if (P.add_minmax) {  // half 0: max over the visible views, half 1: min; zero when no view sees the voxel
      float ext[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) ext[j] = half ? INFINITY : -INFINITY;
#pragma unroll
      for (int v = 0; v < LIFT_MAX_VIEWS; ++v)
        if (vis_mask & (1u << v)) {
#pragma unroll
          for (int j = 0; j < 8; ++j) ext[j] = half ? fminf(ext[j], fv[v][j]) : fmaxf(ext[j], fv[v][j]);
        }
      if (vis_mask == 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) ext[j] = 0.f;
      }
      *reinterpret_cast<uint4*>(row + off_minmax + half * P.D + c8 * 8) =
          make_uint4(pack_bf16(ext[0], ext[1]), pack_bf16(ext[2], ext[3]), pack_bf16(ext[4], ext[5]),
                     pack_bf16(ext[6], ext[7]));
    }
    const int tail_vecs = (P.stats_ld - off_score) / 8;
    for (int tv = lane; tv < tail_vecs; tv += 32) {
      uint4 z = make_uint4(0, 0, 0, 0);
      if (tv == 0) z.x = pack_bf16(smax, 0.f);
      *reinterpret_cast<uint4*>(row + off_score + tv * 8) = z;
    }
    if (lane == 0) valid[n] = vis_mask != 0 ? 1 : 0;
  }
}

// max over the Z axis where valid; zero (and invalid) where no z is valid.  [cells, Z, C] -> [cells, C]
__global__ void vertical_max_kernel(const __nv_bfloat16* __restrict__ vol, const uint8_t* __restrict__ valid,
                                    long long cells, int Z, int C, __nv_bfloat16* __restrict__ plane,
                                    uint8_t* __restrict__ pvalid) {
  const int cv = C / 8;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= cells * cv) return;
  const int c8 = (int)(idx % cv);
  const long long cell = idx / cv;
  float m[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) m[j] = -INFINITY;
  bool any = false;
  for (int z = 0; z < Z; ++z) {
    if (!valid[cell * Z + z]) continue;
    any = true;
    float f[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(vol + ((size_t)cell * Z + z) * C + c8 * 8)), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) m[j] = fmaxf(m[j], f[j]);
  }
  if (!any) {
#pragma unroll
    for (int j = 0; j < 8; ++j) m[j] = 0.f;
  }
  *reinterpret_cast<uint4*>(plane + cell * C + c8 * 8) =
      make_uint4(pack_bf16(m[0], m[1]), pack_bf16(m[2], m[3]), pack_bf16(m[4], m[5]), pack_bf16(m[6], m[7]));
  if (c8 == 0) pvalid[cell] = any ? 1 : 0;
}

// Dense(C -> DM) + bias, (bf16 round), L2-normalise (eps 1e-5), mask.  One warp per cell, DM == 32.
__global__ void __launch_bounds__(256)
match_head_kernel(const __nv_bfloat16* __restrict__ plane, const uint8_t* __restrict__ valid,
                  long long cells, int C, const float* __restrict__ kernel /*[C,32]*/,
                  const float* __restrict__ bias, __nv_bfloat16* __restrict__ out) {
  extern __shared__ float wsm[];  // [C][32]
  for (int i = threadIdx.x; i < C * 32; i += blockDim.x) wsm[i] = kernel[i];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float b = bias[lane];
  for (long long cell = (long long)blockIdx.x * 8 + warp; cell < cells; cell += (long long)gridDim.x * 8) {
    float acc = 0.f;
    for (int k0 = 0; k0 < C; k0 += 32) {
      const float xk = __bfloat162float(plane[cell * C + k0 + lane]);
#pragma unroll 8
      for (int k = 0; k < 32; ++k) acc += __shfl_sync(0xffffffffu, xk, k) * wsm[(k0 + k) * 32 + lane];
    }
    const float y = bf16_round(bf16_round(acc) + b);  // nn.Dense: dot -> dtype, + bias -> dtype
    float ss = y * y;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float nrm = sqrtf(ss);
    float z = (nrm < 1e-5f) ? 0.f : y / nrm;
    if (!valid[cell]) z = 0.f;
    out[cell * 32 + lane] = __float2bfloat16(z);
  }
}

// General matching head (bev_mapper.py:284-291 with any matching_dim <= 256, optional normalisation): one warp per cell,
// lane l computes outputs l, l + 32, ...; the kernel [C, DM] is read through L1.  The default configuration
// (matching_dim 32, normalised) takes match_head_kernel above.
__global__ void __launch_bounds__(256)
match_head_general_kernel(const __nv_bfloat16* __restrict__ plane, const uint8_t* __restrict__ valid, long long cells,
                          int C, const float* __restrict__ kernel, const float* __restrict__ bias, int DM, int normalize,
                          __nv_bfloat16* __restrict__ out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (long long cell = (long long)blockIdx.x * 8 + warp; cell < cells; cell += (long long)gridDim.x * 8) {
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int k0 = 0; k0 < C; k0 += 32) {
      const float xk = k0 + lane < C ? __bfloat162float(plane[cell * C + k0 + lane]) : 0.f;
      for (int k = 0; k < 32 && k0 + k < C; ++k) {
        const float x = __shfl_sync(0xffffffffu, xk, k);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int n = lane + 32 * j;
          if (n < DM) acc[j] += x * __ldg(kernel + (size_t)(k0 + k) * DM + n);
        }
      }
    }
    float y[8], ss = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = lane + 32 * j;
      y[j] = n < DM ? bf16_round(bf16_round(acc[j]) + __ldg(bias + n)) : 0.f;  // nn.Dense: dot -> dtype, + bias -> dtype
      ss += y[j] * y[j];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float nrm = sqrtf(ss);
    const bool ok = valid[cell] != 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = lane + 32 * j;
      if (n >= DM) continue;
      float z = y[j];
      if (normalize) z = (nrm < 1e-5f) ? 0.f : y[j] / nrm;   // layers.normalize (layers.py:45-52)
      out[cell * DM + n] = __float2bfloat16(ok ? z : 0.f);
    }
  }
}

// modality fusion: max over up to 3 planes where valid (VerticalPooling('max') over the modality axis)
__global__ void fuse_max_kernel(const __nv_bfloat16* __restrict__ a, const uint8_t* __restrict__ va,
                                const __nv_bfloat16* __restrict__ b, const uint8_t* __restrict__ vb,
                                long long cells, int C, __nv_bfloat16* __restrict__ out,
                                uint8_t* __restrict__ vout) {
  const int cv = C / 8;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= cells * cv) return;
  const int c8 = (int)(idx % cv);
  const long long cell = idx / cv;
  const bool oa = va[cell] != 0, ob = (vb == nullptr) ? true : vb[cell] != 0;
  float fa[8], fb[8], m[8];
  unpack8(__ldg(reinterpret_cast<const uint4*>(a + cell * C + c8 * 8)), fa);
  unpack8(__ldg(reinterpret_cast<const uint4*>(b + cell * C + c8 * 8)), fb);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float v = -INFINITY;
    if (oa) v = fmaxf(v, fa[j]);
    if (ob) v = fmaxf(v, fb[j]);
    m[j] = (oa || ob) ? v : 0.f;
  }
  *reinterpret_cast<uint4*>(out + cell * C + c8 * 8) =
      make_uint4(pack_bf16(m[0], m[1]), pack_bf16(m[2], m[3]), pack_bf16(m[4], m[5]), pack_bf16(m[6], m[7]));
  if (c8 == 0) vout[cell] = (oa || ob) ? 1 : 0;
}

}  // namespace snapb200

using namespace snapb200;

extern "C" {

/* Host-side mirror of LiftParams/LiftView: the caller fills SnapLiftParams (include/snapb200.h). */
int snapb200_lift_gather_pool(const SnapLiftParams* q, const SnapLiftView* views, const void* fimg,
                              const float* xs, const float* ys,
                              const float* zs, void* stats, uint8_t* valid, uint8_t* dbg_vis,
                              int* dbg_taps, void* stream) {
  SNAP_REQUIRE(q && views && fimg && xs && ys && zs && stats && valid, "null pointer");
  SNAP_REQUIRE(q->V >= 1 && q->V <= LIFT_MAX_VIEWS, "1 <= V <= %d required (got %d)", LIFT_MAX_VIEWS, q->V);
  SNAP_REQUIRE(q->D == 128, "feature_dim must be 128 (got %d)", q->D);
  SNAP_REQUIRE(q->S >= 2 && q->CF == q->D + q->S && q->CF % 8 == 0, "bad channel split");
  {
    const int width = q->D * (1 + (q->no_variance ? 0 : 1) + (q->add_minmax ? 2 : 0));
    SNAP_REQUIRE(q->stats_ld % 32 == 0 && q->stats_ld >= width + 8 && q->stats_ld <= width + 256,
                 "stats_ld must be a multiple of 32 in [%d, %d] (got %d)", width + 8, width + 256, q->stats_ld);
  }
  SNAP_REQUIRE((dbg_vis == nullptr) == (dbg_taps == nullptr), "debug outputs come as a pair");
  static_assert(sizeof(SnapLiftView) == sizeof(LiftView), "SnapLiftView layout");
  static_assert(sizeof(SnapLiftParams) == sizeof(LiftParams), "SnapLiftParams layout");
  LiftParams P;
  memcpy(&P, q, sizeof(P));
  const unsigned grid = (unsigned)(q->X * q->Y);
  lift_gather_pool_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
      P, reinterpret_cast<const LiftView*>(views), (const __nv_bfloat16*)fimg, xs, ys, zs,
      (__nv_bfloat16*)stats, valid, dbg_vis, dbg_taps);
  return check_launch("lift_gather_pool_kernel");
}

int snapb200_vertical_max(const void* volume, const uint8_t* valid, long long cells, int Z, int C,
                          void* plane, uint8_t* plane_valid, void* stream) {
  SNAP_REQUIRE(volume && valid && plane && plane_valid && C % 8 == 0, "bad arguments");
  const long long total = cells * (C / 8);
  vertical_max_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)volume, valid, cells, Z, C, (__nv_bfloat16*)plane, plane_valid);
  return check_launch("vertical_max_kernel");
}

int snapb200_match_head(const void* plane, const uint8_t* valid, long long cells, int C,
                        const float* kernel, const float* bias, int dm, void* out, void* stream) {
  SNAP_REQUIRE(plane && valid && kernel && bias && out, "null pointer");
  SNAP_REQUIRE(dm == 32 && C % 32 == 0 && C <= 256, "matching head needs matching_dim 32, C %% 32 == 0, C <= 256");
  const size_t smem = (size_t)C * 32 * sizeof(float);
  unsigned grid = (unsigned)((cells + 7) / 8);
  if (grid > 4u * 148u) grid = 4u * 148u;
  match_head_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)plane, valid, cells, C, kernel, bias, (__nv_bfloat16*)out);
  return check_launch("match_head_kernel");
}

int snapb200_match_head_ex(const void* plane, const uint8_t* valid, long long cells, int C, const float* kernel,
                           const float* bias, int dm, int normalize, void* out, void* stream) {
  SNAP_REQUIRE(plane && valid && kernel && bias && out, "null pointer");
  SNAP_REQUIRE(dm >= 1 && dm <= 256 && C >= 1, "matching head needs 1 <= matching_dim <= 256 (got %d)", dm);
  if (dm == 32 && normalize && C % 32 == 0 && C <= 256)
    return snapb200_match_head(plane, valid, cells, C, kernel, bias, dm, out, stream);
  unsigned grid = (unsigned)((cells + 7) / 8);
  if (grid > 8u * 148u) grid = 8u * 148u;
  match_head_general_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)plane, valid, cells, C, kernel, bias,
                                                                    dm, normalize, (__nv_bfloat16*)out);
  return check_launch("match_head_general_kernel");
}

int snapb200_fuse_max(const void* a, const uint8_t* va, const void* b, const uint8_t* vb, long long cells,
                      int C, void* out, uint8_t* vout, void* stream) {
  SNAP_REQUIRE(a && va && b && out && vout && C % 8 == 0, "bad arguments");
  const long long total = cells * (C / 8);
  fuse_max_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)a, va, (const __nv_bfloat16*)b, vb, cells, C, (__nv_bfloat16*)out, vout);
  return check_launch("fuse_max_kernel");
}

}  // extern "C"
