// Backward of the (unfused, all-views, default-statistics) camera -> BEV lift: the "true scatter-add" of SURVEY 8(f)1.
//
//   vertical max (snap/models/bev_mapper.py:56-88, 'max')          -> vertical_max_bwd_kernel
//   pooling (snap/models/streetview_encoder.py:141-178, weighted)  \
//   depth-score interpolation (:109-124)                            > lift_gather_pool_bwd_kernel
//   bilinear gather (:69-76, snap/utils/grids.py:116-137)          /
//
// lift_gather_pool_bwd_kernel mirrors lift_gather_pool_kernel (lift_kernels.cu): one warp per voxel, the two half-warps
// own the two tap rows, a lane owns 8 feature channels.  It RECOMPUTES the forward of its voxel (projection, taps,
// interpolated features, depth scores, soft-max weights, mean) from the projected feature maps -- nothing per
// (voxel, view) is saved by the forward -- and evaluates the closed form of tools/design/backward_formulas.py
// (pool_multiview_backward + lift_gather_backward, checked against autograd on the CPU):
//     w = softmax_valid(s), mean = sum w f, var = sum w (f - mean)^2, smax = max_valid s
//     d f_k = w_k dmean + 2 w_k (f_k - mean) dvar
//     d w_k = f_k . dmean + (f_k - mean)^2 . dvar          d s_k = w_k (d w_k - sum_j w_j d w_j) + [s_k = smax] dsmax / #ties
// and scatter-adds w_tap * d f_k into the D feature channels of the four taps and w_tap * (1 - wb1 | wb1) * d s_k into
// the two scale-bin channels: fp32 atomics into a gradient image gimg f32 [V, Hf, Wf, D + S] (zeroed by the caller).
// Rounding of the forward (bf16 materialisation points) is treated as the identity (straight-through), the geometry
// (projection, tap positions, depth) is not differentiated -- the reference's camera poses are data, not parameters.
// Float atomics make the summation order (not the set of addends) run-dependent: gradients agree to fp32 round-off.
#include <cuda_bf16.h>
#include <math.h>
#include <string.h>

#include "common.cuh"
#include "host_common.h"
#include "lift_common.cuh"

namespace snapb200 {

__device__ __forceinline__ void atomic_add8(float* dst, const float (&v)[8]) {
#if __CUDA_ARCH__ >= 900
  atomicAdd(reinterpret_cast<float4*>(dst), make_float4(v[0], v[1], v[2], v[3]));
  atomicAdd(reinterpret_cast<float4*>(dst) + 1, make_float4(v[4], v[5], v[6], v[7]));
#else
#pragma unroll
  for (int j = 0; j < 8; ++j) atomicAdd(dst + j, v[j]);
#endif
}

__global__ void __launch_bounds__(256)
lift_gather_pool_bwd_kernel(const __grid_constant__ LiftParams P, const LiftView* __restrict__ views,
                            const __nv_bfloat16* __restrict__ fimg, const float* __restrict__ xs,
                            const float* __restrict__ ys, const float* __restrict__ zs,
                            const __nv_bfloat16* __restrict__ dstats, float* __restrict__ gimg) {
  __shared__ LiftView sview[LIFT_MAX_VIEWS];
  for (int i = threadIdx.x; i < P.V * (int)(sizeof(LiftView) / 4); i += blockDim.x)
    reinterpret_cast<uint32_t*>(sview)[i] = reinterpret_cast<const uint32_t*>(views)[i];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int col_id = blockIdx.x;
  const int ix = col_id / P.Y, iy = col_id - ix * P.Y;
  const float px = xs[P.xy_paired ? col_id : ix], py = ys[P.xy_paired ? col_id : iy];
  const int half = lane >> 4, c8 = lane & 15;
  const float score_scale = (float)(P.S - 1);

  for (int iz = warp; iz < P.Z; iz += 8) {
    const long long n = (long long)col_id * P.Z + iz;
    const float pz = zs[iz];
    float fv[LIFT_MAX_VIEWS][8], score[LIFT_MAX_VIEWS];
    unsigned vis_mask = 0;
#pragma unroll
    for (int v = 0; v < LIFT_MAX_VIEWS; ++v) {
      score[v] = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) fv[v][j] = 0.f;
    }
    // ---- forward recomputation (same operation sequence as lift_gather_pool_kernel) ---------------------------
#pragma unroll
    for (int v = 0; v < LIFT_MAX_VIEWS; ++v) {
      if (v >= P.V) break;
      const Proj pr = project_point(sview[v], px, py, pz);
      if (!pr.vis) continue;  // warp-uniform
      const Taps t = make_taps(pr.row, pr.col, P.Hf, P.Wf);
      vis_mask |= 1u << v;
      const __nv_bfloat16* img = fimg + (size_t)v * P.Hf * P.Wf * P.CF;
      const int rr = half ? t.r1 : t.r0;
      const float wr = half ? t.wr1 : __fadd_rn(1.0f, -t.wr1);
      const float wc0 = __fadd_rn(1.0f, -t.wc1);
      const uint4 ua = __ldg(reinterpret_cast<const uint4*>(img + ((size_t)rr * P.Wf + t.c0) * P.CF + c8 * 8));
      const uint4 ub = __ldg(reinterpret_cast<const uint4*>(img + ((size_t)rr * P.Wf + t.c1) * P.CF + c8 * 8));
      float fa[8], fb[8];
      unpack8(ua, fa);
      unpack8(ub, fb);
      const float wx0 = __fmul_rn(wr, wc0), wx1 = __fmul_rn(wr, t.wc1);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float part = __fmaf_rn(wx1, fb[j], __fmul_rn(wx0, fa[j]));
        part = __fadd_rn(part, __shfl_xor_sync(0xffffffffu, part, 16));
        fv[v][j] = bf16_round(part);
      }
      const float d = fminf(fmaxf(pr.depth, P.depth_min), P.depth_max);
      const float tt = logf(d / P.depth_min) * P.inv_log_range;
      const float bi = tt * score_scale;
      const float bf = floorf(bi);
      const int b0 = min(max((int)bf, 0), P.S - 1), b1 = min(max((int)bf + 1, 0), P.S - 1);
      const float wb1 = bi - bf;
      float sp = 0.f;
      if (lane < 8) {
        const int tap = lane >> 1, bsel = lane & 1;
        const int r = (tap & 2) ? t.r1 : t.r0;
        const int c = (tap & 1) ? t.c1 : t.c0;
        const float wt = ((tap & 2) ? t.wr1 : 1.0f - t.wr1) * ((tap & 1) ? t.wc1 : 1.0f - t.wc1);
        sp = wt * __bfloat162float(img[((size_t)r * P.Wf + c) * P.CF + P.D + (bsel ? b1 : b0)]);
      }
      sp += __shfl_xor_sync(0xffffffffu, sp, 2);
      sp += __shfl_xor_sync(0xffffffffu, sp, 4);
      sp = bf16_round(sp) * ((lane & 1) ? wb1 : 1.0f - wb1);
      sp += __shfl_xor_sync(0xffffffffu, sp, 1);
      sp = bf16_round(sp);
      score[v] = __shfl_sync(0xffffffffu, sp, 0);
    }
    if (vis_mask == 0) continue;  // statistics of unseen voxels are the constant 0 (:177): no gradient
    float mx = 0.f, smax = -INFINITY, nmax = 0.f;
#pragma unroll
    for (int v = 0; v < LIFT_MAX_VIEWS; ++v)
      if (vis_mask & (1u << v)) {
        mx = fmaxf(mx, score[v]);
        smax = fmaxf(smax, score[v]);
      }
#pragma unroll
    for (int v = 0; v < LIFT_MAX_VIEWS; ++v)  // jnp.max splits the cotangent of score_max evenly among tied views
      if ((vis_mask & (1u << v)) && score[v] == smax) nmax += 1.f;
    float wv[LIFT_MAX_VIEWS], den = 0.f, mean[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) mean[j] = 0.f;
#pragma unroll
    for (int v = 0; v < LIFT_MAX_VIEWS; ++v) {
      wv[v] = (vis_mask & (1u << v)) ? expf(score[v] - mx) : 0.f;
      den += wv[v];
    }
#pragma unroll
    for (int v = 0; v < LIFT_MAX_VIEWS; ++v) {
      wv[v] = __fdiv_rn(wv[v], den);
#pragma unroll
      for (int j = 0; j < 8; ++j) mean[j] += wv[v] * fv[v][j];
    }
    // ---- cotangents of this voxel's statistics row [mean(D) | var(D) | score_max] ---------------------------------
    const __nv_bfloat16* drow = dstats + n * P.stats_ld;
    float dmean[8], dvar[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(drow + c8 * 8)), dmean);
    unpack8(__ldg(reinterpret_cast<const uint4*>(drow + P.D + c8 * 8)), dvar);
    const float dsmax = __bfloat162float(drow[2 * P.D]);
    // d w_k: dot products over the D channels = 16 lanes x 8 channels (both half-warps hold the same features)
    float dw[LIFT_MAX_VIEWS], wdw = 0.f;
#pragma unroll
    for (int v = 0; v < LIFT_MAX_VIEWS; ++v) {
      float s = 0.f;
      if (vis_mask & (1u << v)) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float dd = fv[v][j] - mean[j];
          s += fv[v][j] * dmean[j] + dd * dd * dvar[j];
        }
#pragma unroll
        for (int o = 1; o < 16; o <<= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      }
      dw[v] = s;
      wdw += wv[v] * s;
    }
    // ---- scatter ------------------------------------------------------------------------------------------------------
#pragma unroll
    for (int v = 0; v < LIFT_MAX_VIEWS; ++v) {
      if (v >= P.V) break;
      if (!(vis_mask & (1u << v))) continue;
      const float ds = wv[v] * (dw[v] - wdw) + (score[v] == smax ? dsmax / nmax : 0.f);
      // the tap geometry again (cheap, and keeps no per-view state alive across the pooling math)
      const Proj pr = project_point(sview[v], px, py, pz);
      const Taps t = make_taps(pr.row, pr.col, P.Hf, P.Wf);
      float* gv = gimg + (size_t)v * P.Hf * P.Wf * P.CF;
      const int rr = half ? t.r1 : t.r0;
      const float wr = half ? t.wr1 : __fadd_rn(1.0f, -t.wr1);
      const float wx0 = __fmul_rn(wr, __fadd_rn(1.0f, -t.wc1)), wx1 = __fmul_rn(wr, t.wc1);
      float g0[8], g1[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float df = wv[v] * (dmean[j] + 2.f * (fv[v][j] - mean[j]) * dvar[j]);
        g0[j] = wx0 * df;
        g1[j] = wx1 * df;
      }
      atomic_add8(gv + ((size_t)rr * P.Wf + t.c0) * P.CF + c8 * 8, g0);
      atomic_add8(gv + ((size_t)rr * P.Wf + t.c1) * P.CF + c8 * 8, g1);
      if (lane < 8) {  // 4 taps x 2 bins of the depth-score interpolation
        const float d = fminf(fmaxf(pr.depth, P.depth_min), P.depth_max);
        const float bi = logf(d / P.depth_min) * P.inv_log_range * score_scale;
        const float bf = floorf(bi);
        const int b0 = min(max((int)bf, 0), P.S - 1), b1 = min(max((int)bf + 1, 0), P.S - 1);
        const float wb1 = bi - bf;
        const int tap = lane >> 1, bsel = lane & 1;
        const int r = (tap & 2) ? t.r1 : t.r0;
        const int c = (tap & 1) ? t.c1 : t.c0;
        const float wt = ((tap & 2) ? t.wr1 : 1.0f - t.wr1) * ((tap & 1) ? t.wc1 : 1.0f - t.wc1);
        atomicAdd(gv + ((size_t)r * P.Wf + c) * P.CF + P.D + (bsel ? b1 : b0), wt * (bsel ? wb1 : 1.0f - wb1) * ds);
      }
    }
  }
}

// dvol[cell, z, :] = dplane[cell, :] / (number of maximal valid z) where vol[cell, z, c] is a maximum over the valid z,
// else 0 (jnp.max splits the cotangent evenly among ties; invalid voxels are masked to -inf before the max and cells
// without a valid z output the constant 0, bev_mapper.py:58-60,80-86).  One thread per (cell, 8 channels).
__global__ void vertical_max_bwd_kernel(const __nv_bfloat16* __restrict__ vol, const uint8_t* __restrict__ valid,
                                        const __nv_bfloat16* __restrict__ dplane, long long cells, int Z, int C,
                                        __nv_bfloat16* __restrict__ dvol) {
  const int cv = C / 8;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= cells * cv) return;
  const int c8 = (int)(idx % cv);
  const long long cell = idx / cv;
  float m[8], cnt[8], g[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    m[j] = -INFINITY;
    cnt[j] = 0.f;
  }
  for (int z = 0; z < Z; ++z) {
    if (!valid[cell * Z + z]) continue;
    float f[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(vol + ((size_t)cell * Z + z) * C + c8 * 8)), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (f[j] > m[j]) {
        m[j] = f[j];
        cnt[j] = 1.f;
      } else if (f[j] == m[j]) {
        cnt[j] += 1.f;
      }
    }
  }
  unpack8(__ldg(reinterpret_cast<const uint4*>(dplane + cell * C + c8 * 8)), g);
#pragma unroll
  for (int j = 0; j < 8; ++j) g[j] = cnt[j] > 0.f ? g[j] / cnt[j] : 0.f;
  for (int z = 0; z < Z; ++z) {
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = 0.f;
    if (valid[cell * Z + z]) {
      float f[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(vol + ((size_t)cell * Z + z) * C + c8 * 8)), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = f[j] == m[j] ? g[j] : 0.f;
    }
    *reinterpret_cast<uint4*>(dvol + ((size_t)cell * Z + z) * C + c8 * 8) =
        make_uint4(pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]), pack_bf16(o[4], o[5]), pack_bf16(o[6], o[7]));
  }
}

}  // namespace snapb200

using namespace snapb200;

extern "C" {

int snapb200_lift_gather_pool_backward(const SnapLiftParams* q, const SnapLiftView* views, const void* fimg,
                                       const float* xs, const float* ys, const float* zs, const void* dstats,
                                       float* gimg, void* stream) {
  SNAP_REQUIRE(q && views && fimg && xs && ys && zs && dstats && gimg, "null pointer");
  SNAP_REQUIRE(q->V >= 1 && q->V <= LIFT_MAX_VIEWS, "1 <= V <= %d required (got %d)", LIFT_MAX_VIEWS, q->V);
  SNAP_REQUIRE(q->D == 128, "feature_dim must be 128 (got %d)", q->D);
  SNAP_REQUIRE(q->S >= 2 && q->CF == q->D + q->S && q->CF % 8 == 0, "bad channel split");
  SNAP_REQUIRE(!q->no_variance && !q->add_minmax, "only the default statistics [mean | var | score_max] have a backward");
  SNAP_REQUIRE(q->stats_ld % 8 == 0 && q->stats_ld >= 2 * q->D + 8, "stats_ld too small");
  static_assert(sizeof(SnapLiftView) == sizeof(LiftView), "SnapLiftView layout");
  static_assert(sizeof(SnapLiftParams) == sizeof(LiftParams), "SnapLiftParams layout");
  LiftParams P;
  memcpy(&P, q, sizeof(P));
  lift_gather_pool_bwd_kernel<<<(unsigned)(q->X * q->Y), 256, 0, (cudaStream_t)stream>>>(
      P, reinterpret_cast<const LiftView*>(views), (const __nv_bfloat16*)fimg, xs, ys, zs,
      (const __nv_bfloat16*)dstats, gimg);
  return check_launch("lift_gather_pool_bwd_kernel");
}

int snapb200_vertical_max_backward(const void* vol, const uint8_t* valid, const void* dplane, long long cells, int Z,
                                   int C, void* dvol, void* stream) {
  SNAP_REQUIRE(vol && valid && dplane && dvol && cells >= 1 && Z >= 1 && C % 8 == 0, "bad arguments");
  const long long total = cells * (C / 8);
  vertical_max_bwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)vol, valid, (const __nv_bfloat16*)dplane, cells, Z, C, (__nv_bfloat16*)dvol);
  return check_launch("vertical_max_bwd_kernel");
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------------------
// Matching head (snap/models/bev_mapper.py:284-291: Dense 128 -> 32, snap/models/layers.py:45-52 L2 normalisation, mask)
// and modality fusion (bev_mapper.py:225-252, 'max' over the stacked modality axis) backward.
// ---------------------------------------------------------------------------------------------------------------------
namespace snapb200 {

// Warp per cell, lane = output channel.  Recomputes y = bf16(bf16(x K) + b) and n = |y| like match_head_kernel, then
//   dy = (dz - z (z . dz)) / n   for valid cells with n >= eps (z = y / n), else 0
// written as bf16 rows [cells, 32]; the Dense backward proper (dK = x^T dy, db, dx = dy K^T) runs on the split-K and
// GEMM kernels.
__global__ void __launch_bounds__(256)
match_head_bwd_kernel(const __nv_bfloat16* __restrict__ plane, const uint8_t* __restrict__ valid, long long cells, int C,
                      const float* __restrict__ kernel /*[C,32]*/, const float* __restrict__ bias,
                      const __nv_bfloat16* __restrict__ dout, __nv_bfloat16* __restrict__ dy) {
  extern __shared__ float wsm_b[];  // [C][32]
  for (int i = threadIdx.x; i < C * 32; i += blockDim.x) wsm_b[i] = kernel[i];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float b = bias[lane];
  for (long long cell = (long long)blockIdx.x * 8 + warp; cell < cells; cell += (long long)gridDim.x * 8) {
    float g = 0.f;
    if (valid[cell]) {  // warp-uniform
      float acc = 0.f;
      for (int k0 = 0; k0 < C; k0 += 32) {
        const float xk = __bfloat162float(plane[cell * C + k0 + lane]);
#pragma unroll 8
        for (int k = 0; k < 32; ++k) acc += __shfl_sync(0xffffffffu, xk, k) * wsm_b[(k0 + k) * 32 + lane];
      }
      const float y = bf16_round(bf16_round(acc) + b);
      const float dz = __bfloat162float(dout[cell * 32 + lane]);
      float ss = y * y, zd = y * dz;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        ss += __shfl_xor_sync(0xffffffffu, ss, o);
        zd += __shfl_xor_sync(0xffffffffu, zd, o);
      }
      const float nrm = sqrtf(ss);
      if (nrm >= 1e-5f) {
        const float z = y / nrm;
        g = (dz - z * (zd / nrm)) / nrm;  // zd / nrm = z . dz
      }
    }
    dy[cell * 32 + lane] = __float2bfloat16(g);
  }
}

// da / db from dout for out = max over the valid ones of (a, b); a tie between two valid modalities splits evenly
__global__ void fuse_max_bwd_kernel(const __nv_bfloat16* __restrict__ a, const uint8_t* __restrict__ va,
                                    const __nv_bfloat16* __restrict__ b, const uint8_t* __restrict__ vb,
                                    const __nv_bfloat16* __restrict__ dout, long long cells, int C,
                                    __nv_bfloat16* __restrict__ da, __nv_bfloat16* __restrict__ db) {
  const int cv = C / 8;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= cells * cv) return;
  const int c8 = (int)(idx % cv);
  const long long cell = idx / cv;
  const bool oa = va[cell] != 0, ob = (vb == nullptr) ? true : vb[cell] != 0;
  float fa[8], fb[8], g[8], ga[8], gb[8];
  unpack8(__ldg(reinterpret_cast<const uint4*>(a + cell * C + c8 * 8)), fa);
  unpack8(__ldg(reinterpret_cast<const uint4*>(b + cell * C + c8 * 8)), fb);
  unpack8(__ldg(reinterpret_cast<const uint4*>(dout + cell * C + c8 * 8)), g);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    ga[j] = gb[j] = 0.f;
    if (oa && ob) {
      if (fa[j] > fb[j]) ga[j] = g[j];
      else if (fb[j] > fa[j]) gb[j] = g[j];
      else ga[j] = gb[j] = 0.5f * g[j];
    } else if (oa) {
      ga[j] = g[j];
    } else if (ob) {
      gb[j] = g[j];
    }
  }
  *reinterpret_cast<uint4*>(da + cell * C + c8 * 8) =
      make_uint4(pack_bf16(ga[0], ga[1]), pack_bf16(ga[2], ga[3]), pack_bf16(ga[4], ga[5]), pack_bf16(ga[6], ga[7]));
  *reinterpret_cast<uint4*>(db + cell * C + c8 * 8) =
      make_uint4(pack_bf16(gb[0], gb[1]), pack_bf16(gb[2], gb[3]), pack_bf16(gb[4], gb[5]), pack_bf16(gb[6], gb[7]));
}

}  // namespace snapb200

extern "C" {

int snapb200_match_head_backward(const void* plane, const uint8_t* valid, long long cells, int C, const float* kernel,
                                 const float* bias, const void* dout, void* dy, void* stream) {
  SNAP_REQUIRE(plane && valid && kernel && bias && dout && dy, "null pointer");
  SNAP_REQUIRE(cells >= 1 && C % 32 == 0 && C >= 32 && C <= 256, "C must be a multiple of 32 in [32, 256]");
  const size_t smem = (size_t)C * 32 * sizeof(float);
  long long blocks = (cells + 7) / 8;
  const long long cap = 8LL * snapb200::num_sms();
  if (blocks > cap) blocks = cap;
  snapb200::match_head_bwd_kernel<<<(unsigned)blocks, 256, smem, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)plane, valid, cells, C, kernel, bias, (const __nv_bfloat16*)dout, (__nv_bfloat16*)dy);
  return snapb200::check_launch("match_head_bwd_kernel");
}

int snapb200_fuse_max_backward(const void* a, const uint8_t* va, const void* b, const uint8_t* vb, const void* dout,
                               long long cells, int C, void* da, void* db, void* stream) {
  SNAP_REQUIRE(a && va && b && dout && da && db && C % 8 == 0 && cells >= 1, "bad arguments");
  const long long total = cells * (C / 8);
  snapb200::fuse_max_bwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)a, va, (const __nv_bfloat16*)b, vb, (const __nv_bfloat16*)dout, cells, C,
      (__nv_bfloat16*)da, (__nv_bfloat16*)db);
  return snapb200::check_launch("fuse_max_bwd_kernel");
}

}  // extern "C"
