// Camera -> BEV lift when a scene has MORE views than `top_k_view_selection` (production scenes carry 10-20
// views, snap/data/types.py:58): per voxel the k nearest visible views are selected and only those are sampled.
//
// Reference: snap/models/streetview_encoder.py:127-138 (view_selection: distance to the camera centre, +inf
// where invisible, lax.top_k(-dist) -> ties and the -inf fill go to the LOWER view index), :66 + :241-249
// (gather of p2d / visible / depth by the selected indices), :80-105 (interpolate_views_selective: the
// coordinates are CAST TO THE FEATURE DTYPE (:88), clamped to [0, size-1], lower = floor, upper = lower + 1,
// four taps summed in the order (0,0),(0,1),(1,0),(1,1) in the feature dtype), :109-124 (depth score),
// :141-178 (weighted pooling), :275-279 (max_view_distance).
//
// One warp per voxel, one CTA (8 warps) per (x, y) column.  Lane v projects view v (V <= 32) with the same
// bit-exact prologue as the all-views kernels; the k rounds of the selection are a warp min-reduction over
// the IEEE bit pattern of the (non-negative) distance followed by a ballot (lowest lane wins ties).  The two
// half-warps then sample two selected views at a time (16 lanes x 8 channels = 128 features each; lanes 0/1
// of each half also sample the two depth-bin logits the score needs), exchange them with one shuffle and pool.
#include <cuda_bf16.h>
#include <math.h>
#include <string.h>

#include "common.cuh"
#include "host_common.h"
#include "lift_common.cuh"

namespace snapb200 {

constexpr int SELECT_MAX_VIEWS = 32;

struct SelTaps {
  int r0, r1, c0, c1;          // lower / upper tap indices (upper clamped: it can only be out of range with weight 0)
  float w00, w01, w10, w11;    // products of the bf16 weights, rounded to bf16 (:102)
};

// interpolate_views_selective (:88-98) in bf16 arithmetic: every intermediate is materialised in the feature dtype
__device__ __forceinline__ void sel_axis(float p, int size, int& lo_i, int& up_i, float& w_lo, float& w_up) {
  float q = bf16_round(p);                       // point.astype(arrays.dtype)
  q = bf16_round(q - 0.5f);                      // point - 0.5
  q = fminf(q, bf16_round((float)(size - 1)));   // jnp.minimum(., size - 1): the int32 size is promoted to bf16
  q = fmaxf(q, 0.f);
  const float lo = floorf(q);
  lo_i = (int)lo;
  w_up = bf16_round(q - lo);
  w_lo = bf16_round(1.0f - w_up);
  up_i = min(lo_i + 1, size - 1);
  lo_i = min(lo_i, size - 1);
}

__device__ __forceinline__ SelTaps make_sel_taps(float row, float col, int Hf, int Wf) {
  SelTaps t;
  float wr0, wr1, wc0, wc1;
  sel_axis(row, Hf, t.r0, t.r1, wr0, wr1);
  sel_axis(col, Wf, t.c0, t.c1, wc0, wc1);
  t.w00 = bf16_round(wr0 * wc0);
  t.w01 = bf16_round(wr0 * wc1);
  t.w10 = bf16_round(wr1 * wc0);
  t.w11 = bf16_round(wr1 * wc1);
  return t;
}

// sum(values) of :102-105 for one channel: ((w00 a00 + w01 a01) + w10 a10) + w11 a11, each op rounded to bf16
__device__ __forceinline__ float sel_sum(const SelTaps& t, float a00, float a01, float a10, float a11) {
  float s = bf16_round(t.w00 * a00);
  s = bf16_round(s + bf16_round(t.w01 * a01));
  s = bf16_round(s + bf16_round(t.w10 * a10));
  s = bf16_round(s + bf16_round(t.w11 * a11));
  return s;
}

__global__ void __launch_bounds__(256)
lift_select_pool_kernel(const __grid_constant__ LiftParams P, const int K, const float max_view_distance,
                        const LiftView* __restrict__ views, const float* __restrict__ centers,
                        const __nv_bfloat16* __restrict__ fimg, const float* __restrict__ xs,
                        const float* __restrict__ ys, const float* __restrict__ zs,
                        __nv_bfloat16* __restrict__ stats, uint8_t* __restrict__ valid,
                        int* __restrict__ dbg_idx, uint8_t* __restrict__ dbg_vis, int* __restrict__ dbg_taps) {
  __shared__ LiftView sview[SELECT_MAX_VIEWS];
  __shared__ float scen[SELECT_MAX_VIEWS * 3];
  for (int i = threadIdx.x; i < P.V * (int)(sizeof(LiftView) / 4); i += blockDim.x)
    reinterpret_cast<uint32_t*>(sview)[i] = reinterpret_cast<const uint32_t*>(views)[i];
  for (int i = threadIdx.x; i < P.V * 3; i += blockDim.x) scen[i] = centers[i];
  __syncthreads();
  const unsigned FULL = 0xffffffffu;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int col_id = blockIdx.x;  // x * Y + y
  const int ix = col_id / P.Y, iy = col_id - ix * P.Y;
  const float px = xs[P.xy_paired ? col_id : ix], py = ys[P.xy_paired ? col_id : iy];
  const int half = lane >> 4;  // which of the two selected views of a round this half-warp samples
  const int c8 = lane & 15;    // feature channels [8*c8, 8*c8+8)
  const float score_scale = (float)(P.S - 1);

  for (int iz = warp; iz < P.Z; iz += 8) {
    const long long n = (long long)col_id * P.Z + iz;
    const float pz = zs[iz];
    // ---- lane v: project view v, distance to its centre (:131-133) -----------------------------------
    Proj pr;
    pr.row = pr.col = pr.depth = 0.f;
    pr.vis = false;
    unsigned key = 0xffffffffu;  // lanes without a view never win
    if (lane < P.V) {
      pr = project_point(sview[lane], px, py, pz);
      const float dx = __fadd_rn(px, -scen[lane * 3 + 0]);
      const float dy = __fadd_rn(py, -scen[lane * 3 + 1]);
      const float dz = __fadd_rn(pz, -scen[lane * 3 + 2]);
      const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
      key = pr.vis ? __float_as_uint(__fsqrt_rn(d2)) : 0x7f800000u;  // +inf where not visible
    }
    const float min_dist = __uint_as_float(__reduce_min_sync(FULL, key));
    // ---- k rounds of arg-min, ties -> lower view index (lax.top_k on -dist) ----------------
    int sel[LIFT_MAX_VIEWS];
    bool taken = false;
#pragma unroll
    for (int k = 0; k < LIFT_MAX_VIEWS; ++k) {
      sel[k] = 0;
      if (k >= K) continue;
      const unsigned m = __reduce_min_sync(FULL, taken ? 0xffffffffu : key);
      const unsigned b = __ballot_sync(FULL, !taken && key == m);
      sel[k] = __ffs(b) - 1;
      if (lane == sel[k]) taken = true;
    }
    // ---- sample the selected views, two per round ------------------------------------------------------
    float fv[LIFT_MAX_VIEWS][8];
    float score[LIFT_MAX_VIEWS];
    unsigned vis_mask = 0;
#pragma unroll
    for (int k = 0; k < LIFT_MAX_VIEWS; ++k) {
      score[k] = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) fv[k][j] = 0.f;
    }
#pragma unroll
    for (int i = 0; i < LIFT_MAX_VIEWS / 2; ++i) {
      if (2 * i >= K) break;
      const int k = 2 * i + half;
      const bool act = k < K;
      const int src = half ? sel[2 * i + 1] : sel[2 * i];
      const float row = __shfl_sync(FULL, pr.row, src);
      const float col = __shfl_sync(FULL, pr.col, src);
      const float depth = __shfl_sync(FULL, pr.depth, src);
      const bool vis = __shfl_sync(FULL, pr.vis ? 1 : 0, src) != 0 && act;
      const SelTaps t = make_sel_taps(row, col, P.Hf, P.Wf);
      if (dbg_idx != nullptr && c8 == 0 && act) {
        dbg_idx[n * K + k] = src;
        dbg_vis[n * K + k] = vis ? 1 : 0;
        dbg_taps[(n * K + k) * 2 + 0] = t.r0;
        dbg_taps[(n * K + k) * 2 + 1] = t.c0;
      }
      float f[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = 0.f;
      float sp = 0.f;
      if (vis) {  // an invisible selected view has pooling weight 0 (:160): its samples are never used
        const __nv_bfloat16* img = fimg + (size_t)src * P.Hf * P.Wf * P.CF;
        const __nv_bfloat16* p00 = img + ((size_t)t.r0 * P.Wf + t.c0) * P.CF;
        const __nv_bfloat16* p01 = img + ((size_t)t.r0 * P.Wf + t.c1) * P.CF;
        const __nv_bfloat16* p10 = img + ((size_t)t.r1 * P.Wf + t.c0) * P.CF;
        const __nv_bfloat16* p11 = img + ((size_t)t.r1 * P.Wf + t.c1) * P.CF;
        float a00[8], a01[8], a10[8], a11[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(p00 + c8 * 8)), a00);
        unpack8(__ldg(reinterpret_cast<const uint4*>(p01 + c8 * 8)), a01);
        unpack8(__ldg(reinterpret_cast<const uint4*>(p10 + c8 * 8)), a10);
        unpack8(__ldg(reinterpret_cast<const uint4*>(p11 + c8 * 8)), a11);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = sel_sum(t, a00[j], a01[j], a10[j], a11[j]);
        // depth score (:109-124): the two log-depth bins around clip(depth), sampled like any other channel
        const float d = fminf(fmaxf(depth, P.depth_min), P.depth_max);
        const float tt = logf(d / P.depth_min) * P.inv_log_range;
        const float bi = tt * score_scale;  // (0.5 + t*(S-1)) - 0.5
        const float bf = floorf(bi);
        const int b0 = min(max((int)bf, 0), P.S - 1), b1 = min(max((int)bf + 1, 0), P.S - 1);
        const float wb1 = bi - bf;
        if (c8 < 2) {
          const int ch = P.D + (c8 ? b1 : b0);
          const float s = sel_sum(t, __bfloat162float(p00[ch]), __bfloat162float(p01[ch]),
                                  __bfloat162float(p10[ch]), __bfloat162float(p11[ch]));
          sp = s * (c8 ? wb1 : 1.0f - wb1);
        }
      }
      sp += __shfl_xor_sync(FULL, sp, 1);
      sp = bf16_round(sp);
      const unsigned bal = __ballot_sync(FULL, vis);
      // exchange: afterwards every lane holds both views of the round for its 8 channels
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float o = __shfl_xor_sync(FULL, f[j], 16);
        fv[2 * i][j] = half ? o : f[j];
        fv[2 * i + 1][j] = half ? f[j] : o;
      }
      score[2 * i] = __shfl_sync(FULL, sp, 0);
      score[2 * i + 1] = __shfl_sync(FULL, sp, 16);
      if (bal & 1u) vis_mask |= 1u << (2 * i);
      if (bal & 0x10000u) vis_mask |= 1u << (2 * i + 1);
    }
    // ---- weighted pooling over the selected views (softmax with where=valid, initial=0) ---------------
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
        wv[v] = __fdiv_rn(wv[v], den);
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
    // ---- statistics row [mean(D) | var(D) | score_max | 0 ...]; half 0 writes the mean, half 1 the variance
    __nv_bfloat16* out_row = stats + n * P.stats_ld;
    const float* srcv = half ? var : mean;
    *reinterpret_cast<uint4*>(out_row + half * P.D + c8 * 8) =
        make_uint4(pack_bf16(srcv[0], srcv[1]), pack_bf16(srcv[2], srcv[3]), pack_bf16(srcv[4], srcv[5]),
                   pack_bf16(srcv[6], srcv[7]));
    const int tail_vecs = (P.stats_ld - 2 * P.D) / 8;
    if (lane < tail_vecs) {
      uint4 z = make_uint4(0, 0, 0, 0);
      if (lane == 0) z.x = pack_bf16(smax, 0.f);
      *reinterpret_cast<uint4*>(out_row + 2 * P.D + lane * 8) = z;
    }
    if (lane == 0) {
      bool ok = vis_mask != 0;
      if (max_view_distance >= 0.f) ok = ok && (min_dist <= max_view_distance);  // :275-279
      valid[n] = ok ? 1 : 0;
    }
  }
}

}  // namespace snapb200

using namespace snapb200;

extern "C" int snapb200_lift_select_pool(const SnapLiftParams* q, int top_k, float max_view_distance,
                                         const SnapLiftView* views, const float* view_centers, const void* fimg,
                                         const float* xs, const float* ys, const float* zs, void* stats,
                                         uint8_t* valid, int* dbg_idx, uint8_t* dbg_vis, int* dbg_taps,
                                         void* stream) {
  SNAP_REQUIRE(q && views && view_centers && fimg && xs && ys && zs && stats && valid, "null pointer");
  SNAP_REQUIRE(top_k >= 1 && top_k <= LIFT_MAX_VIEWS, "1 <= top_k <= %d required (got %d)", LIFT_MAX_VIEWS, top_k);
  SNAP_REQUIRE(q->V > top_k && q->V <= SELECT_MAX_VIEWS,
               "view selection needs top_k < V <= %d (got V=%d, top_k=%d); V <= top_k is the all-views path",
               SELECT_MAX_VIEWS, q->V, top_k);
  SNAP_REQUIRE(q->D == 128, "feature_dim must be 128 (got %d)", q->D);
  SNAP_REQUIRE(!q->no_variance && !q->add_minmax, "the view-selection kernel implements the default statistics only");
  SNAP_REQUIRE(q->S >= 2 && q->CF == q->D + q->S && q->CF % 8 == 0, "bad channel split");
  SNAP_REQUIRE(q->stats_ld % 32 == 0 && q->stats_ld >= 2 * q->D + 8 && q->stats_ld <= 2 * q->D + 256,
               "bad stats_ld %d", q->stats_ld);
  SNAP_REQUIRE((dbg_idx == nullptr) == (dbg_vis == nullptr) && (dbg_idx == nullptr) == (dbg_taps == nullptr),
               "debug outputs come as a triple");
  static_assert(SNAPB200_MAX_SELECT_VIEWS == SELECT_MAX_VIEWS, "header constant");
  LiftParams P;
  memcpy(&P, q, sizeof(P));
  const unsigned grid = (unsigned)(q->X * q->Y);
  lift_select_pool_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
      P, top_k, max_view_distance, reinterpret_cast<const LiftView*>(views), view_centers,
      (const __nv_bfloat16*)fimg, xs, ys, zs, (__nv_bfloat16*)stats, valid, dbg_idx, dbg_vis, dbg_taps);
  return check_launch("lift_select_pool_kernel");
}

// ---------------------------------------------------------------------------------------------------------------------
// Backward of lift_select_pool_kernel (the V > top_k path): same warp / half-warp organisation.  Pass 1 repeats the
// forward of the voxel (projection by lane, k rounds of arg-min, sampling of the selected views two per round, soft-max
// pooling); pass 2 walks over the rounds again and scatter-adds, per selected visible view, w_tap * d f_k into the D
// feature channels of its four taps (w_tap = the bf16-rounded tap weights of :102) and w_tap * (1 - wb1 | wb1) * d s_k
// into the two scale-bin channels of the gradient image gimg f32 [V, Hf, Wf, D + S].  Closed forms as in
// lift_backward.cu; rounding points are straight-through, the selection and the geometry carry no gradient.
// ---------------------------------------------------------------------------------------------------------------------
namespace snapb200 {

__device__ __forceinline__ void sel_atomic_add8(float* dst, float w, const float (&v)[8]) {
#if __CUDA_ARCH__ >= 900
  atomicAdd(reinterpret_cast<float4*>(dst), make_float4(w * v[0], w * v[1], w * v[2], w * v[3]));
  atomicAdd(reinterpret_cast<float4*>(dst) + 1, make_float4(w * v[4], w * v[5], w * v[6], w * v[7]));
#else
#pragma unroll
  for (int j = 0; j < 8; ++j) atomicAdd(dst + j, w * v[j]);
#endif
}

__global__ void __launch_bounds__(256)
lift_select_pool_bwd_kernel(const __grid_constant__ LiftParams P, const int K, const LiftView* __restrict__ views,
                            const float* __restrict__ centers, const __nv_bfloat16* __restrict__ fimg,
                            const float* __restrict__ xs, const float* __restrict__ ys, const float* __restrict__ zs,
                            const __nv_bfloat16* __restrict__ dstats, float* __restrict__ gimg) {
  __shared__ LiftView sview[SELECT_MAX_VIEWS];
  __shared__ float scen[SELECT_MAX_VIEWS * 3];
  for (int i = threadIdx.x; i < P.V * (int)(sizeof(LiftView) / 4); i += blockDim.x)
    reinterpret_cast<uint32_t*>(sview)[i] = reinterpret_cast<const uint32_t*>(views)[i];
  for (int i = threadIdx.x; i < P.V * 3; i += blockDim.x) scen[i] = centers[i];
  __syncthreads();
  const unsigned FULL = 0xffffffffu;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int col_id = blockIdx.x;
  const int ix = col_id / P.Y, iy = col_id - ix * P.Y;
  const float px = xs[P.xy_paired ? col_id : ix], py = ys[P.xy_paired ? col_id : iy];
  const int half = lane >> 4, c8 = lane & 15;
  const float score_scale = (float)(P.S - 1);

  for (int iz = warp; iz < P.Z; iz += 8) {
    const long long n = (long long)col_id * P.Z + iz;
    const float pz = zs[iz];
    Proj pr;
    pr.row = pr.col = pr.depth = 0.f;
    pr.vis = false;
    unsigned key = 0xffffffffu;
    if (lane < P.V) {
      pr = project_point(sview[lane], px, py, pz);
      const float dx = __fadd_rn(px, -scen[lane * 3 + 0]);
      const float dy = __fadd_rn(py, -scen[lane * 3 + 1]);
      const float dz = __fadd_rn(pz, -scen[lane * 3 + 2]);
      const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
      key = pr.vis ? __float_as_uint(__fsqrt_rn(d2)) : 0x7f800000u;
    }
    int sel[LIFT_MAX_VIEWS];
    bool taken = false;
#pragma unroll
    for (int k = 0; k < LIFT_MAX_VIEWS; ++k) {
      sel[k] = 0;
      if (k >= K) continue;
      const unsigned m = __reduce_min_sync(FULL, taken ? 0xffffffffu : key);
      const unsigned b = __ballot_sync(FULL, !taken && key == m);
      sel[k] = __ffs(b) - 1;
      if (lane == sel[k]) taken = true;
    }
    // ---- pass 1: the forward's samples ---------------------------------------------------------------------------
    float fv[LIFT_MAX_VIEWS][8], score[LIFT_MAX_VIEWS];
    unsigned vis_mask = 0;
#pragma unroll
    for (int k = 0; k < LIFT_MAX_VIEWS; ++k) {
      score[k] = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) fv[k][j] = 0.f;
    }
#pragma unroll
    for (int i = 0; i < LIFT_MAX_VIEWS / 2; ++i) {
      if (2 * i >= K) break;
      const int k = 2 * i + half;
      const bool act = k < K;
      const int src = half ? sel[2 * i + 1] : sel[2 * i];
      const float row = __shfl_sync(FULL, pr.row, src);
      const float col = __shfl_sync(FULL, pr.col, src);
      const float depth = __shfl_sync(FULL, pr.depth, src);
      const bool vis = __shfl_sync(FULL, pr.vis ? 1 : 0, src) != 0 && act;
      const SelTaps t = make_sel_taps(row, col, P.Hf, P.Wf);
      float f[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = 0.f;
      float sp = 0.f;
      if (vis) {
        const __nv_bfloat16* img = fimg + (size_t)src * P.Hf * P.Wf * P.CF;
        const __nv_bfloat16* p00 = img + ((size_t)t.r0 * P.Wf + t.c0) * P.CF;
        const __nv_bfloat16* p01 = img + ((size_t)t.r0 * P.Wf + t.c1) * P.CF;
        const __nv_bfloat16* p10 = img + ((size_t)t.r1 * P.Wf + t.c0) * P.CF;
        const __nv_bfloat16* p11 = img + ((size_t)t.r1 * P.Wf + t.c1) * P.CF;
        float a00[8], a01[8], a10[8], a11[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(p00 + c8 * 8)), a00);
        unpack8(__ldg(reinterpret_cast<const uint4*>(p01 + c8 * 8)), a01);
        unpack8(__ldg(reinterpret_cast<const uint4*>(p10 + c8 * 8)), a10);
        unpack8(__ldg(reinterpret_cast<const uint4*>(p11 + c8 * 8)), a11);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = sel_sum(t, a00[j], a01[j], a10[j], a11[j]);
        const float d = fminf(fmaxf(depth, P.depth_min), P.depth_max);
        const float bi = logf(d / P.depth_min) * P.inv_log_range * score_scale;
        const float bf = floorf(bi);
        const int b0 = min(max((int)bf, 0), P.S - 1), b1 = min(max((int)bf + 1, 0), P.S - 1);
        const float wb1 = bi - bf;
        if (c8 < 2) {
          const int ch = P.D + (c8 ? b1 : b0);
          const float s = sel_sum(t, __bfloat162float(p00[ch]), __bfloat162float(p01[ch]),
                                  __bfloat162float(p10[ch]), __bfloat162float(p11[ch]));
          sp = s * (c8 ? wb1 : 1.0f - wb1);
        }
      }
      sp += __shfl_xor_sync(FULL, sp, 1);
      sp = bf16_round(sp);
      const unsigned bal = __ballot_sync(FULL, vis);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float o = __shfl_xor_sync(FULL, f[j], 16);
        fv[2 * i][j] = half ? o : f[j];
        fv[2 * i + 1][j] = half ? f[j] : o;
      }
      score[2 * i] = __shfl_sync(FULL, sp, 0);
      score[2 * i + 1] = __shfl_sync(FULL, sp, 16);
      if (bal & 1u) vis_mask |= 1u << (2 * i);
      if (bal & 0x10000u) vis_mask |= 1u << (2 * i + 1);
    }
    if (vis_mask == 0) continue;  // warp-uniform: the statistics of an unseen voxel are the constant 0
    // ---- pooling backward (closed form) ----------------------------------------------------------------------------
    float mx = 0.f, smax = -INFINITY, nmax = 0.f;
#pragma unroll
    for (int v = 0; v < LIFT_MAX_VIEWS; ++v)
      if (vis_mask & (1u << v)) {
        mx = fmaxf(mx, score[v]);
        smax = fmaxf(smax, score[v]);
      }
#pragma unroll
    for (int v = 0; v < LIFT_MAX_VIEWS; ++v)
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
    const __nv_bfloat16* drow = dstats + n * P.stats_ld;
    float dmean[8], dvar[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(drow + c8 * 8)), dmean);
    unpack8(__ldg(reinterpret_cast<const uint4*>(drow + P.D + c8 * 8)), dvar);
    const float dsmax = __bfloat162float(drow[2 * P.D]);
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
        for (int o = 1; o < 16; o <<= 1) s += __shfl_xor_sync(FULL, s, o);   // both half-warps hold all views
      }
      dw[v] = s;
      wdw += wv[v] * s;
    }
    // ---- pass 2: scatter, two selected views per round (one per half-warp) -----------------------------------------------
#pragma unroll
    for (int i = 0; i < LIFT_MAX_VIEWS / 2; ++i) {
      if (2 * i >= K) break;
      const int k = 2 * i + half;
      const bool act = k < K;
      const int src = half ? sel[2 * i + 1] : sel[2 * i];
      const float row = __shfl_sync(FULL, pr.row, src);
      const float col = __shfl_sync(FULL, pr.col, src);
      const float depth = __shfl_sync(FULL, pr.depth, src);
      const bool vis = __shfl_sync(FULL, pr.vis ? 1 : 0, src) != 0 && act;
      if (!vis) continue;  // uniform within a half-warp; no warp-wide operation follows in this iteration
      const SelTaps t = make_sel_taps(row, col, P.Hf, P.Wf);
      const float wk = half ? wv[2 * i + 1] : wv[2 * i];
      const float sk = half ? score[2 * i + 1] : score[2 * i];
      const float dwk = half ? dw[2 * i + 1] : dw[2 * i];
      const float ds = wk * (dwk - wdw) + (sk == smax ? dsmax / nmax : 0.f);
      float df[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float fk = half ? fv[2 * i + 1][j] : fv[2 * i][j];
        df[j] = wk * (dmean[j] + 2.f * (fk - mean[j]) * dvar[j]);
      }
      float* gv = gimg + (size_t)src * P.Hf * P.Wf * P.CF;
      float* g00 = gv + ((size_t)t.r0 * P.Wf + t.c0) * P.CF;
      float* g01 = gv + ((size_t)t.r0 * P.Wf + t.c1) * P.CF;
      float* g10 = gv + ((size_t)t.r1 * P.Wf + t.c0) * P.CF;
      float* g11 = gv + ((size_t)t.r1 * P.Wf + t.c1) * P.CF;
      sel_atomic_add8(g00 + c8 * 8, t.w00, df);
      sel_atomic_add8(g01 + c8 * 8, t.w01, df);
      sel_atomic_add8(g10 + c8 * 8, t.w10, df);
      sel_atomic_add8(g11 + c8 * 8, t.w11, df);
      if (c8 < 2) {
        const float d = fminf(fmaxf(depth, P.depth_min), P.depth_max);
        const float bi = logf(d / P.depth_min) * P.inv_log_range * score_scale;
        const float bf = floorf(bi);
        const int b0 = min(max((int)bf, 0), P.S - 1), b1 = min(max((int)bf + 1, 0), P.S - 1);
        const float wb1 = bi - bf;
        const int ch = P.D + (c8 ? b1 : b0);
        const float gs = (c8 ? wb1 : 1.0f - wb1) * ds;
        atomicAdd(g00 + ch, t.w00 * gs);
        atomicAdd(g01 + ch, t.w01 * gs);
        atomicAdd(g10 + ch, t.w10 * gs);
        atomicAdd(g11 + ch, t.w11 * gs);
      }
    }
  }
}

}  // namespace snapb200

extern "C" int snapb200_lift_select_pool_backward(const SnapLiftParams* q, int top_k, const SnapLiftView* views,
                                                  const float* view_centers, const void* fimg, const float* xs,
                                                  const float* ys, const float* zs, const void* dstats, float* gimg,
                                                  void* stream) {
  SNAP_REQUIRE(q && views && view_centers && fimg && xs && ys && zs && dstats && gimg, "null pointer");
  SNAP_REQUIRE(top_k >= 1 && top_k <= LIFT_MAX_VIEWS, "1 <= top_k <= %d required (got %d)", LIFT_MAX_VIEWS, top_k);
  SNAP_REQUIRE(q->V > top_k && q->V <= SELECT_MAX_VIEWS, "view selection needs top_k < V <= %d", SELECT_MAX_VIEWS);
  SNAP_REQUIRE(q->D == 128 && !q->no_variance && !q->add_minmax, "default statistics with feature_dim 128 only");
  SNAP_REQUIRE(q->S >= 2 && q->CF == q->D + q->S && q->CF % 8 == 0, "bad channel split");
  SNAP_REQUIRE(q->stats_ld % 8 == 0 && q->stats_ld >= 2 * q->D + 8, "stats_ld too small");
  LiftParams P;
  memcpy(&P, q, sizeof(P));
  lift_select_pool_bwd_kernel<<<(unsigned)(q->X * q->Y), 256, 0, (cudaStream_t)stream>>>(
      P, top_k, reinterpret_cast<const LiftView*>(views), view_centers, (const __nv_bfloat16*)fimg, xs, ys, zs,
      (const __nv_bfloat16*)dstats, gimg);
  return check_launch("lift_select_pool_bwd_kernel");
}
