// The per-observation `depth_mlp` residual of the un-weighted fusion branch (snap/models/streetview_encoder.py:262-267):
//     log_depth = log10(clip(depth, 0.1, 100));  rays = where(visible, rays, 0)
//     f_proj   += depth_mlp(concat([f_proj, log_depth, rays]))          per (voxel, view) observation
// followed by the plain (un-weighted) mean / variance pooling of pool_multiview_features (:153-155).
//
// Two kernels around the tcgen05 GEMM engine (the MLP itself runs there on N * V rows):
//   lift_observe_kernel   projection (bit-exact, lift_common.cuh) + bilinear gather of the encoder features -> one row per
//                         observation: bf16 [N * V, 160] = [f(128) | log10 depth | ray(3) | 0...], vis u8 [N * V]
//   lift_pool_obs_kernel  f' = bf16(f + d), mean / population variance over the visible views -> statistics rows
//                         bf16 [N, stats_ld] = [mean(128) | var(128) | 0...], valid u8 [N]
// A non-default configuration (defaults.py:213: depth_mlp is a placeholder): written for clarity, not tuned.
#include <cuda_bf16.h>
#include <math.h>

#include "common.cuh"
#include "host_common.h"
#include "lift_common.cuh"

namespace snapb200 {

constexpr int OBS_LD = 160;

__global__ void __launch_bounds__(256)
lift_observe_kernel(const __grid_constant__ LiftParams P, const LiftView* __restrict__ views,
                    const __nv_bfloat16* __restrict__ fimg, const float* __restrict__ xs, const float* __restrict__ ys,
                    const float* __restrict__ zs, __nv_bfloat16* __restrict__ obs, uint8_t* __restrict__ vis) {
  __shared__ LiftView sview[LIFT_MAX_VIEWS];
  for (int i = threadIdx.x; i < P.V * (int)(sizeof(LiftView) / 4); i += blockDim.x)
    reinterpret_cast<uint32_t*>(sview)[i] = reinterpret_cast<const uint32_t*>(views)[i];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int col_id = blockIdx.x;
  const int ix = col_id / P.Y, iy = col_id - ix * P.Y;
  const float px = xs[P.xy_paired ? col_id : ix], py = ys[P.xy_paired ? col_id : iy];
  const int half = lane >> 4, c8 = lane & 15;
  for (int iz = warp; iz < P.Z; iz += 8) {
    const long long n = (long long)col_id * P.Z + iz;
    const float pz = zs[iz];
    for (int v = 0; v < P.V; ++v) {
      const LiftView& lv = sview[v];
      const Proj pr = project_point(lv, px, py, pz);
      __nv_bfloat16* row = obs + ((size_t)n * P.V + v) * OBS_LD;
      if (lane == 0) vis[n * P.V + v] = pr.vis ? 1 : 0;
      if (!pr.vis) {  // warp-uniform; the pooling ignores the row, but the MLP must see finite numbers
        if (lane < OBS_LD / 8) reinterpret_cast<uint4*>(row)[lane] = make_uint4(0u, 0u, 0u, 0u);
        continue;
      }
      const Taps t = make_taps(pr.row, pr.col, P.Hf, P.Wf);
      const __nv_bfloat16* img = fimg + (size_t)v * P.Hf * P.Wf * P.CF;
      const int rr = half ? t.r1 : t.r0;
      const float wr = half ? t.wr1 : __fadd_rn(1.0f, -t.wr1);
      const float wc0 = __fadd_rn(1.0f, -t.wc1);
      const uint4 ua = __ldg(reinterpret_cast<const uint4*>(img + ((size_t)rr * P.Wf + t.c0) * P.CF + c8 * 8));
      const uint4 ub = __ldg(reinterpret_cast<const uint4*>(img + ((size_t)rr * P.Wf + t.c1) * P.CF + c8 * 8));
      float fa[8], fb[8], f[8];
      unpack8(ua, fa);
      unpack8(ub, fb);
      const float wx0 = __fmul_rn(wr, wc0), wx1 = __fmul_rn(wr, t.wc1);
#pragma unroll
      for (int j = 0; j < 8; ++j) {   // same operation order as lift_gather_pool_kernel
        float part = __fmaf_rn(wx1, fb[j], __fmul_rn(wx0, fa[j]));
        f[j] = __fadd_rn(part, __shfl_xor_sync(0xffffffffu, part, 16));
      }
      if (half == 0) {
        reinterpret_cast<uint4*>(row)[c8] = make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]),
                                                       pack_bf16(f[6], f[7]));
      } else if (c8 < (OBS_LD - 128) / 8) {
        uint4 o = make_uint4(0u, 0u, 0u, 0u);
        if (c8 == 0) {
          // p_view = R^T (p - t) in the oracle's operation order (geometry.py:52-56,67-69), rays = p_view / max(|p_view|, 1e-5)
          float pv[3];
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            float s = __fmul_rn(lv.Rinv[3 * i + 0], px);
            s = __fadd_rn(s, __fmul_rn(lv.Rinv[3 * i + 1], py));
            s = __fadd_rn(s, __fmul_rn(lv.Rinv[3 * i + 2], pz));
            pv[i] = __fadd_rn(lv.tinv[i], s);
          }
          const float dist = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(pv[0], pv[0]), __fmul_rn(pv[1], pv[1])), __fmul_rn(pv[2], pv[2])));
          const float den = fmaxf(dist, 1e-5f);
          const float ld = log10f(fminf(fmaxf(pr.depth, 0.1f), 100.f));
          o.x = pack_bf16(ld, __fdiv_rn(pv[0], den));
          o.y = pack_bf16(__fdiv_rn(pv[1], den), __fdiv_rn(pv[2], den));
        }
        reinterpret_cast<uint4*>(row)[16 + c8] = o;
      }
    }
  }
}

// warp per voxel, lane = 4 channels
__global__ void __launch_bounds__(256)
lift_pool_obs_kernel(int V, long long N, const __nv_bfloat16* __restrict__ obs, const __nv_bfloat16* __restrict__ d,
                     const uint8_t* __restrict__ vis, int stats_ld, __nv_bfloat16* __restrict__ stats,
                     uint8_t* __restrict__ valid) {
  const int lane = threadIdx.x & 31;
  const long long n = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (n >= N) return;
  float f[LIFT_MAX_VIEWS][4];
  unsigned mask = 0;
  int cnt = 0;
  for (int v = 0; v < V; ++v) {
    if (!vis[n * V + v]) continue;   // warp-uniform
    mask |= 1u << v;
    ++cnt;
    const uint2 a = *reinterpret_cast<const uint2*>(obs + ((size_t)n * V + v) * OBS_LD + lane * 4);
    const uint2 b = *reinterpret_cast<const uint2*>(d + ((size_t)n * V + v) * 128 + lane * 4);
    const uint32_t s0 = hadd2_bf16_rn(a.x, b.x), s1 = hadd2_bf16_rn(a.y, b.y);   // f_proj + depth_mlp(...) -> dtype
    f[v][0] = bf16_lo(s0); f[v][1] = bf16_hi(s0); f[v][2] = bf16_lo(s1); f[v][3] = bf16_hi(s1);
  }
  float mean[4] = {0.f, 0.f, 0.f, 0.f}, var[4] = {0.f, 0.f, 0.f, 0.f};
  if (cnt > 0) {
    const float c = (float)cnt;
    for (int v = 0; v < V; ++v)
      if (mask & (1u << v))
#pragma unroll
        for (int j = 0; j < 4; ++j) mean[j] += f[v][j];
#pragma unroll
    for (int j = 0; j < 4; ++j) mean[j] = __fdiv_rn(mean[j], c);
    for (int v = 0; v < V; ++v)
      if (mask & (1u << v))
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float dd = f[v][j] - mean[j];
          var[j] += dd * dd;
        }
#pragma unroll
    for (int j = 0; j < 4; ++j) var[j] = __fdiv_rn(var[j], c);
  }
  __nv_bfloat16* row = stats + n * stats_ld;
  *reinterpret_cast<uint2*>(row + lane * 4) = make_uint2(pack_bf16(mean[0], mean[1]), pack_bf16(mean[2], mean[3]));
  *reinterpret_cast<uint2*>(row + 128 + lane * 4) = make_uint2(pack_bf16(var[0], var[1]), pack_bf16(var[2], var[3]));
  for (int c8 = lane; c8 < (stats_ld - 256) / 8; c8 += 32) reinterpret_cast<uint4*>(row + 256)[c8] = make_uint4(0u, 0u, 0u, 0u);
  if (lane == 0) valid[n] = cnt > 0 ? 1 : 0;
}

}  // namespace snapb200

using namespace snapb200;

extern "C" {

int snapb200_lift_observe(const SnapLiftParams* q, const SnapLiftView* views, const void* fimg, const float* xs,
                          const float* ys, const float* zs, void* obs, uint8_t* vis, void* stream) {
  SNAP_REQUIRE(q && views && fimg && xs && ys && zs && obs && vis, "null pointer");
  SNAP_REQUIRE(q->V >= 1 && q->V <= LIFT_MAX_VIEWS && q->D == 128 && q->CF >= 128 && q->CF % 8 == 0, "bad views / channels");
  LiftParams P;
  static_assert(sizeof(LiftParams) == sizeof(SnapLiftParams), "layout");
  memcpy(&P, q, sizeof(P));
  lift_observe_kernel<<<(unsigned)(q->X * q->Y), 256, 0, (cudaStream_t)stream>>>(
      P, reinterpret_cast<const LiftView*>(views), (const __nv_bfloat16*)fimg, xs, ys, zs, (__nv_bfloat16*)obs, vis);
  return check_launch("lift_observe_kernel");
}

int snapb200_lift_pool_observations(int V, long long N, const void* obs, const void* d, const uint8_t* vis, int stats_ld,
                                    void* stats, uint8_t* valid, void* stream) {
  SNAP_REQUIRE(obs && d && vis && stats && valid, "null pointer");
  SNAP_REQUIRE(V >= 1 && V <= LIFT_MAX_VIEWS && stats_ld >= 256 && stats_ld % 8 == 0, "bad arguments");
  lift_pool_obs_kernel<<<(unsigned)((N + 7) / 8), 256, 0, (cudaStream_t)stream>>>(
      V, N, (const __nv_bfloat16*)obs, (const __nv_bfloat16*)d, vis, stats_ld, (__nv_bfloat16*)stats, valid);
  return check_launch("lift_pool_obs_kernel");
}

}  // extern "C"
