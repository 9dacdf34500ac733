// The sampling localizer the reference trains with (snap/models/bev_localizer.py:137-218 and
// snap/models/pose_estimation.py:50-205), downstream of the two BEV planes:
//
//   sim[n,i,j]  = relu(bf16(f_q[n] . f_m[i,j]))            -> snapb200_gemm_bf16 (relu epilogue, bf16 out, K = 32)
//   prob[n,:,:] = softmax(exp(T) * sim[n])                  -> loc_softmax_stats_kernel (row max + per-map-row sums;
//                                                              the N x H x W probability tensor is never materialised)
//   point weights (1/num_valid or masked softmax of conf)   -> loc_point_weights_kernel (+ inclusive CDF over points)
//   correspondences ~ prob (jax.random.choice, :139-146)    -> loc_sample_kernel (nested inverse CDF: point, map row, cell)
//   minimal sets -> retries -> Kabsch (:103-165)            -> loc_ransac_poses_kernel (closed-form 2-D Procrustes)
//   pose_scoring_many (:65-85, 206)                         -> loc_pose_scoring_kernel (one similarity map resident in
//                                                              smem per step, 8/16 poses per thread in registers)
//   grid_refinement (:168-203)                              -> loc_refine_poses_kernel + scoring + argmax
//   loss / metrics (bev_localizer.py:244-278)               -> loc_nll_kernel
//
// The similarity tensor sim is kept in HBM as bf16 [B,N,H,W]: the reference rounds the einsum to the feature dtype
// (:157) before the fp32 soft-max, so bf16 storage is lossless; exp(T) and the per-point weight are applied on the fly.
#include <cuda_bf16.h>
#include <math.h>
#include <stdlib.h>

#include "common.cuh"
#include "host_common.h"

namespace snapb200 {

namespace {
constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(FULL, v, o));
  return v;
}
// inclusive scan over the lanes of a warp (lane order)
__device__ __forceinline__ float warp_scan_incl(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float t = __shfl_up_sync(FULL, v, o);
    if (lane >= o) v += t;
  }
  return v;
}
__device__ __forceinline__ float bf16_bits_to_float(unsigned short u) { return __uint_as_float((unsigned)u << 16); }
}  // namespace

// ---------------------------------------------------------------------------------------------------------------------
// soft-max statistics of one similarity map per block: row_max[b,n] = max_ij scale*sim, chunk_sum[b,n,i] =
// sum_j exp(scale*sim[n,i,j] - max), row_sum[b,n] = sum_i chunk_sum (jax.nn.softmax over (-1,-2), bev_localizer.py:163)
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
loc_softmax_stats_kernel(const __nv_bfloat16* __restrict__ sim, int H, int W, float scale,
                         float* __restrict__ row_max, float* __restrict__ row_sum, float* __restrict__ chunk_sum) {
  __shared__ float red[8];
  __shared__ float s_max;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = blockIdx.x;
  const int HW = H * W;
  const __nv_bfloat16* p = sim + (size_t)row * HW;
  // pass 1: maximum (16-byte vectors; HW % 8 == 0 is checked by the host)
  float mx = -INFINITY;
  const uint4* p4 = reinterpret_cast<const uint4*>(p);
  for (int v = threadIdx.x; v < HW / 8; v += 256) {
    const uint4 u = __ldg(p4 + v);
    mx = fmaxf(mx, fmaxf(fmaxf(bf16_lo(u.x), bf16_hi(u.x)), fmaxf(bf16_lo(u.y), bf16_hi(u.y))));
    mx = fmaxf(mx, fmaxf(fmaxf(bf16_lo(u.z), bf16_hi(u.z)), fmaxf(bf16_lo(u.w), bf16_hi(u.w))));
  }
  mx = warp_max(mx);
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    float m = red[0];
    for (int w = 1; w < 8; ++w) m = fmaxf(m, red[w]);
    s_max = __fmul_rn(m, scale);  // scale > 0: max(scale * s) = scale * max(s)
  }
  __syncthreads();
  const float m = s_max;
  // pass 2 (L1/L2 hits): one warp per map row i
  float tot = 0.f;
  for (int i = warp; i < H; i += 8) {
    const __nv_bfloat16* r = p + (size_t)i * W;
    float acc = 0.f;
    for (int j = lane; j < W; j += 32) acc += expf(__fmul_rn(__bfloat162float(r[j]), scale) - m);
    acc = warp_sum(acc);
    if (lane == 0) chunk_sum[(size_t)row * H + i] = acc;
    tot += acc;
  }
  __syncthreads();
  if (lane == 0) red[warp] = tot;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w];
    row_max[row] = m;
    if (row_sum != nullptr) row_sum[row] = t;
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// per-point weights (bev_localizer.py:165-172): without confidences every point weighs 1 / max(num_valid, 1); with
// confidences w = masked_softmax(conf, valid) (layers.py:36-42: an all-invalid mask becomes all-valid).
//   row_cdf[b,n]     = inclusive prefix of the mass of point n in prob_points (NOT masked by validity, as in :170-172)
//   point_scale[b,n] = valid ? exp(T) * w_n : 0   (the factor of sim_points in pose_scoring; :78-84 multiply by valid)
// one block of 1024 threads per example
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
loc_point_weights_kernel(const uint8_t* __restrict__ valid, const float* __restrict__ conf, int N, float exp_t,
                         float* __restrict__ point_scale, float* __restrict__ row_cdf) {
  __shared__ float sh[32];
  __shared__ float s_bcast[2];
  __shared__ int s_cnt;
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint8_t* v = valid + (size_t)b * N;
  const float* c = conf ? conf + (size_t)b * N : nullptr;
  float* ps = point_scale + (size_t)b * N;
  float* cdf = row_cdf + (size_t)b * N;
  // number of valid points
  int cnt = 0;
  for (int n = tid; n < N; n += 1024) cnt += v[n] != 0;
  cnt = __reduce_add_sync(FULL, cnt);
  if (tid == 0) s_cnt = 0;
  __syncthreads();
  if (lane == 0 && cnt) atomicAdd(&s_cnt, cnt);
  __syncthreads();
  const int nv = s_cnt;
  const bool any = nv > 0;
  float mx = 0.f, den = 1.f;
  if (c != nullptr) {
    float m = -INFINITY;
    for (int n = tid; n < N; n += 1024)
      if (!any || v[n]) m = fmaxf(m, c[n]);
    m = warp_max(m);
    if (lane == 0) sh[warp] = m;
    __syncthreads();
    if (tid == 0) {
      float t = sh[0];
      for (int w = 1; w < 32; ++w) t = fmaxf(t, sh[w]);
      s_bcast[0] = t;
    }
    __syncthreads();
    mx = s_bcast[0];
    float d = 0.f;
    for (int n = tid; n < N; n += 1024)
      if (!any || v[n]) d += expf(c[n] - mx);
    d = warp_sum(d);
    __syncthreads();
    if (lane == 0) sh[warp] = d;
    __syncthreads();
    if (tid == 0) {
      float t = 0.f;
      for (int w = 0; w < 32; ++w) t += sh[w];
      s_bcast[1] = t;
    }
    __syncthreads();
    den = s_bcast[1];
  }
  const float inv_nv = __fdiv_rn(1.f, (float)max(nv, 1));
  // contiguous chunk per thread -> block scan of the chunk totals -> inclusive prefix
  const int per = (N + 1023) / 1024;
  const int n0 = min(tid * per, N), n1 = min(n0 + per, N);
  float local = 0.f;
  for (int n = n0; n < n1; ++n) {
    const float w = c ? ((!any || v[n]) ? __fdiv_rn(expf(c[n] - mx), den) : 0.f) : inv_nv;
    local += w;
  }
  float incl = warp_scan_incl(local, lane);
  __syncthreads();
  if (lane == 31) sh[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    float t = sh[lane];
    t = warp_scan_incl(t, lane);
    sh[lane] = t;
  }
  __syncthreads();
  float run = incl - local + (warp > 0 ? sh[warp - 1] : 0.f);
  for (int n = n0; n < n1; ++n) {
    const float w = c ? ((!any || v[n]) ? __fdiv_rn(expf(c[n] - mx), den) : 0.f) : inv_nv;
    run += w;
    cdf[n] = run;
    ps[n] = v[n] ? __fmul_rn(exp_t, w) : 0.f;
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// correspondence sampling (pose_estimation.py:139-146): jax.random.choice(p=prob.reshape(-1), replace=True) draws
// r = total * (1 - u) and returns the first flat index whose cumulative sum reaches r.  The flat order is (point, map
// row, map column), so the draw factorises exactly into three nested inverse CDFs; each sample gets two uniforms
// (point; cell inside the point's map) instead of one so that fp32 resolution is not the limit.
// One warp per sample.
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
loc_sample_kernel(const __nv_bfloat16* __restrict__ sim, const float* __restrict__ row_max,
                  const float* __restrict__ chunk_sum, const float* __restrict__ row_cdf,
                  const float* __restrict__ uniforms, int N, int H, int W, int K, long long total, float scale,
                  int* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const long long s = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (s >= total) return;
  const int b = (int)(s / K);
  const float u1 = uniforms[2 * s], u2 = uniforms[2 * s + 1];
  // ---- point ----------------------------------------------------------------------------------------------------
  const float* cdf = row_cdf + (size_t)b * N;
  const float r1 = __fmul_rn(cdf[N - 1], 1.f - u1);
  int lo = 0, hi = N - 1;  // first n with cdf[n] >= r1
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (cdf[mid] >= r1) hi = mid; else lo = mid + 1;
  }
  const int n = lo;
  const size_t row = (size_t)b * N + n;
  // ---- map row ---------------------------------------------------------------------------------------------------
  const float* cs = chunk_sum + row * H;
  float carry = 0.f;
  for (int base = 0; base < H; base += 32) {
    const float v = base + lane < H ? cs[base + lane] : 0.f;
    carry += __shfl_sync(FULL, warp_scan_incl(v, lane), 31);
  }
  const float r2 = __fmul_rn(carry, 1.f - u2);
  int i = H - 1;
  float before = 0.f;  // mass in front of map row i
  carry = 0.f;
  bool found = false;
  for (int base = 0; base < H && !found; base += 32) {
    const float v = base + lane < H ? cs[base + lane] : 0.f;
    const float inc = carry + warp_scan_incl(v, lane);
    const unsigned hit = __ballot_sync(FULL, base + lane < H && inc >= r2);
    if (hit) {
      const int l = __ffs(hit) - 1;
      i = base + l;
      before = __shfl_sync(FULL, inc - v, l);
      found = true;
    } else {
      carry = __shfl_sync(FULL, inc, 31);
    }
  }
  if (!found) before = carry - cs[H - 1];
  // ---- cell inside the map row -------------------------------------------------------------------------------------
  const float r3 = r2 - before;
  const float m = row_max[row];
  const __nv_bfloat16* r = sim + (row * H + i) * (size_t)W;
  int j = W - 1;
  carry = 0.f;
  found = false;
  for (int base = 0; base < W && !found; base += 32) {
    const float v = base + lane < W ? expf(__fmul_rn(__bfloat162float(r[base + lane]), scale) - m) : 0.f;
    const float inc = carry + warp_scan_incl(v, lane);
    const unsigned hit = __ballot_sync(FULL, base + lane < W && inc >= r3);
    if (hit) {
      j = base + __ffs(hit) - 1;
      found = true;
    } else {
      carry = __shfl_sync(FULL, inc, 31);
    }
  }
  if (lane == 0) {
    out[3 * s + 0] = n;
    out[3 * s + 1] = i;
    out[3 * s + 2] = j;
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// minimal sets -> most consistent retry -> rigid 2-D transform (pose_estimation.py:147-165, 103-123).
// One thread per pose.  Kabsch on two correspondences has the closed form
//   angle = atan2(c10 - c01, c00 + c11),  cov = sum_k (j_k - mu_j)(i_k - mu_i)^T,  t = mu_j - R mu_i
// (the SVD + determinant correction of :111-116 selects exactly this proper rotation; cov == 0 gives R = I).
// ---------------------------------------------------------------------------------------------------------------------
__global__ void loc_ransac_poses_kernel(const int* __restrict__ idx, const float* __restrict__ i_xy, int i_xy_batched,
                                        int N, int P, int retries, float cell, long long total,
                                        float* __restrict__ poses) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int b = (int)(t / P);
  const float* pts = i_xy + (i_xy_batched ? (size_t)b * N * 2 : 0);
  const int* q = idx + (size_t)t * retries * 6;
  float best = INFINITY;
  float ix0 = 0, iy0 = 0, ix1 = 0, iy1 = 0, jx0 = 0, jy0 = 0, jx1 = 0, jy1 = 0;
  for (int r = 0; r < retries; ++r) {
    const int* e = q + r * 6;
    const float ax = pts[2 * e[0]], ay = pts[2 * e[0] + 1], bx = pts[2 * e[3]], by = pts[2 * e[3] + 1];
    // grid.index_to_xyz (grids.py:62-63): (idx + 0.5) * cell
    const float cx = __fmul_rn((float)e[1] + 0.5f, cell), cy = __fmul_rn((float)e[2] + 0.5f, cell);
    const float dx = __fmul_rn((float)e[4] + 0.5f, cell), dy = __fmul_rn((float)e[5] + 0.5f, cell);
    float ratio = 0.f;
    if (retries > 1) {
      const float ex = bx - ax, ey = by - ay, fx = dx - cx, fy = dy - cy;
      const float d_i = sqrtf(__fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey)));
      const float d_j = sqrtf(__fadd_rn(__fmul_rn(fx, fx), __fmul_rn(fy, fy)));
      ratio = fmaxf(__fdiv_rn(d_i, fmaxf(d_j, 1e-5f)), __fdiv_rn(d_j, fmaxf(d_i, 1e-5f)));
    }
    if (r == 0 || ratio < best) {  // argmin keeps the first minimum
      best = ratio;
      ix0 = ax; iy0 = ay; ix1 = bx; iy1 = by;
      jx0 = cx; jy0 = cy; jx1 = dx; jy1 = dy;
    }
  }
  // kabsch_algorithm_2d(i_p = j_pool, j_p = i_pool): j ~ R i + t
  const float mjx = __fmul_rn(__fadd_rn(jx0, jx1), 0.5f), mjy = __fmul_rn(__fadd_rn(jy0, jy1), 0.5f);
  const float mix = __fmul_rn(__fadd_rn(ix0, ix1), 0.5f), miy = __fmul_rn(__fadd_rn(iy0, iy1), 0.5f);
  const float a0x = jx0 - mjx, a0y = jy0 - mjy, a1x = jx1 - mjx, a1y = jy1 - mjy;
  const float b0x = ix0 - mix, b0y = iy0 - miy, b1x = ix1 - mix, b1y = iy1 - miy;
  const float c00 = a0x * b0x + a1x * b1x, c01 = a0x * b0y + a1x * b1y;
  const float c10 = a0y * b0x + a1y * b1x, c11 = a0y * b0y + a1y * b1y;
  const float sn = c10 - c01, cs = c00 + c11;
  const float h = sqrtf(sn * sn + cs * cs);
  float c = 1.f, s = 0.f;
  if (h > 0.f) {
    c = cs / h;
    s = sn / h;
  }
  float* o = poses + (size_t)t * 3;
  o[0] = atan2f(s, c);
  o[1] = mjx - (c * mix - s * miy);
  o[2] = mjy - (s * mix + c * miy);
}

// ---------------------------------------------------------------------------------------------------------------------
// grid refinement poses (pose_estimation.py:177-190): j_t_i_init @ Transform2D(deg2rad(dr), (dx, dy)) for every
// (dr, dx, dy) of the offset axes; pose index = (ir * nx + ix) * ny + iy (jnp.mgrid order).
// ---------------------------------------------------------------------------------------------------------------------
__global__ void loc_refine_poses_kernel(const float* __restrict__ init, const float* __restrict__ rot_rad, int nr,
                                        const float* __restrict__ off_x, int nx, const float* __restrict__ off_y,
                                        int ny, long long total, float* __restrict__ poses) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int per = nr * nx * ny;
  const int b = (int)(t / per);
  int k = (int)(t - (long long)b * per);
  const int iy = k % ny;
  k /= ny;
  const int ix = k % nx, ir = k / nx;
  const float a0 = init[3 * b], tx = init[3 * b + 1], ty = init[3 * b + 2];
  float s, c;
  sincosf(a0, &s, &c);
  const float ox = off_x[ix], oy = off_y[iy];
  // compose (geometry.py:142-145): angle = a0 + da, t = t0 + R(a0) t_off
  float* o = poses + (size_t)t * 3;
  o[0] = __fadd_rn(a0, rot_rad[ir]);
  o[1] = __fadd_rn(tx, __fadd_rn(__fmul_rn(c, ox), __fmul_rn(-s, oy)));
  o[2] = __fadd_rn(ty, __fadd_rn(__fmul_rn(s, ox), __fmul_rn(c, oy)));
}

// ---------------------------------------------------------------------------------------------------------------------
// pose_scoring_many (pose_estimation.py:65-85): score[p] = sum_n valid_n * interp(sim_points[n], (T_p i_xy[n]) / cell)
//
// grid (pose chunks, point splits S, B); block = 256 threads (two blocks per SM: one computes while the other waits
// for its next map), PPT poses per thread held in registers as (cos, sin, t) / cell_size.
// The block walks the valid points of its split: the similarity map of one point (H*W bf16, 32 KB at G = 128) is
// streamed into shared memory with cp.async (double-buffered when two maps fit) and every pose of the chunk gathers
// its four bilinear taps from there (grids.interpolate_nd: taps of (uv - 0.5), clamped indices, corner order
// (0,0),(0,1),(1,0),(1,1); the validity of mask_score_out_of_bounds needs the point inside the map and all four
// taps valid).  Partial sums [S] are added in split order by loc_score_reduce_kernel (deterministic).
// ---------------------------------------------------------------------------------------------------------------------
struct LocScoreArgs {
  const __nv_bfloat16* sim;   // [B,N,H,W]
  const float* point_scale;   // [B,N]
  const float* i_xy;          // [N,2] or [B,N,2]
  const uint8_t* valid_j;     // [B,H,W] (only read when mask_oob)
  const float* poses;         // [B,P,3]
  float* partial;             // [B,S,P]
  int N, H, W, P, S, nbuf, i_xy_batched, mask_oob;
  float cell;
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

constexpr int LOC_SCORE_THREADS = 256;

// Shared-memory layout of one similarity map: rows of RS = W + 8 bf16 (16-byte aligned rows for cp.async, 8 zero
// columns on the right) and one zero row below.  Coordinates are clamped to [0, H-1] x [0, W-1] BEFORE the floor:
// a point left of / above the map then has weights (1, 0) on taps (0, 1) and a point right of / below it weights
// (1, 0) on taps (last, zero padding) -- the same value as the reference's index-clamped taps (whose two weights sum
// to 1 on the same edge texel), with no per-tap clamps in the inner loop.
template <int PPT, bool MASK>
__global__ void __launch_bounds__(LOC_SCORE_THREADS, PPT <= 8 ? 3 : 2)
loc_pose_scoring_kernel(const LocScoreArgs A) {
  constexpr int NT = LOC_SCORE_THREADS;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int HW = A.H * A.W;
  const int RS = A.W + 8;                 // padded row stride (elements)
  const int MAPSZ = (A.H + 1) * RS;       // elements per padded map
  const int b = blockIdx.z, split = blockIdx.y;  // x = pose chunk: co-scheduled blocks share their maps in L2
  const int tid = threadIdx.x;
  // shared memory: nbuf padded maps, then the point list of this split (index, x, y, scale), then the map validity
  unsigned short* maps = reinterpret_cast<unsigned short*>(smem_raw);
  const int per = (A.N + A.S - 1) / A.S;
  const int n0 = min(split * per, A.N), n1 = min(n0 + per, A.N);
  float4* plist = reinterpret_cast<float4*>(smem_raw + (size_t)A.nbuf * MAPSZ * 2);
  uint8_t* vj = reinterpret_cast<uint8_t*>(plist + per);
  __shared__ int s_count;
  if (tid == 0) s_count = 0;
  // zero padding (never touched by the cp.async copies): 8 columns right of every row, one row below
  for (int buf = 0; buf < A.nbuf; ++buf) {
    unsigned short* m = maps + (size_t)buf * MAPSZ;
    for (int k = tid; k < A.H * 8; k += NT) m[(k >> 3) * RS + A.W + (k & 7)] = 0;
    for (int k = tid; k < RS; k += NT) m[A.H * RS + k] = 0;
  }
  __syncthreads();
  // compaction of the valid points (order inside the list does not matter for the result up to fp32 summation
  // order; keep it deterministic: warp 0 walks the range in order)
  if (tid < 32) {
    const float* ps = A.point_scale + (size_t)b * A.N;
    const float* xy = A.i_xy + (A.i_xy_batched ? (size_t)b * A.N * 2 : 0);
    int cnt = 0;
    for (int base = n0; base < n1; base += 32) {
      const int n = base + tid;
      const float sc = n < n1 ? ps[n] : 0.f;
      const unsigned m = __ballot_sync(FULL, sc != 0.f);
      if (sc != 0.f) {
        const int pos = cnt + __popc(m & ((1u << tid) - 1u));
        plist[pos] = make_float4(__int_as_float(n), xy[2 * n], xy[2 * n + 1], sc);
      }
      cnt += __popc(m);
    }
    if (tid == 0) s_count = cnt;
  }
  if (MASK) {
    const uint8_t* g = A.valid_j + (size_t)b * HW;
    for (int k = tid; k < HW; k += NT) vj[k] = g[k];
  }
  // poses of this thread, pre-divided by the cell size: uv - 0.5 = (c x - s y + tx) / cell - 0.5
  float pc[PPT], psn[PPT], ptx[PPT], pty[PPT], acc[PPT];
  const int p0 = blockIdx.x * (NT * PPT);
  const float inv_cell = __fdiv_rn(1.f, A.cell);
#pragma unroll
  for (int k = 0; k < PPT; ++k) {
    const int p = p0 + k * NT + tid;
    acc[k] = 0.f;
    if (p < A.P) {
      const float* q = A.poses + ((size_t)b * A.P + p) * 3;
      float sn, cs;
      sincosf(q[0], &sn, &cs);
      pc[k] = cs * inv_cell;
      psn[k] = sn * inv_cell;
      ptx[k] = q[1] * inv_cell - 0.5f;
      pty[k] = q[2] * inv_cell - 0.5f;
    } else {
      pc[k] = inv_cell; psn[k] = 0.f; ptx[k] = 0.f; pty[k] = 0.f;
    }
  }
  __syncthreads();
  const int count = s_count;
  const __nv_bfloat16* simb = A.sim + (size_t)b * A.N * HW;
  const int vpr = A.W >> 3;        // 16-byte vectors per map row
  const int vecs = A.H * vpr;      // ... per map
  // vector v = tid + i * NT of a map goes to padded element offset d0 + i * dstep (+ RS - W when the column wraps)
  const int row0 = tid / vpr, cvec0 = tid - row0 * vpr;
  const int drow = NT / vpr, dcvec = NT - drow * vpr;
  auto prefetch = [&](int k, int buf) {
    const int n = __float_as_int(plist[k].x);
    const uint4* src = reinterpret_cast<const uint4*>(simb + (size_t)n * HW) + tid;
    unsigned short* dst = maps + buf * MAPSZ;
    int off = row0 * RS + cvec0 * 8, cvec = cvec0;
    for (int v = tid; v < vecs; v += NT) {
      cp_async16(dst + off, src);
      src += NT;
      off += drow * RS + dcvec * 8;
      cvec += dcvec;
      if (cvec >= vpr) {
        cvec -= vpr;
        off += RS - A.W;   // next row: +RS, back to column cvec - vpr: -W
      }
    }
    cp_async_commit();
  };
  const float uH = (float)(A.H - 1), uW = (float)(A.W - 1);
  const unsigned magic_off = 0x4B000000u * (unsigned)RS + 0x4B000000u;  // bits of 2^23, folded out of the index (mod 2^32)
  if (count > 0) prefetch(0, 0);
  const int other = A.nbuf - 1;  // double buffering: the other buffer is cur ^ 1; single buffer: always 0
  int cur = 0;
  for (int k = 0; k < count; ++k) {
    if (other && k + 1 < count) {
      prefetch(k + 1, cur ^ 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const unsigned short* mp = maps + cur * MAPSZ;
    const float4 pt = plist[k];
#pragma unroll
    for (int q = 0; q < PPT; ++q) {
      // Transform2D.transform (geometry.py:138-140), / cell_size (:75), - 0.5 (grids.py:129): two FMAs per axis
      const float cu0 = fmaf(pc[q], pt.y, fmaf(-psn[q], pt.z, ptx[q]));
      const float cv0 = fmaf(psn[q], pt.y, fmaf(pc[q], pt.z, pty[q]));
      const float cu = fminf(fmaxf(cu0, 0.f), uH), cv = fminf(fmaxf(cv0, 0.f), uW);
      // floor + float->int without the conversion unit (FRND / F2I are quarter rate): for 0 <= x < 2^22,
      // x + 2^23 rounded toward zero is 2^23 + floor(x), whose mantissa bits are the integer
      const float mu = __fadd_rz(cu, 8388608.f), mv = __fadd_rz(cv, 8388608.f);
      const float fu = mu - 8388608.f, fv = mv - 8388608.f;
      const float whu = cu - fu, whv = cv - fv;
      const float wlu = 1.f - whu, wlv = 1.f - whv;
      const unsigned off = (unsigned)__float_as_int(mu) * (unsigned)RS + (unsigned)__float_as_int(mv) - magic_off;
      const unsigned short* a = mp + off;
      const float s00 = bf16_bits_to_float(a[0]), s01 = bf16_bits_to_float(a[1]);
      const float s10 = bf16_bits_to_float(a[RS]), s11 = bf16_bits_to_float(a[RS + 1]);
      // corner order (0,0),(0,1),(1,0),(1,1) of map_coordinates
      float val = (wlu * wlv) * s00;
      val = fmaf(wlu * whv, s01, val);
      val = fmaf(whu * wlv, s10, val);
      val = fmaf(whu * whv, s11, val);
      if (MASK) {
        // point inside the map (0 <= uv < size <=> -0.5 <= uv - 0.5 < size - 0.5) and the reference's four
        // index-clamped taps valid (grids.py:131-136)
        const float gu = floorf(cu0), gv = floorf(cv0);
        const int iu = (int)fminf(fmaxf(gu, -2.f), uH + 1.f), iv = (int)fminf(fmaxf(gv, -2.f), uW + 1.f);
        const int r0 = min(max(iu, 0), A.H - 1) * A.W, r1 = min(max(iu + 1, 0), A.H - 1) * A.W;
        const int c0 = min(max(iv, 0), A.W - 1), c1 = min(max(iv + 1, 0), A.W - 1);
        const bool ok = cu0 >= -0.5f && cu0 < uH + 0.5f && cv0 >= -0.5f && cv0 < uW + 0.5f &&
                        (vj[r0 + c0] & vj[r0 + c1] & vj[r1 + c0] & vj[r1 + c1]) != 0;
        if (ok) acc[q] = fmaf(val, pt.w, acc[q]);
      } else {
        acc[q] = fmaf(val, pt.w, acc[q]);
      }
    }
    __syncthreads();  // everyone is done with buffer `cur` before it is refilled
    if (!other) {
      if (k + 1 < count) prefetch(k + 1, 0);
    } else {
      cur ^= 1;
    }
  }
#pragma unroll
  for (int k = 0; k < PPT; ++k) {
    const int p = p0 + k * NT + tid;
    if (p < A.P) A.partial[((size_t)b * A.S + split) * A.P + p] = acc[k];
  }
}

__global__ void loc_score_reduce_kernel(const float* __restrict__ partial, int S, int P, long long total,
                                        float* __restrict__ scores) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const long long b = t / P;
  const int p = (int)(t - b * P);
  float a = 0.f;
  for (int s = 0; s < S; ++s) a += partial[(b * S + s) * P + p];
  scores[t] = a;
}

// ---------------------------------------------------------------------------------------------------------------------
// argmax over x[row, start:cols] (jnp.argmax: first maximum), optionally gathering rows of 3 floats
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
argmax_rows_kernel(const float* __restrict__ x, int cols, int start, int* __restrict__ idx,
                   const float* __restrict__ poses, float* __restrict__ best_pose) {
  __shared__ float sv[8];
  __shared__ int si[8];
  const int row = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float* r = x + (size_t)row * cols;
  float bv = -INFINITY;
  int bi = 0x7fffffff;
  for (int k = start + threadIdx.x; k < cols; k += 256) {
    const float v = r[k];
    if (v > bv || (bi == 0x7fffffff)) {
      bv = v;
      bi = k;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(FULL, bv, o);
    const int oi = __shfl_xor_sync(FULL, bi, o);
    if (oi != 0x7fffffff && (bi == 0x7fffffff || ov > bv || (ov == bv && oi < bi))) {
      bv = ov;
      bi = oi;
    }
  }
  if (lane == 0) {
    sv[warp] = bv;
    si[warp] = bi;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w)
      if (si[w] != 0x7fffffff && (bi == 0x7fffffff || sv[w] > bv || (sv[w] == bv && si[w] < bi))) {
        bv = sv[w];
        bi = si[w];
      }
    const int rel = bi - start;
    idx[row] = rel;
    if (poses != nullptr && best_pose != nullptr) {
      const float* q = poses + ((size_t)row * cols + bi) * 3;
      best_pose[3 * row] = q[0];
      best_pose[3 * row + 1] = q[1];
      best_pose[3 * row + 2] = q[2];
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// loss + metrics of BEVLocalizerModel.loss_metrics_function (bev_localizer.py:244-278).  One block per example.
//   samples_t_gt = samples.inv @ gt; (dr, dt) = magnitude (geometry.py:126-136)
//   remove = dr < dr_min & dt < dt_min (never index 0); nll = logsumexp(masked scores) - scores[0]
//   out[b] = {nll, dr(best), dt(best), argmax(scores) == 0, recall_samples @ (0.5 m, 1 deg), (1, 2), (2, 4)}
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void pose_error(float a, float tx, float ty, float ga, float gx, float gy, float* dr,
                                           float* dt) {
  // inv: angle -a, t_inv = -(R^T t); compose with gt: angle = -a + ga, t = t_inv + R(-a) t_gt
  float s, c;
  sincosf(a, &s, &c);
  const float ix = -__fadd_rn(__fmul_rn(c, tx), __fmul_rn(s, ty));
  const float iy = -__fadd_rn(__fmul_rn(-s, tx), __fmul_rn(c, ty));
  float s2, c2;
  sincosf(-a, &s2, &c2);
  const float ox = __fadd_rn(ix, __fadd_rn(__fmul_rn(c2, gx), __fmul_rn(-s2, gy)));
  const float oy = __fadd_rn(iy, __fadd_rn(__fmul_rn(s2, gx), __fmul_rn(c2, gy)));
  const float ang = __fadd_rn(-a, ga);
  float d = fmodf(fabsf(ang) * 57.29577951308232f, 360.f);
  *dr = fminf(d, 360.f - d);
  *dt = sqrtf(__fadd_rn(__fmul_rn(ox, ox), __fmul_rn(oy, oy)));
}

__global__ void __launch_bounds__(256)
loc_nll_kernel(const float* __restrict__ scores, const float* __restrict__ samples, const float* __restrict__ best,
               const float* __restrict__ gt, int P1, int use_remove, float dr_min, float dt_min,
               float* __restrict__ out, float* __restrict__ dr_out, float* __restrict__ dt_out) {
  __shared__ float shf[8];
  __shared__ int shi[8][3];
  __shared__ float s_m;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* sc = scores + (size_t)b * P1;
  const float* sp = samples + (size_t)b * P1 * 3;
  const float ga = gt[3 * b], gx = gt[3 * b + 1], gy = gt[3 * b + 2];
  // pass 1: max of the masked scores, sample recalls, raw arg-max
  float mx = -INFINITY, raw_best = -INFINITY;
  int raw_idx = 0x7fffffff;
  int rc0 = 0, rc1 = 0, rc2 = 0;
  for (int k = tid; k < P1; k += 256) {
    float dr, dt;
    pose_error(sp[3 * k], sp[3 * k + 1], sp[3 * k + 2], ga, gx, gy, &dr, &dt);
    if (dr_out) dr_out[(size_t)b * P1 + k] = dr;
    if (dt_out) dt_out[(size_t)b * P1 + k] = dt;
    const float v = sc[k];
    if (v > raw_best || raw_idx == 0x7fffffff) {
      raw_best = v;
      raw_idx = k;
    }
    const bool rem = use_remove && k > 0 && dr < dr_min && dt < dt_min;
    if (!rem) mx = fmaxf(mx, v);
    if (k > 0) {
      rc0 += dr < 1.f && dt < 0.5f;
      rc1 += dr < 2.f && dt < 1.f;
      rc2 += dr < 4.f && dt < 2.f;
    }
  }
  mx = warp_max(mx);
  rc0 = __reduce_add_sync(FULL, rc0);
  rc1 = __reduce_add_sync(FULL, rc1);
  rc2 = __reduce_add_sync(FULL, rc2);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(FULL, raw_best, o);
    const int oi = __shfl_xor_sync(FULL, raw_idx, o);
    if (oi != 0x7fffffff && (raw_idx == 0x7fffffff || ov > raw_best || (ov == raw_best && oi < raw_idx))) {
      raw_best = ov;
      raw_idx = oi;
    }
  }
  __shared__ float sbv[8];
  __shared__ int sbi[8];
  if (lane == 0) {
    shf[warp] = mx;
    shi[warp][0] = rc0;
    shi[warp][1] = rc1;
    shi[warp][2] = rc2;
    sbv[warp] = raw_best;
    sbi[warp] = raw_idx;
  }
  __syncthreads();
  if (tid == 0) {
    float m = shf[0];
    for (int w = 1; w < 8; ++w) m = fmaxf(m, shf[w]);
    s_m = m;
  }
  __syncthreads();
  const float m = s_m;
  float se = 0.f;
  for (int k = tid; k < P1; k += 256) {
    bool rem = false;
    if (use_remove && k > 0) {
      float dr, dt;
      pose_error(sp[3 * k], sp[3 * k + 1], sp[3 * k + 2], ga, gx, gy, &dr, &dt);
      rem = dr < dr_min && dt < dt_min;
    }
    if (!rem) se += expf(sc[k] - m);
  }
  se = warp_sum(se);
  __syncthreads();
  if (lane == 0) shf[warp] = se;
  __syncthreads();
  if (tid == 0) {
    float t = 0.f;
    int a0 = 0, a1 = 0, a2 = 0;
    float bv = sbv[0];
    int bi = sbi[0];
    for (int w = 0; w < 8; ++w) {
      t += shf[w];
      a0 += shi[w][0];
      a1 += shi[w][1];
      a2 += shi[w][2];
      if (w > 0 && sbi[w] != 0x7fffffff && (bi == 0x7fffffff || sbv[w] > bv || (sbv[w] == bv && sbi[w] < bi))) {
        bv = sbv[w];
        bi = sbi[w];
      }
    }
    float* o = out + (size_t)b * 7;
    o[0] = -(sc[0] - m - logf(t));  // -log_softmax(scores)[0]
    float dr, dt;
    pose_error(best[3 * b], best[3 * b + 1], best[3 * b + 2], ga, gx, gy, &dr, &dt);
    o[1] = dr;
    o[2] = dt;
    o[3] = bi == 0 ? 1.f : 0.f;
    const float den = (float)max(P1 - 1, 1);
    o[4] = (float)a0 / den;
    o[5] = (float)a1 / den;
    o[6] = (float)a2 / den;
  }
}

}  // namespace snapb200

using namespace snapb200;

extern "C" {

int snapb200_loc_softmax_stats(const void* sim, long long rows, int H, int W, float scale, float* row_max,
                               float* row_sum, float* chunk_sum, void* stream) {
  SNAP_REQUIRE(sim && row_max && chunk_sum, "null pointer");
  SNAP_REQUIRE(rows > 0 && rows < (1ll << 31) && H >= 1 && W >= 1 && (H * W) % 8 == 0,
               "need rows > 0 and H*W %% 8 == 0 (got %lld, %d x %d)", rows, H, W);
  SNAP_REQUIRE(scale > 0.f, "scale = exp(temperature) must be positive");
  loc_softmax_stats_kernel<<<(unsigned)rows, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)sim, H, W, scale,
                                                                            row_max, row_sum, chunk_sum);
  return check_launch("loc_softmax_stats_kernel");
}

int snapb200_loc_point_weights(const uint8_t* valid_points, const float* conf, int B, int N, float exp_t,
                               float* point_scale, float* row_cdf, void* stream) {
  SNAP_REQUIRE(valid_points && point_scale && row_cdf, "null pointer");
  SNAP_REQUIRE(B >= 1 && N >= 1, "empty batch");
  loc_point_weights_kernel<<<B, 1024, 0, (cudaStream_t)stream>>>(valid_points, conf, N, exp_t, point_scale, row_cdf);
  return check_launch("loc_point_weights_kernel");
}

int snapb200_loc_sample(const void* sim, const float* row_max, const float* chunk_sum, const float* row_cdf,
                        const float* uniforms, int B, int N, int H, int W, int K, float scale, int* indices,
                        void* stream) {
  SNAP_REQUIRE(sim && row_max && chunk_sum && row_cdf && uniforms && indices, "null pointer");
  SNAP_REQUIRE(B >= 1 && N >= 1 && H >= 1 && W >= 1 && K >= 1, "empty problem");
  const long long total = (long long)B * K;
  loc_sample_kernel<<<(unsigned)((total + 7) / 8), 256, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)sim, row_max, chunk_sum, row_cdf, uniforms, N, H, W, K, total, scale, indices);
  return check_launch("loc_sample_kernel");
}

int snapb200_loc_ransac_poses(const int* indices, const float* i_xy, int i_xy_batched, int B, int N, int num_poses,
                              int num_retries, float cell_size, float* poses, void* stream) {
  SNAP_REQUIRE(indices && i_xy && poses, "null pointer");
  SNAP_REQUIRE(B >= 1 && num_poses >= 1 && num_retries >= 1, "empty problem");
  const long long total = (long long)B * num_poses;
  loc_ransac_poses_kernel<<<(unsigned)((total + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
      indices, i_xy, i_xy_batched, N, num_poses, num_retries, cell_size, total, poses);
  return check_launch("loc_ransac_poses_kernel");
}

int snapb200_loc_refine_poses(const float* init, int B, const float* rot_rad, int nr, const float* off_x, int nx,
                              const float* off_y, int ny, float* poses, void* stream) {
  SNAP_REQUIRE(init && rot_rad && off_x && off_y && poses, "null pointer");
  SNAP_REQUIRE(B >= 1 && nr >= 1 && nx >= 1 && ny >= 1, "empty problem");
  const long long total = (long long)B * nr * nx * ny;
  loc_refine_poses_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      init, rot_rad, nr, off_x, nx, off_y, ny, total, poses);
  return check_launch("loc_refine_poses_kernel");
}

static int loc_score_plan(const SnapLocScoreParams* p, int* ppt, int* chunks, int* splits, int* nbuf, size_t* smem) {
  SNAP_REQUIRE(p != nullptr, "null params");
  SNAP_REQUIRE(p->B >= 1 && p->N >= 1 && p->H >= 1 && p->W >= 1 && p->P >= 1, "empty problem");
  SNAP_REQUIRE(p->W % 8 == 0, "the map width must be a multiple of 8 (got %d)", p->W);
  SNAP_REQUIRE(p->cell_size > 0.f, "cell_size must be positive");
  *ppt = p->P > 2048 ? 16 : 8;
  if (const char* e = getenv("SNAPB200_LOC_PPT")) {  // tuning knob: poses per thread (8: 3 blocks per SM, 16: 2)
    const int v = atoi(e);
    if (v == 8 || v == 16) *ppt = v;
  }
  *chunks = (p->P + LOC_SCORE_THREADS * *ppt - 1) / (LOC_SCORE_THREADS * *ppt);
  int S = (2 * (*ppt <= 8 ? 3 : 2) * num_sms() + *chunks * p->B - 1) / (*chunks * p->B);  // resident blocks x two waves
  S = S < 1 ? 1 : S;
  if (S > (p->N + 15) / 16) S = (p->N + 15) / 16;  // at least ~16 points per split
  *splits = S;
  const int per = (p->N + S - 1) / S;
  const size_t map_bytes = (size_t)(p->H + 1) * (p->W + 8) * 2;  // padded layout of loc_pose_scoring_kernel
  const size_t extra = (size_t)per * 16 + (p->mask_out_of_bounds ? (size_t)p->H * p->W : 0) + 16;
  const size_t cap = 227 * 1024 - 1024;
  // double-buffer when two maps fit; the common 128 x 128 map (32 KB) then also leaves room for two blocks per SM
  *nbuf = 2 * map_bytes + extra <= cap ? 2 : 1;
  *smem = *nbuf * map_bytes + extra;
  SNAP_REQUIRE(*smem <= cap, "similarity map of %d x %d does not fit in shared memory", p->H, p->W);
  return 0;
}

size_t snapb200_loc_pose_scoring_workspace(const SnapLocScoreParams* p) {
  int ppt, chunks, splits, nbuf;
  size_t smem;
  if (loc_score_plan(p, &ppt, &chunks, &splits, &nbuf, &smem) != 0) return 0;
  return (size_t)p->B * splits * p->P * sizeof(float);
}

int snapb200_loc_pose_scoring(const SnapLocScoreParams* p, const void* sim, const float* point_scale,
                              const float* i_xy, const uint8_t* valid_j, const float* poses, void* workspace,
                              size_t workspace_bytes, float* scores, void* stream) {
  int ppt, chunks, splits, nbuf;
  size_t smem;
  if (int rc = loc_score_plan(p, &ppt, &chunks, &splits, &nbuf, &smem)) return rc;
  SNAP_REQUIRE(sim && point_scale && i_xy && poses && workspace && scores, "null pointer");
  SNAP_REQUIRE(!p->mask_out_of_bounds || valid_j, "mask_out_of_bounds needs the map validity");
  SNAP_REQUIRE(workspace_bytes >= (size_t)p->B * splits * p->P * sizeof(float), "workspace too small");
  LocScoreArgs a;
  a.sim = (const __nv_bfloat16*)sim;
  a.point_scale = point_scale;
  a.i_xy = i_xy;
  a.valid_j = valid_j;
  a.poses = poses;
  a.partial = (float*)workspace;
  a.N = p->N; a.H = p->H; a.W = p->W; a.P = p->P; a.S = splits; a.nbuf = nbuf;
  a.i_xy_batched = p->i_xy_batched;
  a.mask_oob = p->mask_out_of_bounds;
  a.cell = p->cell_size;
  cudaStream_t s = (cudaStream_t)stream;
  const dim3 grid(chunks, splits, p->B);
#define SNAP_LOC_SCORE(PPT_, MASK_)                                                                              \
  do {                                                                                                          \
    static DynSmemState smem_state; /* largest size set so far, per device (the attribute is sticky) */         \
    if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(&loc_pose_scoring_kernel<PPT_, MASK_>), smem,    \
                                 &smem_state, "cudaFuncSetAttribute(loc_pose_scoring)"))                        \
      return rc;                                                                                                \
    loc_pose_scoring_kernel<PPT_, MASK_><<<grid, LOC_SCORE_THREADS, smem, s>>>(a);                              \
  } while (0)
  if (ppt == 16) {
    if (p->mask_out_of_bounds) SNAP_LOC_SCORE(16, true); else SNAP_LOC_SCORE(16, false);
  } else {
    if (p->mask_out_of_bounds) SNAP_LOC_SCORE(8, true); else SNAP_LOC_SCORE(8, false);
  }
#undef SNAP_LOC_SCORE
  if (int rc = check_launch("loc_pose_scoring_kernel")) return rc;
  const long long total = (long long)p->B * p->P;
  loc_score_reduce_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(a.partial, splits, p->P, total, scores);
  return check_launch("loc_score_reduce_kernel");
}

int snapb200_argmax_rows(const float* x, int rows, int cols, int start, int* idx, const float* rows3,
                         float* best_row3, void* stream) {
  SNAP_REQUIRE(x && idx, "null pointer");
  SNAP_REQUIRE(rows >= 1 && cols >= 1 && start >= 0 && start < cols, "bad shape");
  argmax_rows_kernel<<<rows, 256, 0, (cudaStream_t)stream>>>(x, cols, start, idx, rows3, best_row3);
  return check_launch("argmax_rows_kernel");
}

int snapb200_loc_nll(const float* scores, const float* samples, const float* best, const float* gt, int B, int P1,
                     int use_remove, float dr_min, float dt_min, float* out, float* dr_samples, float* dt_samples,
                     void* stream) {
  SNAP_REQUIRE(scores && samples && best && gt && out, "null pointer");
  SNAP_REQUIRE(B >= 1 && P1 >= 1, "empty problem");
  loc_nll_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(scores, samples, best, gt, P1, use_remove, dr_min, dt_min, out,
                                                      dr_samples, dt_samples);
  return check_launch("loc_nll_kernel");
}

}  // extern "C"
