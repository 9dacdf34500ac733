// Exhaustive (x, y, theta) pose voting, snap/models/pose_exhaustive_voting.py:37-124.
//   rot_templates : sample_query_templates (:37-69) — R/4 bilinear warps + rot90 index maps, NaN-mask validity
//   pad_map       : jnp.pad(mode='edge') of the map plane (:83-85) into the TMA-friendly padded layout
//   xcorr_count   : overlap count = true convolution of the UN-flipped q_valid with the zero-padded
//                   m_valid (:94-99), computed exactly with bit masks + popcount
//   xcorr_scores  : dense correlation over (i, j, d) per rotation as ONE segmented tcgen05 GEMM
//                   (M = shifts, N = rotations, K = G*G*D) with the -inf mask and the division by
//                   sum(q_valid) fused in the epilogue (:100-103)
#include <cuda_bf16.h>
#include <math.h>

#include "common.cuh"
#include "gemm_launch.h"
#include "gemm_tc.cuh"

namespace snapb200 {

struct RotParams {
  int nq;            // R / 4
  float rot[16][4];  // per quarter rotation: cos, sin, tx, ty of templates_t_grid
};

// (a, b) of template quadrant k  ->  (i, j) of the sampled quarter (jnp.rot90(k, axes=(2,1)))
__device__ __forceinline__ void rot90_src(int k, int G, int a, int b, int& i, int& j) {
  switch (k & 3) {
    case 0: i = a; j = b; break;
    case 1: i = G - 1 - b; j = a; break;
    case 2: i = G - 1 - a; j = G - 1 - b; break;
    default: i = b; j = G - 1 - a; break;
  }
}

__global__ void rot_templates_kernel(const __grid_constant__ RotParams RP_,
                                     const __nv_bfloat16* __restrict__ feats,
                                     const uint8_t* __restrict__ valid, const float* __restrict__ conf,
                                     const float* __restrict__ centers, float cell, int B, int R, int RP,
                                     int G, int D, __nv_bfloat16* __restrict__ templates,
                                     uint8_t* __restrict__ t_valid) {
  const int dv = D / 8;
  const long long total = (long long)B * RP * G * G * dv;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c8 = (int)(idx % dv);
  long long rest = idx / dv;
  const int bb = (int)(rest % G);
  rest /= G;
  const int aa = (int)(rest % G);
  rest /= G;
  const int r = (int)(rest % RP);
  const int b = (int)(rest / RP);
  // templates are stored cell-major for the correlation GEMM: [b][a*G+bb][r (padded to RP)][D]
  const size_t trow = (((size_t)b * G + aa) * G + bb) * RP + r;
  if (r >= R) {  // zero padding rows of the B operand
    *reinterpret_cast<uint4*>(templates + trow * D + c8 * 8) = make_uint4(0, 0, 0, 0);
    return;
  }
  const int k = r / RP_.nq, r0 = r - k * RP_.nq;
  int i, j;
  rot90_src(k, G, aa, bb, i, j);
  // templates_xy = t + R(angle) @ grid_xy (snap/utils/geometry.py:138-140 order), uv = xy / cell
  const float cs = RP_.rot[r0][0], sn = RP_.rot[r0][1], tx = RP_.rot[r0][2], ty = RP_.rot[r0][3];
  const float x = centers[i], y = centers[j];
  const float ux = __fdiv_rn(__fadd_rn(tx, __fadd_rn(__fmul_rn(cs, x), __fmul_rn(-sn, y))), cell);
  const float uy = __fdiv_rn(__fadd_rn(ty, __fadd_rn(__fmul_rn(sn, x), __fmul_rn(cs, y))), cell);
  bool ok = ux >= 0.f && ux < (float)G && uy >= 0.f && uy < (float)G;
  const float px = __fadd_rn(ux, -0.5f), py = __fadd_rn(uy, -0.5f);
  const float fx = floorf(px), fy = floorf(py);
  const float wx1 = __fadd_rn(px, -fx), wy1 = __fadd_rn(py, -fy);
  const int x0 = min(max((int)fx, 0), G - 1), x1 = min(max((int)fx + 1, 0), G - 1);
  const int y0 = min(max((int)fy, 0), G - 1), y1 = min(max((int)fy + 1, 0), G - 1);
  const uint8_t* vb = valid + (size_t)b * G * G;
  // NaN-mask semantics (snap/utils/grids.py:131-136): every clamped tap must be valid, even at weight 0
  ok = ok && vb[x0 * G + y0] && vb[x0 * G + y1] && vb[x1 * G + y0] && vb[x1 * G + y1];
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (ok) {
    const int xs[2] = {x0, x1}, ys[2] = {y0, y1};
    const float wxs[2] = {1.0f - wx1, wx1}, wys[2] = {1.0f - wy1, wy1};
#pragma unroll
    for (int p = 0; p < 2; ++p)
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const size_t cellidx = (size_t)b * G * G + (size_t)xs[p] * G + ys[q];
        float wt = wxs[p] * wys[q];
        if (conf != nullptr) wt *= conf[cellidx];
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(feats + cellidx * D + c8 * 8));
        const uint32_t uu[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 t = unpack_bf16(uu[e]);
          acc[2 * e] += wt * t.x;
          acc[2 * e + 1] += wt * t.y;
        }
      }
  }
  const size_t o = (((size_t)b * R + r) * G + aa) * G + bb;
  *reinterpret_cast<uint4*>(templates + trow * D + c8 * 8) =
      make_uint4(pack_bf16(acc[0], acc[1]), pack_bf16(acc[2], acc[3]), pack_bf16(acc[4], acc[5]),
                 pack_bf16(acc[6], acc[7]));
  if (c8 == 0) t_valid[o] = ok ? 1 : 0;
}

__global__ void pad_map_kernel(const __nv_bfloat16* __restrict__ m, int B, int G, int D, int Prows,
                               int Pal, __nv_bfloat16* __restrict__ out) {
  const int dv = D / 8;
  const long long total = (long long)B * Prows * Pal * dv;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c8 = (int)(idx % dv);
  long long rest = idx / dv;
  const int col = (int)(rest % Pal);
  rest /= Pal;
  const int row = (int)(rest % Prows);
  const int b = (int)(rest / Prows);
  uint4 v = make_uint4(0, 0, 0, 0);
  if (col < 3 * G - 2) {
    const int i = min(max(row - (G - 1), 0), G - 1);
    const int j = min(max(col - (G - 1), 0), G - 1);
    v = __ldg(reinterpret_cast<const uint4*>(m + (((size_t)b * G + i) * G + j) * D + c8 * 8));
  }
  *reinterpret_cast<uint4*>(out + (((size_t)b * Prows + row) * Pal + col) * D + c8 * 8) = v;
}

// cnt[b,r,u,v] = sum_{i,j} qv[b,r,i,j] * mv[b,u-i,v-j]   (mv = 0 outside the map)
//
// flag[b*R] (the `den` output, overwritten afterwards by xcorr_den_kernel) = 1 if the whole map of example b is valid
// (every aerial-fused map is, bev_mapper.py:208-211): then the count is a rectangle sum of q_valid and the kernel
// below leaves the work to xcorr_count_allvalid_kernel.
__global__ void __launch_bounds__(256)
xcorr_allvalid_flag_kernel(const uint8_t* __restrict__ m_valid, int GG, int R, float* __restrict__ flag) {
  __shared__ int s_all;
  if (threadIdx.x == 0) s_all = 1;
  __syncthreads();
  const uint8_t* mv = m_valid + (size_t)blockIdx.x * GG;
  bool all = true;
  for (int k = threadIdx.x; k < GG; k += 256) all &= mv[k] != 0;
  if (!__all_sync(0xffffffffu, all) && (threadIdx.x & 31) == 0) s_all = 0;
  __syncthreads();
  if (threadIdx.x == 0) flag[(size_t)blockIdx.x * R] = s_all ? 1.f : 0.f;
}

constexpr int COUNT_UT = 8;  // consecutive u per block of the generic kernel

// Generic map validity.  One block per (b, r, tile of COUNT_UT shifts u, chunk of 256 v); rows as G-bit masks in shared
// memory.  Thread v walks the map rows a: the row shifted by v (W funnel shifts) is built once and and-popcounted
// against the template rows i = u - a of the COUNT_UT shifts of the tile (one broadcast LDS.128 per row at G = 128).
template <int W>  // words per row, G = 32*W
__global__ void __launch_bounds__(256)
xcorr_count_kernel(const uint8_t* __restrict__ t_valid, const uint8_t* __restrict__ m_valid, int R, int G,
                   int U, const float* __restrict__ flag, float* __restrict__ cnt) {
  __shared__ __align__(16) uint32_t qbits[32 * W][W];  // qbits[i][w] bit t = qv[i][32w+t]
  __shared__ uint32_t mrev[32 * W][W + 2];             // mrev[a][w] bit t = mv[a][G-1-(32w+t)], zero padded words
  const int br = blockIdx.z;  // b*R + r
  const int b = br / R;
  if (flag[(size_t)b * R] != 0.f) return;  // all-valid map: xcorr_count_allvalid_kernel
  const int u0 = blockIdx.y * COUNT_UT;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint8_t* qv = t_valid + (size_t)br * G * G;
  const uint8_t* mv = m_valid + (size_t)b * G * G;
  for (int task = warp; task < G * W; task += 8) {
    const int row = task / W, w = task % W;
    const uint32_t qb = __ballot_sync(0xffffffffu, qv[row * G + 32 * w + lane] != 0);
    const uint32_t mb = __ballot_sync(0xffffffffu, mv[row * G + (G - 1 - (32 * w + lane))] != 0);
    if (lane == 0) {
      qbits[row][w] = qb;
      mrev[row][w] = mb;
    }
  }
  if (threadIdx.x < G) {
    mrev[threadIdx.x][W] = 0;
    mrev[threadIdx.x][W + 1] = 0;
  }
  __syncthreads();
  const int v = blockIdx.x * 256 + threadIdx.x;
  if (v >= U) return;
  // sum_j qv[i][j] * mv[a][v-j] = popc(qrow_i & (mrev_a >> s)), s = G-1-v (left shift if negative)
  const int s = G - 1 - v;
  int total[COUNT_UT];
#pragma unroll
  for (int t = 0; t < COUNT_UT; ++t) total[t] = 0;
  const int a_lo = max(0, u0 - (G - 1)), a_hi = min(G - 1, min(U - 1, u0 + COUNT_UT - 1));
  for (int a = a_lo; a <= a_hi; ++a) {
    uint32_t x[W];
#pragma unroll
    for (int w = 0; w < W; ++w) {
      if (s >= 0) {
        const int ws = s >> 5, bs = s & 31;
        const int w0 = w + ws;
        const uint32_t lo = (w0 < W) ? mrev[a][w0] : 0u;
        const uint32_t hi = (w0 + 1 < W) ? mrev[a][w0 + 1] : 0u;
        x[w] = __funnelshift_r(lo, hi, bs);
      } else {
        const int t = -s;
        const int ws = t >> 5, bs = t & 31;
        const int w0 = w - ws;
        const uint32_t hi = (w0 >= 0) ? mrev[a][w0] : 0u;
        const uint32_t lo = (w0 - 1 >= 0) ? mrev[a][w0 - 1] : 0u;
        x[w] = __funnelshift_l(lo, hi, bs);
      }
    }
#pragma unroll
    for (int t = 0; t < COUNT_UT; ++t) {
      const int i = u0 + t - a;  // template row paired with map row a at shift u0 + t (warp-uniform)
      if (i >= 0 && i < G) {
        int c = 0;
#pragma unroll
        for (int w = 0; w < W; ++w) c += __popc(qbits[i][w] & x[w]);
        total[t] += c;
      }
    }
  }
#pragma unroll
  for (int t = 0; t < COUNT_UT; ++t)
    if (u0 + t < U) cnt[((size_t)br * U + u0 + t) * U + v] = (float)total[t];
}

constexpr int COUNT_USEG = 32;  // consecutive u per block of the all-valid kernel

// All-valid map: cnt[u,v] = sum of qv over rows [max(0,u-G+1), min(G-1,u)] x columns [max(0,v-G+1), min(G-1,v)].
// Thread v holds the column-range mask; the row window slides by one per u (add the entering row, drop the leaving one).
template <int W>
__global__ void __launch_bounds__(256)
xcorr_count_allvalid_kernel(const uint8_t* __restrict__ t_valid, int R, int G, int U, const float* __restrict__ flag,
                            float* __restrict__ cnt) {
  __shared__ __align__(16) uint32_t qbits[32 * W][W];
  const int br = blockIdx.z;
  const int b = br / R;
  if (flag[(size_t)b * R] == 0.f) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint8_t* qv = t_valid + (size_t)br * G * G;
  for (int task = warp; task < G * W; task += 8) {
    const int row = task / W, w = task % W;
    const uint32_t qb = __ballot_sync(0xffffffffu, qv[row * G + 32 * w + lane] != 0);
    if (lane == 0) qbits[row][w] = qb;
  }
  __syncthreads();
  const int v = blockIdx.x * 256 + threadIdx.x;
  if (v >= U) return;
  const int jlo = max(0, v - (G - 1)), jhi = min(G - 1, v);
  uint32_t mask[W];
#pragma unroll
  for (int w = 0; w < W; ++w) {
    const int lo = max(jlo - 32 * w, 0), hi = min(jhi - 32 * w, 31);  // bits [lo, hi] of word w
    mask[w] = hi < lo ? 0u : ((hi == 31 ? 0xffffffffu : ((1u << (hi + 1)) - 1u)) & ~((1u << lo) - 1u));
  }
  auto row_count = [&](int i) {
    int c = 0;
#pragma unroll
    for (int w = 0; w < W; ++w) c += __popc(qbits[i][w] & mask[w]);
    return c;
  };
  const int u0 = blockIdx.y * COUNT_USEG, u1 = min(U, u0 + COUNT_USEG);
  int ilo = max(0, u0 - (G - 1)), ihi = min(G - 1, u0);
  int total = 0;
  for (int i = ilo; i <= ihi; ++i) total += row_count(i);
  cnt[((size_t)br * U + u0) * U + v] = (float)total;
  for (int u = u0 + 1; u < u1; ++u) {
    const int nhi = min(G - 1, u), nlo = max(0, u - (G - 1));
    if (nhi > ihi) total += row_count(nhi);
    if (nlo > ilo) total -= row_count(ilo);
    ihi = nhi;
    ilo = nlo;
    cnt[((size_t)br * U + u) * U + v] = (float)total;
  }
}

__global__ void xcorr_den_kernel(const uint8_t* __restrict__ t_valid, int GG, float* __restrict__ den) {
  __shared__ int red[256];
  const uint8_t* p = t_valid + (size_t)blockIdx.x * GG;
  int s = 0;
  for (int i = threadIdx.x; i < GG; i += 256) s += p[i] != 0;
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) den[blockIdx.x] = (float)red[0];
}

}  // namespace snapb200

using namespace snapb200;

extern "C" {

/* rows per cell of the template tensor: R rounded up to the N tile (48) of the correlation GEMM */
int snapb200_xcorr_padded_rotations(int R) { return (R + 47) / 48 * 48; }

int snapb200_xcorr_padded_cols(int G) {
  const int U = 2 * G - 1;
  const int vt = (U + 127) / 128;
  return ((vt * 128 + G - 1) + 7) / 8 * 8;
}

int snapb200_rot_templates(const void* feats, const uint8_t* valid, const float* conf,
                           const float* rot_host, const float* centers, float cell_size, int B, int R,
                           int G, int D, void* templates, uint8_t* t_valid, void* stream) {
  SNAP_REQUIRE(feats && valid && rot_host && centers && templates && t_valid, "null pointer");
  SNAP_REQUIRE(R % 4 == 0 && R >= 4 && R <= 64, "num_rotations must be a multiple of 4 in [4, 64] (got %d)", R);
  SNAP_REQUIRE(D % 8 == 0, "D must be a multiple of 8");
  RotParams RP;
  RP.nq = R / 4;
  for (int i = 0; i < RP.nq; ++i)
    for (int j = 0; j < 4; ++j) RP.rot[i][j] = rot_host[i * 4 + j];
  const int RPad = snapb200_xcorr_padded_rotations(R);
  const long long total = (long long)B * RPad * G * G * (D / 8);
  rot_templates_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      RP, (const __nv_bfloat16*)feats, valid, conf, centers, cell_size, B, R, RPad, G, D,
      (__nv_bfloat16*)templates, t_valid);
  return check_launch("rot_templates_kernel");
}

int snapb200_xcorr_pad_map(const void* m, int B, int G, int D, void* out, void* stream) {
  SNAP_REQUIRE(m && out && D % 8 == 0, "bad arguments");
  const int Prows = 3 * G - 2, Pal = snapb200_xcorr_padded_cols(G);
  const long long total = (long long)B * Prows * Pal * (D / 8);
  pad_map_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)m, B, G, D, Prows, Pal, (__nv_bfloat16*)out);
  return check_launch("pad_map_kernel");
}

int snapb200_xcorr_count(const uint8_t* t_valid, const uint8_t* m_valid, int B, int R, int G, float* cnt,
                         float* den, void* stream) {
  SNAP_REQUIRE(t_valid && m_valid && cnt && den, "null pointer");
  SNAP_REQUIRE(G % 32 == 0 && G >= 32 && G <= 256, "grid side must be a multiple of 32 in [32, 256] (got %d)", G);
  const int U = 2 * G - 1;
  cudaStream_t s = (cudaStream_t)stream;
  // den[b*R] doubles as the "map b is all valid" flag until xcorr_den_kernel overwrites it at the end
  xcorr_allvalid_flag_kernel<<<B, 256, 0, s>>>(m_valid, G * G, R, den);
  int rc = check_launch("xcorr_allvalid_flag_kernel");
  if (rc) return rc;
  const dim3 grid((U + 255) / 256, (U + COUNT_UT - 1) / COUNT_UT, B * R);
  const dim3 grid_av((U + 255) / 256, (U + COUNT_USEG - 1) / COUNT_USEG, B * R);
#define SNAP_COUNT_CASE(W_)                                                                          \
  case W_:                                                                                           \
    xcorr_count_kernel<W_><<<grid, 256, 0, s>>>(t_valid, m_valid, R, G, U, den, cnt);                \
    rc = check_launch("xcorr_count_kernel");                                                         \
    if (rc) return rc;                                                                               \
    xcorr_count_allvalid_kernel<W_><<<grid_av, 256, 0, s>>>(t_valid, R, G, U, den, cnt);             \
    break;
  switch (G / 32) {
    SNAP_COUNT_CASE(1)
    SNAP_COUNT_CASE(2)
    SNAP_COUNT_CASE(3)
    SNAP_COUNT_CASE(4)
    SNAP_COUNT_CASE(5)
    SNAP_COUNT_CASE(6)
    SNAP_COUNT_CASE(7)
    default:
      xcorr_count_kernel<8><<<grid, 256, 0, s>>>(t_valid, m_valid, R, G, U, den, cnt);
      rc = check_launch("xcorr_count_kernel");
      if (rc) return rc;
      xcorr_count_allvalid_kernel<8><<<grid_av, 256, 0, s>>>(t_valid, R, G, U, den, cnt);
      break;
  }
#undef SNAP_COUNT_CASE
  rc = check_launch("xcorr_count_allvalid_kernel");
  if (rc) return rc;
  xcorr_den_kernel<<<B * R, 256, 0, s>>>(t_valid, G * G, den);
  return check_launch("xcorr_den_kernel");
}

/* scores f32 [B, R, 2G-1, 2G-1] = template_matching(templates bf16 [B,R,G,G,D], m_pad) with
   -inf where cnt <= thr (cnt may be NULL: no mask) and division by den[b,r] (NULL: none). */
int snapb200_xcorr_scores(const void* templates, const void* m_pad, const float* cnt, const float* den,
                          int B, int R, int G, int D, float thr, float* scores, void* stream) {
  SNAP_REQUIRE(templates && m_pad && scores, "null pointer");
  SNAP_REQUIRE(D == 32, "matching_dim must be 32 for the tensor-core correlation (got %d)", D);
  SNAP_REQUIRE(R >= 1 && G >= 8, "bad R/G");
  const int U = 2 * G - 1, Prows = 3 * G - 2, Pal = snapb200_xcorr_padded_cols(G);
  const int bn = 48;
  GemmParams p = {};
  p.xc_G = G;
  p.xc_P = Pal;
  p.xc_U = U;
  p.xc_vt = (U + 127) / 128;
  p.xc_R = R;
  p.xc_rows_per_b = (long long)Prows * Pal;
  const int RPad = snapb200_xcorr_padded_rotations(R);
  p.b_seg_rows = RPad;
  p.m_tiles = B * U * p.xc_vt;
  p.n_tiles = RPad / bn;
  p.nkb = G * G;
  p.kps = 1;
  p.seg_kstride = D;
  p.a_col0 = 0;
  p.seg_mode = SEG_XCORR;
  p.tile_mode = TILE_XCORR;
  p.epi = EPI_XCORR;
  p.out = scores;
  p.xc_cnt = cnt;
  p.xc_den = den;
  p.xc_thr = thr;
  SNAP_REQUIRE((long long)B * Prows * Pal < (1ll << 31), "padded map too large for 32-bit TMA rows");
  return launch_gemm(m_pad, (long long)B * Prows * Pal, D, D, templates, (long long)B * G * G * RPad, D, D, bn, 32,
                     p, (cudaStream_t)stream);
}

}  // extern "C"
