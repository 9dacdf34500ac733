// Exhaustive correlation, map-row-major formulation (snap/models/pose_exhaustive_voting.py:83-103):
//
//   S[r, u, v] = sum_{i,j,d} q_r[i,j,d] * m_pad[u+i, v+j, d]
//
// A CTA owns NU consecutive output rows u0 .. u0+NU-1 and 128 consecutive shifts v.  It walks over the padded-map
// rows t (strip t = m_pad[u0+t, v0 .. v0+127+G-1, :], resident in shared memory, un-swizzled [4 chunks][pixel][16 B])
// and, for every template column j, issues ONE tcgen05.mma of shape M=128 (shifts) x N=NU*R x K=16 whose B operand
// stacks the NU template rows i = t-NU+1+s (s = 0..NU-1) that pair strip t with the CTA's output rows
// u = u0 + NU-1-s: accumulator column s*R + r belongs to (u0 + NU-1-s, r), so every MMA accumulates in place and
// no partial result ever leaves TMEM.  The A descriptor slides along the strip by 16 bytes per template column.
//
// Why this shape: a tcgen05.mma streams its A tile (128 x 16 bf16 = 4 KB) from shared memory at <= 128 B/clk, so an
// N = 48 instruction (24 clk of tensor time) is bound by its operand reads (~44 clk).  With N = NU*R = 144 the A read
// is amortised over 3x more columns and no rotation padding is multiplied (R = 36 is not a multiple of 16, 4*36 is).
// Templates outside [0, G) (first / last strips of a block) are zero-filled by TMA (out-of-bounds box rows).
//
// Warp roles: 0 = TMA producer, 1 = TMEM owner + MMA issuer, 2..5 = epilogue (mask, normalise, store).
#include <cuda_bf16.h>
#include <math.h>

#include "common.cuh"
#include "host_common.h"

namespace snapb200 {

constexpr int XR_NU = 4;        // output rows per CTA block = template rows stacked in one MMA
constexpr int XR_JB = 2;        // template columns per B load
constexpr int XR_SR = 3;        // strip ring slots
constexpr int XR_STRIP_BYTES = 16384;  // 4 chunks x (<= 256 pixels) x 16 B
constexpr int XR_THREADS = 6 * 32;
constexpr int XR_SMEM_MAX = 227 * 1024;

struct XrParams {
  int B, R, G, U, Prows, Pal, N;
  int ublocks, vtiles, total_blocks;
  int SP;          // strip pixels = 128 + G - 1
  int b_bytes;     // bytes of one B ring slot = 4 * JB * N * 16
  int b_slots;     // ring depth
  float* scores;
  const float* cnt;
  const float* den;
  float thr;
};

// K-major, no swizzle: 8-row core matrices of 16 B rows; SBO = 128 B between 8-row groups,
// LBO = byte distance between the two 16 B K-chunks of one K=16 MMA step.
__device__ __forceinline__ uint64_t xr_plain_desc(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)(128u >> 4) << 32;
  d |= (uint64_t)1 << 46;
  return d;  // layout type 0 = SWIZZLE_NONE
}

__device__ __forceinline__ void tma_load_5d(const CUtensorMap* m, uint64_t* bar, void* smem_dst, int c0, int c1, int c2,
                                            int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
        "r"(c3), "r"(c4)
      : "memory");
}

// templates cell-major [B][i][j][RP][D=32] -> chunk-major [B][c = d/8][j][i][R][8]: the 16-byte K-chunk c of the NU
// stacked template rows of one column j is ONE contiguous run of NU*R*16 bytes, so a B-operand plane is fetched by
// TMA as a few long pieces (a TMA box row of 16 bytes per template would be request-rate bound).
__global__ void xr_relayout_kernel(const uint4* __restrict__ src, int B, int G, int RP, int R, uint4* __restrict__ dst) {
  const long long total = (long long)B * G * G * R * 4;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  long long rest = idx;
  const int r = (int)(rest % R);
  rest /= R;
  const int i = (int)(rest % G);
  rest /= G;
  const int j = (int)(rest % G);
  rest /= G;
  const int c = (int)(rest & 3);
  const int b = (int)(rest >> 2);
  dst[idx] = __ldg(src + ((((size_t)b * G + i) * G + j) * RP + r) * 4 + c);
}

__global__ void __launch_bounds__(XR_THREADS, 1)
xcorr_rows_kernel(const __grid_constant__ CUtensorMap tmM, const __grid_constant__ CUtensorMap tmT,
                  const __grid_constant__ XrParams P) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* strips = smem;
  uint8_t* bring = smem + XR_SR * XR_STRIP_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(bring + (size_t)P.b_slots * P.b_bytes);
  uint64_t* s_full = bars;                 // [XR_SR]
  uint64_t* s_empty = s_full + XR_SR;      // [XR_SR]
  uint64_t* b_full = s_empty + XR_SR;      // [16]
  uint64_t* b_empty = b_full + 16;         // [16]
  uint64_t* t_full = b_empty + 16;         // [2]
  uint64_t* t_empty = t_full + 2;          // [2]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(t_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int G = P.G, SP = P.SP, N = P.N, BS = P.b_slots;
  const uint32_t lbo_a = ((uint32_t)SP * 16u + 127u) & ~127u;  // chunk plane pitch of a strip
  const uint32_t lbo_b = (uint32_t)(XR_JB * N) * 16u;          // chunk plane pitch of a B slot
  const int strips_per_block = G + XR_NU - 1;
  const int jgroups = G / XR_JB;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmM);
    tma_prefetch_desc(&tmT);
    for (int s = 0; s < XR_SR; ++s) {
      mbar_init(&s_full[s], 1);
      mbar_init(&s_empty[s], 1);
    }
    for (int s = 0; s < 16; ++s) {
      mbar_init(&b_full[s], 1);
      mbar_init(&b_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&t_full[s], 1);
      mbar_init(&t_empty[s], 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr_smem, 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ============================ TMA producer ============================
    if (elect_one()) {
      int ss = 0, bs = 0;
      uint32_t sph = 0, bph = 0;
      for (int blk = blockIdx.x; blk < P.total_blocks; blk += gridDim.x) {
        const int vt = blk % P.vtiles;
        const int ub = (blk / P.vtiles) % P.ublocks;
        const int b = blk / (P.vtiles * P.ublocks);
        const int u0 = ub * XR_NU;
        const long long pix_base = (long long)b * P.Prows * P.Pal + (long long)vt * 128;
        for (int t = 0; t < strips_per_block; ++t) {
          // strip t = padded-map row u0 + t (rows past the padded map only meet zero templates: clamp the address)
          mbar_wait(&s_empty[ss], sph ^ 1);
          mbar_arrive_expect_tx(&s_full[ss], (uint32_t)SP * 64u);
          uint8_t* sdst = strips + ss * XR_STRIP_BYTES;
          const int prow = min(u0 + t, P.Prows - 1);
          const int row = (int)(pix_base + (long long)prow * P.Pal);
          for (int c = 0; c < 4; ++c) tma_load_2d(&tmM, &s_full[ss], sdst + c * lbo_a, c * 8, row);
          if (++ss == XR_SR) {
            ss = 0;
            sph ^= 1;
          }
          const int i0 = t - (XR_NU - 1);  // template rows i0 .. i0+NU-1 (out of range -> zero fill)
          for (int jg = 0; jg < jgroups; ++jg) {
            mbar_wait(&b_empty[bs], bph ^ 1);
            mbar_arrive_expect_tx(&b_full[bs], (uint32_t)P.b_bytes);
            uint8_t* dst = bring + (size_t)bs * P.b_bytes;
            for (int c = 0; c < 4; ++c) tma_load_5d(&tmT, &b_full[bs], dst + c * lbo_b, 0, 0, i0, jg * XR_JB, b * 4 + c);
            if (++bs == BS) {
              bs = 0;
              bph ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ============================ MMA issuer ============================
    if (elect_one()) {
      const uint32_t idesc = make_idesc_bf16_m128(N);
      int ss = 0, bs = 0;
      uint32_t sph = 0, bph = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      const uint32_t s_strips = smem_u32(strips), s_bring = smem_u32(bring);
      for (int blk = blockIdx.x; blk < P.total_blocks; blk += gridDim.x) {
        mbar_wait(&t_empty[acc], acc_phase ^ 1);
        tc_fence_after_sync();
        const uint32_t tacc = tmem_base + (uint32_t)(acc * N);
        for (int t = 0; t < strips_per_block; ++t) {
          mbar_wait(&s_full[ss], sph);
          tc_fence_after_sync();
          const uint32_t s_strip = s_strips + (uint32_t)ss * XR_STRIP_BYTES;
          for (int jg = 0; jg < jgroups; ++jg) {
            mbar_wait(&b_full[bs], bph);
            tc_fence_after_sync();
            const uint32_t sb = s_bring + (uint32_t)bs * (uint32_t)P.b_bytes;
#pragma unroll
            for (int jj = 0; jj < XR_JB; ++jj) {
              const int j = jg * XR_JB + jj;
              const uint32_t sa = s_strip + (uint32_t)j * 16u;
#pragma unroll
              for (int k = 0; k < 2; ++k) {
                const uint64_t da = xr_plain_desc(sa + (uint32_t)(2 * k) * lbo_a, lbo_a);
                const uint64_t db = xr_plain_desc(sb + (uint32_t)(jj * N) * 16u + (uint32_t)(2 * k) * lbo_b, lbo_b);
                umma_bf16(tacc, da, db, idesc, (t | j | k) != 0 ? 1u : 0u);
              }
            }
            umma_commit(&b_empty[bs]);
            if (++bs == BS) {
              bs = 0;
              bph ^= 1;
            }
          }
          umma_commit(&s_empty[ss]);
          if (++ss == XR_SR) {
            ss = 0;
            sph ^= 1;
          }
        }
        umma_commit(&t_full[acc]);
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else {
    // ============================ epilogue ============================
    const int q = warp & 3;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int blk = blockIdx.x; blk < P.total_blocks; blk += gridDim.x) {
      const int vt = blk % P.vtiles;
      const int ub = (blk / P.vtiles) % P.ublocks;
      const int b = blk / (P.vtiles * P.ublocks);
      mbar_wait(&t_full[acc], acc_phase);
      tc_fence_after_sync();
      const int v = vt * 128 + q * 32 + lane;
      const bool v_ok = v < P.U;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * N);
#pragma unroll 1
      for (int c16 = 0; c16 < N / 16; ++c16) {
        uint32_t vv[16];
        tmem_ld16(taddr + (uint32_t)(c16 * 16), vv);
        tmem_ld_wait();
#pragma unroll
        for (int jx = 0; jx < 16; ++jx) {
          const int col = c16 * 16 + jx;
          const int s = col / P.R, r = col - s * P.R;
          const int u = ub * XR_NU + (XR_NU - 1 - s);
          if (u < P.U && v_ok) {
            const long long o = (((long long)b * P.R + r) * P.U + u) * P.U + v;
            float sc = __uint_as_float(vv[jx]);
            if (P.cnt != nullptr && !(P.cnt[o] > P.thr)) sc = -INFINITY;
            if (P.den != nullptr) sc = sc / P.den[b * P.R + r];
            P.scores[o] = sc;
          }
        }
      }
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&t_empty[acc]);
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace snapb200

using namespace snapb200;

extern "C" int snapb200_xcorr_padded_cols(int G);
extern "C" int snapb200_xcorr_padded_rotations(int R);

extern "C" size_t snapb200_xcorr_scores_rows_workspace(int B, int R, int G, int D) {
  return (size_t)B * G * G * R * D * 2;
}

/* Map-row-major correlation (see the file header): same contract as snapb200_xcorr_scores, plus a caller-provided
   workspace of snapb200_xcorr_scores_rows_workspace() bytes (re-laid-out templates).  Requires D == 32,
   R % 4 == 0, 4 * R <= 256, G even and 128 + G - 1 <= 256. */
extern "C" int snapb200_xcorr_scores_rows(const void* templates, const void* m_pad, const float* cnt, const float* den,
                                          int B, int R, int G, int D, float thr, float* scores, void* workspace,
                                          size_t workspace_bytes, void* stream) {
  SNAP_REQUIRE(templates && m_pad && scores && workspace, "null pointer");
  SNAP_REQUIRE(D == 32, "matching_dim must be 32 (got %d)", D);
  const int N = XR_NU * R;
  SNAP_REQUIRE(R % 4 == 0 && N % 16 == 0 && N >= 16 && N <= 256, "num_rotations must be a multiple of 4, <= 64 (got %d)", R);
  SNAP_REQUIRE(G % XR_JB == 0 && G >= 8 && 128 + G - 1 <= 256, "row-major correlation needs an even G <= 129");
  SNAP_REQUIRE(workspace_bytes >= snapb200_xcorr_scores_rows_workspace(B, R, G, D), "workspace too small");
  SNAP_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 15) == 0, "workspace must be 16-byte aligned");
  static DynSmemState smem_state;
  if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(&xcorr_rows_kernel), XR_SMEM_MAX, &smem_state,
                               "cudaFuncSetAttribute(xcorr_rows)"))
    return rc;
  cudaStream_t s = (cudaStream_t)stream;
  const int RP = snapb200_xcorr_padded_rotations(R);
  {
    const long long total = (long long)B * G * G * R * 4;
    xr_relayout_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>((const uint4*)templates, B, G, RP, R,
                                                                       (uint4*)workspace);
    int rc = check_launch("xr_relayout_kernel");
    if (rc) return rc;
  }
  XrParams P;
  P.B = B; P.R = R; P.G = G; P.U = 2 * G - 1; P.N = N;
  P.Prows = 3 * G - 2; P.Pal = snapb200_xcorr_padded_cols(G);
  P.ublocks = (P.U + XR_NU - 1) / XR_NU;
  P.vtiles = (P.U + 127) / 128;
  P.total_blocks = B * P.ublocks * P.vtiles;
  P.SP = 128 + G - 1;
  P.b_bytes = 4 * XR_JB * N * 16;
  const int avail = XR_SMEM_MAX - XR_SR * XR_STRIP_BYTES - 1024;
  P.b_slots = avail / P.b_bytes;
  if (P.b_slots > 16) P.b_slots = 16;
  SNAP_REQUIRE(P.b_slots >= 2, "not enough shared memory for the template ring");
  P.scores = scores; P.cnt = cnt; P.den = den; P.thr = thr;
  SNAP_REQUIRE((long long)B * P.Prows * P.Pal < (1ll << 31), "tensor too large for 32-bit TMA row coordinates");
  CUtensorMap tmM, tmT;
  int rc = make_tmap_2d_bf16_plain(&tmM, m_pad, (long long)B * P.Prows * P.Pal, 32, 32, P.SP, 8);
  if (rc) return rc;
  {
    // [b*4 + c][j][i][r_hi (2)][r_lo (R/2) x 8 d]: inner pieces of R/2 * 16 bytes
    const unsigned long long half = (unsigned long long)(R / 2) * 8;  // elements of the innermost dimension
    const unsigned long long dims[5] = {half, 2, (unsigned long long)G, (unsigned long long)G, (unsigned long long)B * 4};
    const unsigned long long strides[4] = {half * 2, 16ull * R, 16ull * R * G, 16ull * R * G * G};
    const unsigned box[5] = {(unsigned)half, 2, XR_NU, XR_JB, 1};
    rc = make_tmap_nd_bf16_plain(&tmT, workspace, 5, dims, strides, box);
    if (rc) return rc;
  }
  const int smem_bytes = XR_SR * XR_STRIP_BYTES + P.b_slots * P.b_bytes + 1024;
  const int grid = P.total_blocks < num_sms() ? P.total_blocks : num_sms();
  xcorr_rows_kernel<<<grid, XR_THREADS, smem_bytes, s>>>(tmM, tmT, P);
  return check_launch("xcorr_rows_kernel");
}
