// Exhaustive correlation, sliding-window formulation (snap/models/pose_exhaustive_voting.py:83-103):
//
//   S[r, u, v] = sum_{i,j,d} q_r[i,j,d] * m_pad[u+i, v+j, d]
//
// For a fixed template row i the 128 map windows of consecutive v that the (i, j) terms need are all
// sub-ranges of ONE strip of the padded map, m_pad[u+i, v0 .. v0+127+G-1, :].  A CTA therefore keeps the
// strips of NU consecutive output rows in shared memory, stored un-swizzled as [4 channel chunks][pixel][16 B]
// (K-major "core matrix" layout with a uniform 16-byte row pitch), and slides the tcgen05 A-operand
// descriptor along the strip by 16 bytes per template column j.  Map traffic drops ~128x w.r.t. re-fetching a
// window per (i, j); the template tile of a cell (48 rotations x 32 channels, 3 KB) is fetched once per
// (i, j) and reused by the NU output rows.  Per (i, j): NU x 2 MMAs (M=128 shifts, N=48 rotations, K=16).
//
// Warp roles: 0 = TMA producer, 1 = TMEM owner + MMA issuer, 2..5 = epilogue (mask, normalise, store).
#include <cuda_bf16.h>
#include <math.h>

#include "common.cuh"
#include "host_common.h"

namespace snapb200 {

constexpr int XS_NU = 4;        // output rows (u) per CTA block
constexpr int XS_JB = 4;        // template cells per B load
constexpr int XS_N = 48;        // rotations per cell (padded)
constexpr int XS_SR = 6;        // strip ring slots (>= NU + 1)
constexpr int XS_BR = 8;        // template ring slots
constexpr int XS_STRIP_BYTES = 16384;            // 4 chunks x (<= 256 pixels) x 16 B
constexpr int XS_B_BYTES = 4 * XS_JB * XS_N * 16;  // 12288
constexpr int XS_SMEM = XS_SR * XS_STRIP_BYTES + XS_BR * XS_B_BYTES + 1024 + 512;
constexpr int XS_THREADS = 6 * 32;

struct XsParams {
  int B, R, G, U, Prows, Pal;
  int ublocks, vtiles, total_blocks;
  int SP;  // strip pixels = 128 + G - 1
  float* scores;
  const float* cnt;
  const float* den;
  float thr;
};

// K-major, no swizzle: 8-row core matrices of 16 B rows; SBO = 128 B between 8-row groups,
// LBO = byte distance between the two 16 B K-chunks of one K=16 MMA step.
__device__ __forceinline__ uint64_t make_plain_desc(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)(128u >> 4) << 32;
  d |= (uint64_t)1 << 46;
  return d;  // layout type 0 = SWIZZLE_NONE
}

__global__ void __launch_bounds__(XS_THREADS, 1)
xcorr_sw_kernel(const __grid_constant__ CUtensorMap tmM, const __grid_constant__ CUtensorMap tmT,
                const __grid_constant__ XsParams P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* strips = smem;
  uint8_t* bring = smem + XS_SR * XS_STRIP_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(bring + XS_BR * XS_B_BYTES);
  uint64_t* s_full = bars;                 // [XS_SR]
  uint64_t* s_empty = s_full + XS_SR;      // [XS_SR]
  uint64_t* b_full = s_empty + XS_SR;      // [XS_BR]
  uint64_t* b_empty = b_full + XS_BR;      // [XS_BR]
  uint64_t* t_full = b_empty + XS_BR;      // [2]
  uint64_t* t_empty = t_full + 2;          // [2]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(t_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int G = P.G, SP = P.SP;
  const uint32_t lbo_a = ((uint32_t)SP * 16u + 127u) & ~127u;  // chunk plane pitch of a strip (128 B aligned TMA destinations)
  const uint32_t lbo_b = (uint32_t)(XS_JB * XS_N) * 16u;  // chunk plane pitch of a template group
  const int strips_per_block = G + XS_NU - 1;
  const int jgroups = G / XS_JB;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmM);
    tma_prefetch_desc(&tmT);
    for (int s = 0; s < XS_SR; ++s) {
      mbar_init(&s_full[s], 1);
      mbar_init(&s_empty[s], 1);
    }
    for (int s = 0; s < XS_BR; ++s) {
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
        const int u0 = ub * XS_NU;
        const long long pix_base = (long long)b * P.Prows * P.Pal + (long long)vt * 128;
        const long long cell_base = (long long)b * G * G;
        int next_strip = 0;
        auto load_strip = [&](int t) {  // strip t = padded-map row u0 + t
          mbar_wait(&s_empty[ss], sph ^ 1);
          mbar_arrive_expect_tx(&s_full[ss], (uint32_t)SP * 64u);
          uint8_t* dst = strips + ss * XS_STRIP_BYTES;
          const int row = (int)(pix_base + (long long)(u0 + t) * P.Pal);
          for (int c = 0; c < 4; ++c) tma_load_2d(&tmM, &s_full[ss], dst + c * lbo_a, c * 8, row);
          if (++ss == XS_SR) {
            ss = 0;
            sph ^= 1;
          }
        };
        for (; next_strip < XS_NU; ++next_strip) load_strip(next_strip);
        for (int i = 0; i < G; ++i) {
          // keep the strip ring ahead of the MMA without ever waiting on the strip released by the previous step
          while (next_strip < strips_per_block && next_strip < i + XS_SR - 1) load_strip(next_strip++);
          for (int jg = 0; jg < jgroups; ++jg) {
            mbar_wait(&b_empty[bs], bph ^ 1);
            mbar_arrive_expect_tx(&b_full[bs], XS_B_BYTES);
            uint8_t* dst = bring + bs * XS_B_BYTES;
            const int row = (int)((cell_base + (long long)i * G + jg * XS_JB) * XS_N);
            for (int c = 0; c < 4; ++c) tma_load_2d(&tmT, &b_full[bs], dst + c * lbo_b, c * 8, row);
            if (++bs == XS_BR) {
              bs = 0;
              bph ^= 1;
            }
          }
        }
        while (next_strip < strips_per_block) load_strip(next_strip++);
      }
    }
  } else if (warp == 1) {
    // ============================ MMA issuer ============================
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16_m128(XS_N);
      int bs = 0;
      uint32_t bph = 0;
      long long strip_ctr = 0;   // strips consumed so far by this CTA (ring position of strip 0 of the block)
      long long strips_waited = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      const uint32_t s_strips = smem_u32(strips), s_bring = smem_u32(bring);
      for (int blk = blockIdx.x; blk < P.total_blocks; blk += gridDim.x) {
        mbar_wait(&t_empty[acc], acc_phase ^ 1);
        tc_fence_after_sync();
        const uint32_t tacc = tmem_base + (uint32_t)(acc * XS_NU * XS_N);
        for (int i = 0; i < G; ++i) {
          // strips i .. i+NU-1 must have landed
          while (strips_waited < strip_ctr + i + XS_NU) {
            mbar_wait(&s_full[strips_waited % XS_SR], (uint32_t)((strips_waited / XS_SR) & 1));
            ++strips_waited;
          }
          tc_fence_after_sync();
          for (int jg = 0; jg < jgroups; ++jg) {
            mbar_wait(&b_full[bs], bph);
            tc_fence_after_sync();
            const uint32_t sb = s_bring + bs * XS_B_BYTES;
#pragma unroll
            for (int jj = 0; jj < XS_JB; ++jj) {
              const int j = jg * XS_JB + jj;
#pragma unroll
              for (int uu = 0; uu < XS_NU; ++uu) {
                const uint32_t sa = s_strips + (uint32_t)((strip_ctr + i + uu) % XS_SR) * XS_STRIP_BYTES + (uint32_t)j * 16u;
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                  const uint64_t da = make_plain_desc(sa + (uint32_t)(2 * k) * lbo_a, lbo_a);
                  const uint64_t db = make_plain_desc(sb + (uint32_t)(jj * XS_N) * 16u + (uint32_t)(2 * k) * lbo_b, lbo_b);
                  umma_bf16(tacc + (uint32_t)(uu * XS_N), da, db, idesc, (i | j | k) != 0 ? 1u : 0u);
                }
              }
            }
            umma_commit(&b_empty[bs]);
            if (++bs == XS_BR) {
              bs = 0;
              bph ^= 1;
            }
          }
          umma_commit(&s_empty[(strip_ctr + i) % XS_SR]);  // strip i is not needed by later template rows
        }
        for (int t = G; t < strips_per_block; ++t) umma_commit(&s_empty[(strip_ctr + t) % XS_SR]);
        umma_commit(&t_full[acc]);
        // the producer loads every strip of the block even if unused; make sure they were all waited for
        while (strips_waited < strip_ctr + strips_per_block) {
          mbar_wait(&s_full[strips_waited % XS_SR], (uint32_t)((strips_waited / XS_SR) & 1));
          ++strips_waited;
        }
        strip_ctr += strips_per_block;
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
#pragma unroll 1
      for (int uu = 0; uu < XS_NU; ++uu) {
        const int u = ub * XS_NU + uu;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * XS_NU * XS_N + uu * XS_N);
#pragma unroll 1
        for (int c16 = 0; c16 < XS_N / 16; ++c16) {
          uint32_t vv[16];
          tmem_ld16(taddr + (uint32_t)(c16 * 16), vv);
          tmem_ld_wait();
          if (u < P.U && v_ok) {
#pragma unroll
            for (int jx = 0; jx < 16; ++jx) {
              const int r = c16 * 16 + jx;
              if (r < P.R) {
                const long long o = (((long long)b * P.R + r) * P.U + u) * P.U + v;
                float s = __uint_as_float(vv[jx]);
                if (P.cnt != nullptr && !(P.cnt[o] > P.thr)) s = -INFINITY;
                if (P.den != nullptr) s = s / P.den[b * P.R + r];
                P.scores[o] = s;
              }
            }
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

/* Sliding-window correlation: same contract as snapb200_xcorr_scores; requires padded_rotations(R) == 48,
   G a multiple of 4 and 128 + G - 1 <= 256. */
extern "C" int snapb200_xcorr_scores_sw(const void* templates, const void* m_pad, const float* cnt, const float* den,
                                        int B, int R, int G, int D, float thr, float* scores, void* stream) {
  SNAP_REQUIRE(templates && m_pad && scores, "null pointer");
  SNAP_REQUIRE(D == 32, "matching_dim must be 32 (got %d)", D);
  SNAP_REQUIRE(snapb200_xcorr_padded_rotations(R) == XS_N, "sliding-window correlation needs num_rotations <= 48");
  SNAP_REQUIRE(G % XS_JB == 0 && G >= 8 && 128 + G - 1 <= 256, "sliding-window correlation needs G %% 4 == 0, G <= 129");
  static DynSmemState smem_state;
  if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(&xcorr_sw_kernel), XS_SMEM, &smem_state,
                               "cudaFuncSetAttribute(xcorr_sw)"))
    return rc;
  XsParams P;
  P.B = B; P.R = R; P.G = G; P.U = 2 * G - 1;
  P.Prows = 3 * G - 2; P.Pal = snapb200_xcorr_padded_cols(G);
  P.ublocks = (P.U + XS_NU - 1) / XS_NU;
  P.vtiles = (P.U + 127) / 128;
  P.total_blocks = B * P.ublocks * P.vtiles;
  P.SP = 128 + G - 1;
  P.scores = scores; P.cnt = cnt; P.den = den; P.thr = thr;
  SNAP_REQUIRE((long long)B * P.Prows * P.Pal < (1ll << 31) && (long long)B * G * G * XS_N < (1ll << 31),
               "tensor too large for 32-bit TMA row coordinates");
  CUtensorMap tmM, tmT;
  int rc = make_tmap_2d_bf16_plain(&tmM, m_pad, (long long)B * P.Prows * P.Pal, 32, 32, P.SP, 8);
  if (rc) return rc;
  rc = make_tmap_2d_bf16_plain(&tmT, templates, (long long)B * G * G * XS_N, 32, 32, XS_JB * XS_N, 8);
  if (rc) return rc;
  const int grid = P.total_blocks < num_sms() ? P.total_blocks : num_sms();
  xcorr_sw_kernel<<<grid, XS_THREADS, XS_SMEM, (cudaStream_t)stream>>>(tmM, tmT, P);
  return check_launch("xcorr_sw_kernel");
}
