// tcgen05 GEMM engine (sm_100a): D[M,N] = sum over K-segments of A_seg[M,K] * B[N,K]^T.
//
// One persistent, warp-specialised kernel serves every GEMM-shaped op on the hot path:
//   * 1x1 convs / Dense layers               : one K-segment
//   * 3x3 convs (stride 1 and 2)             : 9 K-segments = 9 row-shifted views of the same
//                                              zero-bordered NHWC activation buffer
//   * exhaustive (x,y,theta) correlation     : G*G K-segments = shifted windows of the edge-padded map
// A and B tiles are fetched by TMA into 128B/64B-swizzled shared memory, multiplied by
// tcgen05.mma (M=128, N=BN, K=16 per instruction) into a double-buffered TMEM accumulator and
// drained by four epilogue warps (tcgen05.ld -> registers -> fused epilogue -> global).
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer,
// warps 2..5 = epilogue (warp w owns TMEM lanes 32*(w%4) .. +31).
#pragma once
#include "common.cuh"

namespace snapb200 {

enum { SEG_TABLE = 0, SEG_XCORR = 1 };
enum { TILE_LINEAR = 0, TILE_XCORR = 1 };
enum { EPI_STORE = 0, EPI_XCORR = 1 };
constexpr int GN_REPLICAS = 8;  // independent accumulator copies; consumers sum them

struct GemmParams {
  int stages;       // smem pipeline depth (2..GemmCfg::MAX_STAGES), chosen by the host per layer
  int m_tiles, n_tiles;
  int nkb;          // K blocks per tile (= num_seg * kps)
  int kps;          // K blocks per segment
  int seg_kstride;  // B column advance per segment (elements)
  int a_col0;       // first A column (elements)
  int seg_mode;
  int seg_off[9];   // SEG_TABLE: A row offset of each segment
  int b_seg_rows;   // 0: B = [N, segments*K] (K-major rows); > 0: B = [segment][b_seg_rows][K] row blocks
  // SEG_XCORR / TILE_XCORR geometry
  int xc_G, xc_P, xc_U, xc_vt, xc_R;
  long long xc_rows_per_b;
  int tile_mode;
  // epilogue
  int epi;
  long long M_valid;
  int N;
  void* out;
  long long ldo;
  int out_f32;
  const __nv_bfloat16* residual;
  long long ldr;
  const float* bias;
  const uint8_t* row_mask;
  int relu;
  // row remap: m -> (img, r, c) over (rm_R, rm_C); valid iff r0<=r<r0+Ho && c0<=c<c0+Wo
  int remap, rm_R, rm_C, rm_r0, rm_c0, rm_Ho, rm_Wo;
  // fused GroupNorm statistics of the stored output (see snapb200.h)
  double* gn_acc;
  double* gn_acc_relu;
  long long gn_rows_per_img;
  int gn_cpg;  // channels per group = N / 32
  int gn_replica_stride;  // doubles between the GN_REPLICAS copies of the accumulator (contention spreading)
  // EPI_XCORR
  const float* xc_cnt;
  const float* xc_den;
  float xc_thr;
};

constexpr int GEMM_EPI_WARPS = 8;                       // two per TMEM lane quadrant
constexpr int GEMM_THREADS = (2 + GEMM_EPI_WARPS) * 32;  // 320

template <int BN, int BK>
struct GemmCfg {
  static constexpr int BM = 128;
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES_RAW = (196 * 1024) / STAGE_BYTES;
  static constexpr int MAX_STAGES = STAGES_RAW > 16 ? 16 : STAGES_RAW;
  static constexpr int TMEM_COLS = (2 * BN <= 32)    ? 32
                                   : (2 * BN <= 64)  ? 64
                                   : (2 * BN <= 128) ? 128
                                   : (2 * BN <= 256) ? 256
                                                     : 512;
  // +1024 for manual alignment, +512 for barriers (2 x 16 stage + 4 accumulator) / tmem pointer
  static constexpr int smem_bytes(int stages) { return stages * STAGE_BYTES + 1024 + 512; }
  static_assert(BN % 16 == 0 && BN >= 16 && BN <= 256, "UMMA N constraint for M=128");
  static_assert(BK == 64 || BK == 32, "BK is one swizzle span: 128B or 64B of bf16");
  static_assert(A_BYTES % 1024 == 0 && B_BYTES % 1024 == 0, "stage tiles must stay 1024B aligned");
  static_assert(MAX_STAGES >= 2, "need at least a double buffer");
};

struct TileCoord {
  int mt, nt;
  int a_row;  // first A row of the tile (before the per-segment offset)
  int b_row;  // first B row
};

__device__ __forceinline__ TileCoord decode_tile(const GemmParams& p, int tile, int bn) {
  TileCoord t;
  t.nt = tile % p.n_tiles;
  t.mt = tile / p.n_tiles;
  if (p.tile_mode == TILE_LINEAR) {
    t.a_row = t.mt * 128;
    t.b_row = t.nt * bn;
  } else {
    int vt = t.mt % p.xc_vt;
    int u = (t.mt / p.xc_vt) % p.xc_U;
    int b = t.mt / (p.xc_vt * p.xc_U);
    t.a_row = (int)(b * p.xc_rows_per_b) + u * p.xc_P + vt * 128;
    t.b_row = p.b_seg_rows == 0 ? b * p.xc_R + t.nt * bn : b * p.xc_G * p.xc_G * p.b_seg_rows + t.nt * bn;
  }
  return t;
}

__device__ __forceinline__ int seg_row_offset(const GemmParams& p, int seg) {
  if (p.seg_mode == SEG_TABLE) return p.seg_off[seg];
  int i = seg / p.xc_G;
  int j = seg - i * p.xc_G;
  return i * p.xc_P + j;
}

// Accumulate (sum, sumsq) of 16 consecutive output columns of one row into the GroupNorm accumulators.
// Rows of a warp normally belong to one image: reduce over the 32 rows with shuffles and issue one
// double atomic per (group, moment); warps straddling an image boundary fall back to per-row atomics.
template <int SPAN>  // columns per group inside the 16-column chunk: 2, 4, 8 or 16
__device__ __forceinline__ void gn_accumulate16_t(const float (&v)[16], bool row_ok, int img, int col,
                                                  int cpg, double* acc, bool uniform, int img_ref,
                                                  int lane) {
#pragma unroll
  for (int g = 0; g < 16 / SPAN; ++g) {
    float s = 0.f, q = 0.f;
#pragma unroll
    for (int j = 0; j < SPAN; ++j) {
      s += v[g * SPAN + j];
      q += v[g * SPAN + j] * v[g * SPAN + j];
    }
    if (!row_ok) s = q = 0.f;
    const int group = (col + g * SPAN) / cpg;
    if (uniform) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        q += __shfl_xor_sync(0xffffffffu, q, o);
      }
      if (lane == 0 && img_ref >= 0) {
        atomicAdd(&acc[((size_t)img_ref * 32 + group) * 2 + 0], (double)s);
        atomicAdd(&acc[((size_t)img_ref * 32 + group) * 2 + 1], (double)q);
      }
    } else if (row_ok) {
      atomicAdd(&acc[((size_t)img * 32 + group) * 2 + 0], (double)s);
      atomicAdd(&acc[((size_t)img * 32 + group) * 2 + 1], (double)q);
    }
  }
}

__device__ __forceinline__ void gn_accumulate16(const float (&v)[16], bool row_ok, int img, int col,
                                                int cpg, double* acc, bool uniform, int img_ref,
                                                int lane) {
  switch (cpg) {
    case 2: gn_accumulate16_t<2>(v, row_ok, img, col, cpg, acc, uniform, img_ref, lane); break;
    case 4: gn_accumulate16_t<4>(v, row_ok, img, col, cpg, acc, uniform, img_ref, lane); break;
    case 8: gn_accumulate16_t<8>(v, row_ok, img, col, cpg, acc, uniform, img_ref, lane); break;
    default: gn_accumulate16_t<16>(v, row_ok, img, col, cpg, acc, uniform, img_ref, lane); break;
  }
}

template <int BN, int BK>
__global__ void __launch_bounds__(GEMM_THREADS, 2)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const GemmParams p) {
  using Cfg = GemmCfg<BN, BK>;
  const int STAGES = p.stages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_tiles = p.m_tiles * p.n_tiles;
  // contiguous tile range per CTA (same image / same A rows stay together: L2 locality, fewer GN flushes)
  const int tile_begin = (int)((long long)blockIdx.x * total_tiles / gridDim.x);
  const int tile_end = (int)((long long)(blockIdx.x + 1) * total_tiles / gridDim.x);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], GEMM_EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr_smem, Cfg::TMEM_COLS);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = tile_begin; tile < tile_end; ++tile) {
        const TileCoord t = decode_tile(p, tile, BN);
        int seg = 0, kc = 0, xi = 0, xj = 0;  // (xi, xj): window shift of the SEG_XCORR segment
        for (int kb = 0; kb < p.nkb; ++kb) {
          const int a_off = p.seg_mode == SEG_TABLE ? p.seg_off[seg] : xi * p.xc_P + xj;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
          uint8_t* sb = sa + Cfg::A_BYTES;
          tma_load_2d(&tmA, &full_bar[stage], sa, p.a_col0 + kc * BK, t.a_row + a_off);
          if (p.b_seg_rows == 0)
            tma_load_2d(&tmB, &full_bar[stage], sb, seg * p.seg_kstride + kc * BK, t.b_row);
          else  // B stored segment-major: [segment][rows][BK] (contiguous tile per segment)
            tma_load_2d(&tmB, &full_bar[stage], sb, kc * BK, t.b_row + seg * p.b_seg_rows);
          if (++kc == p.kps) {
            kc = 0;
            ++seg;
            if (++xj == p.xc_G) {
              xj = 0;
              ++xi;
            }
          }
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16_m128(BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = tile_begin; tile < tile_end; ++tile) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after_sync();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < p.nkb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after_sync();
          const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint32_t sb = sa + Cfg::A_BYTES;
          const uint64_t da = make_kmajor_desc<BK * 2>(sa);
          const uint64_t db = make_kmajor_desc<BK * 2>(sb);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // advancing K by 16 bf16 = 32 B inside the swizzle span: +2 in the (addr >> 4) field
            umma_bf16(tmem_d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc,
                      (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs retire
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tmem_full[acc]);  // accumulator ready for the epilogue
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else {
    // ===================== epilogue warps =====================
    const int q = warp & 3;         // TMEM lane quadrant this warp may access
    const int hw = (warp - 2) >> 2;  // 0/1: which half of the 16-column chunks (interleaved) it drains
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = tile_begin; tile < tile_end; ++tile) {
      const TileCoord t = decode_tile(p, tile, BN);
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after_sync();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN);
      const int row_in_tile = q * 32 + lane;
      const int n0 = t.nt * BN;

      if (p.epi == EPI_STORE) {
        const long long m = (long long)t.mt * 128 + row_in_tile;
        bool row_ok = m < p.M_valid;
        long long orow = m;
        if (p.remap) {
          const long long per_img = (long long)p.rm_R * p.rm_C;
          const long long img = m / per_img;
          const int rem = (int)(m - img * per_img);
          const int r = rem / p.rm_C;
          const int c = rem - r * p.rm_C;
          row_ok = row_ok && r >= p.rm_r0 && r < p.rm_r0 + p.rm_Ho && c >= p.rm_c0 &&
                   c < p.rm_c0 + p.rm_Wo;
          orow = (img * p.rm_Ho + (r - p.rm_r0)) * p.rm_Wo + (c - p.rm_c0);
        }
        if (!row_ok) orow = 0;
        const bool keep = row_ok && (p.row_mask == nullptr || p.row_mask[orow] != 0);
        int gn_img = -1, gn_ref = -1;
        bool gn_uniform = true;
        if (p.gn_acc != nullptr) {
          gn_img = row_ok ? (int)(orow / p.gn_rows_per_img) : -1;
          gn_ref = __reduce_max_sync(0xffffffffu, gn_img);
          gn_uniform = __all_sync(0xffffffffu, gn_img == gn_ref || gn_img == -1);
        }
        const bool rnd = !p.out_f32;
        const bool has_res = p.residual != nullptr;
        const __nv_bfloat16* res_row = has_res ? p.residual + orow * p.ldr : nullptr;

        // One 16-column chunk: all roundings the reference applies (dot -> dtype, + bias -> dtype,
        // + residual -> dtype), fused GroupNorm statistics of exactly what is stored, then the store.
        auto process = [&](const uint32_t (&v)[16], const uint4& r0, const uint4& r1, int c16) {
          const int col = n0 + c16 * 16;
          if (col >= p.N) return;  // warp-uniform
          float f[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) f[j] = rnd ? bf16_round(__uint_as_float(v[j])) : __uint_as_float(v[j]);
          if (p.bias) {
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col) + j4);
              const float bs[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float x = f[j4 * 4 + e] + bs[e];
                f[j4 * 4 + e] = rnd ? bf16_round(x) : x;
              }
            }
          }
          if (has_res) {
            const uint32_t rr[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float2 x = unpack_bf16(rr[j]);
              f[2 * j] += x.x;
              f[2 * j + 1] += x.y;
            }
          }
          if (p.relu) {
#pragma unroll
            for (int j = 0; j < 16; ++j) f[j] = fmaxf(f[j], 0.f);
          }
          if (!keep) {
#pragma unroll
            for (int j = 0; j < 16; ++j) f[j] = 0.f;
          }
          if (p.out_f32) {
            if (row_ok) {
              float4* op = reinterpret_cast<float4*>(static_cast<float*>(p.out) + orow * p.ldo + col);
#pragma unroll
              for (int j = 0; j < 4; ++j)
                op[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
            }
            return;
          }
          uint32_t pk[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) pk[j] = pack_bf16(f[2 * j], f[2 * j + 1]);
          if (row_ok) {
            uint4* op = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.out) + orow * p.ldo + col);
            op[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            op[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
          }
          if (p.gn_acc != nullptr) {
            float g[16];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float2 x = unpack_bf16(pk[j]);  // the stored (rounded) values
              g[2 * j] = x.x;
              g[2 * j + 1] = x.y;
            }
            const size_t rep = (size_t)(blockIdx.x % GN_REPLICAS) * p.gn_replica_stride;
            gn_accumulate16(g, row_ok, gn_img, col, p.gn_cpg, p.gn_acc + rep, gn_uniform, gn_ref, lane);
            if (p.gn_acc_relu != nullptr) {
#pragma unroll
              for (int j = 0; j < 16; ++j) g[j] = fmaxf(g[j], 0.f);
              gn_accumulate16(g, row_ok, gn_img, col, p.gn_cpg, p.gn_acc_relu + rep, gn_uniform, gn_ref, lane);
            }
          }
        };
        auto load_res = [&](int c16, uint4& r0, uint4& r1) {
          const int col = n0 + c16 * 16;
          if (has_res && row_ok && col < p.N) {
            const uint4* rp = reinterpret_cast<const uint4*>(res_row + col);
            r0 = __ldg(rp);
            r1 = __ldg(rp + 1);
          }
        };
        // software pipeline over this warp's chunks (hw, hw+2, ...): the TMEM load and the residual
        // load of the next chunk are in flight while the current chunk is processed
        constexpr int NCH = BN / 16;
        uint32_t va[16], vb[16];
        uint4 ra0 = make_uint4(0, 0, 0, 0), ra1 = ra0, rb0 = ra0, rb1 = ra0;
        int c = hw;
        if (c < NCH) {
          tmem_ld16(taddr + (uint32_t)(c * 16), va);
          load_res(c, ra0, ra1);
        }
#pragma unroll 1
        while (c < NCH) {
          tmem_ld_wait();
          if (c + 2 < NCH) {
            tmem_ld16(taddr + (uint32_t)((c + 2) * 16), vb);
            load_res(c + 2, rb0, rb1);
          }
          process(va, ra0, ra1, c);
          c += 2;
          if (c >= NCH) break;
          tmem_ld_wait();
          if (c + 2 < NCH) {
            tmem_ld16(taddr + (uint32_t)((c + 2) * 16), va);
            load_res(c + 2, ra0, ra1);
          }
          process(vb, rb0, rb1, c);
          c += 2;
        }
      } else {
        // EPI_XCORR: tile = (example b, shift row u, 128 shift columns); column n = rotation r.
        const int vt = t.mt % p.xc_vt;
        const int u = (t.mt / p.xc_vt) % p.xc_U;
        const int b = t.mt / (p.xc_vt * p.xc_U);
        const int vv = vt * 128 + row_in_tile;
        const bool row_ok = vv < p.xc_U;
        float* outp = static_cast<float*>(p.out);
#pragma unroll 1
        for (int c16 = hw; c16 < BN / 16; c16 += 2) {
          uint32_t v[16];
          tmem_ld16(taddr + (uint32_t)(c16 * 16), v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int r = n0 + c16 * 16 + j;
            if (row_ok && r < p.xc_R) {
              const long long o = (((long long)b * p.xc_R + r) * p.xc_U + u) * p.xc_U + vv;
              float s = __uint_as_float(v[j]);
              if (p.xc_cnt != nullptr && !(p.xc_cnt[o] > p.xc_thr)) s = -INFINITY;
              if (p.xc_den != nullptr) s = s / p.xc_den[b * p.xc_R + r];
              outp[o] = s;
            }
          }
        }
      }
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

}  // namespace snapb200
