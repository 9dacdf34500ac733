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
  // implicit root conv (a4d): A is a 4D sliding-window map over the packed image; m-tile mt = (image row block
  // n*Ho + ho, block of 128 output columns), segment = kernel row kh, one 32-element k-block per segment
  int a4d, r4_wtiles, r4_Ho, r4_stride;
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
  int idx32;  // set by the launcher: M, the remap extents and gn_rows_per_img fit 32-bit arithmetic
  // fused GroupNorm statistics of the stored output (see snapb200.h)
  double* gn_acc;
  double* gn_acc_relu;
  long long gn_rows_per_img;
  int gn_cpg;  // channels per group = N / 32 (a power of two)
  int gn_cpg_log;
  int gn_replica_stride;  // doubles between the GN_REPLICAS copies of the accumulator (contention spreading)
  // staged output: the epilogue writes the bf16 tile into (128B-swizzled) shared memory and TMA stores it as whole
  // 128-byte lines (dense row-major outputs without remap; BN = 64 or 128)
  int stage_out;
  int res_tma;  // with stage_out: the residual tile is TMA-loaded into the staging slab and added in place
  // A_GN mode: the A operand is produced in-kernel from the RAW activation tensor [n_img, H, W, C]:
  // GroupNorm (statistics from `g_acc`) + affine + ReLU with the reference's bf16 rounding chain, implicit
  // 1x1 / 3x3 (pad 1) window with stride 1 or 2, zero outside the image.  m = (img*Ho + ho)*Wo + wo.
  const __nv_bfloat16* g_raw;
  const double* g_acc;
  const float* g_scale;
  const float* g_bias;
  int g_rep_stride, g_nimg, g_C, g_H, g_W, g_Ho, g_Wo, g_stride, g_taps, g_pre_relu, g_post_relu;
  int g_cpg_log;   // A_TGN1: log2(C / 32)
  int g_tab_imgs;  // A_TGN1: images a CTA's contiguous tile range can touch (rows of its statistics table)
  // EPI_XCORR
  const float* xc_cnt;
  const float* xc_den;
  float xc_thr;
};

constexpr int GEMM_EPI_WARPS = 8;                       // two per TMEM lane quadrant
constexpr int GEMM_THREADS = (2 + GEMM_EPI_WARPS) * 32;  // 320
constexpr int GEMM_GN_WARPS = 4;                         // A_GN mode: producer warps (one thread per tile row)
constexpr int GEMM_THREADS_GN = GEMM_THREADS + GEMM_GN_WARPS * 32;  // 448
constexpr int GEMM_TGN_WARPS = 8;                        // A_TGN mode: in-smem transformer warps (2 threads per row)
constexpr int GEMM_THREADS_TGN = GEMM_THREADS + GEMM_TGN_WARPS * 32;  // 576
constexpr int GEMM_THREADS_T1 = GEMM_THREADS;  // A_TGN1 mode: the eight epilogue warps also transform the A tiles
enum { AMODE_TMA = 0, AMODE_GN = 1, AMODE_TGN = 2, AMODE_TGN1 = 3 };
constexpr int GEMM_GN_MAX_IMG = 32;
constexpr int GEMM_GN_TABLE_BYTES = GEMM_GN_MAX_IMG * 32 * 8 + 2048 * 4;  // statistics + packed scale/bias

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
  static constexpr int NH = (BN + 63) / 64;            // 64-column boxes per output row (the last one may be partial)
  static constexpr int OUT_STAGE_BYTES = 4 * NH * 4096;  // staged-output buffer: per TMEM quadrant NH slabs of 32 rows x 128 B
  static constexpr int smem_bytes(int stages, bool gn_tables = false, bool stage_out = false) {
    return stages * STAGE_BYTES + (stage_out ? OUT_STAGE_BYTES : 0) + 1024 + 512 +
           (gn_tables ? GEMM_GN_TABLE_BYTES : 0);
  }
  // A_TGN1: staged output + tables sized for the images one CTA touches and the C input channels
  static constexpr int smem_bytes_t1(int stages, int tab_imgs, int C) {
    return stages * STAGE_BYTES + OUT_STAGE_BYTES + 1024 + 512 + tab_imgs * 256 + C * 4;
  }
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
// Rows of a warp normally belong to one image: the NV = 2 * (groups in the chunk) partial moments of the 32 rows
// are reduced with a recursive-halving exchange (NV-1 + (5 - log2 NV) shuffles instead of 5 * NV): after it, lane L
// holds the warp total of value index L >> (5 - log2 NV), and those lanes issue ONE coalesced double atomic
// instruction.  Warps straddling an image boundary fall back to per-row atomics.
template <int NV>  // NV = 2 * (groups in the chunk) partial moments per row: (sum, sumsq) per group
__device__ __forceinline__ void gn_reduce_moments(float (&val)[NV], bool row_ok, int img, int group0, double* acc,
                                                  bool uniform, int img_ref, int lane) {
  if (!row_ok) {
#pragma unroll
    for (int i = 0; i < NV; ++i) val[i] = 0.f;
  }
  if (uniform) {
    int off = 16;
#pragma unroll
    for (int h = NV / 2; h >= 1; h >>= 1) {
      const bool up = (lane & off) != 0;
#pragma unroll
      for (int i = 0; i < h; ++i) {
        const float send = up ? val[i] : val[i + h];
        const float keep = up ? val[i + h] : val[i];
        val[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
      }
      off >>= 1;
    }
    for (; off >= 1; off >>= 1) val[0] += __shfl_xor_sync(0xffffffffu, val[0], off);
    constexpr int LOG = NV == 16 ? 4 : (NV == 8 ? 3 : (NV == 4 ? 2 : 1));
    if ((lane & ((1 << (5 - LOG)) - 1)) == 0 && img_ref >= 0)
      atomicAdd(&acc[((size_t)img_ref * 32 + group0) * 2 + (lane >> (5 - LOG))], (double)val[0]);
  } else if (row_ok) {
#pragma unroll
    for (int i = 0; i < NV; ++i) atomicAdd(&acc[((size_t)img * 32 + group0) * 2 + i], (double)val[i]);
  }
}

template <int SPAN>  // columns per group inside the 16-column chunk: 2, 4, 8 or 16
__device__ __forceinline__ void gn_accumulate16_t(const float (&v)[16], bool row_ok, int img, int group0,
                                                  double* acc, bool uniform, int img_ref, int lane) {
  constexpr int NG = 16 / SPAN;
  constexpr int NV = 2 * NG;
  float val[NV];
#pragma unroll
  for (int g = 0; g < NG; ++g) {
    float s = 0.f, q = 0.f;
#pragma unroll
    for (int j = 0; j < SPAN; ++j) {
      s += v[g * SPAN + j];
      q += v[g * SPAN + j] * v[g * SPAN + j];
    }
    val[2 * g] = s;
    val[2 * g + 1] = q;
  }
  gn_reduce_moments<NV>(val, row_ok, img, group0, acc, uniform, img_ref, lane);
}

// Packed form: the 16 stored bf16 values of the chunk as 8 words.  Sums and sums of squares come straight from the
// packed halves with the mixed-precision adds / FMAs of sm_100 (FHADD.BF16 / FHFMA.BF16): 2 instructions per value
// instead of unpack + FADD + FFMA; same summation order as the fp32 form.
template <int SPAN>
__device__ __forceinline__ void gn_accumulate16p_t(const uint32_t (&pk)[8], bool row_ok, int img, int group0,
                                                   double* acc, bool uniform, int img_ref, int lane) {
  constexpr int NG = 16 / SPAN;
  float val2[2 * NG];
#pragma unroll
  for (int g = 0; g < NG; ++g) {
    float s = 0.f, q = 0.f;
#pragma unroll
    for (int j = 0; j < SPAN / 2; ++j) {
      const uint32_t w = pk[g * (SPAN / 2) + j];
      s = bf16_lo_add(w, s);
      q = bf16_lo_sqacc(w, q);
      s = bf16_hi_add(w, s);
      q = bf16_hi_sqacc(w, q);
    }
    val2[2 * g] = s;
    val2[2 * g + 1] = q;
  }
  gn_reduce_moments<2 * NG>(val2, row_ok, img, group0, acc, uniform, img_ref, lane);
}
__device__ __forceinline__ void gn_accumulate16p(const uint32_t (&pk)[8], bool row_ok, int img, int col, int cpg_log,
                                                 double* acc, bool uniform, int img_ref, int lane) {
  const int group0 = col >> cpg_log;
  switch (cpg_log) {
    case 1: gn_accumulate16p_t<2>(pk, row_ok, img, group0, acc, uniform, img_ref, lane); break;
    case 2: gn_accumulate16p_t<4>(pk, row_ok, img, group0, acc, uniform, img_ref, lane); break;
    case 3: gn_accumulate16p_t<8>(pk, row_ok, img, group0, acc, uniform, img_ref, lane); break;
    default: gn_accumulate16p_t<16>(pk, row_ok, img, group0, acc, uniform, img_ref, lane); break;
  }
}

// cpg_log = log2(channels per group); group of column col = col >> cpg_log
__device__ __forceinline__ void gn_accumulate16(const float (&v)[16], bool row_ok, int img, int col,
                                                int cpg_log, double* acc, bool uniform, int img_ref,
                                                int lane) {
  const int group0 = col >> cpg_log;
  switch (cpg_log) {
    case 1: gn_accumulate16_t<2>(v, row_ok, img, group0, acc, uniform, img_ref, lane); break;
    case 2: gn_accumulate16_t<4>(v, row_ok, img, group0, acc, uniform, img_ref, lane); break;
    case 3: gn_accumulate16_t<8>(v, row_ok, img, group0, acc, uniform, img_ref, lane); break;
    default: gn_accumulate16_t<16>(v, row_ok, img, group0, acc, uniform, img_ref, lane); break;
  }
}

// AMODE_TMA: A tiles come straight from TMA.  AMODE_GN: producer warps build the A tile from the raw tensor through
// registers (any stride).  AMODE_TGN: TMA loads the RAW tile (dense layout, row-shifted per 3x3 tap) and
// transformer warps apply GroupNorm + affine + ReLU IN PLACE in shared memory (stride 1 only).
// AMODE_TGN1: the 1x1-conv form of A_TGN built for two CTAs per SM (320 threads, like the plain TMA mode): the eight
// epilogue warps are "workers" that first normalise the raw K blocks of tile t in shared memory (16 rows per warp,
// a thread owns one 16-byte channel chunk of 4 rows: scale / bias / statistics stay in registers for the K block) and
// then drain the accumulator of tile t - 1 (conv epilogue with staged TMA stores / TMA residual), so the MMA of tile t
// runs under that epilogue.  Statistics tables cover only the images the CTA's contiguous tile range touches.
// CONVEPI: epilogue specialised for plain conv outputs (bf16, staged TMA store, optional TMA residual, optional
// GroupNorm statistics; no bias / activation / mask / remap): the generic epilogue is compiled out.
template <int BN, int BK, int AMODE = AMODE_TMA, bool CONVEPI = false>
__global__ void __launch_bounds__(AMODE == AMODE_GN    ? GEMM_THREADS_GN
                                  : AMODE == AMODE_TGN  ? GEMM_THREADS_TGN
                                  : AMODE == AMODE_TGN1 ? GEMM_THREADS_T1
                                                        : GEMM_THREADS,
                                  (AMODE == AMODE_TMA || AMODE == AMODE_TGN1) ? 2 : 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmR,
               const GemmParams p) {
  using Cfg = GemmCfg<BN, BK>;
  constexpr bool AGN = AMODE == AMODE_GN;
  constexpr bool TGN = AMODE == AMODE_TGN;
  constexpr bool T1 = AMODE == AMODE_TGN1;
  const int STAGES = p.stages;
  // no static shared memory in this kernel: the dynamic window starts 1024-byte aligned (checked), so every
  // buffer below is a constant offset from the window base instead of a runtime-aligned pointer
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* out_stage = smem + STAGES * Cfg::STAGE_BYTES;  // 1024-aligned (stage tiles are multiples of 1024 B)
  const int out_stage_bytes = p.stage_out ? Cfg::OUT_STAGE_BYTES : 0;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(out_stage + out_stage_bytes);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  uint64_t* ready_bar = tmem_empty + 4;  // [16] A_TGN: tile transformed, MMA may consume it
  uint64_t* res_bar = ready_bar + 16;    // [4] staged residual slab of TMEM quadrant q has landed
  // A_GN / A_TGN tables (after the 512 B barrier block): per (image, group) (mean, rstd) and per channel pair packed
  // bf16x2 scale / bias
  float2* g_stat = reinterpret_cast<float2*>(out_stage + out_stage_bytes + 512);
  uint32_t* g_sc2 = reinterpret_cast<uint32_t*>(g_stat + (T1 ? p.g_tab_imgs : GEMM_GN_MAX_IMG) * 32);
  uint32_t* g_bi2 = g_sc2 + (T1 ? p.g_C / 2 : 1024);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr int w_tma = 0, w_mma = 1;
  const int total_tiles = p.m_tiles * p.n_tiles;
  // contiguous tile range per CTA (same image / same A rows stay together: L2 locality, fewer GN flushes)
  const int tile_begin = (int)((long long)blockIdx.x * total_tiles / gridDim.x);
  const int tile_end = (int)((long long)(blockIdx.x + 1) * total_tiles / gridDim.x);

  if (warp == w_tma && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (p.stage_out) tma_prefetch_desc(&tmO);
    if (p.res_tma) tma_prefetch_desc(&tmR);
    for (int s = 0; s < 4; ++s) mbar_init(&res_bar[s], 1);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], AGN ? 1 + GEMM_GN_WARPS * 32 : 1);
      mbar_init(&empty_bar[s], 1);
      if (TGN) mbar_init(&ready_bar[s], GEMM_TGN_WARPS * 32);
      if (T1) mbar_init(&ready_bar[s], GEMM_EPI_WARPS);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], GEMM_EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == w_mma) tmem_alloc(tmem_ptr_smem, Cfg::TMEM_COLS);
  if ((AGN || TGN) && warp >= 2 + GEMM_EPI_WARPS) {
    const int t = threadIdx.x - (2 + GEMM_EPI_WARPS) * 32;
    constexpr int NT = (TGN ? GEMM_TGN_WARPS : GEMM_GN_WARPS) * 32;
    const int cpg = p.g_C / 32;
    const double count = (double)p.g_H * (double)p.g_W * (double)cpg;
    for (int e = t; e < p.g_nimg * 32; e += NT) {
      double su = 0.0, sq = 0.0;
#pragma unroll
      for (int rep = 0; rep < GN_REPLICAS; ++rep) {
        su += p.g_acc[(size_t)rep * p.g_rep_stride + (size_t)e * 2];
        sq += p.g_acc[(size_t)rep * p.g_rep_stride + (size_t)e * 2 + 1];
      }
      const double mu = su / count;
      double var = sq / count - mu * mu;
      if (var < 0.0) var = 0.0;
      g_stat[e] = make_float2((float)mu, (float)(1.0 / sqrt(var + 1e-5)));
    }
    for (int c2 = t; c2 < p.g_C / 2; c2 += NT) {
      g_sc2[c2] = pack_bf16(__ldg(p.g_scale + 2 * c2), __ldg(p.g_scale + 2 * c2 + 1));
      g_bi2[c2] = pack_bf16(__ldg(p.g_bias + 2 * c2), __ldg(p.g_bias + 2 * c2 + 1));
    }
  }
  // A_TGN1: first image of this CTA's statistics table (its tiles are contiguous in m)
  const int t1_img0 = T1 ? (int)(((long long)(tile_begin / p.n_tiles) * 128) / ((long long)p.g_Ho * p.g_Wo)) : 0;
  if (T1) {
    // every thread helps: (mean, rstd) of the images this CTA touches, packed scale / bias of all channels
    const int cpg = p.g_C / 32;
    const double count = (double)p.g_H * (double)p.g_W * (double)cpg;
    const int n_tab = min(p.g_tab_imgs, p.g_nimg - t1_img0);
    for (int e = threadIdx.x; e < n_tab * 32; e += GEMM_THREADS_T1) {
      const size_t src = ((size_t)t1_img0 * 32 + e) * 2;
      double su = 0.0, sq = 0.0;
#pragma unroll
      for (int rep = 0; rep < GN_REPLICAS; ++rep) {
        su += p.g_acc[(size_t)rep * p.g_rep_stride + src];
        sq += p.g_acc[(size_t)rep * p.g_rep_stride + src + 1];
      }
      const double mu = su / count;
      double var = sq / count - mu * mu;
      if (var < 0.0) var = 0.0;
      g_stat[e] = make_float2((float)mu, (float)(1.0 / sqrt(var + 1e-5)));
    }
    for (int c2 = threadIdx.x; c2 < p.g_C / 2; c2 += GEMM_THREADS_T1) {
      g_sc2[c2] = pack_bf16(__ldg(p.g_scale + 2 * c2), __ldg(p.g_scale + 2 * c2 + 1));
      g_bi2[c2] = pack_bf16(__ldg(p.g_bias + 2 * c2), __ldg(p.g_bias + 2 * c2 + 1));
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == w_tma) {
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
          mbar_arrive_expect_tx(&full_bar[stage], AGN ? Cfg::B_BYTES : Cfg::STAGE_BYTES);
          uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
          uint8_t* sb = sa + Cfg::A_BYTES;
          if (!AGN) {
            if (p.a4d) {
              const int rb = t.mt / p.r4_wtiles, wb = t.mt - rb * p.r4_wtiles;
              const int n = rb / p.r4_Ho, ho = rb - n * p.r4_Ho;
              tma_load_4d(&tmA, &full_bar[stage], sa, 0, wb * 128, ho * p.r4_stride + seg, n);
            } else {
              tma_load_2d(&tmA, &full_bar[stage], sa, p.a_col0 + kc * BK, t.a_row + a_off);
            }
          }
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
  } else if (warp == w_mma) {
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
          mbar_wait((TGN || T1) ? &ready_bar[stage] : &full_bar[stage], phase);
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
  } else if (TGN && warp >= 2 + GEMM_EPI_WARPS) {
    // ===================== A_TGN transformer warps: GroupNorm + ReLU in place on the raw smem tile =====
    if constexpr (TGN && BK == 64) {
      const int tt = threadIdx.x - (2 + GEMM_EPI_WARPS) * 32;  // 0..255
      const int t = tt & 127;                                   // tile row
      const int qh = tt >> 7;                                    // which four 16-byte chunks of the row
      const int C = p.g_C, cpg = C / 32;
      const long long per_img = (long long)p.g_Ho * p.g_Wo;
      const __nv_bfloat162 zero2 = __floats2bfloat162_rn(0.f, 0.f);
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = tile_begin; tile < tile_end; ++tile) {
        const int mt = tile / p.n_tiles;
        const long long m = (long long)mt * 128 + t;
        const bool row_ok = m < p.M_valid;
        const int img = row_ok ? (int)(m / per_img) : 0;
        const int rem = row_ok ? (int)(m - (long long)img * per_img) : 0;
        const int ho = rem / p.g_Wo, wo = rem - ho * p.g_Wo;
        const float2* stat = g_stat + img * 32;
        int seg = 0, kc = 0;
        for (int kb = 0; kb < p.nkb; ++kb) {
          bool inb = row_ok;
          if (p.g_taps == 9) {  // stride 1, pad 1: the row-shifted dense tile wraps at image borders -> zero there
            const int kh = seg / 3, kw = seg - 3 * kh;
            const int hi = ho + kh - 1, wi = wo + kw - 1;
            inb = inb && hi >= 0 && hi < p.g_H && wi >= 0 && wi < p.g_W;
          }
          mbar_wait(&full_bar[stage], phase);  // raw tile (and the weights) have landed
          uint8_t* rowp = smem + stage * Cfg::STAGE_BYTES + t * 128;
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) {
            const int c = qh * 4 + cc;
            uint4* slot = reinterpret_cast<uint4*>(rowp + ((c ^ (t & 7)) * 16));
            uint4 o = make_uint4(0, 0, 0, 0);
            if (inb) {
              const uint4 u = *slot;
              const int ch0 = kc * 64 + c * 8;
              const uint32_t w[4] = {u.x, u.y, u.z, u.w};
              uint32_t r[4];
#pragma unroll
              for (int jj = 0; jj < 4; ++jj) {
                const int ch = ch0 + 2 * jj;
                const float2 st = stat[ch / cpg];
                float2 x = unpack_bf16(w[jj]);
                if (p.g_pre_relu) {
                  x.x = fmaxf(x.x, 0.f);
                  x.y = fmaxf(x.y, 0.f);
                }
                __nv_bfloat162 v = __floats2bfloat162_rn((x.x - st.x) * st.y, (x.y - st.x) * st.y);
                const uint32_t s2 = g_sc2[ch >> 1], b2 = g_bi2[ch >> 1];
                v = __hmul2_rn(v, *reinterpret_cast<const __nv_bfloat162*>(&s2));
                v = __hadd2_rn(v, *reinterpret_cast<const __nv_bfloat162*>(&b2));
                if (p.g_post_relu) v = __hmax2(v, zero2);
                r[jj] = *reinterpret_cast<uint32_t*>(&v);
              }
              o = make_uint4(r[0], r[1], r[2], r[3]);
            }
            *slot = o;
          }
          fence_proxy_async_smem();
          mbar_arrive(&ready_bar[stage]);
          if (++kc == p.kps) {
            kc = 0;
            ++seg;
          }
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (AGN && warp >= 2 + GEMM_EPI_WARPS) {
    // ===================== A_GN producer warps: thread t owns tile row t =====================
    if constexpr (AGN && BK == 64) {
      const int t = threadIdx.x - (2 + GEMM_EPI_WARPS) * 32;
      const int C = p.g_C, cpg = C / 32;
      const int pad = p.g_taps == 9 ? 1 : 0;
      const long long per_img = (long long)p.g_Ho * p.g_Wo;
      const __nv_bfloat162 zero2 = __floats2bfloat162_rn(0.f, 0.f);
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = tile_begin; tile < tile_end; ++tile) {
        const int mt = tile / p.n_tiles;
        const long long m = (long long)mt * 128 + t;
        const bool row_ok = m < p.M_valid;
        const int img = row_ok ? (int)(m / per_img) : 0;
        const int rem = row_ok ? (int)(m - (long long)img * per_img) : 0;
        const int ho = rem / p.g_Wo, wo = rem - ho * p.g_Wo;
        const float2* stat = g_stat + img * 32;
        int seg = 0, kc = 0;
        for (int kb = 0; kb < p.nkb; ++kb) {
          const int kh = p.g_taps == 9 ? seg / 3 : 0, kw = p.g_taps == 9 ? seg - 3 * (seg / 3) : 0;
          const int hi = ho * p.g_stride + kh - pad, wi = wo * p.g_stride + kw - pad;
          const bool inb = row_ok && hi >= 0 && hi < p.g_H && wi >= 0 && wi < p.g_W;
          uint4 u[8];
          if (inb) {
            const uint4* src = reinterpret_cast<const uint4*>(
                p.g_raw + (((size_t)img * p.g_H + hi) * p.g_W + wi) * C + kc * 64);
#pragma unroll
            for (int c = 0; c < 8; ++c) u[c] = __ldg(src + c);
          }
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* rowp = smem + stage * Cfg::STAGE_BYTES + t * 128;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            uint4 o = make_uint4(0, 0, 0, 0);
            if (inb) {
              const int ch0 = kc * 64 + c * 8;
              const uint32_t w[4] = {u[c].x, u[c].y, u[c].z, u[c].w};
              uint32_t r[4];
#pragma unroll
              for (int jj = 0; jj < 4; ++jj) {
                const int ch = ch0 + 2 * jj;
                const float2 st = stat[ch / cpg];
                float2 x = unpack_bf16(w[jj]);
                if (p.g_pre_relu) {
                  x.x = fmaxf(x.x, 0.f);
                  x.y = fmaxf(x.y, 0.f);
                }
                // resnet.py:39-41,57-69 in bf16: standardise (fp32) -> bf16, * scale -> bf16, + bias -> bf16
                __nv_bfloat162 v = __floats2bfloat162_rn((x.x - st.x) * st.y, (x.y - st.x) * st.y);
                const uint32_t s2 = g_sc2[ch >> 1], b2 = g_bi2[ch >> 1];
                v = __hmul2_rn(v, *reinterpret_cast<const __nv_bfloat162*>(&s2));
                v = __hadd2_rn(v, *reinterpret_cast<const __nv_bfloat162*>(&b2));
                if (p.g_post_relu) v = __hmax2(v, zero2);
                r[jj] = *reinterpret_cast<uint32_t*>(&v);
              }
              o = make_uint4(r[0], r[1], r[2], r[3]);
            }
            *reinterpret_cast<uint4*>(rowp + ((c ^ (t & 7)) * 16)) = o;  // SWIZZLE_128B position of chunk c
          }
          fence_proxy_async_smem();
          mbar_arrive(&full_bar[stage]);
          if (++kc == p.kps) {
            kc = 0;
            ++seg;
          }
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else {
    // ===================== epilogue warps =====================
    const int q = warp & 3;         // TMEM lane quadrant this warp may access
    const int hw = (warp - 2) >> 2;  // 0/1: which half of the 16-column chunks (interleaved) it drains
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t res_phase = 0;
    uint8_t* slab = out_stage + q * Cfg::NH * 4096;  // this quadrant's 32-row staging slab
    // residual slab of a tile: BN/64 boxes of 32 rows x 64 columns (issued by one thread per quadrant)
    auto issue_residual = [&](int tile_id) {
      if constexpr (BN % 64 == 0) {
        const TileCoord tn = decode_tile(p, tile_id, BN);
        int nbox = 0;
#pragma unroll
        for (int h = 0; h < BN / 64; ++h) nbox += (tn.nt * BN + h * 64 < p.N) ? 1 : 0;
        mbar_arrive_expect_tx(&res_bar[q], (uint32_t)nbox * 4096u);
#pragma unroll
        for (int h = 0; h < BN / 64; ++h)
          if (tn.nt * BN + h * 64 < p.N)
            tma_load_2d(&tmR, &res_bar[q], slab + h * 4096, tn.nt * BN + h * 64, tn.mt * 128 + q * 32);
      }
    };
    if (p.res_tma && hw == 0 && lane == 0 && tile_begin < tile_end) issue_residual(tile_begin);
    if constexpr (CONVEPI && BN % 64 == 0) {
      // ---------------- conv epilogue: warp (q, hw) drains the contiguous column half hw of TMEM quadrant q ----------
      constexpr int NCW = BN / 32;  // 16-column chunks per warp
      const int cbase = hw * NCW;
      uint8_t* rowp = slab + ((cbase * 16) / 64) * 4096 + lane * 128;  // this lane's row inside its 64-column half
      // 16-byte slot of chunk i: (2 * ((cbase & 3) + i)) ^ (row % 8); the chunk part never carries into the rest
      const uint32_t swz16 = (uint32_t)(((lane & 7) ^ ((cbase & 3) * 2)) << 4);
      double* gacc = p.gn_acc != nullptr ? p.gn_acc + (size_t)(blockIdx.x % GN_REPLICAS) * p.gn_replica_stride : nullptr;
      double* gacc_relu =
          p.gn_acc_relu != nullptr ? p.gn_acc_relu + (size_t)(blockIdx.x % GN_REPLICAS) * p.gn_replica_stride : nullptr;
      const int m_valid = (int)p.M_valid;
      const int rpi = (int)p.gn_rows_per_img;
      auto epi_tile = [&](int tile) {
        const TileCoord t = decode_tile(p, tile, BN);
        mbar_wait(&tmem_full[acc], acc_phase);
        tc_fence_after_sync();
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + cbase * 16);
        const int m0 = t.mt * 128 + q * 32;
        const bool row_ok = m0 + lane < m_valid;
        const int col0 = t.nt * BN + cbase * 16;
        int gn_img = -1, gn_ref = -1;
        bool gn_uniform = true;
        if (gacc != nullptr && m0 < m_valid) {
          gn_ref = m0 / rpi;
          const int last = min(m0 + 31, m_valid - 1) / rpi;
          gn_uniform = gn_ref == last;
          if (!gn_uniform) gn_img = row_ok ? (m0 + lane) / rpi : -1;
        }
        if (p.res_tma) {  // residual slab landed (which also implies the previous store released the slab)
          mbar_wait(&res_bar[q], res_phase);
          res_phase ^= 1;
        } else {          // the slab is free once the previous tile's TMA store has read it
          if (hw == 0 && lane == 0) bulk_wait_group_read0();
          pair_bar_sync(q);
        }
        uint32_t v[2][16];
        tmem_ld16(taddr, v[0]);
#pragma unroll
        for (int i = 0; i < NCW; ++i) {
          tmem_ld_wait();
          if (i + 1 < NCW) tmem_ld16(taddr + (uint32_t)((i + 1) * 16), v[(i + 1) & 1]);
          const uint32_t(&vv)[16] = v[i & 1];
          // chunk i of this warp = 16-column chunk cbase + i of the tile: 64-column half (cbase + i) / 4, 16-byte slots
          // 2 ((cbase + i) % 4) and + 1 of this lane's 128-byte row, XOR-swizzled by row % 8 (i is a compile-time constant)
          uint8_t* rp = rowp + ((((cbase & 3) + i) >> 2) * 4096);
          uint8_t* p0 = rp + (((uint32_t)(2 * (i & 3)) << 4) ^ swz16);
          uint8_t* p1 = rp + (((uint32_t)(2 * (i & 3) + 1) << 4) ^ swz16);
          uint32_t pk[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) pk[j] = pack_bf16(__uint_as_float(vv[2 * j]), __uint_as_float(vv[2 * j + 1]));
          if (p.res_tma) {
            const uint4 s0 = *reinterpret_cast<const uint4*>(p0);
            const uint4 s1 = *reinterpret_cast<const uint4*>(p1);
            const uint32_t rr[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
#pragma unroll
            for (int j = 0; j < 8; ++j) pk[j] = hadd2_bf16_rn(pk[j], rr[j]);
          }
          *reinterpret_cast<uint4*>(p0) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          *reinterpret_cast<uint4*>(p1) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
          if (gacc != nullptr) {  // statistics of the stored (rounded) values, straight from the packed words
            gn_accumulate16p(pk, row_ok, gn_img, col0 + i * 16, p.gn_cpg_log, gacc, gn_uniform, gn_ref, lane);
            if (gacc_relu != nullptr) {
#pragma unroll
              for (int j = 0; j < 8; ++j) pk[j] = hmax2_bf16(pk[j], 0u);
              gn_accumulate16p(pk, row_ok, gn_img, col0 + i * 16, p.gn_cpg_log, gacc_relu, gn_uniform, gn_ref, lane);
            }
          }
        }
        fence_proxy_async_smem();
        pair_bar_sync(q);
        if (hw == 0 && lane == 0) {
          if (p.a4d) {   // implicit root conv: tile = (image row block, block of 128 output columns); a 3D map clips at Wo
            const int rb = t.mt / p.r4_wtiles, wb = t.mt - rb * p.r4_wtiles;
#pragma unroll
            for (int h = 0; h < BN / 64; ++h) tma_store_3d(&tmO, slab + h * 4096, t.nt * BN + h * 64, wb * 128 + q * 32, rb);
          } else {
#pragma unroll
            for (int h = 0; h < BN / 64; ++h)
              tma_store_2d(&tmO, slab + h * 4096, t.nt * BN + h * 64, t.mt * 128 + q * 32);
          }
          bulk_commit_group();
          if (p.res_tma && tile + 1 < tile_end) {  // prefetch the next tile's residual behind the MMA wait
            bulk_wait_group_read0();
            issue_residual(tile + 1);
          }
        }
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      };
      if constexpr (!T1) {
        for (int tile = tile_begin; tile < tile_end; ++tile) epi_tile(tile);
      } else if constexpr (BK == 64) {
        // ---------------- A_TGN1 worker loop: normalise tile t's K blocks in place, then drain tile t - 1 ----------
        const int wk = warp - 2;                       // rows 16 wk .. 16 wk + 15 of every K block
        const int c = lane & 7;                        // 16-byte chunk (8 channels) of the 128-byte row
        const int r0 = 16 * wk + (lane >> 3);          // rows r0 + 4 i, i = 0..3: a warp touches 4 whole rows per step
        // SWIZZLE_128B: chunk c of row r sits at c ^ (r % 8); r % 8 = (lane >> 3) + 4 (i & 1)
        const uint32_t off_e = (uint32_t)(r0 * 128 + ((c ^ (lane >> 3)) << 4));
        const uint32_t off_o = (uint32_t)((r0 + 4) * 128 + ((c ^ ((lane >> 3) + 4)) << 4));
        const int cpg_log = p.g_cpg_log;  // channels per group = C / 32, a power of two (checked by the host)
        const int per_img = p.g_Ho * p.g_Wo;
        const uint32_t post_lo = p.g_post_relu ? 0u : 0xff80ff80u;  // optional ReLU = packed max against 0 or -inf
        const bool pre = p.g_pre_relu != 0;
        // resnet.py:39-41,57-69 in bf16: standardise (fp32) -> bf16, * scale -> bf16, + bias -> bf16 (-> ReLU).
        // (x - mean) comes straight from the packed halves (FHADD.BF16): 8 instructions per channel pair.
        auto xform = [&](const uint4& u, const float (&mean)[4], const float (&rstd)[4], const uint4& s4,
                         const uint4& b4, bool pre_relu) -> uint4 {
          const uint32_t w[4] = {u.x, u.y, u.z, u.w};
          const uint32_t sc[4] = {s4.x, s4.y, s4.z, s4.w};
          const uint32_t bi[4] = {b4.x, b4.y, b4.z, b4.w};
          uint32_t r[4];
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            const uint32_t ww = pre_relu ? hmax2_bf16(w[jj], 0u) : w[jj];
            __nv_bfloat162 v = __floats2bfloat162_rn(bf16_lo_sub(ww, mean[jj]) * rstd[jj],
                                                     bf16_hi_sub(ww, mean[jj]) * rstd[jj]);
            v = __hmul2_rn(v, *reinterpret_cast<const __nv_bfloat162*>(&sc[jj]));
            v = __hadd2_rn(v, *reinterpret_cast<const __nv_bfloat162*>(&bi[jj]));
            r[jj] = hmax2_bf16(*reinterpret_cast<uint32_t*>(&v), post_lo);
          }
          return make_uint4(r[0], r[1], r[2], r[3]);
        };
        int stage = 0;
        uint32_t phase = 0;
        for (int tile = tile_begin; tile < tile_end; ++tile) {
          const int mt0 = (tile / p.n_tiles) * 128;
          const int img_a = mt0 / per_img - t1_img0;
          const bool uniform = (min(mt0 + 127, m_valid - 1) / per_img - t1_img0) == img_a;
#pragma unroll 1
          for (int kb = 0; kb < p.nkb; ++kb) {
            // per K block this thread's 8 channels are fixed: scale / bias / group statistics stay in registers
            const int ch0 = kb * 64 + c * 8;
            const uint4 s4 = *reinterpret_cast<const uint4*>(g_sc2 + (ch0 >> 1));
            const uint4 b4 = *reinterpret_cast<const uint4*>(g_bi2 + (ch0 >> 1));
            const int g0 = ch0 >> cpg_log;
            float mean[4], rstd[4];
            auto load_stats = [&](int img) {
              if (cpg_log >= 3) {  // the chunk's 8 channels share one group
                const float2 st = g_stat[img * 32 + g0];
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                  mean[jj] = st.x;
                  rstd[jj] = st.y;
                }
              } else {
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                  const float2 st = g_stat[img * 32 + g0 + ((2 * jj) >> cpg_log)];  // group of channel ch0 + 2 jj
                  mean[jj] = st.x;
                  rstd[jj] = st.y;
                }
              }
            };
            load_stats(img_a);
            mbar_wait(&full_bar[stage], phase);  // raw tile (and the weights) have landed
            uint8_t* base = smem + stage * Cfg::STAGE_BYTES;
            uint4* slot[4] = {reinterpret_cast<uint4*>(base + off_e), reinterpret_cast<uint4*>(base + off_o),
                              reinterpret_cast<uint4*>(base + off_e + 1024), reinterpret_cast<uint4*>(base + off_o + 1024)};
            if (uniform) {
              uint4 u[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) u[j] = *slot[j];
              if (pre) {
#pragma unroll
                for (int j = 0; j < 4; ++j) u[j] = xform(u[j], mean, rstd, s4, b4, true);
              } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) u[j] = xform(u[j], mean, rstd, s4, b4, false);
              }
#pragma unroll
              for (int j = 0; j < 4; ++j) *slot[j] = u[j];
            } else {  // the tile straddles images (or images smaller than a tile): statistics per row
#pragma unroll 1
              for (int i = 0; i < 4; ++i) {
                const int m = min(mt0 + r0 + 4 * i, m_valid - 1);
                load_stats(m / per_img - t1_img0);
                *slot[i] = xform(*slot[i], mean, rstd, s4, b4, pre);
              }
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&ready_bar[stage]);
            if (++stage == STAGES) {
              stage = 0;
              phase ^= 1;
            }
          }
          if (tile > tile_begin) epi_tile(tile - 1);
        }
        if (tile_begin < tile_end) epi_tile(tile_end - 1);
      }
      if (hw == 0 && lane == 0) bulk_wait_group0();
    } else
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
          if (p.idx32) {   // every index fits 32 bits (the host checked): no software 64-bit divisions per tile
            const unsigned per_img = (unsigned)(p.rm_R * p.rm_C);
            const unsigned img = (unsigned)m / per_img;
            const int rem = (int)((unsigned)m - img * per_img);
            const int r = (int)((unsigned)rem / (unsigned)p.rm_C);
            const int c = rem - r * p.rm_C;
            row_ok = row_ok && r >= p.rm_r0 && r < p.rm_r0 + p.rm_Ho && c >= p.rm_c0 && c < p.rm_c0 + p.rm_Wo;
            orow = (long long)((img * (unsigned)p.rm_Ho + (unsigned)(r - p.rm_r0)) * (unsigned)p.rm_Wo + (unsigned)(c - p.rm_c0));
          } else {
            const long long per_img = (long long)p.rm_R * p.rm_C;
            const long long img = m / per_img;
            const int rem = (int)(m - img * per_img);
            const int r = rem / p.rm_C;
            const int c = rem - r * p.rm_C;
            row_ok = row_ok && r >= p.rm_r0 && r < p.rm_r0 + p.rm_Ho && c >= p.rm_c0 &&
                     c < p.rm_c0 + p.rm_Wo;
            orow = (img * p.rm_Ho + (r - p.rm_r0)) * p.rm_Wo + (c - p.rm_c0);
          }
        }
        if (!row_ok) orow = 0;
        const bool keep = row_ok && (p.row_mask == nullptr || p.row_mask[orow] != 0);
        int gn_img = -1, gn_ref = -1;
        bool gn_uniform = true;
        if (p.gn_acc != nullptr) {
          gn_img = !row_ok ? -1 : (p.idx32 ? (int)((unsigned)orow / (unsigned)p.gn_rows_per_img) : (int)(orow / p.gn_rows_per_img));
          gn_ref = __reduce_max_sync(0xffffffffu, gn_img);
          gn_uniform = __all_sync(0xffffffffu, gn_img == gn_ref || gn_img == -1);
        }
        const bool rnd = !p.out_f32;
        // conv fast path: no bias / activation / mask (rows beyond M are clipped by the store, masked in the statistics)
        const bool fast = rnd && p.bias == nullptr && !p.relu && p.row_mask == nullptr;
        const bool has_res = p.residual != nullptr && !p.res_tma;
        const __nv_bfloat16* res_row = has_res ? p.residual + orow * p.ldr : nullptr;

        // One 16-column chunk: all roundings the reference applies (dot -> dtype, + bias -> dtype,
        // + residual -> dtype), fused GroupNorm statistics of exactly what is stored, then the store.
        auto process = [&](const uint32_t (&v)[16], const uint4& r0, const uint4& r1, int c16) {
          const int col = n0 + c16 * 16;
          if (col >= p.N) return;  // warp-uniform
          uint32_t pk[8];
          if (fast) {
            // plain conv output (+ residual): the dot product is rounded to bf16 while packing (one F2FP per pair)
            // and the residual is added as packed bf16 (HADD2.BF16 = exact sum, one rounding: identical to the
            // fp32 add + cast of the reference because both addends are bf16)
#pragma unroll
            for (int j = 0; j < 8; ++j) pk[j] = pack_bf16(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]));
            if (p.res_tma) {
              const uint8_t* rowp = slab + (c16 >> 2) * 4096 + lane * 128;
              const int j = (c16 & 3) * 2, sw = lane & 7;
              const uint4 s0 = *reinterpret_cast<const uint4*>(rowp + ((j ^ sw) << 4));
              const uint4 s1 = *reinterpret_cast<const uint4*>(rowp + (((j + 1) ^ sw) << 4));
              const uint32_t rr[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
#pragma unroll
              for (int j2 = 0; j2 < 8; ++j2) pk[j2] = hadd2_bf16_rn(pk[j2], rr[j2]);
            } else if (has_res) {
              const uint32_t rr[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
              for (int j2 = 0; j2 < 8; ++j2) pk[j2] = hadd2_bf16_rn(pk[j2], rr[j2]);
            }
          } else {
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
            if (p.res_tma) {
              const uint8_t* rowp = slab + (c16 >> 2) * 4096 + lane * 128;
              const int j = (c16 & 3) * 2, sw = lane & 7;
              const uint4 s0 = *reinterpret_cast<const uint4*>(rowp + ((j ^ sw) << 4));
              const uint4 s1 = *reinterpret_cast<const uint4*>(rowp + (((j + 1) ^ sw) << 4));
              const uint32_t rr[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
#pragma unroll
              for (int j2 = 0; j2 < 8; ++j2) {
                const float2 x = unpack_bf16(rr[j2]);
                f[2 * j2] += x.x;
                f[2 * j2 + 1] += x.y;
              }
            } else if (has_res) {
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
#pragma unroll
            for (int j = 0; j < 8; ++j) pk[j] = pack_bf16(f[2 * j], f[2 * j + 1]);
          }
          if (p.stage_out) {
            // slab of this TMEM quadrant: [BN/64 halves][32 rows][128 B], 16-byte chunks XOR-swizzled by row % 8
            // (the TMA SWIZZLE_128B pattern; also makes the row-per-lane 16 B stores bank-conflict free)
            uint8_t* rowp = slab + (c16 >> 2) * 4096 + lane * 128;
            const int j = (c16 & 3) * 2, sw = lane & 7;
            *reinterpret_cast<uint4*>(rowp + ((j ^ sw) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            *reinterpret_cast<uint4*>(rowp + (((j + 1) ^ sw) << 4)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
          } else if (row_ok) {
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
            gn_accumulate16(g, row_ok, gn_img, col, p.gn_cpg_log, p.gn_acc + rep, gn_uniform, gn_ref, lane);
            if (p.gn_acc_relu != nullptr) {
#pragma unroll
              for (int j = 0; j < 16; ++j) g[j] = fmaxf(g[j], 0.f);
              gn_accumulate16(g, row_ok, gn_img, col, p.gn_cpg_log, p.gn_acc_relu + rep, gn_uniform, gn_ref, lane);
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
        if (p.res_tma) {  // residual slab landed (which also implies the previous store released the slab)
          mbar_wait(&res_bar[q], res_phase);
          res_phase ^= 1;
        } else if (p.stage_out) {  // the slab is free once the previous tile's TMA store has read it
          if (hw == 0 && lane == 0) bulk_wait_group_read0();
          pair_bar_sync(q);
        }
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
        if (p.stage_out) {
          fence_proxy_async_smem();
          pair_bar_sync(q);
          if (hw == 0 && lane == 0) {
            if constexpr (BN % 32 == 0) {   // a partial last box (N = 160) is clipped by the tensor map's bounds
#pragma unroll
              for (int h = 0; h < Cfg::NH; ++h)
                if (n0 + h * 64 < p.N)
                  tma_store_2d(&tmO, slab + h * 4096, n0 + h * 64, t.mt * 128 + q * 32);
            }
            bulk_commit_group();
            if (p.res_tma && tile + 1 < tile_end) {  // prefetch the next tile's residual behind the MMA wait
              bulk_wait_group_read0();
              issue_residual(tile + 1);
            }
          }
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
    if (p.stage_out && hw == 0 && lane == 0) bulk_wait_group0();
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == w_mma) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

}  // namespace snapb200
