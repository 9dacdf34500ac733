// Host launcher + C ABI for the tcgen05 GEMM engine (gemm_tc.cuh).
#include <stdlib.h>
#include <string.h>

#include "gemm_launch.h"
#include "gemm_tc.cuh"

namespace snapb200 {

static int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v != nullptr && *v != 0 ? atoi(v) : dflt;
}

template <int BN, int BK, bool CONVEPI = false>
static int launch_inst(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmO,
                       const CUtensorMap& tmR, GemmParams p, int ctas_per_sm, cudaStream_t s) {
  using Cfg = GemmCfg<BN, BK>;
  static DynSmemState smem_state;  // per device (one process may drive several GPUs)
  if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(&gemm_tc_kernel<BN, BK, AMODE_TMA, CONVEPI>), 227 * 1024,
                               &smem_state, "cudaFuncSetAttribute(gemm_tc)"))
    return rc;
  // occupancy plan: `ctas_per_sm` co-resident CTAs share the SM's 227 KB of shared memory and 512 TMEM
  // columns; memory-bound layers (small K) want several CTAs so that epilogues overlap main loops
  int max_by_tmem = 512 / Cfg::TMEM_COLS;
  if (ctas_per_sm < 1) ctas_per_sm = 1;
  if (ctas_per_sm > max_by_tmem) ctas_per_sm = max_by_tmem;
  if (ctas_per_sm > 2) ctas_per_sm = 2;  // register budget: __launch_bounds__(320, 2)
  static const int env_ctas = env_int("SNAPB200_GEMM_CTAS", 0);  // tuning overrides (tools/gemm_sweep.py)
  if (env_ctas > 0) ctas_per_sm = env_ctas > 2 ? 2 : env_ctas;
  const bool so = p.stage_out != 0;
  // the ring may be deeper than one tile's K blocks: the TMA producer then runs whole tiles ahead of the MMA
  // (small-K layers: the root conv has 7 blocks of 12 KB per tile, conv3 of stage 1 a single one)
  static const int env_deep = env_int("SNAPB200_GEMM_DEEP_RING", 1);
  int stages = (env_deep || p.nkb >= Cfg::MAX_STAGES) ? Cfg::MAX_STAGES : p.nkb;
  if (stages < 2) stages = 2;
  p.idx32 = p.M_valid < (1LL << 31) && (long long)p.m_tiles * 128 < (1LL << 31) &&
            (!p.remap || (long long)p.rm_R * p.rm_C < (1LL << 31)) && p.gn_rows_per_img < (1LL << 31);
  while (stages > 2 && ctas_per_sm * Cfg::smem_bytes(stages, false, so) > 225 * 1024) --stages;
  while (ctas_per_sm > 1 && ctas_per_sm * Cfg::smem_bytes(stages, false, so) > 225 * 1024) --ctas_per_sm;
  p.stages = stages;
  const int total = p.m_tiles * p.n_tiles;
  const int cap = num_sms() * ctas_per_sm;
  const int grid = total < cap ? total : cap;
  gemm_tc_kernel<BN, BK, AMODE_TMA, CONVEPI><<<grid, GEMM_THREADS, Cfg::smem_bytes(stages, false, so), s>>>(
      tmA, tmB, tmO, tmR, p);
  return check_launch("gemm_tc_kernel");
}

int launch_gemm(const void* A, long long a_rows, int a_cols, long long a_ld, const void* B,
                long long b_rows, int b_cols, long long b_ld, int bn, int bk, const GemmParams& p_in,
                cudaStream_t s, int ctas_per_sm) {
  GemmParams p = p_in;
  SNAP_REQUIRE(p.m_tiles > 0 && p.n_tiles > 0 && p.nkb > 0, "empty GEMM");
  CUtensorMap tmA, tmB, tmO, tmR;
  memset(&tmO, 0, sizeof(tmO));
  memset(&tmR, 0, sizeof(tmR));
  // dense bf16 outputs leave through shared memory + TMA (whole 128-byte lines instead of 32 B per lane)
  static const int env_so = env_int("SNAPB200_GEMM_STAGE_OUT", 1);
  p.stage_out = env_so && p.epi == EPI_STORE && !p.remap && !p.out_f32 && bk == 64 &&
                (((bn == 64 || bn == 128) && p.N % 64 == 0) || (bn == 160 && p.N == 160 && p.residual == nullptr)) &&
                p.ldo % 8 == 0 && (reinterpret_cast<uintptr_t>(p.out) & 15) == 0;
  if (p.stage_out) {
    int rc = make_tmap_2d_bf16(&tmO, p.out, p.M_valid, p.N, p.ldo, 32, 64);
    if (rc) return rc;
    static const int env_rt = env_int("SNAPB200_GEMM_RES_TMA", 1);
    p.res_tma = env_rt && p.residual != nullptr && p.ldr % 8 == 0 &&
                (reinterpret_cast<uintptr_t>(p.residual) & 15) == 0;
    if (p.res_tma) {
      rc = make_tmap_2d_bf16(&tmR, p.residual, p.M_valid, p.N, p.ldr, 32, 64);
      if (rc) return rc;
    }
  }
  int rc = make_tmap_2d_bf16(&tmA, A, a_rows, a_cols, a_ld, 128, bk);
  if (rc) return rc;
  rc = make_tmap_2d_bf16(&tmB, B, b_rows, b_cols, b_ld, bn, bk);
  if (rc) return rc;
  // plain conv outputs take the specialised epilogue
  const bool conv_epi = p.stage_out && p.bias == nullptr && !p.relu && p.row_mask == nullptr && p.N % bn == 0 &&
                        (p.residual == nullptr || p.res_tma) && p.M_valid < (1LL << 31) &&
                        (p.gn_acc == nullptr || p.gn_rows_per_img < (1LL << 31));
  if (conv_epi && bk == 64 && bn == 64) return launch_inst<64, 64, true>(tmA, tmB, tmO, tmR, p, ctas_per_sm, s);
  if (conv_epi && bk == 64 && bn == 128) return launch_inst<128, 64, true>(tmA, tmB, tmO, tmR, p, ctas_per_sm, s);
#define SNAP_GEMM_CASE(BN_, BK_) \
  if (bn == BN_ && bk == BK_) return launch_inst<BN_, BK_>(tmA, tmB, tmO, tmR, p, ctas_per_sm, s);
  SNAP_GEMM_CASE(16, 64)
  SNAP_GEMM_CASE(64, 64)
  SNAP_GEMM_CASE(128, 64)
  SNAP_GEMM_CASE(160, 64)
  SNAP_GEMM_CASE(256, 64)
  SNAP_GEMM_CASE(48, 32)
  SNAP_GEMM_CASE(64, 32)
  SNAP_GEMM_CASE(128, 32)
  SNAP_GEMM_CASE(256, 32)
#undef SNAP_GEMM_CASE
  return set_error(SNAPB200_ERR_INVALID, "no GEMM instance for bn=%d bk=%d", bn, bk);
}

int pick_bn(int n, int bk) {
  if (n <= 16 && bk == 64) return 16;
  if (n == 160 && bk == 64) return 160;
  if (n <= 64) return 64;
  if (n <= 128) return 128;
  return 256;
}

// Tile-shape / occupancy heuristic for the conv & dense layers (M = pixels is large, N = channels):
// K >= 1024 and N >= 256 are tensor-bound -> 128x256 tiles, one CTA per SM, deep pipeline;
// everything else is bound by streaming A / the output -> 128x128 (or 128x64) tiles, 2-3 CTAs per SM.
void pick_config(int n, int k_total, int bk, int* bn, int* ctas_per_sm) {
  if (bk == 64 && n == 160) { *bn = 160; *ctas_per_sm = 1; return; }
  if (n <= 16 && bk == 64) { *bn = 16; *ctas_per_sm = 2; return; }
  if (n <= 64) { *bn = 64; *ctas_per_sm = 3; return; }
  if (k_total >= 1024 && n >= 256) { *bn = 256; *ctas_per_sm = 1; return; }
  *bn = 128;
  *ctas_per_sm = 2;
}

}  // namespace snapb200

namespace snapb200 {
template <int BN, int AMODE>
static int launch_gn_inst(const CUtensorMap& tmA, const CUtensorMap& tmB, GemmParams p, cudaStream_t s) {
  using Cfg = GemmCfg<BN, 64>;
  static DynSmemState smem_state;
  {
    const int want = Cfg::smem_bytes(Cfg::MAX_STAGES, true);
    if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(&gemm_tc_kernel<BN, 64, AMODE>),
                                 want > 227 * 1024 ? 227 * 1024 : want, &smem_state,
                                 "cudaFuncSetAttribute(gemm_tc<gn>)"))
      return rc;
  }
  int stages = p.nkb < Cfg::MAX_STAGES ? p.nkb : Cfg::MAX_STAGES;
  if (stages < 2) stages = 2;
  while (stages > 2 && Cfg::smem_bytes(stages, true) > 225 * 1024) --stages;
  p.stages = stages;
  const int total = p.m_tiles * p.n_tiles;
  const int grid = total < num_sms() ? total : num_sms();
  const int threads = AMODE == AMODE_TGN ? GEMM_THREADS_TGN : GEMM_THREADS_GN;
  p.stage_out = 0;
  p.res_tma = 0;
  gemm_tc_kernel<BN, 64, AMODE><<<grid, threads, Cfg::smem_bytes(stages, true), s>>>(tmA, tmB, tmA, tmA, p);
  return check_launch("gemm_tc_kernel<gn>");
}
}  // namespace snapb200

namespace snapb200 {
// A_TGN1 launcher (1x1 convs, stride 1): TMA loads the raw tile, two transformer warps normalise it in shared memory,
// conv epilogue with staged TMA stores.  Two CTAs per SM when two pipeline stages fit beside the staging slab.
template <int BN>
static int launch_t1_inst(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmO,
                          const CUtensorMap& tmR, GemmParams p, cudaStream_t s) {
  using Cfg = GemmCfg<BN, 64>;
  static DynSmemState smem_state;
  if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(&gemm_tc_kernel<BN, 64, AMODE_TGN1, true>), 227 * 1024,
                               &smem_state, "cudaFuncSetAttribute(gemm_tc<tgn1>)"))
    return rc;
  const int total = p.m_tiles * p.n_tiles;
  const long long per_img = (long long)p.g_Ho * p.g_Wo;
  static const int env_ctas = env_int("SNAPB200_T1_CTAS", 0);
  int ctas = 2, stages = 2, tab = 1, grid = 1;
  for (;; --ctas) {
    const int cap = num_sms() * ctas;
    grid = total < cap ? total : cap;
    // images a CTA's contiguous tile range can touch: its m-tiles span <= ceil(tiles per CTA / n_tiles) + 1 row blocks
    const long long tiles_per_cta = (total + grid - 1) / grid + 1;
    const long long span_rows = ((tiles_per_cta + p.n_tiles - 1) / p.n_tiles + 1) * 128;
    long long t = (span_rows + per_img - 1) / per_img + 1;
    tab = (int)(t < p.g_nimg ? t : p.g_nimg);
    const int budget = ctas == 2 ? (225 * 1024) / 2 : 225 * 1024;
    stages = p.nkb < Cfg::MAX_STAGES ? p.nkb : Cfg::MAX_STAGES;
    if (stages < 2) stages = 2;
    while (stages > 2 && Cfg::smem_bytes_t1(stages, tab, p.g_C) > budget) --stages;
    const bool fits = Cfg::smem_bytes_t1(stages, tab, p.g_C) <= budget;
    // a deep K loop wants more than two stages: fall back to one CTA per SM with a deep ring
    const bool deep_enough = stages >= 3 || p.nkb <= 2;
    if (ctas == 1) {
      if (!fits) return set_error(SNAPB200_ERR_INVALID, "conv_gn: shared memory budget exceeded (C=%d)", p.g_C);
      break;
    }
    if (env_ctas == 1) continue;
    if (fits && (deep_enough || env_ctas == 2)) break;
  }
  p.stages = stages;
  p.g_tab_imgs = tab;
  p.stage_out = 1;
  gemm_tc_kernel<BN, 64, AMODE_TGN1, true><<<grid, GEMM_THREADS_T1, Cfg::smem_bytes_t1(stages, tab, p.g_C), s>>>(
      tmA, tmB, tmO, tmR, p);
  return check_launch("gemm_tc_kernel<tgn1>");
}
}  // namespace snapb200

using namespace snapb200;

/* conv( relu?( GroupNorm( relu?(x) ) ) ) in ONE launch: the A operand of the tcgen05 GEMM is produced in-kernel
   from the raw activation tensor (see SnapConvGnParams). */
extern "C" int snapb200_conv_gn_bf16(const SnapConvGnParams* q, void* stream) {
  SNAP_REQUIRE(q != nullptr, "null params");
  SNAP_REQUIRE(q->x && q->acc && q->scale && q->bias && q->b && q->out, "null operand");
  SNAP_REQUIRE(q->C % 64 == 0 && q->C <= 2048, "C must be a multiple of 64, <= 2048 (got %d)", q->C);
  SNAP_REQUIRE(q->n_img >= 1, "n_img must be positive");
  SNAP_REQUIRE(q->taps == 1 || q->taps == 9, "taps must be 1 (1x1) or 9 (3x3, pad 1)");
  SNAP_REQUIRE(q->stride == 1 || q->stride == 2, "stride must be 1 or 2");
  SNAP_REQUIRE(q->n >= 16 && q->n % 16 == 0 && q->ldo % 8 == 0, "n must be a multiple of 16, ldo of 8");
  SNAP_REQUIRE(q->replica_stride >= q->n_img * 64, "replica_stride too small");
  const int Ho = (q->H - 1) / q->stride + 1, Wo = (q->W - 1) / q->stride + 1;
  GemmParams p = {};
  int bn = q->n >= 256 ? 256 : (q->n > 64 ? 128 : 64);
  const long long M = (long long)q->n_img * Ho * Wo;
  p.m_tiles = (int)((M + 127) / 128);
  p.n_tiles = (q->n + bn - 1) / bn;
  p.kps = q->C / 64;
  p.nkb = q->taps * p.kps;
  p.seg_kstride = q->C;
  p.seg_mode = SEG_TABLE;
  p.tile_mode = TILE_LINEAR;
  p.epi = EPI_STORE;
  p.M_valid = M;
  p.N = q->n;
  p.out = q->out;
  p.ldo = q->ldo;
  p.residual = static_cast<const __nv_bfloat16*>(q->residual);
  p.ldr = q->ldr;
  p.gn_acc = q->gn_acc;
  p.gn_acc_relu = q->gn_acc_relu;
  p.gn_rows_per_img = (long long)Ho * Wo;
  p.gn_cpg = q->n / 32;
  p.gn_cpg_log = 0;
  while ((1 << p.gn_cpg_log) < p.gn_cpg) ++p.gn_cpg_log;
  if (q->gn_acc != nullptr) SNAP_REQUIRE((1 << p.gn_cpg_log) == p.gn_cpg, "gn_acc needs a power-of-two n / 32");
  p.gn_replica_stride = q->gn_replica_stride;
  if (q->gn_acc != nullptr)
    SNAP_REQUIRE(q->n % 64 == 0 && q->gn_replica_stride > 0, "gn_acc needs n %% 64 == 0 and gn_replica_stride");
  SNAP_REQUIRE(q->gn_acc_relu == nullptr || q->gn_acc != nullptr, "gn_acc_relu needs gn_acc");
  p.g_raw = static_cast<const __nv_bfloat16*>(q->x);
  p.g_acc = q->acc;
  p.g_scale = q->scale;
  p.g_bias = q->bias;
  p.g_rep_stride = q->replica_stride;
  p.g_nimg = q->n_img;
  p.g_C = q->C;
  p.g_H = q->H;
  p.g_W = q->W;
  p.g_Ho = Ho;
  p.g_Wo = Wo;
  p.g_stride = q->stride;
  p.g_taps = q->taps;
  p.g_pre_relu = q->pre_relu;
  p.g_post_relu = q->post_relu;
  CUtensorMap tmA, tmB;
  int rc = make_tmap_2d_bf16(&tmB, q->b, q->b_rows, q->b_cols, q->b_ld, bn, 64);
  if (rc) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  static const int env_t1 = env_int("SNAPB200_CONV_GN_T1", 1);
  const int cpg_in = q->C / 32;
  const bool t1_ok = env_t1 && q->stride == 1 && q->taps == 1 && q->n % 64 == 0 && M < (1LL << 31) &&
                     (cpg_in & (cpg_in - 1)) == 0 &&
                     (reinterpret_cast<uintptr_t>(q->out) & 15) == 0 &&
                     (q->residual == nullptr || (q->ldr % 8 == 0 && (reinterpret_cast<uintptr_t>(q->residual) & 15) == 0));
  if (t1_ok) {
    // 1x1 conv: the two-CTAs-per-SM form with the conv epilogue (staged TMA stores, TMA residual)
    bn = q->n % 128 == 0 ? 128 : 64;
    // wide-K layers with N >= 256 (conv1 of stages 3-4): 128 x 256 tiles at one CTA per SM, so that the A tile is
    // normalised once per 256 output columns instead of once per 128
    static const int env_bn256 = env_int("SNAPB200_T1_BN256", 1);
    if (env_bn256 && q->n % 256 == 0 && q->C >= 512) bn = 256;
    p.n_tiles = q->n / bn;
    while ((1 << p.g_cpg_log) < cpg_in) ++p.g_cpg_log;
    CUtensorMap tmO, tmR;
    memset(&tmR, 0, sizeof(tmR));
    rc = make_tmap_2d_bf16(&tmB, q->b, q->b_rows, q->b_cols, q->b_ld, bn, 64);
    if (rc) return rc;
    rc = make_tmap_2d_bf16(&tmA, q->x, M, q->C, q->C, 128, 64);
    if (rc) return rc;
    rc = make_tmap_2d_bf16(&tmO, q->out, M, q->n, q->ldo, 32, 64);
    if (rc) return rc;
    p.res_tma = q->residual != nullptr;
    if (p.res_tma) {
      rc = make_tmap_2d_bf16(&tmR, q->residual, M, q->n, q->ldr, 32, 64);
      if (rc) return rc;
    }
    if (bn == 64) return launch_t1_inst<64>(tmA, tmB, tmO, tmR, p, s);
    if (bn == 256) return launch_t1_inst<256>(tmA, tmB, tmO, tmR, p, s);
    return launch_t1_inst<128>(tmA, tmB, tmO, tmR, p, s);
  }
  // the round-1 modes below keep the statistics of ALL images in a fixed shared-memory table
  SNAP_REQUIRE(q->n_img <= GEMM_GN_MAX_IMG, "this conv_gn shape (3x3, stride 2 or n %% 64 != 0) handles at most %d images per call",
               GEMM_GN_MAX_IMG);
  if (q->stride == 1) {
    // TMA loads the RAW tile from the dense tensor (row-shifted per tap); transformer warps normalise in smem
    rc = make_tmap_2d_bf16(&tmA, q->x, (long long)q->n_img * q->H * q->W, q->C, q->C, 128, 64);
    if (rc) return rc;
    for (int t = 0; t < q->taps; ++t)
      p.seg_off[t] = q->taps == 9 ? (t / 3 - 1) * q->W + (t % 3 - 1) : 0;
    if (bn == 64) return launch_gn_inst<64, AMODE_TGN>(tmA, tmB, p, s);
    if (bn == 128) return launch_gn_inst<128, AMODE_TGN>(tmA, tmB, p, s);
    return launch_gn_inst<256, AMODE_TGN>(tmA, tmB, p, s);
  }
  if (bn == 64) return launch_gn_inst<64, AMODE_GN>(tmB, tmB, p, s);
  if (bn == 128) return launch_gn_inst<128, AMODE_GN>(tmB, tmB, p, s);
  return launch_gn_inst<256, AMODE_GN>(tmB, tmB, p, s);
}

extern "C" int snapb200_gemm_bf16(const SnapGemmParams* q, void* stream) {
  SNAP_REQUIRE(q != nullptr, "null params");
  SNAP_REQUIRE(q->a && q->b && q->out, "null operand");
  SNAP_REQUIRE(q->num_seg >= 1 && q->num_seg <= 9, "num_seg must be in 1..9 (got %d)", q->num_seg);
  SNAP_REQUIRE(q->seg_k >= 32 && q->seg_k % 32 == 0, "seg_k must be a multiple of 32 (got %d)",
               q->seg_k);
  SNAP_REQUIRE(q->n >= 16 && q->n % 16 == 0, "n must be a multiple of 16 (got %d)", q->n);
  SNAP_REQUIRE(q->m_rows > 0, "m_rows must be positive");
  SNAP_REQUIRE(q->ldo % 8 == 0, "ldo must be a multiple of 8");
  SNAP_REQUIRE(q->residual == nullptr || q->ldr % 8 == 0, "ldr must be a multiple of 8");
  const int bk = (q->seg_k % 64 == 0) ? 64 : 32;
  int bn = 0, ctas = 1;
  pick_config(q->n, q->num_seg * q->seg_k, bk, &bn, &ctas);
  if (q->bn > 0) {
    bn = q->bn;
    ctas = bn <= 64 ? 3 : (bn <= 128 ? 2 : 1);
  }
  GemmParams p = {};
  p.m_tiles = (int)((q->m_rows + 127) / 128);
  p.n_tiles = (q->n + bn - 1) / bn;
  p.kps = q->seg_k / bk;
  p.nkb = q->num_seg * p.kps;
  p.seg_kstride = q->seg_k;
  p.a_col0 = q->a_col0;
  p.seg_mode = SEG_TABLE;
  for (int i = 0; i < 9; ++i) p.seg_off[i] = i < q->num_seg ? q->seg_off[i] : 0;
  p.tile_mode = TILE_LINEAR;
  p.epi = EPI_STORE;
  p.M_valid = q->m_rows;
  p.N = q->n;
  p.out = q->out;
  p.ldo = q->ldo;
  p.out_f32 = q->out_f32;
  p.residual = static_cast<const __nv_bfloat16*>(q->residual);
  p.ldr = q->ldr;
  p.bias = q->bias;
  p.row_mask = q->row_mask;
  p.relu = q->relu;
  p.remap = q->remap;
  p.rm_R = q->rm_R;
  p.rm_C = q->rm_C;
  p.rm_r0 = q->rm_r0;
  p.rm_c0 = q->rm_c0;
  p.rm_Ho = q->rm_Ho;
  p.rm_Wo = q->rm_Wo;
  p.gn_acc = q->gn_acc;
  p.gn_acc_relu = q->gn_acc_relu;
  p.gn_rows_per_img = q->gn_rows_per_img;
  p.gn_cpg = q->n / 32;
  p.gn_cpg_log = 0;
  while ((1 << p.gn_cpg_log) < p.gn_cpg) ++p.gn_cpg_log;
  if (q->gn_acc != nullptr) SNAP_REQUIRE((1 << p.gn_cpg_log) == p.gn_cpg, "gn_acc needs a power-of-two n / 32");
  p.gn_replica_stride = q->gn_replica_stride;
  if (q->gn_acc != nullptr) {
    SNAP_REQUIRE(q->n % 64 == 0 && q->gn_rows_per_img > 0, "gn_acc needs n %% 64 == 0 and gn_rows_per_img");
    SNAP_REQUIRE(q->gn_replica_stride > 0, "gn_replica_stride (doubles between accumulator replicas) required");
    SNAP_REQUIRE(!q->out_f32 && q->bias == nullptr && !q->relu && q->row_mask == nullptr,
                 "gn_acc is only defined for plain bf16 conv outputs (optional residual)");
  }
  SNAP_REQUIRE(q->gn_acc_relu == nullptr || q->gn_acc != nullptr, "gn_acc_relu needs gn_acc");
  if (q->remap) SNAP_REQUIRE(q->rm_R > 0 && q->rm_C > 0 && q->rm_Ho > 0 && q->rm_Wo > 0, "bad remap");
  return launch_gemm(q->a, q->a_rows, q->a_cols, q->a_ld, q->b, q->b_rows, q->b_cols, q->b_ld, bn,
                     bk, p, static_cast<cudaStream_t>(stream), ctas);
}

/* Implicit root conv: see snapb200.h.  M space = [n_img*Ho row blocks] x [wtiles*128 columns]; the epilogue's row remap
   drops the columns >= Wo of the last block. */
extern "C" int snapb200_root_conv_bf16(const SnapRootConvParams* q, void* stream) {
  SNAP_REQUIRE(q != nullptr && q->packed && q->b && q->out, "null operand");
  SNAP_REQUIRE(q->cp == 4 || q->cp == 8, "cp must be 4 or 8");
  SNAP_REQUIRE(q->stride * q->cp * 2 == 16, "window step must be 16 bytes (stride 2 with cp 4, stride 1 with cp 8)");
  SNAP_REQUIRE(q->KH >= 1 && q->KH <= 9, "KH out of range");
  SNAP_REQUIRE(q->n >= 16 && q->n % 16 == 0 && q->n <= 256 && q->ldo % 8 == 0, "bad n / ldo");
  SNAP_REQUIRE((q->Ho - 1) * q->stride + q->KH <= q->Hq, "packed image too short");
  SNAP_REQUIRE(((q->Wo - 1) * q->stride) * q->cp + 32 <= q->Wq * q->cp, "packed image too narrow");
  const int wtiles = (q->Wo + 127) / 128;
  GemmParams p = {};
  p.a4d = 1;
  p.r4_wtiles = wtiles;
  p.r4_Ho = q->Ho;
  p.r4_stride = q->stride;
  const int bn = pick_bn(q->n, 32);
  p.m_tiles = q->n_img * q->Ho * wtiles;
  p.n_tiles = (q->n + bn - 1) / bn;
  p.kps = 1;
  p.nkb = q->KH;
  p.seg_kstride = 32;
  p.seg_mode = SEG_TABLE;
  p.tile_mode = TILE_LINEAR;
  p.epi = EPI_STORE;
  p.M_valid = (long long)p.m_tiles * 128;
  p.N = q->n;
  p.out = q->out;
  p.ldo = q->ldo;
  p.remap = 1;
  p.rm_R = q->Ho;
  p.rm_C = wtiles * 128;
  p.rm_r0 = 0;
  p.rm_c0 = 0;
  p.rm_Ho = q->Ho;
  p.rm_Wo = q->Wo;
  p.gn_acc = q->gn_acc;
  p.gn_rows_per_img = (long long)q->Ho * q->Wo;
  p.gn_cpg = q->n / 32;
  p.gn_cpg_log = 0;
  while ((1 << p.gn_cpg_log) < p.gn_cpg) ++p.gn_cpg_log;
  if (q->gn_acc != nullptr) SNAP_REQUIRE((1 << p.gn_cpg_log) == p.gn_cpg, "gn_acc needs a power-of-two n / 32");
  p.gn_replica_stride = q->gn_replica_stride;
  if (q->gn_acc != nullptr)
    SNAP_REQUIRE(q->n % 64 == 0 && q->gn_replica_stride > 0, "gn_acc needs n %% 64 == 0 and gn_replica_stride");
  CUtensorMap tmA, tmB;
  int rc = make_tmap_window4d_bf16(&tmA, q->packed, q->Wo, (long long)q->stride * q->cp * 2, q->Hq,
                                   (long long)q->Wq * q->cp * 2, q->n_img);
  if (rc) return rc;
  rc = make_tmap_2d_bf16(&tmB, q->b, q->n, q->KH * 32, q->KH * 32, bn, 32);
  if (rc) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  static const int env_r3d = env_int("SNAPB200_ROOT_STAGED", 1);
  if (env_r3d && bn == 64 && q->n == 64 && q->gn_acc == nullptr && q->ldo % 8 == 0 &&
      (reinterpret_cast<uintptr_t>(q->out) & 15) == 0) {
    // staged epilogue: 32-pixel slabs leave through a 3D TMA map over [n_img * Ho][Wo][C] that clips the columns >= Wo of
    // the last 128-column block -- no per-row remap arithmetic, whole 128-byte lines instead of 32 B per lane
    CUtensorMap tmO;
    rc = make_tmap_rows3d_bf16(&tmO, q->out, q->n, q->ldo, q->Wo, (long long)q->n_img * q->Ho);
    if (rc) return rc;
    p.remap = 0;
    p.stage_out = 1;
    return launch_inst<64, 32, true>(tmA, tmB, tmO, tmO, p, 2, s);
  }
  switch (bn) {
    case 64: return launch_inst<64, 32>(tmA, tmB, tmA, tmA, p, 2, s);
    case 128: return launch_inst<128, 32>(tmA, tmB, tmA, tmA, p, 2, s);
    case 256: return launch_inst<256, 32>(tmA, tmB, tmA, tmA, p, 1, s);
    default: return set_error(SNAPB200_ERR_INVALID, "root conv: unsupported channel count %d", q->n);
  }
}
