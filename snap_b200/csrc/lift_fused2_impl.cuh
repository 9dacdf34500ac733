// Body of the second-generation fused lift (see lift_fused2.cu), compiled once per producer-warp count:
// #define F2_NP_VALUE <10|12> and F2_NS <namespace> before including.  No include guard on purpose.
namespace snapb200 {
namespace F2_NS {

constexpr int F2_NP = F2_NP_VALUE;                  // producer warps (10: 16 warps, 128 registers; 12: 18 warps, 96 registers)
constexpr int F2_NC = 4;                           // consumer warps: warp index % 4 == TMEM lane quadrant
constexpr int F2_WARPS = 2 + F2_NP + F2_NC;        // 16
constexpr int F2_THREADS = F2_WARPS * 32;          // 512
constexpr int F2_PTHREADS = F2_NP * 32;            // 320
constexpr int F2_CTHREADS = F2_NC * 32;            // 128
constexpr int F2_BATCH_COLS = F2_PTHREADS / 64;    // BEV columns per visibility batch (one thread per z slot)
constexpr int F2_LIST_CAP = 512;                   // ring of compacted visible voxels (127 pending + 320 new)
constexpr int F2_MAXV = 4;                         // views per scene
constexpr int F2_MAX_VIEWS_TOTAL = 32;             // B * V per launch
constexpr int F2_NHW = 2 * F2_NP;                  // half-warps that gather (one tile row each per iteration)
constexpr int F2_NIT = (128 + F2_NHW - 1) / F2_NHW;  // 7 row iterations per tile
constexpr int F2_NROUND = (F2_NIT + 3) / 4;        // record-prefetch rounds of 4 iterations
constexpr int F2_WSLOTS = 2;                       // weight ring slots (16 KB each)
constexpr int F2_GSLOTS = 2;                       // gather staging slots per half-warp
constexpr int F2_GSLOT_BYTES = 1088;               // 4 taps x 256 B of features + 16 x 4 B depth-score words
// warps 0..3: consumers (warp index = TMEM lane quadrant), warp 4: TMA producer, warp 5: MMA issuer, warps 6..: producers
constexpr int F2_W_TMA = F2_NC, F2_W_MMA = F2_NC + 1, F2_W_P0 = F2_NC + 2;

constexpr int S2_A = 0;                                      // 2 x A / H / volume staging [128 x 256] bf16   (131072)
constexpr int S2_W = 131072;                                 // weight ring
constexpr int S2_STG = S2_W + F2_WSLOTS * 16384;             // gather staging
constexpr int S2_LIST = S2_STG + F2_NHW * F2_GSLOTS * F2_GSLOT_BYTES;  // uint32[F2_LIST_CAP]
constexpr int S2_SMAX = S2_LIST + 4 * F2_LIST_CAP;           // bf16[2][128]
constexpr int S2_RCOL = S2_SMAX + 512;                       // uint32[2][128] global BEV column of every tile row
constexpr int S2_CTL = S2_RCOL + 1024;
constexpr int S2_VIEW = S2_CTL + 512;                        // LiftView[32]
constexpr int S2_CULL = S2_VIEW + 2944;                      // CullView[32]
constexpr int S2_TBL = S2_CULL + 2688;                       // epilogue tables: w256 f32[256], b1 bf16x2[128], b2 bf16x2[64]
constexpr int S2_COORD = S2_TBL + 1024 + 512 + 256;           // voxel coordinates: zs f32[8][64] | xs f32[256] | ys f32[256]
constexpr int F2_ZS_CACHE = 8 * 64;                          // floats: launches of up to 8 scenes (larger ones read global memory)
constexpr int F2_XY_CACHE = 256;                             // floats per axis (separable grids up to 256 cells)
constexpr int F2_SMEM_BYTES = S2_COORD + 4 * (F2_ZS_CACHE + 2 * F2_XY_CACHE);
static_assert(sizeof(LiftView) * F2_MAX_VIEWS_TOTAL <= 2944, "view table");
static_assert(S2_STG % 16 == 0 && F2_GSLOT_BYTES % 16 == 0 && S2_TBL % 16 == 0, "16-byte aligned staging / tables");
static_assert(F2_SMEM_BYTES <= 232448, "shared memory budget (227 KB)");
constexpr int VOL2_STRIDE = 272;                             // bytes per staged volume row (256 + 16)
constexpr int S2_BREC = 128 * VOL2_STRIDE;                   // z-max boundary records inside the tile buffer

struct TapRec2 {
  uint32_t off00, off01, off10, off11;  // element offsets of the four taps into fimg (scene and view base included)
  float wr1, wc1;                       // weights of the upper taps
  float wb1;                            // weight of the upper depth bin
  uint16_t b0, b1;                      // depth bins
};
static_assert(sizeof(TapRec2) == 32, "TapRec2 layout");

struct CullView {   // four frustum planes g.p + h (+ cg * (|p|_1 + 1)) >= 0, conservative
  float g[4][3], h[4], cg[4];
  int enabled;
};
static_assert(sizeof(CullView) == 84 && sizeof(CullView) * F2_MAX_VIEWS_TOTAL <= 2688, "cull table");

struct Fused2Args {
  LiftParams P;
  int B;                      // scenes
  long long views_stride;     // LiftView elements between scenes
  long long fimg_stride;      // bf16 elements between scenes
  long long zs_stride;        // floats between scenes
  const LiftView* views;
  const __nv_bfloat16* fimg;
  const float *xs, *ys, *zs;
  const float* w256;
  const float* b1;
  const float* b2;
  __nv_bfloat16* plane;       // [B * X * Y, 128]
  uint8_t* pvalid;            // [B * X * Y]
  int* col_counter;
  TapRec2* scratch;           // [gridDim.x][F2_LIST_CAP][F2_MAXV]
};

struct Ctl2 {
  uint64_t a_full[2], a_empty[2], w_full[F2_WSLOTS], w_empty[F2_WSLOTS], acc1_full, acc1_empty, h_full[4], acc2_full;
  uint32_t tmem_ptr;
  int batch_col0[2];   // claimed visibility batches, double-buffered (one producer barrier per batch)
  // tiles announced by the producers / no more tiles: plain counters polled by the TMA thread (an mbarrier would lose a
  // phase when the producers run two tiles ahead of it)
  int ann_tiles, ann_done;
  int rows[2];
  int warp_cnt[2][F2_NP];
};
static_assert(sizeof(Ctl2) <= 512, "control block");

__device__ __forceinline__ void pbar() { asm volatile("bar.sync 1, %0;" ::"n"(F2_PTHREADS) : "memory"); }
__device__ __forceinline__ void cbar() { asm volatile("bar.sync 2, %0;" ::"n"(F2_CTHREADS) : "memory"); }
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ uint2 hmax2x2(uint2 a, uint2 b) { return make_uint2(hmax2_bf16(a.x, b.x), hmax2_bf16(a.y, b.y)); }

// weighted pooling of >= 2 views (streetview_encoder.py:156-164): softmax over the visible views' scores (shifted by
// max(0, max score), A.6), mean = sum w f, var = sum w (f - mean)^2 in fp32 -> feature dtype; score_max.
__device__ __forceinline__ void pool_views(const uint32_t (&fvp)[F2_MAXV][4], const float (&score)[F2_MAXV], int nv,
                                        uint32_t (&mean_p)[4], uint32_t (&var_p)[4], float& smaxv) {
  float mean[8], var[8];
#pragma unroll
  for (int jj = 0; jj < 8; ++jj) mean[jj] = var[jj] = 0.f;
  float mx = 0.f;
  smaxv = -INFINITY;
#pragma unroll
  for (int k = 0; k < F2_MAXV; ++k)
    if (k < nv) {
      mx = fmaxf(mx, score[k]);
      smaxv = fmaxf(smaxv, score[k]);
    }
  float wv[F2_MAXV], den = 0.f;
#pragma unroll
  for (int k = 0; k < F2_MAXV; ++k) {
    wv[k] = (k < nv) ? expf(score[k] - mx) : 0.f;
    den += wv[k];
  }
#pragma unroll
  for (int k = 0; k < F2_MAXV; ++k) {
    if (k >= nv) continue;
    wv[k] = __fdiv_rn(wv[k], den);
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      mean[2 * jj] += wv[k] * bf16_lo(fvp[k][jj]);
      mean[2 * jj + 1] += wv[k] * bf16_hi(fvp[k][jj]);
    }
  }
#pragma unroll
  for (int k = 0; k < F2_MAXV; ++k) {
    if (k >= nv) continue;
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const float da = bf16_lo(fvp[k][jj]) - mean[2 * jj], db = bf16_hi(fvp[k][jj]) - mean[2 * jj + 1];
      var[2 * jj] += wv[k] * da * da;
      var[2 * jj + 1] += wv[k] * db * db;
    }
  }
#pragma unroll
  for (int jj = 0; jj < 4; ++jj) {
    mean_p[jj] = pack_bf16(mean[2 * jj], mean[2 * jj + 1]);
    var_p[jj] = pack_bf16(var[2 * jj], var[2 * jj + 1]);
  }
}

__global__ void __launch_bounds__(F2_THREADS, 1)
lift_fused2_kernel(const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmW2,
                   const __grid_constant__ Fused2Args A) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const LiftParams& P = A.P;
  Ctl2* ctl = reinterpret_cast<Ctl2*>(smem + S2_CTL);
  uint32_t* list = reinterpret_cast<uint32_t*>(smem + S2_LIST);
  __nv_bfloat16* smax_s = reinterpret_cast<__nv_bfloat16*>(smem + S2_SMAX);
  uint32_t* rowcol = reinterpret_cast<uint32_t*>(smem + S2_RCOL);
  LiftView* sview = reinterpret_cast<LiftView*>(smem + S2_VIEW);
  CullView* scull = reinterpret_cast<CullView*>(smem + S2_CULL);
  float* tbl_w256 = reinterpret_cast<float*>(smem + S2_TBL);                 // W1[256, :] (score_max input row)
  uint32_t* tbl_b1 = reinterpret_cast<uint32_t*>(smem + S2_TBL + 1024);      // b1 as packed bf16 pairs
  uint32_t* tbl_b2 = reinterpret_cast<uint32_t*>(smem + S2_TBL + 1536);      // b2 as packed bf16 pairs
  // voxel coordinates in shared memory: with ~226 KB of it there is next to no L1, so the three coordinate loads at the top
  // of every visibility batch were exposed L2 round trips
  float* zs_s = reinterpret_cast<float*>(smem + S2_COORD);
  float* xs_s = zs_s + F2_ZS_CACHE;
  float* ys_s = xs_s + F2_XY_CACHE;
  const bool zs_cached = A.B * 64 <= F2_ZS_CACHE;
  const bool xy_cached = !P.xy_paired && P.X <= F2_XY_CACHE && P.Y <= F2_XY_CACHE;
  if (zs_cached)
    for (int i = threadIdx.x; i < A.B * 64; i += blockDim.x) {
      const int sc = i >> 6, z = i & 63;
      zs_s[i] = z < P.Z ? A.zs[(size_t)sc * A.zs_stride + z] : 0.f;
    }
  if (xy_cached) {
    for (int i = threadIdx.x; i < P.X; i += blockDim.x) xs_s[i] = A.xs[i];
    for (int i = threadIdx.x; i < P.Y; i += blockDim.x) ys_s[i] = A.ys[i];
  }

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if ((smem_u32(smem) & 1023u) != 0) __trap();  // swizzled operands need a 1024 B aligned base
  const int nviews = A.B * P.V;

  for (int i = threadIdx.x; i < nviews * (int)(sizeof(LiftView) / 4); i += blockDim.x) {
    const int v = i / (int)(sizeof(LiftView) / 4), w = i - v * (int)(sizeof(LiftView) / 4);
    const int sc = v / P.V, vv = v - sc * P.V;
    reinterpret_cast<uint32_t*>(sview)[i] = reinterpret_cast<const uint32_t*>(A.views + (size_t)sc * A.views_stride + vv)[w];
  }
  // epilogue constants: there is next to no L1 beside 226 KB of shared memory, so a per-chunk __ldg is an L2 round trip
  for (int i = threadIdx.x; i < 256; i += blockDim.x) tbl_w256[i] = A.w256[i];
  for (int i = threadIdx.x; i < 128; i += blockDim.x) tbl_b1[i] = pack_bf16(A.b1[2 * i], A.b1[2 * i + 1]);
  for (int i = threadIdx.x; i < 64; i += blockDim.x) tbl_b2[i] = pack_bf16(A.b2[2 * i], A.b2[2 * i + 1]);
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmW1);
    tma_prefetch_desc(&tmW2);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&ctl->a_full[s], F2_PTHREADS);
      mbar_init(&ctl->a_empty[s], F2_CTHREADS);
    }
    for (int s = 0; s < F2_WSLOTS; ++s) {
      mbar_init(&ctl->w_full[s], 1);
      mbar_init(&ctl->w_empty[s], 1);
    }
    mbar_init(&ctl->acc1_full, 1);
    mbar_init(&ctl->acc1_empty, F2_CTHREADS);
    for (int kc = 0; kc < 4; ++kc) mbar_init(&ctl->h_full[kc], F2_CTHREADS);
    mbar_init(&ctl->acc2_full, 1);
    ctl->rows[0] = ctl->rows[1] = 0;
    ctl->ann_tiles = 0;
    ctl->ann_done = 0;
    fence_barrier_init();
  }
  if (warp == F2_W_MMA) tmem_alloc(&ctl->tmem_ptr, 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  // frustum planes of every view (pinhole cameras with positive focal lengths; anything else is not culled)
  if ((int)threadIdx.x < nviews) {
    const LiftView& v = sview[threadIdx.x];
    CullView c;
    const float fx = v.f[0], fy = v.f[1], cx = v.c[0], cy = v.c[1], W = v.wh[0], H = v.wh[1];
    const float* r0 = v.Rinv;
    const float* r1 = v.Rinv + 3;
    const float* r2 = v.Rinv + 6;
    bool ok = !v.fisheye && fx > 0.f && fy > 0.f && W > 0.f && H > 0.f;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      c.g[0][j] = fx * r0[j] + cx * r2[j];          // u >= 0
      c.g[1][j] = (W - cx) * r2[j] - fx * r0[j];    // u <  W
      c.g[2][j] = fy * r1[j] + cy * r2[j];          // w >= 0
      c.g[3][j] = (H - cy) * r2[j] - fy * r1[j];    // w <  H
    }
    c.h[0] = fx * v.tinv[0] + cx * v.tinv[2];
    c.h[1] = (W - cx) * v.tinv[2] - fx * v.tinv[0];
    c.h[2] = fy * v.tinv[1] + cy * v.tinv[2];
    c.h[3] = (H - cy) * v.tinv[2] - fy * v.tinv[1];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float mag = fabsf(c.g[k][0]) + fabsf(c.g[k][1]) + fabsf(c.g[k][2]) + fabsf(c.h[k]);
      // margin: >= 100 x the fp32 rounding error of this plane test and of the reference's own projection chain
      c.cg[k] = 1e-5f * mag;
      c.h[k] += 1e-3f;
      ok = ok && isfinite(mag);
    }
    c.enabled = ok ? 1 : 0;
    scull[threadIdx.x] = c;
  }
  __syncthreads();
  const uint32_t tmem_base = ctl->tmem_ptr;
  const uint32_t tmem_acc1 = tmem_base;        // 256 columns
  const uint32_t tmem_acc2 = tmem_base + 256;  // 128 columns
  constexpr uint32_t idesc128 = make_idesc_bf16_m128(128);

  if (warp == F2_W_TMA) {
    // ============================ TMA producer: 12 weight chunks per tile ============================
    if (elect_one()) {
      uint32_t g = 0;
      for (uint32_t tile = 0;; ++tile) {
        bool have = false;
        while (true) {  // `done` is read before the count: the producers write the count first
          const int done = *reinterpret_cast<volatile int*>(&ctl->ann_done);
          const int n = *reinterpret_cast<volatile int*>(&ctl->ann_tiles);
          have = n > (int)tile;
          if (have || done) break;
          __nanosleep(128);
        }
        if (!have) break;
        for (int j = 0; j < 12; ++j, ++g) {
          const uint32_t slot = g % F2_WSLOTS, use = g / F2_WSLOTS;
          mbar_wait_sleep(&ctl->w_empty[slot], (use & 1) ^ 1, 32);
          mbar_arrive_expect_tx(&ctl->w_full[slot], 16384);
          if (j < 8)
            tma_load_2d(&tmW1, &ctl->w_full[slot], smem + S2_W + slot * 16384, (j & 3) * 64, (j >> 2) * 128);
          else
            tma_load_2d(&tmW2, &ctl->w_full[slot], smem + S2_W + slot * 16384, (j - 8) * 64, 0);
        }
      }
    }
  } else if (warp == F2_W_MMA) {
    // ============================ MMA issuer ============================
    if (elect_one()) {
      uint32_t g = 0;
      const uint32_t sW = smem_u32(smem + S2_W);
      for (uint32_t tile = 0;; ++tile) {
        const uint32_t b = tile & 1;
        mbar_wait_sleep(&ctl->a_full[b], (tile >> 1) & 1, 64);
        if (*reinterpret_cast<volatile int*>(&ctl->rows[b]) == 0) break;
        mbar_wait_sleep(&ctl->acc1_empty, (tile & 1) ^ 1, 32);  // epilogue 1 of the previous tile has drained acc1
        tc_fence_after_sync();
        const uint32_t sA = smem_u32(smem + S2_A + b * 65536);
        // GEMM1: acc1[128 x 256] = A[128 x 256] * W1^T, two column halves of 128
        for (int j = 0; j < 8; ++j, ++g) {
          const uint32_t slot = g % F2_WSLOTS, use = g / F2_WSLOTS;
          const int nh = j >> 2, kc = j & 3;
          mbar_wait(&ctl->w_full[slot], use & 1);
          tc_fence_after_sync();
          const uint64_t da = make_kmajor_desc<128>(sA + kc * 16384);
          const uint64_t db = make_kmajor_desc<128>(sW + slot * 16384);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem_acc1 + (uint32_t)(nh * 128), da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc128,
                      (kc | k) != 0 ? 1u : 0u);
          umma_commit(&ctl->w_empty[slot]);
        }
        umma_commit(&ctl->acc1_full);
        // GEMM2: acc2[128 x 128] = H[128 x 256] * W2^T, H handed over per K-chunk by the consumers
        for (int kc = 0; kc < 4; ++kc, ++g) {
          const uint32_t slot = g % F2_WSLOTS, use = g / F2_WSLOTS;
          mbar_wait_sleep(&ctl->h_full[kc], tile & 1, 32);
          mbar_wait(&ctl->w_full[slot], use & 1);
          tc_fence_after_sync();
          const uint64_t da = make_kmajor_desc<128>(sA + kc * 16384);
          const uint64_t db = make_kmajor_desc<128>(sW + slot * 16384);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem_acc2, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc128, (kc | k) != 0 ? 1u : 0u);
          umma_commit(&ctl->w_empty[slot]);
        }
        umma_commit(&ctl->acc2_full);
      }
    }
  } else if (warp >= F2_W_P0) {
    // ============================ producers ============================
    const int pw = warp - F2_W_P0;
    const int ptid = threadIdx.x - F2_W_P0 * 32;
    const int half = lane >> 4, l16 = lane & 15;
    const int hw = pw * 2 + half;            // gathering half-warp 0..19
    const int ncols = P.X * P.Y;
    const int total_cols = A.B * ncols;
    const float score_scale = (float)(P.S - 1);
    const unsigned FULL = 0xffffffffu;
    TapRec2* const my_scratch = A.scratch + (size_t)blockIdx.x * F2_LIST_CAP * F2_MAXV;
    const uint32_t stage_base = smem_u32(smem + S2_STG + hw * (F2_GSLOTS * F2_GSLOT_BYTES));
    const uint8_t* const stage_ptr = smem + S2_STG + hw * (F2_GSLOTS * F2_GSLOT_BYTES);
    int list_head = 0, list_count = 0;       // replicated in every producer thread (same decisions)
    bool cols_done = false;
    int n_tiles = 0, n_rows = 0;
    unsigned tprof[3] = {0, 0, 0};           // cycles: fill, wait for a free tile buffer, gather
    unsigned tmark = (unsigned)clock();
#define F2_MARK(arr, i)                      \
  do {                                       \
    const unsigned now_ = (unsigned)clock(); \
    arr[i] += now_ - tmark;                  \
    tmark = now_;                            \
  } while (0)

    if (ptid == 0) ctl->batch_col0[0] = atomicAdd(A.col_counter, F2_BATCH_COLS);  // first batch claim
    uint32_t bpar = 0;   // parity of the batch being processed (selects batch_col0 / warp_cnt copies)
    pbar();

    for (uint32_t tile = 0;; ++tile) {
      // ---------- visibility batches until >= 128 visible voxels are pending ----------
      // One producer barrier per batch: the claim of batch n + 1 (a global atomic, requested before the projection work
      // of batch n and stored just before the barrier) and the per-warp counts are double-buffered by batch parity.
      while (list_count < 128 && !cols_done) {
        const int c0 = ctl->batch_col0[bpar];
        if (c0 >= total_cols) {
          cols_done = true;
          break;
        }
        int next_claim = 0;
        if (ptid == 0) next_claim = atomicAdd(A.col_counter, F2_BATCH_COLS);
        const int cl = ptid >> 6, z = ptid & 63;
        const int gcol = c0 + cl;
        float prow[F2_MAXV], pcol[F2_MAXV], pdep[F2_MAXV];
        uint32_t vm = 0;
        int scene = 0;
        if (gcol < total_cols && z < P.Z) {
          scene = gcol / ncols;
          const int col = gcol - scene * ncols;
          const int ix = col / P.Y, iy = col - ix * P.Y;
          const float px = xy_cached ? xs_s[ix] : A.xs[P.xy_paired ? col : ix];
          const float py = xy_cached ? ys_s[iy] : A.ys[P.xy_paired ? col : iy];
          const float pz = zs_cached ? zs_s[scene * 64 + z] : A.zs[(size_t)scene * A.zs_stride + z];
          const float s1 = fabsf(px) + fabsf(py) + fabsf(pz) + 1.0f;
#pragma unroll
          for (int v = 0; v < F2_MAXV; ++v) {
            prow[v] = pcol[v] = pdep[v] = 0.f;
            if (v < P.V) {
              const CullView& cv = scull[scene * P.V + v];
              bool cand = true;
              if (cv.enabled) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  const float val = fmaf(cv.g[k][0], px, fmaf(cv.g[k][1], py, fmaf(cv.g[k][2], pz, cv.h[k])));
                  cand = cand && (fmaf(cv.cg[k], s1, val) >= 0.f);
                }
              }
              if (cand) {
                const Proj pr = project_point(sview[scene * P.V + v], px, py, pz);
                if (pr.vis) {
                  vm |= 1u << v;
                  prow[v] = pr.row;
                  pcol[v] = pr.col;
                  pdep[v] = pr.depth;
                }
              }
            }
          }
        }
        const bool valid = vm != 0;
        const uint32_t bal = __ballot_sync(FULL, valid);
        if (lane == 0) ctl->warp_cnt[bpar][pw] = __popc(bal);
        if (ptid == 0) ctl->batch_col0[bpar ^ 1] = next_claim;
        pbar();
        int before = 0, tot = 0;
#pragma unroll
        for (int w2 = 0; w2 < F2_NP; ++w2) {
          const int c = ctl->warp_cnt[bpar][w2];
          tot += c;
          if (w2 < pw) before += c;
        }
        if (valid) {
          const int slot = (list_head + list_count + before + __popc(bal & ((1u << lane) - 1u))) & (F2_LIST_CAP - 1);
          list[slot] = ((uint32_t)gcol << 14) | ((uint32_t)z << 8) | vm;
          TapRec2* rec = my_scratch + (size_t)slot * F2_MAXV;
          const uint32_t scene_off = (uint32_t)((long long)scene * A.fimg_stride);
#pragma unroll
          for (int v = 0; v < F2_MAXV; ++v) {
            if (!((vm >> v) & 1u)) continue;
            const Taps t = make_taps(prow[v], pcol[v], P.Hf, P.Wf);
            const float d = fminf(fmaxf(pdep[v], P.depth_min), P.depth_max);
            const float bi = logf(d / P.depth_min) * P.inv_log_range * score_scale;
            const float bf = floorf(bi);
            const uint32_t b0 = (uint32_t)min(max((int)bf, 0), P.S - 1);
            const uint32_t b1 = (uint32_t)min(max((int)bf + 1, 0), P.S - 1);
            const uint32_t row0 = (uint32_t)((v * P.Hf + t.r0) * P.Wf), row1 = (uint32_t)((v * P.Hf + t.r1) * P.Wf);
            uint4* dst = reinterpret_cast<uint4*>(rec + v);
            dst[0] = make_uint4(scene_off + (row0 + (uint32_t)t.c0) * (uint32_t)P.CF, scene_off + (row0 + (uint32_t)t.c1) * (uint32_t)P.CF,
                                scene_off + (row1 + (uint32_t)t.c0) * (uint32_t)P.CF, scene_off + (row1 + (uint32_t)t.c1) * (uint32_t)P.CF);
            dst[1] = make_uint4(__float_as_uint(t.wr1), __float_as_uint(t.wc1), __float_as_uint(bi - bf), b0 | (b1 << 16));
          }
        }
        list_count += tot;
        bpar ^= 1;
      }
      F2_MARK(tprof, 0);

      const int rows = min(128, list_count);
      const int head = list_head;
      const uint32_t b = tile & 1;
      if (ptid == 0) {
        // lets the TMA producer stream this tile's weights under the gather
        if (rows > 0)
          *reinterpret_cast<volatile int*>(&ctl->ann_tiles) = (int)tile + 1;
        else
          *reinterpret_cast<volatile int*>(&ctl->ann_done) = 1;
        mbar_wait(&ctl->a_empty[b], ((tile >> 1) & 1) ^ 1);   // consumers are done with this tile buffer
      }
      pbar();  // list entries / gather records of this round's batches visible; tile buffer b is free
      F2_MARK(tprof, 1);
      if (rows == 0) {
        if (ptid == 0) ctl->rows[b] = 0;
        mbar_arrive(&ctl->a_full[b]);
        break;
      }
      uint8_t* const Ab = smem + S2_A + b * 65536;

      // ---------- gather + pool: one HALF-warp per tile row (16 lanes x 8 channels), rows hw + 20 it ----------
      // Gather records are prefetched by lane (lane 4 j + k of a half holds the record of the k-th visible view of the
      // half's row of iteration 4 round + j); the tap loads of step (row, k) are asynchronous copies into a staging slot
      // issued one step ahead of the arithmetic, so a half-warp always has the next pair's 1 KB in flight.
      uint32_t pvm_n = 0;
      uint4 pq0_n = make_uint4(0u, 0u, 0u, 0u), pq1_n = make_uint4(0u, 0u, 0u, 0u);
      auto prefetch_records = [&](int round) {
        pvm_n = 0;
        pq0_n = make_uint4(0u, 0u, 0u, 0u);
        pq1_n = make_uint4(0u, 0u, 0u, 0u);
        const int pit = 4 * round + (l16 >> 2), pk = l16 & 3;
        const int prow_ = hw + F2_NHW * pit;
        if (pit < F2_NIT && prow_ < rows) {
          const int slot = (head + prow_) & (F2_LIST_CAP - 1);
          const uint32_t e = list[slot];
          pvm_n = e & 0xfu;
          if (pk == 0) rowcol[b * 128 + prow_] = e >> 14;
          uint32_t m = pvm_n;  // drop the pk lowest set bits
          if (pk > 0) m &= m - 1u;
          if (pk > 1) m &= m - 1u;
          if (pk > 2) m &= m - 1u;
          if (m != 0u) {
            const TapRec2* rec = my_scratch + (size_t)slot * F2_MAXV + (__ffs(m) - 1);
            pq0_n = __ldcg(reinterpret_cast<const uint4*>(rec));
            pq1_n = __ldcg(reinterpret_cast<const uint4*>(rec) + 1);
          }
        }
      };
      prefetch_records(0);
#pragma unroll 1
      for (int round = 0; round < F2_NROUND; ++round) {
        const uint32_t pvm = pvm_n;
        const uint4 pq0 = pq0_n, pq1 = pq1_n;
        if (round + 1 < F2_NROUND) prefetch_records(round + 1);
        // per-iteration visible-view counts, 4 bits per iteration: own half / maximum of the two halves (warp-uniform)
        uint32_t nvo_p = 0u, nva_p = 0u;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t o = (uint32_t)__popc(__shfl_sync(FULL, pvm, (lane & 16) + 4 * j));
          const uint32_t a = max(o, __shfl_xor_sync(FULL, o, 16));
          nvo_p |= o << (4 * j);
          nva_p |= a << (4 * j);
        }
        // issue the asynchronous copies of step (j, k) -- four 256-byte feature taps and the 8 depth-score taps (4 taps x
        // 2 bins, one 4-byte word each on the first 8 lanes of the half) -- into staging slot `gs`
        auto issue = [&](int j, int k, int gs) {
          const int src = (lane & 16) + 4 * j + k;
          const uint32_t x0 = __shfl_sync(FULL, pq0.x, src), x1 = __shfl_sync(FULL, pq0.y, src);
          const uint32_t x2 = __shfl_sync(FULL, pq0.z, src), x3 = __shfl_sync(FULL, pq0.w, src);
          const uint32_t y3 = __shfl_sync(FULL, pq1.w, src);
          const int own = (int)((nvo_p >> (4 * j)) & 15u);
          if (k < own) {
            const uint32_t dst = stage_base + gs * F2_GSLOT_BYTES + l16 * 16;
            cp_async16(dst, A.fimg + x0 + l16 * 8);
            cp_async16(dst + 256, A.fimg + x1 + l16 * 8);
            cp_async16(dst + 512, A.fimg + x2 + l16 * 8);
            cp_async16(dst + 768, A.fimg + x3 + l16 * 8);
            if (l16 < 8) {
              const int tap = l16 >> 1;
              const uint32_t xt = (tap & 2) ? ((tap & 1) ? x3 : x2) : ((tap & 1) ? x1 : x0);
              const uint32_t bin = (l16 & 1) ? (y3 >> 16) : (y3 & 0xffffu);
              // the aligned 32-bit word that holds the bf16 logit (tap offsets and D are even: its half is bin & 1)
              cp_async4(stage_base + gs * F2_GSLOT_BYTES + 1024 + l16 * 4, A.fimg + xt + P.D + (bin & ~1u));
            }
          }
          cp_async_commit();
        };
        // iterations of this round with at least one pair, one bit each
        const uint32_t row_mask = ((nva_p & 0xfu) ? 1u : 0u) | ((nva_p & 0xf0u) ? 2u : 0u) | ((nva_p & 0xf00u) ? 4u : 0u) |
                                  ((nva_p & 0xf000u) ? 8u : 0u);
        auto next_row = [&](int j) {  // next such iteration after j, or 4
          const uint32_t m = row_mask >> (j + 1);
          return m != 0u ? j + __ffs(m) : 4;
        };
        int gs = 0;
        int cj = next_row(-1);
        if (cj < 4) issue(cj, 0, gs);
#pragma unroll 1
        while (cj < 4) {
          const int r = hw + F2_NHW * (4 * round + cj);  // this half-warp's row
          const int nva = (int)((nva_p >> (4 * cj)) & 15u);
          const int nv = (int)((nvo_p >> (4 * cj)) & 15u);
          const int nj = next_row(cj);
          uint32_t fvp[F2_MAXV][4];
          float score[F2_MAXV];
#pragma unroll
          for (int k = 0; k < F2_MAXV; ++k) {
            score[k] = 0.f;
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) fvp[k][jj] = 0u;
          }
#pragma unroll
          for (int k = 0; k < F2_MAXV; ++k) {
            if (k >= nva) break;  // warp-uniform
            // next step: (cj, k + 1) or the first view of the next non-empty iteration of this round
            {
              const bool same = k + 1 < nva;
              const int ij = same ? cj : nj, ik = same ? k + 1 : 0;
              if (ij < 4)
                issue(ij, ik, gs ^ 1);
              else
                cp_async_commit();
            }
            cp_async_wait<1>();  // the copies of step (cj, k) have landed (this thread reads only what it copied itself)
            const int src = (lane & 16) + 4 * cj + k;
            const uint32_t y0 = __shfl_sync(FULL, pq1.x, src), y1 = __shfl_sync(FULL, pq1.y, src);
            const uint32_t y2 = __shfl_sync(FULL, pq1.z, src), y3 = __shfl_sync(FULL, pq1.w, src);
            const bool mine = k < nv;
            float sp = 0.f, wb1 = 0.f;
            if (mine) {
              const float wr1 = __uint_as_float(y0), wc1 = __uint_as_float(y1);
              wb1 = __uint_as_float(y2);
              const float wr0 = __fadd_rn(1.0f, -wr1), wc0 = __fadd_rn(1.0f, -wc1);
              const uint8_t* sp_ = stage_ptr + gs * F2_GSLOT_BYTES + l16 * 16;
              const uint4 u00 = *reinterpret_cast<const uint4*>(sp_);
              const uint4 u01 = *reinterpret_cast<const uint4*>(sp_ + 256);
              const uint4 u10 = *reinterpret_cast<const uint4*>(sp_ + 512);
              const uint4 u11 = *reinterpret_cast<const uint4*>(sp_ + 768);
              // tap weights: row weight x column weight (one rounding each), shared with the unfused kernel
              const float w00 = __fmul_rn(wr0, wc0), w01 = __fmul_rn(wr0, wc1);
              const float w10 = __fmul_rn(wr1, wc0), w11 = __fmul_rn(wr1, wc1);
              if (l16 < 8) {
                const int tap = l16 >> 1;
                const float wt = ((tap & 2) ? wr1 : 1.0f - wr1) * ((tap & 1) ? wc1 : 1.0f - wc1);
                const uint32_t word = *reinterpret_cast<const uint32_t*>(stage_ptr + gs * F2_GSLOT_BYTES + 1024 + l16 * 4);
                const uint32_t bin = (l16 & 1) ? (y3 >> 16) : (y3 & 0xffffu);
                sp = wt * ((bin & 1u) ? bf16_hi(word) : bf16_lo(word));
              }
              const uint32_t a00[4] = {u00.x, u00.y, u00.z, u00.w}, a01[4] = {u01.x, u01.y, u01.z, u01.w};
              const uint32_t a10[4] = {u10.x, u10.y, u10.z, u10.w}, a11[4] = {u11.x, u11.y, u11.z, u11.w};
#pragma unroll
              for (int jj = 0; jj < 4; ++jj) {
                // (w00 f00 + w01 f01) + (w10 f10 + w11 f11): lower tap row + upper tap row, as in the unfused kernel
                const float lo_a = __fmaf_rn(w01, bf16_lo(a01[jj]), __fmul_rn(w00, bf16_lo(a00[jj])));
                const float hi_a = __fmaf_rn(w11, bf16_lo(a11[jj]), __fmul_rn(w10, bf16_lo(a10[jj])));
                const float lo_b = __fmaf_rn(w01, bf16_hi(a01[jj]), __fmul_rn(w00, bf16_hi(a00[jj])));
                const float hi_b = __fmaf_rn(w11, bf16_hi(a11[jj]), __fmul_rn(w10, bf16_hi(a10[jj])));
                fvp[k][jj] = pack_bf16(__fadd_rn(lo_a, hi_a), __fadd_rn(lo_b, hi_b));  // -> feature dtype
              }
            }
            // bin-wise spatial interpolation (-> bf16), then interpolation across the two bins (-> bf16);
            // xor 2/4/1 stay inside the 8 score lanes of each half
            sp += __shfl_xor_sync(FULL, sp, 2);
            sp += __shfl_xor_sync(FULL, sp, 4);
            sp = bf16_round(sp) * ((lane & 1) ? wb1 : 1.0f - wb1);
            sp += __shfl_xor_sync(FULL, sp, 1);
            score[k] = __shfl_sync(FULL, bf16_round(sp), lane & 16);  // broadcast from the half's lane 0
            gs ^= 1;
          }
          if (r < rows) {
            uint32_t mean_p[4], var_p[4];
            float smaxv = 0.f;
            const float ssum = (score[0] + score[1]) + (score[2] + score[3]);
            if (nv <= 1 && ssum > -80.f) {
              // one visible view: its softmax weight is exactly 1 (x / x), so mean = its features, variance = +0 and
              // score_max = its score
#pragma unroll
              for (int jj = 0; jj < 4; ++jj) {
                mean_p[jj] = fvp[0][jj];
                var_p[jj] = 0u;
              }
              smaxv = ssum;
            } else {
              pool_views(fvp, score, nv, mean_p, var_p, smaxv);
            }
            // A[r][k]: mean at k = 8*l16.., var at k = 128 + 8*l16..; K-chunk of 64, 16-byte slot (k%64)/8 XOR (r%8)
            // inside the 128-byte row (SWIZZLE_128B, as TMA would write it)
            const int kc_m = l16 >> 3, kc_v = 2 + (l16 >> 3);
            const int slot16 = (l16 & 7) ^ (r & 7);
            *reinterpret_cast<uint4*>(Ab + kc_m * 16384 + r * 128 + slot16 * 16) =
                make_uint4(mean_p[0], mean_p[1], mean_p[2], mean_p[3]);
            *reinterpret_cast<uint4*>(Ab + kc_v * 16384 + r * 128 + slot16 * 16) =
                make_uint4(var_p[0], var_p[1], var_p[2], var_p[3]);
            if (l16 == 0) smax_s[b * 128 + r] = __float2bfloat16(smaxv);
          }
          cj = nj;
        }
        cp_async_wait<0>();
      }
      if (ptid == 0) ctl->rows[b] = rows;
      fence_proxy_async_smem();  // generic-proxy writes of A -> visible to the tensor core (async proxy)
      mbar_arrive(&ctl->a_full[b]);
      list_head = (head + rows) & (F2_LIST_CAP - 1);
      list_count -= rows;
      n_tiles += 1;
      n_rows += rows;
      F2_MARK(tprof, 2);
    }
    if (ptid == 0) {
      atomicAdd(A.col_counter + 1, n_tiles);
      atomicAdd(A.col_counter + 2, n_rows);
#pragma unroll
      for (int i = 0; i < 3; ++i) atomicAdd(A.col_counter + 4 + i, (int)(tprof[i] >> 4));  // units of 16 cycles
    }
  } else {
    // ============================ consumers ============================
    const int q = warp & 3;                     // TMEM lane quadrant = tile rows 32 q .. 32 q + 31
    const int ctid = threadIdx.x;
    uint32_t zm_col = 0xffffffffu;              // z-max carry of warp 0: lane = 4 output channels (2 x bf16x2)
    uint2 zm_val = make_uint2(0u, 0u);
    uint2* const plane64 = reinterpret_cast<uint2*>(A.plane);
    unsigned tprof[5] = {0, 0, 0, 0, 0};        // cycles: wait acc1, epilogue 1, wait acc2, epilogue 2, z-max
    unsigned tmark = (unsigned)clock();
    for (uint32_t tile = 0;; ++tile) {
      const uint32_t b = tile & 1;
      // the consumers idle for a good part of every tile: sleep between probes instead of spinning through the issue slots
      // the producers need (the polling loops were 14 % of all issued instructions)
      mbar_wait_sleep(&ctl->a_full[b], (tile >> 1) & 1, 64);
      const int rows = *reinterpret_cast<volatile int*>(&ctl->rows[b]);
      if (rows == 0) break;
      uint8_t* const Ab = smem + S2_A + b * 65536;
      const int row = q * 32 + lane;
      // ---------- epilogue 1: H = relu(bf16(bf16(acc1 + smax * w256) + b1)) -> smem (over the A tile) ----------
      mbar_wait_sleep(&ctl->acc1_full, tile & 1, 32);
      tc_fence_after_sync();
      F2_MARK(tprof, 0);
      {
        const float sm = __bfloat162float(smax_s[b * 128 + row]);
        const uint32_t taddr = tmem_acc1 + ((uint32_t)(q * 32) << 16);
        uint32_t v[2][16];
        tmem_ld16(taddr, v[0]);
#pragma unroll
        for (int c16 = 0; c16 < 16; ++c16) {
          tmem_ld_wait();
          if (c16 + 1 < 16) tmem_ld16(taddr + (uint32_t)((c16 + 1) * 16), v[(c16 + 1) & 1]);
          if (c16 == 15) {  // acc1 is drained: GEMM1 of the next tile may overwrite it
            tc_fence_before_sync();
            mbar_arrive(&ctl->acc1_empty);
          }
          const uint32_t(&vv)[16] = v[c16 & 1];
          const int n0 = c16 * 16;
          uint32_t h[8];
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            const float4 w4 = *reinterpret_cast<const float4*>(tbl_w256 + n0 + 4 * j4);
            const uint2 b2p = *reinterpret_cast<const uint2*>(tbl_b1 + (n0 >> 1) + 2 * j4);
            // dot over all 257 inputs -> dtype, + bias -> dtype, ReLU (packed bf16 after the first rounding)
            uint32_t p0 = pack_bf16(__fmaf_rn(sm, w4.x, __uint_as_float(vv[j4 * 4 + 0])),
                                    __fmaf_rn(sm, w4.y, __uint_as_float(vv[j4 * 4 + 1])));
            uint32_t p1 = pack_bf16(__fmaf_rn(sm, w4.z, __uint_as_float(vv[j4 * 4 + 2])),
                                    __fmaf_rn(sm, w4.w, __uint_as_float(vv[j4 * 4 + 3])));
            p0 = hadd2_bf16_rn(p0, b2p.x);
            p1 = hadd2_bf16_rn(p1, b2p.y);
            h[j4 * 2 + 0] = hmax2_bf16(p0, 0u);
            h[j4 * 2 + 1] = hmax2_bf16(p1, 0u);
          }
          const int kc = c16 >> 2, s0 = (c16 & 3) * 2;  // K-chunk of H, 16-byte slot of column n0 inside its 128-byte row
          uint8_t* rowp = Ab + kc * 16384 + row * 128;
          *reinterpret_cast<uint4*>(rowp + ((s0 ^ (row & 7)) * 16)) = make_uint4(h[0], h[1], h[2], h[3]);
          *reinterpret_cast<uint4*>(rowp + (((s0 + 1) ^ (row & 7)) * 16)) = make_uint4(h[4], h[5], h[6], h[7]);
          if ((c16 & 3) == 3) {  // K-chunk kc of this warp's rows is complete
            fence_proxy_async_smem();
            tc_fence_before_sync();
            mbar_arrive(&ctl->h_full[kc]);
          }
        }
      }
      F2_MARK(tprof, 1);
      // ---------- epilogue 2: volume rows = bf16(bf16(acc2) + b2) -> smem staging (same tile buffer) ----------
      mbar_wait_sleep(&ctl->acc2_full, tile & 1, 20);
      tc_fence_after_sync();
      F2_MARK(tprof, 2);
      {
        const uint32_t taddr = tmem_acc2 + ((uint32_t)(q * 32) << 16);
        uint32_t v[2][16];
        tmem_ld16(taddr, v[0]);
#pragma unroll
        for (int c16 = 0; c16 < 8; ++c16) {
          tmem_ld_wait();
          if (c16 + 1 < 8) tmem_ld16(taddr + (uint32_t)((c16 + 1) * 16), v[(c16 + 1) & 1]);
          const uint32_t(&vv)[16] = v[c16 & 1];
          const int n0 = c16 * 16;
          const uint4 ba = *reinterpret_cast<const uint4*>(tbl_b2 + (n0 >> 1));
          const uint4 bb = *reinterpret_cast<const uint4*>(tbl_b2 + (n0 >> 1) + 4);
          const uint32_t bp[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
          uint32_t o[8];
#pragma unroll
          for (int j = 0; j < 8; ++j)
            o[j] = hadd2_bf16_rn(pack_bf16(__uint_as_float(vv[2 * j]), __uint_as_float(vv[2 * j + 1])), bp[j]);
          uint8_t* rowp = Ab + row * VOL2_STRIDE + n0 * 2;
          *reinterpret_cast<uint4*>(rowp) = make_uint4(o[0], o[1], o[2], o[3]);
          *reinterpret_cast<uint4*>(rowp + 16) = make_uint4(o[4], o[5], o[6], o[7]);
        }
      }
      tc_fence_before_sync();
      __syncwarp();
      F2_MARK(tprof, 3);
      // ---------- vertical max (bev_mapper.py:56-88) ----------
      // Warp q scans its own 32 staged rows (lane = 4 channels, packed bf16x2 max).  Rows are sorted by BEV column:
      // segments that start and end strictly inside the warp's rows are complete and written directly; the first /
      // last segment go to shared memory and are stitched (with the carry from the previous tile) by warp 0.
      {
        const int r_lo = q * 32;
        const int nr = min(max(rows - r_lo, 0), 32);            // this warp's rows
        const uint32_t mycol = lane < nr ? rowcol[b * 128 + r_lo + lane] : 0xffffffffu;
        const uint32_t prevcol = __shfl_up_sync(0xffffffffu, mycol, 1);
        uint32_t starts = __ballot_sync(0xffffffffu, lane < nr && (lane == 0 || mycol != prevcol));  // segment starts
        uint32_t fcol = 0xffffffffu, lcol = 0xffffffffu;
        uint2 fmax_ = make_uint2(0u, 0u), lmax_ = make_uint2(0u, 0u);
        const uint8_t* base = Ab + r_lo * VOL2_STRIDE + lane * 8;
        while (starts != 0u) {   // warp-uniform
          const int s0 = __ffs(starts) - 1;
          starts &= starts - 1u;
          const int s1 = starts != 0u ? __ffs(starts) - 1 : nr;
          const uint32_t col = __shfl_sync(0xffffffffu, mycol, s0);
          uint2 acc = *reinterpret_cast<const uint2*>(base + s0 * VOL2_STRIDE);
          int r = s0 + 1;
          for (; r + 3 < s1; r += 4) {   // independent loads, max tree
            const uint2 x0 = *reinterpret_cast<const uint2*>(base + r * VOL2_STRIDE);
            const uint2 x1 = *reinterpret_cast<const uint2*>(base + (r + 1) * VOL2_STRIDE);
            const uint2 x2 = *reinterpret_cast<const uint2*>(base + (r + 2) * VOL2_STRIDE);
            const uint2 x3 = *reinterpret_cast<const uint2*>(base + (r + 3) * VOL2_STRIDE);
            acc = hmax2x2(acc, hmax2x2(hmax2x2(x0, x1), hmax2x2(x2, x3)));
          }
          for (; r < s1; ++r) acc = hmax2x2(acc, *reinterpret_cast<const uint2*>(base + r * VOL2_STRIDE));
          const bool first = s0 == 0, last = starts == 0u;
          if (first) {
            fcol = col;
            fmax_ = acc;
          }
          if (last) {
            lcol = col;
            lmax_ = acc;
          }
          if (!first && !last) {  // a segment that starts and ends inside this warp's rows: complete
            plane64[(size_t)col * 32 + lane] = acc;
            if (lane == 0) A.pvalid[col] = 1;
          }
        }
        uint4* brA = reinterpret_cast<uint4*>(Ab + S2_BREC);
        uint2* brB = reinterpret_cast<uint2*>(Ab + S2_BREC + 2048);
        brA[q * 32 + lane] = make_uint4(fcol, fmax_.x, fmax_.y, lcol);
        brB[q * 32 + lane] = lmax_;
      }
      cbar();
      if (q == 0) {
        const uint4* brA = reinterpret_cast<const uint4*>(Ab + S2_BREC);
        const uint2* brB = reinterpret_cast<const uint2*>(Ab + S2_BREC + 2048);
#pragma unroll
        for (int part = 0; part < 4; ++part) {
          const uint4 a4 = brA[part * 32 + lane];
          const uint2 l2 = brB[part * 32 + lane];
          const uint32_t fcol = a4.x, lcol = a4.w;
          if (fcol == 0xffffffffu) continue;  // part without rows
          const uint2 f2 = make_uint2(a4.y, a4.z);
          if (fcol == zm_col) {
            zm_val = hmax2x2(zm_val, f2);
          } else {
            if (zm_col != 0xffffffffu) {
              plane64[(size_t)zm_col * 32 + lane] = zm_val;
              if (lane == 0) A.pvalid[zm_col] = 1;
            }
            zm_col = fcol;
            zm_val = f2;
          }
          if (lcol != fcol) {  // the first segment ended inside this part; the last one stays open
            plane64[(size_t)zm_col * 32 + lane] = zm_val;
            if (lane == 0) A.pvalid[zm_col] = 1;
            zm_col = lcol;
            zm_val = l2;
          }
        }
      }
      mbar_arrive(&ctl->a_empty[b]);   // tile buffer b (staging rows, boundary records) is free again
      F2_MARK(tprof, 4);
    }
    if (q == 0 && zm_col != 0xffffffffu) {
      plane64[(size_t)zm_col * 32 + lane] = zm_val;
      if (lane == 0) A.pvalid[zm_col] = 1;
    }
    if (ctid == 0) {
#pragma unroll
      for (int i = 0; i < 5; ++i) atomicAdd(A.col_counter + 7 + i, (int)(tprof[i] >> 4));
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == F2_W_MMA) {
    __syncwarp();
    tmem_dealloc(tmem_base, 512);
  }
}


static int launch(const SnapLiftParams* q, int B, const SnapLiftView* views, long long views_stride,
                                           const void* fimg, long long fimg_stride, const float* xs, const float* ys,
                                           const float* zs, long long zs_stride, const void* w1t, long long ldw1,
                                           const float* w256, const float* b1, const void* w2t, const float* b2,
                                           void* plane, uint8_t* pvalid, int* col_counter, void* scratch,
                                           size_t scratch_bytes, void* stream) {
  SNAP_REQUIRE(q && views && fimg && xs && ys && zs && w1t && w256 && b1 && w2t && b2 && plane && pvalid && col_counter,
               "null pointer");
  SNAP_REQUIRE(B >= 1, "empty batch");
  SNAP_REQUIRE(q->V >= 1 && q->V <= F2_MAXV, "fused lift handles 1..%d views per scene (got %d)", F2_MAXV, q->V);
  SNAP_REQUIRE(B * q->V <= F2_MAX_VIEWS_TOTAL, "fused lift handles B * V <= %d views per launch (got %d x %d)",
               F2_MAX_VIEWS_TOTAL, B, q->V);
  SNAP_REQUIRE(q->D == 128 && q->S >= 2 && q->CF == q->D + q->S && q->CF % 8 == 0, "bad channel split");
  SNAP_REQUIRE(!q->no_variance && !q->add_minmax, "the fused lift implements the default statistics only");
  SNAP_REQUIRE(q->Z >= 1 && q->Z <= 64, "Z must be <= 64 (got %d)", q->Z);
  SNAP_REQUIRE(fimg_stride >= (long long)q->V * q->Hf * q->Wf * q->CF || B == 1, "fimg_stride smaller than one scene");
  SNAP_REQUIRE((long long)(B - 1) * fimg_stride + (long long)q->V * q->Hf * q->Wf * q->CF < (1LL << 31),
               "feature maps too large for 32-bit tap offsets");
  SNAP_REQUIRE((long long)B * q->X * q->Y <= (1LL << 18), "too many BEV columns in one launch (B * X * Y <= 262144)");
  SNAP_REQUIRE(fimg_stride % 8 == 0 && (reinterpret_cast<uintptr_t>(fimg) & 15) == 0, "fimg must be 16-byte aligned per scene");
  static_assert(sizeof(SnapLiftView) == sizeof(LiftView), "SnapLiftView layout");
  static_assert(sizeof(SnapLiftParams) == sizeof(LiftParams), "SnapLiftParams layout");
  cudaStream_t s = (cudaStream_t)stream;
  static DynSmemState smem_state;
  if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(&lift_fused2_kernel), F2_SMEM_BYTES, &smem_state,
                               "cudaFuncSetAttribute(lift_fused2)"))
    return rc;
  CUtensorMap tmW1, tmW2;
  int rc = make_tmap_2d_bf16(&tmW1, w1t, 256, 256, ldw1, 128, 64);
  if (rc) return rc;
  rc = make_tmap_2d_bf16(&tmW2, w2t, 128, 256, 256, 128, 64);
  if (rc) return rc;
  const long long cells = (long long)B * q->X * q->Y;
  rc = check_cuda(cudaMemsetAsync(plane, 0, (size_t)cells * 128 * 2, s), "memset plane");
  if (rc) return rc;
  rc = check_cuda(cudaMemsetAsync(pvalid, 0, (size_t)cells, s), "memset valid");
  if (rc) return rc;
  rc = check_cuda(cudaMemsetAsync(col_counter, 0, 16 * sizeof(int), s), "memset counter");
  if (rc) return rc;
  Fused2Args a;
  memcpy(&a.P, q, sizeof(LiftParams));
  a.B = B;
  a.views_stride = views_stride;
  a.fimg_stride = fimg_stride;
  a.zs_stride = zs_stride;
  a.views = reinterpret_cast<const LiftView*>(views);
  a.fimg = (const __nv_bfloat16*)fimg;
  a.xs = xs;
  a.ys = ys;
  a.zs = zs;
  a.w256 = w256;
  a.b1 = b1;
  a.b2 = b2;
  a.plane = (__nv_bfloat16*)plane;
  a.pvalid = pvalid;
  a.col_counter = col_counter;
  int grid = num_sms();
  const int max_useful = (int)((cells + F2_BATCH_COLS - 1) / F2_BATCH_COLS);
  if (grid > max_useful) grid = max_useful;
  SNAP_REQUIRE(scratch != nullptr && scratch_bytes >= (size_t)grid * F2_LIST_CAP * F2_MAXV * sizeof(TapRec2),
               "scratch too small: need snapb200_lift_fused_batched_scratch_bytes()");
  a.scratch = reinterpret_cast<TapRec2*>(scratch);
  lift_fused2_kernel<<<grid, F2_THREADS, F2_SMEM_BYTES, s>>>(tmW1, tmW2, a);
  return check_launch("lift_fused2_kernel");
}

static size_t scratch_bytes_needed() { return (size_t)num_sms() * F2_LIST_CAP * F2_MAXV * sizeof(TapRec2); }

}  // namespace F2_NS
}  // namespace snapb200
