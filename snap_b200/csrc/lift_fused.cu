// Fused camera->BEV lift as ONE persistent sm_100a kernel:
//
//   voxel visibility (bit-exact projection)  ->  compaction of the visible voxels into 128-row tiles
//   -> bilinear gather + depth score + multi-view softmax pooling -> statistics tile A[128 x 256] in
//   swizzled shared memory -> tcgen05 GEMM1 (x W1, resident in smem, + rank-1 term of the score column)
//   -> bias / ReLU epilogue writes H[128 x 256] back into the same smem tile -> tcgen05 GEMM2 (x W2,
//   streamed by TMA through a 2-slot ring) -> bias epilogue -> running max over z per BEV column
//   -> plane[G*G, 128] + valid[G*G].
//
// Nothing but the projected feature maps is read from HBM/L2 and nothing but the BEV plane is written:
// the [N,V,160] gather result, the [N,257] statistics, the [N,256] hidden layer and the [N,128] feature
// volume of the reference (streetview_encoder.py:251-286, bev_mapper.py:56-88) never exist in memory.
// Voxels no camera sees are defined as zero / invalid by the reference (:282) and are skipped before the
// MLP ("valid-voxel compaction"); results are identical.
//
// Warp roles (18 warps): 0 = TMA producer (W1 once, W2 ring), 1 = TMEM owner + MMA issuer,
// 2..17 = 16 worker warps (visibility pass, gather/pool, both epilogues, z-max).
#include <cuda_bf16.h>
#include <math.h>
#include <string.h>

#include "common.cuh"
#include "host_common.h"
#include "lift_common.cuh"

namespace snapb200 {

constexpr int FL_WORKERS = 16;                       // worker warps
constexpr int FL_THREADS = (2 + FL_WORKERS) * 32;    // 576
constexpr int FL_BATCH_COLS = 4;                     // BEV columns per visibility batch (4 x 64 z-slots)
constexpr int FL_LIST_CAP = 512;                     // ring of compacted visible voxels (>= 128 in flight + 127 + 256)
static_assert(FL_BATCH_COLS * 64 * 2 == FL_WORKERS * 32, "visibility pass: two worker threads per voxel slot");
constexpr int FL_MAXV = 4;                           // views handled by the fused kernel

// shared memory map (bytes); every MMA operand region is 1024-aligned
constexpr int SM_W1 = 0;                  // 4 K-chunks x [256 x 64] bf16, 128B-swizzled   (131072)
constexpr int SM_AH = 131072;             // A / H / volume staging: 4 x [128 x 64] bf16   ( 65536)
constexpr int SM_W2 = 196608;             // ring: 2 x [128 x 64] bf16                      ( 32768)
constexpr int SM_LIST = 229376;           // uint32[FL_LIST_CAP]                            (  2048)
constexpr int SM_SMAX = SM_LIST + 4 * FL_LIST_CAP;   // bf16[128] score_max per tile row    (   256)
constexpr int SM_BAR = SM_SMAX + 256;     // mbarriers + control words                      (   256)
constexpr int SM_VIEW = SM_BAR + 256;     // LiftView[FL_MAXV]                              (   384)
constexpr int FL_SMEM_BYTES = SM_VIEW + 384;
static_assert(sizeof(LiftView) * FL_MAXV <= 384, "view table");
static_assert(FL_SMEM_BYTES <= 232448, "shared memory budget (227 KB)");
constexpr int VOL_STRIDE = 272;           // bytes per staged volume row (256 + 16: conflict-free)

// Per (visible voxel, view) gather record, written by the thread-per-voxel visibility pass and read by the
// half-warp-per-voxel gather pass (global scratch, L2 resident): tap offsets (so that the gather needs one
// 64-bit add per tap instead of the row/column/view address arithmetic), upper-tap weights, depth-score bins.  32 bytes.
struct TapRec {
  uint32_t off00, off01, off10, off11;  // element offsets of the four taps into fimg (view base included)
  float wr1, wc1;                       // weights of the upper taps
  float wb1;                            // weight of the upper depth bin
  uint16_t b0, b1;                      // depth bins
};
static_assert(sizeof(TapRec) == 32, "TapRec layout");

struct FusedArgs {
  LiftParams P;
  const LiftView* views;
  const __nv_bfloat16* fimg;
  const float *xs, *ys, *zs;
  const float* w256;  // W1[256, :] (score_max input row) as f32[256]
  const float* b1;    // f32[256]
  const float* b2;    // f32[128]
  __nv_bfloat16* plane;
  uint8_t* pvalid;
  int* col_counter;   // zeroed by the host wrapper
  TapRec* scratch;    // [gridDim.x][FL_LIST_CAP][FL_MAXV]
};

struct Ctl {  // lives at SM_BAR
  uint64_t w1_full, w2_full[2], w2_empty[2], a_full, acc1_full, h_full[4], acc2_full;  // 12 x 8 B
  uint32_t tmem_ptr;
  int batch_col0, more;
  int warp_cnt[FL_WORKERS];
};
static_assert(sizeof(Ctl) <= 256, "control block");

__device__ __forceinline__ void worker_bar() {  // named barrier 1 over the 16 worker warps
  asm volatile("bar.sync 1, %0;" ::"n"(FL_WORKERS * 32) : "memory");
}

__global__ void __launch_bounds__(FL_THREADS, 1)
lift_fused_kernel(const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmW2,
                  const __grid_constant__ FusedArgs A) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const LiftParams& P = A.P;
  Ctl* ctl = reinterpret_cast<Ctl*>(smem + SM_BAR);
  uint32_t* list = reinterpret_cast<uint32_t*>(smem + SM_LIST);
  __nv_bfloat16* smax_s = reinterpret_cast<__nv_bfloat16*>(smem + SM_SMAX);
  LiftView* sview = reinterpret_cast<LiftView*>(smem + SM_VIEW);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if ((smem_u32(smem) & 1023u) != 0) __trap();  // swizzled operands need a 1024 B aligned base

  for (int i = threadIdx.x; i < P.V * (int)(sizeof(LiftView) / 4); i += blockDim.x)
    reinterpret_cast<uint32_t*>(sview)[i] = reinterpret_cast<const uint32_t*>(A.views)[i];
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmW1);
    tma_prefetch_desc(&tmW2);
    mbar_init(&ctl->w1_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&ctl->w2_full[s], 1);
      mbar_init(&ctl->w2_empty[s], 1);
    }
    mbar_init(&ctl->a_full, FL_WORKERS * 32);
    mbar_init(&ctl->acc1_full, 1);
    for (int kc = 0; kc < 4; ++kc) mbar_init(&ctl->h_full[kc], FL_WORKERS * 32);
    mbar_init(&ctl->acc2_full, 1);
    ctl->more = 0;
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&ctl->tmem_ptr, 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = ctl->tmem_ptr;
  const uint32_t tmem_acc1 = tmem_base;        // 256 columns
  const uint32_t tmem_acc2 = tmem_base + 256;  // 128 columns

  if (warp == 0) {
    // ============================ TMA producer ============================
    if (elect_one()) {
      mbar_arrive_expect_tx(&ctl->w1_full, 4 * 32768);
      for (int kc = 0; kc < 4; ++kc) tma_load_2d(&tmW1, &ctl->w1_full, smem + SM_W1 + kc * 32768, kc * 64, 0);
      uint32_t tile_phase = 0;
      int slot = 0;
      uint32_t ring_phase = 0;
      while (true) {
        mbar_wait_sleep(&ctl->a_full, tile_phase, 256);
        tile_phase ^= 1;
        if (*reinterpret_cast<volatile int*>(&ctl->more) == 0) break;
        for (int kc = 0; kc < 4; ++kc) {
          mbar_wait_sleep(&ctl->w2_empty[slot], ring_phase ^ 1, 64);
          mbar_arrive_expect_tx(&ctl->w2_full[slot], 16384);
          tma_load_2d(&tmW2, &ctl->w2_full[slot], smem + SM_W2 + slot * 16384, kc * 64, 0);
          if (++slot == 2) {
            slot = 0;
            ring_phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ============================ MMA issuer ============================
    if (elect_one()) {
      constexpr uint32_t idesc1 = make_idesc_bf16_m128(256);
      constexpr uint32_t idesc2 = make_idesc_bf16_m128(128);
      mbar_wait(&ctl->w1_full, 0);
      uint32_t tile_phase = 0;
      int slot = 0;
      uint32_t ring_phase = 0;
      const uint32_t sAH = smem_u32(smem + SM_AH), sW1 = smem_u32(smem + SM_W1), sW2 = smem_u32(smem + SM_W2);
      while (true) {
        mbar_wait_sleep(&ctl->a_full, tile_phase, 128);
        if (*reinterpret_cast<volatile int*>(&ctl->more) == 0) break;
        tc_fence_after_sync();
        // GEMM1: acc1[128 x 256] = A[128 x 256] * W1^T
#pragma unroll
        for (int kc = 0; kc < 4; ++kc) {
          const uint64_t da = make_kmajor_desc<128>(sAH + kc * 16384);
          const uint64_t db = make_kmajor_desc<128>(sW1 + kc * 32768);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem_acc1, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc1, (kc | k) != 0 ? 1u : 0u);
        }
        umma_commit(&ctl->acc1_full);
        // GEMM2: acc2[128 x 128] = H[128 x 256] * W2^T  (H written by the workers into the A tile, one K-chunk
        // at a time: the MMAs of chunk kc run while the workers are still writing chunks kc+1..3)
        for (int kc = 0; kc < 4; ++kc) {
          mbar_wait_sleep(&ctl->h_full[kc], tile_phase, 32);
          mbar_wait(&ctl->w2_full[slot], ring_phase);
          tc_fence_after_sync();
          const uint64_t da = make_kmajor_desc<128>(sAH + kc * 16384);
          const uint64_t db = make_kmajor_desc<128>(sW2 + slot * 16384);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem_acc2, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc2, (kc | k) != 0 ? 1u : 0u);
          umma_commit(&ctl->w2_empty[slot]);
          if (++slot == 2) {
            slot = 0;
            ring_phase ^= 1;
          }
        }
        umma_commit(&ctl->acc2_full);
        tile_phase ^= 1;
      }
    }
  } else {
    // ============================ worker warps ============================
    const int ww = warp - 2;             // 0..15
    const int wtid = threadIdx.x - 64;   // 0..511
    const int q = warp & 3;              // TMEM lane quadrant this warp may touch
    const int sub = ww >> 2;             // which quarter of the accumulator columns it drains
    const int ncols = P.X * P.Y;
    const int half = lane >> 4;
    const int l16 = lane & 15;           // channels [8*l16, 8*l16+8) of the half-warp's row
    const float score_scale = (float)(P.S - 1);
    const unsigned FULL = 0xffffffffu;
    TapRec* const my_scratch = A.scratch + (size_t)blockIdx.x * FL_LIST_CAP * FL_MAXV;
    // z-max scratch inside the A/H region (free between the end of GEMM2 and the next gather)
    uint16_t* const rowcol = reinterpret_cast<uint16_t*>(smem + SM_AH + 128 * VOL_STRIDE);        // BEV column of row r
    uint4* const brec = reinterpret_cast<uint4*>(smem + SM_AH + 128 * VOL_STRIDE + 256);          // [8 parts][64]
    uint32_t tile_phase = 0;
    int n_tiles = 0, n_rows = 0;  // work counters (reported through col_counter[1..2])
    // ring state: replicated in the registers of every worker thread (all of them take the same decisions)
    int list_head = 0, list_count = 0;
    bool cols_done = false;
    unsigned tprof[7] = {0, 0, 0, 0, 0, 0, 0};  // cycles: fill, gather, wait acc1, epi1, wait acc2, epi2, z-max
    unsigned tmark = (unsigned)clock();
#define LIFT_MARK(i)                        \
  do {                                      \
    const unsigned now_ = (unsigned)clock(); \
    tprof[i] += now_ - tmark;               \
    tmark = now_;                           \
  } while (0)
    int zm_col = -1;        // z-max carry of thread wtid < 64 (two output channels each, packed bf16x2)
    uint32_t zm_val = 0;

    if (wtid == 0) ctl->batch_col0 = atomicAdd(A.col_counter, FL_BATCH_COLS);  // first batch claim

    // Software pipeline: the visibility batches that top up the list for tile i+1 run while the tensor core computes
    // GEMM1 of tile i (whose rows stay in the ring until its z-max is done); then the epilogues of tile i, then the
    // gather of tile i+1.
    int rows = 0, head = 0;  // tile in flight (gathered, GEMM1 issued); rows == 0: none
    while (true) {
      // ---------- visibility batches: >= 128 visible voxels beyond the tile in flight ----------
      // Two threads per voxel (thread s of the pair projects views 2s, 2s+1 and writes their gather records);
      // the next batch is claimed from the global counter while the current one is being processed.
      while (list_count - rows < 128 && !cols_done) {
        worker_bar();  // batch_col0 of this round visible; previous readers of warp_cnt are done
        const int c0 = ctl->batch_col0;
        if (c0 >= ncols) {
          cols_done = true;
          break;
        }
        const int vox = wtid >> 1, s2 = wtid & 1;
        const int cl = vox >> 6, z = vox & 63;
        const int col = c0 + cl;
        Proj pr[2];
        uint32_t vm = 0;
        if (col < ncols && z < P.Z) {
          const int ix = col / P.Y, iy = col - ix * P.Y;
          const float px = A.xs[P.xy_paired ? col : ix], py = A.ys[P.xy_paired ? col : iy], pz = A.zs[z];
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const int v = 2 * s2 + k;
            if (v < P.V) {
              pr[k] = project_point(sview[v], px, py, pz);
              if (pr[k].vis) vm |= 1u << v;
            }
          }
        }
        vm |= __shfl_xor_sync(FULL, vm, 1);
        const bool valid = vm != 0;
        const uint32_t bal = __ballot_sync(FULL, valid) & 0x55555555u;  // one bit per voxel (its even lane)
        if (lane == 0) ctl->warp_cnt[ww] = __popc(bal);
        worker_bar();
        if (wtid == 0) ctl->batch_col0 = atomicAdd(A.col_counter, FL_BATCH_COLS);  // everyone has read c0
        int before = 0, tot = 0;
#pragma unroll
        for (int w2 = 0; w2 < FL_WORKERS; ++w2) {
          const int c = ctl->warp_cnt[w2];
          tot += c;
          if (w2 < ww) before += c;
        }
        if (valid) {
          const int slot = (list_head + list_count + before + __popc(bal & ((1u << (lane & 30)) - 1u))) & (FL_LIST_CAP - 1);
          if (s2 == 0) list[slot] = ((uint32_t)col << 14) | ((uint32_t)z << 8) | vm;
          // gather records of this thread's visible views (same tap / bin arithmetic as the unfused kernel)
          TapRec* rec = my_scratch + (size_t)slot * FL_MAXV;
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const int v = 2 * s2 + k;
            if (!((vm >> v) & 1u)) continue;
            const Taps t = make_taps(pr[k].row, pr[k].col, P.Hf, P.Wf);
            const float d = fminf(fmaxf(pr[k].depth, P.depth_min), P.depth_max);
            const float bi = logf(d / P.depth_min) * P.inv_log_range * score_scale;
            const float bf = floorf(bi);
            const uint32_t b0 = (uint32_t)min(max((int)bf, 0), P.S - 1);
            const uint32_t b1 = (uint32_t)min(max((int)bf + 1, 0), P.S - 1);
            const uint32_t row0 = (uint32_t)((v * P.Hf + t.r0) * P.Wf), row1 = (uint32_t)((v * P.Hf + t.r1) * P.Wf);
            uint4* dst = reinterpret_cast<uint4*>(rec + v);
            dst[0] = make_uint4((row0 + (uint32_t)t.c0) * (uint32_t)P.CF, (row0 + (uint32_t)t.c1) * (uint32_t)P.CF,
                                (row1 + (uint32_t)t.c0) * (uint32_t)P.CF, (row1 + (uint32_t)t.c1) * (uint32_t)P.CF);
            dst[1] = make_uint4(__float_as_uint(t.wr1), __float_as_uint(t.wc1), __float_as_uint(bi - bf), b0 | (b1 << 16));
          }
        }
        list_count += tot;
      }
      LIFT_MARK(0);

      if (rows > 0) {
      // ---------- epilogue 1: H = relu(bf16(bf16(acc1 + smax * w256) + b1)) -> smem (A tile) ----------
      // Packed bf16 arithmetic after the first rounding: HADD2.BF16 of two bf16 values is the exactly rounded sum,
      // i.e. identical to the fp32 add + round of the reference's "+ bias -> dtype"; biases are bf16 parameters
      // (flax param_dtype = dtype), an fp32 bias is rounded to bf16 first.
      mbar_wait(&ctl->acc1_full, tile_phase);
      tc_fence_after_sync();
      LIFT_MARK(2);
      {
        // warp (q, sub) owns rows 32q..32q+31 and, of every 64-column K-chunk c of H, columns 16 sub..16 sub+15: chunk c
        // is complete (and handed to the tensor core) after the c-th step of all 16 warps
        const int row = q * 32 + lane;
        const float sm = __bfloat162float(smax_s[row]);
        const uint32_t taddr = tmem_acc1 + ((uint32_t)(q * 32) << 16) + (uint32_t)(sub * 16);
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t v[16];
          tmem_ld16(taddr + (uint32_t)(c * 64), v);
          tmem_ld_wait();
          const int n0 = c * 64 + sub * 16;
          uint32_t h[8];
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            const float4 w4 = __ldg(reinterpret_cast<const float4*>(A.w256 + n0) + j4);
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(A.b1 + n0) + j4);
            // dot over all 257 inputs -> dtype, + bias -> dtype, ReLU
            uint32_t p0 = pack_bf16(__fmaf_rn(sm, w4.x, __uint_as_float(v[j4 * 4 + 0])),
                                    __fmaf_rn(sm, w4.y, __uint_as_float(v[j4 * 4 + 1])));
            uint32_t p1 = pack_bf16(__fmaf_rn(sm, w4.z, __uint_as_float(v[j4 * 4 + 2])),
                                    __fmaf_rn(sm, w4.w, __uint_as_float(v[j4 * 4 + 3])));
            p0 = hadd2_bf16_rn(p0, pack_bf16(b4.x, b4.y));
            p1 = hadd2_bf16_rn(p1, pack_bf16(b4.z, b4.w));
            h[j4 * 2 + 0] = hmax2_bf16(p0, 0u);
            h[j4 * 2 + 1] = hmax2_bf16(p1, 0u);
          }
          const int s0 = sub * 2;  // 16-byte slot of column n0 inside the 128-byte row of chunk c
          uint8_t* rowp = smem + SM_AH + c * 16384 + row * 128;
          *reinterpret_cast<uint4*>(rowp + ((s0 ^ (row & 7)) * 16)) = make_uint4(h[0], h[1], h[2], h[3]);
          *reinterpret_cast<uint4*>(rowp + (((s0 + 1) ^ (row & 7)) * 16)) = make_uint4(h[4], h[5], h[6], h[7]);
          fence_proxy_async_smem();
          tc_fence_before_sync();
          mbar_arrive(&ctl->h_full[c]);
        }
      }
      LIFT_MARK(3);

      // ---------- epilogue 2: volume rows = bf16(bf16(acc2) + b2) -> smem staging ----------
      mbar_wait(&ctl->acc2_full, tile_phase);
      tc_fence_after_sync();
      LIFT_MARK(4);
      {
        const int row = q * 32 + lane;
        const uint32_t taddr = tmem_acc2 + ((uint32_t)(q * 32) << 16) + (uint32_t)(sub * 32);
#pragma unroll 1
        for (int c16 = 0; c16 < 2; ++c16) {
          uint32_t v[16];
          tmem_ld16(taddr + (uint32_t)(c16 * 16), v);
          tmem_ld_wait();
          const int n0 = sub * 32 + c16 * 16;
          uint32_t o[8];
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(A.b2 + n0) + j4);
            o[j4 * 2 + 0] = hadd2_bf16_rn(pack_bf16(__uint_as_float(v[j4 * 4 + 0]), __uint_as_float(v[j4 * 4 + 1])),
                                          pack_bf16(b4.x, b4.y));
            o[j4 * 2 + 1] = hadd2_bf16_rn(pack_bf16(__uint_as_float(v[j4 * 4 + 2]), __uint_as_float(v[j4 * 4 + 3])),
                                          pack_bf16(b4.z, b4.w));
          }
          uint8_t* rowp = smem + SM_AH + row * VOL_STRIDE + n0 * 2;
          *reinterpret_cast<uint4*>(rowp) = make_uint4(o[0], o[1], o[2], o[3]);
          *reinterpret_cast<uint4*>(rowp + 16) = make_uint4(o[4], o[5], o[6], o[7]);
        }
        if (sub == 0 && row < rows) {  // BEV column of every tile row, for the z-max scan
          int slot = head + row;
          if (slot >= FL_LIST_CAP) slot -= FL_LIST_CAP;
          rowcol[row] = (uint16_t)(list[slot] >> 14);
        }
      }
      tc_fence_before_sync();
      worker_bar();
      LIFT_MARK(5);

      // ---------- vertical max (bev_mapper.py:56-88) ----------
      // 8 row parts x 64 channel pairs: thread (part, c2) scans 16 rows with packed bf16x2 max.  Column segments that
      // start and end strictly inside a part are complete and written directly; the first / last segment of each part
      // go to shared memory and are stitched (with the carry from the previous tile) by the 64 threads of part 0.
      {
        const int part = wtid >> 6, c2 = wtid & 63;
        const int r_lo = part * 16, r_hi = min(rows, r_lo + 16);
        int fcol = -1, lcol = -1;   // first / last column of this part
        uint32_t fmax_ = 0u, lmax_ = 0u;
        uint32_t* plane32 = reinterpret_cast<uint32_t*>(A.plane);
        const bool one_col = r_lo < r_hi && rowcol[r_lo] == rowcol[r_hi - 1];  // rows are sorted by column
        if (one_col) {  // the common case (a column has up to Z rows): a plain packed max over the part
          const uint8_t* base = smem + SM_AH + c2 * 4;
          uint32_t m = *reinterpret_cast<const uint32_t*>(base + r_lo * VOL_STRIDE);
          for (int r = r_lo + 1; r < r_hi; ++r)
            m = hmax2_bf16(m, *reinterpret_cast<const uint32_t*>(base + r * VOL_STRIDE));
          fcol = lcol = rowcol[r_lo];
          fmax_ = lmax_ = m;
        }
        for (int r = r_lo; r < (one_col ? r_lo : r_hi); ++r) {
          const int col = rowcol[r];
          const uint32_t x = *reinterpret_cast<const uint32_t*>(smem + SM_AH + r * VOL_STRIDE + c2 * 4);
          if (col != lcol) {
            if (lcol >= 0 && lcol != fcol) {  // a middle segment just ended: complete
              plane32[(size_t)lcol * 64 + c2] = lmax_;
              if (c2 == 0) A.pvalid[lcol] = 1;
            }
            if (fcol < 0) fcol = col;
            lcol = col;
            lmax_ = x;
          } else {
            lmax_ = hmax2_bf16(lmax_, x);
          }
          if (lcol == fcol) fmax_ = lmax_;
        }
        brec[part * 64 + c2] = make_uint4((uint32_t)fcol, fmax_, (uint32_t)lcol, lmax_);
      }
      worker_bar();
      if (wtid < 64) {
        uint32_t* plane32 = reinterpret_cast<uint32_t*>(A.plane);
#pragma unroll
        for (int part = 0; part < 8; ++part) {
          const uint4 b4 = brec[part * 64 + wtid];
          const int fcol = (int)b4.x, lcol = (int)b4.z;
          if (fcol < 0) continue;  // part without rows
          if (fcol == zm_col) {
            zm_val = hmax2_bf16(zm_val, b4.y);
          } else {
            if (zm_col >= 0) {
              plane32[(size_t)zm_col * 64 + wtid] = zm_val;
              if (wtid == 0) A.pvalid[zm_col] = 1;
            }
            zm_col = fcol;
            zm_val = b4.y;
          }
          if (lcol != fcol) {  // the first segment ended inside this part; the last one stays open
            plane32[(size_t)zm_col * 64 + wtid] = zm_val;
            if (wtid == 0) A.pvalid[zm_col] = 1;
            zm_col = lcol;
            zm_val = b4.w;
          }
        }
      }
      list_head = (head + rows) & (FL_LIST_CAP - 1);   // release the rows of the finished tile
      list_count -= rows;
      n_tiles += 1;
      n_rows += rows;
      tile_phase ^= 1;
      LIFT_MARK(6);
      }  // rows > 0: epilogues of the tile in flight

      // ---------- next tile ----------
      worker_bar();  // list entries / records of this round's batches visible; z-max scratch of the last tile is dead
      rows = min(128, list_count);
      head = list_head;
      if (wtid == 0) ctl->more = rows > 0 ? 1 : 0;  // published by this thread's arrive (release) on a_full
      if (rows == 0) {
        mbar_arrive(&ctl->a_full);
        break;
      }
      // ---------- gather + pool: one HALF-warp per tile row (16 lanes x 8 channels) ----------
      // Record prefetch: a warp handles rows {2 ww + 32 it + half}, it = 0..3.  Lane 4 it + k of each half loads the
      // gather record of the k-th visible view of the half's row of iteration it, so all record reads of the tile are
      // in flight at once (one exposed L2 round trip instead of one per row and view); the loop below fetches the words
      // with shuffles.  Views are processed by RANK (k-th visible view of the row), not by view index: the two halves
      // of a warp then need max(popc) iterations instead of popc(union), and the pooling is symmetric in the views.
      uint32_t pvm = 0;
      uint4 pq0 = make_uint4(0u, 0u, 0u, 0u), pq1 = make_uint4(0u, 0u, 0u, 0u);
      {
        const int pit = l16 >> 2, pk = l16 & 3;
        const int prow = ww * 2 + 32 * pit + half;
        if (prow < rows) {
          const int slot = (head + prow) & (FL_LIST_CAP - 1);
          pvm = list[slot] & 0xfu;
          uint32_t m = pvm;  // drop the pk lowest set bits
          if (pk > 0) m &= m - 1u;
          if (pk > 1) m &= m - 1u;
          if (pk > 2) m &= m - 1u;
          if (m != 0u) {
            const TapRec* rec = my_scratch + (size_t)slot * FL_MAXV + (__ffs(m) - 1);
            pq0 = __ldcg(reinterpret_cast<const uint4*>(rec));
            pq1 = __ldcg(reinterpret_cast<const uint4*>(rec) + 1);
          }
        }
      }
      for (int it = 0; it < 4; ++it) {
        const int r = ww * 2 + 32 * it + half;  // this half-warp's row
        const int src = (lane & 16) + 4 * it;   // lane of this half that holds rank 0 of iteration it
        const uint32_t vm = __shfl_sync(FULL, pvm, src);
        const int nv = __popc(vm);
        const int nv_any = max(nv, __shfl_xor_sync(FULL, nv, 16));
        uint32_t fvp[FL_MAXV][4];  // interpolated features of the k-th visible view, already in the feature dtype (bf16x2)
        float score[FL_MAXV];
#pragma unroll
        for (int k = 0; k < FL_MAXV; ++k) {
          score[k] = 0.f;
#pragma unroll
          for (int j = 0; j < 4; ++j) fvp[k][j] = 0u;
        }
#pragma unroll
        for (int k = 0; k < FL_MAXV; ++k) {
          if (k >= nv_any) break;  // warp-uniform
          const uint32_t x0 = __shfl_sync(FULL, pq0.x, src + k), x1 = __shfl_sync(FULL, pq0.y, src + k);
          const uint32_t x2 = __shfl_sync(FULL, pq0.z, src + k), x3 = __shfl_sync(FULL, pq0.w, src + k);
          const uint32_t y0 = __shfl_sync(FULL, pq1.x, src + k), y1 = __shfl_sync(FULL, pq1.y, src + k);
          const uint32_t y2 = __shfl_sync(FULL, pq1.z, src + k), y3 = __shfl_sync(FULL, pq1.w, src + k);
          const bool mine = k < nv;
          float sp = 0.f, wb1 = 0.f;
          if (mine) {
            const float wr1 = __uint_as_float(y0), wc1 = __uint_as_float(y1);
            wb1 = __uint_as_float(y2);
            const int b0 = y3 & 0xffff, b1i = y3 >> 16;
            const float wr0 = __fadd_rn(1.0f, -wr1), wc0 = __fadd_rn(1.0f, -wc1);
            const __nv_bfloat16* p00 = A.fimg + x0;
            const __nv_bfloat16* p01 = A.fimg + x1;
            const __nv_bfloat16* p10 = A.fimg + x2;
            const __nv_bfloat16* p11 = A.fimg + x3;
            const uint4 u00 = __ldg(reinterpret_cast<const uint4*>(p00 + l16 * 8));
            const uint4 u01 = __ldg(reinterpret_cast<const uint4*>(p01 + l16 * 8));
            const uint4 u10 = __ldg(reinterpret_cast<const uint4*>(p10 + l16 * 8));
            const uint4 u11 = __ldg(reinterpret_cast<const uint4*>(p11 + l16 * 8));
            // tap weights: row weight x column weight (one rounding each), shared with the unfused kernel
            const float w00 = __fmul_rn(wr0, wc0), w01 = __fmul_rn(wr0, wc1);
            const float w10 = __fmul_rn(wr1, wc0), w11 = __fmul_rn(wr1, wc1);
            if (l16 < 8) {  // depth-score taps: 4 taps x 2 bins on the first 8 lanes of the half
              const int tap = l16 >> 1, bsel = l16 & 1;
              const __nv_bfloat16* pt = (tap & 2) ? ((tap & 1) ? p11 : p10) : ((tap & 1) ? p01 : p00);
              const float wt = ((tap & 2) ? wr1 : 1.0f - wr1) * ((tap & 1) ? wc1 : 1.0f - wc1);
              sp = wt * __bfloat162float(pt[P.D + (bsel ? b1i : b0)]);
            }
            const uint32_t a00[4] = {u00.x, u00.y, u00.z, u00.w}, a01[4] = {u01.x, u01.y, u01.z, u01.w};
            const uint32_t a10[4] = {u10.x, u10.y, u10.z, u10.w}, a11[4] = {u11.x, u11.y, u11.z, u11.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              // (w00 f00 + w01 f01) + (w10 f10 + w11 f11): lower tap row + upper tap row, as in the unfused kernel
              const float lo_a = __fmaf_rn(w01, bf16_lo(a01[j]), __fmul_rn(w00, bf16_lo(a00[j])));
              const float hi_a = __fmaf_rn(w11, bf16_lo(a11[j]), __fmul_rn(w10, bf16_lo(a10[j])));
              const float lo_b = __fmaf_rn(w01, bf16_hi(a01[j]), __fmul_rn(w00, bf16_hi(a00[j])));
              const float hi_b = __fmaf_rn(w11, bf16_hi(a11[j]), __fmul_rn(w10, bf16_hi(a10[j])));
              fvp[k][j] = pack_bf16(__fadd_rn(lo_a, hi_a), __fadd_rn(lo_b, hi_b));  // -> feature dtype
            }
          }
          // bin-wise spatial interpolation (-> bf16), then interpolation across the two bins (-> bf16);
          // xor 2/4/1 stay inside the 8 score lanes of each half
          sp += __shfl_xor_sync(FULL, sp, 2);
          sp += __shfl_xor_sync(FULL, sp, 4);
          sp = bf16_round(sp) * ((lane & 1) ? wb1 : 1.0f - wb1);
          sp += __shfl_xor_sync(FULL, sp, 1);
          score[k] = __shfl_sync(FULL, bf16_round(sp), lane & 16);  // broadcast from the half's lane 0
        }
        uint32_t mean_p[4], var_p[4];
        float smaxv = 0.f;
        const float ssum = (score[0] + score[1]) + (score[2] + score[3]);
        if (nv <= 1 && ssum > -80.f) {
          // one visible view: its softmax weight is exactly 1 (x / x), so mean = its features, variance = +0 and
          // score_max = its score (rank 0; a fully invisible row gives all zeros)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            mean_p[j] = fvp[0][j];
            var_p[j] = 0u;
          }
          smaxv = ssum;
        } else {
          float mean[8], var[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) mean[j] = var[j] = 0.f;
          float mx = 0.f;
          smaxv = -INFINITY;
#pragma unroll
          for (int k = 0; k < FL_MAXV; ++k)
            if (k < nv) {
              mx = fmaxf(mx, score[k]);
              smaxv = fmaxf(smaxv, score[k]);
            }
          float wv[FL_MAXV], den = 0.f;
#pragma unroll
          for (int k = 0; k < FL_MAXV; ++k) {
            wv[k] = (k < nv) ? expf(score[k] - mx) : 0.f;
            den += wv[k];
          }
#pragma unroll
          for (int k = 0; k < FL_MAXV; ++k) {
            if (k >= nv) continue;
            wv[k] = __fdiv_rn(wv[k], den);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              mean[2 * j] += wv[k] * bf16_lo(fvp[k][j]);
              mean[2 * j + 1] += wv[k] * bf16_hi(fvp[k][j]);
            }
          }
#pragma unroll
          for (int k = 0; k < FL_MAXV; ++k) {
            if (k >= nv) continue;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float da = bf16_lo(fvp[k][j]) - mean[2 * j], db = bf16_hi(fvp[k][j]) - mean[2 * j + 1];
              var[2 * j] += wv[k] * da * da;
              var[2 * j + 1] += wv[k] * db * db;
            }
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            mean_p[j] = pack_bf16(mean[2 * j], mean[2 * j + 1]);
            var_p[j] = pack_bf16(var[2 * j], var[2 * j + 1]);
          }
        }
        // A[r][k]: mean at k = 8*l16.., var at k = 128 + 8*l16..; K-chunk of 64, 16-byte slot (k%64)/8 XOR (r%8)
        // inside the 128-byte row (SWIZZLE_128B, as TMA would write it)
        {
          const int kc_m = l16 >> 3, kc_v = 2 + (l16 >> 3);
          const int slot16 = (l16 & 7) ^ (r & 7);
          *reinterpret_cast<uint4*>(smem + SM_AH + kc_m * 16384 + r * 128 + slot16 * 16) =
              make_uint4(mean_p[0], mean_p[1], mean_p[2], mean_p[3]);
          *reinterpret_cast<uint4*>(smem + SM_AH + kc_v * 16384 + r * 128 + slot16 * 16) =
              make_uint4(var_p[0], var_p[1], var_p[2], var_p[3]);
          if (l16 == 0) smax_s[r] = __float2bfloat16(smaxv);
        }
      }
      fence_proxy_async_smem();  // generic-proxy writes of A -> visible to the tensor core (async proxy)
      mbar_arrive(&ctl->a_full);
      LIFT_MARK(1);

    }
    if (wtid == 0) {
      atomicAdd(A.col_counter + 1, n_tiles);
      atomicAdd(A.col_counter + 2, n_rows);
#pragma unroll
      for (int i = 0; i < 7; ++i) atomicAdd(A.col_counter + 4 + i, (int)(tprof[i] >> 4));  // units of 16 cycles
    }
    if (wtid < 64 && zm_col >= 0) {
      reinterpret_cast<uint32_t*>(A.plane)[(size_t)zm_col * 64 + wtid] = zm_val;
      if (wtid == 0) A.pvalid[zm_col] = 1;
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

extern "C" size_t snapb200_lift_fused_scratch_bytes(void) {
  return (size_t)num_sms() * FL_LIST_CAP * FL_MAXV * sizeof(TapRec);
}

extern "C" int snapb200_lift_fused(const SnapLiftParams* q, const SnapLiftView* views, const void* fimg,
                                   const float* xs, const float* ys, const float* zs, const void* w1t,
                                   long long ldw1, const float* w256, const float* b1, const void* w2t,
                                   const float* b2, void* plane, uint8_t* pvalid, int* col_counter,
                                   void* scratch, size_t scratch_bytes, void* stream) {
  SNAP_REQUIRE(q && views && fimg && xs && ys && zs && w1t && w256 && b1 && w2t && b2 && plane && pvalid &&
                   col_counter,
               "null pointer");
  SNAP_REQUIRE(q->V >= 1 && q->V <= FL_MAXV, "fused lift handles 1..%d views (got %d)", FL_MAXV, q->V);
  SNAP_REQUIRE(q->D == 128 && q->S >= 2 && q->CF == q->D + q->S && q->CF % 8 == 0, "bad channel split");
  SNAP_REQUIRE(!q->no_variance && !q->add_minmax, "the fused lift implements the default statistics only");
  SNAP_REQUIRE(q->Z >= 1 && q->Z <= 64, "Z must be <= 64 (got %d)", q->Z);
  SNAP_REQUIRE((long long)q->V * q->Hf * q->Wf * q->CF < (1LL << 31), "feature maps too large for 32-bit tap offsets");
  SNAP_REQUIRE((long long)q->X * q->Y <= 65536, "too many BEV columns (X*Y <= 65536: 16-bit column ids in the z-max scan)");
  static_assert(sizeof(SnapLiftView) == sizeof(LiftView), "SnapLiftView layout");
  static_assert(sizeof(SnapLiftParams) == sizeof(LiftParams), "SnapLiftParams layout");
  cudaStream_t s = (cudaStream_t)stream;
  static DynSmemState smem_state;
  if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(&lift_fused_kernel), FL_SMEM_BYTES, &smem_state,
                               "cudaFuncSetAttribute(lift_fused)"))
    return rc;
  CUtensorMap tmW1, tmW2;
  int rc = make_tmap_2d_bf16(&tmW1, w1t, 256, 256, ldw1, 256, 64);
  if (rc) return rc;
  rc = make_tmap_2d_bf16(&tmW2, w2t, 128, 256, 256, 128, 64);
  if (rc) return rc;
  const long long cells = (long long)q->X * q->Y;
  rc = check_cuda(cudaMemsetAsync(plane, 0, (size_t)cells * 128 * 2, s), "memset plane");
  if (rc) return rc;
  rc = check_cuda(cudaMemsetAsync(pvalid, 0, (size_t)cells, s), "memset valid");
  if (rc) return rc;
  rc = check_cuda(cudaMemsetAsync(col_counter, 0, 16 * sizeof(int), s), "memset counter");
  if (rc) return rc;
  FusedArgs a;
  memcpy(&a.P, q, sizeof(LiftParams));
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
  const int max_useful = (int)((cells + FL_BATCH_COLS - 1) / FL_BATCH_COLS);
  if (grid > max_useful) grid = max_useful;
  SNAP_REQUIRE(scratch != nullptr && scratch_bytes >= (size_t)grid * FL_LIST_CAP * FL_MAXV * sizeof(TapRec),
               "scratch too small: need snapb200_lift_fused_scratch_bytes()");
  a.scratch = reinterpret_cast<TapRec*>(scratch);
  lift_fused_kernel<<<grid, FL_THREADS, FL_SMEM_BYTES, s>>>(tmW1, tmW2, a);
  return check_launch("lift_fused_kernel");
}
