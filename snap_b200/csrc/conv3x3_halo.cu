// 3x3 / stride-1 StdConv on the zero-bordered NHWC layout as a HALO GEMM (resnet.py:117-127: conv2 of a bottleneck unit).
//
// The segmented engine (gemm_tc.cuh) runs a 3x3 conv as 9 K-segments = 9 row-shifted TMA loads of the same matrix: every
// 128-row output tile pulls 9 x 16 KB per 64 input channels from L2 although the nine windows overlap almost completely
// (stage 1-2 of the trunk are bound by exactly that L2 -> shared-memory traffic).  Here ONE block of 128 + 2 (W + 3) rows
// per 64-channel K chunk is loaded (TMA, SWIZZLE_128B) and all nine taps read it by moving the A descriptor's start address
// by (tap offset) x 128 B: the hardware swizzle is a function of the absolute shared-memory address (self-test
// `snapb200_selftest_shifted_desc`), so a row-shifted descriptor addresses the row-shifted window.  The nine weight tiles
// per K chunk stay resident in shared memory for the whole launch when they fit (stage 1: 72 KB); otherwise (stage 2:
// 288 KB) they stream through a ring and a tile covers MT = 2 blocks of 128 rows that share every weight tile (two
// tcgen05.mma per K step into two accumulators), which halves the weight traffic per output row.
//
// Warp roles (320 threads, one persistent CTA per SM): warp 0 = TMA producer (halo blocks, double-buffered), warp 1 = TMEM
// owner + MMA issuer (9 x 4 tcgen05.mma per K chunk into a double-buffered accumulator), warps 2..9 = epilogue: drop the
// border rows (row remap padded -> dense), packed-bf16 stores, fused GroupNorm statistics of the stored values.
#include "gemm_tc.cuh"
#include "host_common.h"

namespace snapb200 {

struct HaloParams {
  int m_tiles;        // tiles of MT x 128 rows over the padded rows n_img * (H + 2) * (W + 2)
  int nb;             // 128-row TMA boxes per halo block
  int off_min;        // first row of the halo block relative to the tile's first row (= -(W + 3))
  int tap_off[9];     // row offset of tap (kh, kw) relative to off_min
  int kps;            // 64-channel K chunks
  int C;
  int n_img, H, W;
  long long M;        // padded rows
  int N;
  __nv_bfloat16* out;
  long long ldo;
  double* gn_acc;
  int gn_replica_stride, gn_cpg_log;
};

constexpr int HALO_SLOTS = 2;
constexpr int HALO_BRING = 4;   // weight ring stages when the weights are streamed

template <int BN, int MT, bool BRES>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
conv3x3_halo_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ HaloParams p) {
  constexpr int B_TILE = BN * 128;            // one (tap, K chunk) weight tile: BN rows x 64 bf16
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  const int a_slot_bytes = p.nb * 16384;
  uint8_t* sA = smem;
  uint8_t* sB = smem + HALO_SLOTS * a_slot_bytes;
  const int b_tiles = BRES ? 9 * p.kps : HALO_BRING;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(sB + b_tiles * B_TILE);   // 8-byte aligned: every block above is a multiple of 1024 B
  uint64_t* a_empty = a_full + HALO_SLOTS;
  uint64_t* b_full = a_empty + HALO_SLOTS;        // [1] resident weights / [HALO_BRING] ring
  uint64_t* b_empty = b_full + HALO_BRING;
  uint64_t* tmem_full = b_empty + HALO_BRING;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  constexpr int TMEM_COLS = 2 * MT * BN <= 128 ? 128 : (2 * MT * BN <= 256 ? 256 : 512);
  static_assert(2 * MT * BN <= 512, "two accumulator sets of MT x BN columns");

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile_begin = (int)((long long)blockIdx.x * p.m_tiles / gridDim.x);
  const int tile_end = (int)((long long)(blockIdx.x + 1) * p.m_tiles / gridDim.x);
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < HALO_SLOTS; ++s) {
      mbar_init(&a_full[s], 1);
      mbar_init(&a_empty[s], 1);
    }
    for (int s = 0; s < HALO_BRING; ++s) {
      mbar_init(&b_full[s], 1);
      mbar_init(&b_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], GEMM_EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr_smem, TMEM_COLS);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      if (BRES) {   // all weight tiles once: tile (tap, kc) = B[0:BN, tap * C + kc * 64 ...]
        mbar_arrive_expect_tx(&b_full[0], (uint32_t)(9 * p.kps * B_TILE));
        for (int tap = 0; tap < 9; ++tap)
          for (int kc = 0; kc < p.kps; ++kc)
            tma_load_2d(&tmB, &b_full[0], sB + (tap * p.kps + kc) * B_TILE, tap * p.C + kc * 64, 0);
      }
      int slot = 0, bs = 0;
      uint32_t phase = 0, bphase = 0;
      for (int tile = tile_begin; tile < tile_end; ++tile) {
        const int row0 = tile * (128 * MT) + p.off_min;   // may be negative / run past the end: TMA zero-fills
        for (int kc = 0; kc < p.kps; ++kc) {
          mbar_wait(&a_empty[slot], phase ^ 1);
          mbar_arrive_expect_tx(&a_full[slot], (uint32_t)a_slot_bytes);
          for (int b = 0; b < p.nb; ++b)
            tma_load_2d(&tmA, &a_full[slot], sA + slot * a_slot_bytes + b * 16384, kc * 64, row0 + b * 128);
          if (++slot == HALO_SLOTS) {
            slot = 0;
            phase ^= 1;
          }
          if (!BRES) {
            for (int tap = 0; tap < 9; ++tap) {
              mbar_wait(&b_empty[bs], bphase ^ 1);
              mbar_arrive_expect_tx(&b_full[bs], (uint32_t)B_TILE);
              tma_load_2d(&tmB, &b_full[bs], sB + bs * B_TILE, tap * p.C + kc * 64, 0);
              if (++bs == HALO_BRING) {
                bs = 0;
                bphase ^= 1;
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16_m128(BN);
      if (BRES) mbar_wait(&b_full[0], 0);
      int slot = 0, acc = 0, bs = 0;
      uint32_t phase = 0, acc_phase = 0, bphase = 0;
      for (int tile = tile_begin; tile < tile_end; ++tile) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after_sync();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * MT * BN);
        for (int kc = 0; kc < p.kps; ++kc) {
          mbar_wait(&a_full[slot], phase);
          tc_fence_after_sync();
          const uint32_t a0 = smem_u32(sA + slot * a_slot_bytes);
#pragma unroll 1
          for (int tap = 0; tap < 9; ++tap) {
            if (!BRES) {
              mbar_wait(&b_full[bs], bphase);
              tc_fence_after_sync();
            }
            const uint64_t db = make_kmajor_desc<128>(smem_u32(sB + (BRES ? (tap * p.kps + kc) : bs) * B_TILE));
#pragma unroll
            for (int st = 0; st < MT; ++st) {
              // the tap's window of sub-tile st = the halo block shifted by tap_off + 128 st rows: move the descriptor
              // start by whole 128-byte rows
              const uint64_t da = make_kmajor_desc<128>(a0 + (uint32_t)(p.tap_off[tap] + 128 * st) * 128u);
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16(tmem_d + (uint32_t)(st * BN), da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc,
                          (kc | tap | k) != 0 ? 1u : 0u);
            }
            if (!BRES) {
              umma_commit(&b_empty[bs]);
              if (++bs == HALO_BRING) {
                bs = 0;
                bphase ^= 1;
              }
            }
          }
          umma_commit(&a_empty[slot]);   // the halo block is free once these MMAs retire
          if (++slot == HALO_SLOTS) {
            slot = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tmem_full[acc]);
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else {
    // ===================== epilogue warps =====================
    const int q = warp & 3;          // TMEM lane quadrant
    const int hw = (warp - 2) >> 2;  // which half of the 16-column chunks
    constexpr int NCW = BN / 32;     // chunks per warp
    const int hp = p.H + 2, wp = p.W + 2;
    const unsigned per_img = (unsigned)(hp * wp);
    double* gacc = p.gn_acc != nullptr ? p.gn_acc + (size_t)(blockIdx.x % GN_REPLICAS) * p.gn_replica_stride : nullptr;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = tile_begin; tile < tile_end; ++tile) {
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after_sync();
#pragma unroll 1
      for (int st = 0; st < MT; ++st) {
        const unsigned m = (unsigned)tile * (128u * MT) + (unsigned)(st * 128 + q * 32 + lane);
        const unsigned img = m / per_img;
        const int rem = (int)(m - img * per_img);
        const int r = rem / wp, c = rem - r * wp;
        const bool row_ok = (long long)m < p.M && r >= 1 && r <= p.H && c >= 1 && c <= p.W;
        const size_t orow = row_ok ? ((size_t)img * p.H + (r - 1)) * p.W + (c - 1) : 0;
        const int gn_img = row_ok ? (int)img : -1;
        const int gn_ref = __reduce_max_sync(0xffffffffu, gn_img);
        const bool gn_uniform = __all_sync(0xffffffffu, gn_img == gn_ref || gn_img == -1);
        if (gn_ref < 0) continue;   // warp-uniform: a quadrant that holds only border rows skips the drain
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((acc * MT + st) * BN + hw * NCW * 16);
        uint32_t v[2][16];
        tmem_ld16(taddr, v[0]);
#pragma unroll
        for (int i = 0; i < NCW; ++i) {
          tmem_ld_wait();
          if (i + 1 < NCW) tmem_ld16(taddr + (uint32_t)((i + 1) * 16), v[(i + 1) & 1]);
          const uint32_t(&vv)[16] = v[i & 1];
          const int col = (hw * NCW + i) * 16;
          uint32_t pk[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) pk[j] = pack_bf16(__uint_as_float(vv[2 * j]), __uint_as_float(vv[2 * j + 1]));
          if (row_ok) {
            uint4* op = reinterpret_cast<uint4*>(p.out + orow * p.ldo + col);
            op[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            op[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
          }
          if (gacc != nullptr) gn_accumulate16p(pk, row_ok, gn_img, col, p.gn_cpg_log, gacc, gn_uniform, gn_ref, lane);
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
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

// shape plan: MT sub-tiles per tile, halo boxes, resident or streamed weights; returns 0 when the shape does not fit
struct HaloPlan {
  int mt, nb, bres, smem;
};
static HaloPlan halo_plan(int C, int n, int W) {
  HaloPlan h = {0, 0, 0, 0};
  if (C < 64 || C % 64 != 0 || (n != 64 && n != 128) || W < 1) return h;
  const int kps = C / 64, b_tile = n * 128, span = 2 * (W + 3);
  const int budget = 225 * 1024 - 512;
  // resident weights with one 128-row block per tile (stage 1) ...
  int nb = (span + 128 + 127) / 128;
  if (HALO_SLOTS * nb * 16384 + 9 * kps * b_tile <= budget) return HaloPlan{1, nb, 1, HALO_SLOTS * nb * 16384 + 9 * kps * b_tile + 512};
  // ... else streamed weights shared by two 128-row blocks (needs 2 x 2 x n TMEM columns)
  nb = (span + 256 + 127) / 128;
  if (4 * n <= 512 && HALO_SLOTS * nb * 16384 + HALO_BRING * b_tile <= budget)
    return HaloPlan{2, nb, 0, HALO_SLOTS * nb * 16384 + HALO_BRING * b_tile + 512};
  return h;
}

template <int BN, int MT, bool BRES>
static int launch_halo(const CUtensorMap& tmA, const CUtensorMap& tmB, const HaloParams& p, int smem, cudaStream_t s) {
  static DynSmemState st;
  if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(&conv3x3_halo_kernel<BN, MT, BRES>), 227 * 1024, &st,
                               "cudaFuncSetAttribute(conv3x3_halo)"))
    return rc;
  const int grid = p.m_tiles < num_sms() ? p.m_tiles : num_sms();
  conv3x3_halo_kernel<BN, MT, BRES><<<grid, GEMM_THREADS, smem, s>>>(tmA, tmB, p);
  return check_launch("conv3x3_halo_kernel");
}

}  // namespace snapb200

using namespace snapb200;

extern "C" int snapb200_conv3x3_halo_supported(int C, int n, int W) { return halo_plan(C, n, W).mt != 0 ? 1 : 0; }

extern "C" int snapb200_conv3x3_halo_bf16(const SnapConv3x3Params* q, void* stream) {
  SNAP_REQUIRE(q != nullptr && q->a && q->b && q->out, "null operand");
  SNAP_REQUIRE(q->n_img >= 1 && q->H >= 1 && q->W >= 1, "empty input");
  const HaloPlan h = halo_plan(q->C, q->n, q->W);
  SNAP_REQUIRE(h.mt != 0, "conv3x3_halo: unsupported shape (C %% 64 == 0, n in {64, 128}, halo blocks + weights within 225 KB): "
               "C=%d n=%d W=%d", q->C, q->n, q->W);
  SNAP_REQUIRE(q->ldo % 8 == 0 && (reinterpret_cast<uintptr_t>(q->out) & 15) == 0, "out must be 16-byte aligned rows");
  const int hp = q->H + 2, wp = q->W + 2;
  HaloParams p = {};
  p.M = (long long)q->n_img * hp * wp;
  SNAP_REQUIRE(p.M + 256 < (1LL << 31), "too many rows");
  p.m_tiles = (int)((p.M + 128 * h.mt - 1) / (128 * h.mt));
  p.off_min = -(wp + 1);
  p.nb = h.nb;
  for (int t = 0; t < 9; ++t) p.tap_off[t] = (t / 3 - 1) * wp + (t % 3 - 1) - p.off_min;
  p.kps = q->C / 64;
  p.C = q->C;
  p.n_img = q->n_img;
  p.H = q->H;
  p.W = q->W;
  p.N = q->n;
  p.out = static_cast<__nv_bfloat16*>(q->out);
  p.ldo = q->ldo;
  p.gn_acc = q->gn_acc;
  p.gn_replica_stride = q->gn_replica_stride;
  const int cpg = q->n / 32;
  while ((1 << p.gn_cpg_log) < cpg) ++p.gn_cpg_log;
  if (q->gn_acc != nullptr) SNAP_REQUIRE(q->gn_replica_stride > 0, "gn_replica_stride required with gn_acc");
  CUtensorMap tmA, tmB;
  // rows beyond the tensor (and before it: negative coordinates) are zero-filled, like the zero border itself
  int rc = make_tmap_2d_bf16(&tmA, q->a, p.M, q->C, q->C, 128, 64);
  if (rc) return rc;
  rc = make_tmap_2d_bf16(&tmB, q->b, q->n, 9 * q->C, q->b_ld, q->n, 64);
  if (rc) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (q->n == 64) return h.bres ? launch_halo<64, 1, true>(tmA, tmB, p, h.smem, s) : launch_halo<64, 2, false>(tmA, tmB, p, h.smem, s);
  return h.bres ? launch_halo<128, 1, true>(tmA, tmB, p, h.smem, s) : launch_halo<128, 2, false>(tmA, tmB, p, h.smem, s);
}
