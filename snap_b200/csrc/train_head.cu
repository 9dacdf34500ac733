// Head-only training step of the semantic fine-tuning configuration (BASELINE configs[4]; snap/configs/
// train_semantics.py freezes `bev_mapper/`, snap/trainer.py:209-247) for the default 'mlp' decoder of
// snap/models/semantic_net.py:147-152 (layers.MLP: Dense -> relu -> Dense -> relu -> Dense):
//
//   d total / d logits (loss of semantic_net.py:300-343, mean over the batch :221)   -> sem_loss_grad_kernel
//   dX = dY W^T                                                                      -> snapb200_gemm_bf16 (B = the Flax
//                                                                                       kernel [in, out] as stored)
//   relu backward                                                                    -> relu_bwd_kernel
//   dW = X^T dY, db = sum_m dY (fp32 accumulation)                                   -> dense_wgrad_kernel (+ reduce)
//   optax.adam update on fp32 master parameters                                      -> adam_kernel
//
// The weight-gradient product contracts over the M = B*G*G cells; its output is at most 256 x 256, so it is a split-K
// reduction: every CTA multiplies a slab of rows with warp-level MMAs (nvcuda::wmma, bf16 in / fp32 out) and writes a
// partial, a second kernel adds the partials in slab order (deterministic).  It is bound by reading X and dY once.
#include <cuda_bf16.h>
#include <math.h>
#include <mma.h>

#include "common.cuh"
#include "host_common.h"

namespace snapb200 {

constexpr int TH_MAXC = 8;

__device__ __forceinline__ float sigmoid_f(float x) { return 1.f / (1.f + expf(-x)); }

// counts[b] = {number of cells with valid_area, number of cells with valid}
__global__ void __launch_bounds__(256)
sem_counts_kernel(const uint8_t* __restrict__ valid_area, const uint8_t* __restrict__ valid, int cells,
                  float* __restrict__ counts) {
  __shared__ int sa[8], sv[8];
  const int b = blockIdx.x;
  int ca = 0, cv = 0;
  for (int c = threadIdx.x; c < cells; c += 256) {
    ca += valid_area[(size_t)b * cells + c] != 0;
    cv += valid[(size_t)b * cells + c] != 0;
  }
  ca = __reduce_add_sync(0xffffffffu, ca);
  cv = __reduce_add_sync(0xffffffffu, cv);
  if ((threadIdx.x & 31) == 0) {
    sa[threadIdx.x >> 5] = ca;
    sv[threadIdx.x >> 5] = cv;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int a = 0, v = 0;
    for (int w = 0; w < 8; ++w) {
      a += sa[w];
      v += sv[w];
    }
    counts[2 * b] = (float)a;
    counts[2 * b + 1] = (float)v;
  }
}

// one thread per cell: gradient of mean_b(total_b) w.r.t. the (masked) logits, as bf16 rows of ld_out columns
__global__ void __launch_bounds__(256)
sem_loss_grad_kernel(const SnapSemLossParams P, const float* __restrict__ logits, const int* __restrict__ labels_area,
                     const uint8_t* __restrict__ valid_area, const int* __restrict__ labels_excl,
                     const uint8_t* __restrict__ masks_indep, const uint8_t* __restrict__ valid,
                     const float* __restrict__ w_area, const float* __restrict__ w_excl, const float* __restrict__ w_pos,
                     const float* __restrict__ w_neg, const float* __restrict__ counts, int ld_out,
                     __nv_bfloat16* __restrict__ dlogits) {
  const long long g = (long long)blockIdx.x * 256 + threadIdx.x;
  if (g >= (long long)P.B * P.cells) return;
  const int b = (int)(g / P.cells);
  const int Ka = P.num_area, Ke = P.num_excl, Ki = P.num_indep;
  __nv_bfloat16* o = dlogits + g * ld_out;
  const bool v = valid[g] != 0;
  float out[3 * TH_MAXC];
#pragma unroll
  for (int c = 0; c < 3 * TH_MAXC; ++c) out[c] = 0.f;
  if (v) {  // logits = where(valid, logits, 0) (:186): no gradient reaches the decoder through invalid cells
    const float* l = logits + g * P.ld;
    const float cells = (float)P.cells;
    const float den_a = counts[2 * b] > 0.f ? counts[2 * b] : cells;      // layers.masked_mean
    const float den_v = counts[2 * b + 1] > 0.f ? counts[2 * b + 1] : cells;
    const bool objects = Ke > 0 || Ki > 0;
    // total = nll_a  or  (nll_a + (nll_e + nll_i) / 2) / 2  (:339); loss = mean over the batch (trainer.py:221)
    const float ca = (objects ? 0.5f : 1.f) / (float)P.B / den_a;
    const float ce = 0.25f / (float)P.B / den_v;
    if (valid_area[g]) {
      const int lab = labels_area[g];
      float mx = l[0];
      for (int c = 1; c < Ka; ++c) mx = fmaxf(mx, l[c]);
      float se = 0.f;
      for (int c = 0; c < Ka; ++c) se += expf(l[c] - mx);
      const float w = (w_area ? w_area[lab] : 1.f) * ca;
#pragma unroll
      for (int c = 0; c < TH_MAXC; ++c)
        if (c < Ka) out[c] = w * (expf(l[c] - mx) / se - (c == lab ? 1.f : 0.f));
    }
    if (Ke > 0) {
      const int lab = labels_excl[g];
      const float* le = l + Ka;
      float mx = le[0];
      for (int c = 1; c < Ke; ++c) mx = fmaxf(mx, le[c]);
      float se = 0.f;
      for (int c = 0; c < Ke; ++c) se += expf(le[c] - mx);
      const float w = (w_excl ? w_excl[lab] : 1.f) * ce;
#pragma unroll
      for (int c = 0; c < TH_MAXC; ++c)
        if (c < Ke) out[TH_MAXC + c] = w * (expf(le[c] - mx) / se - (c == lab ? 1.f : 0.f));
    }
    if (Ki > 0) {
      const float* li = l + Ka + Ke;
#pragma unroll
      for (int c = 0; c < TH_MAXC; ++c)
        if (c < Ki) {
          const bool gt = masks_indep[g * Ki + c] != 0;
          const float s = sigmoid_f(li[c]);
          // d/dx [-y w+ log_sigmoid(x) - (1 - y) w- log_sigmoid(-x)]
          const float d = gt ? -(w_pos ? w_pos[c] : 1.f) * (1.f - s) : (w_neg ? w_neg[c] : 1.f) * s;
          out[2 * TH_MAXC + c] = d * ce / (float)Ki;
        }
    }
  }
  for (int c = 0; c < ld_out; ++c) {
    float x = 0.f;
    if (c < Ka) x = out[c];
    else if (c < Ka + Ke) x = out[TH_MAXC + c - Ka];
    else if (c < Ka + Ke + Ki) x = out[2 * TH_MAXC + c - Ka - Ke];
    o[c] = __float2bfloat16(x);
  }
}

// dx = where(h > 0, dx, 0) on bf16 vectors of 8 (h = the post-ReLU activation: h > 0 <=> pre-activation > 0)
__global__ void relu_bwd_kernel(const __nv_bfloat16* __restrict__ h, __nv_bfloat16* __restrict__ dx, long long vecs) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= vecs) return;
  const uint4 hv = __ldg(reinterpret_cast<const uint4*>(h) + i);
  uint4 d = reinterpret_cast<uint4*>(dx)[i];
  const uint32_t hh[4] = {hv.x, hv.y, hv.z, hv.w};
  uint32_t dd[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const bool lo = bf16_lo(hh[j]) > 0.f, hi = bf16_hi(hh[j]) > 0.f;
    dd[j] = (lo ? (dd[j] & 0x0000ffffu) : 0u) | (hi ? (dd[j] & 0xffff0000u) : 0u);
  }
  reinterpret_cast<uint4*>(dx)[i] = make_uint4(dd[0], dd[1], dd[2], dd[3]);
}

// partial[s] = X[rows of slab s]^T dY[rows of slab s]  (K x N fp32), pbias[s] = column sums of dY over the slab
__global__ void __launch_bounds__(256)
dense_wgrad_kernel(const __nv_bfloat16* __restrict__ X, long long ldx, const __nv_bfloat16* __restrict__ dY,
                   long long ldy, long long M, int K, int N, int rows_per_slab, float* __restrict__ partial,
                   float* __restrict__ pbias) {
  using namespace nvcuda;
  const int slab = blockIdx.x, warp = threadIdx.x >> 5;
  const long long m0 = (long long)slab * rows_per_slab;
  const long long m1 = min(M, m0 + rows_per_slab);   // M and rows_per_slab are multiples of 16 (host)
  const int tk = K / 16, tn = N / 16, tiles = tk * tn;
  float* out = partial + (size_t)slab * K * N;
  for (int t0 = warp; t0 < tiles; t0 += 8 * 4) {      // up to 4 output tiles per warp and pass
    wmma::fragment<wmma::accumulator, 16, 16, 16, float> acc[4];
    int ti[4], tj[4], nt = 0;
    for (int u = 0; u < 4; ++u) {
      const int t = t0 + u * 8;
      if (t < tiles) {
        ti[nt] = t / tn;
        tj[nt] = t % tn;
        wmma::fill_fragment(acc[nt], 0.f);
        ++nt;
      }
    }
    for (long long m = m0; m < m1; m += 16) {
      for (int u = 0; u < nt; ++u) {
        wmma::fragment<wmma::matrix_a, 16, 16, 16, __nv_bfloat16, wmma::col_major> a;  // A(i, m) = X[m][i]
        wmma::fragment<wmma::matrix_b, 16, 16, 16, __nv_bfloat16, wmma::row_major> bfrag;
        wmma::load_matrix_sync(a, X + m * ldx + ti[u] * 16, (unsigned)ldx);
        wmma::load_matrix_sync(bfrag, dY + m * ldy + tj[u] * 16, (unsigned)ldy);
        wmma::mma_sync(acc[u], a, bfrag, acc[u]);
      }
    }
    for (int u = 0; u < nt; ++u)
      wmma::store_matrix_sync(out + (size_t)ti[u] * 16 * N + tj[u] * 16, acc[u], (unsigned)N, wmma::mem_row_major);
  }
  for (int n = threadIdx.x; n < N; n += 256) {
    float s = 0.f;
    for (long long m = m0; m < m1; ++m) s += __bfloat162float(dY[m * ldy + n]);
    pbias[(size_t)slab * N + n] = s;
  }
}

__global__ void wgrad_reduce_kernel(const float* __restrict__ partial, const float* __restrict__ pbias, int slabs,
                                    int KN, int N, float* __restrict__ dW, float* __restrict__ db) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < KN) {
    float s = 0.f;
    for (int k = 0; k < slabs; ++k) s += partial[(size_t)k * KN + i];
    dW[i] = s;
  } else if (i < KN + N && db != nullptr) {
    const int n = i - KN;
    float s = 0.f;
    for (int k = 0; k < slabs; ++k) s += pbias[(size_t)k * N + n];
    db[n] = s;
  }
}

// dst bf16 [rows, ld] = src f32 [rows, cols], zero in the padding columns (the Flax kernel [in, out] as the K-major
// B operand of the dX = dY W^T GEMM)
__global__ void cast_pad_bf16_kernel(const float* __restrict__ src, int rows, int cols, int ld,
                                     __nv_bfloat16* __restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * ld) return;
  const int r = i / ld, c = i - r * ld;
  dst[i] = __float2bfloat16(c < cols ? src[(size_t)r * cols + c] : 0.f);
}

// optax.adam: m = b1 m + (1 - b1) g; v = b2 v + (1 - b2) g^2; p -= lr * (m / (1 - b1^t)) / (sqrt(v / (1 - b2^t)) + eps)
__global__ void adam_kernel(float* __restrict__ p, float* __restrict__ m, float* __restrict__ v,
                            const float* __restrict__ g, long long n, float lr, float b1, float b2, float eps,
                            float bc1, float bc2) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float gi = g[i];
  const float mi = b1 * m[i] + (1.f - b1) * gi;
  const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
  m[i] = mi;
  v[i] = vi;
  p[i] -= lr * (mi / bc1) / (sqrtf(vi / bc2) + eps);
}

}  // namespace snapb200

using namespace snapb200;

extern "C" {

int snapb200_sem_loss_grad(const SnapSemLossParams* p, const float* logits, const int* labels_area,
                           const uint8_t* valid_area, const int* labels_excl, const uint8_t* masks_indep,
                           const uint8_t* valid, const float* w_area, const float* w_excl, const float* w_pos,
                           const float* w_neg, float* counts, int ld_out, void* dlogits, void* stream) {
  SNAP_REQUIRE(p && logits && labels_area && valid_area && valid && counts && dlogits, "null pointer");
  SNAP_REQUIRE(p->B >= 1 && p->cells >= 1, "empty problem");
  SNAP_REQUIRE(p->num_area >= 1 && p->num_area <= TH_MAXC && p->num_excl >= 0 && p->num_excl <= TH_MAXC &&
                   p->num_indep >= 0 && p->num_indep <= TH_MAXC, "at most %d classes per group", TH_MAXC);
  SNAP_REQUIRE(p->ld >= p->num_area + p->num_excl + p->num_indep && ld_out >= p->num_area + p->num_excl + p->num_indep,
               "row pitch too small");
  SNAP_REQUIRE(p->num_excl == 0 || labels_excl, "exclusive-object labels missing");
  SNAP_REQUIRE(p->num_indep == 0 || masks_indep, "independent-object masks missing");
  cudaStream_t s = (cudaStream_t)stream;
  sem_counts_kernel<<<p->B, 256, 0, s>>>(valid_area, valid, p->cells, counts);
  if (int rc = check_launch("sem_counts_kernel")) return rc;
  const long long total = (long long)p->B * p->cells;
  sem_loss_grad_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(*p, logits, labels_area, valid_area, labels_excl,
                                                                        masks_indep, valid, w_area, w_excl, w_pos, w_neg,
                                                                        counts, ld_out, (__nv_bfloat16*)dlogits);
  return check_launch("sem_loss_grad_kernel");
}

int snapb200_relu_bwd(const void* h, void* dx, long long elems, void* stream) {
  SNAP_REQUIRE(h && dx && elems % 8 == 0, "bad arguments");
  const long long vecs = elems / 8;
  relu_bwd_kernel<<<(unsigned)((vecs + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)h, (__nv_bfloat16*)dx, vecs);
  return check_launch("relu_bwd_kernel");
}

static int wgrad_slabs(long long M, int* rows_per_slab) {
  long long rps = (M + 2 * num_sms() - 1) / (2 * num_sms());
  rps = (rps + 15) / 16 * 16;
  if (rps < 64) rps = 64;
  *rows_per_slab = (int)rps;
  return (int)((M + rps - 1) / rps);
}

size_t snapb200_dense_wgrad_workspace(long long M, int K, int N) {
  int rps;
  const int slabs = wgrad_slabs(M, &rps);
  return (size_t)slabs * ((size_t)K * N + N) * sizeof(float);
}

int snapb200_dense_wgrad(const void* x, long long ldx, const void* dy, long long ldy, long long M, int K, int N,
                         float* dW, float* db, void* workspace, size_t workspace_bytes, void* stream) {
  SNAP_REQUIRE(x && dy && dW && workspace, "null pointer");
  SNAP_REQUIRE(M >= 16 && M % 16 == 0, "M must be a positive multiple of 16 (got %lld)", M);
  SNAP_REQUIRE(K % 16 == 0 && N % 16 == 0 && K >= 16 && N >= 16 && K <= 1024 && N <= 1024, "K, N: multiples of 16 in [16, 1024]");
  SNAP_REQUIRE(ldx % 8 == 0 && ldy % 8 == 0 && ldx >= K && ldy >= N, "row pitches must be multiples of 8");
  int rps;
  const int slabs = wgrad_slabs(M, &rps);
  SNAP_REQUIRE(workspace_bytes >= (size_t)slabs * ((size_t)K * N + N) * sizeof(float), "workspace too small");
  float* partial = (float*)workspace;
  float* pbias = partial + (size_t)slabs * K * N;
  cudaStream_t s = (cudaStream_t)stream;
  dense_wgrad_kernel<<<slabs, 256, 0, s>>>((const __nv_bfloat16*)x, ldx, (const __nv_bfloat16*)dy, ldy, M, K, N, rps,
                                           partial, pbias);
  if (int rc = check_launch("dense_wgrad_kernel")) return rc;
  const int total = K * N + N;
  wgrad_reduce_kernel<<<(total + 255) / 256, 256, 0, s>>>(partial, pbias, slabs, K * N, N, dW, db);
  return check_launch("wgrad_reduce_kernel");
}

int snapb200_cast_pad_bf16(const float* src, int rows, int cols, int ld, void* dst, void* stream) {
  SNAP_REQUIRE(src && dst && rows >= 1 && cols >= 1 && ld >= cols, "bad arguments");
  cast_pad_bf16_kernel<<<(rows * ld + 255) / 256, 256, 0, (cudaStream_t)stream>>>(src, rows, cols, ld, (__nv_bfloat16*)dst);
  return check_launch("cast_pad_bf16_kernel");
}

int snapb200_adam_step(float* p, float* m, float* v, const float* g, long long n, float lr, float b1, float b2,
                       float eps, int step, void* stream) {
  SNAP_REQUIRE(p && m && v && g && n >= 1 && step >= 1, "bad arguments");
  const float bc1 = 1.f - powf(b1, (float)step), bc2 = 1.f - powf(b2, (float)step);
  adam_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(p, m, v, g, n, lr, b1, b2, eps, bc1, bc2);
  return check_launch("adam_kernel");
}

}  // extern "C"
