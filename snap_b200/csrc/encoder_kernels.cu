// HBM-bound kernels of the image encoder (everything that is not a GEMM):
//   weight standardisation (StdConv, snap/models/resnet.py:34-41,73-79) -> GEMM B operand
//   root im2col (2x-1 normalisation + zero/"-1" padding, resnet.py:199, image_encoder.py:32-39)
//   3x3/2 max-pool (resnet.py:99), GroupNorm statistics + apply (resnet.py:46-70),
//   x2 bilinear up-sampling (image_encoder.py:86-91), crop(+ReLU) (image_encoder.py:137-141).
// All are coalesced 16-byte-vector streaming kernels over NHWC bf16.
#include <cuda_bf16.h>
#include <math.h>

#include "common.cuh"
#include "host_common.h"

namespace snapb200 {

// ------------------------------------------------------------------------------------------
// weight standardisation (batched over all convs of an encoder in two launches):
//   HWIO fp32 [K, Cout] (K = KH*KW*Cin) -> bf16 [Cout, ldb], column = k, zero padded to ldb.
// Kernel A: partial (sum, sumsq) per (conv, cout tile, k split); kernel B: finalise + transpose.
// ------------------------------------------------------------------------------------------
struct WeightDesc {
  const float* w;
  __nv_bfloat16* out;
  float2* partial;  // [ksplit][Cout]
  int K, Cout, ldb, standardize, ksplit, pad_;
};

__global__ void std_weights_stats_kernel(const WeightDesc* __restrict__ descs,
                                         const int4* __restrict__ map) {
  const int4 m = map[blockIdx.x];  // (desc, cout tile, k split, -)
  const WeightDesc d = descs[m.x];
  __shared__ float2 red[8][33];
  const int co = m.y * 32 + threadIdx.x;
  const bool ok = co < d.Cout;
  const int kc = (d.K + d.ksplit - 1) / d.ksplit;
  const int k0 = m.z * kc, k1 = min(d.K, k0 + kc);
  float s = 0.f, q = 0.f;
  for (int k = k0 + threadIdx.y; k < k1; k += 8) {
    const float v = ok ? d.w[(size_t)k * d.Cout + co] : 0.f;
    s += v;
    q += v * v;
  }
  red[threadIdx.y][threadIdx.x] = make_float2(s, q);
  __syncthreads();
  if (threadIdx.y == 0 && ok) {
    float2 t = make_float2(0.f, 0.f);
    for (int i = 0; i < 8; ++i) {
      t.x += red[i][threadIdx.x].x;
      t.y += red[i][threadIdx.x].y;
    }
    d.partial[(size_t)m.z * d.Cout + co] = t;
  }
}

__global__ void std_weights_write_kernel(const WeightDesc* __restrict__ descs,
                                         const int4* __restrict__ map, float eps) {
  const int4 m = map[blockIdx.x];  // (desc, cout tile, k chunk of 8 x 32, -)
  const WeightDesc d = descs[m.x];
  __shared__ float s_mean[32], s_rstd[32];
  __shared__ float tile[32][33];
  if (threadIdx.y == 0) {
    const int co = m.y * 32 + threadIdx.x;
    float mean = 0.f, rstd = 1.f;
    if (d.standardize && co < d.Cout) {
      double s = 0.0, q = 0.0;
      for (int i = 0; i < d.ksplit; ++i) {
        const float2 t = d.partial[(size_t)i * d.Cout + co];
        s += (double)t.x;
        q += (double)t.y;
      }
      const double mu = s / (double)d.K;
      double var = q / (double)d.K - mu * mu;  // == mean((x - mu)^2)
      if (var < 0.0) var = 0.0;
      mean = (float)mu;
      rstd = (float)(1.0 / sqrt(var + (double)eps));
    }
    s_mean[threadIdx.x] = mean;
    s_rstd[threadIdx.x] = rstd;
  }
  __syncthreads();
  const int co = m.y * 32 + threadIdx.x;
  // a block converts up to 8 consecutive 32 x 32 tiles (256 k) of its 32 output channels: the statistics prologue
  // above (a chain of dependent global loads) is paid once per 8 tiles
  for (int sub = 0; sub < 8; ++sub) {
    const int kb = (m.z * 8 + sub) * 32;
    if (kb >= d.ldb) break;
    for (int kk = threadIdx.y; kk < 32; kk += 8) {
      const int k = kb + kk;
      float v = 0.f;
      if (co < d.Cout && k < d.K) v = (d.w[(size_t)k * d.Cout + co] - s_mean[threadIdx.x]) * s_rstd[threadIdx.x];
      tile[kk][threadIdx.x] = v;
    }
    __syncthreads();
    for (int cc = threadIdx.y; cc < 32; cc += 8) {
      const int c2 = m.y * 32 + cc;
      const int k = kb + threadIdx.x;
      if (c2 < d.Cout && k < d.ldb) d.out[(size_t)c2 * d.ldb + k] = __float2bfloat16(tile[threadIdx.x][cc]);
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------
// implicit root conv: packed image + per-kernel-row weight layout (see snapb200.h)
// ------------------------------------------------------------------------------------------
__global__ void root_pack_image_kernel(const float* __restrict__ img, int Nimg, int H, int W, int Hp, int Wp, int pad,
                                       int cp, int Hq, int Wq, __nv_bfloat16* __restrict__ out) {
  const long long total = (long long)Nimg * Hq * Wq;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int wq = (int)(idx % Wq);
  const long long t = idx / Wq;
  const int hq = (int)(t % Hq);
  const int n = (int)(t / Hq);
  const int h = hq - pad, w = wq - pad;
  float v[3] = {0.f, 0.f, 0.f};
  if (h >= 0 && h < Hp && w >= 0 && w < Wp) {
    if (h < H && w < W) {
      const float* px = img + (((size_t)n * H + h) * W + w) * 3;
#pragma unroll
      for (int c = 0; c < 3; ++c) v[c] = bf16_round(bf16_round(__ldg(px + c)) * 2.0f - 1.0f);  // resnet.py:199
    } else {
      v[0] = v[1] = v[2] = -1.0f;  // zero padding of pad_to_multiple, after 2x - 1
    }
  }
  uint32_t* o = reinterpret_cast<uint32_t*>(out + (size_t)idx * cp);
  o[0] = pack_bf16(v[0], v[1]);
  o[1] = pack_bf16(v[2], 0.f);
  if (cp == 8) {
    o[2] = 0u;
    o[3] = 0u;
  }
}

__global__ void root_pack_weights_kernel(const __nv_bfloat16* __restrict__ b, int Cout, int ldb, int KH, int KW, int cp,
                                         __nv_bfloat16* __restrict__ out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= Cout * KH * 32) return;
  const int e = idx % 32, kh = (idx / 32) % KH, co = idx / (32 * KH);
  const int kw = e / cp, c = e - kw * cp;
  __nv_bfloat16 v = __float2bfloat16(0.f);
  if (kw < KW && c < 3) v = b[(size_t)co * ldb + (kh * KW + kw) * 3 + c];
  out[idx] = v;
}

// ------------------------------------------------------------------------------------------
// root im2col: image f32 [Nimg,H,W,3] in [0,1] -> A [Nimg*Ho*Wo, Kp] bf16 for a KHxKW/stride conv
// over the (Hp,Wp) zero-padded-then-normalised image (pad value becomes -1, resnet.py:199).
// ------------------------------------------------------------------------------------------
// One warp per output pixel; lane l writes the bf16 pairs (k = 2l, 2l+1), (2l+64, ..): the K axis of a row is
// contiguous in the output (coalesced 128 B stores) and nearly contiguous in the NHWC image (kw, c fastest).
template <int KW>
__global__ void __launch_bounds__(256)
root_im2col_kernel(const float* __restrict__ img, int Nimg, int H, int W, int Hp, int Wp, int KH,
                   int stride, int pad, int Ho, int Wo, __nv_bfloat16* __restrict__ out, int Kp) {
  const int lane = threadIdx.x & 31;
  const int warp_id = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int nwarps = (int)((gridDim.x * blockDim.x) >> 5);
  const int total = Nimg * Ho * Wo;  // < 2^31 (checked by the host wrapper)
  constexpr int KW3 = KW * 3;
  const int K = KH * KW3;
  for (int row = warp_id; row < total; row += nwarps) {
    const int wo = row % Wo;
    const int t_ = row / Wo;
    const int ho = t_ % Ho;
    const int n = t_ / Ho;
    const int h0 = ho * stride - pad, w0 = wo * stride - pad;
    const float* base = img + (size_t)n * H * W * 3;
    for (int k2 = lane * 2; k2 < Kp; k2 += 64) {
      float v[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int k = k2 + e;
        float val = 0.f;
        if (k < K) {
          const int kh = k / KW3, r = k - kh * KW3;
          const int kw = r / 3, c = r - kw * 3;
          const int hi = h0 + kh, wi = w0 + kw;
          if (hi >= 0 && hi < Hp && wi >= 0 && wi < Wp) {
            if (hi < H && wi < W) {
              const float x = bf16_round(__ldg(base + ((size_t)hi * W + wi) * 3 + c));  // images.astype(dtype)
              val = bf16_round(x * 2.0f - 1.0f);                                        // resnet.py:199
            } else {
              val = -1.0f;  // zero padding of pad_to_multiple, after 2x - 1
            }
          }
        }
        v[e] = val;
      }
      *reinterpret_cast<uint32_t*>(out + (size_t)row * Kp + k2) = pack_bf16(v[0], v[1]);
    }
  }
}

// ------------------------------------------------------------------------------------------
// 3x3 stride-2 max pool, pad 1 (-inf)
// ------------------------------------------------------------------------------------------
__global__ void maxpool3x3s2_kernel(const __nv_bfloat16* __restrict__ x, int Nimg, int H, int W,
                                    int C, __nv_bfloat16* __restrict__ y, int Ho, int Wo) {
  const int cv = C / 8;
  const long long total = (long long)Nimg * Ho * Wo * cv;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c8 = (int)(idx % cv);
  const long long pix = idx / cv;
  const int wo = (int)(pix % Wo);
  const int ho = (int)((pix / Wo) % Ho);
  const int n = (int)(pix / ((long long)Wo * Ho));
  float m[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) m[j] = -INFINITY;
  for (int kh = 0; kh < 3; ++kh) {
    const int hi = ho * 2 + kh - 1;
    if (hi < 0 || hi >= H) continue;
    for (int kw = 0; kw < 3; ++kw) {
      const int wi = wo * 2 + kw - 1;
      if (wi < 0 || wi >= W) continue;
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(x + (((size_t)n * H + hi) * W + wi) * C + c8 * 8));
      const uint32_t uu[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = unpack_bf16(uu[j]);
        m[2 * j] = fmaxf(m[2 * j], f.x);
        m[2 * j + 1] = fmaxf(m[2 * j + 1], f.y);
      }
    }
  }
  *reinterpret_cast<uint4*>(y + pix * C + c8 * 8) =
      make_uint4(pack_bf16(m[0], m[1]), pack_bf16(m[2], m[3]), pack_bf16(m[4], m[5]), pack_bf16(m[6], m[7]));
}

// ------------------------------------------------------------------------------------------
// GroupNorm (32 groups).  Statistics are RAW accumulators acc[img][g] = (sum, sumsq) in double,
// normally produced by the epilogue of the conv that wrote the tensor (gemm_tc.cuh); this standalone
// kernel covers tensors that no GEMM produced (max-pool output, tests).  Double atomics make the
// result independent of the accumulation order to ~1e-16, i.e. bitwise stable once rounded to fp32.
// ------------------------------------------------------------------------------------------
constexpr int GN_THREADS = 256;

__global__ void __launch_bounds__(GN_THREADS)
gn_accumulate_kernel(const __nv_bfloat16* __restrict__ x, int HW, int C, int pre_relu,
                     int pix_per_block, double* __restrict__ acc) {
  const int CV = C / 8;            // 16B vectors per pixel (<= 256)
  const int PL = GN_THREADS / CV;  // pixel lanes
  const int cpg = C / 32;
  const int nsub = cpg >= 8 ? 1 : 8 / cpg;  // groups touched by one 8-channel vector
  const int chans_per_sub = 8 / nsub;
  const int cv = threadIdx.x % CV;
  const int pl = threadIdx.x / CV;
  const int img = blockIdx.y;
  const int p0 = blockIdx.x * pix_per_block;
  const int p1 = min(HW, p0 + pix_per_block);
  float s[4] = {0, 0, 0, 0}, q[4] = {0, 0, 0, 0};
  if (pl < PL) {
    const __nv_bfloat16* base = x + ((size_t)img * HW) * C + cv * 8;
#pragma unroll 4
    for (int p = p0 + pl; p < p1; p += PL) {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(base + (size_t)p * C));
      const uint32_t uu[4] = {u.x, u.y, u.z, u.w};
      float f[8];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 t = unpack_bf16(uu[j]);
        f[2 * j] = t.x;
        f[2 * j + 1] = t.y;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float v = pre_relu ? fmaxf(f[j], 0.f) : f[j];
        const int sub = j / chans_per_sub;
        s[sub] += v;
        q[sub] += v * v;
      }
    }
  }
  __shared__ float2 sm[GN_THREADS][4];
#pragma unroll
  for (int k = 0; k < 4; ++k) sm[threadIdx.x][k] = make_float2(s[k], q[k]);
  __syncthreads();
  __shared__ float2 sm2[256 * 4];
  const int entries = CV * nsub;
  for (int e = threadIdx.x; e < entries; e += GN_THREADS) {
    const int ecv = e / nsub, esub = e % nsub;
    float2 a = make_float2(0.f, 0.f);
    for (int l = 0; l < PL; ++l) {
      const float2 t = sm[l * CV + ecv][esub];
      a.x += t.x;
      a.y += t.y;
    }
    sm2[e] = a;
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    const int epg = entries / 32;
    float2 a = make_float2(0.f, 0.f);
    for (int e = threadIdx.x * epg; e < (threadIdx.x + 1) * epg; ++e) {
      a.x += sm2[e].x;
      a.y += sm2[e].y;
    }
    atomicAdd(&acc[((size_t)img * 32 + threadIdx.x) * 2 + 0], (double)a.x);
    atomicAdd(&acc[((size_t)img * 32 + threadIdx.x) * 2 + 1], (double)a.y);
  }
}

// layouts written by gn_apply
enum { LAYOUT_DENSE = 0, LAYOUT_PADDED = 1, LAYOUT_PHASE = 2 };

// One thread owns a fixed 8-channel vector (its scale/bias/mean/rstd live in registers) and walks
// over pixels; consecutive threads -> consecutive 16 B chunks (coalesced).  The kernel is issue-sensitive (about 10
// instructions per channel pair of useful math), so the pixel -> (h, w) bookkeeping is incremental (no division in
// the loop) and compiled out of the dense layout.
template <int LAYOUT, bool SUB, bool PRE>
__global__ void __launch_bounds__(256, 3)
gn_apply_kernel(const __nv_bfloat16* __restrict__ x, int Nimg, int H, int W, int C,
                const double* __restrict__ acc, int replica_stride, const float* __restrict__ scale,
                const float* __restrict__ bias, int post_relu,
                int pix_per_block, __nv_bfloat16* __restrict__ out, __nv_bfloat16* __restrict__ out_sub) {
  const int CV = C / 8;
  const int PL = 256 / CV;
  const int cv = threadIdx.x % CV, pl = threadIdx.x / CV;
  const int n = blockIdx.y;
  const int HW = H * W;
  const int c0 = cv * 8;
  const int cpg = C / 32;
  // one double-precision finalisation per group and CTA (mean, 1/sqrt(var + eps)), shared through smem
  __shared__ float2 s_stat[32];
  // per-thread affine parameters do not depend on the statistics: fetch them while the statistics are finalised
  __nv_bfloat162 sc2[4], bi2[4];  // scale / bias as packed bf16 (they are bf16 parameters in the reference)
  if (pl < PL) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      sc2[j] = __floats2bfloat162_rn(__ldg(scale + c0 + 2 * j), __ldg(scale + c0 + 2 * j + 1));
      bi2[j] = __floats2bfloat162_rn(__ldg(bias + c0 + 2 * j), __ldg(bias + c0 + 2 * j + 1));
    }
  }
  if (threadIdx.x < 32) {
    const int g = threadIdx.x;
    const double count = (double)HW * (double)cpg;
    double su = 0.0, sq = 0.0;
#pragma unroll
    for (int rep = 0; rep < 8; ++rep) {  // SNAPB200_GN_REPLICAS copies, fixed summation order
      su += acc[(size_t)rep * replica_stride + ((size_t)n * 32 + g) * 2];
      sq += acc[(size_t)rep * replica_stride + ((size_t)n * 32 + g) * 2 + 1];
    }
    const double mu = su / count;
    double var = sq / count - mu * mu;
    if (var < 0.0) var = 0.0;
    s_stat[g] = make_float2((float)mu, (float)(1.0 / sqrt(var + 1e-5)));
  }
  __syncthreads();
  if (pl >= PL) return;
  float mean[8], rstd[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float2 st = s_stat[(c0 + j) / cpg];
    mean[j] = st.x;
    rstd[j] = st.y;
  }
  const __nv_bfloat162 zero2 = __floats2bfloat162_rn(0.f, 0.f);
  const int p0 = blockIdx.x * pix_per_block;
  const int p1 = min(HW, p0 + pix_per_block);
  int p = p0 + pl;
  const uint4* xin = reinterpret_cast<const uint4*>(x + ((size_t)n * HW + p) * C + c0);
  const size_t xstep = (size_t)PL * CV;  // uint4 units between this thread's consecutive pixels
  // incremental (h, w) of pixel p; per step: w += dw, h += dh, one carry
  int h = 0, w = 0;
  const int dh = PL / W, dw = PL - dh * W;
  if (LAYOUT != LAYOUT_DENSE || SUB) {
    h = p / W;
    w = p - h * W;
  }
  const int Hq = H / 2 + 1, Wq = W / 2 + 1;
  uint4* odense = reinterpret_cast<uint4*>(out + ((size_t)n * HW + p) * C + c0);
  auto transform = [&](const uint4& uin) -> uint4 {
    const uint32_t uu[4] = {uin.x, uin.y, uin.z, uin.w};
    uint32_t ov[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t ww = PRE ? hmax2_bf16(uu[j], 0u) : uu[j];
      // resnet.py:39-41,57-69 with a bf16 dtype: standardise in fp32 -> bf16, * scale -> bf16, + bias -> bf16.
      // x - mean is taken straight from the packed halves (FHADD.BF16: bf16 operand, fp32 result, no unpacking);
      // the two bf16 ops run as packed HMUL2/HADD2.BF16 (exact product / sum, one rounding each).
      __nv_bfloat162 v = __floats2bfloat162_rn(bf16_lo_sub(ww, mean[2 * j]) * rstd[2 * j],
                                               bf16_hi_sub(ww, mean[2 * j + 1]) * rstd[2 * j + 1]);
      v = __hmul2_rn(v, sc2[j]);  // _rn: never contracted into an FMA (two roundings, like the reference)
      v = __hadd2_rn(v, bi2[j]);
      if (post_relu) v = __hmax2(v, zero2);
      ov[j] = *reinterpret_cast<uint32_t*>(&v);
    }
    return make_uint4(ov[0], ov[1], ov[2], ov[3]);
  };
  auto store = [&](const uint4& o, uint4* dense_ptr) {
    if (LAYOUT == LAYOUT_DENSE) {
      *dense_ptr = o;
    } else if (LAYOUT == LAYOUT_PADDED) {
      const size_t orow = ((size_t)n * (H + 2) + (h + 1)) * (W + 2) + (w + 1);
      *reinterpret_cast<uint4*>(out + orow * C + c0) = o;
    } else {
      const int hp = h + 1, wp = w + 1;
      const size_t plane = (size_t)((hp & 1) * 2 + (wp & 1)) * ((size_t)Nimg * Hq * Wq);
      const size_t orow = plane + ((size_t)n * Hq + (hp >> 1)) * Wq + (wp >> 1);
      *reinterpret_cast<uint4*>(out + orow * C + c0) = o;
    }
    if (SUB && (h & 1) == 0 && (w & 1) == 0) {
      const size_t srow = ((size_t)n * (H / 2) + (h >> 1)) * (W / 2) + (w >> 1);
      *reinterpret_cast<uint4*>(out_sub + srow * C + c0) = o;
    }
    if (LAYOUT != LAYOUT_DENSE || SUB) {
      w += dw;
      h += dh;
      if (w >= W) {
        w -= W;
        ++h;
      }
    }
  };
  constexpr int UN = 8;
  for (; p + (UN - 1) * PL < p1; p += UN * PL) {  // full batches: 8 loads in flight, no predicates
    uint4 u[UN];
#pragma unroll
    for (int i = 0; i < UN; ++i) u[i] = __ldg(xin + i * xstep);
#pragma unroll
    for (int i = 0; i < UN; ++i) store(transform(u[i]), odense + i * xstep);
    xin += UN * xstep;
    odense += UN * xstep;
  }
  for (; p < p1; p += PL) {
    store(transform(__ldg(xin)), odense);
    xin += xstep;
    odense += xstep;
  }
}

// x2 bilinear (half-pixel centres, edge clamp): [Nimg,h,w,C] -> [Nimg,2h,2w,C]
__global__ void upsample2x_kernel(const __nv_bfloat16* __restrict__ x, int Nimg, int h, int w, int C,
                                  __nv_bfloat16* __restrict__ y) {
  const int cv = C / 8;
  const int H = 2 * h, W = 2 * w;
  const long long total = (long long)Nimg * H * W * cv;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c8 = (int)(idx % cv);
  const long long pix = idx / cv;
  const int wo = (int)(pix % W);
  const int ho = (int)((pix / W) % H);
  const int n = (int)(pix / ((long long)W * H));
  // source coordinate (o + 0.5)/2 - 0.5 = o/2 - 0.25
  const int h0 = (ho & 1) ? (ho >> 1) : (ho >> 1) - 1;
  const int w0 = (wo & 1) ? (wo >> 1) : (wo >> 1) - 1;
  const float fh = (ho & 1) ? 0.25f : 0.75f;  // weight of the upper tap h0+1
  const float fw = (wo & 1) ? 0.25f : 0.75f;
  const int ha = max(h0, 0), hb = min(h0 + 1, h - 1);
  const int wa = max(w0, 0), wb = min(w0 + 1, w - 1);
  const __nv_bfloat16* base = x + (size_t)n * h * w * C + c8 * 8;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const int hs[2] = {ha, hb}, ws[2] = {wa, wb};
  const float whs[2] = {1.f - fh, fh}, wws[2] = {1.f - fw, fw};
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      const float wt = whs[a] * wws[b];
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(base + ((size_t)hs[a] * w + ws[b]) * C));
      const uint32_t uu[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 t = unpack_bf16(uu[j]);
        acc[2 * j] += wt * t.x;
        acc[2 * j + 1] += wt * t.y;
      }
    }
  *reinterpret_cast<uint4*>(y + pix * C + c8 * 8) =
      make_uint4(pack_bf16(acc[0], acc[1]), pack_bf16(acc[2], acc[3]), pack_bf16(acc[4], acc[5]),
                 pack_bf16(acc[6], acc[7]));
}

// crop top-left (h,w) of [Nimg,Hs,Ws,C] with optional ReLU -> dense [Nimg,h,w,C]
__global__ void crop_relu_kernel(const __nv_bfloat16* __restrict__ x, int Nimg, int Hs, int Ws, int C,
                                 int h, int w, int relu, __nv_bfloat16* __restrict__ y) {
  const int cv = C / 8;
  const long long total = (long long)Nimg * h * w * cv;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c8 = (int)(idx % cv);
  const long long pix = idx / cv;
  const int ww = (int)(pix % w);
  const int hh = (int)((pix / w) % h);
  const int n = (int)(pix / ((long long)w * h));
  uint4 u = __ldg(reinterpret_cast<const uint4*>(x + (((size_t)n * Hs + hh) * Ws + ww) * C + c8 * 8));
  if (relu) {
    uint32_t uu[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float2 t = unpack_bf16(uu[j]);
      uu[j] = pack_bf16(fmaxf(t.x, 0.f), fmaxf(t.y, 0.f));
    }
    u = make_uint4(uu[0], uu[1], uu[2], uu[3]);
  }
  *reinterpret_cast<uint4*>(y + pix * C + c8 * 8) = u;
}

static inline unsigned blocks_for(long long total, int threads) {
  return (unsigned)((total + threads - 1) / threads);
}

}  // namespace snapb200

using namespace snapb200;

extern "C" {

/* Batched StdConv weight standardisation (resnet.py:34-41,73-79; eps 1e-10) + relayout to the GEMM
   B operand.  `descs` is a DEVICE array of SnapWeightDesc; mapA / mapB are DEVICE int4 work lists
   (desc, cout tile, k split | k tile, 0) built by the host once per parameter set. */
int snapb200_std_weights_batched(const void* descs, const void* mapA, int nA, const void* mapB, int nB,
                                 void* stream) {
  SNAP_REQUIRE(descs && mapB && nB > 0, "null pointer");
  dim3 block(32, 8);
  if (nA > 0) {
    SNAP_REQUIRE(mapA != nullptr, "null mapA");
    std_weights_stats_kernel<<<nA, block, 0, (cudaStream_t)stream>>>((const WeightDesc*)descs,
                                                                    (const int4*)mapA);
    int rc = check_launch("std_weights_stats_kernel");
    if (rc) return rc;
  }
  std_weights_write_kernel<<<nB, block, 0, (cudaStream_t)stream>>>((const WeightDesc*)descs,
                                                                  (const int4*)mapB, 1e-10f);
  return check_launch("std_weights_write_kernel");
}

int snapb200_root_pack_image(const float* images, int Nimg, int H, int W, int Hp, int Wp, int pad, int cp, int Hq,
                             int Wq, void* out, void* stream) {
  SNAP_REQUIRE(images && out && Nimg > 0, "bad arguments");
  SNAP_REQUIRE(cp == 4 || cp == 8, "cp must be 4 or 8");
  SNAP_REQUIRE(Hq >= Hp + 2 * pad && Wq >= Wp + 2 * pad && (Wq * cp * 2) % 16 == 0, "bad packed geometry");
  const long long total = (long long)Nimg * Hq * Wq;
  root_pack_image_kernel<<<blocks_for(total, 256), 256, 0, (cudaStream_t)stream>>>(images, Nimg, H, W, Hp, Wp, pad, cp,
                                                                                  Hq, Wq, (__nv_bfloat16*)out);
  return check_launch("root_pack_image_kernel");
}

int snapb200_root_pack_weights(const void* b_std, int Cout, int ldb, int KH, int KW, int cp, void* out, void* stream) {
  SNAP_REQUIRE(b_std && out && KW * cp <= 32 && KH * KW * 3 <= ldb, "bad arguments");
  const long long total = (long long)Cout * KH * 32;
  root_pack_weights_kernel<<<blocks_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)b_std, Cout, ldb, KH, KW, cp, (__nv_bfloat16*)out);
  return check_launch("root_pack_weights_kernel");
}

int snapb200_root_im2col(const float* images, int Nimg, int H, int W, int Hp, int Wp, int KH, int KW,
                         int stride, int pad, void* out_bf16, int Kp, void* stream) {
  SNAP_REQUIRE(images && out_bf16, "null pointer");
  SNAP_REQUIRE(Kp % 32 == 0 && Kp >= KH * KW * 3, "Kp must be a multiple of 32 covering KH*KW*3");
  SNAP_REQUIRE(Hp >= H && Wp >= W && stride >= 1, "bad padded size");
  const int Ho = (Hp + 2 * pad - KH) / stride + 1, Wo = (Wp + 2 * pad - KW) / stride + 1;
  const long long total_warps = (long long)Nimg * Ho * Wo;
  long long blocks = (total_warps + 7) / 8;
  if (blocks > 148LL * 64) blocks = 148LL * 64;  // grid-stride over pixels
  SNAP_REQUIRE(total_warps < (1ll << 31) && (KW == 7 || KW == 3), "root conv must be 7x7 or 3x3; too many pixels");
  if (KW == 7)
    root_im2col_kernel<7><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        images, Nimg, H, W, Hp, Wp, KH, stride, pad, Ho, Wo, (__nv_bfloat16*)out_bf16, Kp);
  else
    root_im2col_kernel<3><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        images, Nimg, H, W, Hp, Wp, KH, stride, pad, Ho, Wo, (__nv_bfloat16*)out_bf16, Kp);
  return check_launch("root_im2col_kernel");
}

int snapb200_maxpool3x3s2(const void* x, int Nimg, int H, int W, int C, void* y, void* stream) {
  SNAP_REQUIRE(x && y, "null pointer");
  SNAP_REQUIRE(C % 8 == 0, "C must be a multiple of 8");
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  const long long total = (long long)Nimg * Ho * Wo * (C / 8);
  maxpool3x3s2_kernel<<<blocks_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)x, Nimg, H, W, C, (__nv_bfloat16*)y, Ho, Wo);
  return check_launch("maxpool3x3s2_kernel");
}

/* Accumulate GroupNorm(32 groups) raw statistics of x [Nimg, HW, C] (optionally of relu(x)) into
   acc[img][g] = (sum, sumsq) doubles.  acc must be zeroed by the caller; conv outputs get the same
   accumulators from the GEMM epilogue (SnapGemmParams.gn_acc) without this extra pass. */
int snapb200_gn_stats(const void* x, int Nimg, int HW, int C, int pre_relu, double* acc, void* stream) {
  SNAP_REQUIRE(x && acc, "null pointer");
  SNAP_REQUIRE(C % 64 == 0 && C <= 2048, "GroupNorm kernel needs C %% 64 == 0 and C <= 2048 (got %d)", C);
  int ppb = 64 * (GN_THREADS / (C / 8));  // 64 pixels per thread lane
  if (ppb > HW) ppb = HW > 0 ? HW : 1;
  dim3 grid((HW + ppb - 1) / ppb, Nimg);
  gn_accumulate_kernel<<<grid, GN_THREADS, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, HW, C,
                                                                     pre_relu, ppb, acc);
  return check_launch("gn_accumulate_kernel");
}

/* y = post_relu?( (pre_relu?(x) - mean) * rstd * scale + bias ) with mean/rstd derived from the raw
   accumulators acc[img][g] = (sum, sumsq) (eps 1e-5, resnet.py:34-70), written in `layout`
   (0 dense [Nimg*H*W, C]; 1 zero-bordered [Nimg,(H+2),(W+2),C]; 2 phase-split);
   out_sub (optional) additionally receives the even-pixel subsample, dense [Nimg,H/2,W/2,C]. */
int snapb200_gn_apply(const void* x, int Nimg, int H, int W, int C, const double* acc, int replica_stride,
                      const float* scale, const float* bias, int pre_relu, int post_relu, int layout,
                      void* out, void* out_sub, void* stream) {
  SNAP_REQUIRE(x && acc && scale && bias && out, "null pointer");
  SNAP_REQUIRE(replica_stride >= Nimg * 64, "replica_stride must cover [Nimg][32][2] doubles");
  SNAP_REQUIRE(C % 64 == 0 && C <= 2048, "C must be a multiple of 64, <= 2048");
  SNAP_REQUIRE(layout >= 0 && layout <= 2, "bad layout");
  SNAP_REQUIRE((layout != LAYOUT_PHASE && out_sub == nullptr) || (H % 2 == 0 && W % 2 == 0),
               "phase / subsampled layouts need even H, W");
  const int HW = H * W;
  // pixels per thread lane: 32 (amortises the per-CTA statistics finalisation) unless that leaves the GPU short of
  // CTAs (8 x 148 resident), then 16 / 8
  const int PL = 256 / (C / 8);
  int ppl = 32;
  while (ppl > 8 && (long long)Nimg * ((HW + PL * ppl - 1) / (PL * ppl)) < 8LL * num_sms()) ppl >>= 1;
  int ppb = ppl * PL;
  if (ppb > HW) ppb = HW;
  dim3 grid((HW + ppb - 1) / ppb, Nimg);
#define SNAP_GN_APPLY(L_, S_, P_)                                                                 \
  gn_apply_kernel<L_, S_, P_><<<grid, 256, 0, (cudaStream_t)stream>>>(                               \
      (const __nv_bfloat16*)x, Nimg, H, W, C, acc, replica_stride, scale, bias, post_relu, ppb,     \
      (__nv_bfloat16*)out, (__nv_bfloat16*)out_sub)
#define SNAP_GN_APPLY2(L_, S_) \
  do { if (pre_relu) SNAP_GN_APPLY(L_, S_, true); else SNAP_GN_APPLY(L_, S_, false); } while (0)
  if (layout == LAYOUT_DENSE) {
    if (out_sub != nullptr) SNAP_GN_APPLY2(LAYOUT_DENSE, true); else SNAP_GN_APPLY2(LAYOUT_DENSE, false);
  } else if (layout == LAYOUT_PADDED) {
    if (out_sub != nullptr) SNAP_GN_APPLY2(LAYOUT_PADDED, true); else SNAP_GN_APPLY2(LAYOUT_PADDED, false);
  } else {
    if (out_sub != nullptr) SNAP_GN_APPLY2(LAYOUT_PHASE, true); else SNAP_GN_APPLY2(LAYOUT_PHASE, false);
  }
#undef SNAP_GN_APPLY2
#undef SNAP_GN_APPLY
  return check_launch("gn_apply_kernel");
}

int snapb200_upsample2x(const void* x, int Nimg, int h, int w, int C, void* y, void* stream) {
  SNAP_REQUIRE(x && y && C % 8 == 0, "bad arguments");
  const long long total = (long long)Nimg * 4 * h * w * (C / 8);
  upsample2x_kernel<<<blocks_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)x, Nimg, h, w, C, (__nv_bfloat16*)y);
  return check_launch("upsample2x_kernel");
}

int snapb200_crop_relu(const void* x, int Nimg, int Hs, int Ws, int C, int h, int w, int relu, void* y,
                       void* stream) {
  SNAP_REQUIRE(x && y && C % 8 == 0 && h <= Hs && w <= Ws, "bad arguments");
  const long long total = (long long)Nimg * h * w * (C / 8);
  crop_relu_kernel<<<blocks_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)x, Nimg, Hs, Ws, C, h, w, relu, (__nv_bfloat16*)y);
  return check_launch("crop_relu_kernel");
}

}  // extern "C"
