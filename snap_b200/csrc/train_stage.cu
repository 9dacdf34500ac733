// Backward building blocks of the 'resnet_stage' semantic decoder (snap/models/semantic_net.py:153-161 with
// snap/models/resnet.py:34-155; BASELINE configs[4], snap/configs/train_semantics.py:27-36) beyond the dense layers of
// train_head.cu:
//
//   GroupNorm (+ReLU) backward (resnet.py:46-70)      -> gn_bwd_reduce_kernel, gn_bwd_apply_kernel, gn_bwd_params_kernel
//   StdConv weight-standardisation backward (:34-41)  -> stdconv_bwd_kernel
//   B operand of dX = dY W^T for 1x1 / 3x3 kernels    -> wt_segments_kernel (transposed taps of the standardised kernel)
//
// The convolutions themselves reuse existing kernels: dX of a 1x1 conv is the tcgen05 GEMM engine with the transposed
// kernel, dX of the 3x3 conv is the engine's 9-segment mode over the zero-bordered dY with mirrored row offsets, dW is
// snapb200_dense_wgrad (nine row-shifted calls for the 3x3 kernel).  Closed forms and the launch decomposition are the
// ones checked against torch autograd in tools/design/backward_formulas.py (tests/test_backward_design_cpu.py).
//
// GroupNorm backward, per image n and group g over m = H*W*C/32 elements, with xhat = (x - mu) * rstd,
// a = relu(xhat * scale + bias) and d = dy * [a > 0]:
//   dbias_c  = sum d            dscale_c = sum d * xhat
//   s1(n,g)  = sum_{c in g} scale_c * sum_p d        s2(n,g) = sum_{c in g} scale_c * sum_p d * xhat
//   dx       = rstd * (d * scale - s1 / m - xhat * s2 / m)
// so ONE reduction per (image, channel) of (sum d, sum d * xhat) feeds both the parameter gradients and the group sums.
// The ReLU mask is recomputed with the forward's bf16 rounding chain (gn_apply_kernel), i.e. it is exactly the mask of
// the activation the forward produced.  Reductions: fixed-order tree inside a CTA, double atomics across CTAs.
#include <cuda_bf16.h>
#include <math.h>

#include "common.cuh"
#include "host_common.h"

namespace snapb200 {

namespace {

struct GnFwdStat {
  float mean, rstd;
};

// finalise the forward statistics of (image n, group g) from the raw (sum, sumsq) double accumulators, exactly as
// gn_apply_kernel does (8 replicas, fixed order, eps 1e-5 inside the square root)
__device__ __forceinline__ GnFwdStat gn_finalize(const double* __restrict__ acc, int replica_stride, int n, int g,
                                                 double count) {
  double su = 0.0, sq = 0.0;
#pragma unroll
  for (int rep = 0; rep < 8; ++rep) {
    su += acc[(size_t)rep * replica_stride + ((size_t)n * 32 + g) * 2];
    sq += acc[(size_t)rep * replica_stride + ((size_t)n * 32 + g) * 2 + 1];
  }
  const double mu = su / count;
  double var = sq / count - mu * mu;
  if (var < 0.0) var = 0.0;
  GnFwdStat s;
  s.mean = (float)mu;
  s.rstd = (float)(1.0 / sqrt(var + 1e-5));
  return s;
}

// forward value of one channel pair (bf16 rounding chain of gn_apply_kernel) -> is the ReLU open?
__device__ __forceinline__ void relu_open(float xh0, float xh1, __nv_bfloat162 sc, __nv_bfloat162 bi, bool& o0, bool& o1) {
  __nv_bfloat162 v = __floats2bfloat162_rn(xh0, xh1);
  v = __hmul2_rn(v, sc);
  v = __hadd2_rn(v, bi);
  o0 = __low2float(v) > 0.f;
  o1 = __high2float(v) > 0.f;
}

// cotangent of the (activated) GroupNorm output at pixel p of image n, channels [c0, c0 + 8): read from the dense layout
// or from the phase-split layout of a stride-2 3x3 conv input (gn_apply's LAYOUT_PHASE: [2, 2, Nimg, H/2+1, W/2+1, C]);
// dy_sub (optional, dense [Nimg, H/2, W/2, C]) is added at the even pixels (the stride-2 projection shortcut reads the
// even-pixel subsample of the same activation, resnet.py:121-122)
__device__ __forceinline__ void load_dy(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ dy_sub,
                                        int dy_phase, int Nimg, int H, int W, int C, int n, int p, int c0,
                                        uint32_t (&du)[4]) {
  const int h = p / W, w = p - h * W;
  size_t row = (size_t)n * H * W + p;
  if (dy_phase) {
    const int hp = h + 1, wp = w + 1, Hq = H / 2 + 1, Wq = W / 2 + 1;
    row = (size_t)((hp & 1) * 2 + (wp & 1)) * ((size_t)Nimg * Hq * Wq) + ((size_t)n * Hq + (hp >> 1)) * Wq + (wp >> 1);
  }
  const uint4 dv = __ldg(reinterpret_cast<const uint4*>(dy + row * C + c0));
  du[0] = dv.x; du[1] = dv.y; du[2] = dv.z; du[3] = dv.w;
  if (dy_sub != nullptr && (h & 1) == 0 && (w & 1) == 0) {
    const size_t srow = ((size_t)n * (H / 2) + (h >> 1)) * (W / 2) + (w >> 1);
    const uint4 sv = __ldg(reinterpret_cast<const uint4*>(dy_sub + srow * C + c0));
    const uint32_t su[4] = {sv.x, sv.y, sv.z, sv.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 a = unpack_bf16(du[j]), b = unpack_bf16(su[j]);
      du[j] = pack_bf16(a.x + b.x, a.y + b.y);
    }
  }
}

}  // namespace

// accb [Nimg, C, 2] (double) += per-(image, channel) sums of (d, d * xhat).  grid (pixel blocks, Nimg), 256 threads:
// thread = (pixel lane pl, channel vector cv of 8 channels).
__global__ void __launch_bounds__(256)
gn_bwd_reduce_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy,
                     const __nv_bfloat16* __restrict__ dy_sub, int dy_phase, int Nimg, int H, int W, int C,
                     const double* __restrict__ acc, int replica_stride, const float* __restrict__ scale,
                     const float* __restrict__ bias, int pre_relu, int post_relu, int pix_per_block,
                     double* __restrict__ accb) {
  const int CV = C / 8, PL = 256 / CV;
  const int cv = threadIdx.x % CV, pl = threadIdx.x / CV;
  const int n = blockIdx.y, c0 = cv * 8, cpg = C / 32, HW = H * W;
  __shared__ float2 s_stat[32];
  __shared__ float s_part[256][17];   // [thread][8 x (d, d*xhat)], padded against bank conflicts
  if (threadIdx.x < 32) {
    const GnFwdStat st = gn_finalize(acc, replica_stride, n, threadIdx.x, (double)HW * (double)cpg);
    s_stat[threadIdx.x] = make_float2(st.mean, st.rstd);
  }
  __syncthreads();
  float sd[8], sx[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) sd[j] = sx[j] = 0.f;
  if (pl < PL) {
    float mean[8], rstd[8];
    __nv_bfloat162 sc2[4], bi2[4];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float2 st = s_stat[(c0 + j) / cpg];
      mean[j] = st.x;
      rstd[j] = st.y;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      sc2[j] = __floats2bfloat162_rn(__ldg(scale + c0 + 2 * j), __ldg(scale + c0 + 2 * j + 1));
      bi2[j] = __floats2bfloat162_rn(__ldg(bias + c0 + 2 * j), __ldg(bias + c0 + 2 * j + 1));
    }
    const int p0 = blockIdx.x * pix_per_block, p1 = min(HW, p0 + pix_per_block);
    for (int p = p0 + pl; p < p1; p += PL) {
      const size_t off = ((size_t)n * HW + p) * C + c0;
      const uint4 xv = __ldg(reinterpret_cast<const uint4*>(x + off));
      uint32_t du[4];
      load_dy(dy, dy_sub, dy_phase, Nimg, H, W, C, n, p, c0, du);
      const uint32_t xu[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float2 xf = unpack_bf16(xu[j]);
        const float2 df = unpack_bf16(du[j]);
        if (pre_relu) {  // GroupNorm of relu(x) (FPN: image_encoder.py:86-88); the statistics in acc are those of relu(x)
          xf.x = fmaxf(xf.x, 0.f);
          xf.y = fmaxf(xf.y, 0.f);
        }
        const float h0 = (xf.x - mean[2 * j]) * rstd[2 * j], h1 = (xf.y - mean[2 * j + 1]) * rstd[2 * j + 1];
        bool o0 = true, o1 = true;
        if (post_relu) relu_open(h0, h1, sc2[j], bi2[j], o0, o1);
        const float d0 = o0 ? df.x : 0.f, d1 = o1 ? df.y : 0.f;
        sd[2 * j] += d0;
        sx[2 * j] += d0 * h0;
        sd[2 * j + 1] += d1;
        sx[2 * j + 1] += d1 * h1;
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    s_part[threadIdx.x][2 * j] = sd[j];
    s_part[threadIdx.x][2 * j + 1] = sx[j];
  }
  __syncthreads();
  // fixed-order sum over the PL pixel lanes of every (channel, statistic); 2 * C <= 512 entries for 256 threads
  for (int e = threadIdx.x; e < 2 * C; e += 256) {
    const int c = e >> 1, which = e & 1;
    const int tcv = c / 8, j = c % 8;
    float s = 0.f;
    for (int l = 0; l < PL; ++l) s += s_part[l * CV + tcv][2 * j + which];
    atomicAdd(&accb[((size_t)n * C + c) * 2 + which], (double)s);
  }
}

// dx = rstd * (d * scale - s1 / m - xhat * s2 / m) (+ add), written dense [Nimg*HW, C] or zero-bordered
// [Nimg, H+2, W+2, C] (the layout the 3x3 dX GEMM and the shifted dW products read)
// OUT: 0 dense [Nimg*H*W, C]; 1 zero-bordered [Nimg, H+2, W+2, C]; 2 bottom/right zero-extended [Nimg, H+1, W+1, C] (the
// row indexing of the stride-2 3x3 conv's output GEMM, whose A operand is the phase-split input with H/2+1 x W/2+1 planes)
template <int OUT>
__global__ void __launch_bounds__(256)
gn_bwd_apply_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy,
                    const __nv_bfloat16* __restrict__ dy_sub, int dy_phase, int Nimg,
                    const __nv_bfloat16* __restrict__ add, int H, int W, int C, const double* __restrict__ acc,
                    int replica_stride, const float* __restrict__ scale, const float* __restrict__ bias, int pre_relu,
                    int post_relu, const double* __restrict__ accb, int pix_per_block, __nv_bfloat16* __restrict__ dx) {
  const int CV = C / 8, PL = 256 / CV;
  const int cv = threadIdx.x % CV, pl = threadIdx.x / CV;
  const int n = blockIdx.y, c0 = cv * 8, cpg = C / 32, HW = H * W;
  __shared__ float2 s_stat[32];
  __shared__ float2 s_grp[32];   // (s1 / m, s2 / m)
  if (threadIdx.x < 32) {
    const int g = threadIdx.x;
    const double m = (double)HW * (double)cpg;
    const GnFwdStat st = gn_finalize(acc, replica_stride, n, g, m);
    s_stat[g] = make_float2(st.mean, st.rstd);
    double s1 = 0.0, s2 = 0.0;
    for (int k = 0; k < cpg; ++k) {
      const int c = g * cpg + k;
      const double sc = (double)__bfloat162float(__float2bfloat16_rn(__ldg(scale + c)));
      s1 += sc * accb[((size_t)n * C + c) * 2];
      s2 += sc * accb[((size_t)n * C + c) * 2 + 1];
    }
    s_grp[g] = make_float2((float)(s1 / m), (float)(s2 / m));
  }
  __syncthreads();
  if (pl >= PL) return;
  float mean[8], rstd[8], g1[8], g2[8], scf[8];
  __nv_bfloat162 sc2[4], bi2[4];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int g = (c0 + j) / cpg;
    mean[j] = s_stat[g].x;
    rstd[j] = s_stat[g].y;
    g1[j] = s_grp[g].x;
    g2[j] = s_grp[g].y;
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    sc2[j] = __floats2bfloat162_rn(__ldg(scale + c0 + 2 * j), __ldg(scale + c0 + 2 * j + 1));
    bi2[j] = __floats2bfloat162_rn(__ldg(bias + c0 + 2 * j), __ldg(bias + c0 + 2 * j + 1));
    scf[2 * j] = __low2float(sc2[j]);
    scf[2 * j + 1] = __high2float(sc2[j]);
  }
  const int p0 = blockIdx.x * pix_per_block, p1 = min(HW, p0 + pix_per_block);
  for (int p = p0 + pl; p < p1; p += PL) {
    const size_t off = ((size_t)n * HW + p) * C + c0;
    const uint4 xv = __ldg(reinterpret_cast<const uint4*>(x + off));
    uint32_t du[4];
    load_dy(dy, dy_sub, dy_phase, Nimg, H, W, C, n, p, c0, du);
    uint4 av = make_uint4(0u, 0u, 0u, 0u);
    if (add != nullptr) av = __ldg(reinterpret_cast<const uint4*>(add + off));
    const uint32_t xu[4] = {xv.x, xv.y, xv.z, xv.w}, au[4] = {av.x, av.y, av.z, av.w};
    uint32_t ov[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float2 xf = unpack_bf16(xu[j]);
      const float2 df = unpack_bf16(du[j]), af = unpack_bf16(au[j]);
      const bool p0 = !pre_relu || xf.x > 0.f, p1 = !pre_relu || xf.y > 0.f;   // the ReLU in front of the GroupNorm
      if (pre_relu) {
        xf.x = fmaxf(xf.x, 0.f);
        xf.y = fmaxf(xf.y, 0.f);
      }
      const float h0 = (xf.x - mean[2 * j]) * rstd[2 * j], h1 = (xf.y - mean[2 * j + 1]) * rstd[2 * j + 1];
      bool o0 = true, o1 = true;
      if (post_relu) relu_open(h0, h1, sc2[j], bi2[j], o0, o1);
      const float d0 = o0 ? df.x : 0.f, d1 = o1 ? df.y : 0.f;
      float r0 = rstd[2 * j] * (d0 * scf[2 * j] - g1[2 * j] - h0 * g2[2 * j]);
      float r1 = rstd[2 * j + 1] * (d1 * scf[2 * j + 1] - g1[2 * j + 1] - h1 * g2[2 * j + 1]);
      r0 = (p0 ? r0 : 0.f) + af.x;
      r1 = (p1 ? r1 : 0.f) + af.y;
      ov[j] = pack_bf16(r0, r1);
    }
    size_t orow = (size_t)n * HW + p;
    if (OUT == 1) {
      const int h = p / W, w = p - h * W;
      orow = ((size_t)n * (H + 2) + (h + 1)) * (W + 2) + (w + 1);
    } else if (OUT == 2) {
      const int h = p / W, w = p - h * W;
      orow = ((size_t)n * (H + 1) + h) * (W + 1) + w;
    }
    *reinterpret_cast<uint4*>(dx + orow * C + c0) = make_uint4(ov[0], ov[1], ov[2], ov[3]);
  }
}

// dscale_c = sum_n accb[n][c][1], dbias_c = sum_n accb[n][c][0]
__global__ void gn_bwd_params_kernel(const double* __restrict__ accb, int Nimg, int C, float* __restrict__ dscale,
                                     float* __restrict__ dbias) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double sb = 0.0, ss = 0.0;
  for (int n = 0; n < Nimg; ++n) {
    sb += accb[((size_t)n * C + c) * 2];
    ss += accb[((size_t)n * C + c) * 2 + 1];
  }
  dscale[c] = (float)ss;
  dbias[c] = (float)sb;
}

// out[cin, tap * Cout + cout] = in[cout, tap * Cin + cin]: the engine's B operand [N = Cin, K = taps * Cout] of
// dX = sum_tap dY[. - off_tap] W_tap^T from the forward operand [Cout, taps * Cin] (taps = 1: a plain transpose)
__global__ void wt_segments_kernel(const __nv_bfloat16* __restrict__ in, int ld_in, int Cout, int Cin, int taps,
                                   __nv_bfloat16* __restrict__ out, int ld_out) {
  const int total = Cin * taps * Cout;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int cout = i % Cout;
  const int tap = (i / Cout) % taps;
  const int cin = i / (Cout * taps);
  out[(size_t)cin * ld_out + tap * Cout + cout] = in[(size_t)cout * ld_in + tap * Cin + cin];
}

// StdConv backward: one CTA (128 threads) per output channel o over the K = kh*kw*in elements of column o of the
// [K, Cout] kernel: ws = (w - mu) * rstd (eps 1e-10), dw = rstd * (dws - mean(dws) - ws * mean(dws * ws)).
__global__ void __launch_bounds__(128)
stdconv_bwd_kernel(const float* __restrict__ w, const float* __restrict__ dws, int K, int Cout, float* __restrict__ dw) {
  __shared__ double s_red[128];
  __shared__ double s_val[4];
  const int o = blockIdx.x, t = threadIdx.x;
  auto block_sum = [&](double v, int slot) {
    s_red[t] = v;
    __syncthreads();
    for (int s = 64; s > 0; s >>= 1) {
      if (t < s) s_red[t] += s_red[t + s];
      __syncthreads();
    }
    if (t == 0) s_val[slot] = s_red[0];
    __syncthreads();
  };
  double a = 0.0;
  for (int k = t; k < K; k += 128) a += (double)w[(size_t)k * Cout + o];
  block_sum(a, 0);
  const double mu = s_val[0] / K;
  a = 0.0;
  for (int k = t; k < K; k += 128) {
    const double c = (double)w[(size_t)k * Cout + o] - mu;
    a += c * c;
  }
  block_sum(a, 1);
  const double rstd = 1.0 / sqrt(s_val[1] / K + 1e-10);
  double sd = 0.0, sdw = 0.0;
  for (int k = t; k < K; k += 128) {
    const double d = (double)dws[(size_t)k * Cout + o];
    sd += d;
    sdw += d * ((double)w[(size_t)k * Cout + o] - mu) * rstd;
  }
  block_sum(sd, 2);
  block_sum(sdw, 3);
  const double md = s_val[2] / K, mdw = s_val[3] / K;
  for (int k = t; k < K; k += 128) {
    const double ws = ((double)w[(size_t)k * Cout + o] - mu) * rstd;
    dw[(size_t)k * Cout + o] = (float)(rstd * ((double)dws[(size_t)k * Cout + o] - md - ws * mdw));
  }
}

// Backward of the x2 bilinear up-sampling of the FPN (image_encoder.py:86-91; upsample2x_kernel): dx[n, hc, wc, :] = sum
// over the (at most 4 x 4) fine pixels whose clamped taps hit the coarse pixel, with the forward's weights.  One thread
// per (coarse pixel, 8 channels); a gather, so the result is deterministic.
__global__ void upsample2x_bwd_kernel(const __nv_bfloat16* __restrict__ dy, int Nimg, int h, int w, int C,
                                      __nv_bfloat16* __restrict__ dx) {
  const int cv = C / 8;
  const int H = 2 * h, W = 2 * w;
  const long long total = (long long)Nimg * h * w * cv;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c8 = (int)(idx % cv);
  const long long pix = idx / cv;
  const int wc = (int)(pix % w);
  const int hc = (int)((pix / w) % h);
  const int n = (int)(pix / ((long long)w * h));
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int ho = max(2 * hc - 1, 0); ho <= min(2 * hc + 2, H - 1); ++ho) {
    const int h0 = (ho & 1) ? (ho >> 1) : (ho >> 1) - 1;
    const float fh = (ho & 1) ? 0.25f : 0.75f;
    const float wh = (max(h0, 0) == hc ? 1.f - fh : 0.f) + (min(h0 + 1, h - 1) == hc ? fh : 0.f);
    if (wh == 0.f) continue;
    for (int wo = max(2 * wc - 1, 0); wo <= min(2 * wc + 2, W - 1); ++wo) {
      const int w0 = (wo & 1) ? (wo >> 1) : (wo >> 1) - 1;
      const float fw = (wo & 1) ? 0.25f : 0.75f;
      const float ww = (max(w0, 0) == wc ? 1.f - fw : 0.f) + (min(w0 + 1, w - 1) == wc ? fw : 0.f);
      if (ww == 0.f) continue;
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(dy + (((size_t)n * H + ho) * W + wo) * C + c8 * 8));
      const uint32_t uu[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 t = unpack_bf16(uu[j]);
        acc[2 * j] += wh * ww * t.x;
        acc[2 * j + 1] += wh * ww * t.y;
      }
    }
  }
  *reinterpret_cast<uint4*>(dx + pix * C + c8 * 8) =
      make_uint4(pack_bf16(acc[0], acc[1]), pack_bf16(acc[2], acc[3]), pack_bf16(acc[4], acc[5]), pack_bf16(acc[6], acc[7]));
}

// Backward of the 3x3 / stride-2 / pad-1 max pool of the root block (resnet.py:99): dx[n, h, w, :] = sum of dy over the
// (at most 2 x 2) windows that contain the pixel AND whose first maximum (row-major scan, strict >) it is.  A gather: one
// thread per (input pixel, 8 channels) re-derives the arg-max of each candidate window, so the result is deterministic.
__global__ void maxpool3x3s2_bwd_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy, int Nimg,
                                        int H, int W, int C, __nv_bfloat16* __restrict__ dx) {
  const int cv = C / 8;
  const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
  const long long total = (long long)Nimg * H * W * cv;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c8 = (int)(idx % cv);
  const long long pix = idx / cv;
  const int w = (int)(pix % W);
  const int h = (int)((pix / W) % H);
  const int n = (int)(pix / ((long long)W * H));
  const __nv_bfloat16* xb = x + (size_t)n * H * W * C + c8 * 8;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  // window ho covers input rows 2ho-1 .. 2ho+1, so row h lies in the windows ho = h/2 .. (h+1)/2 (likewise for columns)
  for (int ho = h / 2; ho <= min((h + 1) / 2, Ho - 1); ++ho) {
    for (int wo = w / 2; wo <= min((w + 1) / 2, Wo - 1); ++wo) {
      float best[8];
      int bh[8], bw[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        best[j] = -INFINITY;
        bh[j] = bw[j] = -1;
      }
      for (int hi = max(2 * ho - 1, 0); hi <= min(2 * ho + 1, H - 1); ++hi)
        for (int wi = max(2 * wo - 1, 0); wi <= min(2 * wo + 1, W - 1); ++wi) {
          const uint4 u = __ldg(reinterpret_cast<const uint4*>(xb + ((size_t)hi * W + wi) * C));
          const uint32_t uu[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 t = unpack_bf16(uu[j]);
            if (t.x > best[2 * j]) { best[2 * j] = t.x; bh[2 * j] = hi; bw[2 * j] = wi; }
            if (t.y > best[2 * j + 1]) { best[2 * j + 1] = t.y; bh[2 * j + 1] = hi; bw[2 * j + 1] = wi; }
          }
        }
      const uint4 g = __ldg(reinterpret_cast<const uint4*>(dy + (((size_t)n * Ho + ho) * Wo + wo) * C + c8 * 8));
      const uint32_t gu[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 t = unpack_bf16(gu[j]);
        if (bh[2 * j] == h && bw[2 * j] == w) acc[2 * j] += t.x;
        if (bh[2 * j + 1] == h && bw[2 * j + 1] == w) acc[2 * j + 1] += t.y;
      }
    }
  }
  *reinterpret_cast<uint4*>(dx + pix * C + c8 * 8) =
      make_uint4(pack_bf16(acc[0], acc[1]), pack_bf16(acc[2], acc[3]), pack_bf16(acc[4], acc[5]), pack_bf16(acc[6], acc[7]));
}

}  // namespace snapb200

using namespace snapb200;

extern "C" {

int snapb200_gn_backward(const void* x, const void* dy, const void* dy_sub, int dy_phase, const void* add, int Nimg,
                         int H, int W, int C, const double* acc, int replica_stride, const float* scale,
                         const float* bias, int pre_relu, int post_relu, int out_layout, double* accb, void* dx,
                         float* dscale, float* dbias, void* stream) {
  SNAP_REQUIRE(x && dy && acc && scale && bias && accb && dx && dscale && dbias, "null pointer");
  SNAP_REQUIRE(Nimg >= 1 && H >= 1 && W >= 1, "empty problem");
  SNAP_REQUIRE(C == 64 || C == 128 || C == 256 || C == 512 || C == 1024 || C == 2048, "C must be a power of two in [64, 2048]");
  SNAP_REQUIRE(replica_stride >= Nimg * 64, "replica_stride must cover [Nimg][32][2] doubles");
  SNAP_REQUIRE(out_layout >= 0 && out_layout <= 2, "out_layout: 0 dense, 1 zero-bordered, 2 bottom/right extended");
  SNAP_REQUIRE((!dy_phase && dy_sub == nullptr) || (H % 2 == 0 && W % 2 == 0), "phase / subsampled cotangents need even H, W");
  cudaStream_t s = (cudaStream_t)stream;
  if (int rc = check_cuda(cudaMemsetAsync(accb, 0, (size_t)Nimg * C * 2 * sizeof(double), s), "cudaMemsetAsync(accb)"))
    return rc;
  const int HW = H * W;
  const int PL = 256 / (C / 8);
  int ppl = 32;
  while (ppl > 4 && (long long)Nimg * ((HW + PL * ppl - 1) / (PL * ppl)) < 4LL * num_sms()) ppl >>= 1;
  int ppb = ppl * PL;
  if (ppb > HW) ppb = HW;
  dim3 grid((HW + ppb - 1) / ppb, Nimg);
  gn_bwd_reduce_kernel<<<grid, 256, 0, s>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)dy,
                                            (const __nv_bfloat16*)dy_sub, dy_phase, Nimg, H, W, C, acc, replica_stride,
                                            scale, bias, pre_relu, post_relu, ppb, accb);
  if (int rc = check_launch("gn_bwd_reduce_kernel")) return rc;
#define SNAP_GN_BWD_APPLY(OUT_)                                                                                    \
  gn_bwd_apply_kernel<OUT_><<<grid, 256, 0, s>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)dy,                \
                                                 (const __nv_bfloat16*)dy_sub, dy_phase, Nimg,                       \
                                                 (const __nv_bfloat16*)add, H, W, C, acc, replica_stride, scale, bias, \
                                                 pre_relu, post_relu, accb, ppb, (__nv_bfloat16*)dx)
  if (out_layout == 1) SNAP_GN_BWD_APPLY(1);
  else if (out_layout == 2) SNAP_GN_BWD_APPLY(2);
  else SNAP_GN_BWD_APPLY(0);
#undef SNAP_GN_BWD_APPLY
  if (int rc = check_launch("gn_bwd_apply_kernel")) return rc;
  gn_bwd_params_kernel<<<(C + 127) / 128, 128, 0, s>>>(accb, Nimg, C, dscale, dbias);
  return check_launch("gn_bwd_params_kernel");
}

int snapb200_upsample2x_backward(const void* dy, int Nimg, int h, int w, int C, void* dx, void* stream) {
  SNAP_REQUIRE(dy && dx && Nimg >= 1 && h >= 1 && w >= 1 && C % 8 == 0, "bad arguments");
  const long long total = (long long)Nimg * h * w * (C / 8);
  upsample2x_bwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)dy, Nimg, h, w, C, (__nv_bfloat16*)dx);
  return check_launch("upsample2x_bwd_kernel");
}

int snapb200_maxpool3x3s2_backward(const void* x, const void* dy, int Nimg, int H, int W, int C, void* dx, void* stream) {
  SNAP_REQUIRE(x && dy && dx && Nimg >= 1 && H >= 1 && W >= 1 && C % 8 == 0, "bad arguments");
  const long long total = (long long)Nimg * H * W * (C / 8);
  maxpool3x3s2_bwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)x, (const __nv_bfloat16*)dy, Nimg, H, W, C, (__nv_bfloat16*)dx);
  return check_launch("maxpool3x3s2_bwd_kernel");
}

int snapb200_wt_segments(const void* in, int ld_in, int Cout, int Cin, int taps, void* out, int ld_out, void* stream) {
  SNAP_REQUIRE(in && out && Cout >= 1 && Cin >= 1 && taps >= 1, "bad arguments");
  SNAP_REQUIRE(ld_in >= taps * Cin && ld_out >= taps * Cout, "row pitch too small");
  const int total = Cin * taps * Cout;
  wt_segments_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)in, ld_in, Cout, Cin,
                                                                           taps, (__nv_bfloat16*)out, ld_out);
  return check_launch("wt_segments_kernel");
}

int snapb200_stdconv_backward(const float* w, const float* dws, int K, int Cout, float* dw, void* stream) {
  SNAP_REQUIRE(w && dws && dw && K >= 1 && Cout >= 1, "bad arguments");
  stdconv_bwd_kernel<<<Cout, 128, 0, (cudaStream_t)stream>>>(w, dws, K, Cout, dw);
  return check_launch("stdconv_bwd_kernel");
}

}  // extern "C"
