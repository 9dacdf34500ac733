// Losses and metrics of SemanticNetModel.loss_metrics_function (snap/models/semantic_net.py:56-110, 300-343) on the
// logits of the semantic head: class-balanced soft-max cross-entropy over the area classes and over the exclusive object
// classes (+ void), class-balanced sigmoid cross-entropy over the independent object classes, masked means per example
// (layers.masked_mean, layers.py:30-33: an empty mask divides by the number of cells), accuracy and per-class recall.
//
// One block per example; every thread walks a strided slice of the cells with sequential fp32 sums, then a fixed-order
// warp / block reduction (deterministic).
#include <math.h>

#include "common.cuh"
#include "host_common.h"

namespace snapb200 {

constexpr int SEM_MAXC = 8;       // classes per group
constexpr int SEM_NACC = 7 + 6 * SEM_MAXC;

__device__ __forceinline__ float log_sigmoid_f(float x) {  // jax.nn.log_sigmoid = -softplus(-x)
  return -(fmaxf(-x, 0.f) + log1pf(expf(-fabsf(x))));
}

// soft-max cross-entropy with integer labels (optax): logsumexp(logits) - logits[label]; also the arg-max (first maximum)
__device__ __forceinline__ float softmax_xent(const float* __restrict__ l, int n, int label, int* argmax) {
  float mx = l[0];
  int am = 0;
  for (int c = 1; c < n; ++c)
    if (l[c] > mx) {
      mx = l[c];
      am = c;
    }
  float se = 0.f;
  for (int c = 0; c < n; ++c) se += expf(l[c] - mx);
  *argmax = am;
  return logf(se) - (l[label] - mx);
}

__global__ void __launch_bounds__(256)
sem_loss_kernel(const SnapSemLossParams P, const float* __restrict__ logits, const int* __restrict__ labels_area,
                const uint8_t* __restrict__ valid_area, const int* __restrict__ labels_excl,
                const uint8_t* __restrict__ masks_indep, const uint8_t* __restrict__ valid,
                const float* __restrict__ w_area, const float* __restrict__ w_excl, const float* __restrict__ w_pos,
                const float* __restrict__ w_neg, float* __restrict__ out) {
  __shared__ float red[8][SEM_NACC];
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int Ka = P.num_area, Ke = P.num_excl, Ki = P.num_indep;
  // accumulators: 0 nll_a, 1 nll_e, 2 nll_i, 3 cnt valid_area, 4 cnt valid, 5 correct_a, 6 correct_e,
  // then per class (correct, count) for area / excl / indep
  float acc[SEM_NACC];
#pragma unroll
  for (int i = 0; i < SEM_NACC; ++i) acc[i] = 0.f;
  float* ca = acc + 7;                 // area: correct[8], count[8]
  float* ce = acc + 7 + 2 * SEM_MAXC;  // excl
  float* ci = acc + 7 + 4 * SEM_MAXC;  // indep
  for (int cell = tid; cell < P.cells; cell += 256) {
    const size_t g = (size_t)b * P.cells + cell;
    const float* l = logits + g * P.ld;
    const bool v = valid[g] != 0;
    {  // areas (:300-310): valid = bev valid & any area label
      const bool va = valid_area[g] != 0;
      const int lab = labels_area[g];
      int am;
      float nll = softmax_xent(l, Ka, lab, &am);
      if (w_area != nullptr) nll *= w_area[lab];
      if (va) {
        acc[0] += nll;
        acc[3] += 1.f;
        acc[5] += am == lab ? 1.f : 0.f;
#pragma unroll
        for (int c = 0; c < SEM_MAXC; ++c)
          if (c == lab) {
            ca[c] += am == lab ? 1.f : 0.f;
            ca[SEM_MAXC + c] += 1.f;
          }
      }
    }
    if (v) acc[4] += 1.f;
    if (Ke > 0) {  // exclusive objects + void (:312-321)
      const int lab = labels_excl[g];
      int am;
      float nll = softmax_xent(l + Ka, Ke, lab, &am);
      if (w_excl != nullptr) nll *= w_excl[lab];
      if (v) {
        acc[1] += nll;
        acc[6] += am == lab ? 1.f : 0.f;
#pragma unroll
        for (int c = 0; c < SEM_MAXC; ++c)
          if (c == lab) {
            ce[c] += am == lab ? 1.f : 0.f;
            ce[SEM_MAXC + c] += 1.f;
          }
      }
    }
    if (Ki > 0) {  // independent objects (:322-330): sigmoid cross-entropy, mean over the classes
      float s = 0.f;
#pragma unroll
      for (int c = 0; c < SEM_MAXC; ++c) {
        if (c >= Ki) break;
        const float x = l[Ka + Ke + c];
        const bool gt = masks_indep[g * Ki + c] != 0;
        float nll = gt ? -log_sigmoid_f(x) : -log_sigmoid_f(-x);
        if (w_pos != nullptr) nll *= gt ? w_pos[c] : w_neg[c];
        s += nll;
        const bool pred = 1.f / (1.f + expf(-x)) > 0.5f;  // jax.nn.sigmoid(logits) > 0.5
        if (v && gt) {
          ci[c] += pred == gt ? 1.f : 0.f;
          ci[SEM_MAXC + c] += 1.f;
        }
      }
      if (v) acc[2] += s / (float)Ki;
    }
  }
#pragma unroll
  for (int i = 0; i < SEM_NACC; ++i) {
    float x = acc[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if (lane == 0) red[warp][i] = x;
  }
  __syncthreads();
  if (tid < SEM_NACC) {
    float x = 0.f;
    for (int w = 0; w < 8; ++w) x += red[w][tid];
    red[0][tid] = x;
  }
  __syncthreads();
  if (tid == 0) {
    const float* r = red[0];
    const float cells = (float)P.cells;
    auto mm = [&](float num, float cnt) { return num / (cnt > 0.f ? cnt : cells); };  // layers.masked_mean
    float* o = out + (size_t)b * SNAPB200_SEM_OUT;
    for (int i = 0; i < SNAPB200_SEM_OUT; ++i) o[i] = 0.f;
    const float nll_a = mm(r[0], r[3]);
    o[0] = nll_a;
    o[4] = mm(r[5], r[3]);
    float total = nll_a;
    float avg = 0.f;
    for (int c = 0; c < Ka; ++c) {
      const float rc = mm(r[7 + c], r[7 + SEM_MAXC + c]);
      o[16 + c] = rc;
      avg += rc;
    }
    o[6] = avg / (float)Ka;
    if (Ke > 0 || Ki > 0) {
      const float nll_e = Ke > 0 ? mm(r[1], r[4]) : 0.f, nll_i = Ki > 0 ? mm(r[2], r[4]) : 0.f;
      o[1] = nll_e;
      o[2] = nll_i;
      o[5] = mm(r[6], r[4]);
      total = (total + (nll_e + nll_i) / 2.f) / 2.f;  // :339
      avg = 0.f;
      for (int c = 0; c < Ke; ++c) {
        const float rc = mm(r[7 + 2 * SEM_MAXC + c], r[7 + 3 * SEM_MAXC + c]);
        o[24 + c] = rc;
        avg += rc;
      }
      o[7] = Ke > 0 ? avg / (float)Ke : 0.f;
      avg = 0.f;
      for (int c = 0; c < Ki; ++c) {
        const float rc = mm(r[7 + 4 * SEM_MAXC + c], r[7 + 5 * SEM_MAXC + c]);
        o[32 + c] = rc;
        avg += rc;
      }
      o[8] = Ki > 0 ? avg / (float)Ki : 0.f;
    }
    o[3] = total;
  }
}

// Label preparation of SemanticNetModel (semantic_net.py:254-298) on the device: thread per cell.
//   sel_area [Ka][4], sel_excl [Ke-1][4]: ground-truth mask channels OR-ed into each selected class (-1 = unused; 'line'
//   absorbs 'stopline' / 'otherlanemarking', :263-270); sel_indep [Ki]: one channel each.
//   labels = argmax over the selected masks (first true, 0 if none); the exclusive objects get the void index when none.
struct SemLabelSel {
  int area[SEM_MAXC][4];
  int excl[SEM_MAXC][4];
  int indep[SEM_MAXC];
  int Ka, Ke, Ki, ngt;
};

__global__ void sem_labels_kernel(const SemLabelSel S, const uint8_t* __restrict__ masks, const uint8_t* __restrict__ bev_valid,
                                  long long rows, int* __restrict__ labels_area, uint8_t* __restrict__ valid_area,
                                  int* __restrict__ labels_excl, uint8_t* __restrict__ masks_indep) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= rows) return;
  const uint8_t* m = masks + g * S.ngt;
  auto on = [&](const int* sel) {
    bool any = false;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (sel[j] >= 0) any |= m[sel[j]] != 0;
    return any;
  };
  int la = 0;
  bool va = false;
  for (int c = S.Ka - 1; c >= 0; --c)
    if (on(S.area[c])) {
      la = c;
      va = true;
    }
  labels_area[g] = la;
  valid_area[g] = (va && bev_valid[g] != 0) ? 1 : 0;   // pred['bev_features'].valid & valid (:302)
  if (labels_excl != nullptr) {
    int le = S.Ke;                                     // void = len(classes) (:274-275)
    for (int c = S.Ke - 1; c >= 0; --c)
      if (on(S.excl[c])) le = c;
    labels_excl[g] = le;
  }
  if (masks_indep != nullptr)
    for (int c = 0; c < S.Ki; ++c) masks_indep[g * S.Ki + c] = m[S.indep[c]] != 0 ? 1 : 0;
}

}  // namespace snapb200

using namespace snapb200;

extern "C" int snapb200_sem_labels(const int* sel_area, int num_area, const int* sel_excl, int num_excl_classes,
                                   const int* sel_indep, int num_indep, int num_gt, const uint8_t* masks,
                                   const uint8_t* bev_valid, long long rows, int* labels_area, uint8_t* valid_area,
                                   int* labels_excl, uint8_t* masks_indep, void* stream) {
  SNAP_REQUIRE(sel_area && masks && bev_valid && labels_area && valid_area, "null pointer");
  SNAP_REQUIRE(num_area >= 1 && num_area <= SEM_MAXC && num_excl_classes >= 0 && num_excl_classes <= SEM_MAXC - 1 &&
                   num_indep >= 0 && num_indep <= SEM_MAXC, "at most %d classes per group", SEM_MAXC);
  SNAP_REQUIRE((num_excl_classes == 0 && num_indep == 0) || (sel_excl || num_excl_classes == 0), "selection tables missing");
  SemLabelSel S;
  for (int c = 0; c < SEM_MAXC; ++c) {
    for (int j = 0; j < 4; ++j) {
      S.area[c][j] = c < num_area ? sel_area[c * 4 + j] : -1;
      S.excl[c][j] = c < num_excl_classes ? sel_excl[c * 4 + j] : -1;
      SNAP_REQUIRE(S.area[c][j] < num_gt && S.excl[c][j] < num_gt, "mask channel out of range");
    }
    S.indep[c] = c < num_indep ? sel_indep[c] : 0;
    SNAP_REQUIRE(S.indep[c] >= 0 && S.indep[c] < num_gt, "mask channel out of range");
  }
  S.Ka = num_area; S.Ke = num_excl_classes; S.Ki = num_indep; S.ngt = num_gt;
  sem_labels_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      S, masks, bev_valid, rows, labels_area, valid_area, labels_excl, masks_indep);
  return check_launch("sem_labels_kernel");
}

extern "C" int snapb200_sem_loss(const SnapSemLossParams* p, const float* logits, const int* labels_area,
                                 const uint8_t* valid_area, const int* labels_excl, const uint8_t* masks_indep,
                                 const uint8_t* valid, const float* w_area, const float* w_excl, const float* w_pos,
                                 const float* w_neg, float* out, void* stream) {
  SNAP_REQUIRE(p && logits && labels_area && valid_area && valid && out, "null pointer");
  SNAP_REQUIRE(p->B >= 1 && p->cells >= 1, "empty problem");
  SNAP_REQUIRE(p->num_area >= 1 && p->num_area <= SEM_MAXC && p->num_excl >= 0 && p->num_excl <= SEM_MAXC &&
                   p->num_indep >= 0 && p->num_indep <= SEM_MAXC,
               "at most %d classes per group", SEM_MAXC);
  SNAP_REQUIRE(p->ld >= p->num_area + p->num_excl + p->num_indep, "logits pitch too small");
  SNAP_REQUIRE(p->num_excl == 0 || labels_excl, "exclusive-object labels missing");
  SNAP_REQUIRE(p->num_indep == 0 || masks_indep, "independent-object masks missing");
  SNAP_REQUIRE((w_pos == nullptr) == (w_neg == nullptr), "w_pos and w_neg go together");
  sem_loss_kernel<<<p->B, 256, 0, (cudaStream_t)stream>>>(*p, logits, labels_area, valid_area, labels_excl, masks_indep,
                                                          valid, w_area, w_excl, w_pos, w_neg, out);
  return check_launch("sem_loss_kernel");
}
