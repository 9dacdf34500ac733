// Device code shared by the unfused (lift_kernels.cu) and fused (lift_fused.cu) camera->BEV lift:
// parameter structs, the bit-exact projection prologue and the bilinear tap set-up.
// Reference: snap/models/streetview_encoder.py:42-65, snap/utils/geometry.py:52-56,193-221,260-280,
// snap/utils/grids.py:116-137.
#pragma once
#include <cuda_bf16.h>
#include <math.h>

#include "common.cuh"

namespace snapb200 {

constexpr int LIFT_MAX_VIEWS = 8;

struct LiftView {
  float Rinv[9];  // row-major inverse rotation (scene -> view)
  float tinv[3];
  float f[2], c[2], wh[2];  // camera scaled to feature resolution, (x, y) order
  float k_radial[3];
  float tan_half_fov;
  int fisheye;
};

struct LiftParams {
  int V, Hf, Wf, CF, D, S;  // CF = D + S channels per texel
  int X, Y, Z;
  float depth_min, depth_max, inv_log_range;  // 1 / log(max/min)
  int stats_ld;                               // row pitch of the stats matrix (>= 2*D + 1, mult of 32)
  int xy_paired;                              // 0: xs[X], ys[Y] (separable grid); 1: xs[X*Y], ys[X*Y] per column
  int no_variance, add_minmax;                // statistics blocks of the unfused gather/pool kernel (0, 0 = default)
};

struct Proj {
  float row, col, depth;
  bool vis;
};

__device__ __forceinline__ Proj project_point(const LiftView& v, float px, float py, float pz) {
  float pv[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    float s = __fmul_rn(v.Rinv[3 * i + 0], px);
    s = __fadd_rn(s, __fmul_rn(v.Rinv[3 * i + 1], py));
    s = __fadd_rn(s, __fmul_rn(v.Rinv[3 * i + 2], pz));
    pv[i] = __fadd_rn(v.tinv[i], s);
  }
  const float eps = 1e-3f;
  bool vis = pv[2] >= eps;
  const float zc = fmaxf(pv[2], eps);
  float x = __fdiv_rn(pv[0], zc);
  float y = __fdiv_rn(pv[1], zc);
  if (v.fisheye) {  // snap/utils/geometry.py:260-272 (float tolerance only: atan differs in ulps)
    const float eps2 = __fmul_rn(eps, eps);
    const float r2 = __fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y));
    const bool in_center = r2 < eps2;
    const float radius = sqrtf(in_center ? eps2 : r2);
    const float theta = atanf(radius);
    const float t2 = theta * theta;
    const float offset = v.k_radial[0] * t2 + v.k_radial[1] * t2 * t2 + v.k_radial[2] * t2 * t2 * t2;
    float dist = (offset + 1.0f) * theta / radius;
    if (in_center) dist = 1.0f;
    x *= dist;
    y *= dist;
    vis = vis && (in_center || (radius < v.tan_half_fov && dist > 0.f));
  }
  const float u = __fadd_rn(__fmul_rn(x, v.f[0]), v.c[0]);
  const float w = __fadd_rn(__fmul_rn(y, v.f[1]), v.c[1]);
  vis = vis && u >= 0.f && u < v.wh[0] && w >= 0.f && w < v.wh[1];
  Proj p;
  p.row = w;  // flipped to (row, col), streetview_encoder.py:59
  p.col = u;
  p.depth = pv[2];
  p.vis = vis;
  return p;
}

struct Taps {
  int r0, r1, c0, c1;  // clamped tap indices
  int rlo, clo;        // unclamped floor (for the parity debug output)
  float wr1, wc1;      // weights of the upper taps
};

__device__ __forceinline__ Taps make_taps(float row, float col, int Hf, int Wf) {
  Taps t;
  const float pr = __fadd_rn(row, -0.5f), pc = __fadd_rn(col, -0.5f);
  const float fr = floorf(pr), fc = floorf(pc);
  t.rlo = (int)fr;
  t.clo = (int)fc;
  t.wr1 = __fadd_rn(pr, -fr);
  t.wc1 = __fadd_rn(pc, -fc);
  t.r0 = min(max(t.rlo, 0), Hf - 1);
  t.r1 = min(max(t.rlo + 1, 0), Hf - 1);
  t.c0 = min(max(t.clo, 0), Wf - 1);
  t.c1 = min(max(t.clo + 1, 0), Wf - 1);
  return t;
}

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  const uint32_t uu[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 t = unpack_bf16(uu[j]);
    f[2 * j] = t.x;
    f[2 * j + 1] = t.y;
  }
}

}  // namespace snapb200
