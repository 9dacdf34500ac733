// XLA FFI (jax.ffi) registration shim over the C ABI of include/snapb200.h.
//
// SOURCE-ONLY in this image: jaxlib (and therefore xla/ffi/api/ffi.h) is not installed, so this file is
// compile-guarded and untested here.  With jaxlib present:
//   g++ -std=c++17 -shared -fPIC -I$(python -c "import jaxlib,os;print(os.path.join(os.path.dirname(jaxlib.__file__),'include'))") \
//       -Iinclude bindings/xla_ffi_shim.cc -Lsnap_b200 -lsnapb200 -o libsnapb200_xla.so
// and on the Python side:
//   jax.ffi.register_ffi_target(name, jax.ffi.pycapsule(lib.<name>), platform="CUDA")
//   jax.ffi.ffi_call(name, out_shapes)(*args, **attrs)
#if __has_include("xla/ffi/api/ffi.h")
#include <cuda_runtime.h>

#include "snapb200.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

static ffi::Error to_error(int rc) {
  if (rc == SNAPB200_OK) return ffi::Error::Success();
  return ffi::Error(rc == SNAPB200_ERR_INVALID ? ffi::ErrorCode::kInvalidArgument : ffi::ErrorCode::kInternal,
                    snapb200_last_error());
}

// ---- lift_gather_pool: (fimg bf16[V,Hf,Wf,CF], views i32[V*23], xs f32[X], ys f32[Y], zs f32[Z])
//                        -> (stats bf16[N, stats_ld], valid u8[N])
static ffi::Error LiftGatherPoolImpl(cudaStream_t stream, ffi::AnyBuffer fimg, ffi::AnyBuffer views,
                                     ffi::AnyBuffer xs, ffi::AnyBuffer ys, ffi::AnyBuffer zs,
                                     ffi::Result<ffi::AnyBuffer> stats, ffi::Result<ffi::AnyBuffer> valid,
                                     float depth_min, float depth_max, int32_t feature_dim) {
  auto d = fimg.dimensions();
  SnapLiftParams p{};
  p.V = (int)d[0]; p.Hf = (int)d[1]; p.Wf = (int)d[2]; p.CF = (int)d[3];
  p.D = feature_dim; p.S = p.CF - p.D;
  p.X = (int)xs.dimensions()[0]; p.Y = (int)ys.dimensions()[0]; p.Z = (int)zs.dimensions()[0];
  p.depth_min = depth_min; p.depth_max = depth_max;
  p.inv_log_range = 1.0f / logf(depth_max / depth_min);
  p.stats_ld = (int)stats->dimensions()[1];
  return to_error(snapb200_lift_gather_pool(
      &p, static_cast<const SnapLiftView*>(views.untyped_data()), fimg.untyped_data(),
      static_cast<const float*>(xs.untyped_data()), static_cast<const float*>(ys.untyped_data()),
      static_cast<const float*>(zs.untyped_data()), stats->untyped_data(),
      static_cast<uint8_t*>(valid->untyped_data()), nullptr, nullptr, stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(snapb200_xla_lift_gather_pool, LiftGatherPoolImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::AnyBuffer>().Arg<ffi::AnyBuffer>().Arg<ffi::AnyBuffer>()
                                  .Arg<ffi::AnyBuffer>().Arg<ffi::AnyBuffer>()
                                  .Ret<ffi::AnyBuffer>().Ret<ffi::AnyBuffer>()
                                  .Attr<float>("depth_min").Attr<float>("depth_max")
                                  .Attr<int32_t>("feature_dim"));

// ---- dense / 1x1 conv: (a bf16[M,K], b bf16[N,K], bias f32[N]) -> out bf16[M,N]
static ffi::Error DenseImpl(cudaStream_t stream, ffi::AnyBuffer a, ffi::AnyBuffer b, ffi::AnyBuffer bias,
                            ffi::Result<ffi::AnyBuffer> out, bool relu) {
  SnapGemmParams p{};
  p.a = a.untyped_data(); p.a_rows = a.dimensions()[0]; p.a_cols = (int)a.dimensions()[1]; p.a_ld = p.a_cols;
  p.b = b.untyped_data(); p.b_rows = b.dimensions()[0]; p.b_cols = (int)b.dimensions()[1]; p.b_ld = p.b_cols;
  p.m_rows = p.a_rows; p.n = (int)p.b_rows; p.num_seg = 1; p.seg_k = p.b_cols;
  p.out = out->untyped_data(); p.ldo = p.n;
  p.bias = static_cast<const float*>(bias.untyped_data());
  p.relu = relu ? 1 : 0;
  return to_error(snapb200_gemm_bf16(&p, stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(snapb200_xla_dense, DenseImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::AnyBuffer>().Arg<ffi::AnyBuffer>().Arg<ffi::AnyBuffer>()
                                  .Ret<ffi::AnyBuffer>().Attr<bool>("relu"));

// ---- template matching: (templates bf16[B,R,G,G,D], m_pad bf16[B,3G-2,P,D], cnt f32[B,R,U,U], den f32[B,R])
//                         -> scores f32[B,R,U,U]
static ffi::Error XcorrScoresImpl(cudaStream_t stream, ffi::AnyBuffer templates, ffi::AnyBuffer m_pad,
                                  ffi::AnyBuffer cnt, ffi::AnyBuffer den, ffi::Result<ffi::AnyBuffer> scores,
                                  float thr) {
  auto d = templates.dimensions();
  return to_error(snapb200_xcorr_scores(templates.untyped_data(), m_pad.untyped_data(),
                                        static_cast<const float*>(cnt.untyped_data()),
                                        static_cast<const float*>(den.untyped_data()), (int)d[0], (int)d[1],
                                        (int)d[2], (int)d[4], thr, static_cast<float*>(scores->untyped_data()),
                                        stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(snapb200_xla_xcorr_scores, XcorrScoresImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::AnyBuffer>().Arg<ffi::AnyBuffer>().Arg<ffi::AnyBuffer>()
                                  .Arg<ffi::AnyBuffer>().Ret<ffi::AnyBuffer>().Attr<float>("thr"));
// ---- loc_pose_scoring (pose_estimation.py:65-85,206-209): (sim bf16[B,N,H,W], point_scale f32[B,N], i_xy f32[N,2],
//      valid_j u8[B,H,W], poses f32[B,P,3], workspace u8[ws]) -> scores f32[B,P].  The workspace is an extra operand so
//      that XLA owns the memory (size = snapb200_loc_pose_scoring_workspace(), computed at trace time on the host).
static ffi::Error LocPoseScoringImpl(cudaStream_t stream, ffi::AnyBuffer sim, ffi::AnyBuffer point_scale,
                                     ffi::AnyBuffer i_xy, ffi::AnyBuffer valid_j, ffi::AnyBuffer poses,
                                     ffi::AnyBuffer workspace, ffi::Result<ffi::AnyBuffer> scores, float cell_size,
                                     int32_t mask_out_of_bounds) {
  auto d = sim.dimensions();
  SnapLocScoreParams p{};
  p.B = (int)d[0]; p.N = (int)d[1]; p.H = (int)d[2]; p.W = (int)d[3];
  p.P = (int)poses.dimensions()[1];
  p.cell_size = cell_size;
  p.mask_out_of_bounds = mask_out_of_bounds;
  p.i_xy_batched = i_xy.dimensions().size() == 3;
  return to_error(snapb200_loc_pose_scoring(
      &p, sim.untyped_data(), static_cast<const float*>(point_scale.untyped_data()),
      static_cast<const float*>(i_xy.untyped_data()), static_cast<const uint8_t*>(valid_j.untyped_data()),
      static_cast<const float*>(poses.untyped_data()), workspace.untyped_data(), workspace.size_bytes(),
      static_cast<float*>(scores->untyped_data()), stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(snapb200_xla_loc_pose_scoring, LocPoseScoringImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::AnyBuffer>().Arg<ffi::AnyBuffer>().Arg<ffi::AnyBuffer>()
                                  .Arg<ffi::AnyBuffer>().Arg<ffi::AnyBuffer>().Arg<ffi::AnyBuffer>()
                                  .Ret<ffi::AnyBuffer>()
                                  .Attr<float>("cell_size").Attr<int32_t>("mask_out_of_bounds"));


// The remaining entry points (gn_stats/gn_apply, root_im2col, maxpool, upsample2x, crop_relu, vertical_max,
// match_head, fuse_max, rot_templates, xcorr_pad_map, xcorr_count, std_weights_batched) bind the same way:
// AnyBuffer pointers + dims -> the C ABI call, PlatformStream -> `stream`.
#else
// xla/ffi/api/ffi.h not available: nothing to compile (see the header comment).

#endif
