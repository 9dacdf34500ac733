// XLA FFI (jax.ffi) registration shim over the C ABI of include/snapb200.h: one handler per entry point that the
// DEFAULT product path of BEVMapper.__call__ (snap/models/bev_mapper.py:254-296, called under jax.pmap from
// snap/trainer.py:452-464) and exhaustive_pose_voting (snap/models/pose_exhaustive_voting.py:107-124) launch.
//
// jaxlib (and therefore xla/ffi/api/ffi.h) is not installed in this image.  The file is therefore compile-guarded for the
// real header and TYPE-CHECKED in the CPU test suite against a stand-in header (tests/stubs/xla/ffi/api/ffi.h, which
// enforces the same "implementation signature == Ctx/Arg/Ret/Attr list" contract): tests/test_ffi_shim.py.
// With jaxlib present:
//   g++ -std=c++17 -shared -fPIC -I<jaxlib>/include -I/usr/local/cuda/include -Iinclude bindings/xla_ffi_shim.cc
//       -Lsnap_b200 -lsnapb200 -o libsnapb200_xla.so        (<jaxlib> = os.path.dirname(jaxlib.__file__))
// and on the Python side:
//   jax.ffi.register_ffi_target(name, jax.ffi.pycapsule(lib.<name>), platform="CUDA")
//   jax.ffi.ffi_call(name, out_shapes)(*args, **attrs)
//
// Conventions: every buffer is an xla::ffi::AnyBuffer (device memory owned by XLA); scratch memory is an extra OPERAND or
// RESULT sized at trace time with the library's *_workspace / *_scratch_bytes functions, so the library never allocates;
// the handler only enqueues on the PlatformStream (command-buffer compatible); per-device kernel attributes are set by
// the library on first use of each device, so one process may drive all 8 devices of a box (jax.pmap's host model).
#if __has_include("xla/ffi/api/ffi.h")
#include <cuda_runtime.h>
#include <math.h>

#include "snapb200.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;
using Buf = ffi::AnyBuffer;
using Out = ffi::Result<ffi::AnyBuffer>;

static ffi::Error to_error(int rc) {
  if (rc == SNAPB200_OK) return ffi::Error::Success();
  return ffi::Error(rc == SNAPB200_ERR_INVALID ? ffi::ErrorCode::kInvalidArgument : ffi::ErrorCode::kInternal,
                    snapb200_last_error());
}
template <typename T>
static const T* in(const Buf& b) { return static_cast<const T*>(b.untyped_data()); }
template <typename T>
static T* out(Out& b) { return static_cast<T*>(b->untyped_data()); }
static int dim(const Buf& b, int i) { return (int)b.dimensions()[i]; }

#define SNAP_BIND() ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>()

// =====================================================================================================================
// image encoder (resnet.py:34-216, image_encoder.py:32-144)
// =====================================================================================================================
// StdConv standardisation + GEMM-operand relayout of ALL kernels of an encoder (resnet.py:73-79): descs u8[n*sizeof
// (SnapWeightDesc)] (device table built once per parameter tree), mapA i32[nA,4], mapB i32[nB,4]; the operands it writes
// are addressed by the table, `token` is a dummy result that orders the call before its consumers.
static ffi::Error StdWeightsImpl(cudaStream_t s, Buf descs, Buf mapA, Buf mapB, Out token) {
  (void)token;
  return to_error(snapb200_std_weights_batched(descs.untyped_data(), mapA.untyped_data(), dim(mapA, 0), mapB.untyped_data(),
                                               dim(mapB, 0), s));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(snapb200_xla_std_weights_batched, StdWeightsImpl,
                              SNAP_BIND().Arg<Buf>().Arg<Buf>().Arg<Buf>().Ret<Buf>());

// root block, step 1: images f32[N,H,W,3] -> packed bf16 image (resnet.py:199, image_encoder.py:32-39)
static ffi::Error RootPackImageImpl(cudaStream_t s, Buf images, Out packed, int32_t Hp, int32_t Wp, int32_t pad, int32_t cp,
                                    int32_t Hq, int32_t Wq) {
  return to_error(snapb200_root_pack_image(in<float>(images), dim(images, 0), dim(images, 1), dim(images, 2), Hp, Wp, pad, cp,
                                           Hq, Wq, packed->untyped_data(), s));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(snapb200_xla_root_pack_image, RootPackImageImpl,
                              SNAP_BIND().Arg<Buf>().Ret<Buf>().Attr<int32_t>("Hp").Attr<int32_t>("Wp").Attr<int32_t>("pad")
                                  .Attr<int32_t>("cp").Attr<int32_t>("Hq").Attr<int32_t>("Wq"));

// root block, step 2: standardised kernel bf16[Cout, ldb] -> per-kernel-row layout bf16[Cout, KH*32]
static ffi::Error RootPackWeightsImpl(cudaStream_t s, Buf b_std, Out packed, int32_t KH, int32_t KW, int32_t cp) {
  return to_error(snapb200_root_pack_weights(b_std.untyped_data(), dim(b_std, 0), dim(b_std, 1), KH, KW, cp,
                                             packed->untyped_data(), s));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(snapb200_xla_root_pack_weights, RootPackWeightsImpl,
                              SNAP_BIND().Arg<Buf>().Ret<Buf>().Attr<int32_t>("KH").Attr<int32_t>("KW").Attr<int32_t>("cp"));

// root block, step 3: implicit 7x7/2 (or 3x3/1) conv; gn_acc f64[8, N, 32, 2] is updated in place (input/output alias)
static ffi::Error RootConvImpl(cudaStream_t s, Buf packed, Buf b, Buf gn_acc_in, Out y, Out gn_acc, int32_t n_img, int32_t Hq,
                               int32_t Wq, int32_t cp, int32_t KH, int32_t stride, int32_t Ho, int32_t Wo, int32_t with_stats) {
  (void)gn_acc_in;  // aliased to gn_acc by the caller (jax.ffi.ffi_call(..., input_output_aliases={2: 1}))
  SnapRootConvParams p{};
  p.packed = packed.untyped_data();
  p.n_img = n_img; p.Hq = Hq; p.Wq = Wq; p.cp = cp; p.KH = KH; p.stride = stride; p.Ho = Ho; p.Wo = Wo;
  p.b = b.untyped_data();
  p.n = dim(b, 0);
  p.out = y->untyped_data();
  p.ldo = y->dimensions()[1];
  p.gn_acc = with_stats ? out<double>(gn_acc) : nullptr;
  p.gn_replica_stride = n_img * 64;
  return to_error(snapb200_root_conv_bf16(&p, s));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(snapb200_xla_root_conv, RootConvImpl,
                              SNAP_BIND().Arg<Buf>().Arg<Buf>().Arg<Buf>().Ret<Buf>().Ret<Buf>().Attr<int32_t>("n_img")
                                  .Attr<int32_t>("Hq").Attr<int32_t>("Wq").Attr<int32_t>("cp").Attr<int32_t>("KH")
                                  .Attr<int32_t>("stride").Attr<int32_t>("Ho").Attr<int32_t>("Wo").Attr<int32_t>("with_stats"));

static ffi::Error MaxPoolImpl(cudaStream_t s, Buf x, Out y, int32_t n_img, int32_t H, int32_t W) {
  return to_error(snapb200_maxpool3x3s2(x.untyped_data(), n_img, H, W, dim(x, 1), y->untyped_data(), s));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(snapb200_xla_maxpool3x3s2, MaxPoolImpl,
                              SNAP_BIND().Arg<Buf>().Ret<Buf>().Attr<int32_t>("n_img").Attr<int32_t>("H").Attr<int32_t>("W"));

// GroupNorm statistics of a tensor no conv produced (resnet.py:34-41): acc f64[8, N, 32, 2] accumulated in place
static ffi::Error GnStatsImpl(cudaStream_t s, Buf x, Buf acc_in, Out acc, int32_t n_img, int32_t HW, int32_t pre_relu) {
  (void)acc_in;
  return to_error(snapb200_gn_stats(x.untyped_data(), n_img, HW, dim(x, 1), pre_relu, out<double>(acc), s));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(snapb200_xla_gn_stats, GnStatsImpl,
                              SNAP_BIND().Arg<Buf>().Arg<Buf>().Ret<Buf>().Attr<int32_t>("n_img").Attr<int32_t>("HW")
                                  .Attr<int32_t>("pre_relu"));

// GroupNorm + affine (+ ReLU) with the reference's bf16 rounding chain (resnet.py:46-70), written dense (layout 0),
// zero-bordered for a following 3x3 conv (1) or phase-split for a stride-2 3x3 conv (2); out_sub: optional subsampled copy
static ffi::Error GnApplyImpl(cudaStream_t s, Buf x, Buf acc, Buf scale, Buf bias, Out y, Out y_sub, int32_t n_img, int32_t H,
                              int32_t W, int32_t pre_relu, int32_t post_relu, int32_t layout, int32_t with_sub) {
  return to_error(snapb200_gn_apply(x.untyped_data(), n_img, H, W, dim(x, 1), in<double>(acc), n_img * 64, in<float>(scale),
                                    in<float>(bias), pre_relu, post_relu, layout, y->untyped_data(),
                                    with_sub ? y_sub->untyped_data() : nullptr, s));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(snapb200_xla_gn_apply, GnApplyImpl,
                              SNAP_BIND().Arg<Buf>().Arg<Buf>().Arg<Buf>().Arg<Buf>().Ret<Buf>().Ret<Buf>().Attr<int32_t>("n_img")
                                  .Attr<int32_t>("H").Attr<int32_t>("W").Attr<int32_t>("pre_relu").Attr<int32_t>("post_relu")
                                  .Attr<int32_t>("layout").Attr<int32_t>("with_sub"));

// conv as segmented tcgen05 GEMM (resnet.py:73-79,103-134; image_encoder.py:69-77): a bf16[rows, K] (dense, zero-bordered
// or phase-split activation), b bf16[N, num_seg*K], optional residual bf16[M, N]; seg_off i64[num_seg] row offsets of the
// 3x3 taps; remap = (R, C, r0, c0, Ho, Wo) or empty; the GroupNorm statistics of the output are accumulated into
// gn_acc / gn_acc_relu (aliased in/out, f64[8, N_img, 32, 2]) when rows_per_img > 0.
static ffi::Error ConvImpl(cudaStream_t s, Buf a, Buf b, Buf residual, Buf gn_acc_in, Buf gn_acc_relu_in, Out y, Out gn_acc,
                           Out gn_acc_relu, ffi::Span<const int64_t> seg_off, ffi::Span<const int64_t> remap, int64_t m_rows,
                           int32_t seg_k, int32_t with_residual, int64_t rows_per_img, int32_t with_relu_stats, int32_t n_img) {
  (void)gn_acc_in;
  (void)gn_acc_relu_in;
  SnapGemmParams p{};
  p.a = a.untyped_data(); p.a_rows = a.dimensions()[0]; p.a_cols = dim(a, 1); p.a_ld = p.a_cols;
  p.b = b.untyped_data(); p.b_rows = b.dimensions()[0]; p.b_cols = dim(b, 1); p.b_ld = p.b_cols;
  p.m_rows = m_rows; p.n = (int)p.b_rows;
  p.num_seg = seg_off.size() ? (int)seg_off.size() : 1;
  p.seg_k = seg_k;
  for (int i = 0; i < p.num_seg && i < 9; ++i) p.seg_off[i] = seg_off.size() ? (int)seg_off[i] : 0;
  p.out = y->untyped_data(); p.ldo = y->dimensions()[1];
  if (with_residual) { p.residual = residual.untyped_data(); p.ldr = residual.dimensions()[1]; }
  if (remap.size() == 6) {
    p.remap = 1; p.rm_R = (int)remap[0]; p.rm_C = (int)remap[1]; p.rm_r0 = (int)remap[2]; p.rm_c0 = (int)remap[3];
    p.rm_Ho = (int)remap[4]; p.rm_Wo = (int)remap[5];
  }
  if (rows_per_img > 0) {
    p.gn_acc = out<double>(gn_acc);
    p.gn_acc_relu = with_relu_stats ? out<double>(gn_acc_relu) : nullptr;
    p.gn_rows_per_img = rows_per_img;
    p.gn_replica_stride = n_img * 64;
  }
  return to_error(snapb200_gemm_bf16(&p, s));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(snapb200_xla_conv, ConvImpl,
                              SNAP_BIND().Arg<Buf>().Arg<Buf>().Arg<Buf>().Arg<Buf>().Arg<Buf>().Ret<Buf>().Ret<Buf>().Ret<Buf>()
                                  .Attr<ffi::Span<const int64_t>>("seg_off").Attr<ffi::Span<const int64_t>>("remap")
                                  .Attr<int64_t>("m_rows").Attr<int32_t>("seg_k").Attr<int32_t>("with_residual")
                                  .Attr<int64_t>("rows_per_img").Attr<int32_t>("with_relu_stats").Attr<int32_t>("n_img"));

// GroupNorm -> ReLU -> 1x1 conv in one launch (resnet.py:117-133, image_encoder.py:79-85): x bf16[n_img*H*W, C] RAW
// activation, acc f64[8, n_img, 32, 2] its statistics, scale / bias f32[C], b bf16[N, C], optional residual bf16[M, N];
// statistics of the output into gn_acc / gn_acc_relu (aliased in/out) when with_stats.
static ffi::Error ConvGnImpl(cudaStream_t s, Buf x, Buf acc, Buf scale, Buf bias, Buf b, Buf residual, Buf gn_acc_in,
                             Buf gn_acc_relu_in, Out y, Out gn_acc, Out gn_acc_relu, int32_t n_img, int32_t H, int32_t W,
                             int32_t pre_relu, int32_t post_relu, int32_t with_residual, int32_t with_stats,
                             int32_t with_relu_stats) {
  (void)gn_acc_in;
  (void)gn_acc_relu_in;
  SnapConvGnParams p{};
  p.x = x.untyped_data(); p.n_img = n_img; p.H = H; p.W = W; p.C = dim(x, 1);
  p.acc = in<double>(acc); p.replica_stride = n_img * 64;
  p.scale = in<float>(scale); p.bias = in<float>(bias);
  p.pre_relu = pre_relu; p.post_relu = post_relu; p.taps = 1; p.stride = 1;
  p.b = b.untyped_data(); p.b_rows = b.dimensions()[0]; p.b_cols = dim(b, 1); p.b_ld = p.b_cols; p.n = (int)p.b_rows;
  p.out = y->untyped_data(); p.ldo = y->dimensions()[1];
  if (with_residual) { p.residual = residual.untyped_data(); p.ldr = residual.dimensions()[1]; }
  if (with_stats) {
    p.gn_acc = out<double>(gn_acc);
    p.gn_acc_relu = with_relu_stats ? out<double>(gn_acc_relu) : nullptr;
    p.gn_replica_stride = n_img * 64;
  }
  return to_error(snapb200_conv_gn_bf16(&p, s));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(snapb200_xla_conv_gn, ConvGnImpl,
                              SNAP_BIND().Arg<Buf>().Arg<Buf>().Arg<Buf>().Arg<Buf>().Arg<Buf>().Arg<Buf>().Arg<Buf>().Arg<Buf>()
                                  .Ret<Buf>().Ret<Buf>().Ret<Buf>().Attr<int32_t>("n_img").Attr<int32_t>("H").Attr<int32_t>("W")
                                  .Attr<int32_t>("pre_relu").Attr<int32_t>("post_relu").Attr<int32_t>("with_residual")
                                  .Attr<int32_t>("with_stats").Attr<int32_t>("with_relu_stats"));

// 3x3 / stride-1 conv on the zero-bordered layout as a halo GEMM (resnet.py:117-127): a bf16[n_img*(H+2)*(W+2), C],
// b bf16[N, 9*C]; y bf16[n_img*H*W, N]; statistics of the output into gn_acc (aliased in/out).  Shapes the halo kernel
// does not cover (snapb200_conv3x3_halo_supported) go through snapb200_xla_conv with nine segments.
static ffi::Error Conv3x3HaloImpl(cudaStream_t s, Buf a, Buf b, Buf gn_acc_in, Out y, Out gn_acc, int32_t n_img, int32_t H,
                                  int32_t W, int32_t with_stats) {
  (void)gn_acc_in;
  SnapConv3x3Params p{};
  p.a = a.untyped_data(); p.n_img = n_img; p.H = H; p.W = W; p.C = dim(a, 1);
  p.b = b.untyped_data(); p.b_ld = dim(b, 1); p.n = dim(b, 0);
  p.out = y->untyped_data(); p.ldo = y->dimensions()[1];
  if (with_stats) { p.gn_acc = out<double>(gn_acc); p.gn_replica_stride = n_img * 64; }
  if (!snapb200_conv3x3_halo_supported(p.C, p.n, W))
    return ffi::Error(ffi::ErrorCode::kInvalidArgument, "conv3x3_halo: shape not covered, use snapb200_xla_conv");
  return to_error(snapb200_conv3x3_halo_bf16(&p, s));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(snapb200_xla_conv3x3_halo, Conv3x3HaloImpl,
                              SNAP_BIND().Arg<Buf>().Arg<Buf>().Arg<Buf>().Ret<Buf>().Ret<Buf>().Attr<int32_t>("n_img")
                                  .Attr<int32_t>("H").Attr<int32_t>("W").Attr<int32_t>("with_stats"));

// dense / 1x1 conv with bias (+ ReLU): (a bf16[M,K], b bf16[N,K], bias f32[N]) -> out bf16[M,N]  (layers.py:55-78)
static ffi::Error DenseImpl(cudaStream_t s, Buf a, Buf b, Buf bias, Out y, bool relu) {
  SnapGemmParams p{};
  p.a = a.untyped_data(); p.a_rows = a.dimensions()[0]; p.a_cols = dim(a, 1); p.a_ld = p.a_cols;
  p.b = b.untyped_data(); p.b_rows = b.dimensions()[0]; p.b_cols = dim(b, 1); p.b_ld = p.b_cols;
  p.m_rows = p.a_rows; p.n = (int)p.b_rows; p.num_seg = 1; p.seg_k = p.b_cols;
  p.out = y->untyped_data(); p.ldo = p.n;
  p.bias = in<float>(bias);
  p.relu = relu ? 1 : 0;
  return to_error(snapb200_gemm_bf16(&p, s));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(snapb200_xla_dense, DenseImpl,
                              SNAP_BIND().Arg<Buf>().Arg<Buf>().Arg<Buf>().Ret<Buf>().Attr<bool>("relu"));

static ffi::Error Upsample2xImpl(cudaStream_t s, Buf x, Out y, int32_t n_img, int32_t h, int32_t w) {
  return to_error(snapb200_upsample2x(x.untyped_data(), n_img, h, w, dim(x, 1), y->untyped_data(), s));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(snapb200_xla_upsample2x, Upsample2xImpl,
                              SNAP_BIND().Arg<Buf>().Ret<Buf>().Attr<int32_t>("n_img").Attr<int32_t>("h").Attr<int32_t>("w"));

// crop of the finest FPN level to ceil(input / stride) (image_encoder.py:137-141) + the proj MLP's input ReLU (layers.py:73)
static ffi::Error CropReluImpl(cudaStream_t s, Buf x, Out y, int32_t h, int32_t w, int32_t relu) {
  return to_error(snapb200_crop_relu(x.untyped_data(), dim(x, 0), dim(x, 1), dim(x, 2), dim(x, 3), h, w, relu,
                                     y->untyped_data(), s));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(snapb200_xla_crop_relu, CropReluImpl,
                              SNAP_BIND().Arg<Buf>().Ret<Buf>().Attr<int32_t>("h").Attr<int32_t>("w").Attr<int32_t>("relu"));

// =====================================================================================================================
// camera -> BEV lift (streetview_encoder.py:217-287, bev_mapper.py:56-88)
// =====================================================================================================================
// The whole lift of a batch as ONE launch (the default path): fimg bf16[B, V*Hf*Wf, CF], views i32[B, V*23], xs f32[X],
// ys f32[Y], zs f32[B, Z], fusion weights (w1t bf16[256, ldw1], w256 f32[256], b1 f32[256], w2t bf16[128, 256], b2 f32[128])
// -> plane bf16[B, X*Y, 128], valid u8[B, X*Y]; scratch u8[snapb200_lift_fused_batched_scratch_bytes()] and counters
// i32[16] are results so that XLA owns them.
static ffi::Error LiftFusedImpl(cudaStream_t s, Buf fimg, Buf views, Buf xs, Buf ys, Buf zs, Buf w1t, Buf w256, Buf b1, Buf w2t,
                                Buf b2, Out plane, Out valid, Out counters, Out scratch, int32_t V, int32_t Hf, int32_t Wf,
                                float depth_min, float depth_max, int32_t feature_dim, int32_t xy_paired, int32_t X, int32_t Y) {
  const int B = dim(fimg, 0);
  SnapLiftParams p{};
  p.V = V; p.Hf = Hf; p.Wf = Wf; p.CF = dim(fimg, 2);
  p.D = feature_dim; p.S = p.CF - p.D;
  p.X = X; p.Y = Y; p.Z = dim(zs, 1);
  p.depth_min = depth_min; p.depth_max = depth_max;
  p.inv_log_range = 1.0f / logf(depth_max / depth_min);
  p.stats_ld = 288;
  p.xy_paired = xy_paired;
  return to_error(snapb200_lift_fused_batched(
      &p, B, in<SnapLiftView>(views), views.dimensions()[1] * 4 / (long long)sizeof(SnapLiftView), fimg.untyped_data(),
      fimg.dimensions()[1] * fimg.dimensions()[2], in<float>(xs), in<float>(ys), in<float>(zs), zs.dimensions()[1],
      w1t.untyped_data(), w1t.dimensions()[1], in<float>(w256), in<float>(b1), w2t.untyped_data(), in<float>(b2),
      plane->untyped_data(), out<uint8_t>(valid), out<int>(counters), scratch->untyped_data(), scratch->size_bytes(), s));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(snapb200_xla_lift_fused, LiftFusedImpl,
                              SNAP_BIND().Arg<Buf>().Arg<Buf>().Arg<Buf>().Arg<Buf>().Arg<Buf>().Arg<Buf>().Arg<Buf>().Arg<Buf>()
                                  .Arg<Buf>().Arg<Buf>().Ret<Buf>().Ret<Buf>().Ret<Buf>().Ret<Buf>().Attr<int32_t>("V")
                                  .Attr<int32_t>("Hf").Attr<int32_t>("Wf").Attr<float>("depth_min").Attr<float>("depth_max")
                                  .Attr<int32_t>("feature_dim").Attr<int32_t>("xy_paired").Attr<int32_t>("X").Attr<int32_t>("Y"));

// unfused gather + pooling of one scene (for consumers of 'feature_volume'):
// (fimg bf16[V,Hf,Wf,CF], views i32[V*23], xs, ys, zs) -> (stats bf16[N, stats_ld], valid u8[N])
static ffi::Error LiftGatherPoolImpl(cudaStream_t s, Buf fimg, Buf views, Buf xs, Buf ys, Buf zs, Out stats, Out valid,
                                     float depth_min, float depth_max, int32_t feature_dim) {
  SnapLiftParams p{};
  p.V = dim(fimg, 0); p.Hf = dim(fimg, 1); p.Wf = dim(fimg, 2); p.CF = dim(fimg, 3);
  p.D = feature_dim; p.S = p.CF - p.D;
  p.X = dim(xs, 0); p.Y = dim(ys, 0); p.Z = dim(zs, 0);
  p.depth_min = depth_min; p.depth_max = depth_max;
  p.inv_log_range = 1.0f / logf(depth_max / depth_min);
  p.stats_ld = (int)stats->dimensions()[1];
  return to_error(snapb200_lift_gather_pool(&p, in<SnapLiftView>(views), fimg.untyped_data(), in<float>(xs), in<float>(ys),
                                            in<float>(zs), stats->untyped_data(), out<uint8_t>(valid), nullptr, nullptr, s));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(snapb200_xla_lift_gather_pool, LiftGatherPoolImpl,
                              SNAP_BIND().Arg<Buf>().Arg<Buf>().Arg<Buf>().Arg<Buf>().Arg<Buf>().Ret<Buf>().Ret<Buf>()
                                  .Attr<float>("depth_min").Attr<float>("depth_max").Attr<int32_t>("feature_dim"));

// modality fusion 'max' (bev_mapper.py:225-252) and matching head (bev_mapper.py:284-291)
static ffi::Error FuseMaxImpl(cudaStream_t s, Buf a, Buf va, Buf b, Buf vb, Out y, Out vy) {
  const long long cells = (long long)va.element_count();
  return to_error(snapb200_fuse_max(a.untyped_data(), in<uint8_t>(va), b.untyped_data(), in<uint8_t>(vb), cells,
                                    (int)(a.element_count() / (size_t)cells), y->untyped_data(), out<uint8_t>(vy), s));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(snapb200_xla_fuse_max, FuseMaxImpl,
                              SNAP_BIND().Arg<Buf>().Arg<Buf>().Arg<Buf>().Arg<Buf>().Ret<Buf>().Ret<Buf>());

static ffi::Error MatchHeadImpl(cudaStream_t s, Buf plane, Buf valid, Buf kernel, Buf bias, Out y) {
  const long long cells = (long long)valid.element_count();
  return to_error(snapb200_match_head(plane.untyped_data(), in<uint8_t>(valid), cells, (int)(plane.element_count() / (size_t)cells),
                                      in<float>(kernel), in<float>(bias), dim(bias, 0), y->untyped_data(), s));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(snapb200_xla_match_head, MatchHeadImpl,
                              SNAP_BIND().Arg<Buf>().Arg<Buf>().Arg<Buf>().Arg<Buf>().Ret<Buf>());

// =====================================================================================================================
// exhaustive (x, y, theta) voting (pose_exhaustive_voting.py:37-124)
// =====================================================================================================================
// sample_query_templates: rot f32[R/4, 4] is a HOST constant of the grid (cos, sin, tx, ty), passed as an attribute
static ffi::Error RotTemplatesImpl(cudaStream_t s, Buf feats, Buf valid, Buf centers, Out templates, Out t_valid,
                                   ffi::Span<const float> rot, float cell_size, int32_t R) {
  return to_error(snapb200_rot_templates(feats.untyped_data(), in<uint8_t>(valid), nullptr, &rot[0], in<float>(centers),
                                         cell_size, dim(feats, 0), R, dim(feats, 1), dim(feats, 3), templates->untyped_data(),
                                         out<uint8_t>(t_valid), s));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(snapb200_xla_rot_templates, RotTemplatesImpl,
                              SNAP_BIND().Arg<Buf>().Arg<Buf>().Arg<Buf>().Ret<Buf>().Ret<Buf>().Attr<ffi::Span<const float>>("rot")
                                  .Attr<float>("cell_size").Attr<int32_t>("R"));

static ffi::Error XcorrPadMapImpl(cudaStream_t s, Buf m, Out m_pad) {
  return to_error(snapb200_xcorr_pad_map(m.untyped_data(), dim(m, 0), dim(m, 1), dim(m, 3), m_pad->untyped_data(), s));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(snapb200_xla_xcorr_pad_map, XcorrPadMapImpl, SNAP_BIND().Arg<Buf>().Ret<Buf>());

static ffi::Error XcorrCountImpl(cudaStream_t s, Buf t_valid, Buf m_valid, Out cnt, Out den) {
  return to_error(snapb200_xcorr_count(in<uint8_t>(t_valid), in<uint8_t>(m_valid), dim(t_valid, 0), dim(t_valid, 1),
                                       dim(t_valid, 2), out<float>(cnt), out<float>(den), s));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(snapb200_xla_xcorr_count, XcorrCountImpl,
                              SNAP_BIND().Arg<Buf>().Arg<Buf>().Ret<Buf>().Ret<Buf>());

// template_matching, the default kernel (xcorr_rows_kernel): templates bf16[B, G*G, RP, D] (cell-major), m_pad from
// xcorr_pad_map, cnt f32[B,R,U,U], den f32[B,R] -> scores f32[B,R,U,U]; workspace u8[snapb200_xcorr_scores_rows_workspace()]
static ffi::Error XcorrScoresRowsImpl(cudaStream_t s, Buf templates, Buf m_pad, Buf cnt, Buf den, Out scores, Out workspace,
                                      int32_t G, float thr) {
  return to_error(snapb200_xcorr_scores_rows(templates.untyped_data(), m_pad.untyped_data(), in<float>(cnt), in<float>(den),
                                             dim(cnt, 0), dim(cnt, 1), G, dim(templates, 3), thr, out<float>(scores),
                                             workspace->untyped_data(), workspace->size_bytes(), s));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(snapb200_xla_xcorr_scores_rows, XcorrScoresRowsImpl,
                              SNAP_BIND().Arg<Buf>().Arg<Buf>().Arg<Buf>().Arg<Buf>().Ret<Buf>().Ret<Buf>().Attr<int32_t>("G")
                                  .Attr<float>("thr"));

// the segmented-GEMM fallback for shapes the rows kernel does not cover (R > 64, G > 129)
static ffi::Error XcorrScoresImpl(cudaStream_t s, Buf templates, Buf m_pad, Buf cnt, Buf den, Out scores, int32_t G, float thr) {
  return to_error(snapb200_xcorr_scores(templates.untyped_data(), m_pad.untyped_data(), in<float>(cnt), in<float>(den),
                                        dim(cnt, 0), dim(cnt, 1), G, dim(templates, 3), thr, out<float>(scores), s));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(snapb200_xla_xcorr_scores, XcorrScoresImpl,
                              SNAP_BIND().Arg<Buf>().Arg<Buf>().Arg<Buf>().Arg<Buf>().Ret<Buf>().Attr<int32_t>("G").Attr<float>("thr"));

// =====================================================================================================================
// sampling localizer: the handler pattern for an op with a caller-owned workspace OPERAND (pose_estimation.py:65-85,206-209)
// =====================================================================================================================
static ffi::Error LocPoseScoringImpl(cudaStream_t s, Buf sim, Buf point_scale, Buf i_xy, Buf valid_j, Buf poses, Buf workspace,
                                     Out scores, float cell_size, int32_t mask_out_of_bounds) {
  SnapLocScoreParams p{};
  p.B = dim(sim, 0); p.N = dim(sim, 1); p.H = dim(sim, 2); p.W = dim(sim, 3);
  p.P = dim(poses, 1);
  p.cell_size = cell_size;
  p.mask_out_of_bounds = mask_out_of_bounds;
  p.i_xy_batched = i_xy.dimensions().size() == 3;
  return to_error(snapb200_loc_pose_scoring(&p, sim.untyped_data(), in<float>(point_scale), in<float>(i_xy), in<uint8_t>(valid_j),
                                            in<float>(poses), workspace.untyped_data(), workspace.size_bytes(),
                                            out<float>(scores), s));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(snapb200_xla_loc_pose_scoring, LocPoseScoringImpl,
                              SNAP_BIND().Arg<Buf>().Arg<Buf>().Arg<Buf>().Arg<Buf>().Arg<Buf>().Arg<Buf>().Ret<Buf>()
                                  .Attr<float>("cell_size").Attr<int32_t>("mask_out_of_bounds"));
#else
// xla/ffi/api/ffi.h not available: nothing to compile (see the header comment).
#endif
