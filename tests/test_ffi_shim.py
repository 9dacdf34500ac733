"""The jax.ffi registration shim (bindings/xla_ffi_shim.cc) cannot rot: jaxlib's xla/ffi/api/ffi.h is absent from this
image, so the shim is type-checked against a stand-in header that enforces the same contract (the implementation's
parameter list must equal the Ctx / Arg / Ret / Attr list of its binding), compiled against include/snapb200.h (every C
entry point it calls must exist with that signature) and linked against the built library."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "bindings", "xla_ffi_shim.cc")
INC = ["-I" + os.path.join(ROOT, "tests", "stubs"), "-I" + os.path.join(ROOT, "include"), "-I/usr/local/cuda/include"]

# what BEVMapper.apply (default configuration, street-view + aerial) and exhaustive_pose_voting launch
DEFAULT_PATH_ENTRY_POINTS = {
    "snapb200_std_weights_batched", "snapb200_root_pack_image", "snapb200_root_pack_weights", "snapb200_root_conv_bf16",
    "snapb200_maxpool3x3s2", "snapb200_gn_stats", "snapb200_gn_apply", "snapb200_gemm_bf16", "snapb200_upsample2x",
    "snapb200_crop_relu", "snapb200_lift_fused_batched", "snapb200_fuse_max", "snapb200_match_head", "snapb200_rot_templates",
    "snapb200_xcorr_pad_map", "snapb200_xcorr_count", "snapb200_xcorr_scores_rows",
    "snapb200_conv_gn_bf16", "snapb200_conv3x3_halo_bf16",   # round 2: GroupNorm fused into the 1x1 convs, halo 3x3 conv
}


@pytest.mark.skipif(shutil.which("g++") is None, reason="g++ not available")
def test_shim_type_checks_against_the_ffi_contract_and_the_c_abi(tmp_path):
    r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-Wno-comment", *INC, SHIM], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    lib = os.path.join(ROOT, "snap_b200", "libsnapb200.so")
    if os.path.exists(lib):     # link: every entry point the handlers call is exported by the built library
        so = str(tmp_path / "libsnapb200_xla.so")
        r = subprocess.run(["g++", "-std=c++17", "-shared", "-fPIC", "-Wno-comment", *INC, SHIM, "-L" + os.path.dirname(lib),
                            "-lsnapb200", "-Wl,--no-undefined", "-L/usr/local/cuda/lib64", "-o", so], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-3000:]
        syms = subprocess.run(["nm", "-D", so], capture_output=True, text=True).stdout
        assert len(re.findall(r" T snapb200_xla_\w+", syms)) >= 20


def test_shim_binds_every_entry_point_of_the_default_path():
    src = open(SHIM).read()
    called = set(re.findall(r"\b(snapb200_[a-z0-9_]+)\(", src))
    missing = DEFAULT_PATH_ENTRY_POINTS - called
    assert not missing, f"entry points of the default path without a jax.ffi handler: {sorted(missing)}"
    handlers = re.findall(r"XLA_FFI_DEFINE_HANDLER_SYMBOL\((snapb200_xla_\w+),", src)
    assert len(handlers) == len(set(handlers)) >= 20


def test_a_wrong_binding_is_rejected_by_the_stand_in_header(tmp_path):
    """The stand-in header really checks: a handler whose Arg list does not match its implementation fails to compile."""
    bad = tmp_path / "bad.cc"
    bad.write_text('#include <cuda_runtime.h>\n#include "xla/ffi/api/ffi.h"\nnamespace ffi = xla::ffi;\n'
                   "static ffi::Error Impl(cudaStream_t, ffi::AnyBuffer, ffi::Result<ffi::AnyBuffer>, float) { return ffi::Error::Success(); }\n"
                   "XLA_FFI_DEFINE_HANDLER_SYMBOL(bad, Impl, ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>()"
                   ".Arg<ffi::AnyBuffer>().Arg<ffi::AnyBuffer>().Attr<float>(\"x\"));\n")
    r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", *INC, str(bad)], capture_output=True, text=True)
    assert r.returncode != 0 and "does not match its xla::ffi binding" in r.stderr
