"""The C-ABI library loads on a CPU-only box and exports every symbol include/snapb200.h declares;
argument validation fails loudly before any CUDA call.  (No compute without a GPU.)"""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "snapb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(snapb200_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from snap_b200 import _lib
    lib = _lib.lib()
    declared = _header_symbols()
    assert len(declared) >= 20
    for sym in declared:
        assert hasattr(lib, sym), f"{sym} declared in include/snapb200.h but not exported"
    assert sorted(_lib.EXPORTED_SYMBOLS) == declared, "python binding list and header disagree"


def test_struct_layouts_match_header():
    from snap_b200 import _lib
    # sizes asserted in csrc/lift_kernels.cu (static_assert) and mirrored here
    assert C.sizeof(_lib.LiftView) == 92
    assert C.sizeof(_lib.LiftParams) == 64
    assert C.sizeof(_lib.WeightDesc) == 48
    assert C.sizeof(_lib.GemmParams) % 8 == 0


def test_invalid_arguments_fail_loudly_without_gpu():
    from snap_b200 import _lib
    lib = _lib.lib()
    assert lib.snapb200_version() == 100
    p = _lib.GemmParams()
    rc = lib.snapb200_gemm_bf16(C.byref(p), None)
    assert rc == -1 and b"null operand" in lib.snapb200_last_error()
    with pytest.raises(_lib.SnapB200Error):
        _lib.check(rc)
    assert lib.snapb200_xcorr_padded_cols(128) == 384 and lib.snapb200_xcorr_padded_cols(64) == 192
    rc = lib.snapb200_xcorr_count(None, None, 1, 36, 128, None, None, None)
    assert rc == -1


def test_ops_refuse_cpu_tensors():
    import torch
    from snap_b200 import _lib, ops
    a = torch.zeros((128, 64), dtype=torch.bfloat16)
    with pytest.raises(_lib.SnapB200Error):
        ops.gemm(a, a, torch.zeros((128, 128)))


def test_plain_c_client(tmp_path):
    """The drop-in boundary is a C ABI: a C99 program with nothing but include/snapb200.h and dlopen binds and calls it
    (no Python, no torch); the library itself links only the C/C++ runtime (the CUDA runtime is linked statically)."""
    import shutil
    import subprocess
    from snap_b200 import _lib
    _lib.lib()
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    exe = tmp_path / "abi_client"
    src = os.path.join(ROOT, "tests", "c", "abi_client.c")
    subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), src, "-ldl", "-o", str(exe)], check=True)
    res = subprocess.run([str(exe), str(_lib._LIB_PATH)], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "null operand" in res.stdout and "multiple of 8" in res.stdout and res.stdout.strip().endswith("ok")
    deps = subprocess.run(["ldd", str(_lib._LIB_PATH)], capture_output=True, text=True).stdout
    assert "torch" not in deps and "python" not in deps


def test_host_side_planning_functions_without_gpu():
    """Workspace / plan queries are host-only (148 SMs assumed without a device): caller-owned workspaces can be sized at
    trace time, and shapes the kernels cannot take are refused with a message."""
    from snap_b200 import _lib
    lib = _lib.lib()
    ws = lib.snapb200_loc_pose_scoring_workspace
    ws.restype = C.c_size_t
    p = _lib.LocScoreParams()
    p.B, p.N, p.H, p.W, p.P, p.cell_size = 1, 4652, 128, 128, 10001, 0.2
    # 16 poses per thread x 256 threads -> 3 pose chunks; 2 blocks per SM x 2 waves over 148 SMs -> 198 point splits
    assert ws(C.byref(p)) == 198 * 10001 * 4
    p.P = 68921          # 41^3 refinement lattice: 17 chunks -> 35 splits
    assert ws(C.byref(p)) == 35 * 68921 * 4
    p.H = p.W = 256      # one map no longer fits twice: single-buffer path, still supported
    assert ws(C.byref(p)) > 0
    p.H = p.W = 512
    assert ws(C.byref(p)) == 0 and b"does not fit in shared memory" in lib.snapb200_last_error()
    p.H, p.W = 128, 100
    assert ws(C.byref(p)) == 0 and b"multiple of 8" in lib.snapb200_last_error()
    wg = lib.snapb200_dense_wgrad_workspace
    wg.restype = C.c_size_t
    assert wg(C.c_longlong(65536), 128, 128) == 293 * (128 * 128 + 128) * 4      # 224 rows per slab -> 293 slabs of partials


def test_backward_entry_points_reject_null_arguments_without_touching_the_gpu():
    """Every backward entry point added for SURVEY 8(f)1 validates its arguments before any CUDA call: NULL pointers /
    empty problems return SNAPB200_ERR_INVALID with a message (no crash, no device needed)."""
    import ctypes as C
    from snap_b200 import _lib
    lib = _lib.lib()
    lib.snapb200_last_error.restype = C.c_char_p
    null, i0 = C.c_void_p(0), 0
    ll = C.c_longlong
    lp, sp = _lib.LiftParams(), _lib.LocScoreParams()
    calls = {
        "snapb200_gn_backward": (null, null, null, i0, null, 1, 4, 4, 64, null, 64, null, null, i0, 1, i0, null, null, null, null, null),
        "snapb200_upsample2x_backward": (null, 1, 4, 4, 128, null, null),
        "snapb200_maxpool3x3s2_backward": (null, null, 1, 4, 4, 64, null, null),
        "snapb200_wt_segments": (null, 64, 64, 64, 1, null, 64, null),
        "snapb200_stdconv_backward": (null, null, 64, 64, null, null),
        "snapb200_lift_gather_pool_backward": (C.byref(lp), null, null, null, null, null, null, null, null),
        "snapb200_lift_select_pool_backward": (C.byref(lp), 4, null, null, null, null, null, null, null, null, null),
        "snapb200_vertical_max_backward": (null, null, null, ll(16), 4, 128, null, null),
        "snapb200_match_head_backward": (null, null, ll(16), 128, null, null, null, null, null),
        "snapb200_fuse_max_backward": (null, null, null, null, null, ll(16), 128, null, null, null),
        "snapb200_loc_nll_backward": (null, null, null, 1, 8, i0, C.c_float(0), C.c_float(0), null, null, null),
        "snapb200_loc_pose_scoring_backward": (C.byref(sp), null, null, null, null, null, null, 1, 8, null, null),
    }
    for name, args in calls.items():
        rc = getattr(lib, name)(*args)
        msg = lib.snapb200_last_error().decode()
        assert rc != 0 and msg, (name, rc, msg)


def test_round2_encoder_entry_points_plan_and_validate_without_a_gpu():
    """`snapb200_conv3x3_halo_supported` is a host-only plan query (stage 1 / 2 shapes fit, stage 3 does not: its weight tiles
    alone exceed shared memory); the halo conv, the fused GroupNorm conv and the descriptor self-test reject NULL / invalid
    arguments with a message before any CUDA call."""
    from snap_b200 import _lib
    lib = _lib.lib()
    lib.snapb200_last_error.restype = C.c_char_p
    sup = lib.snapb200_conv3x3_halo_supported
    assert sup(64, 64, 168) == 1 and sup(128, 128, 84) == 1          # conv2 of stages 1 and 2 at 480 x 640 inputs
    assert sup(256, 256, 42) == 0 and sup(512, 512, 21) == 0         # stages 3 and 4: 9-segment GEMM
    assert sup(64, 64, 100000) == 0 and sup(96, 64, 40) == 0 and sup(64, 32, 40) == 0
    p = _lib.Conv3x3Params()
    assert lib.snapb200_conv3x3_halo_bf16(C.byref(p), None) == -1 and b"null operand" in lib.snapb200_last_error()
    q = _lib.ConvGnParams()
    assert lib.snapb200_conv_gn_bf16(C.byref(q), None) == -1 and b"null operand" in lib.snapb200_last_error()
    assert lib.snapb200_selftest_shifted_desc(None, None, 0, 0, None, None) == -1
