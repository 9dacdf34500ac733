"""Shared helpers for the parity tests (CUDA library vs the CPU oracle)."""
import numpy as np
import torch

F = np.float32


def rd_bf16(t):
    """Rounding hook of the oracle's bf16-emulation mode (torch tensors)."""
    return t.to(torch.bfloat16).float()


def bf16_np(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=F)).to(torch.bfloat16).float().numpy()


def to_oracle_geometry(data, b=None):
    """snap_b200.types containers -> oracle.geometry containers (same fp32 numbers)."""
    from oracle import geometry
    cam, T = data["camera"], data["T_view2scene"]
    sl = (lambda a: a) if b is None else (lambda a: a[b])
    if hasattr(cam, "k_radial") and cam.k_radial is not None:
        ocam = geometry.FisheyeCamera(wh=sl(cam.wh), f=sl(cam.f), c=sl(cam.c), k_radial=sl(cam.k_radial),
                                      max_fov=sl(cam.max_fov))
    else:
        ocam = geometry.Camera(wh=sl(cam.wh), f=sl(cam.f), c=sl(cam.c))
    return ocam, geometry.Transform3D(R=sl(T.R), t=sl(T.t))


def rel_l2(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30))


def assert_close_bf16(out, ref, what, atol_scale=2e-3, rtol=2.0 ** -7):
    """out: values that went through ONE bf16 rounding of an fp32-accumulated result; ref: fp32 oracle
    on identical inputs.  |err| <= rtol*|ref| + atol_scale*max|ref|."""
    out = np.asarray(out, dtype=F)
    ref = np.asarray(ref, dtype=F)
    assert out.shape == ref.shape, (what, out.shape, ref.shape)
    scale = float(np.abs(ref).max()) + 1e-12
    err = np.abs(out - ref)
    bad = err > rtol * np.abs(ref) + atol_scale * scale
    assert not bad.any(), f"{what}: {int(bad.sum())}/{bad.size} mismatches, max err {err.max():.4g}, scale {scale:.4g}"


def record_parity(block, what, err, tol):
    """Appends one measured error to gpurun_out/parity_table.jsonl (the per-block error table of profiles/rNN_parity_table.md
    is generated from it by tools/parity_table.py); silently does nothing when the directory is not writable."""
    import json, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    try:
        os.makedirs(os.path.join(root, "gpurun_out"), exist_ok=True)
        with open(os.path.join(root, "gpurun_out", "parity_table.jsonl"), "a") as f:
            f.write(json.dumps({"block": block, "what": what, "err": float(err), "tol": float(tol)}) + "\n")
    except OSError:
        pass
