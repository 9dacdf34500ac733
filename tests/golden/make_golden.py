"""Generate golden fixtures by executing the REFERENCE'S OWN source (pure functions of /root/reference) under the
NumPy/SciPy stand-in for jax (tests/golden/jaxshim).  Run in the build container only:

    python tests/golden/make_golden.py            # writes tests/golden/*.npz

tests/test_golden.py then checks the oracle (oracle/) against these fixtures anywhere (no /root/reference needed).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
REF = os.environ.get("SNAP_REFERENCE", "/root/reference")

import jaxshim  # noqa: E402

jaxshim.install(REF)

from snap.utils import geometry, grids  # noqa: E402  (the reference's own files)
from snap.models import layers, pose_exhaustive_voting as pev, streetview_encoder as sve  # noqa: E402
from snap.models import image_encoder as ie, resnet  # noqa: E402

F = np.float32
rng = np.random.default_rng(20240925)
out = {}


def rot(yaw, pitch):
    cy, sy, cp, sp = np.cos(yaw), np.sin(yaw), np.cos(pitch), np.sin(pitch)
    return (np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]]) @ np.array([[1, 0, 0], [0, cp, -sp], [0, sp, cp]])).astype(F)


# ---- grids.interpolate_nd -----------------------------------------------------------------------
arr = rng.standard_normal((9, 11, 4)).astype(F)
pts = (rng.random((200, 2)) * [11, 13] - 1).astype(F)
mask = rng.random((9, 11)) > 0.2
v, ok = grids.interpolate_nd(arr, pts)
v2, ok2 = grids.interpolate_nd(arr, pts, mask)
out["interp"] = dict(arr=arr, pts=pts, mask=mask, val=v, valid=ok, valid_masked=ok2)

# ---- geometry -------------------------------------------------------------------------------------
V, N = 3, 300
R = np.stack([rot(rng.uniform(-3, 3), rng.uniform(-0.3, 0.3)) @ np.array([[0, 0, 1], [-1, 0, 0], [0, -1, 0]], F) for _ in range(V)])
t = (rng.standard_normal((V, 3)) * [3, 3, 0.3] + [5, 5, 2]).astype(F)
T = geometry.Transform3D(R=R, t=t)
pts3 = (rng.random((N, 3)) * [10, 10, 6] + [0, 0, -1]).astype(F)
cam = geometry.Camera(wh=np.tile(F([160, 120]), (V, 1)), f=np.tile(F([110, 108]), (V, 1)), c=np.tile(F([80.5, 59.5]), (V, 1)))
fcam = geometry.FisheyeCamera(wh=cam.wh, f=cam.f, c=cam.c, k_radial=np.tile(F([-0.03, 0.005, 0.0]), (V, 1)),
                              max_fov=np.full((V,), np.deg2rad(115.0), F))
Ti = T.inv
out["geometry"] = dict(R=R, t=t, pts=pts3, Rinv=Ti.R, tinv=Ti.t, wh=cam.wh, f=cam.f, c=cam.c,
                       cam_scaled_f=cam.scale(F([0.25, 0.25])).f, k_radial=fcam.k_radial, max_fov=fcam.max_fov)
p2d, vis, depth, rays = sve.project_points_to_views(T, cam, pts3)
out["project"] = dict(p2d=p2d, vis=vis, depth=depth, rays=rays)
p2df, visf, _, _ = sve.project_points_to_views(T, fcam, pts3)
out["project_fisheye"] = dict(p2d=p2df, vis=visf)

# ---- lift pieces -----------------------------------------------------------------------------------
B, Hf, Wf, D, S = 1, 30, 40, 6, 8
fimg = rng.standard_normal((B, V, Hf, Wf, D + S)).astype(F)
cams = cam.scale(F([0.25, 0.25]))
p2d_s, vis_s, depth_s, _ = sve.project_points_to_views(T, cams, pts3)
f_all = sve.interpolate_views_all(fimg, p2d_s[None].swapaxes(1, 2))  # [B,N,V,D+S]
out["interp_views_all"] = dict(fimg=fimg, p2d=p2d_s, out=f_all)
scores = sve.interpolate_depth_score(f_all[..., D:], depth_s[None], (1.0, 32.0))
out["depth_score"] = dict(scales=f_all[..., D:], depth=depth_s, out=scores)
for tag, kw in (("weighted", dict(scores=scores)), ("plain", dict(scores=None))):
    st, va = sve.pool_multiview_features(f_all[..., :D], vis_s[None], kw["scores"], False, True)
    out["pool_" + tag] = dict(feats=f_all[..., :D], valid=vis_s, stats=st, valid_any=va)
st, va = sve.pool_multiview_features(f_all[..., :D], vis_s[None], scores, True, True)
out["pool_minmax"] = dict(stats=st)
idx, mind = sve.view_selection(pts3[None], geometry.Transform3D(R=R[None], t=t[None]), vis_s[None], 2)
out["view_selection"] = dict(idx=idx, min_dist=mind)
p2d_sel = np.take_along_axis(p2d_s[None], idx[..., None], 2)
f_sel = sve.interpolate_views_selective(jaxshim.clamp_indexing(fimg), p2d_sel, idx)  # JAX clamps OOB gather indices
out["interp_views_selective"] = dict(p2d=p2d_sel, idx=idx, out=f_sel)

# ---- layers ----------------------------------------------------------------------------------------
x = rng.standard_normal((50, 7)).astype(F); x[3] = 0; x[4] = 1e-7
out["normalize"] = dict(x=x, out=layers.normalize(x))
m = rng.random((50, 7)) > 0.5; m[5] = False
out["masked"] = dict(x=x, mask=m, mean=layers.masked_mean(x, m, -1), softmax=layers.masked_softmax(x, m, -1))

# ---- resnet.standardize / pad_to_multiple --------------------------------------------------------------
w = (rng.standard_normal((3, 3, 5, 4)) * 2 + 0.3).astype(F)
out["standardize"] = dict(w=w, out=resnet.standardize(w, axis=(0, 1, 2), eps=1e-10),
                          gn=resnet.standardize(w.reshape(1, 3, 3, 5, 4), axis=(1, 2, 4), eps=1e-5))
img = rng.random((2, 10, 16, 3)).astype(F)
out["pad"] = dict(img=img, p8=ie.pad_to_multiple(img, 8), p32=ie.pad_to_multiple(img, 32))

# ---- exhaustive pose voting ------------------------------------------------------------------------------
G, Rr, Dm = 12, 8, 5
grid = grids.Grid2D((G, G), 0.2)
fq = rng.standard_normal((G, G, Dm)).astype(F); fm = rng.standard_normal((G, G, Dm)).astype(F)
vq = rng.random((G, G)) > 0.15; vm = rng.random((G, G)) > 0.1
fq *= vq[..., None]; fm *= vm[..., None]
tq, tv = pev.sample_query_templates(fq, vq, Rr, grid)
sc = pev.template_matching(tq, tv, fm, vm)
conf = rng.random((G, G)).astype(F)
sc_conf = pev.exhaustive_pose_voting(pev.types.FeaturePlane(features=fq.copy(), valid=vq),  # copy: `feats_q *= conf` is in-place under NumPy (JAX arrays are immutable)
                                     pev.types.FeaturePlane(features=fm, valid=vm),
                                     Rr, grid, conf)
tfm = pev.exhaustive_index_to_tfm(np.array([3, 14, 9]), grid, Rr)
back = pev.exhaustive_tfm_to_index(tfm, grid, Rr)
out["voting"] = dict(fq=fq, vq=vq, fm=fm, vm=vm, templates=tq, t_valid=tv, scores=sc, conf=conf, scores_conf=sc_conf,
                     tfm_angle=tfm.angle, tfm_t=tfm.t, index_back=back)

for name, d in out.items():
    np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **{k: np.asarray(v) for k, v in d.items()})
    print(name, {k: np.asarray(v).shape for k, v in d.items()})
