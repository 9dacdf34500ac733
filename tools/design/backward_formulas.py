"""Closed forms and launch decompositions for the backward kernels of round 2, written as NumPy so that
tests/test_backward_design_cpu.py can check them against torch autograd BEFORE any CUDA is written:

  * GroupNorm backward (resnet.py:46-70) as two per-(image, group) reductions + one elementwise pass,
  * StdConv weight-standardisation backward (resnet.py:34-41,73-79),
  * 3x3 / stride-1 conv backward on the zero-bordered activation layout of the forward GEMM engine:
      dX = a 9-segment GEMM over the zero-bordered dY with the taps mirrored (same engine, same segment mechanism),
      dW = 9 split-K products  X_bordered[m + off_tap]^T dY_bordered[m]  (snapb200_dense_wgrad with a row offset).

Design notes only: nothing here is imported by the product package.
"""
import numpy as np


def groupnorm_backward(x, dy, scale, groups=32, eps=1e-5):
    """x, dy [N,H,W,C]; y = (x - mu) * rstd * scale + bias with statistics over (H, W, C/groups).
    Returns dx, dscale, dbias.  Kernel plan: pass 1 accumulates per (n, g): s1 = sum(dy*scale), s2 = sum(dy*scale*xhat)
    (and per channel dscale, dbias); pass 2: dx = rstd * (dy*scale - s1/m - xhat * s2/m)."""
    N, H, W, C = x.shape
    cpg = C // groups
    xg = x.reshape(N, H * W, groups, cpg).astype(np.float64)
    mu = xg.mean(axis=(1, 3), keepdims=True)
    var = ((xg - mu) ** 2).mean(axis=(1, 3), keepdims=True)
    rstd = 1.0 / np.sqrt(var + eps)
    xhat = (xg - mu) * rstd
    g = (dy.astype(np.float64) * scale.reshape(1, 1, 1, C)).reshape(N, H * W, groups, cpg)
    m = H * W * cpg
    s1 = g.sum(axis=(1, 3), keepdims=True)
    s2 = (g * xhat).sum(axis=(1, 3), keepdims=True)
    dx = rstd * (g - s1 / m - xhat * s2 / m)
    dscale = (dy.astype(np.float64) * xhat.reshape(N, H, W, C)).sum(axis=(0, 1, 2))
    dbias = dy.astype(np.float64).sum(axis=(0, 1, 2))
    return dx.reshape(N, H, W, C), dscale, dbias


def stdconv_weight_backward(w, dws, eps=1e-10):
    """w [kh,kw,in,out] raw kernel, dws = gradient w.r.t. the standardised kernel ws = (w - mu) / sqrt(var + eps) with
    statistics over (kh, kw, in) per output channel.  dw = (dws - mean(dws) - ws * mean(dws * ws)) / sqrt(var + eps)."""
    w64, d64 = w.astype(np.float64), dws.astype(np.float64)
    mu = w64.mean(axis=(0, 1, 2), keepdims=True)
    var = ((w64 - mu) ** 2).mean(axis=(0, 1, 2), keepdims=True)
    rstd = 1.0 / np.sqrt(var + eps)
    ws = (w64 - mu) * rstd
    return rstd * (d64 - d64.mean(axis=(0, 1, 2), keepdims=True) - ws * (d64 * ws).mean(axis=(0, 1, 2), keepdims=True))


def bordered(x):
    """[N,H,W,C] -> zero-bordered rows [N*(H+2)*(W+2), C] (the layout gn_apply writes for the forward 3x3 GEMM)."""
    N, H, W, C = x.shape
    out = np.zeros((N, H + 2, W + 2, C), x.dtype)
    out[:, 1:-1, 1:-1] = x
    return out.reshape(-1, C)


def segment_gemm(a_rows, b, seg_off, seg_k, m_rows):
    """The engine's semantics (include/snapb200.h): out[m] = sum_s A[m + seg_off[s], :seg_k] @ B[:, s*seg_k:(s+1)*seg_k]^T,
    rows outside A read as zero (TMA zero fill)."""
    R = a_rows.shape[0]
    out = np.zeros((m_rows, b.shape[0]), np.float64)
    for s, off in enumerate(seg_off):
        idx = np.arange(m_rows) + off
        ok = (idx >= 0) & (idx < R)
        a = np.zeros((m_rows, seg_k))
        a[ok] = a_rows[idx[ok], :seg_k]
        out += a @ b[:, s * seg_k:(s + 1) * seg_k].T.astype(np.float64)
    return out


def conv3x3_forward_segments(x, w):
    """Forward as the engine runs it: bordered input rows, tap (kh,kw) = row offset (kh-1)(W+2)+(kw-1), B = [out, 9*in];
    returns the bordered-layout output [N,(H+2),(W+2),out] (only the interior is meaningful / stored)."""
    N, H, W, Cin = x.shape
    Cout = w.shape[-1]
    xb = bordered(x)
    seg = [(kh - 1) * (W + 2) + (kw - 1) for kh in range(3) for kw in range(3)]
    b = w.reshape(9 * Cin, Cout).T            # [out, tap*in + c]
    y = segment_gemm(xb, b, seg, Cin, xb.shape[0]).reshape(N, H + 2, W + 2, Cout)
    return y[:, 1:-1, 1:-1]


def conv3x3_dx_segments(dy, w):
    """dX[p] = sum_taps dY[p - off_tap] @ W[tap]^T: the SAME 9-segment GEMM over the zero-bordered dY with mirrored
    offsets and B = [in, 9*out] (tap-major, W[tap] transposed)."""
    N, H, W_, Cout = dy.shape
    Cin = w.shape[2]
    dyb = bordered(dy)
    seg = [-((kh - 1) * (W_ + 2) + (kw - 1)) for kh in range(3) for kw in range(3)]
    b = np.concatenate([w[kh, kw] for kh in range(3) for kw in range(3)], axis=1)   # [in, 9*out]
    dx = segment_gemm(dyb, b, seg, Cout, dyb.shape[0]).reshape(N, H + 2, W_ + 2, Cin)
    return dx[:, 1:-1, 1:-1]


def conv3x3_dw_shifted(x, dy):
    """dW[tap] = X_bordered[m + off_tap]^T dY_bordered[m], summed over all bordered rows m (the zero borders of dY kill
    the rows where the shifted read would wrap): nine calls of the split-K weight-gradient kernel with a row offset."""
    N, H, W_, Cin = x.shape
    Cout = dy.shape[-1]
    xb, dyb = bordered(x).astype(np.float64), bordered(dy).astype(np.float64)
    R = xb.shape[0]
    dw = np.zeros((3, 3, Cin, Cout))
    for kh in range(3):
        for kw in range(3):
            off = (kh - 1) * (W_ + 2) + (kw - 1)
            idx = np.arange(R) + off
            ok = (idx >= 0) & (idx < R)
            xs = np.zeros_like(xb)
            xs[ok] = xb[idx[ok]]
            dw[kh, kw] = xs.T @ dyb
    return dw


def pool_multiview_backward(feats, scores, valid, dmean, dvar, dsmax):
    """Backward of the weighted branch of pool_multiview_features (streetview_encoder.py:156-177) for ONE voxel:
    feats [V,D], scores [V], valid [V] (at least one valid) -> d feats [V,D], d scores [V].
      w = softmax over the valid views;  mean = sum w f;  var = sum w (f - mean)^2;  smax = max valid s
      d f_k = w_k dmean + 2 w_k (f_k - mean) dvar          (d var / d mean = -2 sum w (f - mean) = 0)
      d w_k = f_k . dmean + (f_k - mean)^2 . dvar
      d s_k = w_k (d w_k - sum_j w_j d w_j) + [k = argmax] dsmax
    The lift's backward kernel evaluates this per (voxel, view) from the gather records of the forward pass and
    scatter-adds d f_k through the four bilinear tap weights into the feature maps."""
    f, s = feats.astype(np.float64), scores.astype(np.float64)
    v = valid.astype(bool)
    mx = max(0.0, s[v].max())                        # jax.nn.softmax(where=, initial=0)
    e = np.where(v, np.exp(s - mx), 0.0)
    w = e / e.sum()
    mean = (w[:, None] * f).sum(0)
    df = w[:, None] * dmean[None] + 2 * w[:, None] * (f - mean) * dvar[None]
    dw = (f * dmean[None]).sum(-1) + ((f - mean) ** 2 * dvar[None]).sum(-1)
    ds = w * (dw - (w * dw).sum())
    k = int(np.argmax(np.where(v, s, -np.inf)))
    ds[k] += dsmax
    df[~v] = 0
    ds[~v] = 0
    return df, ds


def xcorr_backward(q, m, dS):
    """Backward of template_matching's score volume (pose_exhaustive_voting.py:72-104, before masking / normalisation):
    S_r[u,v] = sum_{i,j,d} q_r[i,j,d] m_pad[u+i, v+j, d] with the edge-padded map m_pad [(3G-2)^2, D].
      dq_r[i,j,d]    = sum_{u,v} dS_r[u,v] m_pad[u+i, v+j, d]     -> a correlation of m_pad with the (2G-1)^2 'template' dS_r:
                        the forward's sliding-window kernel with the roles (template rows <-> shifts) exchanged
      dm_pad[a,b,d]  = sum_r sum_{i,j} dS_r[a-i, b-j] q_r[i,j,d]   -> a full convolution = correlation of the zero-padded dS_r
                        with the flipped templates, summed over r (one GEMM with K = R * G^2)
      dm             = fold of dm_pad through the edge padding (every padded cell adds to the map cell it replicates).
    q [R,G,G,D], m [G,G,D], dS [R,2G-1,2G-1] -> dq [R,G,G,D], dm [G,G,D]."""
    R, G, _, D = q.shape
    U = 2 * G - 1
    P = 3 * G - 2
    idx = np.clip(np.arange(P) - (G - 1), 0, G - 1)          # padded row/col -> map row/col (jnp.pad mode='edge')
    m_pad = m[idx][:, idx].astype(np.float64)
    dq = np.zeros((R, G, G, D))
    for i in range(G):
        for j in range(G):
            dq[:, i, j] = np.einsum("ruv,uvd->rd", dS, m_pad[i:i + U, j:j + U])
    dm_pad = np.zeros((P, P, D))
    for i in range(G):
        for j in range(G):
            dm_pad[i:i + U, j:j + U] += np.einsum("ruv,rd->uvd", dS, q[:, i, j].astype(np.float64))
    dm = np.zeros((G, G, D))
    np.add.at(dm, (idx[:, None], idx[None, :]), dm_pad)
    return dq, dm


# ----------------------------------------------------------------------------------------------------------------------
# A whole pre-activation bottleneck unit (resnet.py:103-134, stride 1, identity shortcut) backward, composed ONLY of the
# building blocks above, in the order the round-2 launch plan will run them.  Saved from the forward: x, the three
# GroupNorm inputs (x, c1, c2) and their normalised+ReLU'd outputs (a1, a2, a3) -- exactly the tensors the forward plan
# already materialises (the raw conv outputs and the gn_apply outputs).
# ----------------------------------------------------------------------------------------------------------------------
def _gn_relu_forward(x, scale, bias, groups=32, eps=1e-5):
    N, H, W, C = x.shape
    xg = x.reshape(N, H * W, groups, C // groups).astype(np.float64)
    mu = xg.mean(axis=(1, 3), keepdims=True)
    var = ((xg - mu) ** 2).mean(axis=(1, 3), keepdims=True)
    y = ((xg - mu) / np.sqrt(var + eps)).reshape(N, H, W, C) * scale.reshape(1, 1, 1, C) + bias.reshape(1, 1, 1, C)
    return np.maximum(y, 0)


def _std(w, eps=1e-10):
    w = w.astype(np.float64)
    mu = w.mean(axis=(0, 1, 2), keepdims=True)
    return (w - mu) / np.sqrt(((w - mu) ** 2).mean(axis=(0, 1, 2), keepdims=True) + eps)


def residual_unit_forward(x, p):
    a1 = _gn_relu_forward(x, p["gn1"]["scale"], p["gn1"]["bias"])
    c1 = np.einsum("nhwi,io->nhwo", a1, _std(p["conv1"]["kernel"])[0, 0])
    a2 = _gn_relu_forward(c1, p["gn2"]["scale"], p["gn2"]["bias"])
    c2 = conv3x3_forward_segments(a2, _std(p["conv2"]["kernel"]))
    a3 = _gn_relu_forward(c2, p["gn3"]["scale"], p["gn3"]["bias"])
    c3 = np.einsum("nhwi,io->nhwo", a3, _std(p["conv3"]["kernel"])[0, 0])
    return c3 + x, dict(x=x, a1=a1, c1=c1, a2=a2, c2=c2, a3=a3)


def residual_unit_backward(dy, saved, p):
    """Returns dx and the parameter gradients {gn*/scale,bias, conv*/kernel}.  Launch order of round 2:
    conv3: dW (split-K), dX (engine)  -> ReLU mask + GroupNorm backward (2 reductions + 1 pass)
    conv2: 9 shifted dW, 9-segment dX -> ReLU mask + GroupNorm backward
    conv1: dW, dX                     -> ReLU mask + GroupNorm backward, + dy (identity shortcut)
    and one StdConv weight-standardisation backward per kernel."""
    g = {}
    w3, w2, w1 = _std(p["conv3"]["kernel"]), _std(p["conv2"]["kernel"]), _std(p["conv1"]["kernel"])
    # conv3 (1x1): dW = a3^T dy, da3 = dy W3^T
    dws3 = np.einsum("nhwi,nhwo->io", saved["a3"], dy)[None, None]
    da3 = np.einsum("nhwo,io->nhwi", dy, w3[0, 0])
    d = da3 * (saved["a3"] > 0)                                     # ReLU backward on the saved activation
    dc2, g["gn3/scale"], g["gn3/bias"] = groupnorm_backward(saved["c2"], d, p["gn3"]["scale"].reshape(-1))
    # conv2 (3x3): nine shifted products / nine-segment GEMM on the bordered layout
    dws2 = conv3x3_dw_shifted(saved["a2"], dc2)
    da2 = conv3x3_dx_segments(dc2, w2)
    d = da2 * (saved["a2"] > 0)
    dc1, g["gn2/scale"], g["gn2/bias"] = groupnorm_backward(saved["c1"], d, p["gn2"]["scale"].reshape(-1))
    # conv1 (1x1)
    dws1 = np.einsum("nhwi,nhwo->io", saved["a1"], dc1)[None, None]
    da1 = np.einsum("nhwo,io->nhwi", dc1, w1[0, 0])
    d = da1 * (saved["a1"] > 0)
    dx, g["gn1/scale"], g["gn1/bias"] = groupnorm_backward(saved["x"], d, p["gn1"]["scale"].reshape(-1))
    for name, dws in (("conv1", dws1), ("conv2", dws2), ("conv3", dws3)):
        g[name + "/kernel"] = stdconv_weight_backward(p[name]["kernel"], dws)
    return dx + dy, g                                               # identity shortcut (:134)


def lift_gather_backward(fimg_shape, taps, weights, bins, wb1, dfeat, dscore, D):
    """Backward of the bilinear gather + depth-score interpolation of ONE (voxel, view) pair (streetview_encoder.py:69-76,
    109-124; grids.py:116-137) into the projected feature map fimg [Hf,Wf,D+S]: the forward reads four (clamped) taps
    (r_k, c_k) with weights w_k = w_row * w_col, features f = sum_k w_k fimg[tap_k, :D] and score
    s = (1 - wb1) * sum_k w_k fimg[tap_k, D + b0] + wb1 * sum_k w_k fimg[tap_k, D + b1].
    The backward scatter-adds  w_k * dfeat  into the D feature channels of every tap and  w_k * (1 - wb1) * dscore,
    w_k * wb1 * dscore  into the two bin channels (clamped taps that coincide simply accumulate): one atomicAdd per
    (tap, channel), the true 'scatter' of the lift."""
    g = np.zeros(fimg_shape, np.float64)
    b0, b1 = bins
    for (r, c), w in zip(taps, weights):
        g[r, c, :D] += w * dfeat
        g[r, c, D + b0] += w * (1 - wb1) * dscore
        g[r, c, D + b1] += w * wb1 * dscore
    return g
