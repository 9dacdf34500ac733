"""CPU restatement of the google-research/snap hot path — TEST INFRASTRUCTURE ONLY.

This package restates, in NumPy / SciPy-free explicit loops / torch-CPU, the algorithm of the
reference files named in SURVEY.md §8(c).  It is the parity checker for the CUDA library and the
timed CPU baseline of bench.py.  Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
`--impl reference` legs of `bench.py` may import it; the product package `snap_b200` never does.

Pinning status: the reference ships no tests, golden vectors or fixtures and its
runtime (JAX/Flax/Scenic) is not installable in this image (F4).  The restatement is pinned by the
`tests/golden/make_golden*.py` scripts, which execute the reference's OWN source from /root/reference
under NumPy / SciPy / torch-CPU stand-ins for the `jax`, `flax.linen` and `optax` APIs
(`tests/golden/jaxshim/`) and store the outputs as fixtures:
  * pure functions: `snap/utils/{grids,geometry}.py`, `snap/models/{pose_exhaustive_voting,
    streetview_encoder,layers,pose_estimation}.py`, the losses of `semantic_net.py` / `bev_localizer.py`;
  * whole modules on stand-in `self` objects or the `flax.linen` stand-in: `ResNetV2` / `FPNDecoder` /
    `ImageEncoder` / `ResNetStage` / `MLP`, `StreetViewEncoder.__call__`, `VerticalPooling.__call__`,
    `BEVMapper.__call__`, `BEVLocalizer.__call__`, `SemanticNet.__call__`.
What remains **parity unpinned** are the semantics of the third-party primitives the stand-ins
substitute (jax.scipy.ndimage.map_coordinates, jax.scipy.signal.convolve, lax.top_k,
jax.nn.softmax(where=), jax.image.resize, jax.random.choice's stream, optax cross-entropies, flax
Conv / Dense / max_pool / auto-naming): no JAX build is available to confirm them.
"""
