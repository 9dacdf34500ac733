"""CPU restatement of the google-research/snap hot path — TEST INFRASTRUCTURE ONLY.

This package restates, in NumPy / SciPy-free explicit loops / torch-CPU, the algorithm of the
reference files named in SURVEY.md §8(c).  It is the parity checker for the CUDA library and the
timed CPU baseline of bench.py.  Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
`--impl reference` legs of `bench.py` may import it; the product package `snap_b200` never does.

Pinning status: the reference ships no tests, golden vectors or fixtures and its
runtime (JAX/Flax/Scenic) is not installable in this image (F4).  The restatement is pinned by
`tests/golden/make_golden.py`, which executes the reference's OWN pure functions
(`snap/utils/{grids,geometry}.py`, `snap/models/pose_exhaustive_voting.py`,
`snap/models/streetview_encoder.py`, `snap/models/layers.py`, `snap/models/bev_mapper.py:VerticalPooling`)
from /root/reference under a NumPy/SciPy stand-in for the `jax` API and stores their outputs as
fixtures.  The third-party primitives themselves (jax.scipy.ndimage.map_coordinates,
jax.scipy.signal.convolve, lax.top_k, jax.nn.softmax(where=), flax Conv/Dense) remain
**parity unpinned**: no JAX build is available to confirm their semantics.
The Flax modules (ResNetV2, FPNDecoder, MLP, BEVMapper.__call__) cannot run under the stand-in and
are pinned only by the known-answer tests of SURVEY.md §4.1.
"""
