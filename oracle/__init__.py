"""ORACLE — test infrastructure, NOT product code.

CPU restatements of the reference (autonomousvision/volsurfs) algorithms for the
per-ray rendering hot path.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import anything
from this package; the product (``volsurfs_b200``) never does.

Modules
-------
compositing   dense K-layer torch path (volsurfs.py:601-640,708) + packed
              VolumeRendering ops (VolumeRenderingGPU.cuh) in numpy
packing       RaySamplesPacked container + compact_to_valid_samples + K-layer hit
              packing (RaySamplesPacked.cu, RaySamplesPackedGPU.cuh)
raytrace      ctypes wrapper around raytrace_oracle.c (triangle.cuh,
              bounding_box.cuh, bvh.cu restated in plain C, no FMA contraction)
appearance    SH encoder + GELU MLP + alpha decay (mlp.py, rgb.py,
              sphericalharmonics.py, volsurfs.py:583-594) in torch fp32/fp64
"""
