"""B200-native hot path of Video-to-Video Few-Shot Patch-Based Training.

Import name: ``pbt_b200`` (the on-disk directory carries the reference's name, which is not a
valid Python identifier; ``pbt_b200/__init__.py`` at the repo root aliases it).

Layout
  csrc/       hand-written sm_100a CUDA (tcgen05 implicit-GEMM convs, HBM-bound elementwise, gather) + C-ABI
  _native.py  ctypes binding of libpbt.so (no fallback)
  ops.py      thin op wrappers + weight packing
  generator.py  drop-in GeneratorJ (reference src/models/generator.py)
  sampler.py    drop-in StyleTransferDataset / batched device sampler (reference src/data/dataset.py)
"""
__version__ = "0.1.0"
