"""B200-native (sm_100a) DiT train step behind the reference's ``model.py`` API.

Layout
  csrc/      hand-written CUDA kernels + the C ABI (``include/vds_b200.h``) -> ``libvds_b200.so``
  lib.py     ctypes binding of the C ABI (raw device pointers; no torch types cross the boundary)
  ops.py     thin tensor-level wrappers (allocate outputs, pass pointers + the current stream)
  model.py   drop-in mirror of the reference ``model.py`` (DiT, apply_fsdp, get_mup_setup, ...)
  engine.py  hand-derived forward/backward of the whole DiT on top of ops.py
  train.py   mirror of the arithmetic of the reference ``train.py:forward`` + fused AdamW
"""
PKG_DIR = __import__("os").path.dirname(__import__("os").path.abspath(__file__))
