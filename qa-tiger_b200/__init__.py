"""qa-tiger_b200: B200-native (sm_100a) implementation of the QA-TIGER batched-eval hot path.

Layout (only what the path needs):
  csrc/      hand-written CUDA kernels + the C ABI (include/qat.h) -> libqat_b200.so
  build.py   nvcc recipe (in-tree, sm_100a only)
  _cabi.py   ctypes binding of the C ABI; fails loudly when the library is missing
  models/    drop-in mirrors of the reference's src/models/{modules,net,encoders}.py
  tome/      drop-in mirror of the reference's src/tome/merge.py
  evalloop.py  data-parallel eval driver (mirror of src/trainutils.py:test)
"""
__version__ = "0.1.0"
