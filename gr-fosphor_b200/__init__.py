"""fosphor-b200: B200-native (sm_100a CUDA) engine for gr-fosphor's spectral hot
path, behind the reference's ``fosphor_cl_*`` C surface.

Layout:
  csrc/        CUDA kernels + the C-ABI (include/fosphor_b200.h) -> libfosphor_b200.so
  build.py     nvcc build of the shared library (in-tree)
  engine.py    ctypes mirror of the parameterised ``fosphor_cu_*`` API
  dropin.py    ctypes mirror of the reference's libfosphor facade over ``fosphor_cl_*``

Python is only the test / bench harness; the product is the shared library.
"""
__version__ = "0.1.0"
