"""sgdnet_b200 — B200 (sm_100a) SAGA backend for sgdnet behind a C ABI.

`csrc/` holds the CUDA kernels and the extern "C" library (libsgdnet_b200.so, built in-tree by
`__graft_entry__.build()`); `api.py` is the host-side mirror of sgdnet's R front end; `_abi.py` the
ctypes binding of include/sgdnet_b200.h. There is no CPU fallback: without the built library or
without a CUDA device every fit raises.
"""
from ._abi import SgdnetError, Library, CscMatrix, product, product_library_path  # noqa: F401
from .api import (sgdnet, cv_sgdnet, predict, coef, deviance, score, SgdnetFit, CvSgdnet,  # noqa: F401
                  make_foldid, cv_plan, lambda_interpolate)

__all__ = ["sgdnet", "cv_sgdnet", "predict", "coef", "deviance", "score", "SgdnetFit", "CvSgdnet", "SgdnetError",
           "Library", "CscMatrix", "product", "product_library_path", "make_foldid", "cv_plan", "lambda_interpolate"]
