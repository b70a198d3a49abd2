"""Single-fit epoch rate of the sparse wavefront kernel by penalty (ridge / elastic net / lasso) at config 5's shape."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sgdnet_b200 as sg
from sgdnet_b200 import _abi, api, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 500000
p = int(sys.argv[2]) if len(sys.argv) > 2 else 50000
nnz = int(sys.argv[3]) if len(sys.argv) > 3 else 50
lib = sg.product()
x, y = synth.binomial_sparse(n, p, nnz, seed=1005)
m = _abi.CscMatrix.from_any(x)
ya = np.ascontiguousarray(y.reshape(-1, 1))
for alpha in (0.0, 0.5, 1.0):
    for li in (2, 8):
        ctl, keep = api.build_control("binomial", 1, alpha=alpha, nlambda=10, lambda_min_ratio=1e-4, lambda_=None, maxit=1000,
                                      standardize=False, intercept=True, thresh=1e-3, standardize_response=False, debug=False)
        sess = C.c_void_p()
        lib.check(lib.sym("session_create_sparse")(_abi._ptr(m.i, _abi.c_int32_p), _abi._ptr(m.p, _abi.c_int32_p),
                                                  _abi._ptr(m.x, _abi.c_double_p), C.c_int64(n), C.c_int64(p),
                                                  _abi._ptr(ya, _abi.c_double_p), C.c_int32(1), C.byref(ctl), C.byref(sess)), "create")
        rng = lib.rng_from_seed(1)
        ms = C.c_float(0)
        for it in range(3):
            lib.check(lib.sym("session_run_epochs")(sess, li, 1, C.byref(rng), C.byref(ms)), "run")
        print(f"alpha={alpha} lambda_ind={li}: {n / (ms.value * 1e-3) / 1e6:.2f} M updates/s ({ms.value:.1f} ms/epoch)", flush=True)
        lib.sym("session_destroy")(sess)
