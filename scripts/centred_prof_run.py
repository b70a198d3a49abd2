"""Sparse + standardize = TRUE at reduced p (20000 x 2000, 16 nnz/row): epochs of saga_sparse_centred_kernel (for ncu)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from sgdnet_b200 import _abi, api, synth
import sgdnet_b200 as sg
lib = sg.product() if len(sys.argv) < 2 else _abi.Library(os.path.join(ROOT, "sgdnet_b200", f"libsgdnet_b200_{sys.argv[1]}.so"), "sgdnet_")
x, y = synth.binomial_sparse(20_000, 2000, 16, seed=1012)
m = _abi.CscMatrix.from_any(x)
n, p = m.shape
ya = np.ascontiguousarray(y.reshape(-1, 1))
ctl, keep = api.build_control("binomial", 1, alpha=1.0, nlambda=100, lambda_min_ratio=1e-4, lambda_=None, maxit=1000, standardize=True,
                              intercept=True, thresh=1e-3, standardize_response=False, debug=False)
sess = C.c_void_p()
lib.check(lib.sym("session_create_sparse")(_abi._ptr(m.i, _abi.c_int32_p), _abi._ptr(m.p, _abi.c_int32_p), _abi._ptr(m.x, _abi.c_double_p),
                                           C.c_int64(n), C.c_int64(p), _abi._ptr(ya, _abi.c_double_p), C.c_int32(1), C.byref(ctl), C.byref(sess)), "create")
rng = lib.rng_from_seed(1)
ms = C.c_float(0)
for _ in range(3):
    lib.check(lib.sym("session_run_epochs")(sess, 30, 1, C.byref(rng), C.byref(ms)), "run")
    print(f"epoch {ms.value:.2f} ms = {ms.value * 1e-3 * 1.965e9 / n:.0f} cycles/update")
if hasattr(lib.lib, "sgdnet_debug_centred_trace"):
    buf = (C.c_longlong * 8)()
    lib.lib.sgdnet_debug_centred_trace(buf)
    v = [b / n for b in buf]
    print(f"thread 0, cycles per update: pass over owned features {v[0]:.0f} | barrier 1 {v[1]:.0f} | dot + gradient {v[2]:.0f} | barrier 2 {v[3]:.0f}")
lib.sym("session_destroy")(sess)
