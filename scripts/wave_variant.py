"""Epoch time of the sparse wavefront kernel for a measurement variant of the library
(sgdnet_b200/libsgdnet_b200_<NAME>.so, scripts/build_variant.sh; NAME = product for the shipped library).
Usage: python scripts/wave_variant.py NAME [n] [p] [epochs] [family] [intercept 0/1]"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sgdnet_b200 import _abi, api, synth

name = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 200000
p = int(sys.argv[3]) if len(sys.argv) > 3 else 100000
epochs = int(sys.argv[4]) if len(sys.argv) > 4 else 4
family = sys.argv[5] if len(sys.argv) > 5 else "binomial"
icpt = bool(int(sys.argv[6])) if len(sys.argv) > 6 else True
so = "libsgdnet_b200.so" if name == "product" else f"libsgdnet_b200_{name}.so"
lib = _abi.Library(os.path.join(ROOT, "sgdnet_b200", so), "sgdnet_")
x, y = synth.binomial_sparse(n, p, 100, seed=1002)
m = _abi.CscMatrix.from_any(x)
ya = np.ascontiguousarray(y.reshape(-1, 1))
ctl, keep = api.build_control(family, 1, alpha=1.0, nlambda=100, lambda_min_ratio=1e-4, lambda_=None, maxit=1000,
                              standardize=False, intercept=icpt, thresh=1e-3, standardize_response=False, debug=False)
sess = C.c_void_p()
lib.check(lib.sym("session_create_sparse")(_abi._ptr(m.i, _abi.c_int32_p), _abi._ptr(m.p, _abi.c_int32_p),
                                          _abi._ptr(m.x, _abi.c_double_p), C.c_int64(n), C.c_int64(p),
                                          _abi._ptr(ya, _abi.c_double_p), C.c_int32(1), C.byref(ctl), C.byref(sess)), "create")
rng = lib.rng_from_seed(1)
ms = C.c_float(0)
times = []
for it in range(epochs):
    lib.check(lib.sym("session_run_epochs")(sess, 30, 1, C.byref(rng), C.byref(ms)), "run")
    times.append(ms.value)
t = min(times[1:]) if len(times) > 1 else times[0]
S = int(os.environ.get("SGDNET_WAVE_WARPS", "8"))
line = f"{name} {family} intercept={int(icpt)} S={S} n={n} p={p}: epoch {t:.1f} ms = {t * 1e-3 * 1.965e9 / n:.0f} cycles/row, {n / t / 1e3:.3f} M updates/s"
print(line, flush=True)
if os.environ.get("WAVE_VARIANT_ALL"):
    print("  epoch ms:", " ".join(f"{v:.1f}" for v in times), flush=True)
