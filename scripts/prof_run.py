"""Short workloads for ncu captures (profiles/): python scripts/prof_run.py sparse|dense3|dense4|score [epochs]
sparse: config 2 (1M x 100k): `epochs` SAGA epochs at lambda[30] + 3 deviance passes through the stepping interface.
score: held-out scoring of a cv fold of config 5 (50k rows x 50k, 50 nnz/row, 100 lambdas): predict_score_kernel."""
import ctypes as C
import os
import sys

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sgdnet_b200 import _abi, api, synth
import sgdnet_b200 as sg

what = sys.argv[1] if len(sys.argv) > 1 else "sparse"
epochs = int(sys.argv[2]) if len(sys.argv) > 2 else 3
lib = sg.product() if not os.environ.get("SGDNET_VARIANT") else _abi.Library(os.path.join(ROOT, "sgdnet_b200", "libsgdnet_b200_" + os.environ["SGDNET_VARIANT"] + ".so"), "sgdnet_")
ms = C.c_float(0)
sess = C.c_void_p()
if what == "score":
    import time
    x, y = synth.binomial_sparse(50_000, 50_000, 50, seed=1005)
    L, p = 100, 50_000
    g = np.random.default_rng(1)
    beta = (g.normal(size=(L, p, 1)) * (g.uniform(size=(L, p, 1)) < np.linspace(0.001, 0.2, L)[:, None, None])).astype(np.float64)
    a0 = g.normal(size=(L, 1))
    for _ in range(2):
        t0 = time.perf_counter()
        sc = lib.score_deviance(x, y, 1, a0, beta)
        print(f"score: 50000 x 50000, 50 nnz/row, {L} lambdas: {1e3 * (time.perf_counter() - t0):.1f} ms per call (host buffers in), dev[0]={sc[0]:.6f}")
    sys.exit(0)
if what == "sparse":
    x, y = synth.binomial_sparse(1_000_000, 100_000, 100, seed=1002)
    m = _abi.CscMatrix.from_any(x)
    n, p = m.shape
    ya = np.ascontiguousarray(y.reshape(-1, 1))
    ctl, keep = api.build_control("binomial", 1, alpha=1.0, nlambda=100, lambda_min_ratio=1e-4, lambda_=None, maxit=1000,
                                  standardize=False, intercept=True, thresh=1e-3, standardize_response=False, debug=False)
    lib.check(lib.sym("session_create_sparse")(_abi._ptr(m.i, _abi.c_int32_p), _abi._ptr(m.p, _abi.c_int32_p), _abi._ptr(m.x, _abi.c_double_p),
                                               C.c_int64(n), C.c_int64(p), _abi._ptr(ya, _abi.c_double_p), C.c_int32(1), C.byref(ctl),
                                               C.byref(sess)), "create")
else:
    if what == "dense3":
        x, y = synth.multinomial_dense(60_000, 784, 10, seed=1003)
        fam, K, alpha = "multinomial", 10, 0.8
    else:
        x, y = synth.mgaussian_dense(200_000, 2000, 4, seed=1004)
        fam, K, alpha = "mgaussian", 4, 1.0
    n, p = x.shape
    ya = np.asfortranarray(np.asarray(y, dtype=np.float64).reshape(n, -1))
    xa = np.asfortranarray(x)
    ctl, keep = api.build_control(fam, K, alpha=alpha, nlambda=100, lambda_min_ratio=1e-4, lambda_=None, maxit=1000, standardize=True,
                                  intercept=True, thresh=1e-3, standardize_response=False, debug=False)
    lib.check(lib.sym("session_create_dense")(_abi._ptr(xa, _abi.c_double_p), C.c_int64(n), C.c_int64(p), _abi._ptr(ya, _abi.c_double_p),
                                              C.c_int32(ya.shape[1]), C.byref(ctl), C.byref(sess)), "create")
rng = lib.rng_from_seed(1)
t = []
for _ in range(epochs):
    lib.check(lib.sym("session_run_epochs")(sess, 30, 1, C.byref(rng), C.byref(ms)), "run")
    t.append(ms.value)
d = []
for _ in range(3):
    lib.check(lib.sym("session_finish_lambda")(sess, 30, C.byref(ms)), "finish")
    d.append(ms.value)
d2 = []
if what == "sparse":      # the state a warm-started lasso path sees: a few epochs at a strong penalty zero most weights
    lib.check(lib.sym("session_run_epochs")(sess, 3, 3, C.byref(rng), C.byref(ms)), "run")
    for _ in range(3):
        lib.check(lib.sym("session_finish_lambda")(sess, 3, C.byref(ms)), "finish")
        d2.append(ms.value)
lib.sym("session_destroy")(sess)
print(f"{what}: n={n} p={p} epoch ms {t} deviance pass ms {d} (all weights live) {d2} (lasso-sparse weights)")
