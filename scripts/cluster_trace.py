"""Timeline of the dense cluster kernel (library built with -DSGD_CL_TRACE as libsgdnet_b200_cltrace.so, or any other
variant named in argv[1]): ten clock64 events per update for 4096 consecutive updates, CTA 0, reduced to where an update's
time goes. Also prints the plain epoch time of the same shapes.
Usage: python scripts/cluster_trace.py [variant] [shape ...]      shape: c3 | c4 | NxPxK:family"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sgdnet_b200 import _abi, api, synth

name = sys.argv[1] if len(sys.argv) > 1 else "cltrace"
so = "libsgdnet_b200.so" if name == "product" else f"libsgdnet_b200_{name}.so"
lib = _abi.Library(os.path.join(ROOT, "sgdnet_b200", so), "sgdnet_")
shapes = sys.argv[2:] or ["c3", "c4"]
for sh in shapes:
    if sh == "c3":
        n, p, K, fam, alpha = 60000, 784, 10, "multinomial", 0.8
    elif sh == "c4":
        n, p, K, fam, alpha = 60000, 2000, 4, "mgaussian", 1.0
    else:
        dims, fam = sh.split(":")
        n, p, K = (int(v) for v in dims.split("x"))
        alpha = 0.8
    gen = {"multinomial": synth.multinomial_dense, "mgaussian": synth.mgaussian_dense}[fam]
    x, y = gen(n, p, K, seed=1003)
    ymat = np.asarray(y, dtype=np.float64).reshape(n, -1)
    ctl, keep = api.build_control(fam, K, alpha=alpha, nlambda=100, lambda_min_ratio=1e-4, lambda_=None, maxit=1000, standardize=True,
                                  intercept=True, thresh=1e-3, standardize_response=False, debug=False)
    xa = np.asfortranarray(x, dtype=np.float64)
    ya = np.asfortranarray(ymat)
    sess = C.c_void_p()
    lib.check(lib.sym("session_create_dense")(_abi._ptr(xa, _abi.c_double_p), C.c_int64(n), C.c_int64(p), _abi._ptr(ya, _abi.c_double_p),
                                              C.c_int32(ya.shape[1]), C.byref(ctl), C.byref(sess)), "session_create_dense")
    rng = lib.rng_from_seed(1)
    ms = C.c_float(0)
    times = []
    for _ in range(3):
        lib.check(lib.sym("session_run_epochs")(sess, 30, 1, C.byref(rng), C.byref(ms)), "run_epochs")
        times.append(ms.value)
    lib.sym("session_destroy")(sess)
    t = min(times[1:])
    print(f"{sh} [{name}]: n={n} p={p} K={K} {fam}: epoch {t:.2f} ms = {n / t * 1e3:.0f} updates/s = {t * 1e-3 * 1.965e9 / n:.0f} cycles/update")
    if not hasattr(lib.lib, "sgdnet_debug_cluster_trace"):
        continue
    R = 4096
    buf = (C.c_longlong * (R * 10))()
    lib.lib.sgdnet_debug_cluster_trace(buf)
    a = np.array(buf[:], dtype=np.float64).reshape(R, 10)[8:-8]
    np.save(os.path.join(ROOT, "gpurun_out", f"cluster_trace_{sh}.npy"), a)
    c0, c1, c2, c3, f4, f5, f6, f7, f8, f9 = [a[:, i] for i in range(10)]
    print(f"  period {np.diff(c0).mean():.0f} (median {np.median(np.diff(c0)):.0f})")
    print(f"  feature warp 0: top->row in ring {np.mean(f5 - f4):.0f}; dot+butterfly {np.mean(f6 - f5):.0f}; ->past barrier 1 + constants {np.mean(f7 - f6):.0f}; "
          f"->past barrier 2 {np.mean(f8 - f7):.0f}; coefficient step {np.mean(f9 - f8):.0f}; ->next top {np.mean(f4[1:] - f9[:-1]):.0f}")
    print(f"  control warp: barrier 1 -> stores issued {np.mean(c1 - c0):.0f}; -> all sums arrived {np.mean(c2 - c1):.0f}; -> g_change stored {np.mean(c3 - c2):.0f}; "
          f"-> past next barrier 1 {np.mean(c0[1:] - c3[:-1]):.0f}")
    print(f"  critical path: f6(warp sums) -> c0 {np.mean(c0 - f6):.0f}; c3 -> f8 {np.mean(f8 - c3):.0f}")
