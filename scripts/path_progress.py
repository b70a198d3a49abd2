"""lambda-path fit time (BASELINE metric ii) at config 2's FULL size, lambda by lambda through the stepping interface
(sgdnet_session_fit_lambda = Saga() + Deviance + Rescale for one lambda, warm-started, thresh 1e-3, maxit 1000), with
the design resident in HBM. Writes one JSON line per lambda (epochs, seconds) as it goes and stops after BUDGET seconds
of solver wall time, so a bounded GPU slot still yields the longest prefix of the 100-lambda path it can.
Usage: python scripts/path_progress.py [BUDGET_SECONDS] [OUT.jsonl]"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sgdnet_b200 import _abi, api, synth
import sgdnet_b200 as sg

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 600.0
out_path = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "gpurun_out", "path_progress.jsonl")
n, p = 1_000_000, 100_000
lib = sg.product()
x, y = synth.binomial_sparse(n, p, 100, seed=1002)
m = _abi.CscMatrix.from_any(x)
ya = np.ascontiguousarray(y.reshape(-1, 1))
ctl, keep = api.build_control("binomial", 1, alpha=1.0, nlambda=100, lambda_min_ratio=1e-4, lambda_=None, maxit=1000,
                              standardize=False, intercept=True, thresh=1e-3, standardize_response=False, debug=False)
t_create = time.perf_counter()
sess = C.c_void_p()
lib.check(lib.sym("session_create_sparse")(_abi._ptr(m.i, _abi.c_int32_p), _abi._ptr(m.p, _abi.c_int32_p),
                                          _abi._ptr(m.x, _abi.c_double_p), C.c_int64(n), C.c_int64(p),
                                          _abi._ptr(ya, _abi.c_double_p), C.c_int32(1), C.byref(ctl), C.byref(sess)), "create")
t_create = time.perf_counter() - t_create
rng = lib.rng_from_seed(1)
ep, conv = C.c_uint32(0), C.c_int32(0)
total_epochs, t0 = 0, time.perf_counter()
with open(out_path, "w") as fh:
    fh.write(json.dumps({"workload": "config 2: binomial lasso, sparse 1M x 100k, 100 nnz/row, 100-lambda path, thresh 1e-3, set.seed(1)",
                         "session_create_s": t_create}) + "\n")
    for li in range(100):
        t1 = time.perf_counter()
        lib.check(lib.sym("session_fit_lambda")(sess, li, C.byref(rng), C.byref(ep), C.byref(conv)), "fit_lambda")
        dt = time.perf_counter() - t1
        total_epochs += ep.value
        rec = {"lambda_ind": li, "epochs": ep.value, "converged": bool(conv.value), "seconds": dt,
               "cum_epochs": total_epochs, "cum_seconds": time.perf_counter() - t0}
        fh.write(json.dumps(rec) + "\n")
        fh.flush()
        if time.perf_counter() - t0 > budget:
            break
    res = _abi.Result()
    lib.check(lib.sym("session_result")(sess, C.byref(res)), "result")
    raw = lib.take_result(res)
    done = li + 1
    fh.write(json.dumps({"lambdas_done": done, "npasses": total_epochs, "wall_s": time.perf_counter() - t0,
                         "updates_per_s": n * total_epochs / (time.perf_counter() - t0),
                         "nonzeros_per_lambda": [int(np.count_nonzero(raw.beta[l])) for l in range(done)],
                         "dev_ratio": [float(v) for v in raw.dev_ratio[:done]]}) + "\n")
lib.sym("session_destroy")(sess)
print(open(out_path).read()[-1500:])
