"""lambda-path fit time (BASELINE metric ii) at config 2's full size, bounded to the first L values of the automatic
100-lambda path (warm-started, thresh = 1e-3): one sgdnet_fit_sparse call through the C ABI, host buffers in, archives
out. With --cpu the oracle (libm arithmetic, one core) runs the same call for comparison.
Usage: python tests/measure/path_bench.py [L] [--cpu]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import sgdnet_b200 as sg
from sgdnet_b200 import synth

L = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 8
n, p = 1_000_000, 100_000
x, y = synth.binomial_sparse(n, p, 100, seed=1002)
yc = y - y.mean()
lmax = float(np.abs(x.T @ yc).max() / n)
lam = np.exp(np.log(lmax) + np.arange(100) * (np.log(lmax * 1e-4) - np.log(lmax)) / 99.0)[:L]
kw = dict(family="binomial", alpha=1.0, standardize=False, intercept=True, lambda_=list(lam), thresh=1e-3, maxit=1000, seed=1)
out = {"workload": f"config 2 (1M x 100k, 100 nnz/row, binomial lasso), first {L} of the 100 automatic lambdas, thresh 1e-3"}
t0 = time.perf_counter()
g = sg.sgdnet(x, y, backend=sg.product(), **kw)
out["gpu_wall_s"] = time.perf_counter() - t0
out["gpu_solver_s"] = g.raw.seconds_solver
out["gpu_setup_s"] = g.raw.seconds_setup
out["epochs_per_lambda"] = [int(e) for e in g.epochs]
out["npasses"] = int(g.npasses)
out["nonzeros_last"] = int(np.count_nonzero(g.raw.beta[-1]))
if "--cpu" in sys.argv:
    from oracle_lib import load_oracle
    oracle = load_oracle()
    oracle.lib.oracle_set_arith(0)
    t0 = time.perf_counter()
    r = sg.sgdnet(x, y, backend=oracle, **kw)
    out["cpu_wall_s"] = time.perf_counter() - t0
    out["cpu_solver_s"] = r.raw.seconds_solver
    out["cpu_npasses"] = int(r.npasses)
    out["max_rel_coef_diff_vs_libm_oracle"] = float(np.max(np.abs(g.raw.beta - r.raw.beta)) / max(np.max(np.abs(r.raw.beta)), 1e-300))
print(json.dumps(out))
