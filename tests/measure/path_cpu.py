"""The same 100-lambda path of config 2 at FULL size on the CPU: the restated oracle in libm arithmetic (bit-identical to
the reference's compiled sources, tests/test_ref_cpu.py) or, with --ref, oracle/_ref itself (the reference's own
src/sgdnet.cpp; returns npasses but no per-lambda epochs). One core: the reference is single-threaded. Takes the better
part of an hour. Usage: python tests/measure/path_cpu.py [--ref] [OUT.json]"""
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
from oracle_lib import load_oracle, load_reference_build

use_ref = "--ref" in sys.argv
outs = [a for a in sys.argv[1:] if not a.startswith("--")]
out_path = outs[0] if outs else os.path.join(ROOT, "gpurun_out", "path_cpu_ref.json" if use_ref else "path_cpu_oracle.json")
n, p = 1_000_000, 100_000
x, y = synth.binomial_sparse(n, p, 100, seed=1002)
if use_ref:
    lib = load_reference_build()
    lib.lib.ref_set_force_debug(0)
else:
    lib = load_oracle()
    lib.lib.oracle_set_arith(0)
t0 = time.perf_counter()
fit = sg.sgdnet(x, y, family="binomial", alpha=1.0, standardize=False, intercept=True, nlambda=100, thresh=1e-3, maxit=1000,
                seed=1, backend=lib)
wall = time.perf_counter() - t0
rec = {"arm": "oracle/_ref (reference src/sgdnet.cpp on the Rcpp/Eigen stand-in)" if use_ref else "oracle, libm arithmetic",
       "workload": "config 2: binomial lasso, sparse 1M x 100k, 100 nnz/row, 100-lambda path, thresh 1e-3, set.seed(1)",
       "host": f"{os.cpu_count()} vCPU build container, 1 thread", "wall_s": wall, "npasses": int(fit.npasses),
       "epochs_per_lambda": [int(e) for e in fit.raw.epochs], "updates_per_s": n * int(fit.npasses) / wall,
       "nonzeros_per_lambda": [int(np.count_nonzero(fit.raw.beta[l])) for l in range(len(fit.lambda_))],
       "dev_ratio": [float(v) for v in fit.raw.dev_ratio], "lambda": [float(v) for v in fit.lambda_]}
json.dump(rec, open(out_path, "w"))
print(json.dumps({k: v for k, v in rec.items() if not isinstance(v, list)}))
