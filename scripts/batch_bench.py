"""Many independent fits on one GPU (alpha grids / cv folds / bootstrap replicates are this shape): F lasso fits of the
same sparse design, different sampling seeds, one CTA per fit through sgdnet_fit_batch_sparse, E epochs each at one
lambda. Prints aggregate sample-updates/s and the algorithmic HBM rate it corresponds to.
Usage: python scripts/batch_bench.py [F] [n] [p] [E]"""
import json
import os
import sys
import time

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")     # one hardware queue per few fit streams

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sgdnet_b200 as sg
from sgdnet_b200 import api, synth

F = int(sys.argv[1]) if len(sys.argv) > 1 else 128
n = int(sys.argv[2]) if len(sys.argv) > 2 else 200_000
p = int(sys.argv[3]) if len(sys.argv) > 3 else 50_000
E = int(sys.argv[4]) if len(sys.argv) > 4 else 6
nnz = 100
lib = sg.product()
x, y = synth.binomial_sparse(n, p, nnz, seed=1002)
yc = y - y.mean()
lmax = float(np.abs(x.T @ yc).max() / n)
lam = [lmax * 0.05]
specs, keeps = [], []
for k in range(F):
    ctl, keep = api.build_control("binomial", 1, alpha=1.0, nlambda=1, lambda_min_ratio=1e-4, lambda_=lam, maxit=E,
                                  standardize=False, intercept=True, thresh=0.0, standardize_response=False, debug=False)
    ctl.tol = 0.0
    keeps.append(keep)
    specs.append(dict(train_rows=None, test_rows=None, control=ctl, rng=lib.rng_from_seed(100 + k)))
from sgdnet_b200 import _abi
m = _abi.CscMatrix.from_any(x)          # the caller's dgCMatrix: conversion from scipy's CSR is not part of the call
t0 = time.perf_counter()
raws, _ = lib.fit_batch(m, y.reshape(-1, 1), specs)
wall = time.perf_counter() - t0
updates = sum(int(r.npasses) for r in raws) * n
solver = max(r.seconds_solver for r in raws)
setup = max(r.seconds_setup for r in raws)
b_upd = 12 * nnz + 8 + 4 + 8 + 16
print(json.dumps({"workload": f"{F} concurrent lasso fits (one CTA each) of binomial sparse {n}x{p}, {nnz} nnz/row, {E} epochs each",
                  "fits": F, "updates": updates, "wall_s": wall, "solver_s": solver, "setup_s": setup, "wall_over_solver": wall / solver,
                  "agg_updates_per_s_solver": updates / solver, "agg_updates_per_s_wall": updates / wall,
                  "algorithmic_GBps_solver": updates / solver * b_upd / 1e9, "frac_of_measured_hbm_peak": updates / solver * b_upd / 1e9 / 6535.7}))
