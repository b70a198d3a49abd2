import sys, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
import sgdnet_b200 as sg
from sgdnet_b200 import synth
from oracle_lib import load_oracle
cuda, oracle = sg.product(), load_oracle()
for (n, p, nnz) in ((40, 30, 4), (100, 50, 5), (270, 18, 13), (2000, 400, 12)):
    x, y = synth.binomial_sparse(n, p, nnz, seed=21)
    for alpha in (1.0, 0.3, 0.0):
        for maxit in (1, 2, 5):
            kw = dict(family="binomial", alpha=alpha, standardize=False, nlambda=3, lambda_min_ratio=0.1, thresh=0.0, maxit=maxit, seed=6)
            g = sg.sgdnet(x, y, backend=cuda, **kw); r = sg.sgdnet(x, y, backend=oracle, **kw)
            db = np.abs(g.raw.beta - r.raw.beta).max(axis=(1, 2)); sb = np.abs(r.raw.beta).max(axis=(1, 2))
            print(f"n={n} p={p} alpha={alpha} maxit={maxit} diff/lambda={db} scale={sb} a0diff={np.abs(g.raw.a0-r.raw.a0).max():.2e}")
