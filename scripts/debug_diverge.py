import sys, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
import sgdnet_b200 as sg
from sgdnet_b200 import synth
from oracle_lib import load_oracle
cuda, oracle = sg.product(), load_oracle()
x, y = synth.random_data(300, 5, "binomial", True, density=0.8, seed=11)
x = x.toarray()
full = sg.sgdnet(x, y, family="binomial", alpha=0.5, standardize=False, nlambda=12, thresh=1e-4, maxit=300, seed=3, backend=oracle)
lam = [full.lambda_[0]]
print("lambda0", lam, "epochs", full.epochs)
for maxit in (1, 2, 3, 5, 10, 20, 50, 100, 170):
    kw = dict(family="binomial", alpha=0.5, standardize=False, lambda_=lam, thresh=0.0, maxit=maxit, seed=3)
    g = sg.sgdnet(x, y, backend=cuda, **kw); r = sg.sgdnet(x, y, backend=oracle, **kw)
    db = np.abs(g.raw.beta - r.raw.beta).max(); sb = np.abs(r.raw.beta).max()
    print(maxit, "beta diff", db, "scale", sb, "a0", g.raw.a0.ravel(), r.raw.a0.ravel(), "epochs", g.epochs, r.epochs)
    print("   gpu", g.raw.beta.ravel()); print("   ref", r.raw.beta.ravel())
