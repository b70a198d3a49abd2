"""Dense solver rates on BASELINE config 3 / 4 shapes (multinomial 60000x784 K=10; mgaussian Nx2000 K=4) and config 1.
Prints GPU sample-updates/s (solver kernels only) and, on a bounded sample, the CPU oracle's."""
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
from oracle_lib import load_oracle

lib, oracle = sg.product(), load_oracle()
oracle.lib.oracle_set_arith(0) if hasattr(oracle.lib, "oracle_set_arith") else None


def run(name, x, y, cpu_rows, **kw):
    t0 = time.perf_counter()
    g = sg.sgdnet(x, y, backend=lib, **kw)
    wall = time.perf_counter() - t0
    n = x.shape[0]
    gpu = n * g.npasses / g.raw.seconds_solver
    xs, ys = x[:cpu_rows], y[:cpu_rows]
    kw2 = dict(kw); kw2["maxit"] = 2; kw2["nlambda"] = min(3, kw.get("nlambda", 3))
    r = sg.sgdnet(xs, ys, backend=oracle, **kw2)
    cpu = cpu_rows * r.npasses / max(r.raw.seconds_solver, 1e-9)
    print(json.dumps({"config": name, "n": n, "p": x.shape[1], "npasses": int(g.npasses), "gpu_updates_per_s": gpu,
                      "gpu_solver_s": g.raw.seconds_solver, "gpu_wall_s": wall, "cpu_oracle_updates_per_s": cpu,
                      "us_per_update_gpu": 1e6 / gpu}), flush=True)


x, y = synth.multinomial_dense(60000, 784, 10, seed=1003)
run("C3 multinomial 60000x784 K=10 alpha=0.8", x, y, 6000, family="multinomial", alpha=0.8, nlambda=5, maxit=5, seed=1)
x, y = synth.mgaussian_dense(50000, 2000, 4, seed=1004)
run("C4 mgaussian 50000x2000 K=4 (n reduced from 200k)", x, y, 4000, family="mgaussian", alpha=1.0, nlambda=5, maxit=5, seed=1)
d = np.load(os.path.join(ROOT, "tests", "golden", "abalone.npz"))
run("C1 abalone gaussian alpha=0.5 100 lambda", d["x"], d["y"], 4177, family="gaussian", alpha=0.5, seed=1)
