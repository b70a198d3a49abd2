"""Small dense fit on the cluster kernel (for ncu). Usage: python scripts/dense_prof_run.py [c3|c4] [n]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sgdnet_b200 as sg
from sgdnet_b200 import synth
shape = sys.argv[1] if len(sys.argv) > 1 else "c3"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 6000
if shape == "c3":
    x, y = synth.multinomial_dense(n, 784, 10, seed=1003)
    g = sg.sgdnet(x, y, backend=sg.product(), family="multinomial", alpha=0.8, nlambda=2, maxit=3, seed=1)
else:
    x, y = synth.mgaussian_dense(n, 2000, 4, seed=1004)
    g = sg.sgdnet(x, y, backend=sg.product(), family="mgaussian", alpha=1.0, nlambda=2, maxit=3, seed=1)
print(shape, g.npasses, n * g.npasses / g.raw.seconds_solver)
