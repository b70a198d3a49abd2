"""Small config-3-shaped dense multinomial fit (for ncu)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sgdnet_b200 as sg
from sgdnet_b200 import synth
x, y = synth.multinomial_dense(6000, 784, 10, seed=1003)
g = sg.sgdnet(x, y, backend=sg.product(), family="multinomial", alpha=0.8, nlambda=2, maxit=3, seed=1)
print(g.npasses, 6000 * g.npasses / g.raw.seconds_solver)
