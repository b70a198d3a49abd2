"""Fixture for score() / predict(): tests/golden/score_fixture.npz.

Independent of oracle/sgdnet_oracle.cpp and of the CUDA library: the COEFFICIENTS are outputs of the reference's own
compiled solver (tests/golden/ref_vectors.npz), the SCORES are computed by tests/r_score.py, a statement-by-statement
numpy rendering of R/score.R and R/predict.sgdnet.R. auc's tie-breaking uniforms (stats::runif inside auc(),
R/score.R:218) come from numpy's own MT19937 seeded the way R's set.seed() seeds it (tests/r_score_rng.py).

Run in the build container:  python tests/golden/make_score_fixture.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import r_score
from r_score_rng import RUnif
from ref_vectors import case

# name -> (reference-vector case that supplies x, y and the coefficients, family, measures, rows scored)
SCORE_CASES = {
    "abalone_gaussian": ("c1_abalone_gaussian_enet", "gaussian", ("deviance", "mse", "mae")),
    "heart_binomial": ("heart_binomial_lasso_sparse", "binomial", ("deviance", "mse", "mae", "class", "auc")),
    "wine_multinomial": ("wine_multinomial_enet", "multinomial", ("deviance", "mse", "mae", "class")),
    "student_mgaussian": ("student_mgaussian_grouplasso", "mgaussian", ("deviance", "mse", "mae")),
    "c2mini_binomial": ("c2mini_sparse_binomial_enet", "binomial", ("deviance", "mse", "mae", "class", "auc")),
    "c3mini_multinomial": ("c3mini_dense_multinomial", "multinomial", ("deviance", "mse", "mae", "class")),
}
AUC_SEED = 7


def score_inputs(name):
    """x, y, family, a0, beta (R shapes; multinomial intercepts centred as R's object has them) and the held-out rows:
    every third row is scored, so the row-subset path of the device code is exercised as well."""
    ref_name, family, measures = SCORE_CASES[name]
    x, y, kw, exp = case(ref_name)
    a0, beta = r_score.fit_to_r_shapes(family, exp["a0"], exp["beta"])
    if family == "multinomial":
        a0 = a0 - a0.mean(axis=0, keepdims=True)          # R/sgdnet.R:409-410
    rows = np.arange(0, x.shape[0], 3)
    return x, np.asarray(y), family, a0, beta, rows, measures


def main():
    store = {}
    for name in SCORE_CASES:
        x, y, family, a0, beta, rows, measures = score_inputs(name)
        import scipy.sparse as sp
        xs = sp.csr_matrix(x)[rows] if sp.issparse(x) else np.asarray(x)[rows]
        ys = y[rows]
        store[f"{name}/link"] = r_score.predict_link(family, a0, beta, xs)
        for m in measures:
            runif = RUnif(AUC_SEED) if m == "auc" else None
            store[f"{name}/{m}"] = r_score.score(family, a0, beta, xs, ys, m, runif=runif)
            print(name, m, store[f"{name}/{m}"][:3])
    np.savez_compressed(os.path.join(HERE, "score_fixture.npz"), **store)


if __name__ == "__main__":
    main()
