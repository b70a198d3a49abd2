"""Shared by tests/test_score_cpu.py and tests/test_parity_gpu.py: a backend's predict / score checked against
tests/golden/score_fixture.npz (R/score.R rendered in numpy, tests/r_score.py; coefficients from the reference build)."""
import os
import sys

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from make_score_fixture import AUC_SEED, SCORE_CASES, score_inputs  # noqa: E402

from sgdnet_b200._abi import FAMILIES

_FIX = None


def fixture():
    global _FIX
    if _FIX is None:
        _FIX = np.load(os.path.join(ROOT, "tests", "golden", "score_fixture.npz"))
    return _FIX


def raw_coefficients(family, a0, beta):
    """R shapes -> the C ABI's a0 (n_lambda, K), beta (n_lambda, p, K)."""
    if family in ("gaussian", "binomial"):
        return np.ascontiguousarray(a0[:, None]), np.ascontiguousarray(beta.T[:, :, None])
    return np.ascontiguousarray(a0.T), np.ascontiguousarray(np.stack([b.T for b in beta], axis=2))


def encoded_y(family, y):
    if family in ("binomial", "multinomial"):
        levels = np.unique(np.asarray(y).reshape(-1))
        return np.searchsorted(levels, np.asarray(y).reshape(-1)).astype(np.float64).reshape(-1, 1)
    return np.asarray(y, dtype=np.float64).reshape(len(y), -1)


def held_out(name):
    x, y, family, a0, beta, rows, measures = score_inputs(name)
    xs = sp.csr_matrix(x)[rows].tocsc() if sp.issparse(x) else np.ascontiguousarray(np.asarray(x)[rows])
    return xs, y[rows], family, a0, beta, measures


def check_backend_against_fixture(lib, name, rtol=1e-10):
    """link and every measure the backend implements, on the held-out rows as a matrix of their own."""
    fx = fixture()
    xs, ys, family, a0, beta, measures = held_out(name)
    a0r, br = raw_coefficients(family, a0, beta)
    link = lib.predict(xs, a0r, br)                                   # (L, K, n)
    exp_link = fx[f"{name}/link"]
    got = link[:, 0, :].T if family in ("gaussian", "binomial") else np.transpose(link, (2, 1, 0))
    scale = np.abs(exp_link).max()
    assert np.abs(got - exp_link).max() <= 1e-12 * scale, f"{name}: link"
    yv = encoded_y(family, ys)
    done = []
    for m in measures:
        if m == "deviance" and not lib.has("score_dense"):
            sc = lib.score_deviance(xs, yv, FAMILIES[family], a0r, br)
        elif lib.has("score_dense"):
            sc = lib.score(xs, yv, FAMILIES[family], m, a0r, br, rng=lib.rng_from_seed(AUC_SEED) if m == "auc" else None)
        else:
            continue
        exp = fx[f"{name}/{m}"]
        assert sc.shape == exp.shape
        assert np.abs(sc - exp).max() <= rtol * np.abs(exp).max(), f"{name}: {m}: {np.abs(sc - exp).max()}"
        done.append(m)
    return done
