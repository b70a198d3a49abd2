"""BASELINE.json's configurations at their FULL sizes, each for a bounded number of epochs: the CUDA path against the
oracle on the same inputs and sampling sequence (bit-identical coefficients, equal supports and epoch counts), plus
the size-independent properties the reference's own tests assert (first lasso solution all-zero at lambda_max,
test-lambda-path.R:131-170; same seed => same fit, :173-198; probabilities sum to one, test-multinomial.R:8-13)."""
import numpy as np
import pytest

import sgdnet_b200 as sg
from parity import assert_fit_parity
import synth

pytestmark = pytest.mark.gpu


def _binomial_lambda_max(x, y):
    """max |x^T (y - ybar)/sd_y| * sd_y / n  (families.h:203-220; standardize = FALSE)."""
    yc = y - y.mean()
    return float(np.abs(x.T @ yc).max() / x.shape[0])


def test_config2_sparse_binomial_lasso_1m_x_100k(cuda, oracle):
    x, y = synth.binomial_sparse(1_000_000, 100_000, 100, seed=1002)
    lmax = _binomial_lambda_max(x, y)
    # lambda = 4 exceeds every per-sample gradient component (|x| <= 1, |g_change| < 2): the prox keeps that solution
    # at exactly zero however few epochs are run; the next two sit inside the automatic path
    kw = dict(family="binomial", alpha=1.0, standardize=False, intercept=True, lambda_=[4.0, lmax * 0.5, lmax * 0.1],
              maxit=2, thresh=1e-3, seed=1)
    g = sg.sgdnet(x, y, backend=cuda, **kw)
    assert not np.any(g.raw.beta[0]), "a penalty above every gradient component must leave the solution all-zero"
    assert np.count_nonzero(g.raw.beta[2]) > 0
    g2 = sg.sgdnet(x, y, backend=cuda, **kw)                      # the overlapped schedule is deterministic
    np.testing.assert_array_equal(g.raw.beta, g2.raw.beta)
    np.testing.assert_array_equal(g.raw.a0, g2.raw.a0)
    r = sg.sgdnet(x, y, backend=oracle, **kw)
    assert_fit_parity(g.raw, r.raw)


def test_config2_full_100_lambda_path_matches_the_reference_build(cuda):
    """BASELINE config 2 END TO END through the plugin call sgdnet_fit_sparse (host CSC in, archives out): the whole
    100-lambda path against what the reference's own compiled code produced on the same inputs and seed
    (tests/golden/c2_full_path_ref.json; 2302 s on one CPU core): lambda path exact, 381 epochs lambda by lambda,
    the same number of nonzero coefficients at every lambda, deviances within 1e-6 relative. Also BASELINE metric (ii),
    the lambda-path fit time, written to gpurun_out/c2_full_path_gpu.json."""
    import json
    import os
    import time
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with open(os.path.join(root, "tests", "golden", "c2_full_path_ref.json")) as fh:
        ref = json.load(fh)
    x, y = synth.binomial_sparse(1_000_000, 100_000, 100, seed=1002)
    t0 = time.perf_counter()
    g = sg.sgdnet(x, y, family="binomial", alpha=1.0, standardize=False, intercept=True, nlambda=100, thresh=1e-3,
                  maxit=1000, seed=1, backend=cuda)
    wall = time.perf_counter() - t0
    np.testing.assert_array_equal(g.raw.lambda_, np.array(ref["lambda"]), err_msg="lambda path")
    np.testing.assert_array_equal(g.raw.epochs, np.array(ref["epochs_per_lambda"]), err_msg="epochs per lambda")
    assert g.npasses == ref["npasses"] == 381
    assert not g.raw.return_codes.any()
    nnz = [int(np.count_nonzero(g.raw.beta[l])) for l in range(100)]
    assert nnz == ref["nonzeros_per_lambda"], "number of nonzero coefficients per lambda"
    dev_g, dev_r = 1.0 - g.raw.dev_ratio, 1.0 - np.array(ref["dev_ratio"])
    assert np.max(np.abs(dev_g - dev_r)) <= 1e-6 * np.max(np.abs(dev_r))
    out = {"workload": ref["what"], "call": "sgdnet_fit_sparse (host buffers)", "wall_s": wall, "npasses": int(g.npasses),
           "updates_per_s": 1_000_000 * int(g.npasses) / wall, "seconds_setup": g.raw.seconds_setup,
           "seconds_solver": g.raw.seconds_solver, "seconds_deviance": g.raw.seconds_deviance,
           "kernel_launches": int(g.raw.kernel_launches),
           "max_rel_deviance_diff_vs_reference_build": float(np.max(np.abs(dev_g - dev_r)) / np.max(np.abs(dev_r)))}
    os.makedirs(os.path.join(root, "gpurun_out"), exist_ok=True)
    with open(os.path.join(root, "gpurun_out", "c2_full_path_gpu.json"), "w") as fh:
        json.dump(out, fh)


def test_config5_shape_cv_folds_500k_x_50k(cuda, oracle):
    """One alpha of the 10-fold grid at config 5's size: 10 fold fits (50k rows each) + the full fit in one batch;
    two of the fold fits and their held-out deviances are checked against the oracle."""
    x, y = synth.binomial_sparse(500_000, 50_000, 50, seed=1005)
    foldid = (np.random.Generator(np.random.PCG64(1005)).permutation(500_000) % 10) + 1
    kw = dict(family="binomial", alpha=[0.5], foldid=foldid, nlambda=4, standardize=False, maxit=3, seed=1000)
    g = sg.cv_sgdnet(x, y, backend=cuda, **kw)
    assert g.cv_raw[0].shape == (10, 4) and np.isfinite(g.cv_raw[0]).all()
    xr = x.tocsr()
    for k in (0, 7):
        tr, te = np.nonzero(foldid == k + 1)[0], np.nonzero(foldid != k + 1)[0]
        r = sg.sgdnet(xr[tr].tocsc(), y[tr], family="binomial", alpha=0.5, lambda_=g.lambda_[0], standardize=False, maxit=3,
                      seed=1000 + 1 + k, backend=oracle)
        assert_fit_parity(g.fold_fits[k].raw, r.raw)
        sc = sg.score(r, xr[te].tocsc(), y[te], "deviance", backend=oracle)
        np.testing.assert_allclose(g.cv_raw[0][k], sc, rtol=1e-9)


def test_config3_dense_multinomial_60000_x_784(cuda, oracle):
    x, y = synth.multinomial_dense(60_000, 784, 10, seed=1003)
    kw = dict(family="multinomial", alpha=0.8, nlambda=2, maxit=1, seed=1)
    g = sg.sgdnet(x, y, backend=cuda, **kw)
    r = sg.sgdnet(x, y, backend=oracle, **kw)
    assert_fit_parity(g.raw, r.raw)
    pr = sg.predict(g, x[:2000], type="response", backend=cuda)
    np.testing.assert_allclose(np.asarray(pr).sum(axis=1), 1.0, rtol=0, atol=1e-12)


def test_config4_dense_mgaussian_200000_x_2000(cuda, oracle):
    x, y = synth.mgaussian_dense(200_000, 2000, 4, seed=1004)
    kw = dict(family="mgaussian", alpha=1.0, nlambda=2, maxit=1, seed=1)      # group lasso
    g = sg.sgdnet(x, y, backend=cuda, **kw)
    r = sg.sgdnet(x, y, backend=oracle, **kw)
    assert_fit_parity(g.raw, r.raw)
