"""Pins the CPU oracle with the reference's own test assertions (tests/testthat/*.R re-expressed without R):
numpy / scipy / sklearn closed forms stand in for lm(), glm() and glmnet. Runs without a GPU."""
import numpy as np
import pytest
import scipy.sparse as sp

import sgdnet_b200 as sg
from conftest import golden
from sgdnet_b200 import synth


def sd2(a, axis=0):
    return np.sqrt(((a - a.mean(axis=axis, keepdims=True)) ** 2).sum(axis=axis) / a.shape[axis])


# ---------------------------------------------------------------- R RNG (SURVEY.md 8c known-answer values)
@pytest.mark.parametrize("seed,expect", [(1, [0.26550866, 0.37212390, 0.57285336, 0.90820779]),
                                         (42, [0.91480604, 0.93707541, 0.28613953, 0.83044763]),
                                         (123, [0.28757752, 0.78830514, 0.40897692, 0.88301740])])
def test_r_set_seed_runif_kat(oracle, seed, expect):
    np.testing.assert_allclose(oracle.unif(oracle.rng_from_seed(seed), 4), expect, atol=5e-9)


def test_sampling_sequence_is_floor_n_u(oracle):
    import ctypes as C
    r1, r2 = oracle.rng_from_seed(7), oracle.rng_from_seed(7)
    out = np.zeros(1000, dtype=np.uint32)
    assert oracle.lib.oracle_draw_indices(C.byref(r1), C.c_uint32(4177), C.c_int64(1000), out.ctypes.data_as(C.c_void_p)) == 0
    u = oracle.unif(r2, 1000)
    np.testing.assert_array_equal(out, np.floor(4177 * u).astype(np.uint32))
    assert out.max() < 4177


# ---------------------------------------------------------------- test-gaussian.R
def test_gaussian_unpenalised_equals_ols(oracle):                       # :3-15 (airquality -> abalone subset)
    d = golden("abalone")
    x, y = d["x"][:500, 2:], d["y"][:500]
    fit = sg.sgdnet(x, y, lambda_=[0.0], thresh=1e-9, maxit=5000, backend=oracle)   # R's thresh there is default; tighter is stricter
    X1 = np.column_stack([np.ones(len(y)), x])
    ols = np.linalg.lstsq(X1, y, rcond=None)[0]
    np.testing.assert_allclose(np.concatenate([[fit.a0[0]], fit.beta[:, 0]]), ols, rtol=1e-3, atol=1e-3)


def test_gaussian_lambda_max_and_null_first_solution(oracle):          # :17-36
    from sklearn.datasets import load_iris
    ir = load_iris()
    x, y = ir.data[:, 1:], ir.data[:, 0]
    fit = sg.sgdnet(x, y, alpha=1.0, backend=oracle)
    xs = (x - x.mean(0)) / sd2(x)
    ys = (y - y.mean()) / sd2(y)
    assert fit.lambda_.max() == pytest.approx(np.abs(xs.T @ ys).max() * sd2(y) / len(y), rel=1.5e-8)
    assert np.all(fit.beta[:, 0] == 0)


def test_gaussian_ridge_closed_form(oracle):                           # :38-60
    rng = np.random.default_rng(3)
    n, p = 200, 3
    x = rng.normal(size=(n, p))
    y = x @ np.array([1.0, -2.0, 0.5]) + rng.normal(size=n)
    lam = 0.3
    fit = sg.sgdnet(x, y, alpha=0.0, lambda_=[lam], standardize=False, intercept=False, thresh=1e-7, maxit=10000, backend=oracle)
    # objective: 1/(2n) ||y~ - Xb||^2 * sd_y-scaling + lambda/2 ||b||^2 in glmnet's parameterisation:
    # y is standardised internally and lambda is divided by sd(y) (utils.h:175-178), coefficients scaled back
    ysd = sd2(y)
    b = np.linalg.solve(x.T @ x / n + (lam / ysd) * np.eye(p), x.T @ ((y - y.mean()) / ysd) / n) * ysd
    np.testing.assert_allclose(fit.beta[:, 0], b, rtol=1e-3, atol=1e-4)


def test_gaussian_constant_response(oracle):                           # :62-71
    x = np.random.default_rng(0).normal(size=(50, 3))
    y = np.full(50, 2.5)
    fit = sg.sgdnet(x, y, nlambda=5, backend=oracle)
    assert np.all(fit.lambda_ == 0) and np.all(fit.beta == 0)
    np.testing.assert_allclose(fit.a0, 2.5)


# ---------------------------------------------------------------- test-binomial.R
def test_binomial_unpenalised_equals_glm(oracle):                      # :3-14
    rng = np.random.default_rng(5)
    x = rng.normal(size=(300, 2))
    y = (rng.uniform(size=300) < 1 / (1 + np.exp(-(0.5 + x @ [1.0, -1.5])))).astype(float)
    fit = sg.sgdnet(x, y, family="binomial", lambda_=[0.0], thresh=1e-9, maxit=10000, backend=oracle)
    X1 = np.column_stack([np.ones(300), x])
    b = np.zeros(3)
    for _ in range(50):                                               # IRLS = glm()
        mu = 1 / (1 + np.exp(-X1 @ b))
        b = b + np.linalg.solve(X1.T @ (X1 * (mu * (1 - mu))[:, None]), X1.T @ (y - mu))
    np.testing.assert_allclose(np.concatenate([[fit.a0[0]], fit.beta[:, 0]]), b, rtol=1e-5, atol=1e-5)


# ---------------------------------------------------------------- test-multinomial.R / test-mgaussian.R
def test_multinomial_probabilities_sum_to_one(oracle):                 # test-multinomial.R:8-13
    d = golden("wine")
    fit = sg.sgdnet(d["x"], d["y"], family="multinomial", nlambda=10, backend=oracle)
    pr = sg.predict(fit, d["x"], type="response", backend=oracle)
    np.testing.assert_allclose(pr.sum(axis=1), 1.0, atol=1e-12)


def test_mgaussian_ridge_closed_form(oracle):                          # test-mgaussian.R:3-28
    rng = np.random.default_rng(8)
    n, p, K = 300, 3, 2
    x = rng.normal(size=(n, p))
    y = x @ rng.normal(size=(p, K)) + rng.normal(size=(n, K))
    lam = 0.1
    fit = sg.sgdnet(x, y, family="mgaussian", alpha=0.0, lambda_=[lam], standardize=False, intercept=False,
                    thresh=1e-8, maxit=10000, backend=oracle)
    # with intercept = FALSE the mgaussian intercept stays fixed at mean(y) (families.h:380-385, quirk Q6)
    b = np.linalg.solve(x.T @ x / n + lam * np.eye(p), x.T @ (y - y.mean(0)) / n)
    got = np.column_stack([fit.beta[k][:, 0] for k in range(K)])
    np.testing.assert_allclose(got, b, rtol=1e-5, atol=1e-6)


# ---------------------------------------------------------------- test-lambda-path.R
def manual_lambda_max(x, y, family, standardize, alpha):               # :52-97
    x2 = (x - x.mean(0)) / sd2(x) if standardize else x
    if family == "binomial":
        y2 = (np.unique(y, return_inverse=True)[1]).astype(float).reshape(-1, 1)
    elif family == "multinomial":
        codes = np.unique(y, return_inverse=True)[1]
        y2 = np.eye(codes.max() + 1)[codes]
    else:
        y2 = np.asarray(y, float).reshape(len(y), -1)
    ys = sd2(y2)
    y3 = (y2 - y2.mean(0)) / ys
    ip = (x2.T @ y3) * ys
    if family == "multinomial":
        return np.abs(ip).max() / (len(y) * max(alpha, 1e-3))
    return np.sqrt((ip ** 2).sum(axis=1)).max() / (len(y) * max(alpha, 1e-3))


@pytest.mark.parametrize("family", ["gaussian", "binomial", "multinomial", "mgaussian"])
@pytest.mark.parametrize("intercept", [True, False])
@pytest.mark.parametrize("alpha", [0.0, 0.5, 1.0])
@pytest.mark.parametrize("standardize", [True, False])
def test_lambda_max_manual_formula(oracle, family, intercept, alpha, standardize):      # :99-129
    rng = np.random.default_rng(1)
    n = 64
    x = np.column_stack([rng.integers(4, 9, n), rng.normal(200, 100, n), rng.normal(150, 60, n), rng.integers(0, 2, n)]).astype(float)
    y = {"gaussian": rng.normal(20, 6, n), "binomial": rng.integers(0, 2, n), "multinomial": rng.integers(3, 6, n),
         "mgaussian": np.column_stack([rng.normal(150, 60, n), rng.normal(3.5, 0.5, n)])}[family]
    fit = sg.sgdnet(x, y, family=family, intercept=intercept, alpha=alpha, standardize=standardize, nlambda=4, maxit=2,
                    backend=oracle)
    assert fit.lambda_.max() == pytest.approx(manual_lambda_max(x, y, family, standardize, alpha), rel=1.5e-8)


def test_lambda_path_is_logspaced_like_glmnet(oracle):                 # :3-47 (glmnet's path = log-linear from lambda_max)
    d = golden("abalone")
    fit = sg.sgdnet(d["x"], d["y"], nlambda=20, lambda_min_ratio=1e-3, maxit=1, backend=oracle)
    np.testing.assert_allclose(np.diff(np.log(fit.lambda_)), np.log(1e-3) / 19, rtol=1e-12)


@pytest.mark.parametrize("family", ["gaussian", "mgaussian"])
def test_first_lasso_solution_is_sparse(oracle, family):               # :131-170
    x, y = synth.random_data(200, 5, family, True, density=1.0, seed=4)
    fit = sg.sgdnet(x.toarray(), y, family=family, alpha=1.0, backend=oracle, nlambda=5)
    first = fit.beta[:, 0] if family == "gaussian" else np.column_stack([b[:, 0] for b in fit.beta])
    assert np.all(np.abs(first) < 1e-5)


def test_same_seed_same_lambda_same_fit(oracle):                       # :173-198
    x, y = synth.random_data(150, 4, "binomial", True, density=1.0, seed=9)
    f1 = sg.sgdnet(x.toarray(), y, family="binomial", seed=2, backend=oracle, nlambda=10)
    f2 = sg.sgdnet(x.toarray(), y, family="binomial", seed=2, lambda_=f1.lambda_, backend=oracle)
    np.testing.assert_array_equal(f1.beta, f2.beta)
    assert f1.npasses == f2.npasses


# ---------------------------------------------------------------- test-sparse.R
@pytest.mark.parametrize("family", ["gaussian", "binomial"])
@pytest.mark.parametrize("intercept", [True, False])
@pytest.mark.parametrize("alpha", [0.0, 0.5, 1.0])
@pytest.mark.parametrize("standardize", [True, False])
def test_sparse_and_dense_solvers_agree(oracle, family, intercept, alpha, standardize):  # :3-35
    x, y = synth.random_data(1000, 2, family, intercept, density=0.5, seed=1)
    kw = dict(family=family, intercept=intercept, alpha=alpha, standardize=standardize, nlambda=5, thresh=1e-6, seed=1,
              backend=oracle)
    fs = sg.sgdnet(x, y, **kw)
    fd = sg.sgdnet(x.toarray(), y, **kw)
    np.testing.assert_allclose(fs.beta, fd.beta, atol=1e-3 * max(1.0, np.abs(fd.beta).max()))
    np.testing.assert_allclose(fs.a0, fd.a0, atol=1e-3 * max(1.0, np.abs(fd.a0).max()))


# ---------------------------------------------------------------- test-deviance.R
@pytest.mark.parametrize("family", ["gaussian", "binomial", "multinomial", "mgaussian"])
@pytest.mark.parametrize("intercept", [True, False])
def test_null_deviance_manual(oracle, family, intercept):              # :8-96
    rng = np.random.default_rng(1)
    n = 100
    x = rng.normal(size=(n, 2))
    y = {"gaussian": rng.normal(10, 2, n), "binomial": (rng.uniform(size=n) < 0.8).astype(float),
         "multinomial": rng.binomial(2, 0.5, n), "mgaussian": np.column_stack([rng.normal(100, 1, n), rng.normal(size=n)])}[family]
    fit = sg.sgdnet(x, y, family=family, intercept=intercept, lambda_=[1.0 / n], thresh=0.1, backend=oracle)
    if family == "gaussian":
        expect = ((y - y.mean()) ** 2).sum()
    elif family == "mgaussian":
        expect = ((y - y.mean(0)) ** 2).sum()
    elif family == "binomial":
        pbar = np.clip(y.mean(), 1e-9, 1 - 1e-9)
        pl = np.log(pbar / (1 - pbar)) if intercept else 0.0
        expect = -2 * (y * pl - np.log(1 + np.exp(pl))).sum()
    else:
        nc = len(np.unique(y))
        pred = np.bincount(y) / n if intercept else np.full(nc, 1.0 / nc)
        pred2 = np.log(pred) - np.log(pred).sum() / nc
        expect = 2 * sum(np.log(np.exp(pred2).sum()) - pred2[c] for c in y)
    assert fit.nulldev == pytest.approx(expect, rel=1e-12)
    np.testing.assert_allclose(sg.deviance(fit), (1 - fit.dev_ratio) * fit.nulldev)


# ---------------------------------------------------------------- test-cross-validation.R / test-predictions.R
@pytest.mark.parametrize("family", ["gaussian", "binomial", "multinomial", "mgaussian"])
def test_cv_runs_for_all_families(oracle, family):                     # test-cross-validation.R:3-46
    x, y = synth.random_data(120, 3, family, True, density=1.0, seed=6)
    cv = sg.cv_sgdnet(x.toarray(), y, family=family, alpha=[0.5, 1.0], nfolds=3, nlambda=6, maxit=50, backend=oracle,
                      batched=False)
    assert len(cv.cv_raw) == 2 and cv.cv_raw[0].shape == (3, 6)
    assert np.isfinite(cv.cv_summary).all()
    assert cv.lambda_1se >= cv.lambda_min


def test_cv_trains_on_the_single_fold(oracle):                         # R/cv_sgdnet.R:182-183 (quirk Q11)
    plan = sg.cv_plan(10, [1.0], np.array([1, 1, 2, 2, 2, 3, 3, 3, 3, 3]))
    assert [len(w["train_rows"]) for w in plan] == [2, 3, 5]
    assert [len(w["test_rows"]) for w in plan] == [8, 7, 5]


def test_foldid_is_cut_of_permutation():                               # R/cv_sgdnet.R:169
    f = sg.make_foldid(10, 3, np.arange(1, 11))
    np.testing.assert_array_equal(f, [1, 1, 1, 1, 2, 2, 2, 3, 3, 3])   # as.numeric(cut(1:10, 3))


def test_gaussian_link_equals_manual_xb(oracle):                       # test-predictions.R
    d = golden("abalone")
    x, y = d["x"][:300], d["y"][:300]
    fit = sg.sgdnet(x, y, nlambda=8, backend=oracle)
    link = sg.predict(fit, x, backend=oracle)
    np.testing.assert_allclose(link, fit.a0[None, :] + x @ fit.beta, rtol=1e-12, atol=1e-12)
    assert np.array_equal(link, sg.predict(fit, x, type="response", backend=oracle))


def test_interpolated_coefficients_between_path_points(oracle):
    d = golden("abalone")
    fit = sg.sgdnet(d["x"][:300], d["y"][:300], nlambda=8, backend=oracle)
    s = 0.5 * (fit.lambda_[2] + fit.lambda_[3])
    c = sg.coef(fit, s=[s])
    lo, hi = np.minimum(fit.beta[:, 2], fit.beta[:, 3]), np.maximum(fit.beta[:, 2], fit.beta[:, 3])
    assert np.all(c[1:, 0] >= lo - 1e-12) and np.all(c[1:, 0] <= hi + 1e-12)


# ---------------------------------------------------------------- test-assertions.R
def test_front_end_assertions(oracle):
    x = np.random.default_rng(0).normal(size=(20, 2))
    y = np.arange(20.0)
    with pytest.raises(ValueError, match="must match"):
        sg.sgdnet(x, y[:-1], backend=oracle)
    with pytest.raises(ValueError, match=r"alpha\) must be in"):
        sg.sgdnet(x, y, alpha=1.5, backend=oracle)
    with pytest.raises(ValueError, match="must be positive"):
        sg.sgdnet(x, y, lambda_=[-1.0], backend=oracle)
    with pytest.raises(ValueError, match="NA values"):
        sg.sgdnet(x, np.where(y == 3, np.nan, y), backend=oracle)
    with pytest.raises(ValueError, match="cannot be negative"):
        sg.sgdnet(x, y, thresh=-1, backend=oracle)
    with pytest.raises(ValueError, match="negative or zero"):
        sg.sgdnet(x, y, maxit=0, backend=oracle)
    with pytest.raises(ValueError, match="one-dimensional"):
        sg.sgdnet(x, np.column_stack([y, y]), backend=oracle)
    with pytest.raises(ValueError, match="more than two classes"):
        sg.sgdnet(x, (y % 3), family="binomial", backend=oracle)
    with pytest.raises(ValueError, match="only two classes"):
        sg.sgdnet(x, (y % 2), family="multinomial", backend=oracle)
    with pytest.raises(ValueError, match="must not be one-dimensional"):
        sg.sgdnet(x, y, family="mgaussian", backend=oracle)
    with pytest.raises(ValueError, match="more folds than samples"):
        sg.cv_sgdnet(x, y, nfolds=21, backend=oracle)


# ---------------------------------------------------------------- test-options.R
def test_debug_losses_positive_and_finite(oracle):
    x, y = synth.random_data(100, 3, "binomial", True, density=1.0, seed=2)
    fit = sg.sgdnet(x.toarray(), y, family="binomial", nlambda=4, debug=True, backend=oracle)
    assert len(fit.diagnostics["loss"]) == 4
    for l, e in zip(fit.diagnostics["loss"], fit.epochs):
        assert len(l) == e and np.all(np.isfinite(l)) and np.all(l > 0)


# ---------------------------------------------------------------- arithmetic modes of the oracle itself
@pytest.mark.parametrize("family,sparse", [("gaussian", False), ("binomial", True), ("multinomial", False)])
def test_portable_and_libm_arithmetic_agree(oracle, family, sparse):
    """The fixed summation order / sgd_exp of include/sgdnet_arith.h only moves results at the ulp level: with the
    path lengths pinned (thresh = 0, exactly `maxit` epochs per lambda, hence the same sampling sequence) the two
    arithmetic modes agree ~1e-12, far inside the 1e-6 bar.

    With a convergence threshold the epoch counts themselves can differ between the modes on binomial / multinomial
    paths (first lambda: coefficients decay geometrically until the soft threshold zeroes them; the epoch where that
    happens is decided at the 1e-16 level) - which is exactly why the GPU library pins its arithmetic."""
    x, y = synth.random_data(300, 6, family, True, density=0.5 if sparse else 1.0, seed=12)
    xx = x if sparse else x.toarray()
    kw = dict(family=family, alpha=0.6, nlambda=6, lambda_min_ratio=0.05, thresh=0.0, maxit=12, standardize=not sparse,
              seed=3, backend=oracle)
    try:
        oracle.lib.oracle_set_arith(1)
        a = sg.sgdnet(xx, y, **kw)
        oracle.lib.oracle_set_arith(0)
        b = sg.sgdnet(xx, y, **kw)
    finally:
        oracle.lib.oracle_set_arith(1)
    np.testing.assert_array_equal(a.epochs, b.epochs)
    assert np.abs(a.raw.beta - b.raw.beta).max() <= 1e-9 * max(np.abs(b.raw.beta).max(), 1e-300)
    np.testing.assert_array_equal(a.raw.beta != 0, b.raw.beta != 0)
