"""GPU parity tests proper: every call goes through the C ABI of libsgdnet_b200.so and is compared with the CPU
oracle on the same seeded inputs and the same R-compatible sampling sequence."""
import ctypes as C

import numpy as np
import pytest
import scipy.sparse as sp

import sgdnet_b200 as sg
from conftest import golden
from parity import assert_fit_parity, rel_close
import synth

pytestmark = pytest.mark.gpu


def both(cuda, oracle, x, y, **kw):
    g = sg.sgdnet(x, y, backend=cuda, **kw)
    r = sg.sgdnet(x, y, backend=oracle, **kw)
    return g, r


def test_c1_abalone_gaussian_elasticnet(cuda, oracle):
    """BASELINE config 1: gaussian, alpha = 0.5, 100-lambda path on the bundled abalone data, set.seed(1)."""
    d = golden("abalone")
    g, r = both(cuda, oracle, d["x"], d["y"], family="gaussian", alpha=0.5, seed=1)
    assert r.npasses == 1143            # pinned by the survey's independent restatement
    assert_fit_parity(g.raw, r.raw)


@pytest.mark.parametrize("alpha", [0.0, 0.5, 1.0])
@pytest.mark.parametrize("intercept", [True, False])
@pytest.mark.parametrize("standardize", [True, False])
def test_dense_binomial_grid(cuda, oracle, alpha, intercept, standardize):
    x, y = synth.random_data(300, 5, "binomial", intercept, density=0.8, seed=11)
    g, r = both(cuda, oracle, x.toarray(), y, family="binomial", alpha=alpha, intercept=intercept,
                standardize=standardize, nlambda=12, thresh=1e-4, maxit=300, seed=3)
    assert_fit_parity(g.raw, r.raw)


@pytest.mark.parametrize("alpha", [0.0, 0.8, 1.0])
def test_dense_multinomial_wine(cuda, oracle, alpha):
    d = golden("wine")
    g, r = both(cuda, oracle, d["x"], d["y"], family="multinomial", alpha=alpha, nlambda=15, seed=2)
    assert_fit_parity(g.raw, r.raw)


@pytest.mark.parametrize("alpha", [0.0, 1.0])
@pytest.mark.parametrize("standardize_response", [False, True])
def test_dense_mgaussian_student(cuda, oracle, alpha, standardize_response):
    d = golden("student")
    g, r = both(cuda, oracle, d["x"], d["y"], family="mgaussian", alpha=alpha, nlambda=15,
                standardize_response=standardize_response, seed=4)
    assert_fit_parity(g.raw, r.raw)


@pytest.mark.parametrize("family,K,alpha", [("multinomial", 4, 0.8), ("mgaussian", 3, 1.0), ("binomial", 1, 0.5), ("gaussian", 1, 0.0),
                                            ("multinomial", 10, 0.0)])
def test_dense_wide_designs_on_the_cluster_kernel(cuda, oracle, family, K, alpha):
    """p >= 512: the thread-block-cluster kernel (saga_dense_cluster.cu) and the eight-block dot product of the
    arithmetic specification; p not a multiple of 256 or 2048, more than one slice per CTA for p = 2500."""
    rng = np.random.default_rng(17)
    for n, p in ((300, 530), (200, 2500)):
        x = rng.normal(size=(n, p)) * (rng.uniform(size=(n, p)) < 0.3)
        if family == "multinomial":
            y = np.argmax(x[:, :K] + rng.gumbel(size=(n, K)), axis=1)
        elif family == "mgaussian":
            y = x[:, :5] @ rng.normal(size=(5, K)) + 0.5 * rng.normal(size=(n, K))
        elif family == "binomial":
            y = (x[:, 0] - x[:, 1] + rng.normal(size=n) > 0).astype(float)
        else:
            y = x[:, :4] @ np.array([1.0, -2.0, 0.5, 3.0]) + rng.normal(size=n)
        g, r = both(cuda, oracle, x, y, family=family, alpha=alpha, nlambda=5, thresh=1e-3, maxit=25, seed=21)
        assert_fit_parity(g.raw, r.raw)


@pytest.mark.gpu
@pytest.mark.parametrize("family,K,p,alpha", [("multinomial", 20, 600, 0.7),      # 32-class bucket, one slice: state in shared memory
                                              ("mgaussian", 8, 5000, 1.0),        # 8-class bucket, four slices: state in shared memory
                                              ("multinomial", 6, 3000, 0.5),      # 8-class bucket, two slices: state in registers
                                              ("gaussian", 1, 7000, 0.3),         # scalar, four slices: state in registers
                                              ("binomial", 1, 2048, 1.0),         # exactly one full slice per CTA
                                              ("mgaussian", 2, 9000, 0.5),        # five slices: the any-shape cluster kernel
                                              ("multinomial", 12, 4500, 0.0)])    # state too large for shared memory: the any-shape kernel
def test_dense_cluster_kernel_instantiations(cuda, oracle, family, K, p, alpha):
    """Every state placement / slice count of saga_dense_cluster.cu and the fallback to saga_dense_cluster_generic.cu."""
    rng = np.random.default_rng(23)
    n = 120
    x = rng.normal(size=(n, p)) * (rng.uniform(size=(n, p)) < 0.2)
    if family == "multinomial":
        y = np.argmax(x[:, :K] + rng.gumbel(size=(n, K)), axis=1)
    elif family == "mgaussian":
        y = x[:, :5] @ rng.normal(size=(5, K)) + 0.5 * rng.normal(size=(n, K))
    elif family == "binomial":
        y = (x[:, 0] - x[:, 1] + rng.normal(size=n) > 0).astype(float)
    else:
        y = x[:, :4] @ np.array([1.0, -2.0, 0.5, 3.0]) + rng.normal(size=n)
    g, r = both(cuda, oracle, x, y, family=family, alpha=alpha, nlambda=4, thresh=1e-3, maxit=12, seed=5)
    assert_fit_parity(g.raw, r.raw)


def _heart():
    d = golden("heart")
    n, p = (int(v) for v in d["x_shape"])
    return sp.csc_matrix((d["x_x"], d["x_i"], d["x_p"]), shape=(n, p)), d["y"]


@pytest.mark.parametrize("alpha", [0.0, 0.5, 1.0])
@pytest.mark.parametrize("standardize", [False, True])
def test_sparse_binomial_heart(cuda, oracle, alpha, standardize):
    x, y = _heart()
    g, r = both(cuda, oracle, x, y, family="binomial", alpha=alpha, standardize=standardize, nlambda=12,
                maxit=200, seed=5)
    assert_fit_parity(g.raw, r.raw)


@pytest.mark.parametrize("family", ["gaussian", "binomial"])
@pytest.mark.parametrize("alpha", [0.0, 0.3, 1.0])
@pytest.mark.parametrize("intercept", [True, False])
def test_sparse_k1_synthetic(cuda, oracle, family, alpha, intercept):
    """Genuinely sparse rows (the lagged path the reference's own tests never reach)."""
    x, yb = synth.binomial_sparse(2000, 400, 12, seed=21)
    y = yb if family == "binomial" else (x @ np.linspace(-1, 1, 400) + 0.1 * yb)
    g, r = both(cuda, oracle, x, y, family=family, alpha=alpha, intercept=intercept, standardize=False,
                nlambda=10, thresh=1e-4, maxit=100, seed=6)
    assert_fit_parity(g.raw, r.raw)


@pytest.mark.parametrize("family,alpha,intercept", [("binomial", 1.0, True), ("binomial", 0.4, True), ("gaussian", 0.0, True),
                                                    ("gaussian", 1.0, False), ("binomial", 0.0, False)])
def test_sparse_wavefront_stress(cuda, oracle, family, alpha, intercept):
    """Many rows in flight with frequent feature conflicts (10 % per row pair), forwarding chains and, for alpha < 1,
    wscale resets inside epochs: the overlapped schedule must leave every coefficient bit-identical."""
    x, yb = synth.binomial_sparse(8000, 1000, 10, seed=41)
    y = yb if family == "binomial" else (x @ np.linspace(-2, 2, 1000) + 0.3 * yb)
    g, r = both(cuda, oracle, x, y, family=family, alpha=alpha, intercept=intercept, standardize=False,
                nlambda=6, thresh=1e-3, maxit=40, seed=12)
    assert_fit_parity(g.raw, r.raw)


@pytest.mark.parametrize("alpha", [1.0, 0.5])
def test_sparse_wavefront_dense_conflicts_and_repeated_samples(cuda, oracle, alpha):
    """Tiny p and n: almost every row shares features with its predecessors and the same sample recurs inside the
    window of rows in flight (gradient-memory hazard)."""
    x, y = synth.binomial_sparse(120, 48, 8, seed=43)
    g, r = both(cuda, oracle, x, y, family="binomial", alpha=alpha, standardize=False, nlambda=8, thresh=1e-5,
                maxit=200, seed=13)
    assert_fit_parity(g.raw, r.raw)


def test_sparse_long_and_empty_rows(cuda, oracle):
    """Rows longer than a ring slot (> 128 nonzeros) and empty rows take the in-place path."""
    rng = np.random.Generator(np.random.PCG64(5))
    x = sp.random(400, 600, density=0.3, random_state=7, format="lil")
    x[3, :] = 0
    x[17, :] = 0
    x[5, :] = rng.uniform(0.1, 1.0, size=600)     # one fully dense row
    x = sp.csc_matrix(x)
    y = (rng.uniform(size=400) < 0.5).astype(float)
    g, r = both(cuda, oracle, x, y, family="binomial", alpha=0.9, standardize=False, nlambda=6, maxit=50, seed=8)
    assert_fit_parity(g.raw, r.raw)


@pytest.mark.parametrize("family", ["multinomial", "mgaussian"])
@pytest.mark.parametrize("standardize", [False, True])
def test_sparse_generic_multiclass(cuda, oracle, family, standardize):
    x, y = synth.random_data(250, 6, family, True, density=0.4, seed=13)
    g, r = both(cuda, oracle, x, y, family=family, alpha=0.7, standardize=standardize, nlambda=8, maxit=150, seed=9)
    assert_fit_parity(g.raw, r.raw)


def test_user_lambda_order_and_debug_losses(cuda, oracle):
    d = golden("abalone")
    lam = [0.5, 2.0, 0.01]           # unsorted: used in the given order (quirk Q14)
    g, r = both(cuda, oracle, d["x"][:800], d["y"][:800], family="gaussian", alpha=0.3, lambda_=lam, seed=10,
                debug=True)
    assert_fit_parity(g.raw, r.raw)
    for lg, lr in zip(g.diagnostics["loss"], r.diagnostics["loss"]):
        rel_close(lg, lr, what="debug losses")


@pytest.mark.parametrize("nnz_row", [12, 150])
def test_sparse_debug_losses_and_deviance_through_the_tile_pass(cuda, oracle, nnz_row):
    """The bulk-copy tile form of the loss pass (sparse, K = 1): per-epoch debug losses (mode 1) and per-lambda deviances
    (mode 0, with the nonzero bitmap) against the oracle; 12 nonzeros per row stages every tile, 150 per row overflows
    the stage (the tiles are read in place); n is not a multiple of the 32-row tile and some rows are empty."""
    rng = np.random.default_rng(31)
    n, p = 1000 + 13, 900
    x = sp.random(n, p, density=nnz_row / p, random_state=11, format="lil")
    x[7, :] = 0
    x[n - 1, :] = 0
    x = sp.csc_matrix(x)
    y = (rng.uniform(size=n) < 0.4).astype(float)
    g, r = both(cuda, oracle, x, y, family="binomial", alpha=1.0, standardize=False, nlambda=5, maxit=6, thresh=0.0, seed=12, debug=True)
    assert_fit_parity(g.raw, r.raw)
    assert len(g.diagnostics["loss"]) == len(r.diagnostics["loss"]) > 0
    for lg, lr in zip(g.diagnostics["loss"], r.diagnostics["loss"]):
        rel_close(lg, lr, what="debug losses")


def test_rng_stream_is_advanced_exactly(cuda, oracle):
    """The generator handed back has consumed n * npasses draws: a second fit continues the same stream."""
    x, y = synth.random_data(200, 4, "gaussian", True, density=1.0, seed=3)
    x = x.toarray()
    rg, ro = cuda.rng_from_seed(99), oracle.rng_from_seed(99)
    g1 = sg.sgdnet(x, y, nlambda=5, rng=rg, backend=cuda)
    r1 = sg.sgdnet(x, y, nlambda=5, rng=ro, backend=oracle)
    assert g1.npasses == r1.npasses
    np.testing.assert_array_equal(cuda.unif(rg, 5), oracle.unif(ro, 5))


def test_rng_stream_is_advanced_exactly_long_warm_path(cuda, oracle):
    """100 lambdas on a small problem with a loose threshold: 16 epochs are staged per launch and most warm-started
    lambdas use one or two of them, so the generator runs far ahead of what is consumed for many rounds in a row (the
    case in which a bounded list of generator marks lost the one it needed). The handed-back generator must
    still sit exactly n * npasses draws in."""
    x, y = synth.random_data(1000, 6, "gaussian", True, density=1.0, seed=7)
    x = x.toarray()
    rg, ro = cuda.rng_from_seed(5), oracle.rng_from_seed(5)
    g1 = sg.sgdnet(x, y, nlambda=100, thresh=1e-2, rng=rg, backend=cuda)
    r1 = sg.sgdnet(x, y, nlambda=100, thresh=1e-2, rng=ro, backend=oracle)
    assert g1.npasses == r1.npasses and g1.npasses < 400
    np.testing.assert_array_equal(g1.raw.epochs, r1.raw.epochs)
    fresh = oracle.rng_from_seed(5)                       # an untouched generator advanced by n * npasses draws
    skip = np.empty(1000 * int(r1.npasses), dtype=np.uint32)
    oracle.lib.oracle_draw_indices(C.byref(fresh), C.c_uint32(1000), C.c_int64(skip.size), skip.ctypes.data_as(C.POINTER(C.c_uint32)))
    want = oracle.unif(fresh, 5)
    np.testing.assert_array_equal(cuda.unif(rg, 5), want)
    np.testing.assert_array_equal(oracle.unif(ro, 5), want)


def test_device_rng_kernel_matches_the_sequential_generator(cuda, oracle):
    """mt_indices_kernel (rng.cu): indices, per-epoch generator snapshots and the handed-back state, word for word."""
    from test_abi_cpu import _check_device_schedule
    _check_device_schedule(cuda, oracle, on_host=False)


def test_batch_of_mixed_kernel_variants_and_sizes(cuda, oracle):
    """One batch holding fits that need different solver kernels (wavefront K = 1, generic sparse for standardize = TRUE
    and for multinomial) on row subsets of very different sizes: every fit is its own pipeline, none waits for another,
    and each equals the oracle's fit of the same rows with the same seed."""
    from sgdnet_b200 import api
    x, yb = synth.binomial_sparse(3000, 200, 10, seed=77)
    rows_small = np.arange(0, 3000, 7, dtype=np.int32)
    rows_half = np.arange(0, 3000, 2, dtype=np.int32)
    cases = [(None, dict(alpha=1.0, standardize=False)), (rows_small, dict(alpha=0.5, standardize=False)),
             (rows_half, dict(alpha=0.3, standardize=True)), (rows_small, dict(alpha=0.0, standardize=False)),
             (None, dict(alpha=0.7, standardize=True))]
    specs, keeps = [], []
    for k, (rows, kw) in enumerate(cases):
        ctl, keep = api.build_control("binomial", 1, alpha=kw["alpha"], nlambda=6, lambda_min_ratio=1e-3, lambda_=None, maxit=40,
                                      standardize=kw["standardize"], intercept=True, thresh=1e-3, standardize_response=False,
                                      debug=False)
        keeps.append(keep)
        specs.append(dict(train_rows=rows, test_rows=None, control=ctl, rng=cuda.rng_from_seed(50 + k)))
    raws, _ = cuda.fit_batch(x, yb.reshape(-1, 1), specs)
    xr = x.tocsr()
    for k, (rows, kw) in enumerate(cases):
        xs, ys = (x, yb) if rows is None else (xr[rows].tocsc(), yb[rows])
        r = sg.sgdnet(xs, ys, family="binomial", nlambda=6, lambda_min_ratio=1e-3, maxit=40, seed=50 + k, backend=oracle, **kw)
        assert_fit_parity(raws[k], r.raw)


def test_predict_and_score(cuda, oracle):
    x, y = _heart()
    fit = sg.sgdnet(x, y, family="binomial", alpha=0.5, standardize=False, nlambda=10, seed=1, backend=cuda)
    a0 = fit.raw.a0
    beta = fit.raw.beta
    rel_close(cuda.predict(x, a0, beta), oracle.predict(x, a0, beta), rtol=1e-12, what="link")
    rel_close(cuda.predict(x.toarray(), a0, beta), oracle.predict(x.toarray(), a0, beta), rtol=1e-12, what="link dense")
    yy = (y == y.max()).astype(float)
    rel_close(cuda.score_deviance(x, yy, 1, a0, beta), oracle.score_deviance(x, yy, 1, a0, beta), rtol=1e-10, what="score")


from score_cases import SCORE_CASES, check_backend_against_fixture  # noqa: E402


@pytest.mark.parametrize("name", list(SCORE_CASES))
def test_predict_and_score_match_the_r_rendering(cuda, name):
    """predict + score against tests/golden/score_fixture.npz: R/score.R and R/predict.sgdnet.R rendered in numpy
    (tests/r_score.py) on the reference build's coefficients - a pin that does not pass through the oracle."""
    assert set(check_backend_against_fixture(cuda, name)) == set(SCORE_CASES[name][2])


def test_class_measure_when_a_class_is_never_predicted_on_the_device(cuda):
    """the two-pass path of predict_score (engine.cu): classes predicted anywhere -> ids, R/score.R:153."""
    import r_score
    from score_cases import raw_coefficients
    rng = np.random.default_rng(3)
    n, p, K, L = 60, 4, 3, 3
    x = rng.normal(size=(n, p))
    y = np.arange(n) % K
    a0 = np.zeros((K, L))
    a0[1, :] = -50.0
    beta = [rng.normal(size=(p, L)) * (0.0 if k == 1 else 1.0) for k in range(K)]
    a0r, br = raw_coefficients("multinomial", a0, beta)
    got = cuda.score(x, y.astype(float).reshape(-1, 1), 2, "class", a0r, br)
    np.testing.assert_allclose(got, r_score.score("multinomial", a0, beta, x, y, "class"), rtol=1e-12)


def test_cv_batch_matches_sequential_oracle(cuda, oracle):
    x, y = synth.binomial_sparse(1500, 300, 10, seed=31)
    foldid = (np.random.Generator(np.random.PCG64(1)).permutation(1500) % 5) + 1
    kw = dict(family="binomial", alpha=[0.0, 0.5, 1.0], foldid=foldid, nlambda=8, standardize=False, maxit=100, seed=1000)
    g = sg.cv_sgdnet(x, y, backend=cuda, **kw)
    r = sg.cv_sgdnet(x, y, backend=oracle, batched=False, **kw)
    for fg, fr in zip(g.fold_fits, r.fold_fits):
        assert_fit_parity(fg.raw, fr.raw)
    for cg, cr in zip(g.cv_raw, r.cv_raw):
        rel_close(cg, cr, what="cv_raw")
    assert g.alpha_min == r.alpha_min and g.lambda_min == r.lambda_min and g.lambda_1se == r.lambda_1se


@pytest.mark.parametrize("measure", ["mse", "mae", "class"])
def test_cv_other_measures_match_the_oracle(cuda, oracle, measure):
    """type.measure other than deviance through the batch call (scored on each fit's stream as it finishes)."""
    x, y = synth.binomial_sparse(1200, 200, 10, seed=33)
    foldid = (np.random.Generator(np.random.PCG64(2)).permutation(1200) % 4) + 1
    kw = dict(family="binomial", alpha=[0.3, 1.0], foldid=foldid, nlambda=6, standardize=False, maxit=60, seed=700,
              type_measure=measure)
    g = sg.cv_sgdnet(x, y, backend=cuda, **kw)
    r = sg.cv_sgdnet(x, y, backend=oracle, batched=False, **kw)
    for cg, cr in zip(g.cv_raw, r.cv_raw):
        rel_close(cg, cr, rtol=1e-9, what="cv_raw " + measure)
    assert g.name == r.name and g.lambda_min == r.lambda_min


# ---------------------------------------------------------------- against the reference's own compiled code
from ref_vectors import CASES as REF_CASES, assert_matches_reference, case as ref_case  # noqa: E402


@pytest.mark.parametrize("name", list(REF_CASES))
def test_cuda_matches_reference_build_vectors(cuda, name):
    """tests/golden/ref_vectors.npz holds the outputs of /root/reference/src/sgdnet.cpp itself (compiled against the
    Rcpp/Eigen stand-in, oracle/refbuild/). Lambda path exact; the fixed-length cases must agree over the whole path
    (supports exact, coefficients / intercepts / deviances within 1e-6 relative); converging cases up to the first
    lambda whose epoch count the summation order moves (tests/ref_vectors.py)."""
    x, y, kw, exp = ref_case(name)
    n_cmp = assert_matches_reference(sg.sgdnet(x, y, backend=cuda, **kw).raw, exp, exact=False)
    # all committed cases compare over their whole path (see tests/test_ref_cpu.py for the Eigen-association caveat
    # that makes this a property of the chosen inputs): a case that stops comparing is a regression
    assert n_cmp == len(exp["lambda_"])


def test_edge_shapes(cuda, oracle):
    """The shapes tests/test_ref_cpu.py::test_live_edge_cases pins against the reference's code, on the device:
    all-zero and constant columns under standardisation, empty rows, a single feature, lambda = 0, an ascending user
    path, a constant response (all-zero path), more features than samples."""
    rng = np.random.default_rng(5)
    n, p = 90, 7
    x = rng.normal(size=(n, p)) * (rng.uniform(size=(n, p)) < 0.4)
    x[:, 2] = 0.0
    x[[3, 17, 40], :] = 0.0
    x[:, 4] = 1.5
    y = x @ rng.normal(size=p) + 0.1 * rng.normal(size=n)
    yb = (y > np.median(y)).astype(float)
    for std in (True, False):
        for fam, yy in (("gaussian", y), ("binomial", yb)):
            kw = dict(family=fam, alpha=0.6, standardize=std, nlambda=6, maxit=40, seed=3)
            for xx in (x, sp.csc_matrix(x)):
                g, r = both(cuda, oracle, xx, yy, **kw)
                assert_fit_parity(g.raw, r.raw)
    x1 = rng.normal(size=(60, 1))
    y1 = 2.0 * x1[:, 0] + rng.normal(size=60)
    for kw in (dict(family="gaussian", alpha=1.0, nlambda=5, seed=1), dict(family="gaussian", lambda_=[0.0], maxit=50, seed=1),
               dict(family="gaussian", alpha=0.3, lambda_=[0.001, 0.01, 0.1, 1.0], maxit=30, seed=2)):
        for xx in (x1, sp.csc_matrix(x1)):
            g, r = both(cuda, oracle, xx, y1, **kw)
            assert_fit_parity(g.raw, r.raw)
    g, r = both(cuda, oracle, x1, np.full(60, 3.25), family="gaussian", nlambda=4, seed=1)
    assert_fit_parity(g.raw, r.raw)
    assert np.all(g.raw.lambda_ == 0) and np.all(g.raw.beta == 0)
    xw = rng.normal(size=(25, 40))
    yw = xw[:, :3] @ rng.normal(size=(3, 2)) + 0.1 * rng.normal(size=(25, 2))
    g, r = both(cuda, oracle, xw, yw, family="mgaussian", nlambda=5, maxit=30, seed=4)
    assert_fit_parity(g.raw, r.raw)
    g, r = both(cuda, oracle, xw, np.arange(25) % 3, family="multinomial", alpha=0.5, nlambda=5, maxit=30, seed=4)
    assert_fit_parity(g.raw, r.raw)


def test_sparse_very_wide_design_without_the_nonzero_bitmap(cuda, oracle):
    """p = 1.9 M features: the nonzero-coefficient bitmap (238 KB) no longer fits beside the tile ring of the deviance
    pass, which then gathers every weight; the coefficient records (61 MB) spill the solver's L2 working set."""
    rng = np.random.default_rng(41)
    n, p, nnz_row = 600, 1_900_000, 20
    cols = rng.integers(0, p, size=(n, nnz_row))
    cols[:, :3] = rng.integers(0, 40, size=(n, 3))          # a few shared features so that rows conflict and a signal exists
    rows = np.repeat(np.arange(n), nnz_row)
    x = sp.csc_matrix((rng.uniform(0.5, 1.5, size=n * nnz_row), (rows, cols.ravel())), shape=(n, p))
    x.sum_duplicates()
    beta = np.zeros(40)
    beta[:6] = [2.0, -2.0, 1.5, -1.5, 1.0, -1.0]
    y = (np.asarray(x[:, :40] @ beta).ravel() + 0.3 * rng.normal(size=n) > 0).astype(float)
    g, r = both(cuda, oracle, x, y, family="binomial", alpha=1.0, standardize=False, nlambda=4, maxit=8, thresh=1e-3, seed=14)
    assert_fit_parity(g.raw, r.raw)


@pytest.mark.parametrize("family,alpha,intercept", [("binomial", 1.0, True), ("gaussian", 0.0, True), ("binomial", 0.4, False)])
@pytest.mark.parametrize("p", [900, 9000])
def test_sparse_standardized_owner_computes_kernel(cuda, oracle, family, alpha, intercept, p):
    """Sparse input with standardize = TRUE, K = 1 (saga_sparse_centred.cu): virtual centring touches every coefficient on
    every update. p = 900: state in shared memory; p = 9000: state in HBM. Ridge moves wscale (lag-scaling table, resets);
    empty rows, repeated samples and an all-zero column are in."""
    rng = np.random.default_rng(51)
    n = 700
    x = sp.random(n, p, density=14.0 / p, random_state=3, format="lil")
    x[5, :] = 0
    x[:, 7] = 0
    x = sp.csc_matrix(x)
    eta = np.asarray(x[:, :30] @ rng.normal(size=30)).ravel()
    y = (eta + 0.3 * rng.normal(size=n) > 0).astype(float) if family == "binomial" else eta + 0.1 * rng.normal(size=n)
    g, r = both(cuda, oracle, x, y, family=family, alpha=alpha, intercept=intercept, standardize=True, nlambda=6, thresh=1e-4,
                maxit=40, seed=17)
    assert_fit_parity(g.raw, r.raw)


def test_sparse_standardized_long_rows_fall_back_to_the_generic_kernel(cuda, oracle):
    """A row with more nonzeros than the owner-computes kernel stages (256) sends the fit to saga_sparse_generic_kernel."""
    rng = np.random.default_rng(52)
    x = sp.random(300, 400, density=0.05, random_state=5, format="lil")
    x[11, :330] = rng.uniform(0.2, 1.0, size=330)
    x = sp.csc_matrix(x)
    y = (rng.uniform(size=300) < 0.5).astype(float)
    g, r = both(cuda, oracle, x, y, family="binomial", alpha=0.7, standardize=True, nlambda=5, maxit=30, seed=18)
    assert_fit_parity(g.raw, r.raw)
