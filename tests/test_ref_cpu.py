"""Pins the CPU oracle (oracle/sgdnet_oracle.cpp) against the REFERENCE'S OWN CODE.

oracle/_ref/libsgdnet_ref.so is /root/reference/src/sgdnet.cpp with every header it includes, unmodified, compiled from
where it lies against a stand-in for Rcpp/Eigen (oracle/refbuild/). Two layers:
  * golden vectors (tests/golden/ref_vectors.npz, made by tests/golden/make_ref_vectors.py from that library) - these
    travel, so the checks run on any box;
  * a live grid against the library itself wherever it exists (this container builds it; the GPU box receives the
    prebuilt file; nothing reads /root/reference at test time).
The oracle's "libm" arithmetic (std::exp/log, sequential sums) is the stand-in's arithmetic, so that mode must agree
BIT FOR BIT; the "portable" mode (the arithmetic specification the CUDA library implements) must agree within the
north_star tolerance with identical path lengths and supports. Runs without a GPU."""
import os
import subprocess

import numpy as np
import pytest

import sgdnet_b200 as sg
from ref_vectors import CASES, assert_matches_reference, case
from sgdnet_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libsgdnet_ref.so")
REF_SRC = "/root/reference/src/sgdnet.cpp"


@pytest.fixture
def oracle_libm(oracle):
    oracle.lib.oracle_set_arith(0)
    yield oracle
    oracle.lib.oracle_set_arith(1)


@pytest.fixture(scope="module")
def ref():
    if os.path.exists(REF_SRC):     # build container: (re)build from the reference's sources where they lie
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle", "refbuild")])
    if not os.path.exists(REF_SO):
        pytest.skip("oracle/_ref/libsgdnet_ref.so not present (built only where /root/reference exists)")
    from sgdnet_b200._abi import Library
    return Library(REF_SO, "ref_")


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_libm_reproduces_reference_vectors_bit_for_bit(oracle_libm, name):
    x, y, kw, exp = case(name)
    assert_matches_reference(sg.sgdnet(x, y, backend=oracle_libm, **kw).raw, exp, exact=True)


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_portable_matches_reference_vectors(oracle, name):
    x, y, kw, exp = case(name)
    assert oracle.lib.oracle_get_arith() == 1
    n_cmp = assert_matches_reference(sg.sgdnet(x, y, backend=oracle, **kw).raw, exp, exact=False)
    # Every committed case is comparable over its WHOLE path: the fixed-length ones by construction, the converging
    # ones because their epoch counts agree lambda by lambda. That is a property of these inputs, not a guarantee: the
    # arithmetic specification associates dense reductions differently from the reference build's sequential sums (as
    # a real Eigen build would), and a convergence ratio that lands within an ulp-level distance of `thresh` can move
    # an epoch count (mgaussian 300 x 40 with data seed 1004 does: 94 vs 91 epochs at the first lambda). A case that
    # stops comparing is a regression to look at, so the count is asserted rather than tolerated.
    assert n_cmp == len(exp["lambda_"])


def test_c1_path_length_is_the_reference_builds(oracle):
    assert case("c1_abalone_gaussian_enet")[3]["npasses"] == 1143


@pytest.mark.parametrize("seed,expect", [(1, [0.26550866, 0.37212390, 0.57285336, 0.90820779]),
                                         (42, [0.91480604, 0.93707541, 0.28613953, 0.83044763]),
                                         (123, [0.28757752, 0.78830514, 0.40897692, 0.88301740])])
def test_ref_entry_rng_kat_and_same_stream_as_oracle(ref, oracle, seed, expect):
    np.testing.assert_allclose(ref.unif(ref.rng_from_seed(seed), 4), expect, atol=5e-9)
    np.testing.assert_array_equal(ref.unif(ref.rng_from_seed(seed), 2000), oracle.unif(oracle.rng_from_seed(seed), 2000))


@pytest.mark.parametrize("family", ["gaussian", "binomial", "multinomial", "mgaussian"])
@pytest.mark.parametrize("alpha", [0.0, 0.5, 1.0])
def test_live_grid_oracle_equals_reference_code(ref, oracle_libm, family, alpha):
    """dense and sparse, with and without intercept / standardisation: every output identical to the reference's code."""
    for intercept in (True, False):
        for standardize in (True, False):
            xs, y = synth.random_data(120, 6, family, intercept=intercept, density=0.5, seed=3)
            kw = dict(family=family, alpha=alpha, intercept=intercept, standardize=standardize, nlambda=8, maxit=60, seed=2)
            for x in (xs.toarray(), xs):
                a = sg.sgdnet(x, y, backend=oracle_libm, **kw).raw
                b = sg.sgdnet(x, y, backend=ref, **kw).raw
                for f in ("lambda_", "epochs", "return_codes", "beta", "a0", "dev_ratio"):
                    np.testing.assert_array_equal(getattr(a, f), getattr(b, f), err_msg=f"{f} {kw}")
                assert (a.npasses, a.nulldev) == (b.npasses, b.nulldev)


def test_live_user_lambda_debug_losses_and_max_iter(ref, oracle_libm):
    """user lambda in the given order, debug losses (EpochLoss, src/utils.h:185-214) and the max_iter return code."""
    xs, y = synth.random_data(150, 5, "binomial", True, density=0.6, seed=12)
    kw = dict(family="binomial", alpha=0.7, lambda_=[0.02, 0.05, 0.001], maxit=7, debug=True, seed=5)
    a = sg.sgdnet(xs, y, backend=oracle_libm, **kw).raw
    b = sg.sgdnet(xs, y, backend=ref, **kw).raw
    np.testing.assert_array_equal(a.lambda_, b.lambda_)
    np.testing.assert_array_equal(a.return_codes, b.return_codes)
    assert 1 in a.return_codes
    for la, lb in zip(a.losses, b.losses):
        np.testing.assert_array_equal(la, lb)
    np.testing.assert_array_equal(a.beta, b.beta)


def test_live_wscale_reset_path(ref, oracle_libm):
    """ridge with a large step so that wscale falls below SMALL inside an epoch (src/saga-sparse.h:285-295,
    src/saga-dense.h:163-167)."""
    xs, y = synth.random_data(4000, 4, "gaussian", True, density=0.9, seed=21)
    kw = dict(family="gaussian", alpha=0.0, lambda_=[50.0], standardize=False, maxit=3, seed=1)
    for x in (xs, xs.toarray()):
        a = sg.sgdnet(x, y, backend=oracle_libm, **kw).raw
        b = sg.sgdnet(x, y, backend=ref, **kw).raw
        np.testing.assert_array_equal(a.beta, b.beta)
        np.testing.assert_array_equal(a.a0, b.a0)


def _same(a, b, what):
    for f in ("lambda_", "epochs", "return_codes", "beta", "a0", "dev_ratio"):
        np.testing.assert_array_equal(getattr(a, f), getattr(b, f), err_msg=f"{f} {what}")
    assert (a.npasses, a.nulldev) == (b.npasses, b.nulldev), what


def test_live_edge_cases_oracle_equals_reference_code(ref, oracle_libm):
    """Shapes the reference's own tests poke at (tests/testthat/test-gaussian.R:62-71, test-sparse.R, test-lambda-path.R):
    constant and all-zero columns under standardisation (sd == 0 -> 1, src/math.h:108, 128), empty rows, a single
    feature, constant response (lambda_max == 0 -> all-zero path, src/utils.h:160-165), lambda = 0, a user path given
    in ascending order, more features than samples."""
    import scipy.sparse as sp
    rng = np.random.default_rng(5)
    n, p = 90, 7
    x = rng.normal(size=(n, p)) * (rng.uniform(size=(n, p)) < 0.4)
    x[:, 2] = 0.0            # empty column
    x[:, 4] = 1.5            # constant column
    x[[3, 17, 40], :] = 0.0  # empty rows (column 4 included: it is then no longer constant, on purpose for one variant)
    y = x @ rng.normal(size=p) + 0.1 * rng.normal(size=n)
    yb = (y > np.median(y)).astype(float)
    x2 = x.copy()
    x2[:, 4] = 1.5           # truly constant column
    for xm, tag in ((x, "empty rows"), (x2, "constant column")):
        for std in (True, False):
            for fam, yy in (("gaussian", y), ("binomial", yb)):
                kw = dict(family=fam, alpha=0.6, standardize=std, nlambda=6, maxit=40, seed=3)
                for xx in (xm, sp.csc_matrix(xm)):
                    _same(sg.sgdnet(xx, yy, backend=oracle_libm, **kw).raw, sg.sgdnet(xx, yy, backend=ref, **kw).raw,
                          f"{tag} std={std} {fam} sparse={sp.issparse(xx)}")
    # a single feature; lambda = 0; ascending user path
    x1 = rng.normal(size=(60, 1))
    y1 = 2.0 * x1[:, 0] + rng.normal(size=60)
    for kw in (dict(family="gaussian", alpha=1.0, nlambda=5, seed=1), dict(family="gaussian", lambda_=[0.0], maxit=50, seed=1),
               dict(family="gaussian", alpha=0.3, lambda_=[0.001, 0.01, 0.1, 1.0], maxit=30, seed=2)):
        for xx in (x1, sp.csc_matrix(x1)):
            _same(sg.sgdnet(xx, y1, backend=oracle_libm, **kw).raw, sg.sgdnet(xx, y1, backend=ref, **kw).raw, str(kw))
    # constant response: lambda_max == 0, every lambda 0, every coefficient 0 (test-gaussian.R:62-71)
    yc = np.full(60, 3.25)
    a = sg.sgdnet(x1, yc, backend=oracle_libm, family="gaussian", nlambda=4, seed=1).raw
    b = sg.sgdnet(x1, yc, backend=ref, family="gaussian", nlambda=4, seed=1).raw
    _same(a, b, "constant response")
    assert np.all(b.lambda_ == 0) and np.all(b.beta == 0) and np.allclose(b.a0, 3.25)
    # more features than samples (lambda.min.ratio 0.01 branch of R/sgdnet.R:191-192), mgaussian + multinomial
    xw = rng.normal(size=(25, 40))
    yw = xw[:, :3] @ rng.normal(size=(3, 2)) + 0.1 * rng.normal(size=(25, 2))
    _same(sg.sgdnet(xw, yw, backend=oracle_libm, family="mgaussian", nlambda=5, maxit=30, seed=4).raw,
          sg.sgdnet(xw, yw, backend=ref, family="mgaussian", nlambda=5, maxit=30, seed=4).raw, "n < p mgaussian")
    yk = np.arange(25) % 3
    _same(sg.sgdnet(xw, yk, backend=oracle_libm, family="multinomial", alpha=0.5, nlambda=5, maxit=30, seed=4).raw,
          sg.sgdnet(xw, yk, backend=ref, family="multinomial", alpha=0.5, nlambda=5, maxit=30, seed=4).raw, "n < p multinomial")
