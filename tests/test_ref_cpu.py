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
    if name.startswith("fixed_"):
        assert n_cmp == len(exp["lambda_"])     # fixed path lengths: the whole path is comparable


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
