"""The drop-in boundary, end to end: shim/sgdnet.cpp is the file that replaces the reference's src/sgdnet.cpp (same two
exported functions, same signatures). There is no R in the image, so it is compiled against the Rcpp/Eigen stand-in
(oracle/refbuild/, target libsgdnet_shimtest.so, linked to libsgdnet_b200.so) and driven exactly like the reference's
own translation unit is in tests/test_ref_cpu.py: Eigen matrices + an Rcpp::List control in, R's Mersenne-Twister
behind unif_rand(), the list of src/sgdnet.cpp:275-284 out."""
import os
import subprocess

import numpy as np
import pytest

import sgdnet_b200 as sg
from ref_vectors import CASES, assert_matches_reference, case

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM_SO = os.path.join(ROOT, "oracle", "_ref", "libsgdnet_shimtest.so")


def load_shim():
    from sgdnet_b200._abi import Library
    if not os.path.exists(SHIM_SO):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle", "refbuild"), "../_ref/libsgdnet_shimtest.so"])
    return Library(SHIM_SO, "shimtest_")


def test_shim_builds_loads_and_fails_loudly_without_a_gpu():
    """CPU box: the shim library links against libsgdnet_b200.so and exports the entry points; with no CUDA device the
    call comes back as an error carrying the library's message (Rcpp::stop in the shim), never a CPU result."""
    import ctypes
    shim = load_shim()
    cnt = ctypes.c_int(0)
    lib = sg.product()
    has_gpu = lib.sym("device_count")(ctypes.byref(cnt)) == 0 and cnt.value > 0
    if has_gpu:
        pytest.skip("a GPU is present: covered by the gpu tests below")
    x, y, kw, _ = case("fixed_c2mini_sparse_binomial_lasso")
    with pytest.raises(sg._abi.SgdnetError, match="sgdnet_b200"):
        sg.sgdnet(x, y, backend=shim, **kw)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["c1_abalone_gaussian_enet", "heart_binomial_lasso_sparse", "wine_multinomial_enet",
                                  "student_mgaussian_grouplasso", "c2mini_sparse_binomial_lasso",
                                  "fixed_c2mini_sparse_binomial_enet", "fixed_c3mini_dense_multinomial",
                                  "fixed_sparse_multinomial_std"])
def test_shim_returns_what_the_reference_entry_point_returns(cuda, name):
    """SgdnetDense / SgdnetSparse of the shim vs (i) the same fit through the C ABI directly - identical bits, and R's
    generator consumed through the unif_rand callback - and (ii) the reference build's golden vectors."""
    shim = load_shim()
    x, y, kw, exp = case(name)
    via_shim = sg.sgdnet(x, y, backend=shim, **kw).raw
    direct = sg.sgdnet(x, y, backend=cuda, **kw).raw
    for f in ("lambda_", "epochs", "return_codes", "beta", "a0", "dev_ratio"):
        np.testing.assert_array_equal(getattr(via_shim, f), getattr(direct, f), err_msg=f)
    assert (via_shim.npasses, via_shim.nulldev) == (direct.npasses, direct.nulldev)
    assert_matches_reference(via_shim, exp, exact=False)


@pytest.mark.gpu
def test_shim_leaves_r_generator_where_the_reference_does(cuda):
    """n * npasses draws through the unif_rand callback, no read-ahead: after the call R's generator stands exactly
    where n * npasses calls of unif_rand() leave it (what .Random.seed would show in R)."""
    shim = load_shim()
    x, y, kw, exp = case("fixed_heart_binomial_lasso_sparse")
    seed = kw.pop("seed")
    live, counted = shim.rng_from_seed(seed), shim.rng_from_seed(seed)
    fit = sg.sgdnet(x, y, backend=shim, rng=live, **kw)
    assert fit.npasses == exp["npasses"]
    shim.unif(counted, 0)
    import ctypes as C
    f = shim.sym("rng_unif")
    for _ in range(x.shape[0] * fit.npasses):
        f(C.byref(counted))
    assert live.mti == counted.mti and list(live.mt) == list(counted.mt)
