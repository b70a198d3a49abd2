"""Shared parity assertions: CUDA path vs CPU oracle on the same inputs and the same sampling sequence.

The bar (BASELINE.json north_star): path lengths (epochs per lambda, npasses, return codes) and nonzero
supports bit-exact; lambda path exact; coefficients, intercepts and deviance ratios within RTOL = 1e-6
relative (FP64; the only differences allowed are reduction order and libm-vs-CUDA exp/log ulps).
"""
import numpy as np

RTOL = 1e-6


def rel_close(a, b, rtol=RTOL, what=""):
    a = np.asarray(a, dtype=float)
    b = np.asarray(b, dtype=float)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    # 0/0 deviance ratios (constant response: both deviances are exactly zero) are NaN in the reference as well
    np.testing.assert_array_equal(np.isnan(a), np.isnan(b), err_msg=f"{what}: NaN pattern")
    a, b = a[~np.isnan(a)], b[~np.isnan(b)]
    scale = max(np.max(np.abs(b)) if b.size else 0.0, 1e-300)
    err = np.max(np.abs(a - b)) if a.size else 0.0
    assert err <= rtol * scale, f"{what}: max abs diff {err:.3e} > {rtol:g} * {scale:.3e}"


def assert_fit_parity(gpu, ref, check_support=True, exact_beta=True):
    """gpu/ref: _abi.RawFit. With the oracle in its default "portable" arithmetic mode (include/sgdnet_arith.h) the
    solver loops of both arms perform the same IEEE operations in the same order, so the archived coefficients are
    expected to be IDENTICAL (exact_beta); a0 / deviance go through differently-ordered sums and are held to RTOL."""
    np.testing.assert_array_equal(gpu.lambda_, ref.lambda_, err_msg="lambda path differs")
    np.testing.assert_array_equal(gpu.epochs, ref.epochs, err_msg="epochs per lambda differ")
    np.testing.assert_array_equal(gpu.return_codes, ref.return_codes, err_msg="return codes differ")
    assert gpu.npasses == ref.npasses
    assert gpu.nulldev == ref.nulldev or abs(gpu.nulldev - ref.nulldev) <= 1e-12 * abs(ref.nulldev)
    L = len(ref.lambda_)
    for l in range(L):
        rel_close(gpu.beta[l], ref.beta[l], what=f"beta[lambda {l}]")
        if exact_beta:
            np.testing.assert_array_equal(gpu.beta[l], ref.beta[l], err_msg=f"beta not bit-identical at lambda {l}")
        if check_support:
            np.testing.assert_array_equal(gpu.beta[l] != 0, ref.beta[l] != 0, err_msg=f"support differs at lambda {l}")
    rel_close(gpu.a0, ref.a0, what="a0")
    # dev.ratio = 1 - dev/nulldev: compare the deviances it encodes
    rel_close(1.0 - gpu.dev_ratio, 1.0 - ref.dev_ratio, what="deviance")
