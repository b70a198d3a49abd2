"""Access to tests/golden/ref_vectors.npz: outputs of the reference's own code (compiled against the Rcpp/Eigen
stand-in, oracle/refbuild/) on the bundled datasets and small seeded designs; made by tests/golden/make_ref_vectors.py."""
import ast
import os
import sys

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from make_ref_vectors import BUNDLED, CASES, GENERATED, checksum  # noqa: E402  (names and kwargs only; nothing is computed on import)

_NPZ = None


def _npz():
    global _NPZ
    if _NPZ is None:
        _NPZ = np.load(os.path.join(ROOT, "tests", "golden", "ref_vectors.npz"))
    return _NPZ


_GEN_CACHE = {}


def case_input(key):
    if key in GENERATED:
        if key not in _GEN_CACHE:
            x, _ = GENERATED[key]()          # the design from its seed; the response as stored (see make_ref_vectors.py)
            y = _npz()[f"in/{key}/y"]
            np.testing.assert_array_equal(checksum(x, y), _npz()[f"in/{key}/checksum"],
                                          err_msg=f"generated input '{key}' is not the one the reference vectors were made from")
            _GEN_CACHE[key] = (x, y)
        return _GEN_CACHE[key]
    if key in BUNDLED:
        d = np.load(os.path.join(ROOT, "tests", "golden", key + ".npz"))
        pre = ""
    else:
        d, pre = _npz(), f"in/{key}/"
    if pre + "x" in d.files:
        x = d[pre + "x"]
    else:
        x = sp.csc_matrix((d[pre + "x_x"], d[pre + "x_i"], d[pre + "x_p"]), shape=tuple(d[pre + "x_shape"]))
    return x, d[pre + "y"]


def case(name):
    """(x, y, kwargs, expected dict)"""
    key, kw = CASES[name]
    x, y = case_input(key)
    d = _npz()
    exp = {f: d[f"out/{name}/{f}"] for f in ("a0", "beta", "lambda_", "dev_ratio", "return_codes", "epochs")}
    exp["nulldev"] = float(d[f"out/{name}/nulldev"])
    exp["npasses"] = int(d[f"out/{name}/npasses"])
    return x, y, dict(kw), exp


def assert_matches_reference(fit, exp, exact, rtol=1e-6):
    """fit: _abi.RawFit; exp: the reference build's outputs.

    exact: every output bit-identical (same arithmetic as the reference build: the oracle's libm mode).

    Otherwise (the arithmetic specification of include/sgdnet_arith.h: another summation association, another exp/log):
    the lambda path is exact; wherever the path lengths agree the sampling sequences are the same and supports must be
    exact, coefficients / intercepts / deviances within rtol (north_star: 1e-6 relative, double precision). The
    convergence test (src/utils.h:240-262) is a threshold on a ratio, so an epoch count can legitimately move when a
    sum is associated differently - a real Eigen build would differ from this reference build in the same way; from the
    first lambda at which that happens the two runs draw different samples and only the fixed-length cases (thresh = 0)
    are comparable. Returns the number of lambdas compared."""
    np.testing.assert_array_equal(fit.lambda_, exp["lambda_"], err_msg="lambda path")
    if exact:
        np.testing.assert_array_equal(fit.epochs, exp["epochs"], err_msg="epochs per lambda")
        np.testing.assert_array_equal(fit.return_codes, exp["return_codes"], err_msg="return codes")
        assert fit.npasses == exp["npasses"]
        assert fit.nulldev == exp["nulldev"]
        for f in ("beta", "a0", "dev_ratio"):
            np.testing.assert_array_equal(getattr(fit, f), exp[f], err_msg=f)
        return len(exp["lambda_"])
    assert abs(fit.nulldev - exp["nulldev"]) <= 1e-12 * abs(exp["nulldev"])
    same = fit.epochs == exp["epochs"]
    n_cmp = len(same) if same.all() else int(np.argmin(same))
    for l in range(n_cmp):
        assert fit.return_codes[l] == exp["return_codes"][l]
        np.testing.assert_array_equal(fit.beta[l] != 0, exp["beta"][l] != 0, err_msg=f"nonzero support at lambda {l}")
        for f, a, b in (("beta", fit.beta[l], exp["beta"][l]), ("a0", fit.a0[l], exp["a0"][l]),
                        ("deviance", 1.0 - fit.dev_ratio[l], 1.0 - exp["dev_ratio"][l])):
            scale = max(np.max(np.abs(b)), 1e-300)
            err = np.max(np.abs(a - b))
            assert err <= rtol * scale, f"{f} at lambda {l}: {err:.3e} > {rtol:g} * {scale:.3e}"
    return n_cmp
