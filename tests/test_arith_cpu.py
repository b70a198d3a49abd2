"""include/sgdnet_arith.h: the specified exp/log are faithful (< 1 ulp against 200-bit references) and handle the
IEEE special cases."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def arith():
    src = '#include "include/sgdnet_arith.h"\nextern "C" double t_exp(double x){return sgd_exp(x);}\nextern "C" double t_log(double x){return sgd_log(x);}\n'
    so = os.path.join(ROOT, "build", "arith_test.so")
    os.makedirs(os.path.dirname(so), exist_ok=True)
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-x", "c++", "-", "-I", ROOT, "-o", so],
                   input=src.encode(), cwd=ROOT, check=True)
    lib = C.CDLL(so)
    for f in (lib.t_exp, lib.t_log):
        f.restype = C.c_double
        f.argtypes = [C.c_double]
    return lib


def ulp_error(got, ref_mp):
    import mpmath as mp
    r = float(ref_mp)
    return abs(float((mp.mpf(got) - ref_mp) / mp.mpf(float(np.spacing(abs(r))))))


def test_exp_is_faithful(arith):
    import mpmath as mp
    mp.mp.prec = 200
    rng = np.random.default_rng(0)
    xs = np.concatenate([rng.uniform(-40, 40, 4000), rng.uniform(-745, 709.7, 1500), rng.uniform(-1e-3, 1e-3, 500),
                         [0.0, 1.0, -1.0, 709.782712893384, -745.13, -708.4, -740.0]])
    worst = max(ulp_error(arith.t_exp(float(x)), mp.exp(mp.mpf(float(x)))) for x in xs)
    assert worst < 1.0
    assert arith.t_exp(0.0) == 1.0 and arith.t_exp(710.0) == np.inf and arith.t_exp(-746.0) == 0.0
    assert np.isnan(arith.t_exp(float("nan")))


def test_log_is_faithful(arith):
    import mpmath as mp
    mp.mp.prec = 200
    rng = np.random.default_rng(1)
    ys = np.concatenate([np.exp(rng.uniform(-700, 700, 3000)), rng.uniform(0.5, 2.0, 3000), rng.uniform(0.99, 1.01, 1000),
                         [2.0, 0.5, 5e-324, 1e-310, 1.7976931348623157e308]])
    worst = max(ulp_error(arith.t_log(float(y)), mp.log(mp.mpf(float(y)))) for y in ys)
    assert worst < 1.0
    assert arith.t_log(1.0) == 0.0 and arith.t_log(0.0) == -np.inf and arith.t_log(np.inf) == np.inf
    assert np.isnan(arith.t_log(-1.0))


def test_soft_threshold_select_form_equals_the_reference_form():
    """sgdnet_b200/csrc/common.cuh evaluates SoftThreshold (reference src/prox.h:32-39, std::max semantics) as
    t = |x| - s; t <= 0 ? +0 : copysign(t, x). Same bits for every finite x and s >= 0, NaN stays NaN."""
    rng = np.random.Generator(np.random.PCG64(7))
    x = np.concatenate([rng.normal(size=200000) * 10.0 ** rng.integers(-12, 12, size=200000),
                        [0.0, 1.0, -1.0, 1e-320, -1e-320, 1e308, -1e308, np.inf, -np.inf, 3.0, -3.0, 2.5, np.nan]])
    s = np.concatenate([np.abs(rng.normal(size=200000)) * 10.0 ** rng.integers(-12, 12, size=200000),
                        [0.0, 1.0, 1.0, 0.0, 1e-320, 1e308, 1e308, 1.0, np.inf, 3.0, 3.0, 0.0, 1.0]])
    # sweep near-ties: x within a few ulps of +-s
    xt = np.nextafter(s[:1000], np.inf) * np.where(np.arange(1000) % 2 == 0, 1.0, -1.0)
    x, s = np.concatenate([x, xt, s[:1000], -s[:1000]]), np.concatenate([s, s[:1000], s[:1000], s[:1000]])
    with np.errstate(invalid="ignore", over="ignore"):
        a, b = x - s, -x - s
        ref = np.where(a < 0.0, 0.0, a) - np.where(b < 0.0, 0.0, b)        # std::max(v, 0.0) is (v < 0.0) ? 0.0 : v
        t = np.abs(x) - s
        alt = np.where(t <= 0.0, 0.0, np.copysign(t, x))
    nan = np.isnan(ref)
    np.testing.assert_array_equal(nan, np.isnan(alt))
    np.testing.assert_array_equal(ref[~nan].view(np.int64), alt[~nan].view(np.int64))


def test_late_lane_butterfly_identity():
    """The fast conflict path of the wavefront kernel (saga_sparse.cu) closes a row's dot product from ONE lane: it runs
    the 32-lane xor-butterfly with that lane contributing nothing, keeps what the lane RECEIVES at the five levels, and
    later forms ((((a + r1) + r2) + r3) + r4) + r5 with the lane's real running sum a. That must be the bits every lane
    of the full butterfly ends with (sgdnet_arith.h, item 2), for every lane and any values."""
    rng = np.random.default_rng(7)

    def butterfly(v):
        v = np.array(v, dtype=np.float64)
        received = np.zeros((5, 32))
        for i, o in enumerate((16, 8, 4, 2, 1)):
            x = v[np.arange(32) ^ o]
            received[i] = x
            v = v + x
        return v, received

    for trial in range(200):
        a = rng.normal(size=32) * 10.0 ** rng.integers(-8, 8, size=32)
        full, _ = butterfly(a)
        assert np.all(full == full[0])                      # every lane ends with the same bits
        lane = int(rng.integers(0, 32))
        dry = a.copy()
        dry[lane] = 0.0
        _, rc = butterfly(dry)
        closed = a[lane]
        for i in range(5):
            closed = closed + rc[i, lane]
        assert closed == full[0]


def test_lazy_virtual_centring_does_not_preserve_the_reference_bits():
    """SURVEY.md section 8f-4 (H2) asked for an "exact lazy form" of sparse + standardize = TRUE: keep a running scalar
    for the dense correction a.row(k) -= x_center_scaled * g_change(k) * scaling that AddWeighted applies to EVERY
    feature on EVERY update (reference src/saga-sparse.h:127-128), and apply it just in time when a feature is next
    gathered. The reference rounds w_j after each of those T subtractions; the lazy form subtracts c_j * sum_t(...) once.
    Those are different IEEE operation sequences, and they do give different bits (so do the two dot products
    w . x_center_scaled of :276-277 taken over a lazily and an eagerly corrected w). Supports and path lengths are
    decided at that level (include/sgdnet_arith.h), so the library keeps the reference's per-update sweeps for this
    mode (saga_sparse_generic_kernel) instead of a lazy form that would only be tolerance-equal. This test is the
    counterexample that decision rests on."""
    rng = np.random.default_rng(11)
    T, p = 64, 4096
    c = rng.normal(size=p)
    w0 = rng.normal(size=p)
    gch = rng.normal(size=T) * 1e-2
    gamma, r = 0.3, 1.0 - 1e-4
    wscale = 1.0
    eager = w0.copy()
    run = 0.0
    for t in range(T):
        wscale *= r
        scaling = -gamma / wscale
        eager -= c * gch[t] * scaling                       # (x_center_scaled * g_change(k)) * scaling, then -=
        run += gch[t] * scaling
    lazy = w0 - c * run
    differ = np.mean(eager != lazy)
    assert differ > 0.5                                     # most elements end with different bits
    np.testing.assert_allclose(lazy, eager, rtol=0, atol=1e-13)    # while agreeing to rounding level


def test_recursive_halving_equals_the_xor_butterfly():
    """saga_dense_cluster.cu reduces the K class sums of a warp by recursive halving: at offset o = 16, 8, ... a lane keeps
    one half of its values and hands the other half to lane ^ o, so a level moves K/2, K/4, ... values instead of K. Every
    pair sum a_i + a_(i^o) is still formed exactly once (by one of the two partners; addition commutes), so lane
    (k << (5 - log2 K)) must end with the bits the plain 32-lane xor-butterfly of class k ends with (sgdnet_arith.h
    item 2), for every power-of-two K <= 32."""
    rng = np.random.default_rng(3)

    def butterfly(v):                       # v: [32] values of one class, one per lane
        v = np.array(v, dtype=np.float64)
        for o in (16, 8, 4, 2, 1):
            v = v + v[np.arange(32) ^ o]
        return v

    def halving(vals):                      # vals: [32 lanes][K]
        lanes = np.arange(32)
        cur = [np.array(vals[:, k]) for k in range(vals.shape[1])]       # cur[i][lane]
        o = 16
        while o > 0:
            n = len(cur)
            if n == 1:
                cur = [cur[0] + cur[0][lanes ^ o]]
            else:
                up = (lanes & o) != 0
                h = n // 2
                nxt = []
                for i in range(h):
                    keep = np.where(up, cur[h + i], cur[i])
                    send = np.where(up, cur[i], cur[h + i])
                    nxt.append(keep + send[lanes ^ o])
                cur = nxt
            o //= 2
        return cur[0]                       # lane L holds the total of class L >> (5 - log2 K)

    for K in (1, 2, 4, 8, 16, 32):
        shift = 5 - int(np.log2(K))
        for trial in range(20):
            vals = rng.normal(size=(32, K)) * 10.0 ** rng.integers(-6, 6, size=(32, K))
            got = halving(vals)
            for k in range(K):
                ref = butterfly(vals[:, k])
                assert np.all(ref == ref[0])
                lanes_of_k = [L for L in range(32) if (L >> shift) == k]
                assert all(got[L] == ref[0] for L in lanes_of_k), (K, k)
