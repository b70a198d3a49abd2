"""Host-only checks of the product library: it loads, exports every declared symbol, its host-side RNG matches R, and
it fails loudly (no CPU fallback) when no CUDA device is usable."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import sgdnet_b200 as sg
from sgdnet_b200 import _abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    lib = sg.product()
    hdr = open(os.path.join(ROOT, "include", "sgdnet_b200.h")).read()
    names = sorted(set(re.findall(r"\b(sgdnet_[a-z_]+)\s*\(", hdr)))
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib.lib, n), n
    assert lib.lib.sgdnet_abi_version() == 3


def test_struct_sizes_match_the_header():
    src = r'''
    #include "include/sgdnet_b200.h"
    #include <stdio.h>
    int main(void){ printf("%zu %zu %zu %zu\n", sizeof(sgdnet_control), sizeof(sgdnet_rng), sizeof(sgdnet_result), sizeof(sgdnet_fit_spec)); return 0; }
    '''
    exe = os.path.join(ROOT, "build", "abi_sizes")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    subprocess.run(["gcc", "-x", "c", "-", "-I", ROOT, "-o", exe], input=src.encode(), cwd=ROOT, check=True)
    sizes = [int(v) for v in subprocess.check_output([exe]).split()]
    assert sizes == [C.sizeof(_abi.Control), C.sizeof(_abi.Rng), C.sizeof(_abi.Result), C.sizeof(_abi.FitSpec)]


@pytest.mark.parametrize("seed,expect", [(1, [0.26550866, 0.37212390, 0.57285336, 0.90820779]),
                                         (42, [0.91480604, 0.93707541, 0.28613953, 0.83044763]),
                                         (123, [0.28757752, 0.78830514, 0.40897692, 0.88301740])])
def test_product_rng_matches_r(seed, expect):
    lib = sg.product()
    np.testing.assert_allclose(lib.unif(lib.rng_from_seed(seed), 4), expect, atol=5e-9)


def test_product_and_oracle_generators_are_the_same_stream(oracle):
    lib = sg.product()
    a, b = lib.rng_from_seed(2024), oracle.rng_from_seed(2024)
    np.testing.assert_array_equal(lib.unif(a, 2000), oracle.unif(b, 2000))     # crosses 3 twists of the state
    assert a.mti == b.mti and list(a.mt) == list(b.mt)


def _check_device_schedule(lib, oracle, on_host):
    """the block-parallel MT19937 regeneration (rng.cu) against the sequential generator: indices, the generator state
    at every epoch boundary, and the state handed back; starting positions inside, at the end of and before a block,
    n smaller and larger than a block of 624 words."""
    for seed, burn, n, ne in ((1, 0, 1000, 3), (7, 5, 10, 16), (42, 623, 624, 2), (3, 624, 1247, 4), (9, 100, 50_000, 2), (11, 17, 1, 5)):
        a, b = lib.rng_from_seed(seed), oracle.rng_from_seed(seed)
        lib.unif(a, burn)
        oracle.unif(b, burn)
        seq, states = lib.rng_indices(a, n, ne, on_host=on_host)
        want = np.empty(n * ne, dtype=np.uint32)
        c = oracle.rng_from_seed(seed)
        oracle.unif(c, burn)
        for e in range(ne + 1):
            assert states[e].mti % 624 == c.mti % 624 or (states[e].mti, c.mti) in ((624, 624),), (seed, e, states[e].mti, c.mti)
            # same stream from here on (the state may be held before or after a pending block regeneration)
            s_copy = _abi.Rng.from_buffer_copy(states[e])
            c_copy = _abi.Rng.from_buffer_copy(c)
            np.testing.assert_array_equal(lib.unif(s_copy, 700), oracle.unif(c_copy, 700))
            if e < ne:
                oracle.lib.oracle_draw_indices(C.byref(c), C.c_uint32(n), C.c_int64(n), want[e * n:].ctypes.data_as(C.POINTER(C.c_uint32)))
        np.testing.assert_array_equal(seq, want)
        np.testing.assert_array_equal(lib.unif(a, 10), oracle.unif(c, 10))     # handed back advanced by n * ne draws


def test_device_rng_schedule_on_the_host(oracle):
    _check_device_schedule(sg.product(), oracle, on_host=True)


def test_bad_arguments_are_rejected_before_any_device_work():
    lib = sg.product()
    ctl, _ = _abi.make_control(0, alpha=1.0, intercept=True, standardize=True, standardize_response=False, n_lambda=3,
                               n_classes=1, debug=False, max_iter=10, lambda_min_ratio=1e-2, tol=1e-3, lambda_=None)
    res = _abi.Result()
    rng = lib.rng_from_seed(1)
    rc = lib.lib.sgdnet_fit_dense(None, C.c_int64(4), C.c_int64(2), None, C.c_int32(1), C.byref(ctl), C.byref(rng), C.byref(res))
    assert rc == 1 and b"null" in lib.lib.sgdnet_last_error()


def test_no_cpu_fallback_without_a_device():
    """Run in a subprocess with the GPUs hidden: the fit must fail with SGDNET_ERR_CUDA, never compute on the host."""
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "import numpy as np, sgdnet_b200 as sg\n"
        "try:\n"
        "    sg.sgdnet(np.random.rand(20, 2), np.random.rand(20), nlambda=2)\n"
        "except sg.SgdnetError as e:\n"
        "    print('RAISED', e); sys.exit(0)\n"
        "sys.exit(3)\n" % ROOT)
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "status 2" in out.stdout
