"""ctypes view of include/sgdnet_b200.h.

`Library` binds one shared object that exports the C ABI under a symbol prefix. The product binds
`libsgdnet_b200.so` with prefix ``sgdnet_``; tests bind the CPU oracle (same signatures, prefix
``oracle_``) through the same class so both arms run under the identical front end. This module
never loads anything under ``oracle/`` by itself.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np

GAUSSIAN, BINOMIAL, MULTINOMIAL, MGAUSSIAN = 0, 1, 2, 3
FAMILIES = {"gaussian": GAUSSIAN, "binomial": BINOMIAL, "multinomial": MULTINOMIAL, "mgaussian": MGAUSSIAN}
RNG_MT, RNG_CALLBACK, RNG_SEQUENCE = 0, 1, 2

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)
c_uint32_p = C.POINTER(C.c_uint32)
c_int64_p = C.POINTER(C.c_int64)

UNIF_FN = C.CFUNCTYPE(C.c_double, C.c_void_p)


class Control(C.Structure):
    """struct sgdnet_control (the `control` list of R/sgdnet.R:346-359)."""
    _fields_ = [
        ("family", C.c_int32),
        ("intercept", C.c_int32),
        ("standardize", C.c_int32),
        ("standardize_response", C.c_int32),
        ("n_lambda", C.c_int32),
        ("n_classes", C.c_int32),
        ("debug", C.c_int32),
        ("grouped_multinomial", C.c_int32),
        ("max_iter", C.c_uint32),
        ("lambda_len", C.c_int32),
        ("elasticnet_mix", C.c_double),
        ("lambda_min_ratio", C.c_double),
        ("tol", C.c_double),
        ("lambda_", c_double_p),
    ]


class Rng(C.Structure):
    """struct sgdnet_rng."""
    _fields_ = [
        ("kind", C.c_int32),
        ("mti", C.c_int32),
        ("mt", C.c_uint32 * 624),
        ("unif_rand", UNIF_FN),
        ("ctx", C.c_void_p),
        ("seq", c_uint32_p),
        ("seq_len", C.c_int64),
        ("seq_pos", C.c_int64),
    ]


class Result(C.Structure):
    """struct sgdnet_result (the list of src/sgdnet.cpp:275-284 plus measurement extensions)."""
    _fields_ = [
        ("n_lambda", C.c_int32),
        ("n_classes", C.c_int32),
        ("n_features", C.c_int64),
        ("a0", c_double_p),
        ("beta", c_double_p),
        ("lambda_", c_double_p),
        ("dev_ratio", c_double_p),
        ("return_codes", c_uint32_p),
        ("epochs", c_uint32_p),
        ("losses", c_double_p),
        ("losses_ptr", c_int64_p),
        ("nulldev", C.c_double),
        ("npasses", C.c_uint32),
        ("seconds_total", C.c_double),
        ("seconds_setup", C.c_double),
        ("seconds_solver", C.c_double),
        ("seconds_deviance", C.c_double),
        ("kernel_launches", C.c_uint64),
    ]


class FitSpec(C.Structure):
    """struct sgdnet_fit_spec."""
    _fields_ = [
        ("train_rows", c_int32_p),
        ("n_train", C.c_int64),
        ("test_rows", c_int32_p),
        ("n_test", C.c_int64),
        ("lambda_from", C.c_int32),
        ("measure", C.c_int32),
        ("path_only", C.c_int32),
        ("pad_", C.c_int32),
        ("control", Control),
        ("rng", Rng),
    ]


@dataclass
class RawFit:
    """Plain-numpy copy of a sgdnet_result."""
    a0: np.ndarray          # (n_lambda, K)
    beta: np.ndarray        # (n_lambda, p, K)
    lambda_: np.ndarray
    dev_ratio: np.ndarray
    return_codes: np.ndarray
    epochs: np.ndarray
    losses: list
    nulldev: float
    npasses: int
    seconds_total: float = 0.0
    seconds_setup: float = 0.0
    seconds_solver: float = 0.0
    seconds_deviance: float = 0.0
    kernel_launches: int = 0


class SgdnetError(RuntimeError):
    pass


def _as_f64(a, order="F"):
    return np.require(np.asarray(a, dtype=np.float64), requirements=["A", "F_CONTIGUOUS" if order == "F" else "C_CONTIGUOUS"])


def _ptr(a, typ):
    return a.ctypes.data_as(typ)


@dataclass
class CscMatrix:
    """A dgCMatrix: 0-based int32 row ids `i`, column pointers `p`, values `x` (n x p)."""
    i: np.ndarray
    p: np.ndarray
    x: np.ndarray
    shape: tuple

    @staticmethod
    def from_any(m) -> "CscMatrix":
        import scipy.sparse as sp
        m = sp.csc_matrix(m)
        m.sort_indices()
        m.sum_duplicates()
        return CscMatrix(np.ascontiguousarray(m.indices, dtype=np.int32), np.ascontiguousarray(m.indptr, dtype=np.int32),
                         np.ascontiguousarray(m.data, dtype=np.float64), m.shape)


def is_sparse(x) -> bool:
    if isinstance(x, CscMatrix):
        return True
    try:
        import scipy.sparse as sp
        return sp.issparse(x)
    except ImportError:  # pragma: no cover
        return False


def make_control(family: int, *, alpha: float, intercept: bool, standardize: bool, standardize_response: bool,
                 n_lambda: int, n_classes: int, debug: bool, max_iter: int, lambda_min_ratio: float, tol: float,
                 lambda_: Optional[Sequence[float]]):
    """Build a Control and the numpy buffer keeping its lambda pointer alive."""
    ctl = Control()
    ctl.family = family
    ctl.intercept = int(bool(intercept))
    ctl.standardize = int(bool(standardize))
    ctl.standardize_response = int(bool(standardize_response))
    ctl.n_lambda = int(n_lambda)
    ctl.n_classes = int(n_classes)
    ctl.debug = int(bool(debug))
    ctl.grouped_multinomial = 0
    ctl.max_iter = int(max_iter)
    ctl.elasticnet_mix = float(alpha)
    ctl.lambda_min_ratio = float(lambda_min_ratio)
    ctl.tol = float(tol)
    keep = None
    if lambda_ is not None and len(lambda_) > 0:
        keep = np.ascontiguousarray(lambda_, dtype=np.float64)
        ctl.lambda_ = _ptr(keep, c_double_p)
        ctl.lambda_len = int(keep.size)
    else:
        ctl.lambda_ = c_double_p()
        ctl.lambda_len = 0
    return ctl, keep


class Library:
    """One loaded implementation of the C ABI."""

    def __init__(self, path: str, prefix: str):
        if not os.path.exists(path):
            raise SgdnetError(f"shared library not found: {path}")
        self.path = path
        self.prefix = prefix
        self.lib = C.CDLL(path, mode=C.RTLD_GLOBAL if False else C.DEFAULT_MODE)
        self._bind()

    # -- symbol plumbing ----------------------------------------------------------------------
    def sym(self, name):
        return getattr(self.lib, self.prefix + name)

    def has(self, name) -> bool:
        try:
            self.sym(name)
            return True
        except AttributeError:
            return False

    def _bind(self):
        f = self.sym("rng_set_seed"); f.argtypes = [C.POINTER(Rng), C.c_uint32]; f.restype = None
        f = self.sym("rng_unif"); f.argtypes = [C.POINTER(Rng)]; f.restype = C.c_double
        f = self.sym("last_error"); f.argtypes = []; f.restype = C.c_char_p
        f = self.sym("result_free"); f.argtypes = [C.POINTER(Result)]; f.restype = None
        dense_x = [c_double_p, C.c_int64, C.c_int64]
        sparse_x = [c_int32_p, c_int32_p, c_double_p, C.c_int64, C.c_int64]
        fit_tail = [c_double_p, C.c_int32, C.POINTER(Control), C.POINTER(Rng), C.POINTER(Result)]
        f = self.sym("fit_dense"); f.argtypes = dense_x + fit_tail; f.restype = C.c_int
        f = self.sym("fit_sparse"); f.argtypes = sparse_x + fit_tail; f.restype = C.c_int
        coef = [c_double_p, c_double_p, C.c_int32, C.c_int32]
        for nm, xa in (("dense", dense_x), ("sparse", sparse_x)):
            if self.has("predict_" + nm):
                f = self.sym("predict_" + nm); f.argtypes = xa + coef + [c_double_p]; f.restype = C.c_int
            if self.has("score_deviance_" + nm):
                f = self.sym("score_deviance_" + nm)
                f.argtypes = xa + [c_double_p, C.c_int32, C.c_int32] + coef + [c_double_p]; f.restype = C.c_int
            if self.has("score_" + nm):
                f = self.sym("score_" + nm)
                f.argtypes = xa + [c_double_p, C.c_int32, C.c_int32, C.c_int32] + coef + [C.POINTER(Rng), c_double_p]; f.restype = C.c_int
            if self.has("fit_batch_" + nm):
                f = self.sym("fit_batch_" + nm)
                f.argtypes = xa + [c_double_p, C.c_int32, C.POINTER(FitSpec), C.c_int32, C.POINTER(Result), c_double_p]
                f.restype = C.c_int
            if self.has("session_create_" + nm):
                f = self.sym("session_create_" + nm)
                f.argtypes = xa + [c_double_p, C.c_int32, C.POINTER(Control), C.POINTER(C.c_void_p)]; f.restype = C.c_int
        if self.has("session_run_epochs"):
            f = self.sym("session_run_epochs")
            f.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(Rng), C.POINTER(C.c_float)]; f.restype = C.c_int
            f = self.sym("session_fit_lambda")
            f.argtypes = [C.c_void_p, C.c_int32, C.POINTER(Rng), c_uint32_p, c_int32_p]; f.restype = C.c_int
            f = self.sym("session_finish_lambda")
            f.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_float)]; f.restype = C.c_int
            f = self.sym("session_result"); f.argtypes = [C.c_void_p, C.POINTER(Result)]; f.restype = C.c_int
            f = self.sym("session_destroy"); f.argtypes = [C.c_void_p]; f.restype = None
        if self.has("rng_indices"):
            f = self.sym("rng_indices")
            f.argtypes = [C.POINTER(Rng), C.c_uint32, C.c_int32, c_uint32_p, C.POINTER(Rng), C.c_int32]; f.restype = C.c_int
        if self.has("set_device"):
            f = self.sym("set_device"); f.argtypes = [C.c_int]; f.restype = C.c_int
            f = self.sym("device_count"); f.argtypes = [C.POINTER(C.c_int)]; f.restype = C.c_int

    def check(self, rc: int, what: str):
        if rc != 0:
            msg = self.sym("last_error")()
            raise SgdnetError(f"{what} failed with status {rc}: {msg.decode() if msg else ''}")

    # -- RNG ----------------------------------------------------------------------------------
    def rng_from_seed(self, seed: int) -> Rng:
        r = Rng()
        self.sym("rng_set_seed")(C.byref(r), C.c_uint32(seed & 0xFFFFFFFF))
        return r

    @staticmethod
    def rng_from_sequence(seq: np.ndarray):
        seq = np.ascontiguousarray(seq, dtype=np.uint32)
        r = Rng()
        r.kind = RNG_SEQUENCE
        r.seq = _ptr(seq, c_uint32_p)
        r.seq_len = int(seq.size)
        r.seq_pos = 0
        r._keep = seq  # keep the buffer alive with the struct
        return r

    def unif(self, rng: Rng, count: int) -> np.ndarray:
        f = self.sym("rng_unif")
        return np.array([f(C.byref(rng)) for _ in range(count)])

    def rng_indices(self, rng: Rng, n: int, n_epochs: int, on_host: bool = False):
        """(indices [n_epochs * n], list of n_epochs + 1 generator states); `rng` is advanced in place."""
        seq = np.empty(n * n_epochs, dtype=np.uint32)
        states = (Rng * (n_epochs + 1))()
        self.check(self.sym("rng_indices")(C.byref(rng), C.c_uint32(n), C.c_int32(n_epochs), _ptr(seq, c_uint32_p), states,
                                           C.c_int32(1 if on_host else 0)), "rng_indices")
        return seq, states

    # -- fits ---------------------------------------------------------------------------------
    @staticmethod
    def _x_args(x):
        """(kind, ctypes args, keepalive)"""
        if is_sparse(x):
            m = x if isinstance(x, CscMatrix) else CscMatrix.from_any(x)
            n, p = m.shape
            return "sparse", [_ptr(m.i, c_int32_p), _ptr(m.p, c_int32_p), _ptr(m.x, c_double_p), C.c_int64(n), C.c_int64(p)], m
        xa = _as_f64(x)
        if xa.ndim != 2:
            raise SgdnetError("x must be two-dimensional")
        n, p = xa.shape
        return "dense", [_ptr(xa, c_double_p), C.c_int64(n), C.c_int64(p)], xa

    def take_result(self, res: Result, free: bool = True) -> RawFit:
        L, K, p = res.n_lambda, res.n_classes, res.n_features
        arr = lambda ptr, n, dt: np.ctypeslib.as_array(ptr, shape=(n,)).astype(dt, copy=True) if n > 0 else np.zeros(0, dt)
        lp = arr(res.losses_ptr, L + 1, np.int64)
        flat = arr(res.losses, int(lp[-1]) if L > 0 else 0, np.float64)
        out = RawFit(
            a0=arr(res.a0, L * K, np.float64).reshape(L, K),
            beta=arr(res.beta, L * K * p, np.float64).reshape(L, p, K),
            lambda_=arr(res.lambda_, L, np.float64),
            dev_ratio=arr(res.dev_ratio, L, np.float64),
            return_codes=arr(res.return_codes, L, np.uint32),
            epochs=arr(res.epochs, L, np.uint32),
            losses=[flat[lp[i]:lp[i + 1]].copy() for i in range(L)],
            nulldev=float(res.nulldev), npasses=int(res.npasses),
            seconds_total=res.seconds_total, seconds_setup=res.seconds_setup, seconds_solver=res.seconds_solver,
            seconds_deviance=res.seconds_deviance, kernel_launches=int(res.kernel_launches))
        if free:
            self.sym("result_free")(C.byref(res))
        return out

    def fit(self, x, y: np.ndarray, ctl: Control, rng: Rng) -> RawFit:
        kind, xargs, _keep = self._x_args(x)
        ya = _as_f64(np.asarray(y, dtype=np.float64).reshape(len(y), -1))
        res = Result()
        rc = self.sym("fit_" + kind)(*xargs, _ptr(ya, c_double_p), C.c_int32(ya.shape[1]), C.byref(ctl), C.byref(rng),
                                     C.byref(res))
        self.check(rc, "fit_" + kind)
        return self.take_result(res)

    def fit_batch(self, x, y: np.ndarray, specs: Sequence[dict]):
        """specs: dicts with train_rows, test_rows (or None), control, rng. Returns (fits, scores)."""
        kind, xargs, _keep = self._x_args(x)
        ya = _as_f64(np.asarray(y, dtype=np.float64).reshape(len(y), -1))
        nf = len(specs)
        arr = (FitSpec * nf)()
        keep = []
        n_lambda = 0
        for i, s in enumerate(specs):
            tr = s.get("train_rows")
            te = s.get("test_rows")
            if tr is not None:
                tr = np.ascontiguousarray(tr, dtype=np.int32); keep.append(tr)
                arr[i].train_rows = _ptr(tr, c_int32_p); arr[i].n_train = tr.size
            if te is not None and len(te) > 0:
                te = np.ascontiguousarray(te, dtype=np.int32); keep.append(te)
                arr[i].test_rows = _ptr(te, c_int32_p); arr[i].n_test = te.size
            arr[i].lambda_from = int(s.get("lambda_from", -1))
            arr[i].measure = int(s.get("measure", 0))
            arr[i].path_only = int(bool(s.get("path_only", False)))
            arr[i].control = s["control"]
            arr[i].rng = s["rng"]
            n_lambda = max(n_lambda, s["control"].n_lambda)
        results = (Result * nf)()
        scores = np.full((nf, n_lambda), np.nan)
        rc = self.sym("fit_batch_" + kind)(*xargs, _ptr(ya, c_double_p), C.c_int32(ya.shape[1]), arr, C.c_int32(nf),
                                           results, _ptr(scores, c_double_p))
        self.check(rc, "fit_batch_" + kind)
        for i, s in enumerate(specs):       # hand the advanced RNG state back
            s["rng"] = arr[i].rng
        return [self.take_result(results[i]) for i in range(nf)], scores

    def predict(self, x, a0: np.ndarray, beta: np.ndarray) -> np.ndarray:
        """link with shape (n_lambda, K, n)."""
        kind, xargs, _keep = self._x_args(x)
        L, p, K = beta.shape
        n = xargs[-2].value
        a0c = np.ascontiguousarray(a0, dtype=np.float64).reshape(L, K)
        bc = np.ascontiguousarray(beta, dtype=np.float64)
        out = np.empty((L, K, n))
        rc = self.sym("predict_" + kind)(*xargs, _ptr(a0c, c_double_p), _ptr(bc, c_double_p), C.c_int32(L), C.c_int32(K),
                                         _ptr(out, c_double_p))
        self.check(rc, "predict_" + kind)
        return out

    def score(self, x, y, family: int, measure: str, a0: np.ndarray, beta: np.ndarray, rng: Optional[Rng] = None) -> np.ndarray:
        """score() for one type.measure (R/score.R); `rng` supplies auc's tie-breaking draws and is advanced in place."""
        kind, xargs, _keep = self._x_args(x)
        L, p, K = beta.shape
        ya = _as_f64(np.asarray(y, dtype=np.float64).reshape(len(y), -1))
        a0c = np.ascontiguousarray(a0, dtype=np.float64).reshape(L, K)
        bc = np.ascontiguousarray(beta, dtype=np.float64)
        out = np.empty(L)
        rc = self.sym("score_" + kind)(*xargs, _ptr(ya, c_double_p), C.c_int32(ya.shape[1]), C.c_int32(family),
                                       C.c_int32(MEASURES[measure]), _ptr(a0c, c_double_p), _ptr(bc, c_double_p), C.c_int32(L),
                                       C.c_int32(K), C.byref(rng) if rng is not None else None, _ptr(out, c_double_p))
        self.check(rc, "score_" + kind)
        return out

    def score_deviance(self, x, y, family: int, a0: np.ndarray, beta: np.ndarray) -> np.ndarray:
        kind, xargs, _keep = self._x_args(x)
        L, p, K = beta.shape
        ya = _as_f64(np.asarray(y, dtype=np.float64).reshape(len(y), -1))
        a0c = np.ascontiguousarray(a0, dtype=np.float64).reshape(L, K)
        bc = np.ascontiguousarray(beta, dtype=np.float64)
        out = np.empty(L)
        rc = self.sym("score_deviance_" + kind)(*xargs, _ptr(ya, c_double_p), C.c_int32(ya.shape[1]), C.c_int32(family),
                                                _ptr(a0c, c_double_p), _ptr(bc, c_double_p), C.c_int32(L), C.c_int32(K),
                                                _ptr(out, c_double_p))
        self.check(rc, "score_deviance_" + kind)
        return out


MEASURES = {"deviance": 0, "mse": 1, "mae": 2, "class": 3, "auc": 4}

_PRODUCT: Optional[Library] = None


def product_library_path() -> str:
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "libsgdnet_b200.so")


def product() -> Library:
    """The CUDA backend. Fails loudly when the extension has not been built (no CPU fallback)."""
    global _PRODUCT
    if _PRODUCT is None:
        path = product_library_path()
        if not os.path.exists(path):
            raise SgdnetError(f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(there is no CPU fallback)")
        _PRODUCT = Library(path, "sgdnet_")
    return _PRODUCT
