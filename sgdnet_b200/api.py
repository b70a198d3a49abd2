"""Host-side mirror of sgdnet's R front end, over the C ABI.

There is no R in the build image, so the reference's callers are restated here with the same names,
argument meaning and error behaviour, so the parity tests read like tests/testthat/*.R:

    sgdnet()          R/sgdnet.R:183-433          (validation, response encoding, `control`, reshaping)
    predict()/coef()  R/predict.sgdnet.R:347-566  (link / response / class / coefficients / nonzero)
    deviance()        R/deviance.sgdnet.R:33-41
    score()           R/score.R:55-178
    cv_sgdnet()       R/cv_sgdnet.R:113-254       (incl. the train-on-fold-j quirk, :182-183)

Every numerical call goes through a `_abi.Library`; the default is the CUDA library (no CPU fallback).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

from . import _abi
from ._abi import FAMILIES, Library, SgdnetError


@dataclass
class SgdnetFit:
    """Object of class c("sgdnet_<family>", "sgdnet") (R/sgdnet.R:412-425)."""
    family: str
    a0: np.ndarray            # (n_lambda,) for gaussian/binomial; (K, n_lambda) otherwise
    beta: object              # (p, n_lambda) array, or list of K such arrays
    lambda_: np.ndarray
    dev_ratio: np.ndarray
    df: np.ndarray
    nulldev: float
    npasses: int
    alpha: float
    classnames: Optional[list]
    nobs: int
    grouped: bool = False
    offset: bool = False
    dfmat: Optional[np.ndarray] = None
    diagnostics: Optional[dict] = None
    # not part of the R object (R drops them, R/sgdnet.R:412-425) but kept for parity checks
    return_codes: Optional[np.ndarray] = None
    epochs: Optional[np.ndarray] = None
    raw: Optional[_abi.RawFit] = None
    call: dict = field(default_factory=dict)

    @property
    def n_classes(self):
        return 1 if self.family in ("gaussian", "binomial") else len(self.beta)


def _encode_response(y, family: str):
    """R/sgdnet.R:277-339. Returns (y matrix n x cols, n_classes, class_names)."""
    y = np.asarray(y)
    n_targets = 1 if y.ndim == 1 else y.shape[1]
    if family == "gaussian":
        if n_targets > 1:
            raise ValueError("response for Gaussian regression must be one-dimensional.")
        if not np.issubdtype(y.dtype, np.number):
            raise ValueError("non-numeric response.")
        return np.asarray(y, dtype=np.float64).reshape(-1, 1), 1, None
    if family == "binomial":
        levels, codes, counts = np.unique(y.reshape(-1), return_inverse=True, return_counts=True)
        if len(levels) > 2:
            raise ValueError("more than two classes in response. Are you looking for family = 'multinomial'?")
        if len(levels) == 1:
            raise ValueError("only one class in response.")
        if counts.min() <= 1:
            raise ValueError(f"one class only has {counts.min()} observations.")
        return codes.astype(np.float64).reshape(-1, 1), 1, [str(v) for v in levels]
    if family == "multinomial":
        levels, codes, counts = np.unique(y.reshape(-1), return_inverse=True, return_counts=True)
        if len(levels) == 2:
            raise ValueError("only two classes in response. Are you looking for family = 'binomial'?")
        if len(levels) == 1:
            raise ValueError("only one class in response.")
        if counts.min() <= 1:
            raise ValueError(f"one class only has {counts.min()} observations.")
        return codes.astype(np.float64).reshape(-1, 1), len(levels), [str(v) for v in levels]
    if family == "mgaussian":
        if n_targets == 1:
            raise ValueError("response for multivariate Gaussian regression must not be one-dimensional; "
                             "try family = 'gaussian'.")
        if not np.issubdtype(y.dtype, np.number):
            raise ValueError("non-numeric response.")
        return np.asarray(y, dtype=np.float64), n_targets, [f"y{i + 1}" for i in range(n_targets)]
    raise ValueError(f"'arg' should be one of {list(FAMILIES)}")


def _shape(x):
    return x.shape


def _rows(x, idx):
    """x[idx, , drop = FALSE] for dense arrays and scipy sparse matrices."""
    if _abi.is_sparse(x):
        import scipy.sparse as sp
        if isinstance(x, _abi.CscMatrix):
            x = sp.csc_matrix((x.x, x.i, x.p), shape=x.shape)
        return sp.csr_matrix(x)[idx].tocsc()
    return np.asarray(x)[idx]


def validate(x, y, alpha, nlambda, lambda_, maxit, thresh):
    """The stop() conditions of R/sgdnet.R:211-263 (same messages)."""
    n = _shape(x)[0]
    if len(y) != n:
        raise ValueError("the number of samples in 'x' and 'y' must match")
    if len(y) == 0:
        raise ValueError("the response (y) is empty.")
    if n == 0:
        raise ValueError("the predictor matrix (x) is empty.")
    if lambda_ is not None and len(lambda_) > 0:
        nlambda = len(lambda_)
    if nlambda == 0:
        raise ValueError("lambda path cannot be of zero length.")
    if alpha < 0 or alpha > 1:
        raise ValueError("elastic net mixing parameter (alpha) must be in [0, 1].")
    if lambda_ is not None and np.any(np.asarray(lambda_) < 0):
        raise ValueError("penalty strengths (lambdas) must be positive.")
    yy = np.asarray(y)
    if np.issubdtype(yy.dtype, np.floating) and np.isnan(yy).any():
        raise ValueError("NA values are not allowed.")
    xv = x.data if _abi.is_sparse(x) and not isinstance(x, _abi.CscMatrix) else (x.x if isinstance(x, _abi.CscMatrix) else np.asarray(x))
    if np.isnan(xv).any():
        raise ValueError("NA values are not allowed.")
    if thresh < 0:
        raise ValueError("threshold for stopping criteria cannot be negative.")
    if maxit <= 0:
        raise ValueError("maximum number of iterations cannot be negative or zero.")
    return nlambda


def build_control(family: str, n_classes: int, *, alpha, nlambda, lambda_min_ratio, lambda_, maxit, standardize,
                  intercept, thresh, standardize_response, debug):
    return _abi.make_control(FAMILIES[family], alpha=alpha, intercept=intercept, standardize=standardize,
                             standardize_response=standardize_response, n_lambda=nlambda, n_classes=n_classes,
                             debug=debug, max_iter=maxit, lambda_min_ratio=lambda_min_ratio, tol=thresh,
                             lambda_=lambda_)


def wrap_fit(raw: _abi.RawFit, family: str, alpha: float, class_names, nobs: int, debug: bool = False) -> SgdnetFit:
    """R/sgdnet.R:368-432: reshape the raw list into the glmnet-like object."""
    L, p, K = raw.beta.shape
    if family in ("gaussian", "binomial"):
        a0 = raw.a0[:, 0].copy()
        beta = raw.beta[:, :, 0].T.copy()                      # (p, n_lambda)
        df = (beta != 0).sum(axis=0)
        dfmat = None
    else:
        a0 = raw.a0.T.copy()                                    # (K, n_lambda)
        beta = [raw.beta[:, :, k].T.copy() for k in range(K)]
        df = (sum(beta) != 0).sum(axis=0)                       # Reduce("+", beta) != 0  (quirk Q10)
        dfmat = np.stack([(np.abs(b) > 0).sum(axis=0) for b in beta])
        if family == "multinomial":
            a0 = a0 - a0.mean(axis=0, keepdims=True)            # R/sgdnet.R:409-410
    fit = SgdnetFit(family=family, a0=a0, beta=beta, lambda_=raw.lambda_.copy(), dev_ratio=raw.dev_ratio.copy(), df=df,
                    nulldev=raw.nulldev, npasses=raw.npasses, alpha=alpha, classnames=class_names, nobs=nobs,
                    grouped=(family == "mgaussian"), dfmat=dfmat, return_codes=raw.return_codes, epochs=raw.epochs,
                    raw=raw)
    if debug:
        fit.diagnostics = {"loss": raw.losses}
    return fit


def sgdnet(x, y, family: str = "gaussian", alpha: float = 1.0, nlambda: int = 100,
           lambda_min_ratio: Optional[float] = None, lambda_: Optional[Sequence[float]] = None, maxit: int = 1000,
           standardize: bool = True, intercept: bool = True, thresh: float = 0.001,
           standardize_response: bool = False, *, seed: Optional[int] = None, rng: Optional[_abi.Rng] = None,
           debug: bool = False, backend: Optional[Library] = None) -> SgdnetFit:
    """sgdnet.default (R/sgdnet.R:183-433). `seed` plays `set.seed(seed)` before the call; `rng` is a live
    R-compatible generator that is advanced in place (R's global RNG)."""
    lib = backend or _abi.product()
    if family not in FAMILIES:
        raise ValueError(f"'arg' should be one of {list(FAMILIES)}")
    if not all(isinstance(v, (bool, np.bool_)) for v in (intercept, standardize, debug)):
        raise TypeError("is.logical(intercept), is.logical(standardize), is.logical(debug) are not all TRUE")
    n, p = _shape(x)
    if lambda_min_ratio is None:
        lambda_min_ratio = 0.01 if n < p else 0.0001
    if lambda_ is False:
        lambda_ = None
    nlambda = validate(x, y, alpha, nlambda, lambda_, maxit, thresh)
    ymat, n_classes, class_names = _encode_response(y, family)
    ctl, keep = build_control(family, n_classes, alpha=alpha, nlambda=nlambda, lambda_min_ratio=lambda_min_ratio,
                              lambda_=lambda_, maxit=maxit, standardize=standardize, intercept=intercept,
                              thresh=thresh, standardize_response=standardize_response, debug=debug)
    if rng is None:
        rng = lib.rng_from_seed(1 if seed is None else seed)
    raw = lib.fit(x, ymat, ctl, rng)
    fit = wrap_fit(raw, family, alpha, class_names, n, debug)
    fit.call = dict(family=family, alpha=alpha, nlambda=nlambda, lambda_min_ratio=lambda_min_ratio, maxit=maxit,
                    standardize=standardize, intercept=intercept, thresh=thresh,
                    standardize_response=standardize_response)
    return fit


# ------------------------------------------------------------------------------------------- predict / coef
def lambda_interpolate(lambda_, s):
    """R/predict.sgdnet.R:144-169."""
    lambda_ = np.asarray(lambda_, dtype=float)
    s = np.atleast_1d(np.asarray(s, dtype=float))
    if len(lambda_) == 1:
        nums = len(s)
        return dict(left=np.zeros(nums, int), right=np.zeros(nums, int), frac=np.ones(nums))
    k = len(lambda_)
    s = np.clip(s, lambda_.min(), lambda_.max())
    sfrac = (lambda_[0] - s) / (lambda_[0] - lambda_[k - 1])
    lam = (lambda_[0] - lambda_) / (lambda_[0] - lambda_[k - 1])
    coord = np.interp(sfrac, lam, np.arange(k))
    left = np.floor(coord).astype(int)
    right = np.ceil(coord).astype(int)
    with np.errstate(invalid="ignore", divide="ignore"):
        frac = (sfrac - lam[right]) / (lam[left] - lam[right])
    frac[left == right] = 1.0
    return dict(left=left, right=right, frac=frac)


def _stack_coef(fit: SgdnetFit) -> List[np.ndarray]:
    """bind_intercept: one (p+1, n_lambda) matrix per class."""
    if fit.family in ("gaussian", "binomial"):
        return [np.vstack([fit.a0[None, :], fit.beta])]
    return [np.vstack([fit.a0[k][None, :], fit.beta[k]]) for k in range(len(fit.beta))]


def coef(fit: SgdnetFit, s=None):
    mats = _stack_coef(fit)
    if s is not None:
        s = np.atleast_1d(np.asarray(s, dtype=float))
        if np.any(s < 0):
            raise ValueError("s (lambda penalty) cannot be negative")
        li = lambda_interpolate(fit.lambda_, s)
        mats = [m[:, li["left"]] * li["frac"] + m[:, li["right"]] * (1 - li["frac"]) for m in mats]
    return mats[0] if fit.family in ("gaussian", "binomial") else mats


def predict(fit: SgdnetFit, newx=None, s=None, type: str = "link", backend: Optional[Library] = None):
    """predict.sgdnet_* (R/predict.sgdnet.R:347-566). The X*beta product runs on the device."""
    allowed = {"gaussian": ("link", "response", "coefficients", "nonzero"),
               "mgaussian": ("link", "response", "coefficients", "nonzero")}.get(
        fit.family, ("link", "response", "coefficients", "nonzero", "class"))
    if type not in allowed:
        raise ValueError(f"'arg' should be one of {allowed}")
    mats = coef(fit, s)
    if type == "coefficients":
        return mats
    mats_l = [mats] if fit.family in ("gaussian", "binomial") else mats
    if type == "nonzero":
        src = mats_l[0] if (fit.family in ("gaussian", "binomial") or fit.grouped) else None
        if src is not None:
            return [np.nonzero(src[1:, l])[0] for l in range(src.shape[1])]
        return [[np.nonzero(m[1:, l])[0] for l in range(m.shape[1])] for m in mats_l]
    if newx is None:
        raise ValueError(f"you need to supply a value for 'newx' for type = '{type}'")
    lib = backend or _abi.product()
    K = len(mats_l)
    L = mats_l[0].shape[1]
    a0 = np.stack([m[0, :] for m in mats_l], axis=1)                    # (L, K)
    beta = np.stack([m[1:, :].T for m in mats_l], axis=2)               # (L, p, K)
    link = lib.predict(newx, a0, beta)                                  # (L, K, n)
    if fit.family in ("gaussian", "binomial"):
        eta = link[:, 0, :].T                                           # (n, L)
        if fit.family == "binomial":
            if type == "response":
                return 1.0 / (1.0 + np.exp(-eta))
            if type == "class":
                return np.asarray(fit.classnames, dtype=object)[(eta > 0).astype(int)]
        return eta
    dp = np.transpose(link, (2, 1, 0))                                  # (n, K, L)
    if fit.family == "mgaussian" or type == "link":
        return dp
    if type == "response":
        pp = np.exp(dp)
        return pp / pp.sum(axis=1, keepdims=True)
    return np.asarray(fit.classnames, dtype=object)[dp.argmax(axis=1)]  # class


def deviance(fit: SgdnetFit):
    """deviance.sgdnet (R/deviance.sgdnet.R:33-41)."""
    return (1.0 - fit.dev_ratio) * fit.nulldev


MEASURES_OF = {"gaussian": ("deviance", "mse", "mae"), "binomial": ("deviance", "mse", "mae", "class", "auc"),
               "multinomial": ("deviance", "mse", "mae", "class"), "mgaussian": ("deviance", "mse", "mae")}


def _encode_for_score(family: str, y):
    if family in ("binomial", "multinomial"):
        levels = np.unique(np.asarray(y).reshape(-1))
        return np.searchsorted(levels, np.asarray(y).reshape(-1)).astype(np.float64).reshape(-1, 1)
    return np.asarray(y, dtype=np.float64).reshape(len(y), -1)


def score(fit: SgdnetFit, x, y, type_measure: str = "deviance", backend: Optional[Library] = None,
          rng: Optional[_abi.Rng] = None):
    """score.sgdnet_* (R/score.R:55-178), every type.measure of the family, on the device: X * beta and the per-sample
    terms in one pass over the rows. `rng` (R's generator) supplies auc's tie-breaking draws (R/score.R:218)."""
    lib = backend or _abi.product()
    if type_measure not in MEASURES_OF[fit.family]:
        raise ValueError(f"'arg' should be one of {MEASURES_OF[fit.family]}")
    mats_l = _stack_coef(fit)
    a0 = np.stack([m[0, :] for m in mats_l], axis=1)
    beta = np.stack([m[1:, :].T for m in mats_l], axis=2)
    yv = _encode_for_score(fit.family, y)
    if lib.has("score_dense"):
        if type_measure == "auc" and rng is None:
            rng = lib.rng_from_seed(1)
        return lib.score(x, yv, FAMILIES[fit.family], type_measure, a0, beta, rng=rng)
    if type_measure != "deviance":
        raise NotImplementedError(f"this backend only scores type.measure = 'deviance'")
    return lib.score_deviance(x, yv, FAMILIES[fit.family], a0, beta)


# ------------------------------------------------------------------------------------------- cv
@dataclass
class CvSgdnet:
    alpha: list
    lambda_: list
    cv_raw: list              # per alpha: (nfolds, n_lambda)
    cv_summary: np.ndarray    # columns alpha, lambda, mean, sd, ci_lo, ci_up
    fit: SgdnetFit
    alpha_min: float
    lambda_min: float
    lambda_1se: float
    name: str
    fits: list = field(default_factory=list)
    fold_fits: list = field(default_factory=list)
    owned_full: list = field(default_factory=list)     # alphas whose full-data fit THIS rank ran (sharded runs)


def _summarize(cv_raw):
    """summarize_cv_raw (R/cv_sgdnet.R:293-299) with col_sd = n-1 sample sd (R/utils.R:38-46)."""
    m = cv_raw.mean(axis=0)
    sd = cv_raw.std(axis=0, ddof=1)
    return np.stack([m, sd, m - sd, m + sd], axis=1)


def _find_optimum(block):
    """find_optimum (R/cv_sgdnet.R:265-282): lambda_1se uses mean + 1*sd (quirk Q11)."""
    i = int(np.argmin(block[:, 2]))
    within = block[:, 2] <= block[i, 2] + block[i, 3]
    return dict(alpha_min=block[i, 0], lambda_min=block[i, 1], lambda_1se=block[within, 1].max(), error_min=block[i, 2])


def make_foldid(n: int, nfolds: int, perm: np.ndarray) -> np.ndarray:
    """as.numeric(cut(perm, nfolds)) for a permutation `perm` of 1..n (R/cv_sgdnet.R:169)."""
    lo, hi = 1.0, float(n)
    width = (hi - lo) / nfolds
    breaks = lo + width * np.arange(nfolds + 1)
    breaks[0] = lo - (hi - lo) / 1000.0
    breaks[-1] = hi + (hi - lo) / 1000.0
    return np.searchsorted(breaks, perm, side="left").astype(int)     # (a, b] intervals -> 1..nfolds


def cv_plan(n: int, alphas: Sequence[float], foldid: np.ndarray):
    """The (alpha, fold) work list of R/cv_sgdnet.R:178-200: fold j TRAINS on foldid == j."""
    folds = np.unique(foldid)
    # one row-id array per fold, shared by every alpha: the backend keys its prepared designs on it
    split = {int(j): (np.ascontiguousarray(np.nonzero(foldid == j)[0], dtype=np.int32),
                      np.ascontiguousarray(np.nonzero(foldid != j)[0], dtype=np.int32)) for j in folds}
    plan = []
    for i, a in enumerate(alphas):
        for j in folds:
            tr, te = split[int(j)]
            plan.append(dict(alpha_index=i, alpha=a, fold=int(j), train_rows=tr, test_rows=te))
    return plan


def cv_sgdnet(x, y, alpha=1.0, lambda_=None, nfolds: int = 10, foldid=None, type_measure: str = "deviance", *,
              family: str = "gaussian", seed: int = 1, fit_seeds: Optional[Sequence[int]] = None,
              backend: Optional[Library] = None, batched: bool = True, shard=None, **kwargs) -> CvSgdnet:
    """cv_sgdnet (R/cv_sgdnet.R:113-254).

    Deviations, both documented in DESIGN.md: sparse x stays sparse (the reference densifies it,
    :130, quirk Q12), and every fit gets its own generator `set.seed(fit_seeds[k])` (default
    seed + k, full fits first) so that folds can run concurrently. `shard` (a
    `sgdnet_b200.shard.Shard`) deals ALL the fits - full-data and fold - to the ranks longest-first; one all_gather of
    the per-fit score rows ends the run, plus a broadcast of the selected alpha's full fit from the rank that owns it.
    `fits[i]` is None on ranks that do not own alpha i's full fit.
    """
    lib = backend or _abi.product()
    if type_measure not in MEASURES_OF[family]:
        raise ValueError(f"'arg' should be one of {MEASURES_OF[family]}")
    if type_measure == "auc":
        batched = False          # auc is scored one fit at a time: its tie-breaking draws come from the caller's generator
    alphas = list(np.atleast_1d(alpha))
    if not (nfolds > 2 and len(alphas) > 0):
        raise ValueError("nfolds > 2, is.numeric(alpha), length(alpha) > 0 are not all TRUE")
    n, p = _shape(x)
    if nfolds > n:
        raise ValueError("you cannot have more folds than samples.")
    if isinstance(lambda_, list) and len(lambda_) > 0 and isinstance(lambda_[0], (list, np.ndarray)):
        if len(lambda_) != len(alphas):
            raise ValueError("the length of the lambda list needs to match the number of alpha.")
        lambdas = lambda_
    elif lambda_ is not None:
        if len(alphas) > 1:
            raise ValueError("you need a list of lambdas (or have it set at NULL) when you have multiple alphas.")
        lambdas = [lambda_]
    else:
        lambdas = [None] * len(alphas)
    n_full = len(alphas)
    if foldid is None:
        perm = np.random.Generator(np.random.PCG64(seed)).permutation(n) + 1
        foldid = make_foldid(n, nfolds, perm)
    else:
        foldid = np.asarray(foldid)
        if len(foldid) != n:
            raise ValueError("the length of `foldid` must match the number of samples")
    plan = cv_plan(n, alphas, foldid)
    folds = np.unique(foldid)
    n_fits = n_full + len(plan)
    if fit_seeds is None:
        fit_seeds = [seed + k for k in range(n_fits)]

    ymat, n_classes, class_names = _encode_response(y, family)
    opts = dict(nlambda=kwargs.get("nlambda", 100), lambda_min_ratio=kwargs.get("lambda_min_ratio"),
                maxit=kwargs.get("maxit", 1000), standardize=kwargs.get("standardize", True),
                intercept=kwargs.get("intercept", True), thresh=kwargs.get("thresh", 0.001),
                standardize_response=kwargs.get("standardize_response", False))

    if shard is None:
        from .shard import Shard
        shard = Shard()
    # Every fit of the call - the #alpha full-data fits and the #alpha x #folds fold fits - is dealt to the ranks
    # longest-first by its number of training rows (R/cv_sgdnet.R:160-200 runs them one after the other). A fold fit
    # needs its alpha's lambda path, which comes from the full fit's SETUP, not its solution (src/utils.h:157-165), so a
    # rank that does not own an alpha's full fit asks for that setup only (`path_only`).
    plan_costs = [float(len(w["train_rows"])) for w in plan]
    use_batch = batched and lib.has("fit_batch_dense")
    if use_batch:
        all_costs = [float(n)] * n_full + plan_costs
        owned = shard.mine(all_costs)
        my_full = [k for k in owned if k < n_full]
        mine = [k - n_full for k in owned if k >= n_full]
    else:      # one call per fit: the full fits run on every rank (each needs their lambda paths), the fold fits are dealt
        my_full = list(range(n_full))
        mine = shard.mine(plan_costs)
    cv_rows = np.full((len(plan), opts["nlambda"] if lambdas[0] is None else max(len(l) for l in lambdas)), np.nan)
    fold_fits = [None] * len(plan)
    fits = [None] * n_full

    def control_for(a, lam, n_rows):
        lmr = opts["lambda_min_ratio"] if opts["lambda_min_ratio"] is not None else (0.01 if n_rows < p else 0.0001)
        return build_control(family, n_classes, alpha=a, nlambda=opts["nlambda"] if lam is None else len(lam),
                             lambda_min_ratio=lmr, lambda_=lam, maxit=opts["maxit"], standardize=opts["standardize"],
                             intercept=opts["intercept"], thresh=opts["thresh"],
                             standardize_response=opts["standardize_response"], debug=False)

    if use_batch:
        # ONE batch per rank: its full-data fits (or just their lambda paths) and its fold fits, which take that path
        # through `lambda_from` - it is known after the full fit's setup, so all the fits run concurrently, each its
        # own pipeline on the device
        specs, keeps = [], []
        for i, a in enumerate(alphas):
            ctl, keep = control_for(a, lambdas[i], n)
            keeps.append(keep)
            specs.append(dict(train_rows=None, test_rows=None, control=ctl, rng=lib.rng_from_seed(fit_seeds[i]),
                              path_only=i not in my_full))
        for k in mine:
            w = plan[k]
            ctl, keep = control_for(w["alpha"], lambdas[w["alpha_index"]], len(w["train_rows"]))
            keeps.append(keep)
            specs.append(dict(train_rows=w["train_rows"], test_rows=w["test_rows"], control=ctl,
                              rng=lib.rng_from_seed(fit_seeds[n_full + k]), measure=_abi.MEASURES[type_measure],
                              lambda_from=w["alpha_index"] if lambdas[w["alpha_index"]] is None else -1))
        raws, scores = lib.fit_batch(x, ymat, specs)
        lambdas = [raw.lambda_.copy() for raw in raws[:n_full]]
        for i in my_full:
            fits[i] = wrap_fit(raws[i], family, alphas[i], class_names, n)
        for k, raw, sc in zip(mine, raws[n_full:], scores[n_full:]):
            fold_fits[k] = wrap_fit(raw, family, plan[k]["alpha"], class_names, len(plan[k]["train_rows"]))
            cv_rows[k, :len(raw.lambda_)] = sc[:len(raw.lambda_)]
    else:
        # one call per fit (a backend without the batch entry point: the CPU oracle in the tests); unsharded full fits
        fits = [sgdnet(x, y, family=family, alpha=a, lambda_=lambdas[i], seed=fit_seeds[i], backend=lib, **opts)
                for i, a in enumerate(alphas)]
        lambdas = [f.lambda_ for f in fits]
        cv_rows = np.full((len(plan), max(len(l) for l in lambdas)), np.nan)
        yarr = np.asarray(y)
        for k in mine:
            w = plan[k]
            f = sgdnet(_rows(x, w["train_rows"]), yarr[w["train_rows"]], family=family, alpha=w["alpha"],
                       lambda_=lambdas[w["alpha_index"]], seed=fit_seeds[n_full + k], backend=lib, **opts)
            fold_fits[k] = f
            cv_rows[k, :len(f.lambda_)] = score(f, _rows(x, w["test_rows"]), yarr[w["test_rows"]], type_measure, backend=lib)
    cv_rows = shard.all_gather_rows(cv_rows, mine)

    cv_raw = []
    for i in range(len(alphas)):
        rows = [k for k in range(len(plan)) if plan[k]["alpha_index"] == i]
        cv_raw.append(cv_rows[rows, :len(lambdas[i])])
    blocks = []
    for i, a in enumerate(alphas):
        blocks.append(np.column_stack([np.full(len(lambdas[i]), a), lambdas[i], _summarize(cv_raw[i])]))
    optima = [_find_optimum(b) for b in blocks]
    best = int(np.argmin([o["error_min"] for o in optima]))
    if shard.world > 1 and use_batch:     # the model the call returns lives on the rank that fitted it
        owner = int(shard.owner_of(all_costs)[best])
        fits[best] = shard.broadcast_object(fits[best], src=owner)
    name = {"deviance": {"gaussian": "Mean-Squared Error", "mgaussian": "Mean-Squared Error", "binomial": "Binomial Deviance",
                         "multinomial": "Multnomial Deviance"}[family],
            "mse": "Mean-Squared Error", "mae": "Mean Absolute Error", "class": "Misclassification Error", "auc": "AUC"}[type_measure]
    return CvSgdnet(alpha=alphas, lambda_=lambdas, cv_raw=cv_raw, cv_summary=np.vstack(blocks), fit=fits[best],
                    alpha_min=optima[best]["alpha_min"], lambda_min=optima[best]["lambda_min"],
                    lambda_1se=optima[best]["lambda_1se"], name=name, fits=fits, fold_fits=fold_fits,
                    owned_full=list(my_full))
