"""score() and predict() of the reference, rendered line by line in numpy with R's array semantics.

TEST INFRASTRUCTURE. This file is the independent pin of SURVEY.md section 8a rows 14 / 15: it follows the R SOURCE
(R/score.R:55-232, R/predict.sgdnet.R:104-128, 347-566) statement by statement with R's own array operations - column-major
arrays, `diag(K)[as.numeric(y), ]`, `pmin/pmax`, `colMeans(apply(a, 3, rowSums))`, `array(y, dim(y_hat))` recycling,
`as.numeric(as.factor(...))` - and shares no code with oracle/sgdnet_oracle.cpp or the CUDA library; both of those are
checked against it (tests/test_score_cpu.py, tests/test_parity_gpu.py) and against the fixture it produced
(tests/golden/score_fixture.npz, made by tests/golden/make_score_fixture.py).

Conventions: a "fit" is (family, a0, beta): a0 (n_lambda,) and beta (p, n_lambda) for gaussian / binomial; a0
(K, n_lambda) and beta a list of K (p, n_lambda) arrays otherwise - the shapes R's object has (R/sgdnet.R:368-410).
"""
import numpy as np
import scipy.sparse as sp


# --------------------------------------------------------------------------------------------- small pieces of R
def as_factor_codes(y):
    """as.numeric(as.factor(y)) - 1, and the levels (sorted unique values, as R sorts a numeric / character vector)."""
    levels, codes = np.unique(np.asarray(y).reshape(-1), return_inverse=True)
    return codes, levels


def cbind2_one_times(x, b):
    """as.matrix(cbind2(1, newx) %*% beta)  (R/predict.sgdnet.R:377, 510): beta is (p + 1, n_lambda), row 0 the intercept."""
    n = x.shape[0]
    ones = np.ones((n, 1))
    if sp.issparse(x):
        xx = sp.hstack([sp.csc_matrix(ones), sp.csc_matrix(x)], format="csr")
        return np.asarray(xx @ b)
    return np.hstack([ones, np.asarray(x, dtype=float)]) @ b


def bind_intercept(beta, a0):
    """rbind(a0, beta) per class."""
    if isinstance(beta, list):
        return [np.vstack([a0[k][None, :], beta[k]]) for k in range(len(beta))]
    return np.vstack([np.asarray(a0)[None, :], beta])


def softmax_class(x):
    """softmax() of R/predict.sgdnet.R:104-128 on an (n, K) matrix: 1-based class with the largest value, FIRST wins ties
    (`l <- x[, i] > maxdist`)."""
    maxdist = x[:, 0].copy()
    pclass = np.ones(x.shape[0], dtype=int)
    for i in range(1, x.shape[1]):
        l = x[:, i] > maxdist
        pclass[l] = i + 1
        maxdist[l] = x[l, i]
    return pclass


# --------------------------------------------------------------------------------------------- predict
def predict_link(family, a0, beta, newx):
    """type = "link". gaussian / binomial: (n, n_lambda). multinomial / mgaussian: (n, K, n_lambda) = aperm(dp, c(3,1,2))."""
    nb = bind_intercept(beta, a0)
    if family in ("gaussian", "binomial"):
        return cbind2_one_times(newx, nb)
    K, L, n = len(nb), nb[0].shape[1], newx.shape[0]
    dp = np.zeros((K, L, n))
    for i in range(K):                                   # for (i in seq(nclass)) dp[i, , ] <- t(fitk)
        dp[i, :, :] = dp[i, :, :] + cbind2_one_times(newx, nb[i]).T
    return np.transpose(dp, (2, 0, 1))


def predict_response(family, a0, beta, newx):
    if family == "gaussian" or family == "mgaussian":
        return predict_link(family, a0, beta, newx)
    if family == "binomial":
        return 1.0 / (1.0 + np.exp(-predict_link(family, a0, beta, newx)))     # R/predict.sgdnet.R:437
    dp = np.transpose(predict_link(family, a0, beta, newx), (1, 2, 0))          # (K, L, n) as in R
    pp = np.exp(dp)
    psum = pp.sum(axis=0)                                                       # apply(pp, c(2, 3), sum)
    return np.transpose(pp / psum[None, :, :], (2, 0, 1))                       # aperm(..., c(3, 1, 2))


# --------------------------------------------------------------------------------------------- auc (R/score.R:203-232)
def auc_r(y2, prob, runif):
    """auc(y, prob) with y the n x 2 indicator matrix: the matrix branch doubles the data - rep(c(0, 1), c(ny, ny)),
    c(prob, prob), weights as.vector(1 * y) - and calls the weighted branch, which breaks ties in `prob` with
    stats::runif (one draw per doubled observation, from R's generator: `runif(k)` returns the next k uniforms)."""
    ny = y2.shape[0]
    y = np.repeat([0, 1], [ny, ny])
    prob = np.concatenate([prob, prob])
    weights = (1.0 * y2).reshape(-1, order="F")          # as.vector: column-major
    rprob = np.asarray(runif(len(prob)))
    op = np.lexsort((rprob, prob))                       # order(prob, rprob); ties of both keep their original order
    y = y[op]
    weights = weights[op]
    cw = np.cumsum(weights)
    w1 = weights[y == 1]
    cw1 = np.cumsum(w1)
    wauc = np.log(np.sum(w1 * (cw[y == 1] - cw1)))
    sumw1 = cw1[-1]
    sumw2 = cw[-1] - sumw1
    return np.exp(wauc - np.log(sumw1) - np.log(sumw2))


# --------------------------------------------------------------------------------------------- score (R/score.R:55-178)
def score(family, a0, beta, x, y, type_measure="deviance", runif=None):
    prob_min = 1e-05
    prob_max = 1 - prob_min
    if family == "gaussian":
        y = np.asarray(y, dtype=float).reshape(-1)                      # as.vector(y)
        y_hat = predict_link(family, a0, beta, x)
        d = y_hat - y[:, None]
        return {"deviance": (d ** 2).mean(axis=0), "mse": (d ** 2).mean(axis=0), "mae": np.abs(d).mean(axis=0)}[type_measure]

    if family == "binomial":
        codes, _ = as_factor_codes(y)
        y = np.eye(2)[codes, :]                                         # diag(2)[as.numeric(y), ]
        y_hat = predict_response(family, a0, beta, x)
        if type_measure == "auc":
            return np.array([auc_r(y, y_hat[:, i], runif) for i in range(y_hat.shape[1])])
        if type_measure == "mse":
            return ((y_hat + y[:, [0]] - 1) ** 2 + (y_hat - y[:, [1]]) ** 2).mean(axis=0)
        if type_measure == "mae":
            return (np.abs(y_hat + y[:, [0]] - 1) + np.abs(y_hat - y[:, [1]])).mean(axis=0)
        if type_measure == "deviance":
            y_hat = np.minimum(np.maximum(y_hat, prob_min), prob_max)
            lp = y[:, [0]] * np.log(1 - y_hat) + y[:, [1]] * np.log(y_hat)
            with np.errstate(divide="ignore"):
                ly = np.log(y)
            ly[y == 0] = 0
            ly = (y * ly) @ np.array([1.0, 1.0])
            return (2 * (ly[:, None] - lp)).mean(axis=0)
        if type_measure == "class":
            return (y[:, [0]] * (y_hat > 0.5) + y[:, [1]] * (y_hat <= 0.5)).mean(axis=0)
        raise ValueError(type_measure)

    if family == "multinomial":
        codes, levels = as_factor_codes(y)
        n_classes = len(levels)
        y1 = np.eye(n_classes)[codes, :]
        y_hat = predict_response(family, a0, beta, x)                   # (n, K, L)
        L = y_hat.shape[2]
        yy = np.resize(y1.reshape(-1, order="F"), y_hat.size).reshape(y_hat.shape, order="F")   # array(y, dim(y_hat))
        apply3_rowsums = lambda a: a.sum(axis=1)                        # apply(a, 3, rowSums): (n, L)
        if type_measure == "mse":
            return apply3_rowsums((yy - y_hat) ** 2).mean(axis=0)
        if type_measure == "mae":
            return apply3_rowsums(np.abs(yy - y_hat)).mean(axis=0)
        if type_measure == "deviance":
            y_hat = np.minimum(np.maximum(y_hat, prob_min), prob_max)
            lp = yy * np.log(y_hat)
            with np.errstate(divide="ignore", invalid="ignore"):
                ly = yy * np.log(yy)
            ly[yy == 0] = 0
            return apply3_rowsums(2 * (ly - lp)).mean(axis=0)
        if type_measure == "class":
            # classid <- as.numeric(as.factor(apply(y_hat, 3, softmax))): the predicted classes of ALL samples and ALL
            # lambdas go through as.factor, so the ids are RANKS among the classes that were predicted at least once -
            # when some class is never predicted they are not class numbers (kept as the reference has it)
            pred = np.stack([softmax_class(y_hat[:, :, l]) for l in range(L)], axis=1)      # (n, L)
            flat = pred.reshape(-1, order="F")
            classid, _ = as_factor_codes(flat)
            classid = classid + 1
            yperm = np.transpose(yy, (0, 2, 1)).reshape(-1, n_classes, order="F")           # matrix(aperm(y, c(1,3,2)), ncol = K)
            picked = yperm[np.arange(len(classid)), classid - 1]
            return (1 - picked).reshape(-1, L, order="F").mean(axis=0)
        raise ValueError(type_measure)

    if family == "mgaussian":
        y_hat = predict_link(family, a0, beta, x)                       # (n, K, L)
        y = np.asarray(y, dtype=float)
        yy = np.resize(y.reshape(-1, order="F"), y_hat.size).reshape(y_hat.shape, order="F")
        apply3_colsums = lambda a: a.sum(axis=0)                        # apply(a, 3, colSums): (K, L)  [sic: colSums]
        if type_measure in ("deviance", "mse"):
            return apply3_colsums((y_hat - yy) ** 2).mean(axis=0)       # colMeans over the K responses of sums over samples
        if type_measure == "mae":
            return apply3_colsums(np.abs(y_hat - yy)).mean(axis=0)
        raise ValueError(type_measure)
    raise ValueError(family)


def fit_to_r_shapes(family, a0_raw, beta_raw):
    """(n_lambda, K) / (n_lambda, p, K) raw arrays -> the shapes of R's object. Multinomial intercepts are centred
    (R/sgdnet.R:409-410) by the caller if wanted; score() is invariant to it except through rounding."""
    if family in ("gaussian", "binomial"):
        return a0_raw[:, 0].copy(), beta_raw[:, :, 0].T.copy()
    K = beta_raw.shape[2]
    return a0_raw.T.copy(), [beta_raw[:, :, k].T.copy() for k in range(K)]
