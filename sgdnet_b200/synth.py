"""Seeded synthetic inputs shared by tests, smoke() and bench.py (BASELINE configs 2-5 generators)."""
import numpy as np
import scipy.sparse as sp


def sparse_rows(n, p, nnz_row, seed, normalise=True):
    """CSR n x p, exactly nnz_row distinct sorted column ids per row, values U(0.05, 1.05), rows L2-normalised."""
    rng = np.random.Generator(np.random.PCG64(seed))
    # distinct ids per row: sample with replacement in blocks and repair the few duplicates
    cols = rng.integers(0, p, size=(n, nnz_row), dtype=np.int64)
    cols.sort(axis=1)
    for _ in range(64):
        dup = np.zeros_like(cols, dtype=bool)
        dup[:, 1:] = cols[:, 1:] == cols[:, :-1]
        nd = int(dup.sum())
        if nd == 0:
            break
        cols[dup] = rng.integers(0, p, size=nd, dtype=np.int64)
        cols.sort(axis=1)
    vals = rng.uniform(0.05, 1.05, size=(n, nnz_row))
    if normalise:
        vals /= np.sqrt((vals ** 2).sum(axis=1, keepdims=True))
    indptr = np.arange(0, (n + 1) * nnz_row, nnz_row, dtype=np.int64)
    m = sp.csr_matrix((vals.reshape(-1), cols.reshape(-1).astype(np.int32), indptr), shape=(n, p))
    return m, rng


def binomial_sparse(n, p, nnz_row, seed):
    """BASELINE config 2/5 shape: y ~ Bernoulli(sigmoid(x beta* - 0.2)), beta* 1% nonzero ~ N(0, 3^2)."""
    x, rng = sparse_rows(n, p, nnz_row, seed)
    beta = np.zeros(p)
    nz = rng.choice(p, size=max(1, p // 100), replace=False)
    beta[nz] = rng.normal(0.0, 3.0, size=nz.size)
    z = x @ beta - 0.2
    y = (rng.uniform(size=n) < 1.0 / (1.0 + np.exp(-z))).astype(np.float64)
    return x.tocsc(), y


def multinomial_dense(n, p, K, seed):
    """BASELINE config 3 shape: relu(N(0,1)) * Bernoulli(0.19) scaled to [0,1]; y = argmax(X B* + Gumbel)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    x = np.maximum(rng.normal(size=(n, p)), 0.0) * (rng.uniform(size=(n, p)) < 0.19)
    x /= max(x.max(), 1e-12)
    B = rng.normal(size=(p, K)) / np.sqrt(p)
    y = np.argmax(x @ B + rng.gumbel(size=(n, K)), axis=1)
    return np.asfortranarray(x), y


def mgaussian_dense(n, p, K, seed):
    """BASELINE config 4 shape: X ~ N(0,1); Y = X B* + N(0,1), B* with 5% nonzero rows."""
    rng = np.random.Generator(np.random.PCG64(seed))
    x = rng.normal(size=(n, p))
    B = np.zeros((p, K))
    nz = rng.choice(p, size=max(1, p // 20), replace=False)
    B[nz] = rng.normal(size=(nz.size, K))
    y = x @ B + rng.normal(size=(n, K))
    return np.asfortranarray(x), y


def random_data(n=100, p=2, family="gaussian", intercept=True, density=0.5, seed=0):
    """Python rendering of tests/testthat/setup.R:6-54 `random_data` (numpy RNG instead of R's)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    x = rng.normal(0, 0.01, size=(n, p))
    kill = rng.choice(n * p, size=int((1 - density) * n * p), replace=False)
    x.reshape(-1)[kill] = 0.0
    k = 3 if family in ("multinomial", "mgaussian") else 1
    grid = np.linspace(-1, 1, k * p * 10)
    beta = rng.choice(grid, size=(k, p))
    if intercept:
        beta = np.hstack([rng.choice(grid, size=(k, 1)), beta])
    center = rng.choice(np.linspace(-0.1, 0.1, p * 10), size=p, replace=False)
    scale = rng.choice(np.linspace(1.05, 0.95, p * 10), size=p, replace=False)
    x = (x + center) * scale
    xx = np.hstack([np.ones((n, 1)), x]) if intercept else x
    z = xx @ beta.T
    if family == "gaussian":
        y = rng.normal(z[:, 0], 0.01)
    elif family == "binomial":
        y = (rng.uniform(size=n) < 1 / (1 + np.exp(-z[:, 0]))).astype(float)
    elif family == "multinomial":
        pr = np.exp(z) / np.exp(z).sum(axis=1, keepdims=True)
        y = np.array([rng.choice(k, p=pr[i]) for i in range(n)])
    else:
        y = rng.normal(z, 0.01)
    return sp.csc_matrix(x), y
