"""Pins prediction and scoring to R/score.R and R/predict.sgdnet.R through a
rendering that is independent of oracle/sgdnet_oracle.cpp: tests/r_score.py (numpy, R's array operations statement by
statement) and the fixture it produced from the reference build's coefficients (tests/golden/score_fixture.npz)."""
import numpy as np
import pytest

import r_score
from r_score_rng import RUnif
from score_cases import AUC_SEED, SCORE_CASES, check_backend_against_fixture, fixture, held_out


@pytest.mark.parametrize("name", list(SCORE_CASES))
def test_numpy_rendering_reproduces_the_committed_fixture(name):
    fx = fixture()
    xs, ys, family, a0, beta, measures = held_out(name)
    np.testing.assert_allclose(r_score.predict_link(family, a0, beta, xs), fx[f"{name}/link"], rtol=1e-13, atol=0)
    for m in measures:
        sc = r_score.score(family, a0, beta, xs, ys, m, runif=RUnif(AUC_SEED) if m == "auc" else None)
        np.testing.assert_allclose(sc, fx[f"{name}/{m}"], rtol=1e-13, atol=0, err_msg=f"{name} {m}")


@pytest.mark.parametrize("name", list(SCORE_CASES))
def test_oracle_predict_and_score_match_the_r_rendering(oracle, name):
    done = check_backend_against_fixture(oracle, name)
    assert set(done) == set(SCORE_CASES[name][2])


def test_r_rendering_closed_forms():
    """Sanity of the rendering itself on hand-computable inputs."""
    x = np.array([[1.0, 0.0], [0.0, 2.0], [1.0, 1.0], [2.0, 1.0]])
    a0 = np.array([0.5, 0.0])
    beta = np.array([[1.0, 0.0], [-1.0, 0.0]])          # (p, n_lambda)
    link = r_score.predict_link("gaussian", a0, beta, x)
    np.testing.assert_allclose(link[:, 0], 0.5 + x[:, 0] - x[:, 1])
    np.testing.assert_allclose(link[:, 1], 0.0)
    y = np.array([1.0, -1.0, 0.0, 2.0])
    np.testing.assert_allclose(r_score.score("gaussian", a0, beta, x, y, "mse")[1], np.mean(y ** 2))
    yb = np.array([0, 1, 1, 0])
    # all-zero model: p = 1/2 everywhere -> deviance 2 log 2, misclassification = share of the second class (p <= 0.5)
    np.testing.assert_allclose(r_score.score("binomial", a0, beta, x, yb, "deviance")[1], 2 * np.log(2))
    np.testing.assert_allclose(r_score.score("binomial", a0, beta, x, yb, "class")[1], 0.5)
    # a perfectly separating score gives auc 1, its negation 0
    bsep = np.array([[3.0, -3.0], [0.0, 0.0]])
    with np.errstate(divide="ignore"):
        auc = r_score.score("binomial", np.zeros(2), bsep, x[[0, 1, 3]], np.array([0, 0, 1]), "auc", runif=RUnif(1))
    np.testing.assert_allclose(auc, [1.0, 0.0])


def test_class_measure_when_a_class_is_never_predicted(oracle):
    """as.numeric(as.factor(predicted classes)) numbers the classes that occur, so with a class that no sample at no
    lambda is assigned to, the ids are ranks and not class numbers (R/score.R:153): rendering and backend agree on it."""
    rng = np.random.default_rng(3)
    n, p, K, L = 60, 4, 3, 3
    x = rng.normal(size=(n, p))
    y = np.arange(n) % K
    a0 = np.zeros((K, L))
    a0[1, :] = -50.0                      # class 2 of 3 is never the most probable one
    beta = [rng.normal(size=(p, L)) * (0.0 if k == 1 else 1.0) for k in range(K)]
    want = r_score.score("multinomial", a0, beta, x, y, "class")
    from score_cases import raw_coefficients
    a0r, br = raw_coefficients("multinomial", a0, beta)
    got = oracle.score(x, y.astype(float).reshape(-1, 1), 2, "class", a0r, br)
    np.testing.assert_allclose(got, want, rtol=1e-12)
    plain = np.array([np.mean(np.argmax(r_score.predict_response("multinomial", a0, beta, x)[:, :, l], axis=1) != y) for l in range(L)])
    assert not np.allclose(want, plain)   # the quirk is visible: it is not the plain misclassification rate
