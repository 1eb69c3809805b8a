"""CPU tests: the oracle (oracle/tls_oracle.py) against every reference-authored fixture for the hot path
 and against its own committed regression vectors."""
import json
import os
import warnings

import numpy as np
import pytest

import tls_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
SQRT_EPS = np.sqrt(np.finfo(float).eps)


def golden5():
    g = json.load(open(os.path.join(HERE, "golden", "rpca_5x5.json")))
    return np.array(g["D"]), np.array(g["A"]), np.array(g["E"]), g["atol"]


def test_rpca_known_answer_5x5():
    """test/runtests.jl:143-165"""
    D, A, E, atol = golden5()
    r = O.rpca(D, nonnegE=True, nonnegA=True)
    assert np.abs(r.A - A).max() < atol
    assert np.abs(r.E - E).max() < atol
    assert np.linalg.norm(D - (r.A + r.E)) / np.linalg.norm(D) < SQRT_EPS


def test_rpca_residual_no_clamps():
    """test/runtests.jl:168-169"""
    D, _, _, _ = golden5()
    r = O.rpca(D)
    assert np.linalg.norm(D - (r.A + r.E)) / np.linalg.norm(D) < SQRT_EPS


def test_hankel_exact():
    """test/runtests.jl:293-294"""
    x = np.arange(1, 21)
    assert np.array_equal(O.hankel(x, 2), np.stack([np.arange(1, 20), np.arange(2, 21)], axis=1))
    assert np.array_equal(O.hankel(x, 3, 2), np.stack([np.arange(1, 18, 2), np.arange(2, 19, 2), np.arange(3, 20, 2)], axis=1))
    with pytest.raises(AssertionError):
        O.hankel(x, 11)          # L <= N/2
    with pytest.raises(AssertionError):
        O.hankel(x, 2, 3)        # lag <= L


def test_ishankel_and_soft_hankel():
    """test/runtests.jl:296-307"""
    rng = np.random.default_rng(0)
    A = O.hankel(np.arange(1.0, 9.0), 4)
    assert O.ishankel(A)
    An = A + 0.1 * rng.standard_normal(A.shape)
    assert not O.ishankel(An)
    Anc = An.copy()
    O.soft_hankel(An, 0.1)
    assert np.sum((An - A) ** 2) < np.sum((Anc - A) ** 2)


def test_unhankel_roundtrips():
    """test/runtests.jl:356-376"""
    T = 1000
    y = np.sin(0.1 * np.arange(1, T + 1))
    y = y / np.quantile(np.abs(y), 0.9)
    H = O.hankel(y, 2)
    assert O.ishankel(H)
    assert np.array_equal(O.unhankel(H), y)
    assert np.array_equal(O.unhankel(O.hankel(y, 2, 2), 2, T), y)
    yh = O.unhankel(O.hankel(y, 5, 2), 2, T)
    assert np.allclose(yh[:-1], y[:-1])
    y2 = np.random.default_rng(1).standard_normal(T)
    Y = np.stack([y, y2], axis=1)
    yh = O.unhankel(O.hankel(Y, 5, 2), 2, T, 2)
    assert np.allclose(yh[:-1], Y[:-1])


def test_mu_is_weighted_mean():
    """test/runtests.jl:469-480"""
    rng = np.random.default_rng(2)
    U = rng.standard_normal((10, 10))
    s = np.zeros(10)
    assert np.allclose(O.mu_mean(s, np.ones(10), U), U.mean(axis=1))
    w = rng.standard_normal(10)
    assert np.allclose(O.mu_mean(s, w, U), (U * w).sum(axis=1) / w.sum())


@pytest.mark.parametrize("shape", [(10, 40), (40, 10)])
def test_rpca_ga_orthonormal(shape):
    """test/runtests.jl:447-464 (a subset of the 400 cases)"""
    rng = np.random.default_rng(3)
    for r in (1, 3, 6, 10):
        for eps in (1e-8, 1e-4, 1.0):
            U, S, Vt = np.linalg.svd(rng.standard_normal(shape), full_matrices=False)
            A = (U[:, :r] * (10.0 * np.arange(1, r + 1))) @ Vt[:r] + eps * rng.standard_normal(shape)
            Q = O.rpca_ga(A, r, q0=rng.standard_normal((shape[0], r)))
            assert np.linalg.norm(Q.T @ Q - np.eye(r)) < SQRT_EPS


def test_lowrankfilter_missing_values():
    """test/runtests.jl:172-185 (README.md:85-92): mean normalised MSE < 0.025 over 20 draws"""
    rng = np.random.default_rng(4)
    res = []
    for _ in range(20):
        N = 500
        y = np.sin(0.1 * np.arange(1, N + 1)) + 0.1 * rng.standard_normal(N)
        yn = y + (rng.random(N) < 0.1) * 1e2
        yf = O.lowrankfilter(yn, 40)
        res.append(np.mean((y - yf) ** 2) / np.mean(y ** 2))
    assert np.mean(res) < 0.025


def test_lowrankfilter_noise_ratio():
    """test/runtests.jl:378-380"""
    rng = np.random.default_rng(5)
    T = 1000
    qn = lambda x: x / np.quantile(np.abs(x), 0.9)
    y = qn(np.sin(0.1 * np.arange(1, T + 1)))
    n = 20 * rng.standard_normal(T) * (rng.random(T) < 0.01) + 0.1 * rng.standard_normal(T)
    yf = qn(O.lowrankfilter(y + n))
    assert np.mean((y - yf) ** 2) / np.mean(n ** 2) < 0.001


def test_readme_imputation():
    """README.md:94-106: rpca(Hn) recovers H (MSE ~0.06) and E correlates with the missing mask (~1.00)"""
    rng = np.random.default_rng(6)
    N = 500
    H = O.hankel(np.sin(0.1 * np.arange(1, N + 1)), 5)
    miss = rng.random(H.shape) < 0.1
    Hn = H + 0.1 * rng.standard_normal(H.shape) + miss * 1e2
    r = O.rpca(Hn)
    assert np.mean((H - r.A) ** 2) / np.mean(H ** 2) < 0.12
    corr = (r.E.ravel() @ miss.ravel()) / (np.linalg.norm(r.E) * np.linalg.norm(miss))
    assert corr > 0.99


def test_oracle_regression_vectors():
    """The committed vectors (tests/golden/oracle_vectors.npz) are what the CUDA path is checked against on the
    GPU box; make sure the oracle still reproduces them."""
    g = np.load(os.path.join(HERE, "golden", "oracle_vectors.npz"))
    warnings.simplefilter("ignore")
    D = g["rpca_D"]
    r = O.rpca(D, iters=10, tol=0.0)
    assert np.allclose(r.A, g["rpca_plain_A"], rtol=0, atol=1e-12) and np.allclose(r.E, g["rpca_plain_E"], rtol=0, atol=1e-12)
    r = O.rpca(np.abs(D), iters=10, tol=0.0, nonnegA=True, nonnegE=True)
    assert np.allclose(r.A, g["rpca_nonneg_A"], rtol=0, atol=1e-12)
    r = O.rpca(g["rpca_wide_D"], iters=8, tol=0.0)
    assert np.allclose(r.A, g["rpca_wide_A"], rtol=0, atol=1e-12)
    Q = O.rpca_ga(g["ga_X"], 3, q0=g["ga_q0"])
    assert np.allclose(Q, g["ga_Q"], rtol=0, atol=1e-12)
    assert np.allclose(O.lowrankfilter(g["lrf_y"], 10), g["lrf_yf"], rtol=0, atol=1e-12)


def test_max_iterations_warning():
    """src/robustPCA.jl:232"""
    D, _, _, _ = golden5()
    with pytest.warns(UserWarning, match="Maximum number of iterations"):
        O.rpca(D, iters=3)


def test_robust_averages_match_the_reference_identities():
    """test/runtests.jl:469-488: trimmed mean with P = 0 is the weighted mean; with P = 0.1 and unit weights it is the
    plain trimmed mean of every row (StatsBase.trim drops floor(P N) entries on each side)."""
    rng = np.random.default_rng(0)
    U = rng.standard_normal((10, 10))
    s = np.zeros(10)
    w = np.ones(10)
    assert np.allclose(O.entrywise_trimmed_mean(s, w, U, 0).copy(), U.mean(axis=1))
    w = rng.standard_normal(10)
    assert np.allclose(O.entrywise_trimmed_mean(s, w, U, 0).copy(), (U * w).sum(axis=1) / w.sum())
    w = np.ones(10)
    m2 = O.entrywise_trimmed_mean(s, w, U, 0.1).copy()
    for i in range(10):
        assert np.isclose(m2[i], np.sort(U[i])[1:9].mean())
    med = O.entrywise_median(np.zeros(10), np.ones(10), U)
    for i in range(10):
        assert med[i] == np.sort(U[i])[10 // 2 - 1]
    # pluggable into rpca_ga like the reference's mu keyword (:255, :294)
    X = rng.standard_normal((6, 200))
    q0 = rng.standard_normal((6, 2))
    Q = O.rpca_ga(X, 2, q0=q0, mu=O.entrywise_trimmed_mean, exact_order=False, iters=50)
    assert np.all(np.isfinite(Q)) and np.allclose(np.linalg.norm(Q, axis=0), 1.0)


def test_oracle_reproduces_the_extended_golden_file():
    """tests/golden/oracle_vectors_ext.npz (hankel=true, channels, plain SSA, robust averages) is what the CUDA path is
    compared with on the GPU box; the oracle must keep reproducing it."""
    import warnings
    g = np.load(os.path.join(HERE, "golden", "oracle_vectors_ext.npz"))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        r = O.rpca(g["hk_H"], hankel=True, nukeA=False, iters=12, tol=0.0)
    assert np.allclose(r.A, g["hk_A"], rtol=0, atol=1e-12) and np.allclose(r.E, g["hk_E"], rtol=0, atol=1e-12)
    assert O.ishankel(O.rpca(g["hk_H"], hankel=True, nukeA=False).A)          # test/runtests.jl:331
    assert np.array_equal(O.hankel(g["mc_Yn"], 5, 2), g["mc_H"])
    assert np.allclose(O.unhankel(g["mc_H"], 2, 300, 2), g["mc_unh"], rtol=0, atol=1e-14)
    assert np.allclose(O.lowrankfilter(g["mc_Yn"], 12), g["mc_yf"], rtol=0, atol=1e-10)
    assert np.allclose(O.lowrankfilter(g["mc_Yn"], 12, lag=2, sv=4), g["mc_yf_ssa"], rtol=0, atol=1e-10)
    assert np.allclose(O.lowrankfilter(g["ssa_y"], 20, sv=2), g["ssa_yf"], rtol=0, atol=1e-10)
    for name, mu in (("trimmed", O.entrywise_trimmed_mean), ("median", O.entrywise_median)):
        Q, its = O.rpca_ga(g["ra_X"], 2, q0=g["ra_q0"], mu=mu, exact_order=False, iters=25, return_iters=True)
        assert list(its) == list(g[f"ra_its_{name}"]) and np.allclose(Q, g[f"ra_Q_{name}"], rtol=0, atol=1e-10)


def test_soft_hankel_reference_properties():
    """test/runtests.jl:296-307: soft_hankel! pulls a noisy Hankel matrix towards the clean one (both signs)."""
    rng = np.random.default_rng(5)
    A = O.hankel(np.arange(1.0, 9.0), 4)
    assert O.ishankel(A)
    for sgn in (1.0, -1.0):
        An = sgn * A + 0.1 * rng.standard_normal(A.shape)
        assert not O.ishankel(An)
        Anc = An.copy()
        O.soft_hankel(An, 0.1)
        assert np.sum((An - sgn * A) ** 2) < np.sum((Anc - sgn * A) ** 2)
    assert np.array_equal(O.hankel(np.arange(1, 21), 2), np.column_stack([np.arange(1, 20), np.arange(2, 21)]))
    assert np.array_equal(O.hankel(np.arange(1, 21), 3, 2),
                          np.column_stack([np.arange(1, 18, 2), np.arange(2, 19, 2), np.arange(3, 20, 2)]))
