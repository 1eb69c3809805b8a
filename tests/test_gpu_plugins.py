"""GPU tests of the reference's plugin hooks and element types:
svd / opnorm callables (src/robustPCA.jl:168-169, 177, 193-197, 225; test/runtests.jl:384-398) crossing the C ABI as
function pointers, and Float32 inputs (promoted to Float64 on the way in, returned as Float32)."""
import warnings

import numpy as np
import pytest

import tls_oracle as O
import tlsq_b200 as T

pytestmark = pytest.mark.gpu
TOL = 1e-9


def relF(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def lapack_svd(Z, sv):
    U, S, Vt = np.linalg.svd(Z, full_matrices=False)
    return U, S, Vt


def top_svd(Z, sv):
    """a truncated plugin: only sv + 2 triplets, like the randomised svd of the reference's test (:384-398)"""
    U, S, Vt = np.linalg.svd(Z, full_matrices=False)
    k = min(sv + 2, len(S))
    return O.SVD(U[:, :k], S[:k], Vt[:k])


@pytest.mark.parametrize("M,N,kw", [(400, 30, {}), (60, 90, {"nonnegA": True}), (300, 20, {"nukeA": False})])
def test_rpca_with_exact_callables_matches_the_oracle(M, N, kw):
    D = T.synth.lowrank_sparse_np(M, N, 3, 0.05, seed=M + N, nonneg=bool(kw.get("nonnegA")))
    opn = lambda Z: float(np.linalg.norm(Z, 2))                                  # noqa: E731
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = O.rpca(D, svd=lapack_svd, opnorm=opn, **kw)
        A, E, s, sv, info = T.rpca(D, svd=lapack_svd, opnorm=opn, return_info=True, **kw)
    assert info["iters"] == ref.iters and sv == ref.sv
    assert relF(A, ref.A) < TOL and relF(E, ref.E) < TOL
    assert np.allclose(info["hist"][:, 2], ref.hist[:, 2], rtol=1e-9)           # the user's opnorm IS the cost
    assert np.allclose(s.S, ref.s.S, rtol=0, atol=1e-12 * ref.s.S[0])
    # only one of the two hooks: the other one is the built-in device implementation
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        A2, E2, s2, sv2, info2 = T.rpca(D, opnorm=opn, return_info=True, **kw)
        A3, E3, s3, sv3, info3 = T.rpca(D, svd=lapack_svd, return_info=True, **kw)
    for (Ax, Ex, ix) in ((A2, E2, info2), (A3, E3, info3)):
        assert ix["iters"] == ref.iters and relF(Ax, ref.A) < TOL and relF(Ex, ref.E) < TOL
    assert np.allclose(s2.S, ref.s.S, rtol=0, atol=1e-12 * ref.s.S[0])          # built-in SVD of the last SVT input


def test_truncated_svd_plugin_and_maxrank_like_the_reference_test():
    """test/runtests.jl:384-398: lowrankfilter(y + n, 50; opnorm = randomised norm, svd = randomised svd, maxrank = 5)"""
    rng = np.random.default_rng(5)
    Tn = 1000
    t = np.arange(1, Tn + 1)
    qn = lambda x: x / np.quantile(np.abs(x), 0.9)                               # noqa: E731
    y = qn(np.sin(0.1 * t))
    n = 20 * rng.standard_normal(Tn) * (rng.random(Tn) < 0.01) + 0.1 * rng.standard_normal(Tn)
    opn = lambda Z: float(np.linalg.norm(Z, 2))                                  # noqa: E731
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        yf = T.lowrankfilter(y + n, 50, svd=top_svd, opnorm=opn, maxrank=5)
        yo = O.lowrankfilter(y + n, 50, svd=top_svd, opnorm=opn, maxrank=5)
    assert relF(yf, yo) < TOL
    assert np.mean((y - qn(yf)) ** 2) / np.mean(n ** 2) < 0.05
    # an exception inside the callable propagates
    def bad(Z, sv):
        raise ValueError("boom")
    with pytest.raises(ValueError, match="boom"):
        T.rpca(np.asfortranarray(rng.standard_normal((40, 8))), svd=bad, iters=3)


def test_float32_inputs_are_promoted_and_returned_as_float32():
    D = T.synth.lowrank_sparse_np(3000, 64, 4, 0.05, seed=3)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        A64, E64, s64, sv64 = T.rpca(D, iters=10, tol=0.0)
        A32, E32, s32, sv32 = T.rpca(D.astype(np.float32), iters=10, tol=0.0)
    assert A32.dtype == np.float32 and E32.dtype == np.float32 and s32.S.dtype == np.float32 and sv32 == sv64
    assert relF(A32, A64) < 1e-5 and relF(E32, E64) < 1e-4
    y, yn = T.synth.sinusoid_np(2000, seed=3)
    yf32 = T.lowrankfilter(yn.astype(np.float32), 40)
    assert yf32.dtype == np.float32 and relF(yf32, T.lowrankfilter(yn, 40)) < 1e-4
    X, q0 = T.synth.ga_data_np(500, 64, 3, seed=2)
    Q32 = T.rpca_ga(X.astype(np.float32), 2, q0=q0[:, :2].astype(np.float32))
    Q64 = T.rpca_ga(X, 2, q0=q0[:, :2])
    assert Q32.dtype == np.float32 and np.abs(np.abs(Q32) - np.abs(Q64)).max() < 1e-4
