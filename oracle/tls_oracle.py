"""CPU oracle for the robust-PCA hot path of TotalLeastSquares.jl  --  TEST INFRASTRUCTURE ONLY.

This module is a CPU restatement (NumPy + SciPy/LAPACK) of the reference algorithm.  It is the
*checker* for the CUDA path: only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it.  Nothing under ``totalleastsquares.jl_b200/``
imports it, and the product path raises if the CUDA library is missing (no CPU fallback).

Pinning status
--------------
* The reference is pure Julia and Julia is not installed in the build image nor on the GPU box, so the
  reference itself cannot be executed here (``oracle/_ref`` does not exist for this project).
* The heavy arithmetic of the reference lives in Julia's stdlib ``LinearAlgebra`` -> OpenBLAS/LAPACK
  (``dgesdd`` jobz='S' for ``svd!``, jobz='N' for ``opnorm`` -> ``svdvals``; ``dgemm`` for ``mul!``).  The repo
  has no Manifest.toml, hence no pinned version (Project.toml:18 says julia = "1.0"; CI uses 1.12).  This
  restatement calls the *same LAPACK routines* through SciPy's bundled OpenBLAS.
* Pinned against every reference-authored fixture for the path (tests/test_oracle.py):
  the 5x5 known-answer test (test/runtests.jl:143-165, atol 1e-6), the exact hankel/unhankel vectors
  (test/runtests.jl:293-294, 361-376), the mu! weighted-mean identity (:469-480), the Q'Q=I invariant
  (:447-464), the statistical bounds (:172-185, :378-380, README.md:85-106), the soft_hankel! properties
  (:293-307) and the robust-average identities (:469-488).
* Beyond those fixtures numerical parity at the LAPACK boundary is **unpinned by the reference's own
  tests**; the oracle run is the pin ("parity unpinned" beyond the 5x5 golden).

All ``file:line`` citations are relative to /root/reference/.
Arrays are handled in Fortran (column-major) order like Julia's.
"""
from __future__ import annotations

import math
import warnings
from typing import Callable, NamedTuple, Optional

import numpy as np
import scipy.linalg as sla

__all__ = [
    "soft_th", "soft_th_level", "soft_hankel", "ishankel", "hankel", "unhankel", "lowrankfilter",
    "entrywise_trimmed_mean", "entrywise_median",
    "rpca", "rpca_ga", "rpca_ga_1", "mu_mean", "SVD", "RpcaResult",
]


class SVD(NamedTuple):
    """Mirror of Julia's LinearAlgebra.SVD (fields U, S, Vt) -- src/robustPCA.jl:194,238."""
    U: np.ndarray
    S: np.ndarray
    Vt: np.ndarray

    @property
    def V(self):
        return self.Vt.T


class RpcaResult(NamedTuple):
    A: np.ndarray
    E: np.ndarray
    s: SVD
    sv: int
    iters: int          # number of ALM iterations executed (not returned by the reference; for tests)
    hist: np.ndarray    # (iters, 3): k, svp, cost


# ----------------------------------------------------------------------------------------------
# soft thresholds -- src/robustPCA.jl:1-2
# ----------------------------------------------------------------------------------------------
def soft_th(x, eps):
    """src/robustPCA.jl:1   soft_th(x, e) = max(x-e, 0) + min(x+e, 0)."""
    return np.maximum(x - eps, 0.0) + np.minimum(x + eps, 0.0)


def soft_th_level(x, eps, l):
    """src/robustPCA.jl:2   soft_th(x, e, l) = max(x-e, l) + min(x+e, l) - l."""
    return np.maximum(x - eps, l) + np.minimum(x + eps, l) - l


def _antidiag_indices(K, L, k):
    """Index ranges of anti-diagonal k (1-based) -- src/robustPCA.jl:13-14 / 32-33."""
    ri = np.arange(min(K, k), max(k - L, 1) - 1, -1)          # min(K,k):-1:max(k-L,1)
    ci = np.arange(max(1, k - K + 1), L + 1)                   # max(1,k-K+1):L
    n = min(len(ri), len(ci))                                  # zip truncates to the shorter
    return ri[:n] - 1, ci[:n] - 1


def soft_hankel(A, eps):
    """src/robustPCA.jl:9-21  soft_hankel!(A, e): threshold every anti-diagonal towards its mean (in place)."""
    K, L = A.shape
    for k in range(1, K + L):
        r, c = _antidiag_indices(K, L, k)
        m = A[r, c].mean()
        A[r, c] = soft_th_level(A[r, c], eps, m)
    return A


def ishankel(A):
    """src/robustPCA.jl:94-106."""
    K, L = A.shape
    for k in range(1, K + L):
        r, c = _antidiag_indices(K, L, k)
        if np.any(A[r, c] != A[r[0], c[0]]):
            return False
    return True


# ----------------------------------------------------------------------------------------------
# hankel / unhankel -- src/robustPCA.jl:28-39, 53-68, 76-92
# ----------------------------------------------------------------------------------------------
def hankel(x, L, lag=1):
    """src/robustPCA.jl:76-92.  x: (N,) or (N,D).  Returns K x (L*D), K=(N-L)//lag+1,
    X[k,(l-1)D+d] = x[(k-1)lag+l, d]."""
    x = np.asarray(x)
    x2 = x.reshape(x.shape[0], -1)
    N, D = x2.shape
    assert L <= N / 2, f"L has to be less than N/2 = {N / 2}"      # :79
    assert lag <= L, "lag must be <= L"                            # :80
    K = (N - L) // lag + 1                                         # :81
    X = np.empty((K, L * D), dtype=x2.dtype, order="F")
    for d in range(D):
        for k in range(K):
            X[k, d::D] = x2[k * lag:k * lag + L, d]               # :87-88
    return X


def unhankel(A, lag=1, N=None, D=1):
    """src/robustPCA.jl:28-39 (lag==1 && D==1) and :53-68 (general)."""
    A = np.asarray(A)
    if lag == 1 and D == 1:
        K, L = A.shape
        n = L + (K - 1)
        y = np.empty(n, dtype=A.dtype)
        for k in range(1, n + 1):
            r, c = _antidiag_indices(K, L, k)
            y[k - 1] = A[r, c].mean()                              # :35-36
        return y
    K = A.shape[0]
    L = A.shape[1] // D
    y = np.zeros((N, D), dtype=A.dtype)
    counts = np.zeros((N, D), dtype=np.int64)
    # inds = hankel(indmat, L, lag)  (:61) -- enumerate in column-major order of A (:62)
    rows = np.arange(N)
    for d in range(D):
        for l in range(L):
            col = l * D + d
            for k in range(K):
                t = k * lag + l
                y[t, d] += A[k, col]
                counts[t, d] += 1
    y /= np.maximum(counts, 1)                                     # :66
    return y[:, 0] if D == 1 else y


def unhankel_fast(A):
    """Vectorised anti-diagonal mean (same result as unhankel(A) up to summation order); for big K."""
    K, L = A.shape
    n = K + L - 1
    s = np.zeros(n)
    cnt = np.zeros(n)
    for c in range(L):
        s[c:c + K] += A[:, c]
        cnt[c:c + K] += 1
    return s / cnt


# ----------------------------------------------------------------------------------------------
# rpca -- src/robustPCA.jl:156-239
# ----------------------------------------------------------------------------------------------
def _opnorm(X):
    """LinearAlgebra.opnorm(X) = largest singular value (svdvals -> LAPACK dgesdd jobz='N')."""
    return float(sla.svdvals(X, check_finite=False)[0])


def _svd(Z):
    """LinearAlgebra.svd!(Z): thin SVD via LAPACK dgesdd jobz='S' -- src/robustPCA.jl:194."""
    U, S, Vt = sla.svd(Z, full_matrices=False, lapack_driver="gesdd", check_finite=False, overwrite_a=False)
    return SVD(U, S, Vt)


def rpca(D, lam=None, maxrank=None, iters=1000, tol=None, rho=1.5, verbose=False, nonnegA=False,
         nonnegE=False, hankel=False, nukeA=True, svd: Optional[Callable] = None,
         opnorm: Optional[Callable] = None, **kwargs) -> RpcaResult:
    """Inexact-ALM robust PCA, literal to src/robustPCA.jl:156-239 (unknown kwargs swallowed, :170).

    Operation order preserved: ``(D - A) + (1/mu)*Y`` (:188), ``>=`` in the rank count (:198), mu bumped before
    the cost (:223-225), break before warn (:228-232).
    """
    D = np.asarray(D, dtype=np.float64, order="F")
    M, N = D.shape
    lam = 1.0 / math.sqrt(max(M, N)) if lam is None else float(lam)          # :157
    maxrank = np.iinfo(np.int64).max if maxrank is None else int(maxrank)    # :158
    tol = math.sqrt(np.finfo(np.float64).eps) if tol is None else float(tol) # :160
    opn = _opnorm if opnorm is None else opnorm
    d = min(M, N)                                                            # :173
    A = np.zeros((M, N), order="F")                                          # :174
    E = np.zeros((M, N), order="F")
    Z = np.empty((M, N), order="F")                                          # :175
    Y = D.copy(order="F")                                                    # :176
    norm2 = float(opn(Y))                                                    # :177
    norminf = float(np.max(np.abs(Y))) / lam if Y.size else 0.0              # :178  norm(Y, Inf): max|Y_ij|
    dual_norm = max(norm2, norminf)                                          # :179
    d_norm = norm2                                                           # :180
    Y /= dual_norm                                                           # :181
    mu = 1.25 / norm2                                                        # :182
    mubar = mu * 1.0e7                                                       # :183
    sv = svp = 10                                                            # :184
    s = None
    hist = []
    k_done = 0
    tmp = np.empty((M, N), order="F")
    for k in range(1, iters + 1):                                            # :186
        imu = 1.0 / mu
        # E .= soft_th.(D .- A .+ (1/mu) .* Y, lam/mu)                       # :188
        np.multiply(Y, imu, out=tmp)
        np.subtract(D, A, out=E)
        np.add(E, tmp, out=E)
        eps_ = lam / mu
        np.subtract(E, eps_, out=Z)
        np.maximum(Z, 0.0, out=Z)
        np.add(E, eps_, out=E)
        np.minimum(E, 0.0, out=E)
        np.add(Z, E, out=E)
        if nonnegE:
            np.maximum(E, 0.0, out=E)                                        # :189-191
        # Z .= D .- E .+ (1/mu) .* Y                                         # :192
        np.subtract(D, E, out=Z)
        np.add(Z, tmp, out=Z)
        if svd is None or k == 1:
            s = _svd(Z)                                                      # :193-194
        else:
            s = SVD(*svd(Z, sv))                                             # :196
        svp = int(np.sum(s.S >= imu))                                        # :198
        sv = svp                                                             # :199-203 (both branches)
        sv = min(max(sv, 1), maxrank)                                        # :204 clamp
        if nukeA:
            Zs = s.U[:, :svp] * (s.S[:svp] - imu)                            # :207
        else:
            Zs = s.U[:, :svp] * s.S[:svp]                                    # :211
        np.matmul(Zs, s.Vt[:svp, :], out=A) if svp > 0 else A.fill(0.0)      # :208 / :212
        if hankel:
            soft_hankel(A, lam / mu)                                         # :214-216
        if nonnegA:
            np.maximum(A, 0.0, out=A)                                        # :217-219
        np.subtract(D, A, out=Z)                                             # :221  @. Z = D - A - E
        np.subtract(Z, E, out=Z)
        np.multiply(Z, mu, out=tmp)                                          # :222  @. Y = Y + mu*Z
        np.add(Y, tmp, out=Y)
        mu = min(mu * rho, mubar)                                            # :223
        cost = float(opn(Z)) / d_norm                                        # :225
        hist.append((k, svp, cost))
        k_done = k
        if verbose:
            print(f"{k} cost: {cost:.4g}")
        if cost < tol:                                                       # :228
            if verbose:
                print("converged")
            break
        if k == iters:
            warnings.warn(f"Maximum number of iterations reached, cost: {cost}, tol: {tol}")  # :232
    if hankel:
        soft_hankel(E, lam / mu)                                             # :234-236
    return RpcaResult(A, E, s, sv, k_done, np.array(hist, dtype=np.float64).reshape(-1, 3))


# ----------------------------------------------------------------------------------------------
# lowrankfilter -- src/robustPCA.jl:119-128
# ----------------------------------------------------------------------------------------------
def lowrankfilter(y, n=None, sv=0, lag=1, tol=1e-3, **kwargs):
    y = np.asarray(y, dtype=np.float64)
    N = y.shape[0]
    Dch = 1 if y.ndim == 1 else y.shape[1]
    n = min(N // 20, 2000) if n is None else int(n)                          # :119
    H = hankel(y, n, lag)                                                    # :120
    if sv <= 0:
        A = rpca(H, tol=tol, **kwargs).A                                     # :122
    else:
        s = _svd(H)                                                          # :124
        A = (s.U[:, :sv] * s.S[:sv]) @ s.Vt[:sv, :]                          # :125
    return unhankel(A, lag, N, Dch)                                          # :127


# ----------------------------------------------------------------------------------------------
# rpca_ga -- src/robustPCA.jl:255-316
# ----------------------------------------------------------------------------------------------
def mu_mean(s, w, U, exact_order=True):
    """src/robustPCA.jl:308-316  mu!(s,w,U): s = sum_n w[n] U[:,n] / sum_n w[n]  (written into s)."""
    if exact_order:
        ws = 0.0
        s[:] = 0.0
        for n in range(U.shape[1]):
            ws += w[n]
            s += w[n] * U[:, n]
    else:
        ws = float(np.sum(w))
        s[:] = U @ w
    s /= ws
    return s


def entrywise_trimmed_mean(s, w, U, P=0.1):
    """src/robustPCA.jl:323-333: per row, drop a fraction P of the sorted entries on each side, then the w-weighted
    mean of the rest.  range = (1 + floor(P N)) : floor((1 - P) N)  (1-based, inclusive)."""
    N = U.shape[1]
    lo, hi = int(math.floor(P * N)), int(math.floor((1 - P) * N))            # 0-based half-open [lo, hi)
    s[:] = 0.0
    for j in range(U.shape[0]):
        I = np.argsort(U[j, :], kind="stable")[lo:hi]                        # sortperm(U[j,:])[range]   :328
        s[j] += (w[I] @ U[j, I]) / np.sum(w[I])                              # :329
    return s


def entrywise_median(s, w, U):
    """src/robustPCA.jl:349-357: s[j] = sign(w[m]) U[j, m] with m the (N ÷ 2)-th entry of sortperm(w .* U[j,:])."""
    N = U.shape[1]
    s[:] = 0.0
    for j in range(U.shape[0]):
        I = np.argsort(w * U[j, :], kind="stable")
        m = I[N // 2 - 1]                                                    # I[end ÷ 2], 1-based
        s[j] = np.sign(w[m]) * U[j, m]
    return s


def rpca_ga_1(Xnorms, U, w, q0, tol=1e-7, iters=1000, verbose=False, mu=None, exact_order=True):
    """src/robustPCA.jl:283-306.  q0 replaces the global-RNG draw randn(d) (:286); it is normalised here (:287)."""
    d, N = U.shape
    q = np.array(q0, dtype=np.float64, copy=True)
    q /= np.linalg.norm(q)                                                   # :287
    qold = q.copy()                                                          # :288
    its = 0
    for i in range(1, iters + 1):                                            # :290
        if exact_order:
            for n in range(N):
                w[n] = np.sign(U[:, n] @ q) * Xnorms[n]                      # :291-293
        else:
            w[:] = np.sign(U.T @ q) * Xnorms
        mui = (mu or (lambda s_, w_, U_: mu_mean(s_, w_, U_, exact_order)))(q, w, U)   # :294 (writes into q)
        q[:] = mui / np.linalg.norm(mui)                                     # :295
        dq = math.sqrt(float(np.sum((q - qold) ** 2)))                       # :296
        its = i
        if dq < tol:                                                         # :298
            break
        qold[:] = q                                                          # :302
        if i == iters:
            warnings.warn("Reached maximum number of iterations")            # :303
    return q, its


def rpca_ga(X, r=None, q0=None, tol=1e-7, iters=1000, verbose=False, mu=None, exact_order=True,
            return_iters=False):
    """src/robustPCA.jl:255-278.  X is d x N (columns are observations).  q0: d x r start vectors
    (column i is the randn(d) the reference would draw for component i)."""
    X = np.array(X, dtype=np.float64, order="F", copy=True)                  # :257
    d, N = X.shape
    r = min(d, N) if r is None else int(r)
    Q = np.zeros((d, r), order="F")                                          # :260
    w = np.zeros(N)
    Xnorms = np.zeros(N)
    U = np.empty_like(X)
    its = []
    for i in range(r):                                                       # :263
        Xnorms[:] = np.sqrt(np.sum(X * X, axis=0))                           # :265
        with np.errstate(invalid="ignore", divide="ignore"):
            np.divide(X, Xnorms, out=U)                                      # :266
        q, it = rpca_ga_1(Xnorms, U, w, q0[:, i], tol=tol, iters=iters, verbose=verbose, mu=mu,
                          exact_order=exact_order)                           # :268
        its.append(it)
        Q[:, i] = q                                                          # :269
        Xs1 = q @ X                                                          # :271
        X -= np.outer(q, Xs1)                                                # :272
    return (Q, its) if return_iters else Q
