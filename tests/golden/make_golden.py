"""Regenerates tests/golden/oracle_vectors.npz with the CPU oracle (oracle/tls_oracle.py).

The reference itself (Julia) cannot run in this image, so these vectors pin the ORACLE's behaviour (regression
fixtures) and are what the CUDA path is compared with on the GPU box, where /root/reference does not exist.
The only reference-authored known-answer vector is tests/golden/rpca_5x5.json (test/runtests.jl:143-165).

    python tests/golden/make_golden.py
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, ROOT)
import tls_oracle as O  # noqa: E402
import tlsq_b200 as T  # noqa: E402  (only the NumPy generators in synth.py are used)


def main():
    out = {}
    warnings.simplefilter("ignore")
    # rpca, fixed iteration count, three flag combinations
    D = T.synth.lowrank_sparse_np(60, 12, 3, 0.08, seed=101)
    out["rpca_D"] = D
    for name, kw in [("plain", {}), ("nonneg", {"nonnegA": True, "nonnegE": True}), ("nonuke", {"nukeA": False})]:
        Dk = np.abs(D) if name == "nonneg" else D
        r = O.rpca(Dk, iters=10, tol=0.0, **kw)
        out[f"rpca_{name}_A"], out[f"rpca_{name}_E"], out[f"rpca_{name}_S"] = r.A, r.E, r.s.S
        out[f"rpca_{name}_hist"] = r.hist
    # wide matrix (M < N)
    Dw = T.synth.lowrank_sparse_np(9, 40, 2, 0.08, seed=102)
    r = O.rpca(Dw, iters=8, tol=0.0)
    out["rpca_wide_D"], out["rpca_wide_A"], out["rpca_wide_E"] = Dw, r.A, r.E
    # converged run
    r = O.rpca(D)
    out["rpca_conv_A"], out["rpca_conv_E"], out["rpca_conv_iters"], out["rpca_conv_sv"] = r.A, r.E, r.iters, r.sv
    # rpca_ga
    X, q0 = T.synth.ga_data_np(12, 30, 3, seed=103)
    Q, its = O.rpca_ga(X, 3, q0=q0, return_iters=True)
    out["ga_X"], out["ga_q0"], out["ga_Q"], out["ga_iters"] = X, q0, Q, np.array(its)
    # lowrankfilter
    y, yn = T.synth.sinusoid_np(200, seed=104, noise=0.05)
    out["lrf_y"], out["lrf_yf"] = yn, O.lowrankfilter(yn, 10)
    out["lrf_yf_lag2"] = O.lowrankfilter(yn, 10, lag=2)
    np.savez_compressed(os.path.join(HERE, "oracle_vectors.npz"), **out)
    print("wrote", os.path.join(HERE, "oracle_vectors.npz"), {k: np.shape(v) for k, v in out.items()})
    main_ext()


def main_ext():
    """Second fixture file: the branches added after the first one (hankel=true, channels, plain SSA, robust averages)."""
    out = {}
    rng = np.random.default_rng(2024)
    # rpca(...; hankel=true, nukeA=false)  (src/robustPCA.jl:214-216, 234-236)
    H = O.hankel(rng.standard_normal(300), 12)
    r = O.rpca(H, hankel=True, nukeA=False, iters=12, tol=0.0)
    out["hk_H"], out["hk_A"], out["hk_E"], out["hk_hist"] = H, r.A, r.E, r.hist
    # channels: hankel / unhankel / lowrankfilter on an N x 2 signal  (:53-68, :83-90, :119-128)
    t = np.arange(1, 301)
    Y = np.column_stack([np.sin(0.1 * t), np.sin(0.3 * t) + 0.5 * np.sin(0.1 * t)])
    Yn = Y + 0.1 * rng.standard_normal(Y.shape) + (rng.random(Y.shape) < 0.02) * 5.0
    out["mc_Yn"] = Yn
    out["mc_H"] = O.hankel(Yn, 5, 2)
    out["mc_unh"] = O.unhankel(out["mc_H"], 2, 300, 2)
    out["mc_yf"] = O.lowrankfilter(Yn, 12)
    out["mc_yf_ssa"] = O.lowrankfilter(Yn, 12, lag=2, sv=4)
    # plain SSA on one channel  (:123-125)
    yn = np.sin(0.1 * t) + rng.standard_normal(300)
    out["ssa_y"], out["ssa_yf"] = yn, O.lowrankfilter(yn, 20, sv=2)
    # robust averages of rpca_ga  (:323-333, :349-357)
    U0, S0, Vt0 = np.linalg.svd(rng.standard_normal((8, 120)), full_matrices=False)
    X = (U0[:, :2] * S0[:2]) @ Vt0[:2] + 1e-3 * rng.standard_normal((8, 120))
    X += 1000.0 * rng.standard_normal((8, 120)) * (rng.random((8, 120)) < 0.01)
    q0 = rng.standard_normal((8, 2))
    out["ra_X"], out["ra_q0"] = np.asfortranarray(X), np.asfortranarray(q0)
    Q, its = O.rpca_ga(X, 2, q0=q0, mu=O.entrywise_trimmed_mean, exact_order=False, iters=25, return_iters=True)
    out["ra_Q_trimmed"], out["ra_its_trimmed"] = Q, np.array(its)
    Q, its = O.rpca_ga(X, 2, q0=q0, mu=O.entrywise_median, exact_order=False, iters=25, return_iters=True)
    out["ra_Q_median"], out["ra_its_median"] = Q, np.array(its)
    np.savez_compressed(os.path.join(HERE, "oracle_vectors_ext.npz"), **out)
    print("wrote", os.path.join(HERE, "oracle_vectors_ext.npz"), {k: np.shape(v) for k, v in out.items()})


if __name__ == "__main__":
    main()
