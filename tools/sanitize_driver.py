"""Small-shape run of every kernel family for compute-sanitizer (tools/gpu_sanitize.sh):
   python tools/sanitize_driver.py [fused|legacy|largen|ga|lrf|plugins ...]"""
import os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import tlsq_b200 as T

which = sys.argv[1:] or ["fused", "legacy", "largen", "ga", "lrf", "plugins"]
warnings.simplefilter("ignore")
if "fused" in which:                      # one-pass cluster kernel (DSMEM, TMA, remote stores) + cluster Jacobi
    os.environ["TLSQ_FUSED"] = "1"
    D = T.synth.lowrank_sparse_np(8192, 256, 5, 0.05, seed=1)
    A, E, s, sv = T.rpca(D, iters=3, tol=0.0)
    print("fused ok", sv)
    del os.environ["TLSQ_FUSED"]
if "legacy" in which:                     # TMA SYRK + TMA-staged streaming epilogue + fast eigen path + SVD refinement
    os.environ["TLSQ_FUSED"] = "0"
    D = T.synth.lowrank_sparse_np(16384, 256, 5, 0.05, seed=2)
    A, E, s, sv = T.rpca(D, iters=4, tol=0.0)
    print("legacy ok", sv)
    del os.environ["TLSQ_FUSED"]
if "largen" in which:                     # large-n subspace iteration, loop Jacobi, chunked epilogue, GEMM projection
    D = T.synth.lowrank_sparse_np(2000, 600, 4, 0.05, seed=3)
    A, E, s, sv = T.rpca(D, iters=2, tol=0.0)
    print("largen ok", sv)
if "ga" in which:                         # TMA-ring Grassmann sweep + fused deflation
    X, q0 = T.synth.ga_data_np(8200, 64, 3, seed=4)
    Q = T.rpca_ga(X, 2, q0=q0[:, :2], iters=4)
    print("ga ok", Q.shape)
if "lrf" in which:                        # implicit Hankel, factored unhankel, in-place dual variable
    y, yn = T.synth.sinusoid_np(9000, seed=5)
    yf = T.lowrankfilter(yn, 256, iters=3)
    os.environ["TLSQ_FUSED"] = "1"; os.environ["TLSQ_INPLACE_Y"] = "1"
    yf = T.lowrankfilter(yn, 256, iters=3)
    del os.environ["TLSQ_FUSED"], os.environ["TLSQ_INPLACE_Y"]
    print("lrf ok", yf.shape)
if "plugins" in which:
    D = T.synth.lowrank_sparse_np(300, 20, 2, 0.05, seed=6)
    A, E, s, sv = T.rpca(D, iters=2, tol=0.0, opnorm=lambda Z: float(np.linalg.norm(Z, 2)))
    print("plugins ok", sv)
