import os, sys, time, warnings
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import tlsq_b200 as T
import tls_oracle as O
from tools.gpu_check import relF, section
section("rpca parity with the fast eigen path (n in (64,256])")
for (M, N, r, kw, its) in [(8192, 256, 10, {}, 14), (10000, 128, 6, {"nonnegA": True, "nonnegE": True}, 14), (5001, 200, 5, {"nukeA": False}, 9),
                           (3000, 96, 40, {}, 10)]:
    D = T.synth.lowrank_sparse_np(M, N, r, 0.05, seed=M + N, nonneg=bool(kw.get("nonnegA")))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        A, E, s, sv, info = T.rpca(D, iters=its, tol=0.0, return_info=True, **kw)
        ro = O.rpca(D, iters=its, tol=0.0, **kw)
    supp = int(np.sum((E != 0) != (ro.E != 0)))
    print(f"{M}x{N} r={r} {kw} its={its}: relF A={relF(A, ro.A):.2e} E={relF(E, ro.E):.2e} supp_mismatch={supp} sv={sv}/{ro.sv} "
          f"svp_equal={np.array_equal(info['hist'][:, 1], ro.hist[:, 1])} S={relF(s.S, ro.s.S):.1e}", flush=True)
    if not np.array_equal(info['hist'][:, 1], ro.hist[:, 1]):
        print("   svp gpu", info['hist'][:, 1].tolist(), "oracle", ro.hist[:, 1].tolist())
section("converging solve 20000x256 nonnegA")
D = T.synth.lowrank_sparse_np(20000, 256, 10, 0.05, seed=4, nonneg=True)
A, E, s, sv, info = T.rpca(D, nonnegA=True, return_info=True)
ro = O.rpca(D, nonnegA=True)
print("iters", info["iters"], ro.iters, "relF A", relF(A, ro.A), "E", relF(E, ro.E), "svp", info["hist"][:, 1].tolist())
