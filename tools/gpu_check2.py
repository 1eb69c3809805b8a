import os, sys, time, warnings
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import tlsq_b200 as T
import tls_oracle as O
from tools.gpu_check import colmajor, relF, t_gram, run, section
dev = torch.device("cuda", 0)
run(t_gram)
section("rpca parity on the W/SYRK fast path")
for (M, N, r, kw, its) in [(8192, 256, 10, {}, 9), (10000, 128, 6, {"nonnegA": True, "nonnegE": True}, 11), (6000, 512, 8, {}, 5), (5001, 256, 5, {}, 7)]:
    D = T.synth.lowrank_sparse_np(M, N, r, 0.05, seed=M + N, nonneg=bool(kw.get("nonnegA")))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        A, E, s, sv, info = T.rpca(D, iters=its, tol=0.0, return_info=True, **kw)
        ro = O.rpca(D, iters=its, tol=0.0, **kw)
    supp = int(np.sum((E != 0) != (ro.E != 0)))
    print(f"{M}x{N} {kw} its={its}: relF A={relF(A, ro.A):.2e} E={relF(E, ro.E):.2e} supp_mismatch={supp} sv={sv}/{ro.sv} "
          f"svp_equal={np.array_equal(info['hist'][:, 1], ro.hist[:, 1])} U-orth={np.abs(s.U.T @ s.U - np.eye(s.U.shape[1])).max():.1e}")
section("eig sweeps during a converging solve (TLSQ_DEBUG_EIG)")
D = T.synth.lowrank_sparse_np(20000, 256, 10, 0.05, seed=4, nonneg=True)
t0 = time.perf_counter()
A, E, s, sv, info = T.rpca(D, nonnegA=True, return_info=True)
print("iters", info["iters"], "time", time.perf_counter() - t0)
