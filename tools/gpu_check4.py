import os, sys, time, warnings
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import tlsq_b200 as T
import tls_oracle as O
from tools.gpu_check import relF, section
section("converging solves")
for (M, N, r, kw) in [(20000, 256, 10, {"nonnegA": True}), (20000, 256, 10, {}), (12000, 128, 20, {})]:
    D = T.synth.lowrank_sparse_np(M, N, r, 0.05, seed=4, nonneg=bool(kw.get("nonnegA")))
    A, E, s, sv, info = T.rpca(D, return_info=True, **kw)
    ro = O.rpca(D, **kw)
    print(M, N, r, kw, "iters", info["iters"], ro.iters, "relF A", relF(A, ro.A), "E", relF(E, ro.E), "svp equal", np.array_equal(info["hist"][:, 1], ro.hist[:, 1]), flush=True)
