"""Short single-GPU workloads for ncu captures (tools/gpu_r2p.sh):  python tools/prof_driver.py c4|c4fused|ga|c2 [solves]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tlsq_b200 as T

what = sys.argv[1]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
dev = torch.device("cuda", 0)
if what in ("c4", "c4fused"):
    if what == "c4fused":
        os.environ["TLSQ_FUSED"] = "1"
    D = T.synth.lowrank_sparse_cuda(0, 1_000_000, 256, dev, 10, 0.05, 4, True)
    for _ in range(reps):
        A, E, s, sv, info = T.rpca(D, nonnegA=True, lam=1e-3, return_info=True)
    print(what, info["iters"], sv)
elif what == "c2":
    D = T.synth.lowrank_sparse_cuda(0, 100_000, 512, dev, 10, 0.05, 2, False)
    for _ in range(reps):
        A, E, s, sv, info = T.rpca(D, lam=1.0 / 100_000 ** 0.5, return_info=True)
    print(what, info["iters"], sv)
else:
    X, q0 = T.synth.ga_data_cuda(0, 2_000_000, 256, dev, 10, 3)
    for _ in range(reps):
        Q, info = T.rpca_ga(X, 2, q0=q0[:, :2].contiguous() if False else q0.t()[:2].t(), return_info=True)
    print(what, info["iters"])
torch.cuda.synchronize()
