import os, sys, torch, numpy as np
sys.path.insert(0, ".")
import tlsq_b200 as T
dev = torch.device("cuda", 0)
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
D = T.synth.lowrank_sparse_cuda(0, rows, 256, dev, 10, 0.05, 4, True)
kw = dict(nonnegA=True, lam=1.0 / np.sqrt(rows))
for name, env in [("tma", {}), ("notma", {"TLSQ_STREAM_TMA": "0"}), ("fused", {"TLSQ_FUSED": "1"})]:
    for k, v in env.items(): os.environ[k] = v
    hs = []
    for r in range(reps):
        A, E, s, sv, info = T.rpca(D, return_info=True, want_svd=False, exact_cost=True, **kw)
        hs.append(info["hist"][:, 2].copy())
    for k in env: del os.environ[k]
    n = min(len(h) for h in hs)
    H = np.stack([h[:n] for h in hs])
    med = np.median(H, axis=0)
    dev_ = np.abs(H / med - 1.0)
    bad = [(int(r), int(k) + 1, float(H[r, k]), float(med[k])) for r, k in zip(*np.where(dev_ > 1e-4))]
    print(name, "iters", [len(h) for h in hs], "max rel dev", float(dev_.max()), "outliers (rep, k, cost, median):", bad[:8], flush=True)
