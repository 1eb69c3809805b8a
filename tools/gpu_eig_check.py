"""Full symmetric eigensolver check + timing (run on the GPU box; TLSQ_JACOBI_FLAT=1 selects the flat tournament)."""
import os, sys, time
sys.path.insert(0, '.')
import numpy as np, torch
import tlsq_b200 as T
rng = np.random.default_rng(0)
for n in (5, 16, 17, 40, 64, 100, 130, 200, 256, 300, 512):
    W = rng.standard_normal((4 * n + 7, n)) * np.logspace(0, -3, n)
    G = W.T @ W
    Gd = torch.from_numpy(G).cuda()
    lam, V = T.eigh(Gd)
    lam, V = lam.cpu().numpy(), V.cpu().numpy()
    lo = np.linalg.eigvalsh(G)[::-1]
    err = np.abs(lam - lo).max() / lo[0]
    orth = np.abs(V.T @ V - np.eye(n)).max()
    res = np.abs(G @ V - V * lam).max() / lo[0]
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5):
        T.eigh(Gd)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
    print(f"n={n:4d} lam err {err:.1e} orth {orth:.1e} resid {res:.1e}  {dt*1e3:.2f} ms", flush=True)
