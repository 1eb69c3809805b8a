"""Developer smoke script for a GPU box: exercises every kernel against torch/oracle and prints diagnostics.
Not part of the test-suite (tests/ holds the real parity tests); used to get maximum information per gpurun call."""
import os
import sys
import time
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch  # noqa: E402

import tlsq_b200 as T  # noqa: E402
import tls_oracle as O  # noqa: E402

dev = torch.device("cuda", 0)


def colmajor(a):
    """numpy (M,N) -> CUDA tensor with column-major storage"""
    return torch.from_numpy(np.ascontiguousarray(a.T)).to(dev).t()


def section(name):
    print(f"\n=== {name} ===", flush=True)


def run(fn):
    try:
        fn()
    except Exception:
        traceback.print_exc()
    sys.stdout.flush()


def t_gram():
    section("gram")
    rng = np.random.default_rng(0)
    for (M, n) in [(33, 5), (1000, 5), (5000, 40), (777, 64), (20000, 256), (3001, 300), (4000, 512), (100000, 256)]:
        X = rng.standard_normal((M, n))
        Xd = colmajor(X)
        G = T.gram(Xd)
        ref = Xd.t() @ Xd
        err = (G - ref).abs().max().item() / ref.abs().max().item()
        sym = (G - G.t()).abs().max().item()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            T.gram(Xd)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 3
        print(f"M={M:7d} n={n:4d} relerr={err:.2e} asym={sym:.1e} time={dt * 1e3:.3f} ms "
              f"({M * n * (n + 1) / dt * 1e-12:.2f} TFLOP/s syrk-count)")


def t_eigh():
    section("eigh")
    rng = np.random.default_rng(1)
    for n in [2, 5, 16, 17, 40, 64, 100, 128, 200, 256, 300, 512]:
        for kind in ["gauss", "lowrank"]:
            if kind == "gauss":
                X = rng.standard_normal((4 * n + 3, n))
            else:
                r = max(1, n // 8)
                X = rng.standard_normal((4 * n + 3, r)) @ rng.standard_normal((r, n)) + 1e-6 * rng.standard_normal((4 * n + 3, n))
            G = X.T @ X
            Gd = torch.from_numpy(G).to(dev)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            lam, V = T.eigh(Gd)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            ref = torch.linalg.eigvalsh(Gd).flip(0)
            e_l = (lam - ref).abs().max().item() / ref.abs().max().item()
            res = (Gd @ V - V * lam).abs().max().item() / ref.abs().max().item()
            orth = (V.t() @ V - torch.eye(n, device=dev, dtype=torch.float64)).abs().max().item()
            print(f"n={n:4d} {kind:8s} lam_err={e_l:.2e} resid={res:.2e} orth={orth:.2e} time={dt * 1e3:.2f} ms")


def relF(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def t_rpca_small():
    section("rpca golden 5x5")
    import json
    g = json.load(open(os.path.join(ROOT, "tests/golden/rpca_5x5.json")))
    D = np.array(g["D"])
    A, E, s, sv, info = T.rpca(D, nonnegA=True, nonnegE=True, return_info=True)
    print("iters", info["iters"], "sv", sv, "maxerr A", np.abs(A - np.array(g["A"])).max(), "E",
          np.abs(E - np.array(g["E"])).max(), "resid", np.linalg.norm(D - A - E) / np.linalg.norm(D))
    ro = O.rpca(D, nonnegA=True, nonnegE=True)
    print("oracle iters", ro.iters, "relF A", relF(A, ro.A), "relF E", relF(E, ro.E))
    print("hist gpu", info["hist"][-3:].tolist(), "oracle", ro.hist[-3:].tolist())
    W = (s.U * s.S) @ s.Vt
    print("svd recon vs oracle svd recon", relF(W, (ro.s.U * ro.s.S) @ ro.s.Vt), "S", s.S, ro.s.S)


def synth_lr(M, N, r, seed, frac=0.05, nonneg=False):
    rng = np.random.default_rng(seed)
    G1 = rng.standard_normal((M, r))
    G2 = rng.standard_normal((r, N))
    if nonneg:
        G1, G2 = np.abs(G1), np.abs(G2)
    L = G1 @ G2
    mask = rng.random((M, N)) < frac
    mag = 10 * np.sqrt(10.0)
    S = (rng.random((M, N)) if nonneg else rng.uniform(-1, 1, (M, N))) * mag * mask
    return np.asfortranarray(L + S)


def t_rpca_parity():
    section("rpca parity vs oracle (fixed iterations, tol=0)")
    for (M, N, r, kw, its) in [(2000, 64, 5, {}, 12), (3000, 40, 3, {"nonnegA": True}, 12),
                               (496, 5, 2, {}, 20), (5000, 256, 10, {}, 10), (5000, 256, 10, {"nonnegA": True, "nonnegE": True}, 10),
                               (300, 500, 4, {}, 8), (2000, 320, 6, {"nukeA": False}, 6)]:
        D = synth_lr(M, N, r, seed=M + N, nonneg=bool(kw.get("nonnegA")))
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            t0 = time.perf_counter()
            A, E, s, sv, info = T.rpca(D, iters=its, tol=0.0, return_info=True, **kw)
            dt = time.perf_counter() - t0
            ro = O.rpca(D, iters=its, tol=0.0, **kw)
        supp = int(np.sum((E != 0) != (ro.E != 0)))
        print(f"{M}x{N} r={r} {kw} its={its}: relF A={relF(A, ro.A):.2e} E={relF(E, ro.E):.2e} supp_mismatch={supp} "
              f"sv={sv}/{ro.sv} svp_hist_equal={np.array_equal(info['hist'][:, 1], ro.hist[:, 1])} "
              f"S={relF(s.S, ro.s.S):.1e} time={dt:.3f}s")
        if not np.array_equal(info['hist'][:, 1], ro.hist[:, 1]):
            print("   svp gpu", info['hist'][:, 1].tolist(), "oracle", ro.hist[:, 1].tolist())
    section("rpca converge vs oracle (default tol)")
    for (M, N, r, kw) in [(2000, 64, 5, {}), (5000, 256, 10, {"nonnegA": True})]:
        D = synth_lr(M, N, r, seed=7, nonneg=bool(kw.get("nonnegA")))
        A, E, s, sv, info = T.rpca(D, return_info=True, **kw)
        ro = O.rpca(D, **kw)
        print(f"{M}x{N}: iters gpu={info['iters']} oracle={ro.iters} relF A={relF(A, ro.A):.2e} E={relF(E, ro.E):.2e} "
              f"cost_last gpu={info['hist'][-1, 2]:.3e} oracle={ro.hist[-1, 2]:.3e}")
        Ae, Ee, _, _, infoe = T.rpca(D, return_info=True, exact_cost=True, **kw)
        print("   exact-cost hist max rel diff vs oracle:",
              np.abs(infoe['hist'][:, 2] - ro.hist[:infoe['iters'], 2]).max() / ro.hist[:, 2].max() if infoe['iters'] == ro.iters else "iters differ")


def t_lowrank():
    section("lowrankfilter README example")
    rng = np.random.default_rng(0)
    N = 500
    y = np.sin(0.1 * np.arange(1, N + 1))
    miss = rng.random(N) < 0.1
    yn = y + miss * 1e2 + 0.1 * rng.standard_normal(N)
    yf, info = T.lowrankfilter(yn, 40, return_info=True)
    yo = O.lowrankfilter(yn, 40)
    print("iters", info["iters"], "nmse", np.mean((y - yf) ** 2) / np.mean(y ** 2), "vs oracle rel", relF(yf, yo))
    H = T.hankel(np.arange(1.0, 21.0), 3, 2)
    print("hankel ok", np.array_equal(H, O.hankel(np.arange(1.0, 21.0), 3, 2)),
          "unhankel ok", np.allclose(T.unhankel(O.hankel(y, 5)), y),
          "lag2", np.allclose(T.unhankel(O.hankel(y, 5, 2), 2, N)[:-1], y[:-1]))
    yf2 = T.lowrankfilter(yn, 40, lag=3)
    yo2 = O.lowrankfilter(yn, 40, lag=3)
    print("lag=3 rel", relF(yf2, yo2))


def t_ga():
    section("rpca_ga vs oracle")
    rng = np.random.default_rng(3)
    for (d, N, r) in [(10, 40, 3), (40, 10, 5), (1000, 256, 4), (5000, 300, 3), (777, 1000, 2)]:
        rr = min(d, N, 10)
        X = (rng.standard_normal((d, rr)) * np.arange(1, rr + 1)) @ rng.standard_normal((rr, N)) + 0.01 * rng.standard_normal((d, N))
        out = rng.random(N) < 0.1
        X[:, out] += 100 * rng.standard_normal((d, int(out.sum())))
        q0 = rng.standard_normal((d, r))
        Q, info = T.rpca_ga(X, r, q0=q0, return_info=True)
        Qo, its = O.rpca_ga(X, r, q0=q0, exact_order=False, return_iters=True)
        sgn = np.sign(np.sum(Q * Qo, axis=0))
        print(f"d={d} N={N} r={r}: iters gpu={info['iters']} oracle={its} max|Q - s Qo|={np.abs(Q * sgn - Qo).max():.2e} "
              f"orth={np.abs(Q.T @ Q - np.eye(r)).max():.2e}")


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0))
    which = sys.argv[1:] or ["gram", "eigh", "rpca_small", "rpca_parity", "lowrank", "ga"]
    for w in which:
        run(globals()["t_" + w])
