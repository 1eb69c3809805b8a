"""world_size-2 gloo tests (CPU) of the row-sharded decomposition the multi-GPU path relies on:
every rank holds a contiguous row shard; only the n x n Gram, length-N vectors and scalars are all-reduced.
A NumPy model of one sharded ALM iteration / Grassmann sweep must reproduce the single-process oracle."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch
    import torch.distributed as dist
    import tls_oracle as O
    import tlsq_b200 as T
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        M, N = 301, 12
        D = T.synth.lowrank_sparse_np(M, N, 3, 0.1, seed=9)
        r0, r1 = T.synth.shard_rows(M, world, rank)
        Dl = D[r0:r1]
        # --- sharded Gram + all-reduce == full Gram --------------------------------------------------------
        G = torch.from_numpy(Dl.T @ Dl)
        dist.all_reduce(G)
        assert np.allclose(G.numpy(), D.T @ D, rtol=1e-13, atol=1e-11)
        # --- one sharded ALM iteration (Gram -> eig -> local epilogue) vs the oracle ------------------------
        lam = 1.0 / np.sqrt(max(M, N))
        norm2 = np.sqrt(np.linalg.eigvalsh(G.numpy())[-1])
        mx = torch.tensor([np.abs(Dl).max()])
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dual = max(norm2, mx.item() / lam)
        mu = 1.25 / norm2
        Yl, Al = Dl / dual, np.zeros_like(Dl)
        El = O.soft_th((Dl - Al) + Yl / mu, lam / mu)
        Wl = (Dl - El) + Yl / mu
        Gw = torch.from_numpy(Wl.T @ Wl)
        dist.all_reduce(Gw)
        w, V = np.linalg.eigh(Gw.numpy())
        sig = np.sqrt(np.maximum(w[::-1], 0))
        V = V[:, ::-1]
        svp = int(np.sum(sig >= 1 / mu))
        A1 = (Wl @ V[:, :svp]) * ((sig[:svp] - 1 / mu) / sig[:svp]) @ V[:, :svp].T
        ref = O.rpca(D, iters=1, tol=0.0)
        assert svp == int(ref.hist[0, 1])
        assert np.allclose(A1, ref.A[r0:r1], rtol=0, atol=1e-10 * np.abs(ref.A).max())
        assert np.array_equal(El != 0, ref.E[r0:r1] != 0)
        zz = torch.tensor([np.sum((Dl - A1 - El) ** 2)])
        dist.all_reduce(zz)
        assert np.isclose(zz.item(), np.sum((D - ref.A - ref.E) ** 2), rtol=1e-8)
        # --- one sharded Grassmann sweep: (N+1)-vector all-reduce -------------------------------------------
        X, q0 = T.synth.ga_data_np(200, 16, 3, seed=10)
        a0, a1 = T.synth.shard_rows(200, world, rank)
        Xl, ql = X[a0:a1], q0[a0:a1, 0]
        ss = torch.tensor([ql @ ql]); dist.all_reduce(ss)
        ql = ql / np.sqrt(ss.item())
        n2 = torch.from_numpy((Xl * Xl).sum(axis=0)); dist.all_reduce(n2)
        t = torch.from_numpy(Xl.T @ ql); dist.all_reduce(t)
        s = np.sign(t.numpy())
        sumw = float(s @ np.sqrt(n2.numpy()))
        mul = (Xl @ s) / sumw
        pack = torch.from_numpy(np.concatenate([Xl.T @ mul, [mul @ mul]])); dist.all_reduce(pack)
        qn = mul / np.sqrt(pack[-1].item())
        # oracle: one iteration of rpca_ga_1
        norms = np.sqrt((X * X).sum(axis=0))
        qo, _ = O.rpca_ga_1(norms, X / norms, np.zeros(16), q0[:, 0], iters=1, exact_order=False)
        assert np.allclose(qn, qo[a0:a1], rtol=0, atol=1e-13)
        # --- pipeline vote of a sharded solve: ranks with DIFFERENT local capabilities must end up with the same
        #     choice (solver.cu sums {M, can two-kernel, can one-pass, wants one-pass, wants in-place Y} with NCCL) -----
        import ctypes
        lib = T.load()
        for local, expect in [
            # rank 0 owns an even shard (both pipelines), rank 1 an odd one (one-pass kernel only)
            (([1, 1, 0, 0], [0, 1, 0, 0]), (1, 0, 0)),
            # both can do both, rank 1 is short of memory -> one-pass kernel and in-place Y on BOTH ranks
            (([1, 1, 0, 0], [1, 1, 1, 1]), (1, 0, 1)),
            # rank 1 cannot run the one-pass kernel (n != 256 shard too short ...) -> two-kernel pipeline on both
            (([1, 1, 1, 0], [1, 0, 0, 0]), (0, 1, 0)),
        ]:
            v = torch.tensor(local[rank], dtype=torch.float64)
            dist.all_reduce(v)
            f, w, ip = ctypes.c_int(-1), ctypes.c_int(-1), ctypes.c_int(-1)
            vv = (ctypes.c_double * 4)(*v.tolist())
            assert lib.tlsq_plan_pipeline(world, vv, -1, ctypes.byref(f), ctypes.byref(w), ctypes.byref(ip)) == 0
            assert (f.value, w.value, ip.value) == expect, (rank, local, (f.value, w.value, ip.value))
        # Hankel row shards of the two ranks tile [0, K) and agree on the "last shard" every rank tests
        K = 49_999_745
        r0, kl = ctypes.c_int64(), ctypes.c_int64()
        assert lib.tlsq_plan_hankel_shard(K, world, rank, ctypes.byref(r0), ctypes.byref(kl)) == 0
        span = torch.tensor([float(r0.value), float(kl.value)], dtype=torch.float64)
        both = [torch.zeros(2, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(both, span)
        assert both[0][0].item() == 0 and both[0][1].item() == both[1][0].item()
        assert both[1][0].item() + both[1][1].item() == K and both[0][1].item() % 32 == 0
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_row_sharded_model_world2():
    import warnings
    warnings.simplefilter("ignore")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res
