"""Multi-GPU parity check (run under torchrun): every rank solves its row shard; shards are compared with the CPU
oracle run on the full matrix."""
import os, sys, warnings
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import torch.distributed as dist
import tlsq_b200 as T
import tls_oracle as O

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
T.init_distributed(lr)
relF = lambda a, b: float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))
ok = True
for (M, N, r, kw, its) in [(16384, 256, 10, {"nonnegA": True}, 10), (16384, 256, 10, {"nonnegA": True, "fused": 1}, 10),
                           (9000, 128, 5, {}, 10), (3000, 40, 3, {}, 8)]:
    kw = dict(kw)
    os.environ["TLSQ_FUSED"] = "1" if kw.pop("fused", 0) else "0"       # n = 256: both per-iteration pipelines
    D = T.synth.lowrank_sparse_np(M, N, r, 0.05, seed=M + N, nonneg=bool(kw.get("nonnegA")))
    r0, r1 = T.synth.shard_rows(M, world, rank, align=2)
    Dl = torch.from_numpy(np.ascontiguousarray(D[r0:r1].T)).to(dev).t()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        A, E, s, sv, info = T.rpca(Dl, iters=its, tol=0.0, return_info=True, **kw)
        ref = O.rpca(D, iters=its, tol=0.0, **kw)
    ea, ee = relF(A.cpu().numpy(), ref.A[r0:r1]), relF(E.cpu().numpy(), ref.E[r0:r1])
    eu = relF((s.U.cpu().numpy() * s.S.cpu().numpy()) @ s.Vt.cpu().numpy(), ((ref.s.U * ref.s.S) @ ref.s.Vt)[r0:r1])
    good = ea < 1e-9 and ee < 1e-9 and sv == ref.sv and np.array_equal(info["hist"][:, 1], ref.hist[:, 1])
    ok &= good
    print(f"[rank {rank}] rpca {M}x{N} rows[{r0},{r1}) relF A={ea:.1e} E={ee:.1e} svdrecon={eu:.1e} sv={sv}/{ref.sv} ok={good}", flush=True)
# converged run: same iteration count on every rank as the oracle
D = T.synth.lowrank_sparse_np(20000, 256, 10, 0.05, seed=4, nonneg=True)
r0, r1 = T.synth.shard_rows(20000, world, rank, align=2)
Dl = torch.from_numpy(np.ascontiguousarray(D[r0:r1].T)).to(dev).t()
ref = O.rpca(D, nonnegA=True)
for fused in ("0", "1"):
    os.environ["TLSQ_FUSED"] = fused
    A, E, s, sv, info = T.rpca(Dl, nonnegA=True, return_info=True, want_svd=False)
    good = info["iters"] == ref.iters and relF(A.cpu().numpy(), ref.A[r0:r1]) < 1e-9
    ok &= good
    print(f"[rank {rank}] converge (fused={fused}) iters {info['iters']}/{ref.iters} ok={good}", flush=True)
del os.environ["TLSQ_FUSED"]
# Grassmann averages, d sharded
X, q0 = T.synth.ga_data_np(8000, 200, 6, seed=3)
a0, a1 = T.synth.shard_rows(8000, world, rank)
Xl = torch.from_numpy(np.ascontiguousarray(X[a0:a1].T)).to(dev).t()
ql = torch.from_numpy(np.ascontiguousarray(q0[a0:a1, :3].T)).to(dev).t()
Q, inf = T.rpca_ga(Xl, 3, q0=ql, return_info=True)
Qo, its = O.rpca_ga(X, 3, q0=q0[:, :3], exact_order=False, return_iters=True)
Qn = Q.cpu().numpy()
sg = np.sign(np.sum(Qo[a0:a1] * Qn, axis=0)); sg[sg == 0] = 1
good = np.abs(Qn * sg - Qo[a0:a1]).max() < 1e-9 and inf["iters"] == its
ok &= good
print(f"[rank {rank}] rpca_ga iters {inf['iters']}/{its} maxdiff {np.abs(Qn * sg - Qo[a0:a1]).max():.1e} ok={good}", flush=True)
# lowrankfilter: replicated signal, sharded Hankel rows
y, yn = T.synth.sinusoid_np(4000, seed=2, noise=0.05)
yd = torch.from_numpy(yn).to(dev)
for (nn, lag) in [(100, 1), (37, 3)]:
    yf, inf = T.lowrankfilter(yd, nn, lag=lag, return_info=True)
    yo = O.lowrankfilter(yn, nn, lag=lag)
    err = relF(yf.cpu().numpy(), yo)
    good = err < 1e-9
    ok &= good
    print(f"[rank {rank}] lowrankfilter n={nn} lag={lag} iters {inf['iters']} rel {err:.1e} ok={good}", flush=True)
# n = 256: factored unhankel, one-pass kernel on the sharded implicit Hankel matrix (when the shard is tall enough)
y, yn = T.synth.sinusoid_np(24831, seed=7, noise=0.05)
yd = torch.from_numpy(yn).to(dev)
yo = O.lowrankfilter(yn, 256)
for fused in ("0", "1"):
    os.environ["TLSQ_FUSED"] = fused
    yf, inf = T.lowrankfilter(yd, 256, return_info=True)
    err = relF(yf.cpu().numpy(), yo)
    good = err < 1e-9
    ok &= good
    print(f"[rank {rank}] lowrankfilter n=256 (fused={fused}) iters {inf['iters']} rel {err:.1e} ok={good}", flush=True)
del os.environ["TLSQ_FUSED"]
# large embedding (n = 600 > 512), sharded Hankel rows: large-n eigen path + column-chunked epilogue under NCCL
y, yn = T.synth.sinusoid_np(12001, seed=8, noise=0.02)
yd = torch.from_numpy(yn).to(dev)
yf, inf = T.lowrankfilter(yd, 600, return_info=True)
yo = O.lowrankfilter(yn, 600)
err = relF(yf.cpu().numpy(), yo)
good = err < 1e-9
ok &= good
print(f"[rank {rank}] lowrankfilter n=600 iters {inf['iters']} rel {err:.1e} ok={good}", flush=True)
# host-buffer entry on a shard with fewer local rows than columns (M_local < N <= M_global)
Mg, Ng = 200 * world, 256
D = T.synth.lowrank_sparse_np(Mg, Ng, 4, 0.05, seed=77)
r0, r1 = T.synth.shard_rows(Mg, world, rank)
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    A, E, s, sv, info = T.rpca(np.asfortranarray(D[r0:r1]), iters=6, tol=0.0, return_info=True)
    ref = O.rpca(D, iters=6, tol=0.0)
good = relF(A, ref.A[r0:r1]) < 1e-9 and relF(E, ref.E[r0:r1]) < 1e-9 and s.U.shape == (r1 - r0, Ng) and s.Vt.shape == (Ng, Ng) \
    and np.allclose(s.S, ref.s.S, rtol=0, atol=1e-12 * ref.s.S[0])
ok &= bool(good)
print(f"[rank {rank}] host-path shard {r1 - r0}x{Ng} of {Mg}x{Ng}: relF A={relF(A, ref.A[r0:r1]):.1e} ok={good}", flush=True)
t = torch.tensor([1 if ok else 0], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print("MGPU_CHECK", "PASS" if t.item() == 1 else "FAIL", flush=True)
dist.destroy_process_group()
