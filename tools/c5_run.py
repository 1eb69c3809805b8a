"""BASELINE configs[4] (C5): lowrankfilter on a 50M-sample sinusoid sum with 10 % missing values, lag n = 256 -- the
implicit Hankel matrix is 49 999 745 x 256 (102.4 GB if it were materialised; the reference cannot run this shape).
Single GPU: one-pass kernel, Y updated in place, factored unhankel.  Run on the GPU box:

    python tools/c5_run.py [Ns] [--gpus via torchrun]

Prints one JSON line: time to converge, ALM iterations/s, the denoising quality (size-independent property) and the
per-phase device times.
"""
import json, math, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tlsq_b200 as T

Ns = int(sys.argv[1]) if len(sys.argv) > 1 else 50_000_000
n = 256
rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    import torch.distributed as dist
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
    T.init_distributed(local)
g = torch.Generator(device=dev); g.manual_seed(5)
t = torch.arange(1, Ns + 1, device=dev, dtype=torch.float64)
y = torch.sin(0.1 * t) + 0.5 * torch.sin(0.37 * t + 1.0) + 0.25 * torch.sin(0.013 * t + 2.0)
del t
mask = torch.rand(Ns, device=dev, generator=g) < 0.1
yn = y + 1e2 * mask.double()
del mask
torch.cuda.synchronize()
res = {}
for rep in range(2):
    T.set_profiling(rep == 1, local)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    yf, info = T.lowrankfilter(yn, n, return_info=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    res = {"seconds": dt, "iters": info["iters"], "converged": bool(info["converged"]), "sv": int(info.get("sv", 0))}
prof = T.get_profile(local)
mse = (torch.mean((y - yf) ** 2) / torch.mean(y ** 2)).item()
mse_in = (torch.mean((y - yn) ** 2) / torch.mean(y ** 2)).item()
K = Ns - n + 1
S = K * n * 8
if rank == 0:
    print(json.dumps({"workload": f"lowrankfilter Ns={Ns} n={n} (implicit Hankel {K}x{n}, {S/1e9:.1f} GB if materialised)",
                      "n_gpus": world, **res, "its_per_s": res["iters"] / res["seconds"],
                      "normalised_mse_out": mse, "normalised_mse_in": mse_in,
                      "phase_ms": {k: round(v[0], 2) for k, v in prof.items() if v[1]},
                      "mem_GB": torch.cuda.max_memory_allocated() / 1e9}))
if world > 1:
    dist.destroy_process_group()
