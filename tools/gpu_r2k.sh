#!/bin/bash
# round-2 GPU call K: cooperative block Jacobi at 256 < n <= 512 -- parity at n = 512 / 384 / 300, C2 bench A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=12 -k "512 or c2_full or factored_unhankel or 5001 or robust_averages" > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2k_pytest.log
tail -8 gpurun_out/r2k_pytest.log
python - <<'PY' > gpurun_out/r2k_eigh.log 2>&1
import sys, time, torch
sys.path.insert(0, ".")
import tlsq_b200 as T
for n in (257, 300, 384, 512):
    torch.manual_seed(n)
    X = torch.randn(4000, n, dtype=torch.float64, device="cuda")
    X[:, :5] *= 100.0
    G = X.t() @ X
    torch.cuda.synchronize(); t0 = time.perf_counter()
    lam, V = T.eigh(G.t().contiguous().t())
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    ref = torch.linalg.eigvalsh(G).flip(0)
    err = ((lam - ref).abs().max() / ref[0]).item()
    orth = (V.t() @ V - torch.eye(n, dtype=torch.float64, device="cuda")).abs().max().item()
    res = (G @ V - V * lam).abs().max().item() / ref[0].item()
    print(f"n={n}: {dt*1e3:.2f} ms  eig err {err:.1e}  orth {orth:.1e}  resid {res:.1e}")
PY
cat gpurun_out/r2k_eigh.log
timeout 600 python bench.py --workload c2 --steps 5 --warmup 3 --no-cpu > gpurun_out/r2k_bench_c2.json 2> gpurun_out/r2k_bench_c2.err; echo "c2 rc=$?"
TLSQ_JACOBI_GLOBAL=1 timeout 600 python bench.py --workload c2 --steps 5 --warmup 3 --no-cpu > gpurun_out/r2k_bench_c2_old.json 2> gpurun_out/r2k_bench_c2_old.err; echo "c2 old rc=$?"
python - <<'PY'
import json
for f in ("r2k_bench_c2", "r2k_bench_c2_old"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, "value", round(d["value"], 2), "ms/step", round(d["ms_per_step"], 2), "iters", d["iters_per_step"], "phases", {k: round(v, 3) for k, v in d["iteration_roofline"]["phase_ms_per_iter"].items()})
    except Exception as e:
        print(f, "ERR", e)
PY
