#!/bin/bash
# round-2 GPU call A: full GPU test-suite, C4 bench (run-ahead on/off), C5 bench on one GPU
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --maxfail=12 > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -25 gpurun_out/r2a_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2a_bench_c4.json 2> gpurun_out/r2a_bench_c4.err; echo "bench c4 rc=$?"
TLSQ_NO_RUNAHEAD=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/r2a_bench_c4_sync.json 2> gpurun_out/r2a_bench_c4_sync.err; echo "bench c4 sync rc=$?"
timeout 900 python bench.py --workload c5 --steps 2 --warmup 1 --no-cpu > gpurun_out/r2a_bench_c5.json 2> gpurun_out/r2a_bench_c5.err; echo "bench c5 rc=$?"
python - <<'PY'
import json
for f in ("r2a_bench_c4", "r2a_bench_c4_sync", "r2a_bench_c5"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, "value", round(d["value"], 2), "ms/step", round(d["ms_per_step"], 2), "e2e", d["e2e"]["value"],
              "iterfrac", d.get("iteration_roofline", {}).get("frac"), "phases", {k: round(v, 3) for k, v in d.get("iteration_roofline", {}).get("phase_ms_per_iter", {}).items()})
        print("   parity", d.get("parity"))
    except Exception as e:
        print(f, "ERR", e)
PY
