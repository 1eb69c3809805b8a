#!/bin/bash
# round-2 GPU call H: plugin / Float32 tests, full suite, C4 bench with the D2H overlap, trace of one solve
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_plugins.py -m gpu -q --maxfail=6 > gpurun_out/r2h_plugins.log 2>&1; echo "plugins rc=$?" >> gpurun_out/r2h_plugins.log
tail -30 gpurun_out/r2h_plugins.log
timeout 1200 python -m pytest tests -m gpu -q --maxfail=12 --deselect tests/test_gpu_plugins.py > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2h_pytest.log
tail -8 gpurun_out/r2h_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2h_bench_c4.json 2> gpurun_out/r2h_bench_c4.err; echo "bench rc=$?"
TLSQ_TRACE=1 TLSQ_DEBUG_EIG=1 timeout 300 python tools/prof_driver.py c4 2 2>&1 | grep -E "tlsq" | tail -60 > gpurun_out/r2h_trace.log
python - <<'PY'
import json
for f in ("r2h_bench_c4",):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, "value", round(d["value"], 2), "ms/step", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 2), round(d["e2e_no_svd"]["value"], 2),
              "iterfrac", round(d["iteration_roofline"]["frac"], 3), "phases", {k: round(v, 3) for k, v in d["iteration_roofline"]["phase_ms_per_iter"].items()})
        print("cpu", d.get("cpu_baseline"))
    except Exception as e:
        print(f, "ERR", e)
PY
