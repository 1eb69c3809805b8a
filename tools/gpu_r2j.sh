#!/bin/bash
# round-2 GPU call J: TMA-ring Grassmann sweep + run-ahead GA loop -- parity, C3 bench A/B, ncu of the sweep
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=12 -k "ga or golden or c3 or float32" > gpurun_out/r2j_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2j_pytest.log
tail -8 gpurun_out/r2j_pytest.log
timeout 600 python bench.py --workload c3 --steps 5 --warmup 3 --no-cpu > gpurun_out/r2j_bench_c3.json 2> gpurun_out/r2j_bench_c3.err; echo "bench rc=$?"
TLSQ_GA_TMA=0 TLSQ_NO_RUNAHEAD=1 timeout 600 python bench.py --workload c3 --steps 5 --warmup 3 --no-cpu > gpurun_out/r2j_bench_c3_old.json 2> gpurun_out/r2j_bench_c3_old.err; echo "bench old rc=$?"
TLSQ_GA_NO_FUSE=1 timeout 600 python bench.py --workload c3 --steps 5 --warmup 3 --no-cpu > gpurun_out/r2j_bench_c3_notma.json 2> gpurun_out/r2j_bench_c3_notma.err; echo "bench nofuse rc=$?"
timeout 300 ncu --clock-control none --set full --import-source on -f -k regex:'ga_sweep_tma_kernel' --launch-skip 6 --launch-count 1 -o gpurun_out/r2j_ga python tools/prof_driver.py ga 1 > gpurun_out/r2j_ncu.log 2>&1
python - <<'PY'
import json
for f in ("r2j_bench_c3", "r2j_bench_c3_old", "r2j_bench_c3_notma"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, "value", round(d["value"], 2), "ms/step", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 2), "roofline", round(d["roofline"]["frac"], 3), round(d["roofline"]["avg_launch_ms"], 4), d["parity"]["iters"])
    except Exception as e:
        print(f, "ERR", e)
PY
