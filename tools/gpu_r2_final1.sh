#!/bin/bash
# round-2 final single-GPU evidence: full GPU test-suite, smoke, bench lines of all four workloads, reference arm,
# host trace, ncu launch list and full captures of the hot kernels
mkdir -p gpurun_out
O=gpurun_out/r02
timeout 1500 python -m pytest tests -m gpu -q --maxfail=12 > ${O}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> ${O}_pytest_gpu.log
tail -6 ${O}_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > ${O}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 ${O}_smoke.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > ${O}_bench_c4_n1.json 2> ${O}_bench_c4_n1.err; echo "c4 rc=$?"
timeout 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > ${O}_ref_c4.json 2> ${O}_ref_c4.err; echo "ref rc=$?"
timeout 600 python bench.py --workload c2 --steps 5 --warmup 3 > ${O}_bench_c2_n1.json 2> ${O}_bench_c2_n1.err; echo "c2 rc=$?"
timeout 600 python bench.py --workload c3 --steps 5 --warmup 3 > ${O}_bench_c3_n1.json 2> ${O}_bench_c3_n1.err; echo "c3 rc=$?"
timeout 900 python bench.py --workload c5 --steps 2 --warmup 1 > ${O}_bench_c5_n1.json 2> ${O}_bench_c5_n1.err; echo "c5 rc=$?"
TLSQ_TRACE=1 TLSQ_DEBUG_EIG=1 timeout 300 python tools/prof_driver.py c4 2 2>&1 | grep -E "tlsq" | tail -62 > ${O}_trace_c4.log
python - <<'PY'
import json
for f in ("bench_c4_n1", "ref_c4", "bench_c2_n1", "bench_c3_n1", "bench_c5_n1"):
    try:
        d = json.load(open(f"gpurun_out/r02_{f}.json"))
        print(f, "value", round(d["value"], 3), "ms/step", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 3),
              "roof", d.get("roofline", {}).get("frac"), "iterfrac", d.get("iteration_roofline", {}).get("frac"))
    except Exception as e:
        print(f, "ERR", e)
PY
ls -la gpurun_out/r02_* | awk '{print $5, $9}'
