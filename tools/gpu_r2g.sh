#!/bin/bash
# round-2 GPU call G: TMA-staged streaming kernels v2 (512 threads) -- quick parity subset, A/B bench, ncu
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_scale.py tests/test_gpu_parity.py -m gpu -q --maxfail=12 -k "c4_shaped or runahead or inplace or fixed_iterations or converges or fused_pipeline or lowrankfilter_n256" > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2g_pytest.log
tail -8 gpurun_out/r2g_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r2g_bench_c4.json 2> gpurun_out/r2g_bench_c4.err; echo "bench rc=$?"
TLSQ_STREAM_TMA=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r2g_bench_c4_notma.json 2> gpurun_out/r2g_bench_c4_notma.err; echo "bench notma rc=$?"
timeout 300 ncu --clock-control none --set full --import-source on -f -k regex:'alm_ew_tma_kernel|tproj_tma_kernel' --launch-skip 30 --launch-count 2 -o gpurun_out/r2g_streamtma python tools/prof_driver.py c4 1 > gpurun_out/r2g_ncu.log 2>&1
python - <<'PY'
import json
for f in ("r2g_bench_c4", "r2g_bench_c4_notma"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, "value", round(d["value"], 2), "ms/step", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 2), round(d["e2e_no_svd"]["value"], 2),
              "iterfrac", round(d["iteration_roofline"]["frac"], 3), "phases", {k: round(v, 3) for k, v in d["iteration_roofline"]["phase_ms_per_iter"].items()})
    except Exception as e:
        print(f, "ERR", e)
PY
