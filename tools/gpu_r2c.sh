#!/bin/bash
# round-2 GPU call C: large-n tests first (new code), then the full suite, then C4 bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_scale.py -m gpu -q -x -k "embedding or 3000" > gpurun_out/r2c_largen.log 2>&1; echo "largen rc=$?" >> gpurun_out/r2c_largen.log
tail -30 gpurun_out/r2c_largen.log
timeout 900 python -m pytest tests -m gpu -q --maxfail=12 --deselect tests/test_gpu_scale.py::test_lowrankfilter_default_embedding_50k_samples --deselect tests/test_gpu_scale.py::test_rpca_3000_x_1000_with_returned_svd --deselect tests/test_gpu_scale.py::test_lowrankfilter_default_embedding_16k_samples_live_oracle > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
tail -15 gpurun_out/r2c_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r2c_bench_c4.json 2> gpurun_out/r2c_bench_c4.err; echo "bench c4 rc=$?"
TLSQ_NO_RUNAHEAD=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r2c_bench_c4_noahead.json 2> gpurun_out/r2c_bench_c4_noahead.err; echo "bench c4 noahead rc=$?"
python - <<'PY'
import json
for f in ("r2c_bench_c4", "r2c_bench_c4_noahead"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, "value", round(d["value"], 2), "ms/step", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 2), round(d["e2e_no_svd"]["value"], 2),
              "iterfrac", d.get("iteration_roofline", {}).get("frac"), "phases", {k: round(v, 3) for k, v in d.get("iteration_roofline", {}).get("phase_ms_per_iter", {}).items()})
    except Exception as e:
        print(f, "ERR", e)
PY
