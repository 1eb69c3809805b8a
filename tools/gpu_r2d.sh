#!/bin/bash
# round-2 GPU call D (2 GPUs): multi-GPU parity + benches at N = 2
mkdir -p gpurun_out
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 900 $TR tools/mgpu_check.py > gpurun_out/r2d_mgpu_check_n$N.log 2>&1; echo "mgpu rc=$?"
grep -E "MGPU_CHECK|ok=False|Error|error" gpurun_out/r2d_mgpu_check_n$N.log | head -20
timeout 600 $TR bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2d_bench_c4_n$N.json 2> gpurun_out/r2d_bench_c4_n$N.err; echo "c4 rc=$?"
timeout 600 $TR bench.py --gpus $N --workload c3 --steps 3 --warmup 2 > gpurun_out/r2d_bench_c3_n$N.json 2> gpurun_out/r2d_bench_c3_n$N.err; echo "c3 rc=$?"
timeout 900 $TR bench.py --gpus $N --workload c5 --steps 2 --warmup 1 > gpurun_out/r2d_bench_c5_n$N.json 2> gpurun_out/r2d_bench_c5_n$N.err; echo "c5 rc=$?"
python - <<PY
import json
for f in ("r2d_bench_c4_n$N", "r2d_bench_c3_n$N", "r2d_bench_c5_n$N"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, "value", round(d["value"], 2), "ms/step", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 2),
              "phases", {k: round(v, 3) for k, v in d.get("iteration_roofline", {}).get("phase_ms_per_iter", {}).items()})
        print("   parity", d.get("parity"))
    except Exception as e:
        print(f, "ERR", e)
PY
