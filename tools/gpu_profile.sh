#!/bin/bash
# Run on the GPU box through gpurun: GPU test-suite, headline bench, ncu launch list and full captures of the top kernels.
set -u
mkdir -p gpurun_out
TAG=${1:-m}
python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$TAG.log
tail -3 gpurun_out/pytest_gpu_$TAG.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
cat gpurun_out/bench_$TAG.json
# launch list (per-launch device time) of one short bench run: skip the warm-up solve, list one full solve
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 700 --csv \
    --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_launch_$TAG.log 2>&1
# full captures of the two dominant kernels
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'syrk_tma_kernel|alm_stream_kernel|alm_fused_kernel' -s 12 -c 4 \
    -o gpurun_out/prof_$TAG -f python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_full_$TAG.log 2>&1
ls -la gpurun_out | tail -8
