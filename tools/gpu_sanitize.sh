#!/bin/bash
# compute-sanitizer (memcheck on every family, racecheck on the cluster / TMA-ring kernels) at small shapes
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
timeout 900 $CS --tool memcheck --print-limit 20 python tools/sanitize_driver.py fused legacy largen ga lrf plugins > gpurun_out/r2_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"
tail -5 gpurun_out/r2_sanitizer_memcheck.log
timeout 900 $CS --tool racecheck --print-limit 20 python tools/sanitize_driver.py fused legacy ga > gpurun_out/r2_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"
tail -5 gpurun_out/r2_sanitizer_racecheck.log
