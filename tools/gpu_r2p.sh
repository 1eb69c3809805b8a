#!/bin/bash
# round-2 profiling call: launch list of one C4 solve + full captures (with source) of the hot kernels
mkdir -p gpurun_out
NCU="ncu --clock-control none"
timeout 300 $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r2p_launches_c4.csv python tools/prof_driver.py c4 1 > gpurun_out/r2p_launch.log 2>&1
FULL="$NCU --set full --import-source on -f"
timeout 300 $FULL -k regex:alm_fused_kernel --launch-skip 14 --launch-count 1 -o gpurun_out/r2p_fused python tools/prof_driver.py c4fused 1 > gpurun_out/r2p_fused.log 2>&1
timeout 300 $FULL -k regex:alm_stream_kernel --launch-skip 30 --launch-count 2 -o gpurun_out/r2p_stream python tools/prof_driver.py c4 1 > gpurun_out/r2p_stream.log 2>&1
timeout 300 $FULL -k regex:syrk_tma_kernel --launch-skip 14 --launch-count 1 -o gpurun_out/r2p_syrk python tools/prof_driver.py c4 1 > gpurun_out/r2p_syrk.log 2>&1
timeout 300 $FULL -k regex:'jacobi_cluster_block_kernel|chol_upper_kernel|gemm_xb_kernel' --launch-count 6 -o gpurun_out/r2p_final python tools/prof_driver.py c4 1 > gpurun_out/r2p_final.log 2>&1
timeout 300 $FULL -k regex:'si_jacobi_kernel' --launch-skip 40 --launch-count 2 -o gpurun_out/r2p_si python tools/prof_driver.py c4 1 > gpurun_out/r2p_si.log 2>&1
timeout 300 $FULL -k regex:ga_sweep_kernel --launch-skip 6 --launch-count 1 -o gpurun_out/r2p_ga python tools/prof_driver.py ga 1 > gpurun_out/r2p_ga.log 2>&1
ls -la gpurun_out/r2p_*
