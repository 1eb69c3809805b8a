#!/bin/bash
# round-2 final ncu evidence (single GPU): launch lists and full captures of the hot kernels (sizes kept below the 64 MiB
# gpurun_out limit: a handful of launches per capture)
mkdir -p gpurun_out
O=gpurun_out/r02
NCU="ncu --clock-control none"
timeout 300 $NCU --metrics gpu__time_duration.sum --csv --log-file ${O}_launches_c4.csv python tools/prof_driver.py c4 1 > /dev/null 2>&1
timeout 300 $NCU --metrics gpu__time_duration.sum --csv --log-file ${O}_launches_c3.csv python tools/prof_driver.py ga 1 > /dev/null 2>&1
FULL="$NCU --set full --import-source on -f"
timeout 300 $FULL -k regex:'syrk_tma_kernel|tproj_tma_kernel|alm_ew_tma_kernel' --launch-skip 40 --launch-count 3 -o ${O}_ncu_c4_hot python tools/prof_driver.py c4 1 > /dev/null 2>&1
timeout 300 $FULL -k regex:alm_fused_kernel --launch-skip 14 --launch-count 1 -o ${O}_ncu_fused python tools/prof_driver.py c4fused 1 > /dev/null 2>&1
timeout 300 $FULL -k regex:'ga_sweep_tma_kernel' --launch-skip 6 --launch-count 2 -o ${O}_ncu_ga python tools/prof_driver.py ga 1 > /dev/null 2>&1
timeout 300 $FULL -k regex:'chol_upper_kernel|gemm_xb_kernel' --launch-count 4 -o ${O}_ncu_svd_refine python tools/prof_driver.py c4 1 > /dev/null 2>&1
timeout 300 $FULL -k regex:'jacobi_cluster_block_kernel' --launch-skip 24 --launch-count 8 -o ${O}_ncu_jacobi python tools/prof_driver.py c4 1 > /dev/null 2>&1
timeout 300 $FULL -k regex:'si_jacobi_kernel' --launch-skip 100 --launch-count 2 -o ${O}_ncu_si python tools/prof_driver.py c4 1 > /dev/null 2>&1
ls -la gpurun_out/r02_* | awk '{print $5, $9}'
du -sh gpurun_out
