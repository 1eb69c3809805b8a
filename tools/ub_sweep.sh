#!/bin/bash
# A/B of the streaming epilogue's load batch (run on the GPU box): rebuild stream.o with -DTLSQ_STREAM_UB=n, relink, bench
cd totalleastsquares.jl_b200/csrc
for ub in 8 6 4; do
  nvcc -c stream.cu -o stream.o -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --cudart static -gencode arch=compute_100a,code=sm_100a -DTLSQ_STREAM_UB=$ub -Xptxas -v 2>&1 | grep -A1 "alm_stream_kernelILi12ELb0ELb1ELi2" | grep registers
  nvcc -shared -o ../libtlsq_b200.so gram.o syrk_tma.o eig.o eig_fast.o epilogue.o stream.o fused.o gemm.o elementwise.o ga.o solver.o --cudart static -ldl -lpthread -lrt -gencode arch=compute_100a,code=sm_100a
  (cd ../.. && python bench.py --steps 3 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('UB=$ub', round(d['value'],1), d['iteration_roofline']['phase_ms_per_iter']['epilogue'])")
done
