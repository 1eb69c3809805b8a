#!/bin/bash
# round-2 GPU call B: why is the run-ahead schedule slower?  host-side trace + variants
mkdir -p gpurun_out
cat > /tmp/trace.py <<'PY'
import os, sys, time, torch
sys.path.insert(0, ".")
import tlsq_b200 as T
dev = torch.device("cuda", 0)
D = T.synth.lowrank_sparse_cuda(0, 1_000_000, 256, dev, 10, 0.05, 4, True)
kw = dict(nonnegA=True, lam=1e-3)
for _ in range(2):
    T.rpca(D, **kw)
torch.cuda.synchronize()
for rep in range(3):
    t0 = time.perf_counter(); A, E, s, sv, info = T.rpca(D, return_info=True, **kw); torch.cuda.synchronize()
    print("solve", rep, round((time.perf_counter() - t0) * 1e3, 2), "ms", info["iters"], "iters", file=sys.stderr)
t0 = time.perf_counter(); A, E, s, sv, info = T.rpca(D, return_info=True, want_svd=False, **kw); torch.cuda.synchronize()
print("solve no-svd", round((time.perf_counter() - t0) * 1e3, 2), "ms", file=sys.stderr)
PY
for v in "" "TLSQ_NO_RUNAHEAD=1" "TLSQ_NO_RUNAHEAD=1 TLSQ_NO_SPECULATE=1" "TLSQ_NO_SVD_REFINE=1"; do
  echo "=== variant: $v" >> gpurun_out/r2b_trace.log
  env $v python /tmp/trace.py 2>> gpurun_out/r2b_trace.log
done
echo "=== TRACE (default)" >> gpurun_out/r2b_trace.log
TLSQ_TRACE=1 TLSQ_DEBUG_EIG=1 python /tmp/trace.py 2>&1 | tail -150 >> gpurun_out/r2b_trace.log
grep -E "^solve|variant" gpurun_out/r2b_trace.log
