python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "converges_like" 2>&1 | grep -E "assert|Error|error|relF|iters" | head -20
TLSQ_DEBUG_EIG=1 python - <<'PY' 2>&1 | tail -40 | cut -c1-170
import sys; sys.path.insert(0,'.'); sys.path.insert(0,'oracle')
import numpy as np, tlsq_b200 as T, tls_oracle as O
D = T.synth.lowrank_sparse_np(12000, 128, 20, 0.05, seed=4)
A,E,s,sv,info = T.rpca(D, return_info=True)
ref = O.rpca(D)
print(info['iters'], ref.iters, np.linalg.norm(A-ref.A)/np.linalg.norm(ref.A), info['hist'][:,1].tolist(), ref.hist[:,1].tolist())
PY
