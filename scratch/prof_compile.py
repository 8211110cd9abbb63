"""Time the fused key compile (CSR out) of one conv layer: usage C M U [perm]"""
import sys
sys.path.insert(0, '.')
import numpy as np, torch
from keynet_b200 import sparse
from keynet_b200.sparse import MonomialKey
(C, M, U) = [int(v) for v in sys.argv[1:4]]
perm = len(sys.argv) > 4
rs = np.random.RandomState(0)
f = (rs.randn(M, C, 3, 3) * 0.05).astype(np.float32); b = rs.randn(M).astype(np.float32)
K = C * U * U + 1
Ainv = MonomialKey(np.concatenate([rs.permutation(K - 1), [K - 1]]) if perm else np.arange(K))
for rep in range(3):
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    W = sparse.keyed_toeplitz_conv2d((C, U, U), f, b, 1, None, Ainv, build_groups=False)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    nnz = W.nnz()
    del W
print('C=%d M=%d U=%d perm=%s: nnz %.1f M, %.2f ms (whole call), %.0f GB/s of CSR written' % (C, M, U, perm, nnz / 1e6, ms, nnz * 8 / ms / 1e6))
