"""One conv layer (C -> M at U x U, permutation-free keys) at batch N through spmm: target for ncu captures of the tensor-core kernels.
usage: python scratch/prof_conv.py C M U N [stride]      (KEYNET_B200_TILES=0 selects the per-pixel kernel)"""
import sys, time
sys.path.insert(0, '.')
import numpy as np, torch
from keynet_b200 import sparse
from keynet_b200.sparse import MonomialKey
(C, M, U, N) = [int(v) for v in sys.argv[1:5]]
stride = int(sys.argv[5]) if len(sys.argv) > 5 else 1
rs = np.random.RandomState(0)
f = (rs.randn(M, C, 3, 3) * 0.05).astype(np.float32); b = rs.randn(M).astype(np.float32)
K = C * U * U + 1
W = sparse.keyed_toeplitz_conv2d((C, U, U), f, b, stride, None, MonomialKey(np.arange(K)), want_csr=False)
X = torch.randn(K, N, device='cuda'); X[-1] = 1
Y = torch.empty((W.shape[0], N), device='cuda')
for _ in range(3):
    sparse.spmm(W, X, relu=True, out=Y)
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    sparse.spmm(W, X, relu=True, out=Y)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
fl = 2.0 * W.nnz() * N
print('C=%d M=%d U=%d N=%d stride=%d tiles=%s: %.3f ms  %.1f TFLOP/s (algorithmic)  tile=%s' % (C, M, U, N, stride, sparse.tiles_enabled(), ms, fl / ms / 1e9, W._pg.classes[0].get('tile') is not None))
