"""Per-role cycle accounting of ONE CTA of the tiled tensor-core kernel (library built with -DKN_TILE_PROF)."""
import sys, ctypes
sys.path.insert(0, '.')
import numpy as np, torch
from keynet_b200 import sparse, _native
from keynet_b200.sparse import MonomialKey
(C, M, U, N) = [int(v) for v in sys.argv[1:5]]
rs = np.random.RandomState(0)
f = (rs.randn(M, C, 3, 3) * 0.05).astype(np.float32); b = rs.randn(M).astype(np.float32)
K = C * U * U + 1
W = sparse.keyed_toeplitz_conv2d((C, U, U), f, b, 1, None, MonomialKey(np.arange(K)), want_csr=False)
X = torch.randn(K, N, device='cuda'); X[-1] = 1
Y = torch.empty((W.shape[0], N), device='cuda')
L = _native.lib()
for _ in range(2):
    sparse.spmm(W, X, relu=True, out=Y)
torch.cuda.synchronize()
out = (ctypes.c_int64 * 64)()
L.kn_debug_tile_prof.argtypes = [ctypes.c_void_p, ctypes.c_int32]
L.kn_debug_tile_prof(out, 1)
sparse.spmm(W, X, relu=True, out=Y)
torch.cuda.synchronize()
L.kn_debug_tile_prof(out, 0)
t = list(out)
print('CTA total %d cycles: prologue %d, issuer0 done at %d, producer loop done at %d, accum ready at %d, epilogue done at %d' % (t[6] - t[0], t[1] - t[0], t[2] - t[0], t[3] - t[0], t[4] - t[0], t[5] - t[0]))
print('issuer 0: wait fullB %d, wait fullA %d, fence %d, MMA issue %d, commits %d, arrive (unused stage) %d' % (t[10], t[11], t[13], t[12], t[14], t[15]))
print('issuer 0: whole valid iterations %d, table loads at the top %d' % (t[16], t[17]))
print('splitter thread 256 (half the stages): wait raw_full %d, LDS+split %d, wait emptyA %d, STTM+wait+arrive %d' % (t[20], t[21], t[23], t[24]))
print('gather warp: wait raw_empty %d, issue %d' % (t[30], t[31]))
print('epilogue thread 256: tcgen05.ld + wait %d, stores %d' % (t[40], t[41]))
