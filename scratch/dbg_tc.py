import sys, numpy as np, torch
sys.path.insert(0, '.')
from keynet_b200 import sparse, _native
from keynet_b200._native import ptr, stream_ptr, check
rs = np.random.RandomState(0)
(M, C, U, V, N) = (32, 3, 6, 8, 128)
W = sparse.keyed_toeplitz_conv2d((C, U, V), np.ones((M, C, 3, 3), dtype=np.float32), np.ones(M, dtype=np.float32), 1, None, sparse.sparse_identity_matrix(C*U*V+1))
W._pg = sparse.PatternGroups.build(W, min_group=4)
print(W._pg.summary())
X = torch.ones(W.shape[1], N, device='cuda')
for relu in (False,):
    y = torch.full((W.shape[0], N), float('nan'), device='cuda')
    sparse.spmm(W, X, relu=relu, out=y)
    torch.cuda.synchronize()
    yc = y.cpu().numpy()
    print('nan count', np.isnan(yc).sum(), 'of', yc.size, 'min', np.nanmin(yc), 'max', np.nanmax(yc))
    print(yc[:4, :8]); print(yc[-3:, :4])
    sparse.tensor_cores_enabled(False)
    y2 = sparse.spmm(W, X, relu=relu).cpu().numpy(); sparse.tensor_cores_enabled(True)
    print('simt', y2[:4, :8])
