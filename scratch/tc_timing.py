import sys, ctypes, numpy as np, torch
sys.path.insert(0, '.')
from keynet_b200 import sparse, _native
from keynet_b200._native import check
L = _native.lib()
rs = np.random.RandomState(0)
for (M, C, U) in [(96, 96, 32), (192, 192, 16)]:
    W = sparse.keyed_toeplitz_conv2d((C, U, U), rs.randn(M, C, 3, 3).astype(np.float32) * 0.05, rs.randn(M).astype(np.float32), 1, None, sparse.sparse_identity_matrix(C * U * U + 1))
    print(W._pg.summary())
    X = torch.randn(W.shape[1], 4096, device='cuda')
    Y = torch.empty(W.shape[0], 4096, device='cuda')
    for it in range(3):
        sparse.spmm(W, X, relu=True, out=Y)
    torch.cuda.synchronize()
    for cta in (100, 3000):
        check(L.kn_debug_tc_timing(cta, None))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); sparse.spmm(W, X, relu=True, out=Y); e1.record(); torch.cuda.synchronize()
        out = (ctypes.c_int64 * 8)()
        check(L.kn_debug_tc_timing(-1, out))
        t = np.array(list(out)[:7], dtype=np.int64)
        names = ['prologue', 'first stage full', 'main loop (issue)', 'drain to accum ready', 'epilogue', 'teardown sync']
        print('M=%d cta=%d launch %.3f ms: ' % (M, cta, e0.elapsed_time(e1)) + ', '.join('%s=%d' % (n, d) for (n, d) in zip(names, np.diff(t))) + ' | total=%d cycles' % (t[6] - t[0]))
