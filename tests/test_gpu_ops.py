"""torch.library ops of the keynet_b200 namespace (keynet_b200/ops.py): same results as the host mirror's direct calls, usable
under torch.compile's tracing (fake implementations) and visible in the dispatcher."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_ops_are_registered_and_match_the_direct_path():
    from keynet_b200 import sparse, ops  # noqa: F401
    from oracle import keynet_oracle as ko
    rs = np.random.RandomState(0)
    f = rs.randn(5, 3, 3, 3).astype(np.float32); b = rs.randn(5).astype(np.float32)
    (fq, bq, _) = sparse._conv_weights_rounded((3, 8, 8), f, b, 1)
    (ip, ix, dt) = torch.ops.keynet_b200.toeplitz_conv2d_csr(torch.from_numpy(fq).cuda(), torch.from_numpy(bq).cuda(), 8, 8, 1)
    ref = ko.toeplitz_conv2d((3, 8, 8), f, b, 1)
    assert np.array_equal(ip.cpu().numpy(), ref.indptr) and np.array_equal(ix.cpu().numpy(), ref.indices)
    assert np.array_equal(dt.cpu().numpy().view(np.uint32), ref.data.view(np.uint32))
    # key compile: column permutation + gains
    K = ref.shape[1]
    perm = np.concatenate([rs.permutation(K - 1), [K - 1]]).astype(np.int32)
    cs = np.concatenate([rs.rand(K - 1) + 0.5, [1.0]]).astype(np.float32)
    rsc = np.concatenate([rs.rand(ref.shape[0] - 1) + 0.5, [1.0]]).astype(np.float32)
    (ip2, ix2, dt2) = torch.ops.keynet_b200.keycompile_monomial(ip, ix, dt, K, K, torch.from_numpy(perm).cuda(), torch.from_numpy(rsc).cuda(), torch.from_numpy(cs).cuda())
    o = ko.sort_indices(ko.key_compile(ko.monomial_key(np.arange(ref.shape[0]), rsc), ref, ko.monomial_key(perm, cs)))
    assert np.array_equal(ip2.cpu().numpy(), o.indptr) and np.array_equal(ix2.cpu().numpy(), o.indices)
    assert np.array_equal(dt2.cpu().numpy().view(np.uint32), o.data.view(np.uint32))
    # SpMM through the dispatcher == oracle
    X = rs.randn(K, 7).astype(np.float32)
    y = torch.ops.keynet_b200.spmm_csr(ip2, ix2, dt2, K, torch.from_numpy(X).cuda(), True)
    yr = ko.spmm(o, X, relu=True)
    assert np.allclose(y.cpu().numpy(), yr, rtol=1e-4, atol=1e-5 * np.abs(yr).max())
    # SparseMatrix.torchdot (the reference-facing operator method) is that op
    W = sparse.SparseMatrix((o.shape, o.indptr, o.indices, o.data))
    assert torch.equal(W.torchdot(torch.from_numpy(X).cuda(), relu=True), y)


def test_spmm_op_traces_with_fake_tensors():
    from keynet_b200 import ops  # noqa: F401
    from torch._subclasses.fake_tensor import FakeTensorMode
    with FakeTensorMode():
        ip = torch.empty(11, dtype=torch.int64, device='cuda'); ix = torch.empty(30, dtype=torch.int32, device='cuda'); dt = torch.empty(30, device='cuda')
        y = torch.ops.keynet_b200.spmm_csr(ip, ix, dt, 20, torch.empty(20, 6, device='cuda'), False)
        assert tuple(y.shape) == (10, 6)
    torch.library.opcheck(torch.ops.keynet_b200.spmm_csr.default, (torch.tensor([0, 1, 2], device='cuda'), torch.tensor([0, 1], dtype=torch.int32, device='cuda'),
                                                                   torch.ones(2, device='cuda'), 2, torch.ones(2, 3, device='cuda'), False), test_utils=('test_schema', 'test_faketensor'))
