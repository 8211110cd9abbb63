"""torch.library registration of the C-ABI entry points (north star: "a thin C-ABI PyTorch op"): dispatcher-visible ops
in the `keynet_b200` namespace, each a direct call into libkeynet_b200.so on the tensors' device pointers and the
current CUDA stream, with fake (meta) implementations so they trace / compile / capture like any other op.

  torch.ops.keynet_b200.spmm_csr(indptr, indices, data, n_cols, x, relu) -> y          SparseMatrix.torchdot (keynet/sparse.py:488-492)
  torch.ops.keynet_b200.spmm_csr_out(indptr, indices, data, n_cols, x, y, relu)        the same into a caller-owned buffer
  torch.ops.keynet_b200.keycompile_monomial(indptr, indices, data, n_cols_in, n_cols_out, col_map, row_scale, col_scale)
                                              -> (indptr, indices, data)                 A.dot(W).dot(Ainv) for one-entry-per-row keys (keynet/layer.py:35,59,70)
  torch.ops.keynet_b200.toeplitz_conv2d_csr(weight, bias, U, V, stride) -> (indptr, indices, data)    sparse_toeplitz_conv2d (keynet/sparse.py:163-203)

`SparseMatrix.torchdot` (the reference-facing operator method) goes through spmm_csr; the batched engine calls the library
directly (one ctypes call per layer, no dispatcher in the loop)."""
from typing import Optional, Tuple

import torch

from . import _native
from ._native import check, kn_conv2d_desc, ptr, stream_ptr


def _cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _native.NativeError('keynet_b200 ops need CUDA tensors (there is no CPU fallback)')


@torch.library.custom_op('keynet_b200::spmm_csr_out', mutates_args=('y',))
def spmm_csr_out(indptr: torch.Tensor, indices: torch.Tensor, data: torch.Tensor, n_cols: int, x: torch.Tensor, y: torch.Tensor, relu: bool) -> None:
    _cuda(indptr, indices, data, x, y)
    assert x.dtype == torch.float32 and y.dtype == torch.float32 and x.is_contiguous() and y.is_contiguous() and x.shape[0] == n_cols
    n_rows = indptr.numel() - 1
    assert tuple(y.shape) == (n_rows, x.shape[1])
    N = x.shape[1]
    check(_native.lib().kn_spmm_csr_f32(ptr(indptr), ptr(indices), ptr(data), n_rows, n_cols, ptr(x), N, ptr(y), N, N,
                                        _native.KN_SPMM_RELU if relu else 0, None, stream_ptr()))


@torch.library.custom_op('keynet_b200::spmm_csr', mutates_args=())
def spmm_csr(indptr: torch.Tensor, indices: torch.Tensor, data: torch.Tensor, n_cols: int, x: torch.Tensor, relu: bool) -> torch.Tensor:
    y = torch.empty((indptr.numel() - 1, x.shape[1]), dtype=torch.float32, device=x.device)
    spmm_csr_out(indptr, indices, data, n_cols, x, y, relu)
    return y


@spmm_csr.register_fake
def _(indptr, indices, data, n_cols, x, relu):
    return x.new_empty((indptr.numel() - 1, x.shape[1]), dtype=torch.float32)


@torch.library.custom_op('keynet_b200::keycompile_monomial', mutates_args=())
def keycompile_monomial(indptr: torch.Tensor, indices: torch.Tensor, data: torch.Tensor, n_cols_in: int, n_cols_out: int,
                        col_map: Optional[torch.Tensor], row_scale: Optional[torch.Tensor], col_scale: Optional[torch.Tensor]) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Row-gathered CSR in, canonical keyed CSR out: columns through col_map, values fl32(fl32(row_scale*v)*col_scale), exact
    zeros dropped, columns ascending."""
    _cuda(indptr, indices, data, col_map, row_scale, col_scale)
    L = _native.lib()
    n_rows = indptr.numel() - 1
    out_ip = torch.zeros(n_rows + 1, dtype=torch.int64, device=data.device)
    if n_rows > 0:
        check(L.kn_keycompile_count(ptr(indptr), ptr(indices), ptr(data), n_rows, ptr(row_scale), ptr(col_scale), None, None, n_cols_in, 0, ptr(out_ip[1:]), stream_ptr()))
        check(L.kn_exclusive_scan_i64(ptr(out_ip[1:]), ptr(out_ip), n_rows, stream_ptr()))
    nnz = int(out_ip[-1].item())
    out_ix = torch.empty(nnz, dtype=torch.int32, device=data.device)
    out_dt = torch.empty(nnz, dtype=torch.float32, device=data.device)
    if nnz > 0:
        check(L.kn_keycompile_fill(ptr(indptr), ptr(indices), ptr(data), n_rows, n_cols_out, ptr(col_map), ptr(row_scale), ptr(col_scale), None, None, n_cols_in, 0,
                                   ptr(out_ip), ptr(out_ix), ptr(out_dt), stream_ptr()))
    return (out_ip, out_ix, out_dt)


@keycompile_monomial.register_fake
def _(indptr, indices, data, n_cols_in, n_cols_out, col_map, row_scale, col_scale):
    nnz = torch.library.get_ctx().new_dynamic_size()
    return (indptr.new_empty(indptr.shape), indices.new_empty((nnz,)), data.new_empty((nnz,)))


@torch.library.custom_op('keynet_b200::toeplitz_conv2d_csr', mutates_args=())
def toeplitz_conv2d_csr(weight: torch.Tensor, bias: Optional[torch.Tensor], U: int, V: int, stride: int) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Un-keyed sparse Toeplitz matrix of a 'same'-padded cross-correlation, homogeneous form when bias is given (weights as
    given: the caller applies the reference's offset rounding, sparse._conv_weights_rounded)."""
    _cuda(weight, bias)
    L = _native.lib()
    (M, C, P, Q) = [int(v) for v in weight.shape]
    w = weight.detach().to(torch.float32).contiguous()
    b = None if bias is None else bias.detach().to(torch.float32).contiguous()
    desc = kn_conv2d_desc(C, int(U), int(V), M, P, Q, int(stride), 0, 1 if b is not None else 0)
    n_rows = M * (int(U) // int(stride)) * (int(V) // int(stride)) + (1 if b is not None else 0)
    ip = torch.zeros(n_rows + 1, dtype=torch.int64, device=w.device)
    check(L.kn_toeplitz_conv2d_count(desc, None, n_rows, ptr(ip[1:]), stream_ptr()))
    check(L.kn_exclusive_scan_i64(ptr(ip[1:]), ptr(ip), n_rows, stream_ptr()))
    nnz = int(ip[-1].item())
    ix = torch.empty(nnz, dtype=torch.int32, device=w.device)
    dt = torch.empty(nnz, dtype=torch.float32, device=w.device)
    check(L.kn_toeplitz_conv2d_fill(desc, ptr(w), ptr(b), None, n_rows, ptr(ip), ptr(ix), ptr(dt), stream_ptr()))
    return (ip, ix, dt)


@toeplitz_conv2d_csr.register_fake
def _(weight, bias, U, V, stride):
    nnz = torch.library.get_ctx().new_dynamic_size()
    (M, C, P, Q) = weight.shape
    n_rows = M * (U // stride) * (V // stride) + (1 if bias is not None else 0)
    return (weight.new_empty((n_rows + 1,), dtype=torch.int64), weight.new_empty((nnz,), dtype=torch.int32), weight.new_empty((nnz,), dtype=torch.float32))
