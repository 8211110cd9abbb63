"""CPU ORACLE for the keyed-layer forward path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

numpy + ctypes glue around oracle/keynet_oracle.c.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / `--impl reference` legs may import this module; keynet_b200/ never does.

The reference (visym/keynet) is pure Python over scipy.sparse; this oracle restates its
algorithm for the path (Toeplitz build -> key compile A.W.Ainv -> CSR x dense forward) without
importing scipy or the reference, so it can run on the GPU box where /root/reference does not
exist.  Parity is PINNED: tests/test_oracle_golden.py checks every function here bit-for-bit
against golden vectors produced by the unmodified reference (tests/golden/make_golden.py) and
against scipy's own kernels.

Reference call sites restated (paths relative to /root/reference):
  toeplitz_conv2d      keynet/sparse.py:163-203   (offset trick :184-187, bias column :190-198)
  toeplitz_avgpool2d   keynet/sparse.py:206-212
  linear_matrix        keynet/torch.py:80-89 + keynet/layer.py:69
  matmat / key_compile keynet/layer.py:35,46,59,70   (scipy csr_matmat x2, left product first)
  spmm                 keynet/sparse.py:488-492     (scipy csr_matvecs)
  affine_to_linear / linear_to_affine   keynet/torch.py:65-77
"""
import ctypes
import os
import subprocess
from collections import namedtuple

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(HERE, 'keynet_oracle.c')
_LIB = os.path.join(HERE, '_build', 'libkeynet_oracle.so')

CSR = namedtuple('CSR', ['shape', 'indptr', 'indices', 'data'])   # indptr int64, indices int32, data float32


def build(force=False):
    """Compile the C restatement with gcc (no FMA contraction; OpenMP for the SpMM row loop)."""
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(_SRC):
        os.makedirs(os.path.dirname(_LIB), exist_ok=True)
        subprocess.check_call(['gcc', '-O2', '-fPIC', '-shared', '-fopenmp', '-ffp-contract=off', '-fvisibility=hidden',
                               '-o', _LIB, _SRC])
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(build())
        i64, i32p, i64p, f32p, f64p = ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p
        L.ko_toeplitz_conv2d_coo.restype = ctypes.c_int64
        L.ko_toeplitz_conv2d_coo.argtypes = [ctypes.c_int] * 7 + [f32p, i32p, i32p, f32p]
        L.ko_toeplitz_conv2d_coo_pixels.restype = ctypes.c_int64
        L.ko_toeplitz_conv2d_coo_pixels.argtypes = [ctypes.c_int] * 7 + [f32p, i64, i64p, i32p, i32p, f32p]
        L.ko_toeplitz_conv2d_emitted_min.restype = ctypes.c_float
        L.ko_toeplitz_conv2d_emitted_min.argtypes = [ctypes.c_int] * 7 + [f32p]
        L.ko_coo_tocsr.restype = ctypes.c_int64
        L.ko_coo_tocsr.argtypes = [i64, i64, i32p, i32p, f32p, i64p, i32p, f32p]
        L.ko_csr_matmat_maxnnz.restype = ctypes.c_int64
        L.ko_csr_matmat_maxnnz.argtypes = [i64, i64, i64p, i32p, i64p, i32p]
        L.ko_csr_matmat.restype = ctypes.c_int64
        L.ko_csr_matmat.argtypes = [i64, i64, i64p, i32p, f32p, i64p, i32p, f32p, i64p, i32p, f32p]
        L.ko_csr_matvecs.restype = None
        L.ko_csr_matvecs.argtypes = [i64, i64, i64p, i32p, f32p, f32p, f32p, ctypes.c_int, ctypes.c_int]
        L.ko_csr_matvecs_f64.restype = None
        L.ko_csr_matvecs_f64.argtypes = [i64, i64, i64p, i32p, f64p, f64p, f64p]
        L.ko_csr_sort_indices.restype = None
        L.ko_csr_sort_indices.argtypes = [i64, i64p, i32p, f32p]
        L.ko_max_threads.restype = ctypes.c_int
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def max_threads():
    return int(lib().ko_max_threads())


# ---------------------------------------------------------------------------------------------
# containers
def csr(shape, indptr, indices, data):
    return CSR((int(shape[0]), int(shape[1])), np.ascontiguousarray(indptr, dtype=np.int64),
               np.ascontiguousarray(indices, dtype=np.int32), np.ascontiguousarray(data, dtype=np.float32))


def csr_from_coo(shape, row, col, val):
    """scipy coo_matrix(...).tocsr(): canonical (sorted, duplicates summed), explicit zeros kept."""
    row = np.ascontiguousarray(row, dtype=np.int32); col = np.ascontiguousarray(col, dtype=np.int32)
    val = np.ascontiguousarray(val, dtype=np.float32)
    n = len(val)
    indptr = np.zeros(shape[0] + 1, dtype=np.int64)
    indices = np.zeros(max(n, 1), dtype=np.int32); data = np.zeros(max(n, 1), dtype=np.float32)
    nnz = lib().ko_coo_tocsr(shape[0], n, _p(row), _p(col), _p(val), _p(indptr), _p(indices), _p(data))
    return csr(shape, indptr, indices[:nnz], data[:nnz])


def csr_from_dense(D):
    """scipy coo_matrix(dense): keeps only non-zero entries, row-major order."""
    D = np.asarray(D, dtype=np.float32)
    (r, c) = np.nonzero(D)
    return csr_from_coo(D.shape, r, c, D[r, c])


def todense(A):
    D = np.zeros(A.shape, dtype=np.float64)
    rows = np.repeat(np.arange(A.shape[0]), np.diff(A.indptr))
    np.add.at(D, (rows, A.indices), A.data.astype(np.float64))
    return D


def sort_indices(A):
    """Canonical form used for parity: column indices sorted within each row."""
    indices = A.indices.copy(); data = A.data.copy()
    lib().ko_csr_sort_indices(A.shape[0], _p(A.indptr), _p(indices), _p(data))
    return CSR(A.shape, A.indptr, indices, data)


def transpose(A):
    rows = np.repeat(np.arange(A.shape[0], dtype=np.int32), np.diff(A.indptr))
    return csr_from_coo((A.shape[1], A.shape[0]), A.indices, rows, A.data)


# ---------------------------------------------------------------------------------------------
# key matrices in homogeneous (N+1)x(N+1) form, as CSR (any sparse key works with matmat)
def monomial_key(perm, scale=None):
    """A[r, perm[r]] = scale[r]   (keynet/sparse.py:280-285 permutation, :318-321 diagonal)."""
    perm = np.asarray(perm, dtype=np.int32)
    n = len(perm)
    scale = np.ones(n, dtype=np.float32) if scale is None else np.asarray(scale, dtype=np.float32)
    return csr((n, n), np.arange(n + 1, dtype=np.int64), perm, scale)


def identity_key(n):
    return monomial_key(np.arange(n))


# ---------------------------------------------------------------------------------------------
def toeplitz_conv2d(inshape, f, bias=None, stride=1):
    """Sparse Toeplitz matrix of a 'same'-padded cross-correlation, homogeneous form when bias is given.

    Restates keynet/sparse.py:163-203: values pass through fl32(fl32(w+off)-off) with
    off = fl32(|min(emitted values)|+1) so that zero coefficients stay stored; the bias column is
    built the same way with its own offset; last row is e_last; result is canonical CSR.
    """
    (C, U, V) = [int(s) for s in inshape]
    f = np.ascontiguousarray(f, dtype=np.float32)
    (M, C2, P, Q) = f.shape
    assert C2 == C and P == Q and P % 2 == 1
    (Uo, Vo) = (U // stride, V // stride)
    cap = Uo * Vo * C * M * P * Q
    rows = np.zeros(max(cap, 1), dtype=np.int32); cols = np.zeros(max(cap, 1), dtype=np.int32); vals = np.zeros(max(cap, 1), dtype=np.float32)
    n = lib().ko_toeplitz_conv2d_coo(C, U, V, M, P, Q, int(stride), _p(f), _p(rows), _p(cols), _p(vals))
    (rows, cols, vals) = (rows[:n], cols[:n], vals[:n].copy())
    offset = np.float32(np.abs(np.min(vals)) + np.float32(1.0))
    vals += offset
    vals -= offset
    (R, K) = (M * Uo * Vo, C * U * V)
    if bias is None:
        return csr_from_coo((R, K), rows, cols, vals)
    bias = np.asarray(bias, dtype=np.float32)
    UV = Uo * Vo
    boff = np.float32(np.abs(np.min(bias)) + np.float32(1.0))
    bvals = np.repeat(bias, UV).astype(np.float32)
    bvals = (bvals + boff).astype(np.float32)
    bvals -= boff
    brows = np.arange(M * UV, dtype=np.int32)
    bcols = np.full(M * UV, K, dtype=np.int32)
    rows = np.concatenate([rows, brows, np.array([R], dtype=np.int32)])
    cols = np.concatenate([cols, bcols, np.array([K], dtype=np.int32)])
    vals = np.concatenate([vals, bvals, np.array([1.0], dtype=np.float32)])
    return csr_from_coo((R + 1, K + 1), rows, cols, vals)


def toeplitz_conv2d_pixels(inshape, f, bias, stride, pixels):
    """ROW BAND of toeplitz_conv2d(inshape, f, bias, stride): only the rows of the output pixels `pixels` (ku*Vo + kv) are
    populated -- all M channel rows of each -- plus the last row e_last; every other row is empty.  Same shape, same
    global row / column numbers and the same offset rounding as the full matrix (the offset uses the minimum over the
    FULL emission, keynet/sparse.py:184), so  band.rows == full.rows  on the populated rows (tests/test_oracle_bands.py).
    For layers the host cannot hold (VGG16 conv1_2: 1.8 G entries)."""
    (C, U, V) = [int(s) for s in inshape]
    f = np.ascontiguousarray(f, dtype=np.float32)
    (M, C2, P, Q) = f.shape
    assert C2 == C and P == Q and P % 2 == 1 and bias is not None
    (Uo, Vo) = (U // stride, V // stride)
    pixels = np.unique(np.asarray(pixels, dtype=np.int64))
    assert len(pixels) == 0 or (pixels[0] >= 0 and pixels[-1] < Uo * Vo)
    cap = len(pixels) * C * M * P * Q
    rows = np.zeros(max(cap, 1), dtype=np.int32); cols = np.zeros(max(cap, 1), dtype=np.int32); vals = np.zeros(max(cap, 1), dtype=np.float32)
    n = lib().ko_toeplitz_conv2d_coo_pixels(C, U, V, M, P, Q, int(stride), _p(f), len(pixels), _p(pixels), _p(rows), _p(cols), _p(vals))
    (rows, cols, vals) = (rows[:n], cols[:n], vals[:n].copy())
    mn = np.float32(lib().ko_toeplitz_conv2d_emitted_min(C, U, V, M, P, Q, int(stride), _p(f)))
    offset = np.float32(np.abs(mn) + np.float32(1.0))
    vals += offset
    vals -= offset
    (R, K) = (M * Uo * Vo, C * U * V)
    bias = np.asarray(bias, dtype=np.float32)
    boff = np.float32(np.abs(np.min(bias)) + np.float32(1.0))
    brows = (np.arange(M, dtype=np.int64).reshape(-1, 1) * (Uo * Vo) + pixels.reshape(1, -1)).reshape(-1).astype(np.int32)
    bvals = np.repeat(bias, len(pixels)).astype(np.float32)
    bvals = (bvals + boff).astype(np.float32)
    bvals -= boff
    rows = np.concatenate([rows, brows, np.array([R], dtype=np.int32)])
    cols = np.concatenate([cols, np.full(len(brows), K, dtype=np.int32), np.array([K], dtype=np.int32)])
    vals = np.concatenate([vals, bvals, np.array([1.0], dtype=np.float32)])
    return csr_from_coo((R + 1, K + 1), rows, cols, vals)


def toeplitz_avgpool2d_pixels(inshape, kernel_size, stride, pixels):
    """Row band of toeplitz_avgpool2d (dense-channel filter: C*C channel pairs per pixel, mostly explicit zeros)."""
    C = int(inshape[0])
    F = np.zeros((C, C, kernel_size, kernel_size), dtype=np.float32)
    for k in range(C):
        F[k, k, :, :] = 1.0 / (kernel_size * kernel_size)
    return toeplitz_conv2d_pixels(inshape, F, np.zeros(C, dtype=np.float32), stride, pixels)


def linear_matrix_rows(weight, bias, rows):
    """Row band of linear_matrix: rows `rows` (< out) of [[W, b],[0, 1]] plus the last row, every other row empty."""
    weight = np.asarray(weight, dtype=np.float32)
    (out, inn) = weight.shape
    rows = np.unique(np.asarray(rows, dtype=np.int64))
    D = np.zeros((len(rows) + 1, inn + 1), dtype=np.float32)
    D[:-1, :inn] = weight[rows]
    D[:-1, inn] = 0 if bias is None else np.asarray(bias, dtype=np.float32)[rows]
    D[-1, inn] = 1
    (r, c) = np.nonzero(D)
    grow = np.concatenate([rows, [out]])[r]
    return csr_from_coo((out + 1, inn + 1), grow, c, D[r, c])


def key_compile_rows(A, rows, W, Ainv):
    """Rows `rows` of W_hat = A.dot(W).dot(Ainv): the same two csr_matmat products with A restricted to those rows
    (a row of a product depends only on that row of the left factor).  W may be a row band: rows of W that
    A[rows, :] does not touch are never read."""
    rows = np.asarray(rows, dtype=np.int64)
    if A is None:
        counts = np.diff(W.indptr)[rows]
        indptr = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
        take = np.concatenate([np.arange(W.indptr[r], W.indptr[r + 1]) for r in rows] + [np.zeros(0, dtype=np.int64)]).astype(np.int64)
        T = csr((len(rows), W.shape[1]), indptr, W.indices[take], W.data[take])
    else:
        counts = np.diff(A.indptr)[rows]
        indptr = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
        take = np.concatenate([np.arange(A.indptr[r], A.indptr[r + 1]) for r in rows] + [np.zeros(0, dtype=np.int64)]).astype(np.int64)
        T = matmat(csr((len(rows), A.shape[1]), indptr, A.indices[take], A.data[take]), W)
    return matmat(T, Ainv)


def toeplitz_avgpool2d(inshape, kernel_size, stride):
    """keynet/sparse.py:206-212: dense-channel (C,C,k,k) filter with 1/k^2 on the channel diagonal, zero bias."""
    C = int(inshape[0])
    F = np.zeros((C, C, kernel_size, kernel_size), dtype=np.float32)
    for k in range(C):
        F[k, k, :, :] = 1.0 / (kernel_size * kernel_size)
    return toeplitz_conv2d(inshape, F, bias=np.zeros(C, dtype=np.float32), stride=stride)


def linear_matrix(weight, bias):
    """(out+1)x(in+1) homogeneous matrix [[W, b],[0, 1]] with exact zeros dropped
    (keynet/torch.py:80-89 builds its transpose densely; keynet/layer.py:69 takes coo_matrix(...).transpose())."""
    weight = np.asarray(weight, dtype=np.float32)
    (out, inn) = weight.shape
    D = np.zeros((out + 1, inn + 1), dtype=np.float32)
    D[:out, :inn] = weight
    D[:out, inn] = 0 if bias is None else np.asarray(bias, dtype=np.float32)
    D[out, inn] = 1
    return csr_from_dense(D)


# ---------------------------------------------------------------------------------------------
def matmat(A, B):
    """C = A.dot(B) exactly as scipy csr_matmat: zero sums dropped, columns in reverse first-touch order."""
    assert A.shape[1] == B.shape[0], (A.shape, B.shape)
    L = lib()
    cap = L.ko_csr_matmat_maxnnz(A.shape[0], B.shape[1], _p(A.indptr), _p(A.indices), _p(B.indptr), _p(B.indices))
    Cp = np.zeros(A.shape[0] + 1, dtype=np.int64)
    Cj = np.zeros(max(cap, 1), dtype=np.int32); Cx = np.zeros(max(cap, 1), dtype=np.float32)
    nnz = L.ko_csr_matmat(A.shape[0], B.shape[1], _p(A.indptr), _p(A.indices), _p(A.data),
                          _p(B.indptr), _p(B.indices), _p(B.data), _p(Cp), _p(Cj), _p(Cx))
    return CSR((A.shape[0], B.shape[1]), Cp, Cj[:nnz].copy(), Cx[:nnz].copy())


def key_compile(A, W, Ainv):
    """W_hat = A.dot(W).dot(Ainv) (keynet/layer.py:35,59,70); A may be None (last layer: W.dot(Ainv))."""
    T = W if A is None else matmat(A, W)
    return matmat(T, Ainv)


def spmm(W, X, relu=False, threads=1):
    """Y[R,N] = W[R,C] . X[C,N], fp32, sequential per-row accumulation in stored order (scipy csr_matvecs)."""
    X = np.ascontiguousarray(X, dtype=np.float32)
    assert X.ndim == 2 and X.shape[0] == W.shape[1], (W.shape, X.shape)
    Y = np.zeros((W.shape[0], X.shape[1]), dtype=np.float32)
    lib().ko_csr_matvecs(W.shape[0], X.shape[1], _p(W.indptr), _p(W.indices), _p(W.data), _p(X), _p(Y), int(bool(relu)), int(threads))
    return Y


def spmm_f64(shape, indptr, indices, data64, X):
    """fp64-matrix SpMM (challenge pickle conv/pool layers)."""
    X = np.ascontiguousarray(X, dtype=np.float64)
    Y = np.zeros((shape[0], X.shape[1]), dtype=np.float64)
    indptr = np.ascontiguousarray(indptr, dtype=np.int64); indices = np.ascontiguousarray(indices, dtype=np.int32)
    data64 = np.ascontiguousarray(data64, dtype=np.float64)
    lib().ko_csr_matvecs_f64(shape[0], X.shape[1], _p(indptr), _p(indices), _p(data64), _p(X), _p(Y))
    return Y


def affine_to_linear(x):
    """N x C x H x W -> N x (CHW+1), last column one (keynet/torch.py:65-68)."""
    x = np.asarray(x, dtype=np.float32)
    x = x.reshape(x.shape[0], -1) if x.ndim == 4 else x.reshape(1, -1)
    return np.concatenate([x, np.ones((x.shape[0], 1), dtype=np.float32)], axis=1)


def linear_to_affine(x, outshape=None):
    """keynet/torch.py:71-77: last column must be ~1 (atol 1e-3) else ValueError; drop it."""
    x = np.asarray(x)
    assert x.ndim == 2
    if not np.allclose(x[:, -1], 1, atol=1e-3):
        raise ValueError('invalid affine vector')
    y = x[:, :-1]
    return y.reshape(outshape) if outshape is not None else y


def keyed_forward(layers, x_linear, threads=1):
    """layers: list of (CSR W_hat, relu_after: bool); x_linear: [N, D+1].  Returns [N, Dout+1]
    (keynet/layer.py:92 y = W.torchdot(x.t()).t(), then the unkeyed nn.ReLU of system.py:92)."""
    X = np.ascontiguousarray(np.asarray(x_linear, dtype=np.float32).T)
    for (W, relu) in layers:
        X = spmm(W, X, relu=relu, threads=threads)
    return np.ascontiguousarray(X.T)
