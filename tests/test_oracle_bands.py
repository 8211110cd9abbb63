"""Row bands of the oracle (oracle.toeplitz_conv2d_pixels / linear_matrix_rows / key_compile_rows) equal the same rows of
the full oracle matrices -- which are pinned bit-for-bit to the reference goldens (tests/test_oracle_golden.py).  The bands
are what bench.py's CPU legs time for VGG16, whose full matrices (120 GB) no host can hold."""
import numpy as np
import pytest

from oracle import keynet_oracle as ko


def _rows_of(W, rows):
    return [(W.indices[W.indptr[r]:W.indptr[r + 1]].tolist(), W.data[W.indptr[r]:W.indptr[r + 1]].view(np.uint32).tolist()) for r in rows]


@pytest.mark.parametrize('shape', [((3, 8, 8), 5, 3, 1), ((2, 12, 12), 4, 3, 2), ((1, 6, 6), 3, 5, 1), ((4, 4, 4), 6, 1, 1)])
def test_conv_band_equals_full_rows(shape):
    ((C, U, V), M, k, stride) = shape
    rs = np.random.RandomState(7)
    f = rs.randn(M, C, k, k).astype(np.float32)
    f[rs.rand(*f.shape) < 0.1] = 0                      # explicit zeros survive the offset trick
    b = rs.randn(M).astype(np.float32)
    full = ko.toeplitz_conv2d((C, U, V), f, b, stride)
    (Uo, Vo) = (U // stride, V // stride)
    pix = np.array([0, 1, Vo - 1, Vo, Uo * Vo // 2, Uo * Vo - 1])
    band = ko.toeplitz_conv2d_pixels((C, U, V), f, b, stride, pix)
    assert band.shape == full.shape
    rows = sorted(set((np.arange(M).reshape(-1, 1) * Uo * Vo + np.unique(pix).reshape(1, -1)).reshape(-1).tolist()) | {M * Uo * Vo})
    assert _rows_of(band, rows) == _rows_of(full, rows)
    others = np.setdiff1d(np.arange(full.shape[0]), rows)
    assert np.all(np.diff(band.indptr)[others] == 0)


def test_avgpool_and_linear_bands():
    full = ko.toeplitz_avgpool2d((3, 8, 8), 3, 2)
    band = ko.toeplitz_avgpool2d_pixels((3, 8, 8), 3, 2, [0, 5, 15])
    rows = sorted(set((np.arange(3).reshape(-1, 1) * 16 + np.array([0, 5, 15]).reshape(1, -1)).reshape(-1).tolist()) | {48})
    assert _rows_of(band, rows) == _rows_of(full, rows)
    rs = np.random.RandomState(0)
    (w, b) = (rs.randn(9, 20).astype(np.float32), rs.randn(9).astype(np.float32))
    w[2, 3] = 0
    full = ko.linear_matrix(w, b)
    band = ko.linear_matrix_rows(w, b, [1, 2, 7])
    assert _rows_of(band, [1, 2, 7, 9]) == _rows_of(full, [1, 2, 7, 9])


def test_key_compile_rows_equals_rows_of_full_compile():
    rs = np.random.RandomState(3)
    (C, U, V, M) = (3, 8, 8, 4)
    f = rs.randn(M, C, 3, 3).astype(np.float32); b = rs.randn(M).astype(np.float32)
    (R, K) = (M * U * V + 1, C * U * V + 1)
    po = np.concatenate([rs.permutation(R - 1), [R - 1]]); a = np.concatenate([rs.rand(R - 1) + 0.5, [1.0]]).astype(np.float32)
    pi = np.concatenate([rs.permutation(K - 1), [K - 1]]); ai = np.concatenate([rs.rand(K - 1) + 0.5, [1.0]]).astype(np.float32)
    (A, Ainv) = (ko.monomial_key(po, a), ko.monomial_key(pi, ai))
    W = ko.toeplitz_conv2d((C, U, V), f, b, 1)
    full = ko.key_compile(A, W, Ainv)
    pix = np.array([3, 17, 40])
    src = np.concatenate([(np.arange(M).reshape(-1, 1) * U * V + pix.reshape(1, -1)).reshape(-1), [R - 1]])
    inv = np.empty(R, dtype=np.int64); inv[po] = np.arange(R)
    rows = np.sort(inv[src])                                     # rows of W_hat whose Toeplitz row lies in the band
    band = ko.key_compile_rows(A, rows, ko.toeplitz_conv2d_pixels((C, U, V), f, b, 1, pix), Ainv)
    assert band.shape == (len(rows), K)
    got = _rows_of(band, range(len(rows)))
    assert got == _rows_of(full, rows)                           # including scipy's unsorted stored order
    # A = None (last layer)
    band = ko.key_compile_rows(None, src, ko.toeplitz_conv2d_pixels((C, U, V), f, b, 1, pix), Ainv)
    assert _rows_of(band, range(len(src))) == _rows_of(ko.key_compile(None, W, Ainv), src)
