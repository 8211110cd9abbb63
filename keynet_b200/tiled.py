"""Unique-tile views of keyed layer matrices: TiledMatrix / DiagonalTiledMatrix / Conv2dTiledMatrix with the
observable structure of the reference (keynet/sparse.py:517-835): `_blocks` = [(row0, col0, tile_id)] in row-major
order, unique tiles numbered by first appearance, `nnz()` = stored elements of the unique tiles (the paper's
parameter count), `tosparse()/tocsr()/tocoo()`, `torchdot()`.

B200-first differences (not a port):
  * the tile tables are computed on the GPU with sort / unique / scatter-reduce over the COO entries (the reference
    walks every non-zero in a Python loop and hashes `str(sorted(entries))`, sparse.py:547-568);
  * `torchdot` does NOT re-expand tiles to CSR on every call (sparse.py:610, 816): the matrix stays resident on the
    device and runs on the pattern-grouped kernels (csrc/pgroup*.cu), whose unique value blocks are the execution
    form of the same idea -- so tiled and untiled layers execute identically and only `nnz()` / structure differ.
"""
import numpy as np
import torch

from .sparse import SparseMatrix, MonomialKey, PatternGroups


def _mix(x):
    """64-bit mixer on int64 tensors (wrap-around arithmetic); collisions are as (un)likely as with the reference's hash()."""
    x = x * -7046029254386353131          # 0x9E3779B97F4A7C15
    x = x ^ (x >> 29)
    x = x * -4658895280553007687          # 0xBF58476D1CE4E5B9
    x = x ^ (x >> 32)
    return x


def _coo(W):
    """(rows, cols, vals) of a device SparseMatrix in CSR (row-major) order = the order scipy's tocoo() yields."""
    n = W.shape[0]
    counts = W._indptr[1:] - W._indptr[:-1]
    rows = torch.repeat_interleave(torch.arange(n, device=W._data.device, dtype=torch.int64), counts)
    off = int(W._indptr[0].item())
    return (rows, W._indices[off:off + rows.numel()].to(torch.int64), W._data[off:off + rows.numel()])


def _tile_tables(rows, cols, vals, shape, tileshape):
    """Blocks (row-major) and unique-tile ids (numbered by first appearance in COO order) of a sparse matrix.
    Returns dict(block_i0, block_j0, block_tile, n_tiles, tile_nnz, block_of_entry (inverse), rep_block)."""
    (H, W) = shape
    (h, w) = tileshape
    dev = vals.device
    n = rows.numel()
    nbw = -(-W // w)
    b = (rows // h) * nbw + (cols // w)
    (ub, inv) = torch.unique(b, return_inverse=True)                 # ascending block id == row-major (i0, j0) order
    nb = ub.numel()
    vbits = vals.contiguous().view(torch.int32).to(torch.int64)
    eh = _mix((rows % h) * 1000003 + (cols % w) * 7919 + _mix(vbits))
    bh = torch.zeros(nb, dtype=torch.int64, device=dev).scatter_add_(0, inv, eh)      # order-independent: a SET of entries
    cnt = torch.bincount(inv, minlength=nb)
    (bi, bj) = (ub // nbw, ub % nbw)
    sh = torch.clamp(H - bi * h, max=h) * 65537 + torch.clamp(W - bj * w, max=w)       # ragged edge tiles differ by shape
    bh = _mix(bh + cnt * 2654435761 + sh * 40503)
    first = torch.full((nb,), n, dtype=torch.int64, device=dev).scatter_reduce_(0, inv, torch.arange(n, device=dev), reduce='amin')
    order = torch.argsort(first)                                      # blocks in order of first appearance
    (uh, inv2) = torch.unique(bh[order], return_inverse=True)
    nt = uh.numel()
    first_rank = torch.full((nt,), nb, dtype=torch.int64, device=dev).scatter_reduce_(0, inv2, torch.arange(nb, device=dev), reduce='amin')
    tile_rank = torch.empty(nt, dtype=torch.int64, device=dev)
    tile_rank[torch.argsort(first_rank)] = torch.arange(nt, device=dev)
    block_tile = torch.empty(nb, dtype=torch.int64, device=dev)
    block_tile[order] = tile_rank[inv2]
    rep_block = torch.empty(nt, dtype=torch.int64, device=dev)        # representative block of every tile
    rep_block[tile_rank] = order[first_rank]
    return dict(block_i0=bi * h, block_j0=bj * w, block_tile=block_tile, n_tiles=int(nt), tile_nnz=cnt[rep_block],
                block_of_entry=inv, rep_block=rep_block, block_id=ub, nbw=nbw)


class TiledMatrix(SparseMatrix):
    def __init__(self, T, tileshape):
        """T: SparseMatrix (device CSR) or anything SparseMatrix() accepts; tileshape=(h, w) > 0."""
        assert isinstance(tileshape, tuple) and len(tileshape) == 2 and tileshape[0] > 0 and tileshape[1] > 0, "tileshape must be tuple (tileheight, tilewidth) > 0"
        T = T if isinstance(T, SparseMatrix) else SparseMatrix(T)
        SparseMatrix.__init__(self, T)
        self._pg = T._pg
        self._tileshape = (int(tileshape[0]), int(tileshape[1]))
        (r, c, v) = _coo(self)
        self._tab = _tile_tables(r, c, v, self.shape, self._tileshape)
        self._blocks = None

    def __repr__(self):
        return str('<keynet_b200.TiledMatrix: H=%d, W=%d, tileshape=%s, tiles=%d>' % (*self.shape, str(self.tileshape()), self._tab['n_tiles']))

    def tileshape(self):
        return self._tileshape

    def blocks(self):
        if self._blocks is None:
            t = self._tab
            self._blocks = [(int(i), int(j), int(k)) for (i, j, k) in zip(t['block_i0'].tolist(), t['block_j0'].tolist(), t['block_tile'].tolist())]
        return self._blocks

    def __iter__(self):
        for b in self.blocks():
            yield b

    def tiles(self):
        """Unique tiles as scipy COO matrices (host; small matrices / tests only)."""
        import scipy.sparse
        (r, c, v) = [t.cpu().numpy() for t in _coo(self)]
        t = self._tab
        (h, w) = self._tileshape
        (inv, rep) = (t['block_of_entry'].cpu().numpy(), t['rep_block'].cpu().numpy())
        (bi0, bj0) = (t['block_i0'].cpu().numpy(), t['block_j0'].cpu().numpy())
        out = []
        for k in range(t['n_tiles']):
            m = inv == rep[k]
            (i0, j0) = (bi0[rep[k]], bj0[rep[k]])
            shp = (min(h, self.shape[0] - i0), min(w, self.shape[1] - j0))
            A = scipy.sparse.coo_matrix((np.ones(m.sum(), dtype=np.float32), (r[m] - i0, c[m] - j0)), shape=shp)
            A.data = v[m].astype(np.float32)      # keeps explicit zeros
            out.append(A)
        return out

    def nnz(self):
        """Stored elements of the unique tiles (reference: sum(t.nnz for t in self._tiles), sparse.py:649)."""
        return int(self._tab['tile_nnz'].sum().item())

    def expanded_nnz(self):
        return SparseMatrix.nnz(self)

    def tosparse(self, format='coo'):
        A = SparseMatrix.tocoo(self)
        if format not in ('coo', 'csr', 'csc'):
            raise ValueError('Invalid format "%s" - must be ["coo", "csr", "csc"]' % format)
        return A.asformat(format)

    def tocsr(self):
        return self.tosparse(format='csr')

    def tocoo(self):
        return self.tosparse(format='coo')

    def torchdot(self, x, relu=False):
        """(C*H*W+1) x N -> R x N; same shape assertion as the reference (sparse.py:605)."""
        assert self.shape[1] == x.shape[0], "Non-conformal shape for W=%s, x=%s" % (str(self.shape), str(tuple(x.shape)))
        return SparseMatrix.torchdot(self, x, relu=relu)

    def dot(self, x):
        assert isinstance(x, np.ndarray)
        return self.torchdot(torch.as_tensor(x)).numpy()


class DiagonalTiledMatrix(TiledMatrix):
    def __init__(self, B, shape):
        """Key block B repeated down the main diagonal of a `shape` matrix (reference sparse.py:657-687); for monomial
        blocks this is keynet_b200.sparse.sparse_block_diagonal_repeat, wrapped here for API parity."""
        from .sparse import sparse_block_diagonal_repeat
        assert isinstance(B, MonomialKey), 'DiagonalTiledMatrix expects a key block'
        assert isinstance(shape, tuple) and len(shape) == 2 and shape[0] == shape[1], "invalid shape"
        self._key = sparse_block_diagonal_repeat(B, shape)
        TiledMatrix.__init__(self, SparseMatrix(self._key), B.shape)

    def tokey(self):
        return self._key


class Conv2dTiledMatrix(TiledMatrix):
    def __init__(self, T, inshape, outshape, tileshape, bias, sanitycheck=True):
        """Unique spatial tile x dense (Cout, Cin) channel block (reference sparse.py:690-779).  The spatial tile
        structure comes from the channel-(0,0) block of T; every (unique tile, position) carries a dense Cout x Cin
        matrix, so nnz() = entries * Cout * Cin (+ one element per bias-column tile entry)."""
        (Cin, Hin, Win) = [int(s) for s in inshape]
        (Cout, Hout, Wout) = [int(s) for s in outshape]
        T = T if isinstance(T, SparseMatrix) else SparseMatrix(T)
        SparseMatrix.__init__(self, T)
        self._pg = T._pg
        (self._inshape, self._outshape) = (inshape, outshape)
        self._tileshape = (int(tileshape[0]), int(tileshape[1]))
        (h, w) = self._tileshape
        assert h <= self.shape[0] and w <= self.shape[1]
        (R, K) = (Cout * Hout * Wout, Cin * Hin * Win)
        if bias:
            assert self.shape == (R + 1, K + 1) and R % h == 0 and K % w == 0
        else:
            assert self.shape == (R, K) and R % h == 0 and K % w == 0
        (r, c, v) = _coo(self)
        (HoWo, HiWi) = (Hout * Wout, Hin * Win)
        # spatial tile structure of the channel-(0,0) block
        m00 = (r < HoWo) & (c < HiWi)
        t00 = _tile_tables(r[m00], c[m00], v[m00], (HoWo, HiWi), self._tileshape)
        # tile entries: distinct (it, jt, tile) over ALL channel pairs whose spatial block exists in T_00
        main = (r < R) & (c < K)
        (si, sj) = (r[main] % HoWo, c[main] % HiWi)
        sb = (si // h) * t00['nbw'] + (sj // w)
        pos = torch.searchsorted(t00['block_id'], sb)
        pos = pos.clamp(max=t00['block_id'].numel() - 1)
        present = t00['block_id'][pos] == sb
        kt = t00['block_tile'][pos][present]
        key = (kt * h + (si[present] % h)) * w + (sj[present] % w)
        key = torch.cat([key, torch.arange(t00['n_tiles'], device=key.device) * h * w])   # every tile owns a (0,0) entry
        n_entries = int(torch.unique(key).numel())
        self._n_spatial_entries = n_entries
        self._n_tile_entries = n_entries
        self._nnz_tiled = n_entries * Cout * Cin
        blocks = [(int(i), int(j), int(k)) for (i, j, k) in zip(t00['block_i0'].tolist(), t00['block_j0'].tolist(), t00['block_tile'].tolist())]
        if bias:
            # bias column tiled as (h, 1) tiles of 1x1 elements, appended after the spatial tiles (sparse.py:767-773)
            lc = c == K
            tb = _tile_tables(r[lc], torch.zeros_like(r[lc]), v[lc], (self.shape[0], 1), (h, 1))
            self._n_tile_entries += int(tb['tile_nnz'].sum().item())
            self._nnz_tiled += int(tb['tile_nnz'].sum().item())
            blocks += [(int(i), K, int(n_entries + k)) for (i, k) in zip(tb['block_i0'].tolist(), tb['block_tile'].tolist())]   # k_offset = len(tile dict), sparse.py:769
        self._blocks = sorted(blocks, key=lambda x: (x[0], x[1]))
        self._tab = dict(n_tiles=t00['n_tiles'])

    @classmethod
    def from_twin(cls, Wexec, module, inshape, outshape, tileshape, A, Ainv):
        """The same tiled view WITHOUT the expanded matrix (VGG16: 120 GB as CSR).  The reference takes the spatial tile
        structure from the channel-(0,0) block of the keyed matrix (sparse.py:740,752) and assumes channel-repeated keys;
        that block is the keyed Toeplitz matrix of a ONE-channel twin of the layer (first channel block of A and Ainv), a
        (Hout*Wout+1) x (Hin*Win+1) matrix built by the same compiler.  Every tile entry of the twin stands for a dense
        Cout x Cin block; the bias column is tiled from the bias vector.  Wexec: the layer's execution form (pattern
        groups, no CSR).  Equal to Conv2dTiledMatrix(expanded matrix) wherever that fits (tests/test_gpu_tiled_vgg.py)."""
        from . import sparse as _sp
        (Cin, Hin, Win) = [int(v) for v in inshape]
        (Cout, Hout, Wout) = [int(v) for v in outshape]
        (HoWo, HiWi) = (Hout * Wout, Hin * Win)
        (h, w) = (int(tileshape[0]), int(tileshape[1]))

        def first_channel(K, n):
            if K is None:
                return None
            assert isinstance(K, MonomialKey) and K.bias is None, 'tiled layers at this scale need permutation / gain keys'
            p = K.perm[:n]
            assert int(p.max()) < n, 'the key mixes channels: it is not tile-repeated (sparse.py:690-717 assumes it is)'
            return MonomialKey(np.concatenate([p, [n]]), np.concatenate([K.scale[:n], np.ones(1, dtype=np.float32)]))
        stride = module.stride[0]
        w00 = module.weight.detach().cpu().numpy()[0:1, 0:1]
        b0 = module.bias.detach().cpu().numpy()[0:1] if module.bias is not None else np.zeros(1, dtype=np.float32)
        T00 = _sp.keyed_toeplitz_conv2d((1, Hin, Win), w00, b0, stride, first_channel(A, HoWo), first_channel(Ainv, HiWi), build_groups=False)
        twin = cls(T00, (1, Hin, Win), (1, Hout, Wout), (h, w), bias=True, sanitycheck=False)
        self = cls.__new__(cls)
        SparseMatrix.__init__(self)
        (self.shape, self._pg) = (Wexec.shape, Wexec._pg)
        (self._nnz, self._device) = (Wexec.nnz(), Wexec._device if Wexec._data is None else Wexec._data.device)
        (self._inshape, self._outshape, self._tileshape) = (inshape, outshape, (h, w))
        (R, K) = (Cout * HoWo, Cin * HiWi)
        assert self.shape[0] in (R + 1,) and R % h == 0 and K % w == 0
        n_entries = twin._n_spatial_entries
        (self._n_spatial_entries, self._n_tile_entries, self._nnz_tiled) = (n_entries, n_entries, n_entries * Cout * Cin)
        blocks = [b for b in twin._blocks if b[1] < HiWi]                       # spatial blocks of the (0,0) block
        # bias column: value of row r = bias of its source channel (gains applied), last row = 1
        dev = self._device
        bias = module.bias.detach().cpu().numpy().astype(np.float32) if module.bias is not None else np.zeros(Cout, dtype=np.float32)
        src = np.arange(R) if A is None else A.perm[:R]
        v = _sp._offset_round(bias, np.min(bias))[src // HoWo]
        if A is not None:
            v = (A.scale[:R] * v).astype(np.float32)
        v = np.concatenate([v, np.ones(1, dtype=np.float32)])
        nz = np.nonzero(v)[0]
        rb = torch.from_numpy(nz.astype(np.int64)).to(dev)
        tb = _tile_tables(rb, torch.zeros_like(rb), torch.from_numpy(v[nz]).to(dev), (R + 1, 1), (h, 1))
        self._n_tile_entries += int(tb['tile_nnz'].sum().item())
        self._nnz_tiled += int(tb['tile_nnz'].sum().item())
        blocks += [(int(i), K, int(n_entries + k)) for (i, k) in zip(tb['block_i0'].tolist(), tb['block_tile'].tolist())]
        self._blocks = sorted(blocks, key=lambda x: (x[0], x[1]))
        self._tab = dict(n_tiles=twin._tab['n_tiles'])
        return self

    def expanded_nnz(self):
        return int(self._nnz) if self._data is None else SparseMatrix.nnz(self)

    def __repr__(self):
        return str('<keynet_b200.Conv2dTiledMatrix: H=%d, W=%d, tileshape=%s, tile entries=%d>' % (*self.shape, str(self.tileshape()), self._n_tile_entries))

    def blocks(self):
        return self._blocks

    def tiles(self):
        raise NotImplementedError('dense (Cout,Cin) tile dictionaries are not materialised; the unique value blocks live in the pattern groups')

    def nnz(self):
        return int(self._nnz_tiled)


def tile_keyed_layer(W, module, inshape, outshape, tileshape, A=None, Ainv=None):
    """What KeyedLayer does with tileshape (keynet/layer.py:38-41,62-65): conv -> Conv2dTiledMatrix, avgpool -> TiledMatrix.
    A conv layer that was built as pattern groups only (no CSR: VGG16 scale) takes its tile tables from the one-channel twin."""
    from torch import nn
    if isinstance(module, nn.Conv2d) and W._data is None:
        return Conv2dTiledMatrix.from_twin(W, module, inshape, outshape, tileshape, A, Ainv)
    if W._pg is None:
        W.optimize()
    if isinstance(module, nn.Conv2d):
        return Conv2dTiledMatrix(W, inshape, outshape, tileshape, bias=True, sanitycheck=False)
    if isinstance(module, nn.AvgPool2d):
        return TiledMatrix(W, tileshape)
    return W
