"""KeyedLayer: one conv / avgpool / linear / ReLU layer compiled to W_hat = A . W . Ainv on the GPU and
executed as a sparse x dense-batch product (reference: keynet/layer.py:15-106)."""
import time

import numpy as np
import torch
from torch import nn

from . import sparse
from .globals import verbose
from .sparse import SparseMatrix, MonomialKey, SparseKey


class FusedReLU(nn.Module):
    """Placeholder for the un-keyed nn.ReLU the reference appends after a keyed conv/linear
    (keynet/system.py:92).  On B200 the ReLU runs in the epilogue of the preceding SpMM kernel, so this
    module is the identity; it only keeps the layer names of the keyed network unchanged."""

    def forward(self, x):
        return x

    def extra_repr(self):
        return 'fused into the previous KeyedLayer SpMM epilogue'


class KeyedLayer(nn.Module):
    def __init__(self, module, inshape, outshape, A, Ainv, tileshape=None, rows=None, col_remap=None, n_cols_phys=None, keep_csr=True, build_groups=True):
        """module: nn.Conv2d | nn.AvgPool2d | nn.Linear | nn.ReLU; A / Ainv: MonomialKey (A may be None for the
        last layer); rows=(r0, r1) or an index array: build and hold only those rows of W_hat (row shard);
        col_remap / n_cols_phys: physical position of every canonical input column (gathered layout, dist.py);
        keep_csr=False: conv / linear layers are built as pattern groups only and never exist as a CSR (VGG16 scale);
        build_groups=False: canonical CSR only (key-compile sweep, parity / export)."""
        super(KeyedLayer, self).__init__()
        self._layertype = str(type(module))
        self._inshape = inshape
        self._outshape = outshape
        self._tileshape = tileshape
        self._fused_relu = False
        self._rows = rows
        self._build_groups = bool(build_groups)
        want_csr = bool(keep_csr) or not build_groups        # tiled layers without a CSR take their tile tables from a one-channel twin
        (self._A, self._Ainv) = (A, Ainv)
        assert A is None or isinstance(A, (MonomialKey, SparseKey)), 'A must be a key (MonomialKey / SparseKey)'
        assert isinstance(Ainv, (MonomialKey, SparseKey)), 'Ainv must be a key (MonomialKey / SparseKey)'
        t0 = time.time()
        if isinstance(A, SparseKey) or isinstance(Ainv, SparseKey):
            # general keys (Givens-orthogonal / doubly stochastic blocks, keynet/system.py:398-410): the un-keyed layer matrix,
            # then the two products of keynet/layer.py:35,46,59,70 as GPU SpGEMMs (csrc/spgemm.cu)
            self._init_general(module, inshape, outshape, A, Ainv)
            if rows is not None or col_remap is not None:
                # row shard of a general-key layer: the full W_hat is compiled (general keys are used on small networks), then
                # this rank's rows are gathered and the columns moved to their gathered positions
                self.W = sparse.shard_compiled(self.W, rows, col_remap, n_cols_phys)
            self._finish(module, inshape, outshape, tileshape, keep_csr, t0)
            return

        if isinstance(module, nn.Conv2d):
            assert len(module.kernel_size) == 1 or len(module.kernel_size) == 2 and (module.kernel_size[0] == module.kernel_size[1]), "Kernel must be square"
            assert len(module.stride) == 1 or len(module.stride) == 2 and (module.stride[0] == module.stride[1]), "Strides must be isotropic"
            assert len(inshape) == 3, "Inshape must be (C,H,W) for the shape of the tensor at the input to this layer"
            assert module.padding[0] == module.kernel_size[0] // 2 and module.padding[1] == module.kernel_size[1] // 2, "Padding is assumed to be equal to (kernelsize-1)/2"
            stride = module.stride[0]
            self._repr = 'Conv2d: in_channels=%d, out_channels=%d, kernel_size=%s, stride=%s' % (module.in_channels, module.out_channels, str(module.kernel_size), str(stride))
            bias = module.bias.detach().cpu().numpy() if module.bias is not None else np.zeros(module.out_channels, dtype=np.float32)
            self.W = sparse.keyed_toeplitz_conv2d(inshape, module.weight.detach().cpu().numpy(), bias, stride, A, Ainv, rows=rows, col_remap=col_remap, n_cols_phys=n_cols_phys,
                                                  build_groups=build_groups, want_csr=want_csr)

        elif isinstance(module, nn.ReLU):
            # explicit keyed ReLU (only after a batchnorm merge, keynet/system.py:97-99): W = A . Ainv, then ReLU
            self._repr = 'ReLU'
            assert col_remap is None and (rows is None or isinstance(rows, tuple)), 'explicit keyed ReLU layers are not row-sharded by groups'
            self.W = SparseMatrix(A.dot(Ainv))
            if rows is not None:
                self.W = self.W.row_slice(*rows)

        elif isinstance(module, nn.AvgPool2d):
            assert isinstance(module.kernel_size, int) or len(module.kernel_size) == 2 and (module.kernel_size[0] == module.kernel_size[1]), "Kernel must be square"
            assert isinstance(module.stride, int) or len(module.stride) == 2 and (module.stride[0] == module.stride[1]), "Strides must be isotropic"
            assert len(inshape) == 3, "Inshape must be (C,H,W) for the shape of the tensor at the input to this layer"
            stride = module.stride if isinstance(module.stride, int) else module.stride[0]
            kernel_size = module.kernel_size if isinstance(module.kernel_size, int) else module.kernel_size[0]
            self._repr = 'AvgPool2d: kernel_size=%s, stride=%s' % (str(kernel_size), str(stride))
            # as in the reference, padding / ceil_mode of the module are ignored: centred k x k windows, divisor k*k
            self.W = sparse.keyed_toeplitz_avgpool2d(inshape, kernel_size, stride, A, Ainv, rows=rows, col_remap=col_remap, n_cols_phys=n_cols_phys)

        elif isinstance(module, nn.Linear):
            self._repr = 'Linear: in_features=%d, out_features=%d' % (module.in_features, module.out_features)
            self.W = sparse.keyed_linear(module.weight.detach(), module.bias.detach() if module.bias is not None else None, A, Ainv, rows=rows, col_remap=col_remap, n_cols_phys=n_cols_phys,
                                         build_groups=build_groups, want_csr=want_csr)

        elif isinstance(module, nn.BatchNorm2d):
            raise ValueError('batchnorm layer should be named "mylayer_bn" for batchnorm of "mylayer" and should come right before "mylayer" to merge keyed layers')
        elif isinstance(module, nn.Dropout):
            raise ValueError('dropout layer should be skipped during keying and removed from final network')
        else:
            raise ValueError('unsupported layer type "%s"' % str(type(module)))

        self._finish(module, inshape, outshape, tileshape, keep_csr, t0)

    def _init_general(self, module, inshape, outshape, A, Ainv):
        if isinstance(module, nn.Conv2d):
            assert module.kernel_size[0] == module.kernel_size[1], "Kernel must be square"
            assert module.stride[0] == module.stride[1], "Strides must be isotropic"
            assert module.padding[0] == module.kernel_size[0] // 2 and module.padding[1] == module.kernel_size[1] // 2, "Padding is assumed to be equal to (kernelsize-1)/2"
            stride = module.stride[0]
            self._repr = 'Conv2d: in_channels=%d, out_channels=%d, kernel_size=%s, stride=%s' % (module.in_channels, module.out_channels, str(module.kernel_size), str(stride))
            bias = module.bias.detach().cpu().numpy() if module.bias is not None else np.zeros(module.out_channels, dtype=np.float32)
            W = sparse.keyed_toeplitz_conv2d(inshape, module.weight.detach().cpu().numpy(), bias, stride, None, MonomialKey(np.arange(int(np.prod(inshape)) + 1)), build_groups=False)
        elif isinstance(module, nn.AvgPool2d):
            stride = module.stride if isinstance(module.stride, int) else module.stride[0]
            kernel_size = module.kernel_size if isinstance(module.kernel_size, int) else module.kernel_size[0]
            self._repr = 'AvgPool2d: kernel_size=%s, stride=%s' % (str(kernel_size), str(stride))
            W = sparse.keyed_toeplitz_avgpool2d(inshape, kernel_size, stride, None, MonomialKey(np.arange(int(np.prod(inshape)) + 1)))
            W._pg = None
        elif isinstance(module, nn.Linear):
            self._repr = 'Linear: in_features=%d, out_features=%d' % (module.in_features, module.out_features)
            n = module.in_features + 1
            W = sparse.keyed_linear(module.weight.detach(), module.bias.detach() if module.bias is not None else None, None, MonomialKey(np.arange(n)))
        elif isinstance(module, nn.ReLU):
            self._repr = 'ReLU'
            self.W = SparseMatrix(SparseKey.coerce(A).dot(SparseKey.coerce(Ainv)).astype(np.float32))
            return
        else:
            raise ValueError('unsupported layer type "%s"' % str(type(module)))
        if A is not None:
            W = A.dot(W)                  # left product first, like A.dot(W).dot(Ainv)
        self.W = W.matmul(Ainv)

    def _finish(self, module, inshape, outshape, tileshape, keep_csr, t0):
        if tileshape is None and self._build_groups:
            self.W.optimize()          # pattern-grouped execution format for batched forward (no-op if the builder made it)
        if tileshape is not None:
            from .tiled import tile_keyed_layer
            self.W = tile_keyed_layer(self.W, module, inshape, outshape, tileshape, self._A, self._Ainv)
        (self._A, self._Ainv) = (None, None)              # keys are not kept with the layer
        if not keep_csr and getattr(self.W, '_pg', None) is not None and self.W._data is not None:
            self.W.drop_csr()
            torch.cuda.empty_cache()
        if verbose():
            torch.cuda.synchronize()
            print('[KeyedLayer]: %s compiled on GPU in %1.3f seconds, nnz=%d' % (self._repr, time.time() - t0, self.nnz()))

    def extra_repr(self):
        return str('<%s, backend=b200, shape=%s, nnz=%d%s>' % (self._repr, str(self.W.shape), self.nnz(), ', +relu' if self._fused_relu else ''))

    def fuse_relu(self, flag=True):
        self._fused_relu = bool(flag)
        return self

    def forward(self, x_affine):
        """x_affine: N x (Din+1) (CPU or CUDA).  Returns N x (Dout+1); same math as keynet/layer.py:88-93."""
        relu = self._fused_relu or ('ReLU' in self._layertype)
        return self.W.torchdot(x_affine.t(), relu=relu).t()

    def decrypt(self, Ainv, x_affine):
        """Decrypt the output of this layer (x_affine) using supplied key Ainv (keynet/layer.py:95-99)."""
        if not isinstance(Ainv, SparseMatrix):
            Ainv = SparseMatrix(Ainv)
        return Ainv.torchdot(x_affine.t()).t()

    def spy(self, mindim=256, showdim=1024, range=None):
        """Picture of the keyed layer matrix (keynet/layer.py:104-106)."""
        return self.W.spy(mindim, showdim, range)

    def nnz(self):
        assert self.W is not None, "Layer not keyed"
        return self.W.nnz()
