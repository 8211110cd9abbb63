"""Network driver of the keyed path on B200 (reference: keynet/system.py): key generation and chaining,
layer merging, the keyed sensor and the keyed model.  Same public names and keyword arguments as the
reference; keys are MonomialKey objects, matrices are compiled on the GPU, and the forward chain keeps
activations feature-major ([D+1, N]) on the device from encrypt to the last layer."""
import copy
import warnings
from collections import OrderedDict

import numpy as np
import torch
from torch import nn

from . import _native
from . import layer as _layer
from . import sparse as _sparse
from . import torch as _ktorch
from . import util as _util
from .blockpermute import hierarchical_block_permutation_matrix
from .globals import verbose
from .sparse import (MonomialKey, SparseKey, sparse_orthogonal_matrix, sparse_random_diagonally_dominant_doubly_stochastic_matrix, diagonal_affine_to_linear, sparse_permutation_matrix, sparse_identity_matrix, sparse_uniform_random_diagonal_matrix,
                     sparse_channelorder_to_blockorder_matrix, sparse_channelorder_to_pixelorder_matrix, sparse_affine_to_linear,
                     sparse_block_diagonal_repeat)


def _tolist(x):
    if isinstance(x, (list, tuple)):
        return list(x)
    if isinstance(x, np.ndarray):
        return x.reshape(-1).tolist()
    return [x]


# =============================================================================================
def _unlink(table, kind):
    """Take the layers whose name contains `kind` (dropout: the identity in eval()) out of the prev / next links of a netshape
    table.  Their entries stay, so they still draw a key and numpy's RNG stream stays aligned with the reference's
    (keynet/system.py:33-46) -- except names that are substrings of `kind` itself, which the reference's `k not in prefix`
    drops from the table (kept for stream parity).  Chains of such layers are followed to their end."""
    for v in table.values():
        for link in ('nextlayer', 'prevlayer'):
            while v[link] is not None and kind in v[link]:
                v[link] = table[v[link]][link]
    return OrderedDict((k, v) for (k, v) in table.items() if k not in kind)


class _Keying(object):
    """One keyed layer to build: where it goes in the keyed stack, the plain module it is compiled from, its shapes and keys."""
    __slots__ = ('name', 'module', 'inshape', 'outshape', 'A', 'Ainv', 'relu_name', 'relu_module')

    def __init__(self, name, module, inshape, outshape, A, Ainv, relu_name=None, relu_module=None):
        (self.name, self.module, self.inshape, self.outshape, self.A, self.Ainv, self.relu_name, self.relu_module) = (name, module, inshape, outshape, A, Ainv, relu_name, relu_module)


class KeyedModel(object):
    def __init__(self, net, inshape, inkey, f_layername_to_keypair, f_module_to_keyedmodule=None, do_output_encryption=False):
        """Key a plain network (keynet/system.py:27-119): every layer draws an output key pair, the input key of a layer is the
        inverse output key of its predecessor, conv -> batchnorm and layer -> ReLU pairs are merged into one keyed layer,
        dropout disappears, and every keyed layer is compiled on the GPU.

        Three passes: draw the keys in the reference's order (the RNG stream is part of the contract), plan the merges
        (_plan), build the planned layers."""
        net.eval()
        shapes = _unlink(_ktorch.netshape(net, inshape), 'dropout')
        last = shapes['output']['prevlayer']
        pairs = OrderedDict((k, f_layername_to_keypair(k, v['outshape'])) for (k, v) in shapes.items() if k not in ('input', 'output'))
        out_key = {k: (A if (k != last or do_output_encryption) else None) for (k, (A, _)) in pairs.items()}
        in_key = {k: (inkey if shapes[k]['prevlayer'] == 'input' else pairs[shapes[k]['prevlayer']][1]) for k in pairs}

        keyed = OrderedDict()
        for job in self._plan(net, shapes, out_key, in_key):
            if verbose():
                print('[keynet_b200.KeyedModel]: Keying "%s"' % job.name)
            L = f_module_to_keyedmodule(job.module, job.inshape, job.outshape, job.A, job.Ainv)
            if job.relu_name is not None:
                # a keyed layer that can apply the ReLU itself keeps an identity placeholder under the ReLU's name
                fused = hasattr(L, 'fuse_relu')
                keyed[job.name] = L.fuse_relu(True) if fused else L
                keyed[job.relu_name] = _layer.FusedReLU() if fused else copy.deepcopy(job.relu_module)
            else:
                keyed[job.name] = L
            if verbose():
                print('[keynet_b200.KeyedModel]:     %s' % str(keyed[job.name]))

        self._keynet = nn.Sequential(keyed)
        self._embeddingkey = pairs[last][1] if do_output_encryption else None
        self._imagekey = inkey
        self._layernames = set(k for (k, m) in net.named_children())
        self._outshape = shapes['output']['outshape']
        self._netshape = shapes

    @staticmethod
    def _plan(net, shapes, out_key, in_key):
        """The keyed layers to build, in stack order.  A layer followed by its batchnorm or by a ReLU is not keyed by itself:
        the follower keys it with the follower's output key, evaluated as (A_f . A_f_in) . A_layer in fp32 -- the
        reference's association (keynet/system.py:79-80,90-91); with gain keys the diagonal is fl32(fl32(a.(1/d)).d), not a."""
        for (k, m) in net.named_children():
            assert k in shapes and (k in out_key), 'layer "%s" of the network is missing from its shape table' % k
            v = shapes[k]
            if isinstance(m, nn.Dropout):
                continue
            if isinstance(m, nn.BatchNorm2d):
                conv = k.split('_')[0]
                assert '_bn' in k and v['prevlayer'] == conv, \
                    'a batchnorm layer must be named "<layer>_bn" and follow "<layer>" directly (e.g. conv3, conv3_bn); got "%s" after "%s"' % (k, v['prevlayer'])
                folded = copy.deepcopy(getattr(net, conv))
                (w, b) = _ktorch.fuse_conv2d_and_bn(folded.weight, folded.bias, m.running_mean, m.running_var, 1E-5, m.weight, m.bias)
                (folded.weight, folded.bias) = (torch.nn.Parameter(w), torch.nn.Parameter(b))
                yield _Keying(conv, folded, shapes[conv]['inshape'], v['outshape'], out_key[k].dot(in_key[k]).dot(out_key[conv]), in_key[conv])
            elif isinstance(m, nn.ReLU):
                prev = v['prevlayer']
                if '_bn' in prev:
                    warnings.warn('ReLU "%s" follows the batchnorm-merged layer "%s" and is keyed as a layer of its own; avoid batchnorm + ReLU sequences for efficient keying' % (k, prev))
                    yield _Keying(k, m, v['inshape'], v['outshape'], out_key[k], in_key[k])
                else:
                    yield _Keying(prev, getattr(net, prev), shapes[prev]['inshape'], shapes[prev]['outshape'],
                                  out_key[k].dot(in_key[k]).dot(out_key[prev]), in_key[prev], relu_name=k, relu_module=m)
            elif v['nextlayer'] is not None and (v['nextlayer'] == '%s_bn' % k or 'relu' in v['nextlayer']):
                continue                                         # keyed by the batchnorm / ReLU that follows
            else:
                yield _Keying(k, m, v['inshape'], v['outshape'], out_key[k], in_key[k])

    def __repr__(self):
        return self._keynet.__repr__()

    def __getattr__(self, attr):
        try:
            return self.__dict__['_keynet'].__getattr__(attr)
        except (KeyError, AttributeError):
            raise AttributeError(attr)

    def forward_linear(self, x_cipher):
        """Keyed layer stack on an N x (D+1) encrypted batch; returns N x (K+1) (still homogeneous).
        CPU input is staged to the GPU once and the result copied back once."""
        on_host = not x_cipher.is_cuda
        x = x_cipher.to(torch.device('cuda', torch.cuda.current_device()), non_blocking=True) if on_host else x_cipher
        y = self._keynet.forward(x)
        return y.cpu() if on_host else y

    def forward(self, img_cipher, outkey=None):
        """Reference semantics (keynet/system.py:130-133) for N=1; a batch N>1 returns (N, *outshape)
        instead of failing in reshape()."""
        outkey = outkey if outkey is not None else self.embeddingkey()
        y_cipher = self.forward_linear(img_cipher)
        y = self.decrypt(y_cipher, outkey) if outkey is not None else y_cipher
        N = y.shape[0]
        return _ktorch.linear_to_affine(y, self._outshape if N == 1 else (N,) + tuple(self._outshape))

    def decrypt(self, y_cipher, outkey=None):
        outkey = outkey if outkey is not None else self.embeddingkey()
        if outkey is None:
            return y_cipher
        W = outkey if isinstance(outkey, _sparse.SparseMatrix) else _sparse.SparseMatrix(outkey)
        return W.torchdot(y_cipher.t()).t()

    def imagekey(self):
        return self._imagekey

    def embeddingkey(self):
        return self._embeddingkey

    def public(self):
        self._imagekey = None
        self._embeddingkey = None
        return self

    def num_parameters(self):
        return sum([c.nnz() for (k, c) in self._keynet.named_children() if isinstance(c, _layer.KeyedLayer)])

    def layers(self):
        return self._layernames

    def keyedlayers(self):
        return [(k, c) for (k, c) in self._keynet.named_children() if isinstance(c, _layer.KeyedLayer)]


# =============================================================================================
class KeyedSensor(_layer.KeyedLayer):
    def __init__(self, inshape, keypair):
        assert isinstance(inshape, tuple) and len(inshape) == 3
        nn.Module.__init__(self)
        (self._encryptkey, self._decryptkey) = keypair
        self._inshape = (1, *inshape)
        self._tensor = None
        self._im = None
        self.W = _sparse.SparseMatrix(self._encryptkey)
        self._layertype = 'input'
        self._fused_relu = False
        self._repr = 'KeyedSensor'
        self._enc = None

    def encrypt_into(self, images, Y):
        """Y[D+1][N] = A . affine_to_linear(images[N][D])^T on the current stream.  A monomial image key (permutation x gain
        [+ bias column]) is applied while the batch is transposed (kn_encrypt_monomial_t, one pass); a general key takes the
        two-pass route (homogenise + transpose, then the CSR SpMM)."""
        (N, D) = (int(images.shape[0]), int(np.prod(images.shape[1:])))
        assert tuple(Y.shape) == (D + 1, N) and Y.is_contiguous() and images.is_contiguous()
        K = self._encryptkey
        if isinstance(K, MonomialKey) and K.shape[0] == D + 1 and (K.bias is None or K.perm[-1] == D):
            if self._enc is None:
                dev = Y.device
                row_of_col = np.empty(D + 1, dtype=np.int32); row_of_col[K.perm] = np.arange(D + 1, dtype=np.int32)
                scale_of_col = np.empty(D + 1, dtype=np.float32); scale_of_col[K.perm] = K.scale
                self._enc = (torch.from_numpy(row_of_col).to(dev), torch.from_numpy(scale_of_col).to(dev), None if K.bias is None else torch.from_numpy(K.bias).to(dev))
            (rc, sc, rb) = self._enc
            _native.check(_native.lib().kn_encrypt_monomial_t(_native.ptr(images), N, D, _native.ptr(rc), _native.ptr(sc), _native.ptr(rb), _native.ptr(Y), N, _native.stream_ptr()))
            return 1
        X = torch.empty((D + 1, N), dtype=torch.float32, device=Y.device)
        _native.check(_native.lib().kn_affine_to_linear_t(_native.ptr(images), N, D, _native.ptr(X), N, _native.stream_ptr()))
        _sparse.spmm(self.W, X, out=Y)
        return 2

    def __repr__(self):
        return str('<KeyedSensor: height=%d, width=%d, channels=%d>' % (self._inshape[2], self._inshape[3], self._inshape[1]))

    def fromtensor(self, x):
        if x is not None:
            self._tensor = x.detach().clone().to(torch.float32)
        return self

    def tensor(self):
        return self._tensor.unsqueeze(0) if self._tensor.ndim == 3 else self._tensor

    def astensor(self):
        return self.tensor()

    def totensor(self):
        return self.astensor()

    def keypair(self):
        return (self._encryptkey, self._decryptkey)

    def key(self):
        return self._decryptkey

    def isloaded(self):
        return self._tensor is not None

    def isencrypted(self):
        """An encrypted batch is N x (C*H*W+1).  (The reference only recognises N == 1, system.py:243-245,
        and would re-encrypt an encrypted batch; any N is recognised here.)"""
        return self.isloaded() and self._tensor.ndim == 2 and self._tensor.shape[1] == int(np.prod(self._inshape)) + 1

    def encrypt(self):
        """N x C x H x W -> N x (C*H*W+1), homogenised and multiplied by the image key (system.py:250-255).
        The homogeneous coordinate and the batch-major -> feature-major transpose are one fused kernel."""
        assert self.isloaded(), "Load image first"
        if not self.isencrypted():
            x = KeyedSensor.tensor(self)
            on_host = not x.is_cuda
            dev = torch.device('cuda', torch.cuda.current_device())
            xd = x.to(dev, non_blocking=True).contiguous()
            (N, D) = (xd.shape[0], int(np.prod(xd.shape[1:])))
            Y = torch.empty((D + 1, N), dtype=torch.float32, device=dev)
            self.encrypt_into(xd, Y)
            y = Y.t()
            self._tensor = y.cpu() if on_host else y
        return self

    def decrypt(self):
        """N x (C*H*W+1) -> N x C x H x W with the private image key (system.py:257-263)."""
        assert self.isloaded(), "Load image first"
        if self.isencrypted():
            x_raw = _layer.KeyedLayer.decrypt(self, self._decryptkey, self._tensor)
            N = x_raw.shape[0]
            self._tensor = _ktorch.linear_to_affine(x_raw, (N,) + tuple(self._inshape[1:]))
        return self

    # ---- image-side I/O (keynet/system.py:173-235; PIL images where the reference uses vipy) -------------------
    def _from_pil(self, im):
        (C, H, W) = self._inshape[1:]
        im = im.convert('L' if C == 1 else 'RGB')
        a = np.asarray(im, dtype=np.float32)
        a = a[:, :, None] if a.ndim == 2 else a
        self._tensor = torch.from_numpy(np.ascontiguousarray(a.transpose(2, 0, 1))).unsqueeze(0)     # HxWxC [0,255] -> 1xCxHxW
        return self

    def load(self, imgfile, imagekey=None):
        """Load an image file as the sensor's 1xCxHxW tensor (values 0..255, resized to the sensor shape); with
        `imagekey` (returned by save()) the file is an encrypted image and is decrypted while loading."""
        from PIL import Image
        im = Image.open(imgfile)
        (C, H, W) = self._inshape[1:]
        if imagekey is not None:
            self._from_pil(im)
            assert tuple(self._tensor.shape) == tuple(self._inshape), 'encrypted image must have the sensor shape'
            x_linear = _ktorch.affine_to_linear((1.0 / 255.0) * self._tensor)                          # [0,255] -> [0,1]
            x = imagekey.apply(x_linear.numpy().transpose())                                            # DecryptKey . mat2gray^-1
            self._tensor = _ktorch.linear_to_affine(torch.as_tensor(np.ascontiguousarray(x.transpose()), dtype=torch.float32), self._inshape)
        else:
            self._from_pil(im.resize((W, H), Image.BILINEAR))
        self._im = im
        return self

    def fromimage(self, im):
        """im: PIL image with the sensor's height / width."""
        assert (im.size[1], im.size[0]) == tuple(self._inshape[2:]), 'image must have the sensor shape'
        self._im = im
        return self._from_pil(im)

    def save(self, outfile='/tmp/out.png'):
        """Write the ENCRYPTED image as an 8-bit PNG (values min-max normalised by a mat2gray key) and return
        (outfile, imagekey); load(outfile, imagekey) decrypts it again (to 8-bit quantisation)."""
        from PIL import Image
        assert self.isencrypted() and self._tensor.shape[0] == 1
        x_linear = self._tensor.detach().cpu().numpy().transpose()                                      # (D+1) x 1
        (A, Ainv) = _sparse.mat2gray(x_linear.flatten()[:-1])
        x = _ktorch.linear_to_affine(torch.as_tensor(np.ascontiguousarray(A.apply(x_linear).transpose()), dtype=torch.float32), self._inshape)
        a = np.clip(np.rint(255.0 * x[0].numpy().transpose(1, 2, 0)), 0, 255).astype(np.uint8)       # 1xCxHxW [0,1] -> HxWxC uint8
        Image.fromarray(a[:, :, 0] if a.shape[2] == 1 else a).save(outfile)
        return (outfile, self._decryptkey.dot(Ainv))

    def asimage(self):
        """The current tensor (encrypted or not) as a min-max normalised uint8 PIL image."""
        from PIL import Image
        x = self._tensor
        if self.isencrypted():
            x = _ktorch.linear_to_affine(x[0:1].cpu(), self._inshape)
        a = x[0] if x.ndim == 4 else x
        a = a.detach().cpu().numpy().transpose(1, 2, 0).astype(np.float32)
        (lo, hi) = (float(a.min()), float(a.max()))
        a = np.uint8(255 * (a - lo) / (hi - lo)) if hi > lo else np.zeros(a.shape, dtype=np.uint8)
        return Image.fromarray(a[:, :, 0] if a.shape[2] == 1 else a)

    def toimage(self):
        return self.asimage()

    def show(self):
        self.asimage().show()
        return self


class PublicKeyedSensor(KeyedSensor):
    def __init__(self, inshape):
        assert isinstance(inshape, tuple) and len(inshape) == 3
        n = int(np.prod(inshape)) + 1
        super(PublicKeyedSensor, self).__init__(inshape, (sparse_identity_matrix(n), sparse_identity_matrix(n)))

    def __repr__(self):
        return str('<PublicKeyedSensor: height=%d, width=%d, channels=%d>' % (self._inshape[2], self._inshape[3], self._inshape[1]))

    def encrypt(self):
        raise ValueError('PublicKeyedSensor has no encryption keys')

    def decrypt(self):
        raise ValueError('PublicKeyedSensor has no decryption keys')

    def tensor(self):
        assert self.isloaded(), "Load image first"
        if not self.isencrypted():
            KeyedSensor.encrypt(self)
        return self._tensor


# =============================================================================================
def layergen(module, inshape, outshape, A, Ainv, tileshape=None, backend='b200', rows=None, keep_csr=True):
    """Keyed-layer factory (keynet/system.py:303-314).  tileshape is snapped to divisors of the spatial size;
    the only backend is 'b200' ('scipy' is accepted as an alias so reference call sites keep working)."""
    if tileshape is not None:
        new_tileshape = (_util.find_closest_positive_divisor(outshape[1], tileshape[0]), _util.find_closest_positive_divisor(inshape[1], tileshape[1]))
        if verbose() and new_tileshape != tileshape:
            print('[layergen]: Ragged spatial tileshape=%s, forcing non-ragged tileshape "%s" for inshape="%s", outshape="%s"' % (str(tileshape), str(new_tileshape), str(inshape), str(outshape)))
        tileshape = new_tileshape
    if backend in ('b200', 'scipy'):
        return _layer.KeyedLayer(module, inshape, outshape, A, Ainv, tileshape=tileshape, rows=rows, keep_csr=keep_csr)
    raise ValueError('invalid backend "%s"' % backend)


# ---- key generation: one builder per option of each of the four key axes -------------------------------------------------
class _KeySpec(object):
    """Everything the per-axis key builders need for one activation shape: sizes, the block geometry after snapping the
    block size (keynet/system.py:336-343) and the numeric options."""

    def __init__(self, shape, blocksize, tileshape, strict, alpha, beta, gamma, hierarchical_blockshape, hierarchical_permute_at_level, seed):
        (self.channels, self.height, self.width) = [int(v) for v in shape]
        self.shape = (self.channels, self.height, self.width)
        self.N = self.channels * self.height * self.width
        (self.alpha, self.beta, self.gamma, self.seed, self.tileshape) = (alpha, beta, gamma, seed, tileshape)
        (self.hblockshape, self.hlevels) = (hierarchical_blockshape, hierarchical_permute_at_level)
        (self.blocksize, self.plane, self.blocknumel) = (blocksize, None, None)       # plane = size of the matrix one block key is tiled over
        if blocksize is not None:
            assert tileshape is None or (blocksize == tileshape[0] and blocksize == tileshape[1]), 'blocksize must equal the tile size'
            if self.height == 1 and self.width == 1:
                (self.blocksize, self.plane, self.blocknumel) = (self.N, self.N, self.N)              # a vector is one block
            else:
                if not strict and (self.height % blocksize != 0 or self.width % blocksize != 0):
                    assert self.height == self.width, "Image must be square to correct ragged blocksize"
                    self.blocksize = _util.find_closest_positive_divisor(self.height, blocksize)
                (self.plane, self.blocknumel) = (self.height * self.width, self.blocksize * self.blocksize)

    def need(self, **what):
        for (name, ok) in what.items():
            assert ok, 'this key needs a valid "%s"' % name

    def tile_over_image(self, B):
        """Block key B repeated over the spatial plane, then over the channels."""
        return sparse_block_diagonal_repeat(sparse_block_diagonal_repeat(B, (self.plane, self.plane)), (self.N, self.N))

    def tiled_bias(self, b):
        """Per-block bias vector repeated to the full length N (column vector)."""
        return np.tile(b, int(np.ceil(self.N / self.blocknumel)))[0:self.N].reshape(self.N, 1)


def _identity_pair(k):
    return (sparse_identity_matrix(k.N), sparse_identity_matrix(k.N))


def _global_permutation(k):
    assert k.tileshape is None, "Global permutation is not tile compressible"
    return sparse_permutation_matrix(k.N, withinverse=True)


def _hierarchical(twist):
    def build(k, c=None, cinv=None):
        k.need(hierarchical_blockshape=k.hblockshape is not None, hierarchical_permute_at_level=k.hlevels is not None)
        levels = _tolist(k.hlevels)
        if max(k.height, k.width) / np.power(2, max(levels)) < 8 or (k.height == 1 and k.width == 1):
            levels = []                                                     # too small to permute at these levels (system.py:365-366)
        (Q, Qinv) = sparse_channelorder_to_pixelorder_matrix(k.shape, withinverse=True)
        (G, Ginv) = hierarchical_block_permutation_matrix((k.height, k.width, k.channels), k.hblockshape, levels, min_blocksize=8, seed=k.seed,
                                                          twist=twist, withinverse=True, strict=False)
        (G, Ginv) = (Qinv.dot(G).dot(Q), Qinv.dot(Ginv).dot(Q))             # CxHxW -> HxWxC -> permute -> CxHxW
        if c is not None:
            (G, Ginv) = (c.dot(G).dot(cinv), c.dot(Ginv).dot(cinv))         # block memory order
        return (G, Ginv)
    build.wants_memoryorder = True
    return build


def _global_givens(k):
    k.need(alpha=k.alpha is not None)
    assert k.tileshape is None, "Global givens rotation orthogonal matrix is not tile compressible"
    return sparse_orthogonal_matrix(k.N, int(k.alpha), balanced=True, withinverse=True)


def _local_permutation(k):
    k.need(blocksize=k.blocksize is not None and k.height == k.width)
    g = k.tile_over_image(sparse_permutation_matrix(k.blocknumel))
    return (g, g.transpose())


def _local_doubly_stochastic(k):
    k.need(blocksize=k.blocksize is not None and k.height == k.width, alpha=k.alpha is not None)
    assert k.blocksize < 8192, "Blocksize %d must be less than 8192, since doubly_stochastic requires the direct inverse of a dense matrix" % k.blocksize
    (g, ginv) = sparse_random_diagonally_dominant_doubly_stochastic_matrix(k.blocknumel, int(k.alpha), withinverse=True)
    # the reference carries these blocks (and every matrix compiled with them) in float64; this path is fp32 throughout
    return (k.tile_over_image(g.astype(np.float32)), k.tile_over_image(ginv.astype(np.float32)))


def _local_givens(k):
    k.need(blocksize=k.blocksize is not None and k.height == k.width, alpha=k.alpha is not None)
    (g, ginv) = sparse_orthogonal_matrix(k.blocknumel, int(k.alpha), balanced=True, withinverse=True)
    (Pb, Pbinv) = sparse_permutation_matrix(k.blocknumel, withinverse=True)
    return (k.tile_over_image(Pb.dot(g)), k.tile_over_image(ginv.dot(Pbinv)))


# photometric builders return HOMOGENEOUS pairs (the bias column lives in the last column)
def _homogeneous(pair):
    return (sparse_affine_to_linear(pair[0]), sparse_affine_to_linear(pair[1]))


def _photo_identity(k):
    return _homogeneous(_identity_pair(k))


def _global_gain(k):
    assert k.tileshape is None, "Global permutation is not tile compressible"
    k.need(beta=k.beta is not None and k.beta > 0)
    return _homogeneous(sparse_uniform_random_diagonal_matrix(k.N, k.beta, bias=1, withinverse=True))


def _global_bias(draw):
    def build(k):
        k.need(gamma=k.gamma is not None and k.gamma > 0)
        return diagonal_affine_to_linear(sparse_identity_matrix(k.N), draw(k), withinverse=True)
    return build


def _global_affine(k):
    assert k.tileshape is None, "Global permutation is not tile compressible"
    k.need(beta=k.beta is not None and k.beta > 0, gamma=k.gamma is not None and k.gamma > 0)
    D = sparse_uniform_random_diagonal_matrix(k.N, k.beta, bias=1)
    return diagonal_affine_to_linear(D, k.gamma * np.random.rand(k.N, 1), withinverse=True)


def _blockwise_constant_bias(k):
    k.need(blocksize=k.blocksize is not None)
    per_block = k.gamma * np.random.rand(int(np.ceil(k.N // k.blocksize)), 1)          # one draw per block, spread over the block
    return per_block.dot(np.ones((1, k.blocknumel))).flatten()[0:k.N].reshape(k.N, 1)


def _local_gain(k):
    k.need(blocksize=k.blocksize is not None, beta=k.beta is not None and k.beta > 0)
    (d, dinv) = sparse_uniform_random_diagonal_matrix(k.blocknumel, k.beta, bias=1, withinverse=True)
    return _homogeneous((_repeat_diagonal(d, k.N), _repeat_diagonal(dinv, k.N)))


def _local_bias(k):
    k.need(blocksize=k.blocksize is not None, gamma=k.gamma is not None and k.gamma > 0)
    return diagonal_affine_to_linear(sparse_identity_matrix(k.N), bias=k.tiled_bias(k.gamma * np.random.rand(k.blocknumel)), withinverse=True)


def _local_affine(k):
    k.need(blocksize=k.blocksize is not None, beta=k.beta is not None and k.beta > 0, gamma=k.gamma is not None and k.gamma > 0)
    d = sparse_uniform_random_diagonal_matrix(k.blocknumel, k.beta, bias=1)
    return diagonal_affine_to_linear(_repeat_diagonal(d, k.N), bias=k.tiled_bias(k.gamma * np.random.rand(k.blocknumel)), withinverse=True)


def _testing_only(k):
    raise ValueError('blockwise_constant_bias supported for global_photometric testing only')


_KEY_AXES = OrderedDict([      # in the order the reference draws from the RNG (keynet/system.py:357-464)
    ('global geometric', {'identity': _identity_pair, 'permutation': _global_permutation, 'hierarchical_permutation': _hierarchical(False),
                          'hierarchical_rotation': _hierarchical(True), 'givens_orthogonal': _global_givens}),
    ('local geometric', {'identity': _identity_pair, 'permutation': _local_permutation, 'doubly_stochastic': _local_doubly_stochastic, 'givens_orthogonal': _local_givens}),
    ('global photometric', {'identity': _photo_identity, 'uniform_random_gain': _global_gain, 'uniform_random_affine': _global_affine,
                            'uniform_random_bias': _global_bias(lambda k: k.gamma * np.random.rand(k.N, 1)),
                            'linear_bias': _global_bias(lambda k: (k.gamma / float(k.N)) * np.array(range(0, k.N)).reshape(k.N, 1)),
                            'blockwise_constant_bias': _global_bias(_blockwise_constant_bias)}),
    ('local photometric', {'identity': _photo_identity, 'uniform_random_gain': _local_gain, 'uniform_random_affine': _local_affine, 'uniform_random_bias': _local_bias,
                           'blockwise_constant_bias': _testing_only}),
])


def keygen(shape, global_geometric, local_geometric, global_photometric, local_photometric, memoryorder='channel', alpha=None, beta=None, gamma=None, seed=None,
           hierarchical_blockshape=None, hierarchical_permute_at_level=None, blocksize=None, tileshape=None, strict=False):
    """(A, A^-1) for one activation shape: A = C^-1 . p . g . P . G . C with C the memory order, G / g the global / local
    geometric keys and P / p the global / local photometric keys (keynet/system.py:317-469).

    Each axis is one table lookup (_KEY_AXES); the builders draw from numpy's global RNG in the reference's order (global
    geometric, local geometric, global photometric, local photometric).  Permutation / gain keys are MonomialKeys, bias /
    affine keys MonomialKeys with a bias column (a family closed under products), Givens-orthogonal and doubly-stochastic
    keys general SparseKeys composed on the host and compiled into the layers by the GPU SpGEMM."""
    if seed is not None:
        np.random.seed(seed)
    k = _KeySpec(shape, blocksize, tileshape, strict, alpha, beta, gamma, hierarchical_blockshape, hierarchical_permute_at_level, seed)
    if memoryorder == 'channel':
        (c, cinv) = (None, None)
        (C, Cinv) = _photo_identity(k)
    elif memoryorder == 'block':
        assert k.blocksize is not None, 'block memory order needs a blocksize'
        (c, cinv) = sparse_channelorder_to_blockorder_matrix(k.shape, k.blocksize, withinverse=True)
        (C, Cinv) = _homogeneous((c, cinv))
    else:
        raise ValueError("Invalid memory order '%s' - must be in '%s'" % (memoryorder, str(['channel', 'block'])))
    pairs = []
    for ((axis, table), choice) in zip(_KEY_AXES.items(), (global_geometric, local_geometric, global_photometric, local_photometric)):
        if choice not in table:
            raise ValueError("Invalid %s transform '%s' - must be in '%s'" % (axis, choice, str(sorted(table))))
        build = table[choice]
        pair = build(k, c, cinv) if getattr(build, 'wants_memoryorder', False) else build(k)
        pairs.append(_homogeneous(pair) if 'geometric' in axis else pair)
    ((G, Ginv), (g, ginv), (P, Pinv), (p, pinv)) = pairs
    # right-to-left products, in the reference's association (system.py:467-468): the fp32 roundings of gain keys depend on it
    A = C
    for K in (G, P, g, p, Cinv):
        A = K.dot(A)
    Ainv = C
    for K in (pinv, ginv, Pinv, Ginv, Cinv):
        Ainv = K.dot(Ainv)
    return (A, Ainv)


def _repeat_diagonal(D, n):
    """diag block repeated with period len(D), truncated at n (reference: sparse_block_diagonal, sparse.py:215-235)."""
    h = D.shape[0]
    reps = int(np.ceil(n / float(h)))
    return MonomialKey(np.arange(n), np.tile(D.scale, reps)[0:n])


def keypair_policy(global_photometric='identity', local_photometric='identity', global_geometric='identity', local_geometric='identity', memoryorder='channel',
                   alpha=None, beta=None, gamma=None, hierarchical_blockshape=None, hierarchical_permute_at_level=None, blocksize=None, tileshape=None):
    """layername, shape -> (A, Ainv).  Layers named '*relu*' get keys that commute with ReLU: global transforms are
    dropped, a requested local photometric key becomes a local gain and a requested local geometric key a local
    permutation (keynet/system.py:476-482)."""
    def f_keypair(layername, shape):
        relu = 'relu' in layername
        return keygen(shape,
                      global_photometric=global_photometric if not relu or global_photometric == 'identity' else 'identity',
                      local_photometric=local_photometric if not relu or local_photometric == 'identity' else 'uniform_random_gain',
                      global_geometric=global_geometric if not relu or global_geometric == 'identity' else 'identity',
                      local_geometric=local_geometric if not relu or local_geometric == 'identity' else 'permutation',
                      memoryorder=memoryorder, blocksize=blocksize, tileshape=tileshape, alpha=alpha, beta=beta, gamma=gamma,
                      hierarchical_blockshape=hierarchical_blockshape, hierarchical_permute_at_level=hierarchical_permute_at_level)
    return f_keypair


def Keynet(inshape, net=None, backend='b200', global_photometric='identity', local_photometric='identity', global_geometric='identity', local_geometric='identity', memoryorder='channel',
           do_output_encryption=False, alpha=None, beta=None, gamma=None, hierarchical_blockshape=None, hierarchical_permute_at_level=None, blocksize=None, tileshape=None,
           keep_csr=True):
    """(sensor, model) for a plain torch net (keynet/system.py:472-486).  Output keys of layers named '*relu*'
    are restricted to keys that commute with ReLU: no global transforms, local gain / local permutation only."""
    # keep_csr=False (not in the reference): free each layer's canonical CSR once its pattern-grouped form exists
    f_layergen = lambda module, inshape, outshape, A, Ainv: layergen(module, inshape, outshape, A, Ainv, tileshape=tileshape, backend=backend, keep_csr=keep_csr)
    f_keypair = keypair_policy(global_photometric=global_photometric, local_photometric=local_photometric, global_geometric=global_geometric, local_geometric=local_geometric,
                               memoryorder=memoryorder, alpha=alpha, beta=beta, gamma=gamma, hierarchical_blockshape=hierarchical_blockshape,
                               hierarchical_permute_at_level=hierarchical_permute_at_level, blocksize=blocksize, tileshape=tileshape)
    sensor = KeyedSensor(inshape, f_keypair('input', inshape))
    model = KeyedModel(net, inshape, sensor.key(), f_keypair, f_layergen, do_output_encryption=do_output_encryption) if net is not None else None
    return (sensor, model)


def IdentityKeynet(inshape, net, backend='b200'):
    return Keynet(inshape, net, backend=backend)


def PermutationKeynet(inshape, net, do_output_encryption=False):
    return Keynet(inshape, net, global_geometric='permutation', do_output_encryption=do_output_encryption)


def TiledIdentityKeynet(inshape, net, tilesize, keep_csr=True):
    """keep_csr=False (not in the reference): VGG16-scale networks -- conv / linear layers exist only as pattern groups and
    the tiled views come from one-channel twins (tiled.Conv2dTiledMatrix.from_twin), never from a 120 GB expansion."""
    return Keynet(inshape, net, tileshape=(tilesize, tilesize), keep_csr=keep_csr)


def TiledPermutationKeynet(inshape, net, tilesize, keep_csr=True):
    return Keynet(inshape, net, local_geometric='permutation', tileshape=(tilesize, tilesize), blocksize=tilesize, keep_csr=keep_csr)


def TiledOrthogonalKeynet(inshape, net, tilesize, hierarchical_permute_at_level=(0, 1)):
    return Keynet(inshape, net, tileshape=(tilesize, tilesize),
                  global_geometric='hierarchical_permutation', hierarchical_blockshape=(2, 2), hierarchical_permute_at_level=hierarchical_permute_at_level,
                  global_photometric='identity',
                  local_geometric='givens_orthogonal', alpha=tilesize, blocksize=tilesize,
                  local_photometric='uniform_random_affine', beta=0.1, gamma=100.0,
                  memoryorder='block')
