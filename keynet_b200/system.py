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
class KeyedModel(object):
    def __init__(self, net, inshape, inkey, f_layername_to_keypair, f_module_to_keyedmodule=None, do_output_encryption=False):
        """Walk net.named_children(), give every layer an output key, chain them (Ainv of a layer is the
        inverse output key of its predecessor), merge conv->relu and conv->bn, drop dropout, and compile each
        keyed layer on the GPU (keynet/system.py:27-119)."""
        net.eval()
        netshape = _ktorch.netshape(net, inshape)

        # dropout is the identity in eval(): unlink it.  As in the reference the entries stay in the table, so
        # they still draw a key below and the RNG stream stays aligned (system.py:33-46).
        for prefix in ['dropout']:
            netshape = OrderedDict((k, v) for (k, v) in netshape.items() if k not in prefix)
            for (k, v) in netshape.items():
                if v['nextlayer'] is not None and prefix in v['nextlayer']:
                    v['nextlayer'] = netshape[v['nextlayer']]['nextlayer']
                elif v['prevlayer'] is not None and prefix in v['prevlayer']:
                    v['prevlayer'] = netshape[v['prevlayer']]['prevlayer']

        o = netshape['output']['prevlayer']
        outkeys = OrderedDict((k, f_layername_to_keypair(k, v['outshape'])) for (k, v) in netshape.items() if k not in ('input', 'output'))
        layerkey = {k: {'A': outkeys[k][0] if (k != o or do_output_encryption) else None,
                        'Ainv': inkey if netshape[k]['prevlayer'] == 'input' else outkeys[netshape[k]['prevlayer']][1]}
                    for k in outkeys}
        layerkey['input'] = inkey
        layerkey['output'] = outkeys[o][1] if do_output_encryption else None

        keyed = OrderedDict()
        for (k, m) in net.named_children():
            if verbose():
                print('[keynet_b200.KeyedModel]: Keying "%s"' % k)
            assert k in layerkey, 'Key not found for layer "%s"' % k
            assert k in netshape, 'Layer name not found in net shape for layer "%s"' % k
            shp = netshape[k]

            if isinstance(m, nn.BatchNorm2d):
                assert '_bn' in k, "Batchnorm layers must be named 'mylayername_bn' for corresponding linear layer mylayername.  (e.g. 'conv3_bn')"
                k_prev = k.split('_')[0]
                assert shp['prevlayer'] == k_prev, "Batchnorm layer named 'mylayer_bn' must come right after 'mylayer' (e.g. 'conv3_bn' must come right after 'conv3')"
                m_prev = copy.deepcopy(getattr(net, k_prev))
                (w, b) = _ktorch.fuse_conv2d_and_bn(m_prev.weight, m_prev.bias, m.running_mean, m.running_var, 1E-5, m.weight, m.bias)
                (m_prev.weight, m_prev.bias) = (torch.nn.Parameter(w), torch.nn.Parameter(b))
                B = layerkey[k]['A'].dot(layerkey[k]['Ainv'])   # batchnorm out-key times inverse conv out-key
                keyed[k_prev] = f_module_to_keyedmodule(m_prev, netshape[k_prev]['inshape'], shp['outshape'], B.dot(layerkey[k_prev]['A']), layerkey[k_prev]['Ainv'])

            elif isinstance(m, nn.ReLU):
                k_prev = shp['prevlayer']
                if '_bn' not in k_prev:
                    # key the skipped predecessor with the ReLU's out-key.  The effective key is evaluated as
                    # (A_relu . A_prev^-1) . A_prev in fp32, exactly like the reference (system.py:90-91): with
                    # gain keys its diagonal is fl32(fl32(a.(1/d)).d), not a -- simplifying it changes last bits.
                    B = layerkey[k]['A'].dot(layerkey[k]['Ainv'])
                    L = f_module_to_keyedmodule(getattr(net, k_prev), netshape[k_prev]['inshape'], netshape[k_prev]['outshape'], B.dot(layerkey[k_prev]['A']), layerkey[k_prev]['Ainv'])
                    keyed[k_prev] = L.fuse_relu(True) if hasattr(L, 'fuse_relu') else L
                    keyed[k] = _layer.FusedReLU() if hasattr(L, 'fuse_relu') else copy.deepcopy(m)
                else:
                    warnings.warn('Keying ReLU since previous layer "%s" is already keyed - Avoid sequential batchnorm and ReLU layers for efficient keying' % k_prev)
                    keyed[k] = f_module_to_keyedmodule(m, shp['inshape'], shp['outshape'], layerkey[k]['A'], layerkey[k]['Ainv'])

            elif isinstance(m, nn.Dropout):
                pass
            elif shp['nextlayer'] is not None and (('%s_bn' % k) == shp['nextlayer'] or 'relu' in shp['nextlayer']):
                pass    # merged into the batchnorm / ReLU that follows
            else:
                keyed[k] = f_module_to_keyedmodule(m, shp['inshape'], shp['outshape'], layerkey[k]['A'], layerkey[k]['Ainv'])
            if verbose() and k in keyed:
                print('[keynet_b200.KeyedModel]:     %s' % str(keyed[k]))

        self._keynet = nn.Sequential(keyed)
        self._embeddingkey = layerkey['output']
        self._imagekey = layerkey['input']
        self._layernames = set(k for (k, m) in net.named_children())
        self._outshape = netshape['output']['outshape']
        self._netshape = netshape

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


def keygen(shape, global_geometric, local_geometric, global_photometric, local_photometric, memoryorder='channel', alpha=None, beta=None, gamma=None, seed=None,
           hierarchical_blockshape=None, hierarchical_permute_at_level=None, blocksize=None, tileshape=None, strict=False):
    """Compose A = C^-1 . p . g . P . G . C and its inverse for one activation shape (keynet/system.py:317-469).

    RNG draws happen in the reference's order: global geometric, local geometric, global photometric, local
    photometric.  All photometric options are supported (gain keys are monomial; bias / affine keys are monomial
    plus a bias column, a family closed under products).  Givens-orthogonal and doubly-stochastic geometric keys are general
    sparse keys (sparse.SparseKey): composed on the host like the reference's, compiled into the layers by the GPU SpGEMM."""
    allowable_memoryorder = set(['channel', 'block'])
    allowable_global_geometric = set(['identity', 'permutation', 'hierarchical_permutation', 'hierarchical_rotation', 'givens_orthogonal'])
    allowable_local_geometric = set(['identity', 'permutation', 'doubly_stochastic', 'givens_orthogonal'])
    allowable_photometric = set(['identity', 'uniform_random_gain', 'uniform_random_affine', 'uniform_random_bias', 'constant_bias', 'linear_bias', 'blockwise_constant_bias'])

    (channels, height, width) = shape
    N = int(np.prod(shape))
    if seed is not None:
        np.random.seed(seed)

    (H, blocknumel) = (None, None)
    if blocksize is not None:
        if tileshape is not None:
            assert blocksize == tileshape[0] and blocksize == tileshape[1]
        if height == 1 and width == 1:
            (blocksize, H, blocknumel) = (N, N, N)
        elif not strict and (height % blocksize != 0 or width % blocksize != 0):
            assert height == width, "Image must be square to correct ragged blocksize"
            blocksize = _util.find_closest_positive_divisor(height, blocksize)
            (H, blocknumel) = (height * width, blocksize * blocksize)
        else:
            (H, blocknumel) = (height * width, blocksize * blocksize)

    if memoryorder == 'channel':
        (c, cinv) = (sparse_identity_matrix(N), sparse_identity_matrix(N))
    elif memoryorder == 'block':
        assert blocksize is not None
        (c, cinv) = sparse_channelorder_to_blockorder_matrix(shape, blocksize, withinverse=True)
    else:
        raise ValueError("Invalid memory order '%s' - must be in '%s'" % (memoryorder, str(allowable_memoryorder)))
    (C, Cinv) = (sparse_affine_to_linear(c), sparse_affine_to_linear(cinv))

    if global_geometric == 'identity':
        (G, Ginv) = (sparse_identity_matrix(N), sparse_identity_matrix(N))
    elif global_geometric == 'permutation':
        assert tileshape is None, "Global permutation is not tile compressible"
        (G, Ginv) = sparse_permutation_matrix(N, withinverse=True)
    elif global_geometric in ('hierarchical_permutation', 'hierarchical_rotation'):
        assert hierarchical_blockshape is not None and hierarchical_permute_at_level is not None
        levels = _tolist(hierarchical_permute_at_level)
        levels = levels if max(height, width) / np.power(2, max(levels)) >= 8 else []
        levels = [] if (height == 1 and width == 1) else levels
        (Q, Qinv) = sparse_channelorder_to_pixelorder_matrix((channels, height, width), withinverse=True)
        (G, Ginv) = hierarchical_block_permutation_matrix((height, width, channels), hierarchical_blockshape, levels, min_blocksize=8, seed=seed,
                                                          twist=(global_geometric == 'hierarchical_rotation'), withinverse=True, strict=False)
        (G, Ginv) = (Qinv.dot(G).dot(Q), Qinv.dot(Ginv).dot(Q))     # CxHxW -> HxWxC -> permute -> CxHxW
        if memoryorder != 'channel':
            (G, Ginv) = (c.dot(G).dot(cinv), c.dot(Ginv).dot(cinv))
    elif global_geometric == 'givens_orthogonal':
        assert alpha is not None
        assert tileshape is None, "Global givens rotation orthogonal matrix is not tile compressible"
        (G, Ginv) = sparse_orthogonal_matrix(N, int(alpha), balanced=True, withinverse=True)
    else:
        raise ValueError("Invalid global geometric transform '%s' - must be in '%s'" % (global_geometric, str(allowable_global_geometric)))
    (G, Ginv) = (sparse_affine_to_linear(G), sparse_affine_to_linear(Ginv))

    if local_geometric == 'identity':
        (g, ginv) = (sparse_identity_matrix(N), sparse_identity_matrix(N))
    elif local_geometric == 'permutation':
        assert blocksize is not None and height == width
        g = sparse_block_diagonal_repeat(sparse_block_diagonal_repeat(sparse_permutation_matrix(blocknumel), (H, H)), (N, N))   # spatial, then channel repeat
        ginv = g.transpose()
    elif local_geometric == 'doubly_stochastic':
        assert blocksize is not None and alpha is not None and height == width
        assert blocksize < 8192, "Blocksize %d must be less than 8192, since doubly_stochastic requires the direct inverse of a dense matrix" % blocksize
        (g, ginv) = sparse_random_diagonally_dominant_doubly_stochastic_matrix(blocknumel, int(alpha), withinverse=True)
        # the reference carries these blocks (and every matrix compiled with them) in float64; this path is fp32 throughout
        g = sparse_block_diagonal_repeat(sparse_block_diagonal_repeat(g.astype(np.float32), (H, H)), (N, N))        # spatial, then channel repeat
        ginv = sparse_block_diagonal_repeat(sparse_block_diagonal_repeat(ginv.astype(np.float32), (H, H)), (N, N))
    elif local_geometric == 'givens_orthogonal':
        assert alpha is not None and blocksize is not None and height == width
        (g, ginv) = sparse_orthogonal_matrix(blocknumel, int(alpha), balanced=True, withinverse=True)
        (Pb, Pbinv) = sparse_permutation_matrix(blocknumel, withinverse=True)
        (g, ginv) = (Pb.dot(g), ginv.dot(Pbinv))
        g = sparse_block_diagonal_repeat(sparse_block_diagonal_repeat(g, (H, H)), (N, N))                             # spatial, then channel repeat
        ginv = sparse_block_diagonal_repeat(sparse_block_diagonal_repeat(ginv, (H, H)), (N, N))
    else:
        raise ValueError("Invalid local geometric transform '%s' - must be in '%s'" % (local_geometric, str(allowable_local_geometric)))
    (g, ginv) = (sparse_affine_to_linear(g), sparse_affine_to_linear(ginv))

    if global_photometric == 'identity':
        (P, Pinv) = (sparse_affine_to_linear(sparse_identity_matrix(N)), sparse_affine_to_linear(sparse_identity_matrix(N)))
    elif global_photometric == 'uniform_random_gain':
        assert tileshape is None, "Global permutation is not tile compressible"
        assert beta is not None and beta > 0
        (P, Pinv) = sparse_uniform_random_diagonal_matrix(N, beta, bias=1, withinverse=True)
        (P, Pinv) = (sparse_affine_to_linear(P), sparse_affine_to_linear(Pinv))
    elif global_photometric == 'uniform_random_bias':
        assert gamma is not None and gamma > 0
        (P, Pinv) = diagonal_affine_to_linear(sparse_identity_matrix(N), gamma * np.random.rand(N, 1), withinverse=True)
    elif global_photometric == 'linear_bias':
        assert gamma is not None and gamma > 0
        (P, Pinv) = diagonal_affine_to_linear(sparse_identity_matrix(N), (gamma / float(N)) * np.array(range(0, N)).reshape(N, 1), withinverse=True)
    elif global_photometric == 'uniform_random_affine':
        assert tileshape is None, "Global permutation is not tile compressible"
        assert beta is not None and beta > 0 and gamma is not None and gamma > 0
        P = sparse_uniform_random_diagonal_matrix(N, beta, bias=1)
        (P, Pinv) = diagonal_affine_to_linear(P, gamma * np.random.rand(N, 1), withinverse=True)
    elif global_photometric == 'blockwise_constant_bias':
        assert gamma is not None and gamma > 0
        assert blocksize is not None
        bias = gamma * np.random.rand(int(np.ceil(N // blocksize)), 1).dot(np.ones((1, blocknumel))).flatten()[0:N].reshape(N, 1)
        (P, Pinv) = diagonal_affine_to_linear(sparse_identity_matrix(N), bias, withinverse=True)
    else:
        raise ValueError("Invalid global photometric transform '%s' - must be in '%s'" % (global_photometric, str(allowable_photometric)))

    if local_photometric == 'identity':
        (p, pinv) = (sparse_affine_to_linear(sparse_identity_matrix(N)), sparse_affine_to_linear(sparse_identity_matrix(N)))
    elif local_photometric == 'uniform_random_gain':
        assert blocksize is not None
        assert beta is not None and beta > 0
        (p, pinv) = sparse_uniform_random_diagonal_matrix(blocknumel, beta, bias=1, withinverse=True)
        (p, pinv) = (_repeat_diagonal(p, N), _repeat_diagonal(pinv, N))
        (p, pinv) = (sparse_affine_to_linear(p), sparse_affine_to_linear(pinv))
    elif local_photometric == 'uniform_random_bias':
        assert blocksize is not None
        assert gamma is not None and gamma > 0
        bias = np.tile(gamma * np.random.rand(blocknumel), int(np.ceil(N / blocknumel)))[0:N].reshape(N, 1)
        (p, pinv) = diagonal_affine_to_linear(sparse_identity_matrix(N), bias=bias, withinverse=True)
    elif local_photometric == 'uniform_random_affine':
        assert blocksize is not None
        assert beta is not None and beta > 0 and gamma is not None and gamma > 0
        p = sparse_uniform_random_diagonal_matrix(blocknumel, beta, bias=1)
        bias = np.tile(gamma * np.random.rand(blocknumel), int(np.ceil(N / blocknumel)))[0:N].reshape(N, 1)
        (p, pinv) = diagonal_affine_to_linear(_repeat_diagonal(p, N), bias=bias, withinverse=True)
    elif local_photometric == 'blockwise_constant_bias':
        raise ValueError('blockwise_constant_bias supported for global_photometric testing only')
    else:
        raise ValueError("Invalid local photometric transform '%s' - must be in '%s'" % (local_photometric, str(allowable_photometric)))

    A = Cinv.dot(p.dot(g.dot(P.dot(G.dot(C)))))
    Ainv = Cinv.dot(Ginv.dot(Pinv.dot(ginv.dot(pinv.dot(C)))))
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


def TiledIdentityKeynet(inshape, net, tilesize):
    return Keynet(inshape, net, tileshape=(tilesize, tilesize))


def TiledPermutationKeynet(inshape, net, tilesize):
    return Keynet(inshape, net, local_geometric='permutation', tileshape=(tilesize, tilesize), blocksize=tilesize)


def TiledOrthogonalKeynet(inshape, net, tilesize, hierarchical_permute_at_level=(0, 1)):
    return Keynet(inshape, net, tileshape=(tilesize, tilesize),
                  global_geometric='hierarchical_permutation', hierarchical_blockshape=(2, 2), hierarchical_permute_at_level=hierarchical_permute_at_level,
                  global_photometric='identity',
                  local_geometric='givens_orthogonal', alpha=tilesize, blocksize=tilesize,
                  local_photometric='uniform_random_affine', beta=0.1, gamma=100.0,
                  memoryorder='block')
