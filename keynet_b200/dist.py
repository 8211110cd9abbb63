"""Multi-GPU execution of the keyed path on one 8xB200 box, one process per GPU.

  * small networks (LeNet, AllConvNet): data-parallel REPLICAS -- every rank keys the same network (same seed, same
    keys) and runs its own slice of the batch; no collective on the data path (bench.py --gpus N).
  * VGG16-scale keyed layers (15 G non-zeros): ROW-SHARDED.  Every rank compiles and holds only its rows of every
    W_hat; a layer is  Y_local = W_hat[rows_r, :] . X_full  followed by one all-gather of the feature-major
    activations over NVLink (torch.distributed / NCCL all_gather_into_tensor), ReLU fused before the gather.

Sharding is expressed as ONE MORE KEY.  Rows of a conv layer are re-ordered pixel-major before they are cut into
`world` equal chunks, so every rank owns WHOLE pattern groups (all M output channels of its pixels -- the unit the
tensor-core kernel works on); the gathered activation buffer is therefore in shard-major order, and that order is
folded into the next layer's column map at compile time (an integer composition with the input key).  Nothing is
permuted at run time, and the homogeneous coordinate is kept as a local row of ones on every rank.

The planner (`plan_rows`, `LayerShard`) is pure numpy so it is covered by world_size-2 gloo tests on CPU.
"""
import os
import numpy as np
import torch
from torch import nn


def plan_rows(module, outshape, world, A=None):
    """Shard-major order of the output rows of one keyed layer W_hat = A.W.Ainv (homogeneous row excluded).

    Returns (order, chunk): `order` = row indices of W_hat; rank r owns order[r*chunk:(r+1)*chunk] (the last ranks may
    own fewer / no rows).  Spatial layers are cut by the PIXEL of the underlying Toeplitz row (row r of W_hat is row
    A.perm[r] of W), raster order, all channels of a pixel together: chunk boundaries fall between pattern groups, and
    a shard reads only its own pixels plus a halo of the previous layer whatever permutation the keys apply -- which is
    what lets the fused path (peer row masks) replace the all-gather by a halo exchange."""
    (C, H, W) = [int(s) for s in outshape]
    R = C * H * W
    if A is None:
        src = np.arange(R, dtype=np.int64)
    elif hasattr(A, 'perm'):
        src = np.asarray(A.perm[:R], dtype=np.int64)            # Toeplitz row of every W_hat row
        assert len(A.perm) == R + 1, 'output key must be homogeneous'
    else:
        # general key (several entries per row): the first stored column stands in for the row's position -- block-local
        # keys mix rows of one small spatial block, so this keeps the cut spatial; any assignment of rows to ranks is valid
        assert A.shape[0] == R + 1, 'output key must be homogeneous'
        src = np.minimum(np.asarray(A.indices[A.indptr[:-1][:R]], dtype=np.int64), R - 1)
    assert src.max() < R, 'output key must be homogeneous'
    if isinstance(module, (nn.Conv2d, nn.AvgPool2d)) and H * W > 1:
        (channel, pixel) = (src // (H * W), src % (H * W))
        order = np.argsort(pixel * C + channel, kind='stable').astype(np.int64)       # (pixel, channel) <- (channel, pixel)
        px_chunk = -(-(H * W) // world)
        chunk = px_chunk * C
    else:
        order = np.argsort(src, kind='stable').astype(np.int64)
        chunk = -(-R // world)
    return (order, int(chunk))


def peer_row_masks(need, rank, chunk, n_mine):
    """need: bool [world, n_phys] -- need[q, p] = rank q reads gathered position p in its next layer.  Returns the
    uint8 mask of this rank's n_mine output rows: bit q set = store the row into rank q's buffer (own bit always set)."""
    need = np.asarray(need, dtype=bool)
    world = need.shape[0]
    assert world <= 8
    mine = need[:, rank * chunk:rank * chunk + n_mine]
    mask = np.zeros(n_mine, dtype=np.uint8)
    for q in range(world):
        mask |= (mine[q].astype(np.uint8) << q).astype(np.uint8)
    mask |= np.uint8(1 << rank)
    return mask


def sync_sets(reads_k, reads_next, rank):
    """Neighbourhood synchronisation after layer k (fused row-sharded forward).  reads_k[q, p] = rank q reads rows that rank p
    produced in layer k; reads_next the same for layer k+1 (None after the last layer).  Returns (signal_mask, wait_mask):
    this rank waits for the ranks whose layer-k rows it reads (RAW) and for the ranks it will store rows to in layer k+1 --
    layer k+1 on this rank overwrites, on those ranks, the ping-pong buffer their layer k is still reading (WAR) -- and it
    signals every rank that waits for it.  With shards cut by image rows both sets are the two spatial neighbours."""
    reads_k = np.asarray(reads_k, dtype=bool)
    world = reads_k.shape[0]

    def waits(r):
        w = reads_k[r].copy()
        if reads_next is not None:
            w |= np.asarray(reads_next, dtype=bool)[:, r]
        w[r] = False
        return w
    wait = waits(rank)
    signal = np.array([waits(q)[rank] for q in range(world)], dtype=bool)
    signal[rank] = False
    to_mask = lambda v: int(sum(1 << i for i in range(world) if v[i]))
    return (to_mask(signal), to_mask(wait))


class LayerShard(object):
    """Bookkeeping of one row-sharded layer: which canonical rows this rank computes and where every canonical row
    lives in the gathered buffer [world*chunk + 1] (last position = homogeneous coordinate)."""

    def __init__(self, module, outshape, rank, world, A=None):
        (order, chunk) = plan_rows(module, outshape, world, A)
        R = len(order)
        self.chunk = chunk
        self.n_phys = world * chunk + 1
        self.my_rows = order[rank * chunk:min(R, (rank + 1) * chunk)]
        pos = np.empty(R + 1, dtype=np.int64)
        pos[order] = np.arange(R)            # canonical row -> gathered position (rank-major, then order within the chunk)
        pos[R] = world * chunk               # homogeneous coordinate
        self.position = pos
        self.n_rows = R + 1


class ShardedLayerGen(object):
    """f_module_to_keyedmodule callback for system.KeyedModel: compiles, for every keyed layer, only this rank's rows
    with the previous layer's gathered layout folded into the column map."""

    def __init__(self, rank, world, inshape, keep_csr=True):
        (self.rank, self.world) = (int(rank), int(world))
        self.keep_csr = keep_csr
        self.prev_position = None            # first layer reads the sensor output in canonical order
        self.prev_n_phys = int(np.prod(inshape)) + 1
        self.layers = []

    def __call__(self, module, inshape, outshape, A, Ainv):
        from . import layer as _layer
        shard = LayerShard(module, outshape, self.rank, self.world, A)
        L = _layer.KeyedLayer(module, inshape, outshape, A, Ainv, rows=shard.my_rows,
                              col_remap=self.prev_position, n_cols_phys=self.prev_n_phys, keep_csr=self.keep_csr)
        L._shard = shard
        (self.prev_position, self.prev_n_phys) = (shard.position, shard.n_phys)
        self.layers.append(L)
        return L


class ShardedKeyedModel(object):
    """Row-sharded keyed network: same constructor contract as system.Keynet(...)[1] plus (rank, world, group)."""

    def __init__(self, inshape, net, rank, world, group=None, fused=False, selective=True, keep_csr=True, **keynet_kwargs):
        """fused=True: no NCCL on the data path -- every SpMM epilogue stores its rows straight into all ranks' gathered
        activation buffers (torch symmetric memory = NVLink peer mappings, the `kn_peers` argument of every kn_spmm_*), one device-side barrier per
        layer; with selective=True (default) a row is stored only into the buffers of the ranks whose next layer reads it
        (kn_peers.row_mask).  fused=False: local SpMM + torch.distributed all_gather_into_tensor (NCCL)."""
        from . import system
        self.rank, self.world, self.group = int(rank), int(world), group
        self.fused = bool(fused)
        self.selective = bool(selective)       # fused only: store a row only to the peers whose next layer reads it
        self.flag_sync = os.environ.get('KEYNET_B200_FLAG_SYNC', '1') != '0'                  # selective only: neighbourhood flags (kn_peer_sync) instead of a barrier over all ranks per layer
        self._symm = {}
        f_keypair = system.keypair_policy(**keynet_kwargs)
        self.sensor = system.KeyedSensor(inshape, f_keypair('input', inshape))
        self._gen = ShardedLayerGen(rank, world, inshape, keep_csr=keep_csr)
        self._model = system.KeyedModel(net, inshape, self.sensor.key(), f_keypair, self._gen)
        self._outshape = self._model._outshape
        self.layers = self._gen.layers
        self._masks = None
        self.time_layers = False      # record CUDA events around every layer (SpMM + barrier / all-gather)
        self._events = []

    def _peer_masks(self, dev):
        """Per layer: which peers read each of this rank's output rows (one all-gather of need bitmaps at set-up)."""
        if self._masks is None:
            import torch.distributed as dist
            group = self.group if self.group is not None else dist.group.WORLD
            masks = []
            reads = []                      # reads[k][q, p]: rank q reads rows of layer k produced by rank p
            for (k, L) in enumerate(self.layers):
                sh = L._shard
                n_mine = len(sh.my_rows)
                if k + 1 < len(self.layers):
                    need = self.layers[k + 1].W.used_columns().to(torch.uint8)
                    assert need.numel() == sh.n_phys
                    allneed = torch.empty((self.world, sh.n_phys), dtype=torch.uint8, device=dev)
                    dist.all_gather_into_tensor(allneed, need.reshape(1, -1).contiguous(), group=group)
                    an = allneed.cpu().numpy() != 0
                    m = peer_row_masks(an, self.rank, sh.chunk, n_mine)
                    slots = an[:, :self.world * sh.chunk].reshape(self.world, self.world, sh.chunk)
                    reads.append(slots.any(axis=2))
                else:
                    m = np.full(n_mine, (1 << self.world) - 1, dtype=np.uint8)      # the logits go to everyone
                    reads.append(np.ones((self.world, self.world), dtype=bool))
                masks.append(torch.from_numpy(m).to(dev))
            self._masks = masks
            self._sync = [sync_sets(reads[k], reads[k + 1] if k + 1 < len(reads) else None, self.rank) for k in range(len(reads))]
        return self._masks

    def peer_store_fraction(self):
        """Rows x destinations actually stored / (rows x world): 1.0 = full all-gather."""
        if self._masks is None:
            return None
        (sent, full) = (0, 0)
        for (L, m) in zip(self.layers, self._masks):
            bits = m.cpu().numpy()
            sent += int(np.unpackbits(bits.reshape(-1, 1), axis=1).sum())
            full += len(bits) * self.world
        return sent / max(full, 1)

    def num_parameters_local(self):
        return sum(L.nnz() for L in self.layers)

    def _symm_buffers(self, N, dev):
        """Two ping-pong symmetric buffers (layer k writes buffer k%2 on every rank while buffer (k-1)%2 is being read)."""
        if N not in self._symm:
            import torch.distributed as dist
            import torch.distributed._symmetric_memory as symm_mem
            group = self.group if self.group is not None else dist.group.WORLD
            rows = max(L._shard.n_phys for L in self.layers)
            bufs = []
            for _ in range(2):
                t = symm_mem.empty((rows * N,), dtype=torch.float32, device=dev)
                h = symm_mem.rendezvous(t, group)
                bufs.append((t, h))
            self._symm[N] = bufs
        return self._symm[N]

    def _forward_fused(self, X, N, dev):
        """Layer k writes ping-pong buffer (k + parity) % 2 on every rank that needs the row; one symmetric-memory barrier
        per layer.  The parity flips by the layer count from one forward to the next, so the first layer of forward f+1
        never stores into the buffer that still holds the output of forward f (a fast peer may run ahead by up to one
        layer: its stores would otherwise race this rank's read of the logits)."""
        from . import _native
        from .sparse import spmm
        assert N >= 32 and N % 4 == 0, 'fused row-sharded forward runs on batches padded to a multiple of 4, at least 32 (forward_linear pads)'
        bufs = self._symm_buffers(N, dev)
        masks = self._peer_masks(dev) if self.selective else [None] * len(self.layers)
        parity = getattr(self, '_parity', 0)
        for (k, L) in enumerate(self.layers):
            sh = L._shard
            relu = L._fused_relu or ('ReLU' in L._layertype)
            self._stamp()
            (t, h) = bufs[(k + parity) % 2]
            Yfull = t[:sh.n_phys * N].view(sh.n_phys, N)
            n_mine = len(sh.my_rows)
            slot = self.rank * sh.chunk * N * 4                         # byte offset of this rank's slot in every buffer
            if n_mine > 0:
                peers = _native.Peers([int(p) + slot for p in h.buffer_ptrs], masks[k])
                spmm(L.W, X, relu=relu, out=Yfull[self.rank * sh.chunk:self.rank * sh.chunk + n_mine], peers=peers)
            Yfull[-1].fill_(1.0)                                        # homogeneous coordinate: local
            self._stamp()
            if self.selective and self.flag_sync:
                self._peer_sync(k, dev)                                 # signal the ranks that read these rows, wait for the ranks read from
            else:
                h.barrier()                                             # every rank's stores have landed everywhere
            X = Yfull
        self._stamp()
        self._parity = (parity + len(self.layers)) % 2
        return X

    def _peer_sync(self, k, dev):
        """Neighbourhood synchronisation after layer k (kn_peer_sync): flags in symmetric memory, one epoch per call."""
        from . import _native
        import ctypes
        if getattr(self, '_flags', None) is None:
            import torch.distributed as dist
            import torch.distributed._symmetric_memory as symm_mem
            group = self.group if self.group is not None else dist.group.WORLD
            t = symm_mem.empty((16,), dtype=torch.int32, device=dev)
            t.zero_()
            h = symm_mem.rendezvous(t, group)
            h.barrier()                                                 # every rank's flags are zero before anyone signals
            self._flags = (t, h, (ctypes.c_uint64 * 8)(*[int(p) for p in h.buffer_ptrs]))
            self._epoch = torch.zeros(1, dtype=torch.int32, device=dev)      # advanced on the device by every kn_peer_sync
            self._sync_timeout = torch.zeros(1, dtype=torch.int32, device=dev)
        (signal, wait) = self._sync[k]
        _native.check(_native.lib().kn_peer_sync(self._flags[2], self.world, self.rank, signal, wait, _native.ptr(self._epoch), _native.ptr(self._sync_timeout), _native.stream_ptr()))

    def sync_timed_out(self):
        """True if a peer failed to arrive at some neighbourhood synchronisation (checked by tests / the bench after a run)."""
        return getattr(self, '_sync_timeout', None) is not None and bool(self._sync_timeout.item() != 0)

    def _stamp(self):
        if self.time_layers:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            self._events.append(e)

    def layer_times_ms(self):
        """time_layers=True: [(layer, spmm_ms, barrier / all-gather ms)] of the LAST forward."""
        torch.cuda.synchronize()
        n = len(self.layers)
        ev = self._events[-(2 * n + 1):]
        out = []
        for k in range(n):
            out.append((k, ev[2 * k].elapsed_time(ev[2 * k + 1]), ev[2 * k + 1].elapsed_time(ev[2 * k + 2])))
        self._events = []
        return out

    def forward_linear(self, x_cipher):
        """x_cipher: N x (D+1) encrypted batch, identical on every rank.  Returns N x (K+1) on every rank."""
        import torch.distributed as dist
        from .sparse import spmm
        dev = torch.device('cuda', torch.cuda.current_device())
        X = x_cipher.to(dev).t().contiguous()                         # feature-major [D+1, N]
        N = X.shape[1]
        if self.fused:
            Np = max(32, (N + 3) // 4 * 4)                              # granularity of the grouped kernels (peer stores use ld = Np)
            if Np != N:
                Xp = torch.zeros((X.shape[0], Np), dtype=torch.float32, device=dev)
                Xp[:, :N] = X
                X = Xp
            X = self._forward_fused(X, Np, dev)
            return X[self._out_positions(dev)][:, :N].t().contiguous()
        for L in self.layers:
            sh = L._shard
            relu = L._fused_relu or ('ReLU' in L._layertype)
            self._stamp()
            Yfull = torch.empty((sh.n_phys, N), dtype=torch.float32, device=dev)
            Yloc = Yfull[self.rank * sh.chunk:(self.rank + 1) * sh.chunk]       # this rank's slot of the gathered buffer
            n_mine = len(sh.my_rows)
            if n_mine < sh.chunk:
                Yloc[n_mine:].zero_()
            if n_mine > 0:
                spmm(L.W, X, relu=relu, out=Yloc[:n_mine])
            self._stamp()
            if self.world > 1:
                dist.all_gather_into_tensor(Yfull[:self.world * sh.chunk], Yloc, group=self.group)
            Yfull[-1].fill_(1.0)                                      # homogeneous coordinate: local, never communicated
            X = Yfull
        self._stamp()
        # last layer: undo the shard-major order (a gather of K+1 rows)
        return X[self._out_positions(dev)].t().contiguous()

    def capture(self, N):
        """Capture the fused forward for batches of N (a multiple of 4, at least 32) into CUDA graphs -- one per ping-pong parity --
        after it has run eagerly at least once (symmetric buffers, flags, split-K scratch exist).  At 8 GPUs a VGG16 layer
        takes 50-500 us of GPU time, less than the Python + ctypes work to launch it: replaying a graph removes those gaps.
        forward_graph(X) then copies X [D+1, N] into the static input and replays."""
        assert self.fused and self.selective and self.flag_sync and N >= 32 and N % 4 == 0
        dev = torch.device('cuda', torch.cuda.current_device())
        D1 = self.layers[0].W.shape[1]
        self._gx = torch.zeros((D1, N), dtype=torch.float32, device=dev)
        self._graphs = {}
        torch.cuda.synchronize()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        saved = getattr(self, '_parity', 0)
        with torch.cuda.stream(side):
            for parity in (0, 1):
                self._parity = parity
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=side):
                    y = self.forward_linear(self._gx.t())
                self._graphs[parity] = (g, y)
        torch.cuda.current_stream().wait_stream(side)
        self._parity = saved
        self._graph_N = N
        return self

    def forward_graph(self, X):
        """X: encrypted batch, feature-major [D+1, N] on this device (the sensor's encrypt_into layout).  Returns N x (K+1)."""
        parity = getattr(self, '_parity', 0)
        (g, y) = self._graphs[parity]
        if X.data_ptr() != self._gx.data_ptr():
            self._gx.copy_(X, non_blocking=True)
        g.replay()
        self._parity = (parity + len(self.layers)) % 2
        return y

    def _out_positions(self, dev):
        if getattr(self, '_pos', None) is None:
            self._pos = torch.from_numpy(self.layers[-1]._shard.position).to(dev)
        return self._pos

    def forward_host_many(self, batches, outs):
        """End to end over a sequence of pinned HOST batches [N, C, H, W] (the same batch on every rank): H2D, sensor
        encryption, the sharded chain, D2H of the N x K logits into outs[k] (pinned).  The H2D copy of batch k+1 runs on a
        copy stream while the chain of batch k runs.  Checks the homogeneous coordinate of every batch at the end."""
        from . import _native
        dev = torch.device('cuda', torch.cuda.current_device())
        if not hasattr(self, '_h2d'):
            self._h2d = dict(stream=torch.cuda.Stream(), stage=[None, None], done=[torch.cuda.Event() for _ in range(2)], free=[torch.cuda.Event() for _ in range(2)])
        st = self._h2d
        main = torch.cuda.current_stream()
        for b in range(2):
            st['free'][b].record(main)

        band = self._input_band()

        def h2d(k):
            b = k % 2
            with torch.cuda.stream(st['stream']):
                st['stream'].wait_event(st['free'][b])
                if st['stage'][b] is None or st['stage'][b].shape != batches[k].shape:
                    st['stage'][b] = torch.zeros(batches[k].shape, dtype=torch.float32, device=dev)
                if band is None or batches[k].ndim != 4:
                    st['stage'][b].copy_(batches[k], non_blocking=True)
                else:
                    # only the image rows this rank's first layer reads (its band of output pixels + halo) cross PCIe
                    (y0, y1) = band
                    st['stage'][b][:, :, y0:y1, :].copy_(batches[k][:, :, y0:y1, :], non_blocking=True)
                st['done'][b].record(st['stream'])
        if len(batches) > 0:
            h2d(0)
        bad = torch.zeros(1, dtype=torch.int32, device=dev)
        for k in range(len(batches)):
            b = k % 2
            if k + 1 < len(batches):
                h2d(k + 1)
            main.wait_event(st['done'][b])
            x = st['stage'][b]
            (N, D) = (int(x.shape[0]), int(np.prod(x.shape[1:])))
            graphed = getattr(self, '_graphs', None) is not None and self._graph_N == N
            X = self._gx if graphed else torch.empty((D + 1, N), dtype=torch.float32, device=dev)
            self.sensor.encrypt_into(x.reshape(N, D), X)
            st['free'][b].record(main)
            y = self.forward_graph(X) if graphed else self.forward_linear(X.t())      # [N, K+1]; .t() of a transposed view is free
            K = y.shape[1] - 1
            logits = torch.empty((N, K), dtype=torch.float32, device=dev)
            _native.check(_native.lib().kn_linear_to_affine_t(_native.ptr(y.t().contiguous()), N, N, K, _native.ptr(logits), 1e-3, _native.ptr(bad), _native.stream_ptr()))
            outs[k].copy_(logits, non_blocking=True)
        nbad = int(bad.cpu().item())
        if nbad != 0:
            raise ValueError('invalid affine vector: %d outputs lost the homogeneous coordinate' % nbad)
        return outs

    def _input_band(self):
        """Image rows [y0, y1) whose pixels this rank's first keyed layer reads (through the image key), or None when it reads
        (nearly) everything.  The sensor output rows outside the band are never read by this rank's shard."""
        if not hasattr(self, '_band'):
            self._band = None
            (A, _) = self.sensor.keypair()
            (C, H, W) = [int(v) for v in self.sensor._inshape[1:]]
            if self.world > 1 and hasattr(A, 'perm') and getattr(A, 'bias', None) is None and len(self.layers) > 0:
                used = self.layers[0].W.used_columns().cpu().numpy()
                rows = np.nonzero(used[:C * H * W])[0]                     # keyed rows read (the homogeneous row is local)
                if len(rows) > 0:
                    d = np.asarray(A.perm)[rows]                            # raw pixel behind every keyed row
                    y = (d % (H * W)) // W
                    (y0, y1) = (int(y.min()), int(y.max()) + 1)
                    if (y1 - y0) < 0.75 * H:
                        self._band = (y0, y1)
        return self._band

    def h2d_bytes(self, batch_shape):
        """Bytes forward_host_many copies to this rank's GPU for one host batch of this shape."""
        band = self._input_band()
        n = int(np.prod(batch_shape))
        if band is None or len(batch_shape) != 4:
            return 4 * n
        return 4 * n * (band[1] - band[0]) // int(batch_shape[2])

    def forward(self, x_cipher):
        from . import torch as ktorch
        y = self.forward_linear(x_cipher)
        N = y.shape[0]
        return ktorch.linear_to_affine(y, self._outshape if N == 1 else (N,) + tuple(self._outshape))
