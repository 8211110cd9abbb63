"""Multi-process (world_size 2, gloo, CPU) tests of the row-sharding plumbing of keynet_b200/dist.py: the shard
planner (shard-major row order, gathered positions, column remap) and the all-gather data flow.  The local product is
done by the CPU oracle here (the CUDA kernels need a GPU; `-m gpu` covers them); what is checked is that sharded ==
unsharded (to fp32 rounding: the shard-major layout changes the order of the terms inside a row)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from torch import nn

from keynet_b200 import dist as kdist
from keynet_b200 import system, nets
from keynet_b200.sparse import MonomialKey
from oracle import keynet_oracle as ko


def _free_port():
    s = socket.socket(); s.bind(('127.0.0.1', 0)); p = s.getsockname()[1]; s.close()
    return p


def _oracle_net(seed=0):
    """Tiny conv-relu-pool-fc network keyed with permutation + gain keys by the ORACLE (full matrices on the host)."""
    torch.manual_seed(seed)

    class Tiny(nn.Module):
        def __init__(self):
            super(Tiny, self).__init__()
            self.conv1 = nn.Conv2d(2, 4, 3, padding=1); self.relu1 = nn.ReLU()
            self.pool1 = nn.AvgPool2d(3, 2, 1)
            self.conv2 = nn.Conv2d(4, 6, 3, padding=1); self.relu2 = nn.ReLU()
            self.fc1 = nn.Linear(6 * 3 * 3, 7)

        def forward(self, x):
            x = self.relu2(self.conv2(self.pool1(self.relu1(self.conv1(x)))))
            return self.fc1(x.reshape(x.shape[0], -1))
    net = Tiny().eval()
    inshape = (2, 6, 6)
    layers = []       # (module, outshape, W_hat canonical CSR, relu)

    def f_layergen(module, ishape, oshape, A, Ainv):
        k = lambda K: None if K is None else ko.monomial_key(K.perm, K.scale)
        if isinstance(module, nn.Conv2d):
            W = ko.toeplitz_conv2d(ishape, module.weight.detach().numpy(), module.bias.detach().numpy(), module.stride[0])
        elif isinstance(module, nn.Linear):
            W = ko.linear_matrix(module.weight.detach().numpy(), module.bias.detach().numpy())
        else:
            W = ko.toeplitz_avgpool2d(ishape, 3, 2)
        What = ko.sort_indices(ko.key_compile(k(A), W, k(Ainv)))

        class Rec(nn.Module):
            def fuse_relu(self, flag=True):
                self.relu = bool(flag); return self
        r = Rec(); r.W = What; r.relu = False; r.module = module; r.outshape = oshape; r.A = A
        layers.append(r)
        return r
    np.random.seed(seed)
    f_keypair = system.keypair_policy(global_geometric='permutation', global_photometric='uniform_random_gain', beta=1.0)
    (A, Ainv) = f_keypair('input', inshape)
    system.KeyedModel(net, inshape, Ainv, f_keypair, f_layergen)
    return (inshape, ko.monomial_key(A.perm, A.scale), layers)


def _shard_csr(W, my_rows, position_prev, n_phys_prev):
    """Numpy restatement of what the GPU compile does for a shard: select rows, remap columns to gathered positions."""
    rows = []
    (ip, ix, dt) = (W.indptr, W.indices, W.data)
    (nip, nix, ndt) = ([0], [], [])
    for r in my_rows:
        c = ix[ip[r]:ip[r + 1]]; v = dt[ip[r]:ip[r + 1]]
        c = c if position_prev is None else position_prev[c]
        o = np.argsort(c, kind='stable')
        nix.append(c[o]); ndt.append(v[o]); nip.append(nip[-1] + len(c))
    return ko.csr((len(my_rows), n_phys_prev), np.array(nip), np.concatenate(nix) if nix else np.zeros(0), np.concatenate(ndt) if ndt else np.zeros(0))


def _worker(rank, world, port, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'; os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        (inshape, Akey, layers) = _oracle_net()
        rs = np.random.RandomState(1)
        x = ko.affine_to_linear(rs.randn(5, *inshape).astype(np.float32))
        X = ko.spmm(Akey, np.ascontiguousarray(x.T))                     # sensor encrypt, replicated
        # unsharded reference
        Xr = X
        for L in layers:
            Xr = ko.spmm(L.W, Xr, relu=L.relu)
        # sharded: compile every layer's shard first (rows cut by the pixel of the underlying Toeplitz row, the previous
        # layer's gathered layout folded into the columns), then derive who needs which gathered row
        (pos_prev, n_prev) = (None, X.shape[0])
        (shards, Wls) = ([], [])
        for L in layers:
            sh = kdist.LayerShard(L.module, L.outshape, rank, world, L.A)
            Wls.append(_shard_csr(L.W, sh.my_rows, pos_prev, n_prev))
            shards.append(sh)
            (pos_prev, n_prev) = (sh.position, sh.n_phys)
        sent = full = 0
        for (k, (L, sh, Wl)) in enumerate(zip(layers, shards, Wls)):
            Yloc = np.zeros((sh.chunk, X.shape[1]), dtype=np.float32)
            if len(sh.my_rows):
                Yloc[:len(sh.my_rows)] = ko.spmm(Wl, X, relu=L.relu)
            Yfull = torch.zeros((sh.n_phys, X.shape[1]), dtype=torch.float32)
            dist.all_gather_into_tensor(Yfull[:world * sh.chunk], torch.from_numpy(Yloc))
            Yfull[-1] = 1.0
            X = Yfull.numpy()
            if k + 1 < len(layers):
                # selective peer stores (kn_peers.row_mask): a rank only ever receives the rows its next layer reads.
                # Emulated by poisoning everything else after the gather -- the result must not change.
                need = np.zeros(sh.n_phys, dtype=bool)
                need[Wls[k + 1].indices] = True
                allneed = [None] * world
                dist.all_gather_object(allneed, need)
                allneed = np.stack(allneed)
                mask = kdist.peer_row_masks(allneed, rank, sh.chunk, len(sh.my_rows))
                for peer in range(world):
                    want = allneed[peer, rank * sh.chunk:rank * sh.chunk + len(sh.my_rows)] | (peer == rank)
                    assert np.array_equal((mask >> peer) & 1, want.astype(np.uint8))
                sent += int(np.unpackbits(mask.reshape(-1, 1), axis=1).sum()); full += len(mask) * world
                own = np.zeros(sh.n_phys, dtype=bool); own[rank * sh.chunk:(rank + 1) * sh.chunk] = True; own[-1] = True
                X[~(need | own)] = np.nan
        Y = X[pos_prev]
        # the gathered layout re-orders the columns inside a row, i.e. the fp32 summation order: equal to rounding
        q.put((rank, bool(np.allclose(Y, Xr, rtol=1e-5, atol=1e-6)), float(np.abs(Y - Xr).max())))
    finally:
        dist.destroy_process_group()


def test_row_sharded_forward_equals_unsharded_world2():
    world = 2
    port = _free_port()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=60) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for (rank, exact, err) in res:
        assert exact, (rank, err)


def test_shard_plan_properties():
    for (module, outshape) in [(nn.Conv2d(3, 8, 3, padding=1), (8, 5, 7)), (nn.AvgPool2d(3, 2, 1), (4, 6, 6)), (nn.Linear(10, 13), (13, 1, 1)),
                               (nn.Conv2d(3, 8, 3, padding=1), (8, 1, 1))]:
        for (world, keyed) in [(1, False), (2, False), (2, True), (3, True), (8, False), (8, True)]:
            R = int(np.prod(outshape))
            A = MonomialKey(np.concatenate([np.random.RandomState(world).permutation(R), [R]])) if keyed else None
            shards = [kdist.LayerShard(module, outshape, r, world, A) for r in range(world)]
            got = np.concatenate([s.my_rows for s in shards])
            assert np.array_equal(np.sort(got), np.arange(R))                 # every canonical row owned exactly once
            for (r, s) in enumerate(shards):
                assert len(s.my_rows) <= s.chunk and s.n_phys == world * s.chunk + 1
                assert np.array_equal(s.position[s.my_rows], r * s.chunk + np.arange(len(s.my_rows)))
                assert s.position[R] == world * s.chunk
            if isinstance(module, (nn.Conv2d, nn.AvgPool2d)) and outshape[1] * outshape[2] > 1:
                # spatial shards hold whole pattern groups: all output channels of each owned pixel of the underlying
                # Toeplitz matrix, and the pixels of rank r precede those of rank r+1 in raster order
                last = -1
                for s in shards:
                    px = (s.my_rows if A is None else A.perm[s.my_rows]) % (outshape[1] * outshape[2])
                    if len(px):
                        assert px.min() > last
                        last = px.max()
                    assert len(s.my_rows) % outshape[0] == 0
                    assert all(np.sum(px == p) == outshape[0] for p in np.unique(px))


def test_sync_sets_neighbourhood_and_war():
    """kn_peer_sync plan: a rank waits for the ranks it reads from (RAW) and for the ranks it will store to in the next layer
    (WAR on the ping-pong buffer), and signals exactly the ranks that wait for it."""
    from keynet_b200.dist import sync_sets
    W = 4
    line = np.zeros((W, W), dtype=bool)
    for r in range(W):
        for q in (r - 1, r, r + 1):
            if 0 <= q < W:
                line[r, q] = True                          # conv layer cut by image rows: a rank reads itself and its neighbours
    everyone = np.ones((W, W), dtype=bool)
    for r in range(W):
        (sig, wait) = sync_sets(line, line, r)
        nb = sum(1 << q for q in (r - 1, r + 1) if 0 <= q < W)
        assert sig == nb and wait == nb
    # next layer dense (fc6 reads everything): every rank will store to every rank -> waits for all (WAR), signals all
    for r in range(W):
        (sig, wait) = sync_sets(line, everyone, r)
        assert wait == ((1 << W) - 1) & ~(1 << r) and sig == wait
    # last layer: everyone reads the logits
    (sig, wait) = sync_sets(everyone, None, 2)
    assert sig == wait == 0b1011
    # consistency: q is in r's wait set  <=>  r is in q's signal set
    rs = np.random.RandomState(0)
    (a, b) = (rs.rand(W, W) < 0.4, rs.rand(W, W) < 0.4)
    sets = [sync_sets(a, b, r) for r in range(W)]
    for r in range(W):
        for q in range(W):
            if r != q:
                assert bool((sets[r][1] >> q) & 1) == bool((sets[q][0] >> r) & 1)
