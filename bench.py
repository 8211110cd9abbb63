#!/usr/bin/env python
"""bench.py -- encrypted images/s of the keyed-layer forward path on B200.

Workload (BASELINE.json configs[1]): CIFAR-10 AllConvNet 3x32x32, hierarchical block-permutation keys
(`np.random.seed(0); Keynet((3,32,32), net, global_geometric='hierarchical_permutation',
hierarchical_blockshape=(2,2), hierarchical_permute_at_level=(0,1))`), batch 4096 per GPU, synthetic
images, numpy-seeded random-init weights.  One step = sensor.encrypt() + knet.forward() over one batch:
1 encrypt kernel (homogenise + transpose + image key) + one or two SpMM launches per keyed layer (+ReLU fused) + 1 layout kernel.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--net acn|lenet|vgg16]
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...      (N > 1: data-parallel replicas)
  ... bench.py --gpus N --net vgg16 --batch 256 --parallel rows|rows-fused         (every keyed layer row-sharded: strong scaling)
  python bench.py --impl reference ...      (CPU arm: the oracle port of the reference's scipy path, all host threads)

Prints ONE JSON line (rank 0).  `value` = whole-job images/s with inputs resident in HBM; `e2e` = the same
through the host-buffer API (pinned H2D of the images and D2H of the logits inside the timed region);
`roofline` = the dominant SpMM launch against the measured HBM copy bandwidth (MEASURED_PEAKS.json);
`cpu_baseline` = the oracle port timed on the host cores on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np

HBM_FALLBACK_GBS = 6650.0     # /opt/skills/guides/B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent


def numpy_weights(net, seed):
    """Deterministic kaiming-uniform-like init from numpy's legacy RNG (same as tests/golden/make_golden.py)."""
    import torch
    rs = np.random.RandomState(seed)
    with torch.no_grad():
        for (name, p) in net.named_parameters():
            fan_in = int(np.prod(p.shape[1:])) if p.ndim > 1 else int(p.shape[0])
            bound = 1.0 / np.sqrt(max(1, fan_in))
            p.copy_(torch.from_numpy(rs.uniform(-bound, bound, size=tuple(p.shape)).astype(np.float32)))
    return net


def workload(name):
    from keynet_b200 import nets
    if name == 'acn':
        return dict(net=numpy_weights(nets.AllConvNet(batchnorm=False), 0).eval(), inshape=(3, 32, 32),
                    keys=dict(global_geometric='hierarchical_permutation', hierarchical_blockshape=(2, 2), hierarchical_permute_at_level=(0, 1)),
                    label='AllConvNet 3x32x32, hierarchical block-permutation keys (BASELINE configs[1])')
    if name == 'lenet':
        return dict(net=numpy_weights(nets.LeNet_AvgPool(), 0).eval(), inshape=(1, 28, 28), keys=dict(global_geometric='permutation'),
                    label='LeNet_AvgPool 1x28x28, PermutationKeynet (BASELINE configs[0])')
    if name == 'vgg16':
        return dict(net=numpy_weights(nets.VGG16(), 0).eval(), inshape=(3, 224, 224), keys=dict(global_geometric='permutation', keep_csr=False),
                    label='VGG16 3x224x224, PermutationKeynet, pattern groups with unique value blocks on one GPU (BASELINE configs[3])')
    raise ValueError(name)


def hbm_peak():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        try:
            with open(p) as f:
                return (float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json)')
        except Exception:
            pass
    return (HBM_FALLBACK_GBS, 'fallback (B200_PROFILING.md)')


def tensor_peak():
    """Dense bf16 tensor throughput measured on this pool's B200s (sustained figure: the kernel is timed inside a long step)."""
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        try:
            with open(p) as f:
                return (float(json.load(f)['bf16_tflops_sustained']), 'measured bf16 sustained (MEASURED_PEAKS.json)')
        except Exception:
            pass
    return (1400.0, 'fallback (B200_PROFILING.md, sustained)')


def profiled_traffic(net, batch, layer):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed ncu capture
    (profiles/traffic.json); None if no capture matches this workload."""
    p = os.path.join(ROOT, 'profiles', 'traffic.json')
    try:
        with open(p) as f:
            t = json.load(f)
        return t.get('%s:%d:%s' % (net, batch, layer))
    except Exception:
        return None


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled every 50 ms while the timed region runs."""
    Q = 'clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        self.gpu_index = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.gpu_index), '--query-gpu=' + self.Q, '--format=csv,noheader,nounits', '-lms', '50'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        if len(self.rows) < 2:
            # a very short timed region can end before the sampler's second reading: add one taken right now (GPU still loaded)
            try:
                r = subprocess.run(['nvidia-smi', '-i', str(self.gpu_index), '--query-gpu=' + self.Q, '--format=csv,noheader,nounits'],
                                   capture_output=True, text=True, timeout=10)
                self.rows.extend([l.strip() for l in r.stdout.splitlines() if l.strip()])
            except Exception:
                pass
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        (sm, smax, reasons) = ([], [], set())
        for r in self.rows:
            f = [t.strip() for t in r.split(',')]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); smax.append(float(f[1]))
            except ValueError:
                continue
            for (name, v) in zip(['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'], f[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(smax) if smax else None, 'reasons': sorted(reasons), 'samples': len(sm)}


# =================================================================================================
def oracle_layers_from_gpu(sensor, knet):
    """Copy the (bit-exact, verified) compiled matrices to the host for the CPU baseline."""
    from oracle import keynet_oracle as ko
    layers = []
    (ip, ix, dt) = sensor.W.csr_arrays()
    layers.append((ko.csr(sensor.W.shape, ip, ix, dt), False))
    for (k, L) in knet.keyedlayers():
        (ip, ix, dt) = L.W.csr_arrays()
        layers.append((ko.csr(L.W.shape, ip - ip[0], ix, dt), bool(L._fused_relu)))
    return layers


def oracle_layers_on_cpu(wl):
    """Reference arm: key the network on the CPU with the oracle (Toeplitz + two SpGEMMs per layer), no GPU."""
    from oracle import keynet_oracle as ko
    from keynet_b200 import system, torch as ktorch
    from torch import nn
    recorded = []

    def f_layergen(module, inshape, outshape, A, Ainv):
        k = lambda K: None if K is None else ko.monomial_key(K.perm, K.scale)
        if isinstance(module, nn.Conv2d):
            W = ko.toeplitz_conv2d(inshape, module.weight.detach().numpy(), module.bias.detach().numpy(), module.stride[0])
        elif isinstance(module, nn.AvgPool2d):
            ks = module.kernel_size if isinstance(module.kernel_size, int) else module.kernel_size[0]
            st = module.stride if isinstance(module.stride, int) else module.stride[0]
            W = ko.toeplitz_avgpool2d(inshape, ks, st)
        elif isinstance(module, nn.Linear):
            W = ko.linear_matrix(module.weight.detach().numpy(), module.bias.detach().numpy())
        else:
            raise ValueError(str(type(module)))
        What = ko.key_compile(k(A), W, k(Ainv))

        class Rec(nn.Module):       # stands in for a KeyedLayer inside KeyedModel's Sequential
            def fuse_relu(self, flag=True):
                self.relu = bool(flag)
                return self
        r = Rec(); r.W = What; r.relu = False
        recorded.append(r)
        return r
    np.random.seed(0)
    f_keypair = system.keypair_policy(**wl['keys'])
    (A, Ainv) = f_keypair('input', wl['inshape'])
    system.KeyedModel(wl['net'], wl['inshape'], Ainv, f_keypair, f_layergen)
    return [(ko.monomial_key(A.perm, A.scale), False)] + [(r.W, r.relu) for r in recorded]


def host_threads():
    """Host cores this process may use.  (torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU arm runs on rank 0 alone
    and is meant to use all the host threads it can, so the OpenMP default is not what we want there.)"""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def time_oracle(layers, inshape, n_images, threads, repeats=1):
    from oracle import keynet_oracle as ko
    rs = np.random.RandomState(0)
    x = ko.affine_to_linear(rs.randn(n_images, *inshape).astype(np.float32))
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        y = ko.keyed_forward(layers, x, threads=threads)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    ko.linear_to_affine(y)
    return best


def cpu_baseline(layers, inshape, budget_s=15.0):
    threads = host_threads()
    t_probe = time_oracle(layers, inshape, 32, threads)
    n = int(max(32, min(16384, (budget_s / max(t_probe / 32.0, 1e-6)))))
    dt = time_oracle(layers, inshape, n, threads)
    return {'value': n / dt, 'unit': 'images/s', 'cores': threads, 'kind': 'port',
            'sample': '%d images through the same compiled layer stack, oracle csr_matvecs port (OpenMP over rows), %.1f s' % (n, dt)}


def run_reference(args, rank, world):
    """CPU arm: rank 0 only; the oracle port keys and runs the same workload on the host cores."""
    if rank != 0:
        return
    from oracle import keynet_oracle as ko
    wl = workload(args.net)
    layers = oracle_layers_on_cpu(wl)
    threads = host_threads()
    t_probe = time_oracle(layers, wl['inshape'], 2, threads)
    n = int(max(1, min(64, 4.0 / max(t_probe / 2.0, 1e-6))))       # ~4 s of CPU work per step
    for _ in range(max(1, min(args.warmup, 1))):
        time_oracle(layers, wl['inshape'], n, threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        time_oracle(layers, wl['inshape'], n, threads)
    dt = time.perf_counter() - t0
    v = n * args.steps / dt
    out = {'impl': 'reference', 'metric': 'encrypted_images_per_sec', 'value': v, 'unit': 'images/s', 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
           'ms_per_step': 1e3 * dt / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
           'config': {'workload': wl['label'], 'images_per_step': n, 'note': 'CPU arm: oracle port of the reference scipy path; each step is a bounded sample of the workload'},
           'cpu_baseline': {'value': v, 'unit': 'images/s', 'cores': threads, 'kind': 'port', 'sample': '%d images per step, %d steps' % (n, args.steps)},
           'e2e': {'value': v, 'unit': 'images/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(out))


def run_row_sharded(args, rank, world, local_rank):
    """Strong scaling: ONE batch, every keyed layer's rows cut into `world` shards (keynet_b200/dist.py), activations
    all-gathered per layer over NVLink (NCCL, or fused into the SpMM epilogue with --parallel rows-fused)."""
    import torch
    import torch.distributed as dist
    from keynet_b200 import dist as kdist
    wl = workload(args.net)
    keys = {k: v for (k, v) in wl['keys'].items() if k != 'keep_csr'}
    N = args.batch
    t0 = time.perf_counter()
    np.random.seed(0)
    m = kdist.ShardedKeyedModel(wl['inshape'], wl['net'], rank=rank, world=world, fused=(args.parallel == 'rows-fused'), keep_csr=False, **keys)
    torch.cuda.synchronize()
    t_compile = time.perf_counter() - t0
    x = torch.randn((N,) + wl['inshape'], generator=torch.Generator().manual_seed(1)).cuda()      # same batch on every rank
    K = args.steps

    def step():
        return m.forward_linear(m.sensor.fromtensor(x).encrypt().astensor())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    for _ in range(args.warmup):
        y = step()
    barrier()
    sampler = ClockSampler(local_rank).start() if rank == 0 else None
    (e0, e1) = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
    e0.record()
    for _ in range(K):
        y = step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if sampler is not None else None
    layer_ms = None
    if True:                                      # one extra (untimed) forward with per-layer events
        m.time_layers = True
        step()
        layer_ms = [(type(m._model).__name__ and k, round(a, 3), round(b, 3)) for (k, a, b) in m.layer_times_ms()]
        m.time_layers = False
    # correctness of the sharded forward, outside the timed region: the decrypted-by-construction logits of the first images
    # against the plain torch network on the host
    err = None
    if rank == 0:
        import copy
        plain = copy.deepcopy(wl['net'])
        for (k, mod) in list(plain.named_children()):      # the pooling the reference actually keys: centred k x k windows,
            if isinstance(mod, torch.nn.AvgPool2d):        # divisor k*k (keynet/layer.py:48-56 ignores padding / ceil_mode)
                ks = mod.kernel_size if isinstance(mod.kernel_size, int) else mod.kernel_size[0]
                st = mod.stride if isinstance(mod.stride, int) else mod.stride[0]
                setattr(plain, k, torch.nn.AvgPool2d(ks, st, ks // 2, ceil_mode=False, count_include_pad=True))
        with torch.no_grad():
            yp = plain(x[:4].cpu()).numpy()
        yk = y[:4, :-1].cpu().numpy()
        err = {'max_abs_err_vs_plain_net': float(np.abs(yk - yp).max()), 'max_abs_logit': float(np.abs(yp).max()), 'argmax_equal': bool(np.array_equal(yk.argmax(1), yp.argmax(1)))}
    nnz_local = m.num_parameters_local()
    stats = torch.tensor([ms, float(nnz_local), torch.cuda.max_memory_allocated() / 1e9], device='cuda', dtype=torch.float64)
    if world > 1:
        mx = stats.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stats.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
    else:
        (mx, sm) = (stats, stats)
    ms = float(mx[0])
    if rank == 0:
        (peak, peak_src) = hbm_peak()
        nnz = float(sm[1])
        act = sum((L.W.shape[0] * 0 + L._shard.n_rows + L.W.shape[1]) * N * 4 for L in m.layers)          # every rank reads X_full; rows written once
        alg = nnz * 8 + act
        gather = sum((L._shard.n_phys - 1) * N * 4 for L in m.layers)
        hbm = alg / (ms / K * 1e-3) / 1e9
        out = {'metric': 'encrypted_images_per_sec', 'value': N * K / (ms * 1e-3), 'unit': 'images/s', 'n_gpus': world, 'steps': K, 'warmup': args.warmup,
               'ms_per_step': ms / K, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
               'config': {'workload': wl['label'], 'global_batch': N, 'parallelism': 'rows x%d, %s' % (world, 'fused SpMM+all-gather (NVLink peer stores)' if m.fused else 'NCCL all-gather per layer'),
                          'nnz': int(nnz), 'nnz_max_rank': int(mx[1]), 'key_compile_s': round(t_compile, 3), 'hbm_allocated_gb_max_rank': round(float(mx[2]), 2),
                          'all_gather_bytes_per_step': int(gather), 'peer_store_fraction': m.peer_store_fraction() if m.fused else None,
                          'rank0_layer_ms_spmm_barrier': layer_ms, 'check': err, 'l2': 'inputs larger than L2'},
               'gpu_launches': (len(m.layers) + 2) * K, 'clocks': clocks,
               'roofline': {'bound': 'hbm', 'kernel': 'whole network (CSR-equivalent algorithmic bytes: 8 B/nnz + activations)', 'achieved': hbm, 'peak': peak * world, 'unit': 'GB/s',
                            'frac': hbm / (peak * world), 'traffic': None, 'peak_source': peak_src + ' x n_gpus'}}
        print(json.dumps(out))
    if dist.is_initialized():
        dist.barrier()
        dist.destroy_process_group()


# =================================================================================================
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--batch', type=int, default=4096, help='images per GPU per step')
    ap.add_argument('--net', default='acn', choices=['acn', 'lenet', 'vgg16'])
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--parallel', default='dp', choices=['dp', 'rows', 'rows-fused'],
                    help='N > 1: dp = replicas (default); rows = every keyed layer row-sharded + NCCL all-gather; rows-fused = SpMM epilogue stores to NVLink peers')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'b200' else args.warmup

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if args.impl == 'reference':
        return run_reference(args, rank, world)

    import torch
    import torch.distributed as dist
    from keynet_b200 import system, engine, _native
    assert torch.cuda.is_available(), 'bench.py needs a GPU (no CPU fallback)'
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    _native.lib()
    if args.parallel != 'dp':
        if world == 1 and not dist.is_initialized():        # single-rank run of the sharded code path (layout / epilogue A-B)
            os.environ.setdefault('MASTER_ADDR', '127.0.0.1'); os.environ.setdefault('MASTER_PORT', '29533')
            dist.init_process_group('nccl', rank=0, world_size=1, device_id=torch.device('cuda', local_rank))
        return run_row_sharded(args, rank, world, local_rank)

    wl = workload(args.net)
    N = args.batch
    t0 = time.perf_counter()
    np.random.seed(0)
    (sensor, knet) = system.Keynet(wl['inshape'], wl['net'], **wl['keys'])
    torch.cuda.synchronize()
    t_compile = time.perf_counter() - t0

    K = args.steps
    plan = engine.ForwardPlan(sensor, knet, N, use_graph=False, time_layers=True, event_sets=K)
    g = torch.Generator(device='cuda').manual_seed(rank)
    images = torch.randn((N,) + wl['inshape'], device='cuda', generator=g)      # synthetic, resident in HBM
    plan.images.copy_(images.reshape(N, -1))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput ---------------------------------------------------------
    for _ in range(args.warmup):
        plan.run_device()
    barrier()
    sampler = ClockSampler(local_rank).start() if rank == 0 else None
    (e0, e1) = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
    e0.record()
    for k in range(K):
        plan.event_set = k
        plan.run_device()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = plan.launches_per_run * K

    # ---- end to end with host buffers -------------------------------------------------------
    # K batches through the public host-buffer API: every step copies its own images from pinned host memory and reads its
    # logits back; the copy of step k+1 overlaps the chain of step k (ForwardPlan.run_host_many)
    host_in = [torch.randn((N,) + wl['inshape']).pin_memory() for _ in range(2)]
    host_out = [torch.empty((N, plan.K), dtype=torch.float32).pin_memory() for _ in range(2)]
    plan.time_layers = False
    plan.run_host_many([host_in[k % 2] for k in range(2)], [host_out[k % 2] for k in range(2)])
    barrier()
    t0 = time.perf_counter()
    e0.record()
    plan.run_host_many([host_in[k % 2] for k in range(K)], [host_out[k % 2] for k in range(K)])
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    clocks = sampler.stop() if sampler is not None else None

    if world > 1:
        t = torch.tensor([ms, ms_e2e], device='cuda', dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        (ms, ms_e2e) = (float(t[0]), float(t[1]))

    if rank == 0:
        (peak, peak_src) = hbm_peak()
        (tpeak, tpeak_src) = tensor_peak()
        # per-layer launch durations measured inside the timed region (CUDA events on the launch stream)
        per_layer = plan.layer_times_ms_mean(K)
        alg = dict(plan.algorithmic_bytes())
        flops = {name: 2.0 * W.nnz() * N for (name, W, _) in plan.layers}
        (dom, dom_ms) = max(per_layer, key=lambda kv: kv[1])
        domW = [W for (name, W, _) in plan.layers if name == dom][0]
        grouped = domW._pg is not None and N >= 32 and N % 4 == 0
        on_tc = grouped and any(c['tc'] is not None for c in domW._pg.classes) and N >= 128
        clustered = grouped and any(c.get('cg') is not None for c in domW._pg.classes)
        kname = ('pg_tc_kernel (tcgen05 3xTF32)' if on_tc else ('pg_cluster_kernel (fp32 FMA, staged gathers)' if clustered else 'pg_simt_kernel / pg_small_kernel (fp32 FMA)')) if grouped else 'spmm_rowwarp_kernel'
        if grouped:
            kname += ', pattern groups (G, K_pad, n_groups)=%s' % str(domW._pg.summary()['classes'])
        hbm_achieved = alg[dom] / (dom_ms * 1e-3) / 1e9
        total_alg = sum(alg.values())
        traffic = profiled_traffic(args.net, N, dom)
        if on_tc:
            # this launch is bound by the tensor pipe, not HBM: arithmetic intensity 2*nnz*N / bytes is far above the
            # machine balance at batch 4096.  achieved = ALGORITHMIC flops (2*nnz*N) / launch time; peak = measured dense
            # bf16 (the only measured tensor number).  kind::tf32 runs at half the bf16 rate and the 3xTF32 split issues
            # 3 MMAs per product, so the ceiling for this arithmetic is peak/6.
            achieved = flops[dom] / (dom_ms * 1e-3) / 1e12
            roofline = {'bound': 'tensor', 'kernel': '%s on layer %s' % (kname, dom), 'achieved': achieved, 'peak': tpeak, 'unit': 'TFLOP/s', 'frac': achieved / tpeak,
                        'traffic': traffic, 'peak_source': tpeak_src, 'ceiling_3xtf32': tpeak / 6.0, 'frac_of_3xtf32_ceiling': achieved / (tpeak / 6.0),
                        'hbm': {'achieved': hbm_achieved, 'peak': peak, 'unit': 'GB/s', 'frac': hbm_achieved / peak, 'peak_source': peak_src}}
        else:
            roofline = {'bound': 'hbm', 'kernel': '%s on layer %s' % (kname, dom), 'achieved': hbm_achieved, 'peak': peak, 'unit': 'GB/s', 'frac': hbm_achieved / peak,
                        'traffic': traffic, 'peak_source': peak_src}
        roofline.update({'launch_ms': dom_ms, 'algorithmic_bytes_per_launch': alg[dom], 'algorithmic_flops_per_launch': flops[dom],
                         'share_of_step': dom_ms / (ms / K),
                         'network': {'algorithmic_bytes_per_step': total_alg, 'hbm_achieved_gbs': total_alg / (ms / K * 1e-3) / 1e9, 'hbm_frac': total_alg / (ms / K * 1e-3) / 1e9 / peak,
                                     'algorithmic_tflops': sum(flops.values()) / (ms / K * 1e-3) / 1e12},
                         'layers_ms': {k: round(v, 4) for (k, v) in per_layer}})
        out = {'metric': 'encrypted_images_per_sec', 'value': world * N * K / (ms * 1e-3), 'unit': 'images/s', 'n_gpus': world, 'steps': K, 'warmup': args.warmup,
               'ms_per_step': ms / K, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
               'config': {'workload': wl['label'], 'batch_per_gpu': N, 'global_batch': N * world, 'parallelism': 'dp%d replicas, no collective' % world,
                          'l2': 'inputs larger than L2: %.2f GB of CSR + %.2f GB of activations per step' % (sum(L[1].nnz() for L in plan.layers) * 8 / 1e9, sum((L[1].shape[0] + L[1].shape[1]) * N * 4 for L in plan.layers) / 1e9),
                          'nnz': int(sum(L[1].nnz() for L in plan.layers)), 'key_compile_s': round(t_compile, 3), 'hbm_allocated_gb': round(torch.cuda.max_memory_allocated() / 1e9, 2)},
               'e2e': {'value': world * N * K / (ms_e2e * 1e-3), 'unit': 'images/s', 'h2d_bytes_per_step': int(host_in[0].numel() * 4), 'd2h_bytes_per_step': int(host_out[0].numel() * 4),
                       'note': 'ForwardPlan.run_host_many: pinned H2D of step k+1 overlaps the chain of step k; one homogeneous-coordinate check (4 B D2H) per call',
                       'ms_per_step': ms_e2e / K},
               'gpu_launches': launches, 'clocks': clocks, 'roofline': roofline}
        if world == 1 and not args.no_cpu_baseline and args.net != 'vgg16':     # the oracle cannot hold 120 GB of CSR on the host
            out['cpu_baseline'] = cpu_baseline(oracle_layers_from_gpu(sensor, knet), wl['inshape'])
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
