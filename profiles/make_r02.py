"""Summaries of the round-2 ncu captures / bench lines kept under profiles/ (inputs: gpurun_out/final/*, untracked scratch).
  python profiles/make_r02.py"""
import collections, csv, hashlib, json, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, 'gpurun_out', 'final')
OUT = os.path.join(ROOT, 'profiles')


def raw(rep):
    txt = subprocess.run(['ncu', '-i', os.path.join(SRC, rep), '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    return {h: (u, v) for (h, u, v) in zip(rows[0], rows[1], rows[2])}


def fnum(d, k):
    return float(d[k][1].replace(',', ''))


def full_summary(rep, title, alg_bytes=None, alg_flops=None):
    d = raw(rep)
    ms = fnum(d, 'gpu__time_duration.sum') / (1e6 if d['gpu__time_duration.sum'][0] in ('ns', 'nsecond') else (1e3 if d['gpu__time_duration.sum'][0] in ('us', 'usecond') else 1))
    unit = lambda k: {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1}[d[k][0]]
    rd = fnum(d, 'dram__bytes_read.sum') * unit('dram__bytes_read.sum'); wr = fnum(d, 'dram__bytes_write.sum') * unit('dram__bytes_write.sum')
    lines = [title, 'kernel: ' + d['Kernel Name'][1][:160], 'grid %s block %s, registers/thread %s, dynamic smem %s %s' % (d['Grid Size'][1], d['Block Size'][1], d['launch__registers_per_thread'][1], d.get('launch__shared_mem_per_block_dynamic', ('', '?'))[1], d.get('launch__shared_mem_per_block_dynamic', ('', ''))[0]),
             'duration %.3f ms (under ncu: cold caches, serialised)' % ms,
             'DRAM read %.3f GB + write %.3f GB = %.3f GB  (%.0f GB/s)' % (rd / 1e9, wr / 1e9, (rd + wr) / 1e9, (rd + wr) / ms / 1e6)]
    if alg_bytes:
        lines.append('algorithmic bytes %.3f GB -> traffic / algorithmic = %.2f' % (alg_bytes / 1e9, (rd + wr) / alg_bytes))
    if alg_flops:
        lines.append('algorithmic flops %.3f T -> %.1f TFLOP/s' % (alg_flops / 1e12, alg_flops / ms / 1e9))
    for k in ['sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
              'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'sm__warps_active.avg.pct_of_peak_sustained_active',
              'smsp__issue_active.avg.pct_of_peak_sustained_active']:
        hit = [h for h in d if h.endswith(k)]
        if hit:
            lines.append('%-80s %s %s' % (k, d[hit[0]][1], d[hit[0]][0]))
    st = []
    for h in d:
        if 'pcsamp_warps_issue_stalled' in h and 'not_issued' not in h:
            try:
                st.append((float(d[h][1].replace(',', '')), h.replace('smsp__pcsamp_warps_issue_stalled_', '')))
            except ValueError:
                pass
    lines.append('warp stall samples (top): ' + ', '.join('%s %d' % (n, v) for (v, n) in sorted(st, reverse=True)[:7]))
    return ('\n'.join(lines) + '\n', rd + wr)


def launches(csvfile, title, top=25):
    rows = list(csv.reader(open(os.path.join(SRC, csvfile))))
    for (i, r) in enumerate(rows):
        if 'Kernel Name' in r:
            (h, st) = (r, i)
            break
    (ik, iv, iu) = (h.index('Kernel Name'), h.index('Metric Value'), h.index('Metric Unit'))
    agg = collections.defaultdict(list)
    for r in rows[st + 1:]:
        if len(r) > iv:
            v = float(r[iv].replace(',', ''))
            ms = v / 1e6 if r[iu] in ('ns', 'nsecond') else (v / 1e3 if r[iu] in ('us', 'usecond') else v)
            agg[r[ik].split('(')[0].replace('void ', '').replace('<unnamed>::', '')].append(ms)
    tot = sum(sum(v) for v in agg.values())
    lines = [title, '%-70s %6s %10s %9s %7s' % ('kernel', 'n', 'total ms', 'max ms', 'share')]
    for (k, v) in sorted(agg.items(), key=lambda kv: -sum(kv[1]))[:top]:
        lines.append('%-70s %6d %10.3f %9.3f %6.1f%%' % (k[:70], len(v), sum(v), max(v), 100 * sum(v) / tot))
    lines.append('%-70s %6d %10.3f' % ('all kernels in the capture', sum(len(v) for v in agg.values()), tot))
    return '\n'.join(lines) + '\n'


if __name__ == '__main__':
    b = json.load(open(os.path.join(SRC, 'bench_1gpu.json')))
    rf = b['roofline']
    (txt, traffic) = full_summary('vgg16_conv1_2_tile.ncu-rep', 'ncu --set full, dominant launch of the default bench.py run: VGG16 conv1_2 (64 -> 64 channels at 224x224, batch 256) on pg_tile_tc_kernel',
                                  alg_bytes=rf['algorithmic_bytes_per_launch'], alg_flops=rf['algorithmic_flops_per_launch'])
    open(os.path.join(OUT, 'r02_vgg16_dominant_kernel_ncu.txt'), 'w').write(txt)
    (txt2, _) = full_summary('keycompile_rows.ncu-rep', 'ncu --set full, keyed_conv_rows_kernel: fused key compile (canonical CSR) of a VGG16 conv1_2-sized layer (1.84 G entries, identity keys)',
                             alg_bytes=1841905665 * 8 + 3211266 * 8)
    open(os.path.join(OUT, 'r02_keycompile_fill_ncu.txt'), 'w').write(txt2)
    open(os.path.join(OUT, 'r02_vgg16_launches.txt'), 'w').write(launches('vgg16_launches.csv', 'ncu --metrics gpu__time_duration.sum --clock-control none python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra  (VGG16 batch 256: key compile + 6 forward passes)'))
    sys.path.insert(0, ROOT)
    from bench import lib_sha16
    t = {'vgg16:256:conv1_2': {'bytes': traffic, 'lib_sha16': lib_sha16(), 'capture': 'profiles/r02_vgg16_dominant_kernel_ncu.txt'}}
    json.dump(t, open(os.path.join(OUT, 'traffic.json'), 'w'), indent=1)
    for (src, dst) in [('bench_1gpu.json', 'r02_bench_1gpu.json'), ('bench_reference.json', 'r02_bench_reference_arm.json'), ('keycompile.json', 'r02_keycompile_sweep.json')]:
        shutil.copy(os.path.join(SRC, src), os.path.join(OUT, dst))
    print(open(os.path.join(OUT, 'r02_vgg16_dominant_kernel_ncu.txt')).read())
    print(open(os.path.join(OUT, 'r02_keycompile_fill_ncu.txt')).read())
    print(open(os.path.join(OUT, 'r02_vgg16_launches.txt')).read())
