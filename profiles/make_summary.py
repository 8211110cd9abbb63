#!/usr/bin/env python
"""Turn the scratch ncu outputs in gpurun_out/ into the small, tracked summaries under profiles/.

  python profiles/make_summary.py <round tag> <launches.csv> <full.ncu-rep> <net:batch:layer>
"""
import collections
import csv
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
(tag, launches_csv, rep, traffic_key) = sys.argv[1:5]

# ---- launch list: every launch with its device time, aggregated per kernel (shares, not absolutes) ----
rows = list(csv.reader(open(launches_csv)))
hi = [i for (i, r) in enumerate(rows) if 'Kernel Name' in r][0]
hdr = rows[hi]
(ki, vi, gi) = (hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Grid Size'))
agg = collections.OrderedDict()
per_launch = []
for r in rows[hi + 1:]:
    if len(r) <= vi:
        continue
    name = r[ki].split('(')[0].replace('void <unnamed>::', '').replace('<unnamed>::', '')
    t = float(r[vi].replace(',', ''))
    per_launch.append((name, r[gi], t))
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += t
FORWARD = ('pg_tc_kernel', 'pg_simt_kernel', 'pg_small_kernel', 'pg_cluster_kernel', 'spmm_', 'affine_to_linear_t', 'linear_to_affine_t', 'encrypt_monomial_t', 'splitk_reduce')
fwd = collections.OrderedDict((k, v) for (k, v) in agg.items() if k.startswith(FORWARD))
build = collections.OrderedDict((k, v) for (k, v) in agg.items() if not k.startswith(FORWARD))
tot = sum(v[1] for v in fwd.values())
with open(os.path.join(HERE, '%s_launches.csv' % tag), 'w') as f:
    f.write('# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)\n')
    f.write('# forward step kernels (sensor.encrypt + knet.forward), share of the summed forward kernel time\n')
    f.write('kernel,launches,total_us,share\n')
    for (k, (n, t)) in sorted(fwd.items(), key=lambda kv: -kv[1][1]):
        f.write('%s,%d,%.1f,%.4f\n' % (k, n, t / 1e3, t / tot))
    f.write('\n# one-off key-compile kernels of the same process (Toeplitz build, A.W.Ainv, pattern grouping), not part of the step\n')
    f.write('kernel,launches,total_us\n')
    for (k, (n, t)) in sorted(build.items(), key=lambda kv: -kv[1][1]):
        f.write('%s,%d,%.1f\n' % (k, n, t / 1e3))
    f.write('\n# individual launches of the last profiled step (kernel,grid,us)\n')
    last = [p for p in per_launch if p[0].startswith(FORWARD)]
    for (name, grid, t) in last[-max(1, len(last) // 9):]:
        f.write('%s,"%s",%.1f\n' % (name, grid, t / 1e3))

# ---- full capture of the dominant kernel ----
out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rr = list(csv.reader(out.splitlines()))
(h, u) = (rr[0], rr[1])
ti = h.index('gpu__time_duration.sum')
v = max(rr[2:], key=lambda r: float(r[ti]))          # the capture may hold several launches: summarise the longest
want = ['Kernel Name', 'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed', 'sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__pcsamp_warps_issue_stalled_long_scoreboard', 'smsp__pcsamp_warps_issue_stalled_barrier', 'smsp__pcsamp_warps_issue_stalled_wait',
        'smsp__pcsamp_warps_issue_stalled_short_scoreboard', 'smsp__pcsamp_warps_issue_stalled_selected', 'smsp__pcsamp_warps_issue_stalled_branch_resolving']
with open(os.path.join(HERE, '%s_dominant_kernel_ncu.txt' % tag), 'w') as f:
    f.write('# ncu --set full --clock-control none --import-source on (longest of %d captured launches; %s)\n' % (len(rr) - 2, os.path.basename(rep)))
    for k in want:
        if k in h:
            i = h.index(k)
            f.write('%-80s %s %s\n' % (k, v[i], u[i]))
    traffic = None
    try:
        (ri, wi) = (h.index('dram__bytes_read.sum'), h.index('dram__bytes_write.sum'))
        scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
        traffic = float(v[ri]) * scale[u[ri]] + float(v[wi]) * scale[u[wi]]
        f.write('%-80s %.0f byte\n' % ('traffic = dram__bytes_read.sum + dram__bytes_write.sum', traffic))
    except Exception as e:
        f.write('traffic unavailable: %s\n' % e)
tj = os.path.join(HERE, 'traffic.json')
d = json.load(open(tj)) if os.path.exists(tj) else {}
if traffic is not None:
    d[traffic_key] = traffic
json.dump(d, open(tj, 'w'), indent=1)
print('wrote profiles/%s_*' % tag, traffic)
