"""SASS mnemonic counts per kernel of the built library -> profiles/r02_sass_evidence.txt   (python profiles/make_sass.py)"""
import collections, os, re, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'keynet_b200', 'lib', 'libkeynet_b200.so')
MN = ['UTCHMMA', 'UTCBAR', 'UTMALDG', 'UBLKCP', 'STTM', 'LDTM', 'LDGSTS', 'SYNCS', 'UTCATOMSWS', 'ELECT', 'R2UR']
txt = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True).stdout
counts = collections.OrderedDict()
cur = None
for line in txt.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    for k in MN:
        if re.search(r'\b%s\b' % k, line.split('/*')[1] if line.count('/*') >= 2 else line):
            counts[cur][k] += 1


def demangled(n):
    out = subprocess.run(['cu++filt', n], capture_output=True, text=True).stdout.strip()
    out = re.sub(r'\(anonymous namespace\)::|<unnamed>::|\((int|bool)\)', '', out)
    return re.sub(r'\(.*', '', out).replace('void ', '')


lines = ['SASS evidence for libkeynet_b200.so (sm_100a), `cuobjdump -sass keynet_b200/lib/libkeynet_b200.so`, mnemonics counted per kernel (python profiles/make_sass.py).',
         '',
         'UTCHMMA = tcgen05.mma, UTCBAR = tcgen05.commit, UTMALDG = TMA tensor load (cp.async.bulk.tensor), UBLKCP = cp.async.bulk (rows of the tiled',
         "kernel's epilogue), STTM / LDTM = tcgen05.st / .ld (TMEM), LDGSTS = cp.async, SYNCS = mbarrier ops, UTCATOMSWS = tcgen05.alloc, ELECT = elect.sync.",
         'R2UR / ELECT: under `if (lane == 0)` every UTCHMMA sat in an ELECT + 6 x R2UR.BROADCAST + BRA.U.ANY loop; with elect.sync issuers the tiled',
         "kernel's specialised instantiations (<.., 2, 2, 64> etc.) issue their UTCHMMAs back to back.",
         '',
         '%-78s' % 'kernel' + ''.join('%11s' % k for k in MN)]
tot = collections.Counter()
other = collections.Counter()
n_other = 0
for (n, c) in counts.items():
    if not any(c[k] for k in ('UTCHMMA', 'UTMALDG', 'UBLKCP', 'STTM')):
        if any(c[k] for k in MN[:8]):
            other.update(c); n_other += 1
        continue
    lines.append('%-78s' % demangled(n)[:78] + ''.join('%11d' % c[k] for k in MN))
    tot.update(c)
lines.append('%-78s' % ('%d other kernels with cp.async / mbarrier ops (fp32 FMA paths)' % n_other) + ''.join('%11d' % other[k] for k in MN))
tot.update(other)
lines.append('%-78s' % 'all kernels of the library' + ''.join('%11d' % tot[k] for k in MN))
open(os.path.join(ROOT, 'profiles', 'r02_sass_evidence.txt'), 'w').write('\n'.join(lines) + '\n')
print('\n'.join(lines))
