"""cuobjdump -sass vit_search_b200/libvsx.so | python tools/sass_summary.py  ->  profiles/r2_sass_summary.md
Instruction mnemonics per kernel family: the proof that the hot kernels are tcgen05 / TMEM / TMA code and which ones are legacy mma.sync."""
import collections
import os
import re
import subprocess
import sys

KEYS = ['UTCHMMA', 'UTCBAR', 'LDTM', 'UTMALDG', 'UTMASTG', 'UTMAREDG', 'UBLKCP', 'HMMA', 'SYNCS', 'MUFU.EX2']
cur, cnt = None, collections.OrderedDict()
for line in sys.stdin:
    m = re.search(r'Function : (\S+)', line)
    if m:
        cur = m.group(1)
        cnt[cur] = collections.Counter()
        continue
    if cur:
        for k in KEYS:
            if re.search(r'(?<![A-Z])' + re.escape(k), line):      # HMMA must not count UTCHMMA
                cnt[cur][k] += 1
names = subprocess.run(['c++filt'], input='\n'.join(cnt), capture_output=True, text=True).stdout.splitlines()
agg = collections.OrderedDict()
for name, c in zip(names, cnt.values()):
    fam = re.sub(r'\(anonymous namespace\)::', '', name)
    fam = re.sub(r'^void ', '', fam)
    fam = re.sub(r'^vsx::', '', fam)
    fam = re.sub(r'[<(].*', '', fam)
    a = agg.setdefault(fam, [0, collections.Counter()])
    a[0] += 1
    a[1].update(c)
out = ['# SASS evidence (`cuobjdump -sass vit_search_b200/libvsx.so | python tools/sass_summary.py`, sm_100a)', '',
       'Instruction mnemonics per kernel family, summed over the template instantiations.  UTCHMMA = tcgen05.mma, UTCBAR = tcgen05.commit,',
       'LDTM = tcgen05.ld (tensor memory), UTMALDG / UTMASTG / UTMAREDG = TMA tensor load / store / reduce-add, UBLKCP = cp.async.bulk,',
       'HMMA = legacy mma.sync, SYNCS = mbarrier operations.', '',
       '| kernel family | instantiations | ' + ' | '.join(KEYS) + ' |', '|---|---:|' + '---:|' * len(KEYS)]
for fam, (n, c) in sorted(agg.items(), key=lambda kv: -sum(kv[1][1].values())):
    if sum(c.values()):
        out.append('| `%s` | %d | ' % (fam, n) + ' | '.join(str(c[k]) if c[k] else '' for k in KEYS) + ' |')
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'profiles', 'r2_sass_summary.md')
open(path, 'w').write('\n'.join(out) + '\n')
print('\n'.join(out))
