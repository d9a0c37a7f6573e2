"""Per-kernel counts of the Blackwell mnemonics in the built library (cuobjdump -sass; no GPU needed):
UTCHMMA = tcgen05.mma (.2CTA = cta_group::2), LDTM = tcgen05.ld, UTMALDG = TMA tensor load, UBLKRED = TMA bulk reduce-add,
LDGSTS = cp.async.      python scripts/sass_summary.py > profiles/r2_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'emsanet_b200', 'lib', 'libemsanet_b200.so')
sass = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True, check=True).stdout
keys = ['UTCHMMA', '.2CTA', 'LDTM', 'UTMALDG', 'UBLKRED', 'UTMAPF', 'LDGSTS', 'SYNCS']
per = collections.OrderedDict()
fn = None
for line in sass.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        fn = m.group(1)
        per[fn] = collections.Counter()
        continue
    if fn is None:
        continue
    for k in keys:
        if k in line:
            per[fn][k] += 1
demangled = subprocess.run(['c++filt'], input='\n'.join(per), capture_output=True, text=True).stdout.splitlines()
groups = collections.OrderedDict()
for name, d in zip(per, demangled):
    base = re.sub(r'<.*', '', d.split('(')[0]).replace('void ', '').strip()
    g = groups.setdefault(base, [0, collections.Counter()])
    g[0] += 1
    g[1].update(per[name])
tot = collections.Counter()
print(f'{lib.replace(ROOT + "/", "")}: {len(per)} kernels (template instances), arch sm_100a')
print(f'{"kernel (all template instances)":44s} {"inst":>4s} ' + ' '.join(f'{k:>8s}' for k in keys))
for base, (n, c) in sorted(groups.items(), key=lambda kv: -kv[1][1]['UTCHMMA']):
    tot.update(c)
    if any(c[k] for k in keys[:-1]):
        print(f'{base[:44]:44s} {n:4d} ' + ' '.join(f'{c[k]:8d}' for k in keys))
print(f'{"TOTAL":44s} {len(per):4d} ' + ' '.join(f'{tot[k]:8d}' for k in keys))
