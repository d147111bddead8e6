"""SASS evidence of the shipped library: per-kernel instruction histogram (cuobjdump -sass) with the mnemonics that prove
the tcgen05 / TMEM / TMA / packed-fp32 paths called out, and the full listing of the dominant kernel.
usage: python tools/sass_hist.py [lib.so] > profiles/sass_gemm_tc.txt"""
import collections
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(REPO, 'gast-net-3dposeestimation_b200', 'csrc', 'libgast_b200.so')
full_for = sys.argv[2] if len(sys.argv) > 2 else 'gemm_tc_kernelILi0ELi0ELi2ELi2E'
sass = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(['c++filt', n], capture_output=True, text=True).stdout.strip()
KEY = ['UTCHMMA', 'UTCQMMA', 'UTCBAR', 'UTCATOMSWS', 'LDTM', 'STTM', 'UTMALDG', 'UTMASTG', 'UBLKCP', 'SYNCS', 'FFMA2', 'FADD2', 'HMMA', 'MUFU',
       'LDS', 'STS', 'LDG', 'STG', 'FFMA', 'SHFL', 'BAR', 'ATOM', 'RED']
funcs, cur = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        cur = m.group(1)
        funcs[cur] = []
        continue
    m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(.*?);', line)
    if cur and m:
        funcs[cur].append(m.group(1))
print('# SASS instruction histogram of %s (sm_100a), cuobjdump -sass' % os.path.basename(lib))
print('# mnemonics: UTCHMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st (tensor memory), UTMALDG = TMA tensor load,')
print('#            UBLKCP = cp.async.bulk, UTCBAR = tcgen05.commit, SYNCS = mbarrier ops, FFMA2 = fma.rn.f32x2')
tot = collections.Counter()
for name, ins in funcs.items():
    c = collections.Counter()
    for i in ins:
        op = i.split()[0]
        if op.startswith('@'):
            op = i.split()[1]
        c[op.split('.')[0]] += 1
    tot.update(c)
    keys = ' '.join('%s=%d' % (k, c[k]) for k in KEY if c[k])
    print('\n%s\n  %d instructions | %s' % (demangle(name)[:150], len(ins), keys))
print('\n# whole library: ' + ' '.join('%s=%d' % (k, tot[k]) for k in KEY if tot[k]))
print('\n# ---- full listing of %s ----' % full_for)
for name, ins in funcs.items():
    if full_for in name:
        for i in ins:
            print('    ' + i)
        break
