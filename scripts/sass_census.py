"""Per-kernel census of the Blackwell-native SASS opcodes in libladder_sm100.so (cuobjdump -sass): tcgen05.mma -> UTC*MMA,
tcgen05.ld / st -> LDTM / STTM, TMA -> UTMALDG / UBLKCP, tcgen05.commit -> UTCBAR, mbarrier -> SYNCS, ex2 -> MUFU.EX2.
usage: python scripts/sass_census.py > profiles/<round>_sass_census.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'ladder_latent_data_distribution_modelling_b200', 'libladder_sm100.so')
OPS = ['UTCHMMA', 'UTCQMMA', 'UTCMMA', 'LDTM', 'STTM', 'UTMALDG', 'UTMASTG', 'UBLKCP', 'UTCBAR', 'SYNCS', 'MUFU.EX2', 'HMMA', 'RED.E.ADD']


def main():
    sass = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True).stdout
    demangle = lambda n: subprocess.run(['c++filt', n], capture_output=True, text=True).stdout.strip()
    cur, counts, first = None, collections.OrderedDict(), {}
    for line in sass.splitlines():
        m = re.match(r'\s*Function : (\S+)', line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        for op in OPS:
            if re.search(r'\b' + re.escape(op) + r'\b|\b' + re.escape(op) + r'\.', line):
                counts[cur][op] += 1
                first.setdefault((cur, op), line.strip())
                break
    print('# SASS census of libladder_sm100.so (`cuobjdump -sass`, sm_100a)\n')
    print('Only kernels that contain at least one tensor-core / TMEM / TMA opcode or MUFU.EX2 are listed; `HMMA` (legacy mma.sync) must be absent.\n')
    print('| kernel | ' + ' | '.join(OPS) + ' |')
    print('|---|' + '---|' * len(OPS))
    tot = collections.Counter()
    for fn, c in counts.items():
        tot.update(c)
        if not any(c[o] for o in OPS if o not in ('SYNCS', 'RED.E.ADD')):
            continue
        name = demangle(fn)
        name = re.sub(r'\(.*', '', name).replace('void ', '').replace('ladder::', '')
        print('| `%s` | ' % name[:70] + ' | '.join(str(c[o]) if c[o] else '' for o in OPS) + ' |')
    print('| **total** | ' + ' | '.join(str(tot[o]) for o in OPS) + ' |')
    print('\n## First occurrence per opcode in the main kernels\n')
    for key in ('tma_kernel', 'mix_tc_grad_kernel', 'mix_tc_kernel', 'mix_kernel'):
        for fn in counts:
            if key in demangle(fn) and any(counts[fn][o] for o in OPS[:9]):
                print('`%s`' % re.sub(r'\(.*', '', demangle(fn)).replace('void ', '')[:90])
                print('```')
                for op in OPS:
                    if (fn, op) in first:
                        print(first[(fn, op)][:150])
                print('```')
                break


if __name__ == '__main__':
    main()
