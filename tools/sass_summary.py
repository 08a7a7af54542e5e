"""Per-kernel counts of the SASS mnemonics that prove (or disprove) a Blackwell-native kernel (B200_PROFILING.md: tcgen05.mma ->
UTC*MMA, tcgen05.ld -> LDTM, TMA -> UTMALDG / UTMASTG, legacy tensor path -> HMMA) in the shipped library.

    python tools/sass_summary.py > profiles/r02_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'megadetector_b200', 'libmdb200.so')
WATCH = ['UTCHMMA', 'UTCQMMA', 'UTCBAR', 'LDTM', 'STTM', 'UTMALDG', 'UTMASTG', 'UBLKCP', 'UTMAPF', 'SYNCS', 'ELECT', 'HMMA', 'MUFU',
         'LDG', 'STG', 'LDS', 'STS', 'ATOMS', 'ATOMG', 'RED', 'SHFL', 'MATCH', 'VOTE', 'BAR', 'NANOSLEEP', 'ACQBULK', 'FFMA', 'FMUL',
         'FADD', 'IMAD']


def main():
    out = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True).stdout
    arch = re.findall(r'arch = (sm_\w+)', out)
    fn, counts, total = None, collections.defaultdict(collections.Counter), collections.Counter()
    for ln in out.splitlines():
        m = re.search(r'Function : (\S+)', ln)
        if m:
            fn = m.group(1)
            continue
        m = re.match(r'\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)', ln)
        if m and fn:
            op = m.group(1)
            total[fn] += 1
            for w in WATCH:
                if op.startswith(w):
                    counts[fn][w] += 1
                    break
    demangle = subprocess.run(['c++filt'] + list(total), capture_output=True, text=True).stdout.splitlines()
    print('# SASS summary of megadetector_b200/libmdb200.so (cuobjdump -sass), arch {}'.format(sorted(set(arch))))
    print('# tcgen05.mma -> UTCHMMA, tcgen05.ld -> LDTM, tcgen05.commit -> UTCBAR, TMA loads -> UTMALDG, TMA stores -> UTMASTG (none: the')
    print('# epilogue stores with STG.256), mbarrier -> SYNCS, elect.sync -> ELECT, legacy mma.sync -> HMMA (none)')
    for fn, name in zip(total, demangle):
        print('\n{}\n  instructions: {}'.format(name.split('(')[0], total[fn]))
        print('  ' + '  '.join('{} {}'.format(w, counts[fn][w]) for w in WATCH if counts[fn][w]))


if __name__ == '__main__':
    sys.exit(main())
