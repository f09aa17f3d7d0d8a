"""Instruction census of the shipped library: which Blackwell-native SASS instructions each kernel contains.
Usage: python tools/sass_census.py > profiles/sass_census_rN.txt   (needs cuobjdump; no GPU)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'oct_segmentation_b200', 'liboctseg.so')
WATCH = ['UTCHMMA', 'UTCQMMA', 'UTCMMA', 'LDTM', 'STTM', 'UTMALDG', 'UTMASTG', 'UBLKCP', 'UTCBAR', 'UTCATOM', 'SYNCS', 'HMMA', 'HGMMA',
         'FFMA2', 'MUFU.TANH', 'LDGSTS']


def main():
    sass = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True, check=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r'Function : (\S+)', line)
        if m:
            cur = m.group(1)
            per[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r'\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)', line)
        if m:
            op = m.group(1)
            per[cur]['_total'] += 1
            for w in WATCH:
                if op.startswith(w):
                    per[cur][w] += 1
    demangle = subprocess.run(['c++filt'], input='\n'.join(per), capture_output=True, text=True).stdout.splitlines()
    print(f'# SASS census of {os.path.relpath(LIB, ROOT)} (sm_100a), cuobjdump -sass; counts are static instructions per kernel')
    print('# tcgen05.mma -> UTCHMMA, tcgen05.ld -> LDTM, TMA load/store -> UTMALDG/UTMASTG, mbarrier -> SYNCS, legacy mma.sync would be HMMA')
    tot = collections.Counter()
    for (name, c), dn in zip(per.items(), demangle):
        short = re.sub(r'\(.*', '', dn)
        items = ', '.join(f'{w} {c[w]}' for w in WATCH if c[w])
        print(f'{short:90s} total {c["_total"]:6d}  {items}')
        tot.update(c)
    print('# library totals: ' + ', '.join(f'{w} {tot[w]}' for w in WATCH if tot[w]) + f', all instructions {tot["_total"]}')


if __name__ == '__main__':
    main()
