"""SASS opcode histogram of libysmr_b200.so per kernel -> profiles/r2_sass_opcodes.md   (python scripts/sass_histogram.py)"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, 'ysmr_b200', 'libysmr_b200.so')
out = subprocess.run(['cuobjdump', '-sass', so], capture_output=True, text=True).stdout
kern, hist = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        name = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r'\(.*', '', name).replace('void ', '').replace('ysmr::', '').replace('(anonymous namespace)::', '')
        hist[kern] = collections.Counter()
        continue
    m = re.match(r'\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
    if m and kern:
        hist[kern][m.group(1)] += 1
KEY = ['IDP.4A', 'IDP.2A', 'FFMA2', 'FADD2', 'FMUL2', 'DFMA', 'LDG.E.EF', 'LDG', 'STG', 'LDS', 'STS', 'ATOMS', 'ATOMG', 'RED', 'BAR', 'SHFL', 'VOTE',
       'CREDUX', 'REDUX', 'MUFU', 'UTMALDG', 'UBLKCP', 'LDGSTS', 'VIMNMX3', 'PRMT']
with open(os.path.join(ROOT, 'profiles', 'r2_sass_opcodes.md'), 'w') as f:
    f.write('# SASS opcode histogram (static instruction counts), `cuobjdump -sass ysmr_b200/libysmr_b200.so`, sm_100a\n\n')
    f.write('Prefix counts (e.g. `LDG` counts every LDG.* variant; `LDG.E.EF` = evict-first streaming loads of the input pixels). '
            'No TMA (`UTMALDG` / `UBLKCP`): the BGR rows of a 1228-wide frame are 3,684 bytes apart, not a multiple of 16, so neither a '
            'tensor map nor `cp.async.bulk` can address them; the input is read with 32-bit evict-first loads through a three-deep '
            'register pipeline instead (DESIGN.md section 4).\n\n')
    f.write('| kernel | total | ' + ' | '.join(KEY) + ' |\n|---|---|' + '---|' * len(KEY) + '\n')
    for k, h in hist.items():
        if sum(h.values()) < 40:
            continue
        row = [sum(v for op, v in h.items() if op.startswith(key)) for key in KEY]
        f.write(f'| `{k[:70]}` | {sum(h.values())} | ' + ' | '.join(str(v) if v else '' for v in row) + ' |\n')
print(open(os.path.join(ROOT, 'profiles', 'r2_sass_opcodes.md')).read()[:3000])
