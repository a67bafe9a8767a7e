"""Attribute executed instructions / stall samples of an ncu report to CUDA source lines (dev tool).
usage: python tools/ncu_lines.py <report.ncu-rep> <object.o> <mangled kernel substring> <units>"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

rep, obj, kern, units = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
tmp = tempfile.mkdtemp()
subprocess.run(f'cd {tmp} && cuobjdump -xelf all {os.path.abspath(obj)} > /dev/null', shell=True, check=True)
cubin = [f for f in os.listdir(tmp) if f.endswith('.cubin')][0]
dis = subprocess.run(f'nvdisasm -g -c {tmp}/{cubin}', shell=True, capture_output=True, text=True).stdout.split('\n')
start = next(i for i, l in enumerate(dis) if l.startswith('.text.') and kern in l and l.rstrip().endswith(':'))
cur, addr2line = None, {}
for l in dis[start + 1:]:
    if l.startswith('//---------------------'):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m:
        addr2line[int(m.group(1), 16)] = (cur, m.group(2))
src = subprocess.run(f'ncu -i {rep} --page source --csv', shell=True, capture_output=True, text=True).stdout
rows = list(csv.reader(src.split('\n')))
hdr = rows[1]
data = [r for r in rows[2:] if len(r) > 10]
iA, iN, iE, iS = hdr.index('Address'), hdr.index('# Samples'), hdr.index('Instructions Executed'), hdr.index('Source')
base = int(data[0][iA], 16)
byline, sline, ops, sops = collections.Counter(), collections.Counter(), collections.Counter(), collections.Counter()
tot = tots = 0
for r in data:
    ln = addr2line.get(int(r[iA], 16) - base, (None, ''))[0]
    e, s = int(r[iE] or 0), int(r[iN] or 0)
    byline[ln] += e
    sline[ln] += s
    m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)', r[iS])
    op = m.group(2) if m else '?'
    ops[op] += e
    sops[op] += s
    tot += e
    tots += s
print(f'total warp-instructions per unit: {tot / units:.1f}')
print('--- by opcode')
for op, e in ops.most_common(18):
    print(f'{op:10s} {e / units:9.1f}/unit {100 * e / tot:5.1f}% instr {100 * sops[op] / tots:5.1f}% samples')
print('--- by source line')
srcs = {}
for ln, e in byline.most_common(40):
    if ln is None:
        continue
    if ln[0] not in srcs:
        p = os.path.join(os.path.dirname(os.path.abspath(obj)), '..', 'csrc', ln[0])
        srcs[ln[0]] = open(p).read().split('\n') if os.path.exists(p) else []
    txt = srcs[ln[0]][ln[1] - 1].strip()[:64] if len(srcs[ln[0]]) >= ln[1] else ''
    print(f'{ln[0][:20]:20s}:{ln[1]:4d} {e / units:7.1f}/unit {100 * sline[ln] / tots:5.1f}%smp  {txt}')
