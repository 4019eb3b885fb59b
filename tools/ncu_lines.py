#!/usr/bin/env python
"""Join an `ncu --page source --csv --print-source sass` export with `nvdisasm -g` line info of the same build: executed warp
instructions and stall samples per SOURCE LINE of one kernel.  usage: ncu_lines.py src_sass.csv disasm.txt mangled_kernel_name [top]"""
import csv, re, sys, collections
src, dis, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
rows = list(csv.reader(open(src)))
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
base = int(data[0][ix['Address']], 16)
line_of = {}
cur = None
inside = False
for ln in open(dis, errors='replace'):
    if ln.startswith('.text.'):
        inside = ln.startswith('.text.' + kname + ':')
        continue
    if not inside:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    m = re.match(r'\s*/\*([0-9a-f]{4,6})\*/', ln)
    if m:
        line_of[int(m.group(1), 16)] = cur
cols = ['Instructions Executed', '# Samples', 'stall_no_inst', 'stall_long_sb', 'stall_short_sb', 'stall_wait', 'stall_barrier', 'L1 Wavefronts Shared Excessive']
agg = collections.defaultdict(lambda: [0] * len(cols))
tot = [0] * len(cols)
for r in data:
    a = int(r[ix['Address']], 16) - base
    k = line_of.get(a, ('?', 0))
    for j, c in enumerate(cols):
        v = int(r[ix[c]] or 0)
        agg[k][j] += v
        tot[j] += v
print('%-28s' % 'file:line', *['%13s' % c[:13] for c in cols])
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print('%-28s' % ('%s:%d' % k), *['%13d' % x for x in v])
print('%-28s' % 'total', *['%13d' % x for x in tot])
