#!/usr/bin/env python
"""Attribute executed SASS instructions of an ncu capture to source lines of one kernel file.

    ncu -i rep --page source --csv --print-source sass > sass.csv
    cuobjdump -xelf all obj.o ; nvdisasm -gi x.cubin > gi.sass
    python tools/sass_profile.py sass.csv gi.sass <function-substring> <source-file> [top]
"""
import collections
import csv
import re
import sys

sass_csv, gi, fn, srcfile = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 50
INNER = len(sys.argv) > 6 and sys.argv[6] == 'inner'   # attribute to the innermost line of the file instead of the outermost
rows = list(csv.reader(open(sass_csv)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[hi]
iI, iS = hdr.index("Instructions Executed"), hdr.index("# Samples")
ins = [(r[1].strip(), int(r[iI]), int(r[iS])) for r in rows[hi + 1:] if len(r) > iI]
base = srcfile.split('/')[-1]
infn, cur, seq = False, None, []
pending = []
for l in open(gi).read().splitlines():
    if l.startswith('.text.'):
        infn = fn in l
        continue
    if not infn:
        continue
    if '//## File' in l:
        allm = re.findall(r'"([^"]+)", line (\d+)', l)
        if 'inlined at' in l and pending is not None:
            pending.extend(allm)
        else:
            pending = list(allm)
        key = None
        for ff, nn in (reversed(pending) if INNER else pending):     # pending[0] is the innermost location
            if ff.endswith(base):
                key = int(nn)
        cur = key if key is not None else (pending[0][0].split('/')[-1], int(pending[0][1]))
        continue
    m = re.match(r'\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);', l)
    if m:
        seq.append((cur, m.group(2)))
assert len(seq) == len(ins), (len(seq), len(ins))
tot = sum(i[1] for i in ins)
ts = sum(i[2] for i in ins)
agg = collections.defaultdict(lambda: [0, 0])
for (key, _), (s, c, sm) in zip(seq, ins):
    agg[key][0] += c
    agg[key][1] += sm
src = open(srcfile).read().splitlines()
print('total warp instructions', tot, 'samples', ts)
for key, (c, sm) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    text = src[key - 1].strip()[:100] if isinstance(key, int) else str(key)
    print(f"{c / tot * 100:5.2f}% inst {sm / ts * 100:5.2f}% samp  {key if isinstance(key, int) else ''}: {text}")
