"""Aggregate the warp-stall samples of one kernel of an `ncu --set full --import-source on` report per SOURCE LINE: joins the
SASS page of the report (`ncu -i rep --page source --csv`) with nvdisasm's line table of the shipped cubin.
Usage: python tools/ncu_lines.py <sass.csv> <nvdisasm -g output> <mangled kernel name> [top]"""
import collections
import csv
import re
import sys

sass_csv, disasm, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
lines = open(disasm).read().split('\n')
start = next(i for i, l in enumerate(lines) if re.match(r'\s*\.section\s+\.text\.' + re.escape(kern), l) or l.startswith('.text.' + kern))
addr2line, cur = {}, None
for l in lines[start + 1:]:
    if re.match(r'\s*\.section', l) and kern not in l:
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(\S.*?);', l)
    if m:
        addr2line[int(m.group(1), 16)] = (cur, m.group(2))
rows = list(csv.reader(open(sass_csv)))
hdr = rows[1]
ai, si = hdr.index('Address'), hdr.index('# Samples')
agg, tot, base = collections.Counter(), 0, None
for r in rows[2:]:
    try:
        a = int(r[ai], 16) if r[ai].startswith('0x') else int(r[ai])
    except ValueError:
        continue
    base = a if base is None else base
    n = int(r[si] or 0)
    tot += n
    agg[addr2line.get(a - base, (None, ''))[0]] += n
print('total samples', tot)
for k, v in agg.most_common(top):
    print(f'{100 * v / tot:6.2f}%  {k}')
