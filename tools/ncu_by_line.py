#!/usr/bin/env python3
"""Aggregate an `ncu --page source --csv` SASS dump by CUDA source line, using nvdisasm -g line info
of the matching cubin.  usage: ncu_by_line.py <sass.csv> <dis.txt> <kernel mangled-name substring>"""
import csv
import collections
import re
import sys

sass_csv, dis_txt, kname = sys.argv[1:4]
# address -> (file, line) from nvdisasm
addr2line = {}
cur = None
inside = False
for ln in open(dis_txt):
    if ln.startswith(".text."):
        inside = kname in ln
        cur = None
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        addr2line[int(m.group(1), 16)] = (cur, m.group(2).strip())
rows = list(csv.reader(open(sass_csv)))
# the dump may hold several kernels, each introduced by a "Kernel Name" row: keep the one whose name matches
want = sys.argv[4] if len(sys.argv) > 4 else None
blocks, cur_rows = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur_rows = [r]
        blocks.append(cur_rows)
    elif cur_rows is not None:
        cur_rows.append(r)
rows = next((b for b in blocks if want is None or want in b[0][1]), blocks[0])
print("kernel:", rows[0][1])
hdr = rows[1]
ai, ii, ti, si = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
base = None
agg = collections.defaultdict(lambda: [0, 0, 0])
tot = [0, 0, 0]
for r in rows[2:]:
    if len(r) <= ti:
        continue
    a = int(r[ai], 16) if r[ai].startswith("0x") else int(r[ai])
    if base is None:
        base = a
    line, _ = addr2line.get(a - base, (None, None))
    v = [int(r[ii] or 0), int(r[ti] or 0), int(r[si] or 0)]
    for k in range(3):
        agg[line][k] += v[k]
        tot[k] += v[k]
print(f"total warp-instr {tot[0]:,}  thread-instr {tot[1]:,}  avg active threads {tot[1]/max(tot[0],1):.1f}  samples {tot[2]}")
for line, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:60]:
    print(f"{100*v[0]/tot[0]:6.2f}% instr  {100*v[2]/max(tot[2],1):6.2f}% samples  act={v[1]/max(v[0],1):5.1f}  {line}")
