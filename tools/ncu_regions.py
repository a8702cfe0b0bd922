"""Execution profile of a kernel by straight-line region: consecutive SASS instructions with the same executed count,
from `ncu --page source --csv` joined with `nvdisasm -g` line annotations.
Usage: ncu_regions.py src.csv disasm.txt [min_share_pct]"""
import csv
import re
import sys

src_csv, dis_txt = sys.argv[1], sys.argv[2]
min_share = float(sys.argv[3]) if len(sys.argv) > 3 else 0.5
line_of = {}
cur = ("?", 0)
for ln in open(dis_txt):
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        line_of[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
iI, iS = hdr.index("Instructions Executed"), hdr.index("# Samples")
base = None
regions = []  # [count, n_instr, samples, first_off, lines set]
tot = ts = 0
for r in rows[2:]:
    if len(r) <= iI or not r[0].startswith("0x"):
        continue
    addr = int(r[0], 16)
    base = addr if base is None else base
    off = addr - base
    n, s = int(r[iI] or 0), int(r[iS] or 0)
    tot += n
    ts += s
    key = line_of.get(off, ("?", 0))
    if regions and regions[-1][0] == n:
        regions[-1][1] += 1
        regions[-1][2] += s
        regions[-1][4].append(key)
    else:
        regions.append([n, 1, s, off, [key]])
print("total warp-instructions %d, samples %d" % (tot, ts))
for cnt, ni, s, off, keys in regions:
    share = 100.0 * cnt * ni / tot
    if share < min_share and 100.0 * s / ts < min_share:
        continue
    files = {}
    for f, l in keys:
        files.setdefault(f, []).append(l)
    desc = "; ".join("%s:%d-%d" % (f.replace("lmc_", "").replace(".cuh", ""), min(ls), max(ls)) for f, ls in files.items())
    print("%6.2f%% inst %6.2f%% samp  exec %9d x %4d instr  @%05x  %s" % (share, 100.0 * s / ts, cnt, ni, off, desc[:150]))
