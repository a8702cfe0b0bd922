"""Join an `ncu --page source --csv` export (SASS rows with executed counts / stall samples) with `nvdisasm -g`
line annotations, and print the hottest source lines.  Usage: ncu_by_line.py src.csv disasm.txt [top]"""
import collections
import csv
import re
import sys

src_csv, dis_txt = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 60
line_of = {}
cur = ("?", 0)
for ln in open(dis_txt):
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        line_of[int(m.group(1), 16)] = (cur, m.group(2).strip())
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
iI, iS = hdr.index("Instructions Executed"), hdr.index("# Samples")
base = None
agg, sagg, ops = collections.Counter(), collections.Counter(), collections.defaultdict(collections.Counter)
tot = ts = 0
for r in rows[2:]:
    if len(r) <= iI or not r[0].startswith("0x"):
        continue
    addr = int(r[0], 16)
    base = addr if base is None else base
    off = addr - base
    n, s = int(r[iI] or 0), int(r[iS] or 0)
    key, sass = line_of.get(off, (("?", 0), r[1]))
    agg[key] += n
    sagg[key] += s
    op = re.sub(r"^@!?U?P\d+\s+", "", sass).split()[0].split(".")[0]
    ops[key][op] += n
    tot += n
    ts += s
print("total warp-instructions %d, samples %d" % (tot, ts))
for key, n in agg.most_common(top):
    o = ", ".join("%s %.1f" % (k, 100.0 * v / tot) for k, v in ops[key].most_common(4))
    print("%5.2f%% inst %5.2f%% samp  %s:%d   [%s]" % (100.0 * n / tot, 100.0 * sagg[key] / ts, key[0], key[1], o))
