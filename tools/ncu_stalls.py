"""Top stalled SASS instructions of a kernel with their dominant stall reasons and source line.
Usage: ncu_stalls.py src.csv disasm.txt [top]"""
import csv
import re
import sys

src_csv, dis_txt = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
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
iS = hdr.index("# Samples")
reasons = [(i, h[6:]) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
base = None
out = []
ts = 0
for r in rows[2:]:
    if len(r) <= iS or not r[0].startswith("0x"):
        continue
    addr = int(r[0], 16)
    base = addr if base is None else base
    s = int(r[iS] or 0)
    ts += s
    rs = sorted(((int(r[i] or 0), n) for i, n in reasons), reverse=True)[:2]
    out.append((s, addr - base, r[1].strip(), rs))
out.sort(reverse=True)
for s, off, sass, rs in out[:top]:
    f, l = line_of.get(off, ("?", 0))
    print("%5.2f%%  @%05x %-46s %-28s %s:%d" % (100.0 * s / ts, off, sass[:46], " ".join("%s=%d" % (n, c) for c, n in rs if c), f, l))
