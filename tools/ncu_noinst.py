"""Where a kernel waits for instructions: warp samples, `no_instructions` samples and executed instructions per bin of
code addresses, from `ncu -i rep --page source --csv` joined with the `nvdisasm -g` listing of the same object.
Usage: ncu_noinst.py src.csv disasm.txt [bin_bytes=2048]"""
import collections
import csv
import re
import sys

src_csv, dis_txt = sys.argv[1], sys.argv[2]
BIN = int(sys.argv[3]) if len(sys.argv) > 3 else 2048
line_of, cur = {}, ("?", 0)
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
iS, iN, iI = hdr.index("# Samples"), hdr.index("stall_no_inst"), hdr.index("Instructions Executed")
base, tot, totn, toti = None, 0, 0, 0
bins = collections.defaultdict(lambda: [0, 0, 0])
for r in rows[2:]:
    if len(r) <= iS or not r[0].startswith("0x"):
        continue
    a = int(r[0], 16)
    base = a if base is None else base
    s, n, i = int(r[iS] or 0), int(r[iN] or 0), int(r[iI] or 0)
    tot, totn, toti = tot + s, totn + n, toti + i
    b = bins[(a - base) // BIN]
    b[0] += s
    b[1] += n
    b[2] += i
print("samples %d, no_instructions %d (%.1f%%), warp instructions %d" % (tot, totn, 100.0 * totn / tot, toti))
print("offset  samples%  no_inst%  executed%  no_inst/samples  source lines")
for k in sorted(bins):
    b = bins[k]
    fl = [line_of[o] for o in range(k * BIN, (k + 1) * BIN, 16) if o in line_of]
    print("%6x %8.2f %9.2f %10.2f %12.2f     %s:%d .. %s:%d" % ((k * BIN, 100.0 * b[0] / tot, 100.0 * b[1] / max(totn, 1),
          100.0 * b[2] / toti, b[1] / max(b[0], 1)) + (fl[0] + fl[-1] if fl else ("", 0, "", 0))))
