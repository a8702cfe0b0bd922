"""Static SASS instruction count per source line of an `nvdisasm -g` listing: python tools/sass_by_line.py dis.txt [top]"""
import collections
import re
import sys

cur = ("?", 0)
cnt = collections.Counter()
byfile = collections.Counter()
for ln in open(sys.argv[1]):
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+", ln):
        cnt[cur] += 1
        byfile[cur[0]] += 1
tot = sum(cnt.values())
print("total", tot, dict(byfile.most_common(8)))
for k, v in cnt.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 40):
    print("%5d  %s:%d" % (v, k[0], k[1]))
