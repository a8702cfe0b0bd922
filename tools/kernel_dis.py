"""Extract one kernel's `nvdisasm -g` listing from an object file: python tools/kernel_dis.py obj.o <mangled substring> out.txt
(prints the instruction count / code size)."""
import os
import re
import subprocess
import sys
import tempfile

obj, sub, out = os.path.abspath(sys.argv[1]), sys.argv[2], sys.argv[3]
with tempfile.TemporaryDirectory() as d:
    subprocess.run(["cuobjdump", "-xelf", "all", obj], cwd=d, capture_output=True)
    cub = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
    txt = subprocess.run(["nvdisasm", "-g", os.path.join(d, cub)], capture_output=True, text=True).stdout.split("\n")
start = [i for i, l in enumerate(txt) if l.startswith(".text.") and sub in l][0]
end = [i for i, l in enumerate(txt) if i > start and l.strip().startswith(".section")]
end = end[0] if end else len(txt)
body = txt[max(0, start - 8):end]
open(out, "w").write("\n".join(body))
n = sum(1 for l in body if re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+\S", l))
print("%s: %d instructions, %.1f KB" % (txt[start][:110], n, n * 16 / 1024))
