"""ptxas register / spill summary of one translation unit:  python tools/regs.py <file.cu> [extra nvcc flags]"""
import re
import subprocess
import sys

src = sys.argv[1]
cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
       "-Xcompiler", "-fPIC", "-Xptxas", "-v", "-I", "include", "-c", src, "-o", "/tmp/regs_tmp.o"] + sys.argv[2:]
out = subprocess.run(cmd, capture_output=True, text=True)
txt = out.stderr + out.stdout
if out.returncode:
    print(txt[-3000:])
    sys.exit(1)
cur = None
for ln in txt.splitlines():
    m = re.search(r"Compiling entry function '(\S+)'", ln)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(lmc_sampler_args.*", "", cur).replace("lmc::", "").replace("(int)", "")
        spill = ""
        continue
    if cur and "spill" in ln and "stack frame" in ln:
        spill = ln.strip()
    m = re.search(r"Used (\d+) registers", ln)
    if m and cur:
        print("%4s regs  %-90s %s" % (m.group(1), cur[:90], spill if "0 bytes spill stores" not in spill else ""))
        cur = None
