"""Turn what a gpurun visit brought back (gpurun_out/) into the tracked summaries under profiles/.

    python tools/ncu_summary.py <tag> [--rep gpurun_out/prof_sampler.ncu-rep] [--launches gpurun_out/launches.csv]

Writes profiles/<tag>_launches.csv (every launch: kernel, grid, block, ns), profiles/<tag>_launches.md (per-kernel
share of the step) and profiles/<tag>_ncu_full.md (the metrics quoted in DESIGN.md from the --set full capture), and
updates profiles/traffic.json (dram bytes per launch, read by bench.py for roofline.traffic)."""
import argparse
import collections
import csv
import json
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_bytes.sum", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__cycles_active.avg",
    "smsp__pcsamp_sample_count",
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("tag")
    ap.add_argument("--rep", default=os.path.join(ROOT, "gpurun_out", "prof_sampler.ncu-rep"))
    ap.add_argument("--launches", default=os.path.join(ROOT, "gpurun_out", "launches.csv"))
    ap.add_argument("--workload", default="headline")
    ap.add_argument("--note", default="")
    a = ap.parse_args()
    out = os.path.join(ROOT, "profiles")
    os.makedirs(out, exist_ok=True)

    if os.path.exists(a.launches):
        rows = list(csv.DictReader(l for l in open(a.launches) if l.startswith('"')))
        agg = collections.defaultdict(lambda: [0, 0.0])
        with open(os.path.join(out, a.tag + "_launches.csv"), "w") as f:
            f.write("id,kernel,grid,block,ns\n")
            for r in rows:
                f.write('%s,"%s","%s","%s",%s\n' % (r["ID"], r["Kernel Name"][:120].replace('"', "'"), r["Grid Size"],
                                                  r["Block Size"], r["Metric Value"]))
                k = r["Kernel Name"][:100]
                agg[k][0] += 1
                agg[k][1] += float(r["Metric Value"])
        tot = sum(v[1] for v in agg.values())
        with open(os.path.join(out, a.tag + "_launches.md"), "w") as f:
            f.write("# %s: launch list (ncu --metrics gpu__time_duration.sum --clock-control none)\n\n%s\n\n" % (a.tag, a.note))
            f.write("| share | total us | launches | kernel |\n|---:|---:|---:|---|\n")
            for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
                f.write("| %.1f%% | %.1f | %d | `%s` |\n" % (100 * v[1] / tot, v[1] / 1e3, v[0], k))

    if os.path.exists(a.rep):
        raw = subprocess.run(["ncu", "-i", a.rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        hdr, units = rows[0], rows[1]
        with open(os.path.join(out, a.tag + "_ncu_full.md"), "w") as f:
            f.write("# %s: ncu --set full --clock-control none (one launch per row block)\n\n%s\n\n" % (a.tag, a.note))
            for vals in rows[2:]:
                d = dict(zip(hdr, vals))
                u = dict(zip(hdr, units))
                f.write("## %s  grid %s block %s\n\n| metric | value | unit |\n|---|---:|---|\n" % (
                    d.get("Kernel Name", "?")[:100], d.get("Grid Size"), d.get("Block Size")))
                for k in KEYS:
                    if k in d:
                        f.write("| %s | %s | %s |\n" % (k, d[k], u[k]))
                stalls = {h: float(d[h]) for h in hdr if h.startswith("smsp__pcsamp_warps_issue_stalled_")
                          and not h.endswith("_not_issued") and d[h] not in ("", "n/a")}
                tot = sum(stalls.values()) or 1.0
                f.write("\nwarp-state samples: " + ", ".join("%s %.1f%%" % (k.replace("smsp__pcsamp_warps_issue_stalled_", ""), 100 * v / tot)
                                                             for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:9]) + "\n\n")

                def to_bytes(key):
                    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u[key]]
                    return float(d[key]) * mult
                tpath = os.path.join(out, "traffic.json")
                tj = json.load(open(tpath)) if os.path.exists(tpath) else {}
                tj[a.workload] = to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum")
                tj[a.workload + "_capture"] = a.tag + ": " + a.note
                json.dump(tj, open(tpath, "w"), indent=1)


if __name__ == "__main__":
    main()
