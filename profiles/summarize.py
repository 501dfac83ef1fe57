"""Turns gpurun_out/*.ncu-rep and launch lists into the small text summaries committed here.

    python profiles/summarize.py <tag>      # e.g. r01a -> profiles/<tag>_*.txt
"""
import csv
import glob
import os
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
    "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_static", "launch__occupancy_limit_registers",
    "sm__cycles_elapsed.avg", "smsp__cycles_active.avg",
]


def rep_summary(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    lines = []
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        lines.append(f"kernel: {d.get('Kernel Name')}   grid {d.get('Grid Size')} block {d.get('Block Size')}")
        for h, u, v in zip(hdr, units, vals):
            if h in WANT:
                lines.append(f"  {h:72s} {v} {u}")
        # stall reasons (warp states), largest first
        st = [(h, float(v)) for h, v in zip(hdr, vals)
              if h.startswith("smsp__average_warp") and h.endswith("_per_issue_active.ratio") and v not in ("", "n/a")]
        st.sort(key=lambda x: -x[1])
        for h, v in st[:8]:
            lines.append(f"  stall {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):40s} {v:.3f} warps/issue")
    return "\n".join(lines)


def launch_summary(path):
    rows = [r for r in csv.reader(open(path)) if r and not r[0].startswith("==")]
    hdr = rows[0]
    ki, mi, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    ui = hdr.index("Metric Unit")
    agg = defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        if r[mi] != "gpu__time_duration.sum":
            continue
        v = float(r[vi].replace(",", ""))
        scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui], 1e-6)
        name = r[ki].split("(")[0]
        agg[name][0] += 1
        agg[name][1] += v * scale
    total = sum(v[1] for v in agg.values())
    lines = [f"{len(rows) - 1} launches, {total:.3f} ms serialised device time (cold cache; compare SHARES)",
             f"{'kernel':60s} {'launches':>8s} {'ms':>10s} {'share':>7s}"]
    for k, (n, ms) in sorted(agg.items(), key=lambda x: -x[1][1]):
        lines.append(f"{k[:60]:60s} {n:8d} {ms:10.3f} {100 * ms / total:6.1f}%")
    return "\n".join(lines)


def traffic_entries(path, workload, cull_mode, source):
    """{workload/cull_mode/stage: dram bytes per launch} for bench.py's roofline.traffic"""
    stage_of = {"k_visibility_fused": "traverse", "k_sort_write": "compact", "k_cam_plan": "cull",
                "k_traverse": "traverse", "k_cull_exhaustive": "cull"}
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    out = {}
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        u = dict(zip(hdr, units))
        name = d.get("Kernel Name", "").split("(")[0].split("<")[0].replace("void ", "").strip()
        if name not in stage_of:
            continue
        tot = 0.0
        for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u[m]]
            tot += float(d[m].replace(",", "")) * scale
        out[f"{workload}/{cull_mode}/{stage_of[name]}"] = {"dram_bytes": tot, "kernel": name, "source": source}
    return out


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "--traffic":
        # python profiles/summarize.py --traffic <rep> <workload> <cull_mode> <source label>
        import json
        here = os.path.dirname(os.path.abspath(__file__))
        p = os.path.join(here, "traffic.json")
        cur = json.load(open(p)) if os.path.exists(p) else {}
        cur.update(traffic_entries(sys.argv[2], sys.argv[3], sys.argv[4], sys.argv[5]))
        json.dump(cur, open(p, "w"), indent=1, sort_keys=True)
        print(json.dumps(cur, indent=1))
        return
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    here = os.path.dirname(os.path.abspath(__file__))
    for rep in sorted(glob.glob(os.path.join(OUT, "prof_*.ncu-rep"))):
        name = os.path.basename(rep)[5:-8]
        open(os.path.join(here, f"{tag}_ncu_{name}.txt"), "w").write(rep_summary(rep) + "\n")
        print("wrote", f"{tag}_ncu_{name}.txt")
    for lst in sorted(glob.glob(os.path.join(OUT, "launches_*.csv"))):
        name = os.path.basename(lst)[9:-4]
        open(os.path.join(here, f"{tag}_launches_{name}.txt"), "w").write(launch_summary(lst) + "\n")
        print("wrote", f"{tag}_launches_{name}.txt")


if __name__ == "__main__":
    main()
