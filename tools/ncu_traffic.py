#!/usr/bin/env python
"""Summarise `ncu --set full` captures for profiles/: per kernel (mean over the captured launches) duration,
DRAM read/write bytes, DRAM / tensor-pipe / issue utilisation, registers, shared memory.

    python tools/ncu_traffic.py ROUND gpurun_out/prof_train.ncu-rep [gpurun_out/prof_infer.ncu-rep ...]

writes profiles/<ROUND>_ncu_kernels.md and merges profiles/traffic.json (launcher -> dram bytes per launch of its
main kernel; bench.py reports it as roofline.traffic).  Needs the `ncu` CLI (no GPU)."""
import csv
import io
import json
import os
import re
import subprocess
import sys
from collections import OrderedDict

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LAUNCHER = {"head_fwd_kernel": "gg_head_fwd", "hav_ce_stream_kernel": "gg_hav_ce_fwd_bwd", "head_bwd_kernel": "gg_head_bwd",
            "fuse_flat_kernel": "gg_fuse_headings", "fuse_rows_kernel": "gg_fuse_headings (with norms)",
            "fuse_and_cast_kernel": "gg_fuse_and_prepare", "cast_weight_kernel": "gg_prepare_head_weights",
            "proto_retrieve_kernel": "gg_proto_retrieve", "proto_refine_kernel": "gg_proto_refine",
            "hav_row_stats_kernel": "gg_hav_row_stats"}
METRICS = OrderedDict([
    ("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "MB read"), ("dram__bytes_write.sum", "MB written"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps %"),
    ("launch__registers_per_thread", "regs"), ("launch__shared_mem_per_block_dynamic", "dyn smem KB"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block")])
UNIT_SCALE = {"nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "ns": 1e-3, "us": 1.0, "ms": 1e3, "byte": 1e-6, "Kbyte": 1e-3,
              "Mbyte": 1.0, "Gbyte": 1e3}


def load(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    agg = OrderedDict()
    for r in rows[2:]:
        name = re.sub(r"\(.*", "", r[idx["Kernel Name"]]).replace("void ", "")
        base = re.sub(r"<.*", "", name)
        a = agg.setdefault(name, {"base": base, "n": 0, "sum": {m: 0.0 for m in METRICS}})
        a["n"] += 1
        for m in METRICS:
            if m in idx and r[idx[m]]:
                v = float(r[idx[m]].replace(",", ""))
                v *= UNIT_SCALE.get(units[idx[m]], 1.0) if ("bytes" in m or "time" in m) else 1.0
                if m == "launch__shared_mem_per_block_dynamic":
                    v *= {"byte": 1e-3, "Kbyte": 1.0, "Mbyte": 1e3}.get(units[idx[m]], 1.0)
                a["sum"][m] += v
    return agg


def main():
    rnd, reps = sys.argv[1], sys.argv[2:]
    lines = [f"# {rnd}: `ncu --set full --clock-control none` per kernel (means over the captured launches)\n"]
    traffic_path = os.path.join(REPO, "profiles", "traffic.json")
    traffic = json.load(open(traffic_path)) if os.path.exists(traffic_path) else {}
    for rep in reps:
        agg = load(rep)
        lines.append(f"\n## {os.path.basename(rep)}\n")
        lines.append("| kernel | launches | " + " | ".join(METRICS.values()) + " |")
        lines.append("|---|---:|" + "---:|" * len(METRICS))
        for name, a in agg.items():
            mean = {m: a["sum"][m] / a["n"] for m in METRICS}
            lines.append(f"| `{name[:48]}` | {a['n']} | " + " | ".join(
                f"{mean[m]:.0f}" if METRICS[m] in ("regs", "grid", "block") else f"{mean[m]:.2f}" for m in METRICS) + " |")
            if a["base"] in LAUNCHER:
                # captures of the serving workload (file name contains "infer") get their own keys: the same launcher
                # runs on another batch there
                key = ("infer:" if "infer" in os.path.basename(rep) else "") + LAUNCHER[a["base"]]
                traffic[key] = {
                    "kernel": name, "dram_bytes": (mean["dram__bytes_read.sum"] + mean["dram__bytes_write.sum"]) * 1e6,
                    "dram_read_bytes": mean["dram__bytes_read.sum"] * 1e6, "dram_write_bytes": mean["dram__bytes_write.sum"] * 1e6,
                    "ncu_us": mean["gpu__time_duration.sum"], "source": f"profiles/{rnd}_ncu_kernels.md ({os.path.basename(rep)})"}
    with open(os.path.join(REPO, "profiles", f"{rnd}_ncu_kernels.md"), "w") as f:
        f.write("\n".join(lines) + "\n")
    with open(traffic_path, "w") as f:
        json.dump(traffic, f, indent=1, sort_keys=True)
    print("\n".join(lines))


if __name__ == "__main__":
    main()
