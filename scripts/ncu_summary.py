#!/usr/bin/env python
"""Summarise an .ncu-rep (one `ncu --set full` capture) into a small JSON for profiles/.

    python scripts/ncu_summary.py gpurun_out/x.ncu-rep profiles/rNN_name.json --note "..." [--command "..."]
"""
import argparse
import csv
import io
import json
import subprocess

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "sm__cycles_active.avg",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sass__inst_executed_register_spilling", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
    "smsp__average_warp_latency_per_inst_issued.ratio",
]
STALLS = "smsp__average_warps_issue_stalled_"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("out")
    ap.add_argument("--note", default="")
    ap.add_argument("--command", default="")
    ap.add_argument("--workload", default="")
    a = ap.parse_args()
    txt = subprocess.run(["ncu", "-i", a.rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    out = []
    for vals in rows[2:]:
        rec = dict(zip(hdr, zip(vals, units)))
        m = {}
        for k in KEYS:
            if k in rec and rec[k][0] != "":
                m[k] = {"value": rec[k][0], "unit": rec[k][1]}
        stalls = {h[len(STALLS):].replace("_per_issue_active.ratio", ""): float(v[0]) for h, v in rec.items()
                  if h.startswith(STALLS) and h.endswith("_per_issue_active.ratio") and v[0] not in ("", "0")}
        m["stall_warps_per_issue"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1]))
        out.append({"kernel": rec.get("Kernel Name", ("?",))[0], "metrics": m})
    doc = {"source": a.rep, "command": a.command, "workload": a.workload, "note": a.note, "launches": out}
    with open(a.out, "w") as f:
        json.dump(doc, f, indent=1)
    print(json.dumps(doc, indent=1)[:3000])


if __name__ == "__main__":
    main()
