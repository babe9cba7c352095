#!/usr/bin/env python
"""Warp-stall samples of one kernel from `ncu --page source --print-source sass --csv`, summed over the
stretches of SASS between barriers / branches: where the warps of a kernel spend their time, phase by phase.

    ncu -i x.ncu-rep --page source --csv --print-source sass > x_sass.csv
    python scripts/ncu_segments.py x_sass.csv [min_percent]
"""
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    thresh = float(sys.argv[2]) if len(sys.argv) > 2 else 0.3
    hdr, data = rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}

    def f(r, k):
        try:
            return float(r[ix[k]])
        except (ValueError, KeyError):
            return 0.0

    tot = sum(f(r, "# Samples") for r in data)
    stallk = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    new = lambda i: {"start": i, "samples": 0, "n": 0, "fp64": 0, "lds": 0, "sts": 0, "ldg": 0, "stg": 0, "shfl": 0, "exec": 0,
                     "wave": 0, "stalls": {}}
    seg, cur = [], new(0)
    for i, r in enumerate(data):
        src = r[ix["Source"]]
        parts = src.split()
        op = parts[1] if src.startswith("@") and len(parts) > 1 else (parts[0] if parts else "")
        ex = f(r, "Instructions Executed")
        cur["samples"] += f(r, "# Samples"); cur["n"] += 1; cur["exec"] += ex
        cur["wave"] += f(r, "L1 Wavefronts Shared")
        for key, pre in (("fp64", ("DFMA", "DADD", "DMUL")), ("lds", ("LDS",)), ("sts", ("STS",)), ("ldg", ("LDG", "LD.")),
                         ("stg", ("STG", "ST.", "RED")), ("shfl", ("SHFL",))):
            if op.startswith(pre):
                cur[key] += ex
        for k in stallk:
            cur["stalls"][k] = cur["stalls"].get(k, 0) + f(r, k)
        if op.startswith(("BAR", "WARPSYNC", "BRA", "EXIT", "SYNCS")):
            cur["end"], cur["endsrc"] = i, src
            seg.append(cur)
            cur = new(i + 1)
    seg.append(cur)
    print(f"total samples {tot:.0f}, {len(data)} SASS instructions")
    for s in seg:
        if s["samples"] < tot * thresh / 100:
            continue
        top = sorted(s["stalls"].items(), key=lambda kv: -kv[1])[:4]
        print(f"[{s['start']:4d}-{s.get('end', -1):4d}] samples {100 * s['samples'] / tot:5.1f}%  inst {s['exec'] / 1e6:6.2f}M  fp64 {s['fp64'] / 1e6:6.2f}M"
              f"  lds {s['lds'] / 1e6:5.2f} sts {s['sts'] / 1e6:5.2f} ldg {s['ldg'] / 1e6:5.2f} stg {s['stg'] / 1e6:5.2f} shfl {s['shfl'] / 1e6:5.2f}"
              f"  smem wavefronts {s['wave'] / 1e6:6.2f}M | {s.get('endsrc', '')[:34]:34s} | "
              + " ".join(f"{k[6:]}={100 * v / max(1, s['samples']):.0f}%" for k, v in top))


if __name__ == "__main__":
    main()
