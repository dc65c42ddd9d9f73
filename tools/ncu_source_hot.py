#!/usr/bin/env python
"""Aggregate an `ncu -i REP --page source --csv --print-source cuda,sass` export by CUDA source line:
instructions executed, stall samples and the top stall reasons.  Usage: ncu_source_hot.py export.csv [top_n]"""
import csv
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    rows = list(csv.reader(open(path, newline="")))
    cur_file, header = None, None
    agg = defaultdict(lambda: defaultdict(float))
    src_text = {}
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            header = None
            continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            header = r
            continue
        if header is None or len(r) != len(header):
            continue
        line = r[0]
        if not line.strip():
            continue  # SASS rows: already included in their source line's totals
        key = (cur_file, line)
        if r[1].strip():
            src_text[key] = r[1].strip()
        for i, name in enumerate(header):
            if name in ("# Samples", "Instructions Executed", "Warp Stall Sampling (All Samples)") or (name.startswith("stall_") and "Not Issued" not in name):
                try:
                    agg[key][name] += float(r[i])
                except ValueError:
                    pass
    tot_inst = sum(v["Instructions Executed"] for v in agg.values())
    tot_samp = sum(v["# Samples"] for v in agg.values())
    print(f"total instructions {tot_inst:.3e}  total samples {tot_samp:.0f}")
    stall_tot = defaultdict(float)
    for v in agg.values():
        for k, x in v.items():
            if k.startswith("stall_"):
                stall_tot[k] += x
    print("stall totals:", ", ".join(f"{k[6:]}={x / max(tot_samp, 1):.1%}" for k, x in sorted(stall_tot.items(), key=lambda t: -t[1])[:10]))
    print(f"{'file:line':32s} {'inst%':>6s} {'samp%':>6s}  top stalls | source")
    for key, v in sorted(agg.items(), key=lambda t: -t[1]["# Samples"])[:top]:
        st = sorted(((k[6:], x) for k, x in v.items() if k.startswith("stall_") and x > 0), key=lambda t: -t[1])[:3]
        print(f"{key[0][:22]}:{key[1]:<8s} {v['Instructions Executed'] / max(tot_inst, 1):6.1%} {v['# Samples'] / max(tot_samp, 1):6.1%}  "
              + " ".join(f"{k}={x / max(v['# Samples'], 1):.0%}" for k, x in st) + " | " + src_text.get(key, "")[:90])


if __name__ == "__main__":
    main()
