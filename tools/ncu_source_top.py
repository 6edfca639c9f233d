#!/usr/bin/env python3
"""Top stall sites of one kernel from an ncu report's source page (SASS level, grouped by the
instruction's position).  usage: python tools/ncu_source_top.py report.ncu-rep kernel_regex [n]"""
import csv
import io
import subprocess
import sys


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + kern, "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    rows = [r for r in csv.reader(io.StringIO(raw))]
    hdr = None
    data = []
    for r in rows:
        if r and r[0] == "Address":
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            data.append(r)
    ix = {h: i for i, h in enumerate(hdr)}
    tot = sum(int(r[ix["# Samples"]]) for r in data)
    print("kernel", kern, "total samples", tot, "SASS instructions", len(data),
          "executed", sum(int(r[ix["Instructions Executed"]]) for r in data))
    cats = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    agg = {c: sum(int(r[ix[c]]) for r in data) for c in cats}
    print("stall totals:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
    pos = {id(r): k for k, r in enumerate(data)}
    for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:n]:
        main_stall = max(cats, key=lambda c: int(r[ix[c]]))
        print("%5d  #%-5d %-14s x%-8s %s" % (int(r[ix["# Samples"]]), pos[id(r)], main_stall.replace("stall_", ""),
                                           r[ix["Instructions Executed"]], r[ix["Source"]].strip()[:100]))


if __name__ == "__main__":
    main()
