#!/usr/bin/env python3
"""Stall samples of one kernel per CUDA SOURCE LINE (ncu source page, -lineinfo build):
   python tools/ncu_source_lines.py report.ncu-rep kernel_regex [n]   -> the n hottest lines with their two main stall reasons"""
import csv
import io
import subprocess
import sys


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + kern, "--print-source", "sass,cuda"],
                         capture_output=True, text=True).stdout
    cur, hdr, data = None, None, []
    for r in csv.reader(io.StringIO(raw)):
        if len(r) == 2 and r[0] == "File Path":
            cur = r[1]
        elif r and r[0] == "Line No":
            hdr = r
        elif hdr and len(r) == len(hdr) and r[0].isdigit():
            data.append((cur, r))
    ix = {}
    for i, h in enumerate(hdr):
        ix.setdefault(h, i)
    tot = sum(int(r[ix["# Samples"]]) for _, r in data)
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    agg = {s: sum(int(r[ix[s]]) for _, r in data) for s in stalls}
    print("kernel", kern, "samples", tot, "stall totals:", {k.replace("stall_", ""): v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
    for f, r in sorted(data, key=lambda x: -int(x[1][ix["# Samples"]]))[:n]:
        st = sorted(((int(r[ix[s]]), s) for s in stalls), reverse=True)[:2]
        print("%5d %5.1f%% %-26s:%-5s %-34s %s" % (int(r[ix["# Samples"]]), 100.0 * int(r[ix["# Samples"]]) / max(tot, 1), f.split("/")[-1], r[0],
                                                 " ".join("%s=%d" % (s.replace("stall_", ""), v) for v, s in st), r[1].strip()[:100]))


if __name__ == "__main__":
    main()
