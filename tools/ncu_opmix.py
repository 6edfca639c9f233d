#!/usr/bin/env python3
"""Executed-instruction mix of one kernel from an ncu report (SASS opcode histogram weighted by
executed count).  usage: python tools/ncu_opmix.py report.ncu-rep kernel_regex"""
import collections
import csv
import io
import subprocess
import sys


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + kern, "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    hdr, data = None, []
    for r in csv.reader(io.StringIO(raw)):
        if r and r[0] == "Address":
            if hdr is not None:
                break          # the page repeats per view: keep the first copy only
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            data.append(r)
    ix = {h: i for i, h in enumerate(hdr)}
    hist, samp = collections.Counter(), collections.Counter()
    for r in data:
        src = r[ix["Source"]].strip()
        toks = src.split()
        op = toks[1] if toks[0].startswith("@") else toks[0]
        op = op.split(".")[0]
        hist[op] += int(r[ix["Instructions Executed"]])
        samp[op] += int(r[ix["# Samples"]])
    tot, stot = sum(hist.values()), sum(samp.values())
    print("kernel", kern, "executed warp-instructions", tot, "samples", stot)
    for op, n in hist.most_common(30):
        print("  %-12s %12d  %5.1f%%   samples %5.1f%%" % (op, n, 100.0 * n / tot, 100.0 * samp[op] / max(stot, 1)))


if __name__ == "__main__":
    main()
