#!/usr/bin/env python3
"""ncu report -> per-kernel counters JSON that bench.py parses for its `roofline` object (no literals in bench.py).

  python tools/ncu_counters.py gpurun_out/<label>.ncu-rep profiles/kernel_counters_<workload>.json [label]

Per kernel (first captured launch of each name): duration, DRAM bytes read / written, executed FP64 thread
instructions (DFMA counted as 2 flop, DMUL / DADD as 1; predicated-on, i.e. the padding lane of an octet included),
FP64-pipe and issue utilisation, registers, achieved warps.  Runs wherever `ncu` is installed (no GPU needed)."""
import csv
import io
import json
import re
import subprocess
import sys

FIELDS = {
    "gpu__time_duration.sum": "duration_ns",
    "dram__bytes_read.sum": "dram_bytes_read",
    "dram__bytes_write.sum": "dram_bytes_write",
    "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum": "thread_dfma",
    "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum": "thread_dmul",
    "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum": "thread_dadd",
    "sm__inst_executed_pipe_fp64.sum": "warp_inst_fp64",
    "smsp__inst_executed.sum": "warp_inst",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active": "fp64_pipe_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "launch__registers_per_thread": "registers",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "launch__waves_per_multiprocessor": "waves",
    "sass__inst_executed_local_loads": "local_loads",
    "sass__inst_executed_local_stores": "local_stores",
}
UNIT_SCALE = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "usecond": 1e3, "msecond": 1e6, "nsecond": 1.0, "second": 1e9,
              "us": 1e3, "ms": 1e6, "ns": 1.0, "s": 1e9}


def short_name(name):
    m = re.search(r"(k_[a-z0-9_]+)", name)
    return m.group(1) if m else name


def main():
    rep, out = sys.argv[1], sys.argv[2]
    label = sys.argv[3] if len(sys.argv) > 3 else rep
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    kernels = {}
    for r in rows[2:]:
        rec = dict(zip(hdr, r))
        name = short_name(rec.get("Kernel Name", ""))
        full = rec.get("Kernel Name", "")
        key = name
        m = re.search(r"<(.*?)>", full)
        if m:
            key = "%s<%s>" % (name, m.group(1).replace(" ", ""))
        if key in kernels:
            continue
        k = {"kernel_name": full, "grid": rec.get("Grid Size"), "block": rec.get("Block Size")}
        for i, h in enumerate(hdr):
            if h in FIELDS and r[i] != "":
                try:
                    val = float(r[i].replace(",", ""))
                except ValueError:
                    continue
                val *= UNIT_SCALE.get(units[i], 1.0) if FIELDS[h] in ("duration_ns", "dram_bytes_read", "dram_bytes_write") else 1.0
                k[FIELDS[h]] = val
        for op in ("dfma", "dmul", "dadd"):     # --set full only carries the per-cycle form of these counters
            key_pc = "smsp__sass_thread_inst_executed_op_%s_pred_on.sum.per_cycle_elapsed" % op
            if "thread_" + op not in k and rec.get(key_pc, "") != "" and rec.get("smsp__cycles_elapsed.avg", "") != "":
                k["thread_" + op] = float(rec[key_pc].replace(",", "")) * float(rec["smsp__cycles_elapsed.avg"].replace(",", ""))
        if "thread_dfma" in k:
            k["fp64_flop_executed"] = 2.0 * k["thread_dfma"] + k.get("thread_dmul", 0.0) + k.get("thread_dadd", 0.0)
        if "dram_bytes_read" in k:
            k["dram_bytes"] = k["dram_bytes_read"] + k.get("dram_bytes_write", 0.0)
        kernels[key] = k
    with open(out, "w") as f:
        json.dump({"source": label, "how": "ncu --set full --clock-control none, one launch per kernel (tools/ncu_capture.sh)",
                   "kernels": kernels}, f, indent=1, sort_keys=True)
        f.write("\n")
    for key, k in kernels.items():
        print("%-40s %8.3f ms  dram %.3f GB  fp64 %.3f GFLOP  pipe %s %%  issue %s %%  regs %s" % (
            key, k.get("duration_ns", 0) / 1e6, k.get("dram_bytes", 0) / 1e9, k.get("fp64_flop_executed", 0) / 1e9,
            k.get("fp64_pipe_pct"), k.get("issue_active_pct"), k.get("registers")))


if __name__ == "__main__":
    main()
