#!/bin/bash
# ncu --set full capture of the hot kernels of one bench workload, on the GPU box (run under gpurun, ONE GPU):
#   tools/ncu_capture.sh <label> <workload> <kernel-regex> [extra bench.py args]
# -> gpurun_out/<label>_<workload>.ncu-rep, profiles/<label>_ncu_full_<workload>.txt (tools/ncu_summary.py) and
#    profiles/kernel_counters_<workload>.json (tools/ncu_counters.py; parsed by bench.py for roofline.traffic /
#    frac_fp64).  The first launch of every kernel after the warm-up iterations is captured (cold-cache,
#    serialised replays: use the SHARES and the byte / instruction counts, never the durations, as bench values).
set -euo pipefail
LABEL=$1; WORKLOAD=$2; REGEX=$3; shift 3
mkdir -p gpurun_out profiles
REP=gpurun_out/${LABEL}_${WORKLOAD}
ncu --set full --metrics smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum --clock-control none --import-source on -k "regex:${REGEX}" --launch-skip "${NCU_SKIP:-40}" -c "${NCU_COUNT:-8}" \
    -f -o "$REP" python bench.py --workload "$WORKLOAD" --steps 3 --warmup 3 --no-cpu-baseline "$@" > "gpurun_out/${LABEL}_${WORKLOAD}_ncu.log" 2>&1 || { tail -20 "gpurun_out/${LABEL}_${WORKLOAD}_ncu.log"; exit 1; }
python tools/ncu_summary.py "$REP.ncu-rep" "profiles/${LABEL}_ncu_full_${WORKLOAD}.txt" > /dev/null
python tools/ncu_counters.py "$REP.ncu-rep" "profiles/kernel_counters_${WORKLOAD}.json" "profiles/${LABEL}_ncu_full_${WORKLOAD}.txt"
mkdir -p gpurun_out/profiles
# the .ncu-rep is tens of MiB: gpurun merges at most 64 MiB back, so keep it only on request
[ "${KEEP_REP:-0}" = 1 ] || rm -f "$REP.ncu-rep"
cp "profiles/${LABEL}_ncu_full_${WORKLOAD}.txt" "profiles/kernel_counters_${WORKLOAD}.json" gpurun_out/profiles/
