#!/bin/bash
# dev experiment (run under gpurun): register-cap sweep of k_linearize / k_riccati
set -e
cd "$(dirname "$0")/.."
for cfg in "2 2" "3 2" "4 2" "2 3" "2 4" "3 3" "4 4"; do
  set -- $cfg
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false --shared -Xcompiler -fPIC \
     -DIDOCP_LIN_MINB=$1 -DIDOCP_RIC_MINB=$2 -Xptxas -v -o idocp_b200/libidocp_b200.so idocp_b200/csrc/capi.cu 2> gpurun_out/ptxas_$1_$2.log
  python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/sweep_$1_$2.json 2>/dev/null
  python - <<PY
import json
d=json.load(open("gpurun_out/sweep_$1_$2.json"))
print("LIN_MINB=$1 RIC_MINB=$2 ms/step %.4f"%d["ms_per_step"], {k:round(v["ms_per_launch"],4) for k,v in d["roofline"]["kernels"].items()})
PY
  grep -A2 "k_linearizeILb0\|k_riccati" gpurun_out/ptxas_$1_$2.log | grep -E "spill|registers" | tr '\n' ' '; echo
done
