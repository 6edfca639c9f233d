#!/bin/bash
# dev experiment (run under gpurun): register-cap sweep of k_parnmpc_invert; the default build is restored at the end
set -e
cd "$(dirname "$0")/.."
for minb in 2 3 4; do
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false --shared -Xcompiler -fPIC \
     -DIDOCP_INV_MINB=$minb -Xptxas -v -o idocp_b200/libidocp_b200.so idocp_b200/csrc/capi.cu 2> gpurun_out/ptxas_inv_$minb.log
  echo "INV_MINB=$minb"; grep -A2 "k_parnmpc_invert" gpurun_out/ptxas_inv_$minb.log | grep -E "spill|registers" | tr '\n' ' '; echo
  python tools/bench_solver.py --solver unparnmpc --steps 10 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), {k:round(v['ms_per_step'],3) for k,v in d['kernels'].items()})"
done
