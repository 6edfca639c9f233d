#!/usr/bin/env python3
"""Per-kernel-class device time of one solver configuration (development aid, not the judged bench).

  python tools/bench_solver.py --solver unparnmpc --batch 16384 --steps 20 [--line-search]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--solver", default="unocp", choices=["unocp", "unparnmpc"])
    ap.add_argument("--batch", type=int, default=16384)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--N", type=int, default=20)
    ap.add_argument("--line-search", action="store_true")
    args = ap.parse_args()
    import torch
    import bench
    import idocp_b200 as I
    lib = I.default_library()
    prob = I.benchmark_problem(lib, N=args.N, T=0.05 * args.N)
    q0, v0 = bench.initial_states(0, args.batch, list(prob.q_min), list(prob.q_max))
    cls = I.UnOCPSolver if args.solver == "unocp" else I.UnParNMPCSolver
    s = cls(prob, args.batch)
    s.setSolution("q", q0)
    s.setSolution("v", v0)
    if args.solver == "unparnmpc":
        s.initBackwardCorrection(0.0)
    qd, vd = torch.from_numpy(q0).cuda(), torch.from_numpy(v0).cuda()
    torch.cuda.synchronize()
    for _ in range(args.warmup):
        s.updateSolutionDevice(0.0, qd.data_ptr(), vd.data_ptr(), args.line_search)
    s.sync()
    stream = torch.cuda.ExternalStream(s.stream())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.setProfiling(True)
    e0.record(stream)
    for _ in range(args.steps):
        s.updateSolutionDevice(0.0, qd.data_ptr(), vd.data_ptr(), args.line_search)
    e1.record(stream)
    s.sync()
    ms = e0.elapsed_time(e1) / args.steps
    prof = {k: {"ms_per_step": v[0] / args.steps, "launches_per_step": v[1] / args.steps} for k, v in s.getProfile().items()}
    s.computeKKTResidualDevice(0.0, qd.data_ptr(), vd.data_ptr())
    kkt = s.KKTError()
    print(json.dumps({"solver": args.solver, "batch": args.batch, "N": args.N, "line_search": args.line_search,
                      "ms_per_step": ms, "instance_iterations_per_s": args.batch / (ms * 1e-3), "kernels": prof,
                      "kkt_median": float(np.nanmedian(kkt)), "kkt_nan": int(np.isnan(kkt).sum()),
                      "status_nonzero": int((s.getStatus() != 0).sum())}))


if __name__ == "__main__":
    main()
