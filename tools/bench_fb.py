#!/usr/bin/env python3
"""Timing of the batched ANYmal OCPSolver (anymal_trotting problem) per kernel class.  Usage:
   python tools/bench_fb.py [batch] [iters]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import anymal_problems as ap  # noqa: E402
import fb_py  # noqa: E402
import idocp_b200 as I  # noqa: E402
import oracle_py  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    oracle_py.build()
    fb_py.lib()
    pr = ap.TrottingProblem()
    lib = I.default_library()
    rng = np.random.default_rng(0)
    q0 = np.tile(pr.q0, (B, 1))
    q0[:, 7:] += rng.uniform(-0.02, 0.02, (B, 12))
    v0 = rng.uniform(-0.1, 0.1, (B, 18))
    solver = ap.make_product_solver(pr, lib, fb_py, batch=B, q0=q0, v0=v0)
    for _ in range(3):
        solver.updateSolution(0.0, q0, v0)
    solver.sync()
    solver.setProfiling(True)
    t0 = time.perf_counter()
    for _ in range(iters):
        solver.updateSolution(0.0, q0, v0)
    solver.sync()
    wall = (time.perf_counter() - t0) / iters
    prof = solver.getProfile()
    solver.setProfiling(False)
    solver.computeKKTResidual(0.0, q0, v0)
    kkt = solver.KKTError()
    out = {"solver": "ocp_anymal_trotting", "batch": B, "stages": len(solver.chain()), "ms_per_step_wall": wall * 1e3,
           "instance_iterations_per_s": B / wall,
           "kernels": {k: {"ms_per_step": v["ms"] / iters, "launches_per_step": v["calls"] / iters} for k, v in prof.items()},
           "kkt_median": float(np.median(kkt)), "kkt_nan": int(np.isnan(kkt).sum())}
    # CPU oracle on the host cores, a bounded sample
    import multiprocessing
    ncpu = multiprocessing.cpu_count()
    o = pr.make_oracle(fb_py)
    o.set_threads(ncpu)
    t0 = time.perf_counter()
    n = 20
    for _ in range(n):
        o.update_solution(0.0, pr.q0, pr.v0)
    out["cpu_oracle_ms_per_instance_iteration"] = (time.perf_counter() - t0) / n * 1e3
    out["cpu_threads"] = ncpu
    print(json.dumps(out))


if __name__ == "__main__":
    main()
