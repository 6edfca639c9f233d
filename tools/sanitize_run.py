#!/usr/bin/env python3
"""Small runs of every solver path for compute-sanitizer (memcheck / racecheck / initcheck / synccheck):
   compute-sanitizer --tool racecheck python tools/sanitize_run.py [anymal|iiwa|all]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import idocp_b200 as I  # noqa: E402
from idocp_b200 import problems as P  # noqa: E402
from idocp_b200 import solvers as S  # noqa: E402


def anymal(lib):
    for name, cls, ls in (("trotting", P.AnymalTrotting, False), ("running", P.AnymalRunning, True)):
        pr = cls(lib=lib)
        B = 3
        q0, v0 = P.anymal_initial_states(0, B, q_nominal=pr.q_nominal)
        solver = P.make_solver(pr, B, q0, v0, lib=lib)
        for _ in range(2):
            solver.updateSolution(0.0, q0, v0, ls)
        solver.computeKKTResidual(0.0, q0, v0)
        print("anymal", name, solver.KKTError())


def iiwa(lib):
    rng = np.random.default_rng(1)
    for kind in ("unocp", "unparnmpc"):
        for task in (False, True):
            p = S.task_space_problem(lib, N=20, T=1.0) if task else S.benchmark_problem(lib)
            B = 5
            solver = (I.UnOCPSolver if kind == "unocp" else I.UnParNMPCSolver)(p, B, lib=lib)
            q = np.tile(np.array([0, np.pi / 2, 0, np.pi / 2, 0, np.pi / 2, 0]), (B, 1)) + rng.uniform(-0.1, 0.1, (B, 7))
            v = np.zeros((B, 7))
            solver.setSolution("q", q)
            solver.setSolution("v", v)
            if task:
                solver.setTaskReference(S.task_space_circle_ref, 0.0)
            solver.initConstraints()
            if kind == "unparnmpc":
                solver.initBackwardCorrection(0.0)
            for ls in (False, True):
                solver.updateSolution(0.0, q, v, ls)
            solver.computeKKTResidual(0.0, q, v)
            print("iiwa", kind, "task" if task else "config", solver.KKTError())


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    lib = I.default_library()
    if what in ("anymal", "all"):
        anymal(lib)
    if what in ("iiwa", "all"):
        iiwa(lib)
