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
    for name, cls, ls in (("trotting", P.AnymalTrotting, False), ("running", P.AnymalRunning, True), ("trotting-f3", P.AnymalTrotting, True)):
        pr = cls(lib=lib)
        if name.endswith("f3"):   # FrictionCone / ImpulseFrictionCone + JointAcceleration{Lower,Upper}Limit
            pr.problem.cone_nonlinear[0] = pr.problem.cone_nonlinear[1] = 1
            pr.problem.enable_acceleration_limit[0] = pr.problem.enable_acceleration_limit[1] = 1
            for j in range(12):
                pr.problem.a_min[j], pr.problem.a_max[j] = -9.0, 9.0
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
        for task, acc in ((False, False), (True, False), (False, True), (True, True)):
            p = S.task_space_problem(lib, N=20, T=1.0) if task else S.benchmark_problem(lib)
            if acc:   # JointAcceleration{Lower,Upper}Limit: the ACC kernel instantiations and the XA array
                p.enable_acceleration_limit[0] = p.enable_acceleration_limit[1] = 1
                for j in range(7):
                    p.a_min[j], p.a_max[j] = -25.0, 25.0
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
            print("iiwa", kind, "task" if task else "config", "acc" if acc else "", solver.KKTError())


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    lib = I.default_library()
    if what in ("anymal", "all"):
        anymal(lib)
    if what in ("iiwa", "all"):
        iiwa(lib)
