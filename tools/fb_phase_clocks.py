#!/usr/bin/env python3
"""Per-phase cycle sums of k_fb_condense / k_fb_riccati_backward on the anymal_trotting workload.

Builds a VARIANT library (build/libidocp_b200_phase.so, -DFB_PHASE_CLOCKS: thread 0 of every CTA accumulates clock64()
differences at the phase boundaries) and runs a few iterations with it.  `python tools/fb_phase_clocks.py build` only builds
(done here, without a GPU); without arguments it runs on the GPU box with the prebuilt variant."""
import ctypes as C
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
VARIANT = os.path.join(ROOT, "build", "libidocp_b200_phase.so")

CONDENSE = ["load FbLin", "LLT(M)", "M^-1 solves", "J M^-1, S", "LLT(S), S^-1", "TR / TL / BL", "MJtJinv [dIDC, IDC]",
            "Qafqv / Qafu / laf", "condensed products -> FbKKT", "FbExp store"]
ROBOT = ["load FbSol", "SE(3) pairs, dminus, inverse", "cost gradient, constraint residuals", "state equation, Fqq condensing",
         "forward kinematics", "RNEA + derivatives", "(impulse: second kinematics)", "contact rows", "augment (mat-vec from FbLin)",
         "cost Hessian, slack / dual condensing", "switching constraint"]
RICCATI = ["load FbKKT", "A^T P (6x6 part)", "A^T P, B^T P", "F^T P F (6x6 part)", "F^T P F, Qxu, Quu, lu", "LLT(G)",
           "gain K, k (+ Schur)", "P = Q - K^T G K", "symmetrise, sq, sv", "constrained tail", "store FbRic"]


VARIANT_SIMT = os.path.join(ROOT, "build", "libidocp_b200_phase_simt.so")   # dense products on the SIMT path (A/B of the DMMA form)


def build():
    import __graft_entry__ as g
    os.makedirs(os.path.dirname(VARIANT), exist_ok=True)
    for out, extra in ((VARIANT, []), (VARIANT_SIMT, ["-DIDOCP_FB_MM_SIMT"])):
        cmd = ["/usr/local/cuda/bin/nvcc"] + g.NVCC_FLAGS + ["-DFB_PHASE_CLOCKS"] + extra + ["-o", out, os.path.join(g.CSRC, "capi.cu")]
        subprocess.run(cmd, cwd=g.CSRC, check=True, capture_output=True)


def main():
    global VARIANT
    if len(sys.argv) > 1 and sys.argv[1] == "build":
        build()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "simt":
        VARIANT = VARIANT_SIMT
    import anymal_problems as ap
    import fb_py
    import idocp_b200 as I
    from idocp_b200 import capi
    B = 4096
    lib = capi.Library(VARIANT)
    fn = lib.L.idocp_b200_fb_debug_phase_clocks
    pr = ap.TrottingProblem()
    rng = np.random.default_rng(0)
    q0 = np.tile(pr.q0, (B, 1))
    q0[:, 7:] += rng.uniform(-0.02, 0.02, (B, 12))
    v0 = rng.uniform(-0.1, 0.1, (B, 18))
    solver = ap.make_product_solver(pr, lib, fb_py, batch=B, q0=q0, v0=v0)
    for _ in range(2):
        solver.updateSolution(0.0, q0, v0)
    solver.sync()
    buf = (C.c_ulonglong * 96)()
    fn(buf)
    iters = 5
    for _ in range(iters):
        solver.updateSolution(0.0, q0, v0)
    solver.sync()
    assert fn(buf) == 0
    a = np.array(buf[:], dtype=np.float64).reshape(3, 32)
    n_stage = len(solver.chain())
    out = {}
    for k, (names, units) in enumerate(((CONDENSE, B * (n_stage - 1) * iters), (RICCATI, B * (n_stage - 1) * iters),
                                        (ROBOT, B * (n_stage - 1) * iters))):
        tot = a[k].sum()
        out[("condense", "riccati_backward", "robot")[k]] = {
            "cycles_per_stage": tot / units,
            "phases": {names[i]: {"cycles": a[k][i] / units, "share": a[k][i] / tot} for i in range(len(names))}}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
