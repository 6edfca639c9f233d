"""Bodies of the BASELINE-config parity tests (tests/test_gpu_baseline_configs.py) with the library under test and the
batch size as parameters, so that the emulator suite runs the same code at a tiny size on the GPU-less build box."""
import os

import numpy as np

import anymal_problems as ap
import idocp_b200 as I
from fb_scenarios import compare_batch
from helpers import DIR_FIELDS, SOL_FIELDS, copy_problem, rel_close
from test_emu_fb_parity import DIR, EXP, KKT, RIC, SOL

THREADS = os.cpu_count() or 1


# --------------------------------------------------------------------------------------------------
# configs[2]
# --------------------------------------------------------------------------------------------------
def run_config2(gpu_lib, oracle, B, exact_iters=10, max_iter=150):
    import bench
    prob = I.benchmark_problem(gpu_lib)
    assert bench.SEED == 20240001
    q0, v0 = bench.initial_states(0, B, list(prob.q_min), list(prob.q_max))
    s = I.UnOCPSolver(prob, B, lib=gpu_lib)
    s.setSolution("q", q0)
    s.setSolution("v", v0)
    ob = oracle.Batch(copy_problem(prob, oracle.default_problem()), B)
    for b, o in enumerate(ob.solvers):
        o.set_solution("q", q0[b])
        o.set_solution("v", v0[b])
    # (1) ten iterations, every quantity of every instance bit for bit
    for it in range(exact_iters):
        s.computeKKTResidual(0.0, q0, v0)
        assert np.array_equal(s.KKTError(), ob.kkt_error(0.0, q0, v0, THREADS)), it
        s.updateSolution(0.0, q0, v0)
        ob.update_solution(0.0, q0, v0, False, THREADS)
        for name in DIR_FIELDS:
            assert np.array_equal(s.getDirection(name), ob.get_direction(name)), (it, name)
        p, d = s.getStepSizes()
        st = ob.step_sizes()
        assert np.array_equal(p, st[:, 0]) and np.array_equal(d, st[:, 1]), it
    for name in SOL_FIELDS:
        assert np.array_equal(s.getSolution(name), ob.get_solution(name)), name
    # (2) north_star gate on the full batch: iterate to KKT < 1e-8, compare iteration counts and final trajectories
    tol = 1e-8
    it_gpu = np.full(B, -1)
    it_cpu = np.full(B, -1)
    fin_gpu = {n: np.zeros((B, prob.N + 1 if n == "q" else prob.N, 7)) for n in ("q", "u")}
    fin_cpu = {n: np.zeros_like(fin_gpu[n]) for n in ("q", "u")}
    for it in range(exact_iters, max_iter):
        s.computeKKTResidual(0.0, q0, v0)
        kg = s.KKTError()
        kc = ob.kkt_error(0.0, q0, v0, THREADS)
        ng = (it_gpu < 0) & (kg < tol)
        nc = (it_cpu < 0) & (kc < tol)
        if ng.any():
            for n in ("q", "u"):
                fin_gpu[n][ng] = s.getSolution(n)[ng]
        if nc.any():
            for n in ("q", "u"):
                fin_cpu[n][nc] = ob.get_solution(n)[nc]
        it_gpu[ng] = it
        it_cpu[nc] = it
        if (it_cpu >= 0).all() and (it_gpu >= 0).all():
            break
        s.updateSolution(0.0, q0, v0)
        ob.update_solution(0.0, q0, v0, False, THREADS)
    solved = it_cpu >= 0
    same = solved & (it_gpu == it_cpu)
    within = same.copy()
    for b in np.where(same)[0]:
        within[b] = rel_close(fin_gpu["q"][b], fin_cpu["q"][b]) and rel_close(fin_gpu["u"][b], fin_cpu["u"][b])
    frac = within.sum() / max(solved.sum(), 1)
    print("configs[2]: solved %d / %d, identical iteration count %d, within 1e-9: %d (%.2f %%), iterations median %d max %d"
          % (solved.sum(), B, same.sum(), within.sum(), 100 * frac, np.median(it_cpu[solved]), it_cpu.max()))
    assert solved.sum() >= 0.9 * B
    assert frac >= 0.95
    assert np.array_equal(it_gpu, it_cpu)          # canonical arithmetic: in fact identical for every instance
    assert np.all(s.getStatus() == 0)


def run_config2_tolerance(lib, oracle, B, max_iter=150):
    """A library whose arithmetic is NOT canonical (the -fmad=true build, __graft_entry__.build_fmad_variant) against the
    oracle under the north-star tolerances instead of bit equality: first Newton direction within 1e-9 relative, every
    instance iterated to KKT < 1e-8, iteration counts identical for >= 95 % of the batch and the final trajectories of those
    within 1e-9.  Returns the measured figures (what bit-exactness buys: 100 % by construction)."""
    import bench
    prob = I.benchmark_problem(lib)
    q0, v0 = bench.initial_states(0, B, list(prob.q_min), list(prob.q_max))
    s = I.UnOCPSolver(prob, B, lib=lib)
    s.setSolution("q", q0)
    s.setSolution("v", v0)
    ob = oracle.Batch(copy_problem(prob, oracle.default_problem()), B)
    for b, o in enumerate(ob.solvers):
        o.set_solution("q", q0[b])
        o.set_solution("v", v0[b])
    tol = 1e-8
    it_gpu, it_cpu = np.full(B, -1), np.full(B, -1)
    fin_gpu = {n: np.zeros((B, prob.N + 1 if n == "q" else prob.N, 7)) for n in ("q", "u")}
    fin_cpu = {n: np.zeros_like(fin_gpu[n]) for n in ("q", "u")}
    first_dir_dev, bitwise_equal_dirs = 0.0, True
    for it in range(max_iter):
        s.computeKKTResidual(0.0, q0, v0)
        kg = s.KKTError()
        kc = ob.kkt_error(0.0, q0, v0, THREADS)
        ng = (it_gpu < 0) & (kg < tol)
        nc = (it_cpu < 0) & (kc < tol)
        for n in ("q", "u"):
            if ng.any():
                fin_gpu[n][ng] = s.getSolution(n)[ng]
            if nc.any():
                fin_cpu[n][nc] = ob.get_solution(n)[nc]
        it_gpu[ng] = it
        it_cpu[nc] = it
        if (it_cpu >= 0).all() and (it_gpu >= 0).all():
            break
        s.updateSolution(0.0, q0, v0)
        ob.update_solution(0.0, q0, v0, False, THREADS)
        if it == 0:
            for name in DIR_FIELDS:
                g, c = s.getDirection(name), ob.get_direction(name)
                bitwise_equal_dirs &= bool(np.array_equal(g, c))
                scale = np.maximum(np.max(np.abs(c), axis=tuple(range(1, c.ndim)), keepdims=True), 1e-12)
                first_dir_dev = max(first_dir_dev, float(np.max(np.abs(g - c) / scale)))
    solved = (it_cpu >= 0) & (it_gpu >= 0)
    same = solved & (it_gpu == it_cpu)
    within = same.copy()
    for b in np.where(same)[0]:
        within[b] = rel_close(fin_gpu["q"][b], fin_cpu["q"][b]) and rel_close(fin_gpu["u"][b], fin_cpu["u"][b])
    res = {"batch": int(B), "solved_both": int(solved.sum()), "identical_iteration_count": int(same.sum()),
           "within_1e-9": int(within.sum()), "first_direction_max_rel_dev": first_dir_dev,
           "first_direction_bitwise_equal": bitwise_equal_dirs, "kkt_max_final": float(np.nanmax(kg)),
           "max_iteration_count_difference": int(np.max(np.abs(it_gpu - it_cpu)[solved])) if solved.any() else -1}
    print("configs[2] under tolerances:", res)
    return res


# --------------------------------------------------------------------------------------------------
# configs[3], configs[4]: batched comparison of every field of every chain element
# --------------------------------------------------------------------------------------------------
def _anymal_states(pr, count, seed):
    from idocp_b200 import problems as P
    return P.anymal_initial_states(0, count, q_nominal=pr.q0, seed=seed)


def run_config3(gpu_lib, fb, B, iters=25, full_iters=5):
    pr = ap.TrottingProblem()
    q0, v0 = _anymal_states(pr, B, 20240004)
    solver = ap.make_product_solver(pr, gpu_lib, fb, batch=B, q0=q0, v0=v0)
    oracles = [pr.make_oracle(fb, q0=q0[b], v0=v0[b]) for b in range(B)]
    hist_g, hist_c = [], []
    for it in range(iters):
        solver.computeKKTResidual(0.0, q0, v0)
        hist_g.append(solver.KKTError())
        hist_c.append(fb.batch_kkt(oracles, 0.0, q0, v0, THREADS))
        assert np.array_equal(hist_g[-1], hist_c[-1]), it
        solver.updateSolution(0.0, q0, v0)
        fb.batch_update_solution(oracles, 0.0, q0, v0, False, THREADS)
        st = np.array([o.step_sizes() for o in oracles])
        assert np.array_equal(solver.stepSizes(), st), it
        if it < full_iters:
            for names in (KKT + EXP, RIC, DIR, SOL):
                assert compare_batch(oracles, solver, fb, names) == [], it
    assert compare_batch(oracles, solver, fb, SOL) == []
    last = np.array(hist_g[-1])
    print("configs[3]: KKT after 24 iterations: median %.3e max %.3e, converged (<1e-8): %d / %d"
          % (np.median(last), last.max(), (last < 1e-8).sum(), B))
    assert np.all(np.isfinite(last))
    if iters >= 25:
        assert (last < 1e-6).mean() >= 0.95


def run_config4(gpu_lib, fb, line_search, B, iters=None):
    pr = ap.RunningProblem(10)
    q0, v0 = _anymal_states(pr, B, 20240005)
    solver = ap.make_product_solver(pr, gpu_lib, fb, batch=B, q0=q0, v0=v0)
    oracles = [pr.make_oracle(fb, q0=q0[b], v0=v0[b]) for b in range(B)]
    ch = solver.chain()
    kinds = [c["kind"] for c in ch]
    assert len(ch) == 307 and kinds.count(fb.K_IMPULSE) == 26 and kinds.count(fb.K_LIFT) == 14 and pr.N == 240
    assert [(c["kind"], c["index"], c["dimf"], c["dimi"]) for c in ch] == \
           [(c["kind"], c["index"], c["dimf"], c["dimi"]) for c in oracles[0].chain()]
    iters = iters or (25 if line_search else 32)
    nan_c = nan_g = None
    for it in range(iters):
        solver.updateSolution(0.0, q0, v0, line_search)
        fb.batch_update_solution(oracles, 0.0, q0, v0, line_search, THREADS)
        st = np.array([o.step_sizes() for o in oracles])
        assert np.array_equal(solver.stepSizes(), st, equal_nan=True), (it, line_search)
        solver.computeKKTResidual(0.0, q0, v0)
        kg = solver.KKTError()
        kc = fb.batch_kkt(oracles, 0.0, q0, v0, THREADS)
        assert np.array_equal(kg, kc, equal_nan=True), (it, line_search, np.nanmax(np.abs(kg - kc)))
        nan_g, nan_c = np.isnan(kg), np.isnan(kc)
        if it in (0, 1, 4, iters - 1):
            live = ~nan_c
            # instances that left the interior (slack < 0 after the 0.05 floor, upstream behaviour) are NaN on both
            # sides at the same iteration; the live ones agree on every field
            bad = []
            for e in (0, 1, len(ch) // 2, len(ch) - 1):
                for nm in ("q", "v", "lmd", "gmm"):
                    x = fb.batch_get(oracles, e, nm)[live]
                    y = np.asarray(solver.get(e, nm))[live]
                    if not np.array_equal(x, y):
                        bad.append((e, nm))
            assert bad == [], (it, bad)
    print("configs[4] line_search=%s: after %d iterations NaN instances %d / %d (oracle: %d), KKT median %.3e"
          % (line_search, iters, nan_g.sum(), B, nan_c.sum(), np.nanmedian(kg)))
    if not line_search and iters >= 32:
        # the shipped example's setting: every instance stays finite and the KKT error falls by orders of magnitude
        assert nan_g.sum() == 0
        assert np.median(kg) < 1.0


