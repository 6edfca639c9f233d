"""idocp_b200_create_sharded (batch split into contiguous shards, one device + stream per shard, no collective):
the sharded solver gives, instance by instance, the bits of ONE single-device solver over the whole batch -- uneven split,
both solver kinds, setSolution broadcast / per instance, KKT errors, step sizes, full and per-stage getters.

CPU: the SIMT-emulator build (three shards on the emulated device).  GPU (-m gpu): two shards on cuda:0 (two streams of one
device) and, when the box has them, one shard per visible device."""
import numpy as np
import pytest

import idocp_b200 as I
from conftest import make_states
from idocp_b200.capi import SOLVER_UNOCP, SOLVER_UNPARNMPC


def run_sharded_equals_single(lib, devices, kind, batch=7, iters=3, line_search=False):
    prob = I.benchmark_problem(lib)
    q0, v0 = make_states(batch, 41)
    one = (I.UnOCPSolver if kind == SOLVER_UNOCP else I.UnParNMPCSolver)(prob, batch, lib=lib)
    many = I.ShardedSolver(prob, batch, devices, kind=kind, lib=lib)
    assert many.first[0] == 0 and many.first[-1] == batch and len(many.first) == len(devices) + 1
    assert all(b > a for a, b in zip(many.first, many.first[1:]))
    for s in (one, many):
        s.setSolution("q", q0)
        s.setSolution("v", v0)
        s.setSolution("a", np.full(7, 0.1))     # broadcast form
        if kind == SOLVER_UNPARNMPC:
            s.initBackwardCorrection(0.0)
    for it in range(iters):
        for s in (one, many):
            s.computeKKTResidual(0.0, q0, v0)
        assert np.array_equal(one.KKTError(), many.KKTError()), it
        for s in (one, many):
            s.updateSolution(0.0, q0, v0, line_search)
        p1, d1 = one.getStepSizes()
        p2, d2 = many.getStepSizes()
        assert np.array_equal(p1, p2) and np.array_equal(d1, d2), it
    many.sync()
    for name in ("q", "v", "a", "u", "lmd", "gmm"):
        assert np.array_equal(one.getSolution(name), many.getSolution(name)), name
    assert np.array_equal(one.getStageSolution("u", 0), many.getStageSolution("u", 0))
    assert np.array_equal(one.getStatus(), many.getStatus())
    many.close()


@pytest.mark.parametrize("kind", [SOLVER_UNOCP, SOLVER_UNPARNMPC])
def test_sharded_equals_single_emulator(emu_lib, kind):
    run_sharded_equals_single(emu_lib, [0, 0, 0], kind)


def test_sharded_line_search_emulator(emu_lib):
    run_sharded_equals_single(emu_lib, [0, 0], SOLVER_UNOCP, batch=5, iters=2, line_search=True)


def test_sharded_argument_errors(emu_lib):
    prob = I.benchmark_problem(emu_lib)
    with pytest.raises(I.Idocp_b200Error):
        I.ShardedSolver(prob, 2, [0, 0, 0], lib=emu_lib)      # fewer instances than devices
    with pytest.raises(I.Idocp_b200Error):
        I.ShardedSolver(prob, 4, [], lib=emu_lib)
    s = I.ShardedSolver(prob, 4, [0, 0], lib=emu_lib)
    with pytest.raises(ValueError):
        s.updateSolution(0.0, np.zeros((3, 7)), np.zeros((3, 7)))


@pytest.mark.gpu
@pytest.mark.parametrize("kind", [SOLVER_UNOCP, SOLVER_UNPARNMPC])
def test_sharded_equals_single_gpu(gpu_lib, kind):
    import torch
    run_sharded_equals_single(gpu_lib, [0, 0], kind, batch=37, iters=4)
    n = torch.cuda.device_count()
    if n > 1:
        run_sharded_equals_single(gpu_lib, list(range(n)), kind, batch=8 * n + 3, iters=4)


# ---- the hybrid OCPSolver of the floating-base robot over several devices (idocp_b200_fb_create_sharded) ----
def run_fb_sharded_equals_single(lib, devices, batch=5, iters=3):
    import sys
    import os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle"))
    import anymal_problems as ap
    import fb_py
    fb_py.lib()
    pr = ap.with_nonlinear_cones_and_acceleration_limits(ap.TrottingProblem())
    rng = np.random.default_rng(8)
    q0 = np.stack([fb_py.integrate(pr.q0, np.concatenate([rng.uniform(-0.01, 0.01, 3), rng.uniform(-0.02, 0.02, 15)])) for _ in range(batch)])
    v0 = rng.uniform(-0.1, 0.1, (batch, 18))
    one = ap.make_product_solver(pr, lib, fb_py, batch=batch, q0=q0, v0=v0)
    many = ap.make_product_solver(pr, lib, fb_py, batch=batch, q0=q0, v0=v0, devices=devices)
    assert [c["kind"] for c in one.chain()] == [c["kind"] for c in many.chain()]
    for it in range(iters):
        ls = it == iters - 1
        for s in (one, many):
            s.computeKKTResidual(0.0, q0, v0)
        assert np.array_equal(one.KKTError(), many.KKTError(), equal_nan=True), it
        for s in (one, many):
            s.updateSolution(0.0, q0, v0, ls)
        assert np.array_equal(one.stepSizes(), many.stepSizes(), equal_nan=True), it
    many.sync()
    for e in (0, 7, len(one.chain()) - 1):
        for name in ("q", "v", "lmd", "gmm") + (("a", "u", "f", "slack", "dual", "K") if e < len(one.chain()) - 1 else ()):
            assert np.array_equal(one.get(e, name), many.get(e, name), equal_nan=True), (e, name)
    Kq1, Kv1 = one.getStateFeedbackGain(3)
    Kq2, Kv2 = many.getStateFeedbackGain(3)
    assert np.array_equal(Kq1, Kq2) and np.array_equal(Kv1, Kv2)
    assert many.launchCount() > one.launchCount()


def test_fb_sharded_equals_single_emulator(emu_lib):
    run_fb_sharded_equals_single(emu_lib, [0, 0, 0])


@pytest.mark.gpu
def test_fb_sharded_equals_single_gpu(gpu_lib):
    import torch
    run_fb_sharded_equals_single(gpu_lib, [0, 0], batch=9, iters=4)
    n = torch.cuda.device_count()
    if n > 1:
        run_fb_sharded_equals_single(gpu_lib, list(range(n)), batch=2 * n + 1, iters=3)
