"""DerivativeChecker on the device (idocp_b200_check_cost_derivatives) against what the reference's own cost tests expect
of idocp::DerivativeChecker (test/cost/configuration_space_cost_test.cpp:116-124,181-189; task_space_6d_cost_test.cpp:86-90,
139-143; task_space_3d_cost_test.cpp): first-order checks pass for every component; the second-order check passes for the
configuration-space cost of a fixed-base robot and FAILS for the task-space costs (Gauss-Newton Hessian).  The CPU run uses
the SIMT emulator build of the same kernel sources; the GPU run (-m gpu) is the product library."""
import numpy as np
import pytest

import idocp_b200 as I
from idocp_b200.solvers import _fill, _is_approx


def _random_weights(p, rng, task):
    for name in ("q_weight", "v_weight", "a_weight", "u_weight", "qf_weight", "vf_weight"):
        _fill(getattr(p, name), 0.0 if task else rng.uniform(0.1, 1.0, 7))
    for name in ("q_ref", "v_ref", "u_ref"):
        _fill(getattr(p, name), rng.uniform(-1.0, 1.0, 7))
    p.N, p.T = 2, 1.0   # dt = 0.5
    return p


def config_problem(lib, seed=1):
    return _random_weights(I.benchmark_problem(lib), np.random.default_rng(seed), task=False)


def task_problem(lib, kind, seed=2):
    rng = np.random.default_rng(seed)
    p = _random_weights(I.task_space_problem(lib) if kind == 1 else I.task_space_3d_problem(lib), rng, task=True)
    for k in range(6):
        w = rng.uniform(0.1, 1.0) if (kind == 1 or k < 3) else 0.0
        p.task_q_weight[k] = w
        p.task_qf_weight[k] = 0.5 * w
    return p


def run_checks(lib):
    # ConfigurationSpaceCost, fixed base: everything holds, second order included (the cost is quadratic)
    chk = I.DerivativeChecker(config_problem(lib), lib=lib, samples=6, seed=11)
    assert chk.checkFirstOrderStageCostDerivatives(), chk.last_failure
    assert chk.checkSecondOrderStageCostDerivatives(), chk.last_failure
    assert chk.checkFirstOrderTerminalCostDerivatives(), chk.last_failure
    assert chk.checkSecondOrderTerminalCostDerivatives(), chk.last_failure
    r = chk.evaluate()
    p = config_problem(lib)
    dt = p.T / p.N
    for b in range(chk.samples):   # the numbers themselves: cost and gradient of the quadratic form
        w = lambda name: np.array(getattr(p, name)[:7])
        cost = 0.5 * dt * (np.sum(w("q_weight") * (r["q"][b] - w("q_ref")) ** 2) + np.sum(w("v_weight") * (r["v"][b] - w("v_ref")) ** 2)
                           + np.sum(w("a_weight") * r["a"][b] ** 2) + np.sum(w("u_weight") * (r["u"][b] - w("u_ref")) ** 2))
        assert abs(r["cost"][b] - cost) <= 1e-13 * abs(cost)
        assert np.allclose(r["lq"][b], dt * w("q_weight") * (r["q"][b] - w("q_ref")), rtol=1e-14, atol=0)
        assert np.allclose(r["lu"][b], dt * w("u_weight") * (r["u"][b] - w("u_ref")), rtol=1e-14, atol=0)
    # a step far too large for a forward difference: the check has to notice
    chk.setFiniteDifference(0.3)
    chk.setTestTolerance(1e-6)
    assert not chk.checkFirstOrderStageCostDerivatives()
    assert "lq is not correct" in chk.last_failure
    # TimeVaryingTaskSpace6DCost / TaskSpace6DCost and the 3D pair: gradient exact, Hessian Gauss-Newton
    for kind in (1, 2):
        chk = I.DerivativeChecker(task_problem(lib, kind), lib=lib, samples=6, seed=12 + kind, task_ref=I.task_space_circle_ref, t=0.37)
        chk.setTestTolerance(1.0e-03)
        assert chk.checkFirstOrderStageCostDerivatives(), chk.last_failure
        assert chk.checkFirstOrderTerminalCostDerivatives(), chk.last_failure
        assert not chk.checkSecondOrderStageCostDerivatives()
        assert chk.last_failure.startswith("Qqq is not correct")
        assert not chk.checkSecondOrderTerminalCostDerivatives()
        r = chk.evaluate()
        assert np.all(r["cost"] > 0) and np.all(np.abs(r["lq"]).max(axis=1) > 0)
        for b in range(chk.samples):   # Gauss-Newton: symmetric, positive semidefinite, rank <= 6 (3 for the position cost)
            H = r["Qqq"][b]
            assert np.allclose(H, H.T, rtol=1e-12, atol=1e-14)
            ev = np.linalg.eigvalsh(0.5 * (H + H.T))
            assert ev.min() > -1e-12 * ev.max()
            assert np.sum(ev > 1e-10 * ev.max()) <= (6 if kind == 1 else 3)


def test_is_approx_is_eigens():
    a = np.array([1.0, 2.0, 3.0])
    assert _is_approx(a, a * (1 + 5e-5), 1e-4) and not _is_approx(a, a * (1 + 2e-4), 1e-4)
    assert _is_approx(np.zeros(3), np.zeros(3), 1e-4) and not _is_approx(np.zeros(3), np.full(3, 1e-30), 1e-4)


def test_derivative_checker_emulator(emu_lib):
    run_checks(emu_lib)


def test_derivative_checker_rejects_bad_arguments(emu_lib):
    chk = I.DerivativeChecker(config_problem(emu_lib), lib=emu_lib, samples=2)
    s = chk._solver
    x = np.zeros((2, 7))
    out = np.zeros((2, I.capi.DC_DOUBLES))
    d = I.capi.dptr
    for args in ((2, 0, 2, 1e-8), (0, 99, 2, 1e-8), (0, 0, 0, 1e-8), (0, 0, 2, 0.0)):
        rc = s.lib.L.idocp_b200_check_cost_derivatives(s._h, args[0], args[1], args[2], d(x), d(x), d(x), d(x), args[3], d(out))
        assert rc != 0


@pytest.mark.gpu
def test_derivative_checker_gpu():
    run_checks(I.default_library())


CPP_SOURCE = r"""
#include "idocp/robot/robot.hpp"
#include "idocp/cost/configuration_space_cost.hpp"
#include "idocp/cost/task_space_6d_cost.hpp"
#include "idocp/cost/task_space_3d_cost.hpp"
#include "idocp/utils/derivative_checker.hpp"
#include <iostream>
int main() {
  idocp::Robot robot("");
  // test/cost/configuration_space_cost_test.cpp:116-124,181-189 (fixed base)
  auto cost = std::make_shared<idocp::ConfigurationSpaceCost>(robot);
  Eigen::VectorXd w(7), r(7);
  w << 0.9, 0.3, 0.5, 0.7, 0.2, 0.8, 0.4;
  r << 0.1, -0.4, 0.6, -0.2, 0.3, -0.7, 0.5;
  cost->set_q_weight(w); cost->set_v_weight(w); cost->set_a_weight(w); cost->set_u_weight(w); cost->set_qf_weight(w); cost->set_vf_weight(w);
  cost->set_q_ref(r); cost->set_v_ref(r); cost->set_u_ref(r);
  idocp::DerivativeChecker derivative_checker(robot);
  std::cout << derivative_checker.checkFirstOrderStageCostDerivatives(cost) << derivative_checker.checkSecondOrderStageCostDerivatives(cost)
            << derivative_checker.checkFirstOrderTerminalCostDerivatives(cost) << derivative_checker.checkSecondOrderTerminalCostDerivatives(cost) << " ";
  // test/cost/task_space_6d_cost_test.cpp:86-90,139-143
  auto task = std::make_shared<idocp::TaskSpace6DCost>(robot, 22);
  Eigen::Matrix3d R;
  R << 0, 0, 1, 0, 1, 0, -1, 0, 0;
  task->set_q_6d_ref(Eigen::Vector3d(0.5, 0.1, 0.7), R);
  task->set_q_6d_weight(Eigen::Vector3d(0.8, 0.3, 0.6), Eigen::Vector3d(0.4, 0.9, 0.2));
  task->set_qf_6d_weight(Eigen::Vector3d(0.5, 0.7, 0.1), Eigen::Vector3d(0.6, 0.2, 0.3));
  derivative_checker.setTestTolerance(1.0e-03);
  std::cout << derivative_checker.checkFirstOrderStageCostDerivatives(task) << derivative_checker.checkFirstOrderTerminalCostDerivatives(task) << " ";
  // This is due to Gauss-Newton Hessian approximation.
  std::cout << derivative_checker.checkSecondOrderStageCostDerivatives(task) << derivative_checker.checkSecondOrderTerminalCostDerivatives(task) << " ";
  // test/cost/task_space_3d_cost_test.cpp
  auto task3 = std::make_shared<idocp::TaskSpace3DCost>(robot, 22);
  task3->set_q_3d_ref(Eigen::Vector3d(0.5, 0.1, 0.7));
  task3->set_q_3d_weight(Eigen::Vector3d(0.8, 0.3, 0.6));
  task3->set_qf_3d_weight(Eigen::Vector3d(0.5, 0.7, 0.1));
  std::cout << derivative_checker.checkFirstOrderStageCostDerivatives(task3) << derivative_checker.checkFirstOrderTerminalCostDerivatives(task3) << " ";
  std::cout << derivative_checker.checkSecondOrderStageCostDerivatives(task3) << derivative_checker.checkSecondOrderTerminalCostDerivatives(task3) << std::endl;
  return 0;
}
"""


def _run_cpp(tmp_path, libdir, libname):
    import os
    import subprocess
    from conftest import ROOT
    src = tmp_path / "dc.cpp"
    src.write_text(CPP_SOURCE)
    exe = str(tmp_path / "dc")
    subprocess.check_call(["g++", "-std=c++17", "-Wall", "-I" + os.path.join(ROOT, "include", "idocp_b200", "compat"),
                           "-I" + os.path.join(ROOT, "include"), str(src), "-L" + libdir, "-l:" + libname, "-Wl,-rpath," + libdir, "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout
    import re
    assert len(re.findall(r"Qqq is not correct! \(sample \d+\)\n", out)) == 4   # the reference's message, for the Gauss-Newton blocks
    assert re.sub(r"Qqq is not correct! \(sample \d+\)\n", "", out).split() == ["1111", "11", "00", "11", "00"], out


def test_cpp_derivative_checker_on_the_emulator(emu_lib, tmp_path):
    """The C++ host class, written like the reference's cost tests against the reference's include paths, linked with the SIMT
    emulator build of the kernels (CPU-only container)."""
    import os
    from conftest import ROOT
    _run_cpp(tmp_path, os.path.join(ROOT, "tests", "emu"), "libidocp_b200_emu.so")


@pytest.mark.gpu
def test_cpp_derivative_checker_gpu(tmp_path):
    import os
    from conftest import ROOT
    _run_cpp(tmp_path, os.path.join(ROOT, "idocp_b200"), "libidocp_b200.so")
