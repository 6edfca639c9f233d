"""GPU vs oracle on every BASELINE.json config AT ITS DEFINITION (size, seed, solver settings of SURVEY.md 8d):

  configs[2]  iiwa14 UnOCPSolver, all 16384 splitmix64 (seed 20240001) initial states: 10 iterations compared bit for
              bit on every direction / step size / KKT error / iterate, then iterated to tolerance: the north_star gate
              (>= 95 % identical iteration counts, trajectories within 1e-9) on the FULL batch
  configs[3]  ANYmal trotting OCPSolver, all 4096 seed-20240004 states: 5 iterations, every field of every chain
              element, then the KKT-error history of the example's 25 iterations
  configs[4]  ANYmal running OCPSolver (examples/anymal/anymal_running.cpp:119-229: T = 7, N = 240, 26 impulses +
              14 lifts = 307 stages) with line_search = true on 64 seed-20240005 states, per-iteration step sizes,
              KKT errors (NaN pattern included) and iterates; plus the same problem with line_search = false as in the
              shipped example (anymal_running.cpp:228)
  receding horizon / feedback gains on the GPU: popFront / pushBack between updateSolution calls
  and getStateFeedbackGain against the oracle's K.

The oracle is the checker only; the CUDA library runs through the C-ABI (idocp_b200/*.py marshals)."""
import pytest


pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fb(oracle):
    import fb_py
    fb_py.lib()
    return fb_py


def test_config2_full_batch_bit_exact_and_iteration_gate(gpu_lib, oracle):
    import bench
    import config_scenarios
    assert bench.BATCH_PER_GPU == 16384
    config_scenarios.run_config2(gpu_lib, oracle, bench.BATCH_PER_GPU)


def test_config3_trotting_4096_bit_exact(fb, gpu_lib):
    import config_scenarios
    config_scenarios.run_config3(gpu_lib, fb, 4096)


@pytest.mark.parametrize("line_search", [True, False])
def test_config4_running_full_horizon(fb, gpu_lib, line_search):
    import config_scenarios
    config_scenarios.run_config4(gpu_lib, fb, line_search, 64)


def test_state_feedback_gain_matches_oracle(fb, gpu_lib):
    import fb_scenarios
    fb_scenarios.run_state_feedback_gain(gpu_lib, fb, batch=16)


def test_receding_horizon_bit_exact(fb, gpu_lib):
    import fb_scenarios
    fb_scenarios.run_receding_horizon(gpu_lib, fb, batch=16)


def test_event_before_t_is_an_error(fb, gpu_lib):
    import fb_scenarios
    fb_scenarios.run_event_before_t_is_an_error(gpu_lib, fb)


def test_event_entering_horizon_keeps_constraints(fb, gpu_lib):
    import fb_scenarios
    fb_scenarios.run_event_entering_horizon_keeps_constraints(gpu_lib, fb, batch=8)


def test_batched_mpc_ticks(fb, gpu_lib):
    import fb_scenarios
    fb_scenarios.run_mpc_ticks(gpu_lib, fb, batch=16)
