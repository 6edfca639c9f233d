"""The bodies of the BASELINE-config GPU tests (tests/config_scenarios.py) at a tiny size in the SIMT emulator: checks
the test logic and the kernels' lane-parallel code on the GPU-less build box.  The full sizes run with -m gpu."""
import pytest

import config_scenarios


@pytest.fixture(scope="module")
def fb(oracle):
    import fb_py
    fb_py.lib()
    return fb_py


def test_config2_scenario_emu(emu_lib, oracle):
    config_scenarios.run_config2(emu_lib, oracle, 12, exact_iters=3, max_iter=60)


def test_config3_scenario_emu(emu_lib, fb):
    config_scenarios.run_config3(emu_lib, fb, 2, iters=3, full_iters=1)


def test_config4_scenario_emu(emu_lib, fb):
    # the full 307-stage running horizon with the filter line search, two instances, two iterations
    config_scenarios.run_config4(emu_lib, fb, True, 2, iters=2)
