"""bench.py output contract on the CPU box: the reference arm (`--impl reference`, the CPU oracle timed on the host
cores) prints exactly ONE line on stdout, that line is the JSON object the driver parses, and chatter written to
file descriptor 1 by libraries (NCCL's version banner) cannot get in front of it."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None):
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, cwd=ROOT, capture_output=True, text=True,
                         timeout=600, env={**os.environ, **(env or {})})
    assert res.returncode == 0, res.stderr[-2000:]
    return res.stdout


def test_reference_arm_prints_one_json_line():
    out = _run(["--impl", "reference", "--steps", "3", "--warmup", "3"])
    lines = out.splitlines()
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 1 and d["higher_is_better"] is True
    assert d["unit"] == "instance-iterations/s" and d["value"] > 0 and d["dtype"] == "f64"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]


def test_reference_arm_other_ranks_print_nothing():
    out = _run(["--impl", "reference", "--gpus", "2", "--steps", "3", "--warmup", "3"], env={"RANK": "1", "WORLD_SIZE": "2"})
    assert out == ""


def test_stdout_is_claimed_before_any_library_can_write_to_it():
    code = ("import os, sys; sys.path.insert(0, %r); import bench; bench.claim_stdout(); "
            "os.write(1, b'NCCL version 2.28.9+cuda12.9\\n'); print('chatter'); bench.emit_line({'ok': 1})" % ROOT)
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert res.returncode == 0, res.stderr
    assert res.stdout == '{"ok": 1}\n'
    assert "NCCL version" in res.stderr and "chatter" in res.stderr
