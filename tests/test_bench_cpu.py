"""bench.py's reference arm (no GPU involved): the unmodified reference on the host cores.  Under torchrun only rank 0 runs
and prints; the other ranks exit 0 without work (the driver launches both arms the same way)."""
import json
import os
import subprocess
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_EXE = os.path.join(REPO, "oracle", "_ref", "ref_mountaincar")


def _run(env_extra, *args):
    env = dict(os.environ, **env_extra)
    return subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", *args],
                          env=env, capture_output=True, text=True, timeout=300)


def test_reference_arm_other_ranks_exit_without_work():
    r = _run({"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"}, "--gpus", "2")
    assert r.returncode == 0 and r.stdout.strip() == ""


@pytest.mark.skipif(not os.path.exists(REF_EXE), reason="oracle/_ref not built (make -C oracle ref where /root/reference exists)")
def test_reference_arm_line_has_the_contract_keys():
    r = _run({"FRIQ_BENCH_REF_SECONDS": "1"}, "--steps", "2", "--warmup", "1")
    assert r.returncode == 0, r.stderr
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "agent_env_steps_per_sec" and d["unit"] == "env-steps/s"
    assert d["higher_is_better"] is True and d["value"] > 0
    cb = d["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]
