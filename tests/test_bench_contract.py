"""bench.py's reference arm (CPU only): `--impl reference` times the reference's own ParticleFilter::updateFilter on the
host and prints ONE JSON line with the contract's keys; under a multi-rank launch only rank 0 prints.  The engine arm
needs a GPU and must refuse to run without one (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

BENCH = os.path.join(ROOT, "bench.py")


def run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, BENCH] + args, capture_output=True, text=True, env=e, timeout=600)


def test_reference_arm_prints_one_contract_line():
    r = run(["--impl", "reference", "--config", "config2", "--steps", "1", "--warmup", "0"])
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "particle_beam_evals_per_sec" and d["unit"] == "evals/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["n_gpus"] == 1 and d["steps"] == 1
    assert d["value"] > 1e6                                             # ~2-3e7 evals/s on one host core
    assert d["config"]["workload"].startswith("config2: 100000 particles x 360 beams")
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] == 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


def test_reference_arm_other_ranks_stay_silent():
    r = run(["--impl", "reference", "--config", "config2", "--steps", "1", "--warmup", "0", "--gpus", "2"],
            env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_engine_arm_refuses_to_run_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = run(["--config", "config2", "--steps", "1"])
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)
