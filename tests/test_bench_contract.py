"""bench.py's contract where it can be checked without a GPU: the reference arm's JSON line, the rank rule under torchrun, and the
refusal to run the product arm on a CPU (no fallback)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BENCH = os.path.join(ROOT, "bench.py")


def _run(args, env=None, timeout=300):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, BENCH] + args, capture_output=True, text=True, timeout=timeout, env=e, cwd=ROOT)


def test_reference_arm_prints_one_contract_line():
    # C1 = the reference's own CPU-runnable case (64^3, 256x256): a frame takes milliseconds on the host cores
    r = _run(["--impl", "reference", "--workload", "C1", "--steps", "2", "--warmup", "1"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "Mrays/s" and d["unit"] == "Mrays/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 1 and d["value"] > 0 and d["ms_per_step"] > 0  # exactly the K and W asked for
    assert d["vs_baseline"] is None and d["dtype"] == "f32" and d["data"] == "synthetic"
    assert d["config"]["workload"].startswith("C1") and d["config"]["rays_per_step"] == 256 * 256
    assert set(d["config"]) == {"workload", "pose", "rays_per_step", "grid_bricks", "active_bricks"}  # the keys that name the workload in both arms
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["unit"] == "Mrays/s" and cb["sample"]
    # this container builds oracle/_ref from /root/reference (Makefile): the arm must then run the reference's shader text
    if os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libref_shader.so")):
        assert cb["kind"] == "reference"
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == "Mrays/s" and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0


def test_reference_arm_runs_on_rank_0_only():
    # under torchrun (N > 1) the other ranks exit 0 without work and without output
    r = _run(["--impl", "reference", "--workload", "C1", "--gpus", "2", "--steps", "1", "--warmup", "0"], env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""
    r = _run(["--impl", "reference", "--workload", "C1", "--gpus", "2", "--steps", "1", "--warmup", "0"], env={"RANK": "0", "WORLD_SIZE": "2", "LOCAL_RANK": "0"})
    assert r.returncode == 0
    d = json.loads(r.stdout.strip())
    assert d["impl"] == "reference" and d["n_gpus"] == 2


def test_product_arm_refuses_to_run_without_a_gpu():
    import torch

    if torch.cuda.is_available():
        import pytest

        pytest.skip("a CUDA device is present")
    r = _run(["--steps", "1", "--warmup", "0"])
    assert r.returncode != 0
    assert "no CPU fallback" in (r.stderr + r.stdout)
    assert not any(l.startswith("{") for l in r.stdout.splitlines())  # and prints no result line
