"""bench.py contract checks that need no GPU: the reference arm (`--impl reference`) prints exactly one JSON line with
the keys the driver reads, on rank 0 only, and does no work on other ranks."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(args, env_extra=None):
    env = dict(os.environ)
    env.update(env_extra or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, cwd=ROOT, env=env, stdout=subprocess.PIPE,
                          stderr=subprocess.PIPE, text=True, timeout=600)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    r = run(["--impl", "reference", "--batch", "1", "--height", "96", "--width", "128", "--steps", "1", "--warmup", "1"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "images/s" and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0
    assert d["value"] > 0 and abs(d["ms_per_step"] * d["value"] - 1000.0) < 1.0      # one image per step


def test_reference_arm_bounds_its_sample_to_the_budget():
    # 13 steps of a whole image would blow a (deliberately tiny) budget: the step shrinks to the top 1/d of the image and
    # images/s is scaled by the fraction processed
    r = run(["--impl", "reference", "--batch", "1", "--height", "256", "--width", "128", "--steps", "2", "--warmup", "1", "--ref-budget-s", "0.001"])
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    if d["cpu_baseline"]["kind"] != "reference":
        return                                       # the numpy port has no cost probe: whole images
    f = d["config"]["images_per_step"]
    assert 0 < f < 1 and "top" in d["cpu_baseline"]["sample"]
    assert abs(d["ms_per_step"] * d["value"] - 1000.0 * f) < 1.0


def test_reference_arm_is_silent_on_other_ranks():
    r = run(["--impl", "reference", "--gpus", "2", "--batch", "1", "--height", "64", "--width", "64", "--steps", "1", "--warmup", "0"],
            {"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2", "MASTER_ADDR": "127.0.0.1", "MASTER_PORT": "29533"})
    assert r.returncode == 0, r.stderr[-2000:]
    assert r.stdout.strip() == ""
