"""bench.py contract on a GPU (-m gpu): the B200 arm prints exactly one JSON line on stdout with every key the driver reads.
A short run of BASELINE configs[1] (one 3x512x512 image); the numbers themselves are not asserted, their shape is."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_b200_arm_prints_one_json_line_with_the_contract_keys():
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--workload", "cfg1", "--steps", "4", "--warmup", "3", "--no-cpu-baseline"],
                       cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout[:2000]                       # nothing else may reach stdout
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
              "data", "config", "clocks", "gpu_launches", "e2e", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["unit"] == "images/s" and d["n_gpus"] == 1 and d["steps"] == 4 and d["warmup"] >= 3 and d["higher_is_better"] is True
    assert d["scaling"] == "weak" and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert "3x512x512" in d["config"]["workload"] and "model" not in d["config"]
    assert d["value"] > 0 and abs(d["value"] * d["ms_per_step"] / 1e3 - 1.0) < 1e-6          # batch 1: images/s x s/step == 1
    assert d["gpu_launches"] == 4 * 162                                                    # every launch of the timed region is ours
    e = d["e2e"]
    assert e["unit"] == "images/s" and e["value"] > 0 and e["h2d_bytes_per_step"] == 3 * 512 * 512 * 4 and e["d2h_bytes_per_step"] == (14 + 28) * 64 * 64 * 4
    rf = d["roofline"]
    assert rf["bound"] == "tensor" and rf["unit"] == "TFLOP/s" and 0 < rf["frac"] < 1 and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-9
    assert "traffic" in rf
    assert isinstance(d["clocks"], dict) and "reasons" in d["clocks"]
