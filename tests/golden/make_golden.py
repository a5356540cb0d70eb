"""Generates tests/golden/tiny_net.npz: outputs of the CPU oracle for the (1,1,1,1)-stage DeeperCut
topology at 2x3x64x64 with the seeded calibrated weights.  The reference itself cannot be run in
this container (no protobuf/glog/boost/BLAS dev files), so the vectors come from the oracle, which
tests/test_oracle_*.py pin against the reference's own known-answer tests.
Run from the repo root:  python tests/golden/make_golden.py"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import dcutil  # noqa: E402
import netutil  # noqa: E402

if __name__ == "__main__":
    with tempfile.TemporaryDirectory() as tmp:
        path, weights = netutil.build(tmp, (1, 1, 1, 1), 64, 64)
        x = dcutil.synth.images(2, 64, 64, seed=7)
        out = netutil.oracle_forward(path, weights, x, want={"res2a_relu", "res5a_relu"})
    out["res2a_relu"] = out["res2a_relu"][:, ::16]       # keep the fixture small
    np.savez_compressed(os.path.join(HERE, "tiny_net.npz"), **{k: v.astype(np.float32) for k, v in out.items()})
    print({k: (v.shape, float(np.abs(v).mean())) for k, v in out.items()})
