"""Generates tests/golden/tiny_net.npz: outputs of THE REFERENCE ITSELF (its CPU layer code compiled from
/root/reference into oracle/_ref by oracle/build_ref.py) for the (1,1,1,1)-stage DeeperCut topology at
2x3x64x64 with the seeded calibrated weights.  tests/test_oracle_net.py holds the numpy restatement to these
vectors; the GPU tests hold the product to them.
Run from the repo root (needs /root/reference):  python tests/golden/make_golden.py"""
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
        assert netutil.reference_available(), "oracle/_ref is not built"
        # in-place chains: blob "res2a" holds res2a_relu's output after the forward (pycaffe convention)
        r = netutil.reference_forward(path, weights, x, want=["prob", "loc_pred", "next_pred", "res2a", "res5a"])
    out = {"prob": r["prob"], "loc_pred": r["loc_pred"], "next_pred": r["next_pred"],
           "res2a_relu": r["res2a"][:, ::16], "res5a_relu": r["res5a"]}       # keep the fixture small
    np.savez_compressed(os.path.join(HERE, "tiny_net.npz"), **{k: v.astype(np.float32) for k, v in out.items()})
    print({k: (v.shape, float(np.abs(v).mean())) for k, v in out.items()})
