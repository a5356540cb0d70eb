"""Whole-net checks of the oracle: (1) against an INDEPENDENT float64 PyTorch model of the same
prototxt (F.conv2d / conv_transpose2d / max_pool2d(ceil_mode)) so a mis-reading shared with the
im2col+GEMM restatement cannot hide; (2) against the committed golden vectors, which are outputs of the
reference's own CPU code (tests/golden/make_golden.py); tests/test_oracle_ref.py runs the live comparison."""
import os

import numpy as np

import dcutil
import netutil
import torch_model
from oracle import prototxt as opt

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tiny_net.npz")


def test_oracle_matches_independent_fp64_model(tmp_path):
    path, weights = netutil.build(tmp_path, (1, 2, 2, 1), 96, 80)
    x = dcutil.synth.images(2, 96, 80, seed=3)
    got = netutil.oracle_forward(path, weights, x)
    want = torch_model.forward(opt.parse_file(path), weights, x)
    assert sorted(got) == ["loc_pred", "next_pred", "prob"]
    for k in got:
        assert got[k].shape == want[k].shape
        assert netutil.max_err(got[k], want[k]) < 2e-5, k
    assert 0.02 < got["prob"].min() and got["prob"].max() < 0.98      # calibrated weights: unsaturated scoremaps


def test_oracle_reproduces_golden_vectors(tmp_path):
    g = np.load(GOLDEN)
    path, weights = netutil.build(tmp_path, (1, 1, 1, 1), 64, 64)
    x = dcutil.synth.images(2, 64, 64, seed=7)
    out = netutil.oracle_forward(path, weights, x, want={"res2a_relu", "res5a_relu"})
    out["res2a_relu"] = out["res2a_relu"][:, ::16]
    for k in g.files:
        assert netutil.max_err(out[k], g[k]) < 5e-5, k      # fp32 summation order (OpenBLAS vs numpy)
