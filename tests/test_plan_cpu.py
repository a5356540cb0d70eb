"""The fused plan's scheduler and arena placement, planned on the host (no GPU): the L2-resident chunked schedule
(dc_engine.cpp PlanSchedule) must keep every tensor a launch reads intact until that launch -- the planner replays its own
schedule against its arena placement (FusedPlan::VerifySchedule) and caffe_net_describe_plan reports the outcome."""
import re

import pytest

import dcutil


def _net(tmp_path, stages, batch, h, w):
    caffe = dcutil.caffe_module()
    caffe.set_mode_cpu()
    path = dcutil.write_prototxt(tmp_path, stages=tuple(stages), height=h, width=w)
    net = caffe.Net(path, caffe.TEST)
    net.blobs["data"].reshape(batch, 3, h, w)
    return caffe, net


def _segments(desc):
    return [(m.group(1), int(m.group(2)), int(m.group(3)), int(m.group(4)))
            for m in re.finditer(r"segment (\S+) \.\. \S+: (\d+) KiB resident per image, (\d+) of (\d+) images per pass", desc)]


def test_resnet152_at_bench_size_is_chunked_per_stage(tmp_path, monkeypatch):
    monkeypatch.delenv("DC_CHUNK_PLAN", raising=False)
    monkeypatch.setenv("DC_L2_CHUNK_MB", "80")
    _, net = _net(tmp_path, (3, 8, 36, 3), 16, 720, 1280)
    desc = net.describe_plan()
    seg = _segments(desc)
    assert [s[0] for s in seg] == ["res2a_branch1", "res3a_branch1", "res4a_branch1", "res5a_branch1"]
    # per-image working set = block input/shortcut (overwritten in place by the block output) + the two narrow intermediates
    assert [s[1] for s in seg] == [86400, 43200, 21600, 43200]
    # res2 does not fit the budget even for one image -> whole batch per pass; the others run 1 / 3 / 1 images per pass
    assert [s[2] for s in seg] == [16, 1, 3, 1]
    assert "block outputs written in place" in desc


@pytest.mark.parametrize("plan", ["0,0,0,0", "1,1,1,1", "16,2,4,2", "3,5,7,15", "2,2"])
@pytest.mark.parametrize("inplace", ["1", "0"])
def test_forced_chunk_plans_are_consistent(tmp_path, monkeypatch, plan, inplace):
    monkeypatch.setenv("DC_CHUNK_PLAN", plan)
    monkeypatch.setenv("DC_INPLACE_RESIDUAL", inplace)
    _, net = _net(tmp_path, (2, 3, 4, 2), 16, 96, 160)
    desc = net.describe_plan()               # raises if VerifySchedule finds a read of reused storage
    forced = [int(v) for v in plan.split(",")]
    for (name, _, chunk, n), f in zip(_segments(desc), forced):
        want = n if f <= 0 or f >= n else -(-n // -(-n // f))      # even passes
        assert chunk == want, (name, chunk, want)
    if inplace == "0":
        assert "  0 block outputs written in place" in desc


def test_ragged_batches_and_single_image(tmp_path, monkeypatch):
    monkeypatch.setenv("DC_CHUNK_PLAN", "2,2,2,2")
    for batch in (1, 2, 3, 5, 7):
        _, net = _net(tmp_path, (1, 2, 2, 1), batch, 64, 80)
        desc = net.describe_plan()
        for _, _, chunk, n in _segments(desc):
            assert n == batch and 1 <= chunk <= max(1, min(batch, 2))


def test_the_schedule_check_has_teeth(tmp_path, monkeypatch):
    # without the liveness widening for tensors that cross a chunked segment, pass k+1 reads inputs pass k overwrote
    monkeypatch.setenv("DC_CHUNK_PLAN", "1,1,1,1")
    monkeypatch.setenv("DC_PLAN_BREAK_LIVENESS", "1")
    caffe, net = _net(tmp_path, (2, 3, 4, 2), 4, 96, 160)
    with pytest.raises(caffe._caffe.CaffeError, match="inconsistent schedule"):
        net.describe_plan()


def test_materialised_plans_do_not_chunk(tmp_path, monkeypatch):
    # with every named blob materialised each conv is followed by a ToBlob copy: no multi-step segment exists
    monkeypatch.setenv("DC_CHUNK_PLAN", "1,1,1,1")
    _, net = _net(tmp_path, (1, 1, 1, 1), 4, 64, 64)
    net.materialize_intermediates(True)
    desc = net.describe_plan()
    assert "launch groups" in desc


def test_skipped_outputs_drop_their_heads_from_the_plan(tmp_path):
    """Net::set_skipped_outputs (pycaffe shim: net.skip_outputs): a caller that never reads next_pred (estimate_pose.py:231 reads
    prob and loc_pred) gets merged head GEMMs of 42 instead of 406 output channels and no finishing step for the skipped blob."""
    caffe, net = _net(tmp_path, (1, 1, 1, 1), 2, 64, 64)
    full = net.describe_plan()
    assert len(re.findall(r"HeadFinish", full)) == 3 and "heads/deconv_gemm" in full
    net.skip_outputs(["next_pred"])
    trimmed = net.describe_plan()
    assert len(re.findall(r"HeadFinish", trimmed)) == 2
    assert "heads/deconv_gemm-next_pred" in trimmed and "next_pred/" not in trimmed.replace("gemm-next_pred", "")
    net.skip_outputs([])
    assert net.describe_plan() == full
    with pytest.raises(caffe._caffe.CaffeError, match="not an output blob"):
        net.skip_outputs(["res5c"])
    with pytest.raises(caffe._caffe.CaffeError, match="every head output"):
        net.skip_outputs(["prob", "loc_pred", "next_pred"])
        net.describe_plan()
