#!/usr/bin/env python
"""BASELINE.json configs[4]: multi-scale pyramid (0.5 / 1.0 / 1.5x) over 720p, batch 8, on N GPUs of one box.

  python tools/pyramid_bench.py --steps 5                                   (1 GPU: all 24 work items on it)
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/pyramid_bench.py --steps 5

A step = one call of pose.pyramid.estimate_poses_pyramid on 8 synthetic 720p images (rank 0 owns them; NCCL broadcast of the
uint8 batch, LPT assignment of the 24 (image, scale) items, batched forwards per geometry, all-gather of the poses, best of
scales per image) -- end to end, host images in, host poses out.  Prints ONE JSON line (rank 0): images/s and items/s."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "deepcut-cnn_b200", "python"))

_RESULT_FD = os.dup(1)
os.dup2(2, 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--images", type=int, default=8)
    ap.add_argument("--height", type=int, default=720)
    ap.add_argument("--width", type=int, default=1280)
    ap.add_argument("--scales", default="0.5,1.0,1.5")
    ap.add_argument("--model", default="152")
    args = ap.parse_args()
    import importlib
    import numpy as np
    import torch
    import caffe
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    caffe.set_mode_gpu()
    caffe.set_device(local_rank)
    from pose import pyramid
    gen = importlib.import_module("deepcut-cnn_b200.gen_prototxt")
    synth = importlib.import_module("deepcut-cnn_b200.synth")
    ptx = importlib.import_module("deepcut-cnn_b200.prototxt")
    d = os.path.join(ROOT, "models", "_gen")
    os.makedirs(d, exist_ok=True)
    path = os.path.join(d, "pyramid_resnet%s_r%d.prototxt" % (args.model, rank))
    gen.write(path, stages=gen.STAGES_152 if args.model == "152" else gen.STAGES_101)
    weights = synth.calibrated_weights(ptx.parse_file(path))
    scales = tuple(float(s) for s in args.scales.split(","))
    images = list(synth.images_u8(args.images, args.height, args.width, seed=4)) if rank == 0 else None

    def step():
        return pyramid.estimate_poses_pyramid(images, path, None, scales=scales, weights=weights, dist=dist)

    for _ in range(max(args.warmup, 1)):
        best, poses, items = step()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    t0 = time.time()
    for _ in range(args.steps):
        best, poses, items = step()
    caffe.sync()
    wall = time.time() - t0
    if dist is not None:
        t = torch.tensor([wall], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        wall = float(t[0])
    if rank == 0:
        dmod = importlib.import_module("deepcut-cnn_b200.dist")
        bins = dmod.lpt_assign([c for _, _, c in items], world)
        loads = [sum(items[k][2] for k in b) for b in bins]
        line = {"metric": "multi-scale pose images/sec (configs[4])", "value": args.images * args.steps / wall, "unit": "images/s",
                "items_per_s": len(items) * args.steps / wall, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 1),
                "ms_per_step": wall / args.steps * 1e3, "higher_is_better": True, "data": "synthetic",
                "config": {"workload": "DeeperCut ResNet-%s, %d images 3x%dx%d x scales %s, best of scales per image" %
                                       (args.model, args.images, args.height, args.width, args.scales),
                           "assignment": "longest-processing-time-first over (image, scale) items, cost = input pixels",
                           "load_imbalance": max(loads) / (sum(loads) / len(loads))},
                "poses_found": int(sum(b is not None for b in best)),
                "h2d_bytes_per_step": int(args.images * args.height * args.width * 3), "d2h_bytes_per_step": int(len(items) * 280)}
        os.write(_RESULT_FD, (json.dumps(line) + "\n").encode())
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
