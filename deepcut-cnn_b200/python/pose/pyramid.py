"""Multi-scale pose estimation over a batch of images, sharded across the GPUs of one box (BASELINE.json configs[4]:
"multi-scale pyramid (0.5/1.0/1.5x) over 720p, batch 8, 8xB200").

The reference runs the scales of ONE image one after the other on ONE GPU and keeps the scale whose weakest joint is most
confident (python/pose/estimate_pose.py:83-129).  Here the (image, scale) pairs of a whole batch are the work items:

  * cost(item) = pixels of the rescaled net input (the net is fully convolutional: cost is linear in pixels);
    items go to ranks longest-processing-time-first (dist.lpt_assign) -- 8 images x {0.5, 1, 1.5} on 8 GPUs = one item of
    each scale per GPU, 3.5 cost units each;
  * a rank batches its items of one geometry into ONE forward (device pre-processing of each image straight into its slot of
    the `data` blob, one Net::Forward, one dc_pose_from_maps read-out);
  * the only exchange is the batch scatter / result gather: rank 0 owns the decoded images and broadcasts them (NCCL), every
    rank all-gathers 280 bytes of pose per item; best-of-scale is then taken per image exactly like estimate_pose.py:121-126.

With world == 1 (no process group) the same code runs every item on the one GPU.
"""
import ctypes as C
import importlib
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(_HERE))))
import caffe as _caffe  # noqa: E402

from . import estimate_pose as _ep  # noqa: E402

_libdc = importlib.import_module("deepcut-cnn_b200.libdc")
_dist = importlib.import_module("deepcut-cnn_b200.dist")

_NETS = {}      # (model_def, model_bin, batch, out_h, out_w) -> Net
_DEV = {}       # name -> _DeviceBuffer


def _buf(name, nbytes):
    b = _DEV.setdefault(name, _ep._DeviceBuffer())
    return b.reserve(nbytes)


def _net(model_def, model_bin, batch, out_h, out_w, weights):
    key = (model_def, model_bin, batch, out_h, out_w)
    if key not in _NETS:
        net = _caffe.Net(model_def, model_bin, _caffe.TEST) if model_bin else _caffe.Net(model_def, _caffe.TEST)
        if weights is not None:
            net.set_params(weights)
        net.blobs["data"].reshape(batch, 3, out_h, out_w)
        _NETS[key] = net
    return _NETS[key]


def work_items(shapes, scales):
    """[(image index, scale, cost)] for images of the given (h, w) shapes; cost = net-input pixels (estimate_pose.py:83-86)."""
    items = []
    for i, (h, w) in enumerate(shapes):
        for s in scales:
            bh = int(np.ceil(float(h) * s / _ep._STRIDE) * _ep._STRIDE)
            bw = int(np.ceil(float(w) * s / _ep._STRIDE) * _ep._STRIDE)
            items.append((i, float(s), float(bh * bw)))
    return items


def _is_device(img):
    return hasattr(img, "data_ptr")          # a torch CUDA uint8 tensor (what the NCCL broadcast leaves on every rank)


def run_items(images, items, model_def, model_bin, weights=None):
    """Runs (image index, scale) items on THIS rank's GPU; items of one input geometry share a forward.  `images[i]` is a uint8
    HxWx3 numpy array (uploaded once, whatever the number of scales) or a CUDA tensor (used in place).
    -> float32 [len(items), 5, 14] poses in item order."""
    L = _libdc.lib()
    stream = C.c_void_p(_caffe._caffe.lib.caffe_stream())
    out = np.zeros((len(items), 5, 14), np.float32)
    # every distinct image on the device once
    need = sorted({i for i, _, _ in items})
    host = [i for i in need if not _is_device(images[i])]
    dev = {i: int(images[i].data_ptr()) for i in need if _is_device(images[i])}
    keep = []
    if host:
        base = _buf("img", sum(int(np.prod(images[i].shape)) for i in host)).value
        off = 0
        for i in host:
            img = np.ascontiguousarray(images[i], np.uint8)
            keep.append(img)
            _libdc.check(L.dc_memcpy_async(C.c_void_p(base + off), img.ctypes.data_as(C.c_void_p), img.nbytes, 1, stream))
            dev[i] = base + off
            off += img.nbytes
    groups = {}
    for k, (i, s, _) in enumerate(items):
        h, w = images[i].shape[:2]
        groups.setdefault((int(h), int(w), s), []).append(k)
    mean = _ep._MEAN.ctypes.data_as(C.POINTER(C.c_float))
    pending = []
    d_pose_all = _buf("pose", out.nbytes).value
    for (h, w, s), ks in groups.items():
        plan, out_h, out_w, ws = _ep._plan(h, w, s)
        n = len(ks)
        net = _net(model_def, model_bin, n, out_h, out_w, weights)
        base = net.blobs["data"].overwrite_gpu_data_ptr()
        d_ws = _buf("ws", ws) if ws else None
        for slot, k in enumerate(ks):
            _libdc.check(L.dc_preprocess_u8_forward(plan, C.c_void_p(dev[items[k][0]]), mean, C.c_void_p(base + slot * 3 * out_h * out_w * 4), d_ws, stream))
        net.forward()
        prob, loc = net.blobs["prob"], net.blobs["loc_pred"]
        d_pose = d_pose_all + len(pending) * 5 * 14 * 4
        _libdc.check(L.dc_pose_from_maps(prob.gpu_data_ptr(), loc.gpu_data_ptr(), n, 14, prob.shape[2], prob.shape[3], _ep._STRIDE,
                                         _ep._LOCREF_SCALE_MUL, float(s), C.c_void_p(d_pose), stream))
        pending += ks
    # one read-back for all groups (stream-ordered behind every forward), one sync
    got = np.zeros((len(pending), 5, 14), np.float32)
    if len(pending):
        _libdc.check(L.dc_memcpy_async(got.ctypes.data_as(C.c_void_p), C.c_void_p(d_pose_all), got.nbytes, 2, stream))
    _libdc.check(L.dc_stream_sync(stream))
    for slot, k in enumerate(pending):
        out[k] = got[slot]
    del keep
    return out


def best_of_scales(poses, items, n_images):
    """estimate_pose.py:117-126 per image: the scale whose minimum joint confidence is highest (first wins ties; None when no
    scale beats confidence 0, as in the reference)."""
    best = [None] * n_images
    conf = [0.0] * n_images
    for k, (i, _, _) in enumerate(items):
        m = float(poses[k][2].min())
        if m > conf[i]:
            conf[i], best[i] = m, poses[k]
    return best


def estimate_poses_pyramid(images, model_def, model_bin, scales=(0.5, 1.0, 1.5), weights=None, dist=None):
    """images: list of uint8 HxWx3 arrays (channel order of the model; BGR in the demo), held by rank 0 (other ranks may pass
    None when `dist` is an initialised torch.distributed process group: rank 0 broadcasts them).
    -> (list of 5x14 poses, one per image; per-item poses [n_items, 5, 14]; items) on every rank."""
    rank, world = 0, 1
    if dist is not None:
        import torch
        rank, world = dist.get_rank(), dist.get_world_size()
        meta = [None]
        if rank == 0:
            meta[0] = [tuple(im.shape) for im in images]
        dist.broadcast_object_list(meta, src=0)
        shapes = meta[0]
        flat = torch.empty(sum(int(np.prod(s)) for s in shapes), dtype=torch.uint8, device="cuda")
        if rank == 0:
            flat.copy_(torch.from_numpy(np.concatenate([np.ascontiguousarray(im, np.uint8).ravel() for im in images])))
        dist.broadcast(flat, src=0)                                   # the batch scatter (22 MB for 8 x 720p)
        torch.cuda.current_stream().synchronize()                     # NCCL ran on torch's stream; the forwards run on Caffe's
        images, off = [], 0
        for s in shapes:
            n = int(np.prod(s))
            images.append(flat[off:off + n].view(*s))                 # stays on the device: pre-processing reads it in place
            off += n
    items = work_items([im.shape[:2] for im in images], scales)
    bins = _dist.lpt_assign([c for _, _, c in items], world)
    mine = sorted(bins[rank])
    local = run_items(images, [items[k] for k in mine], model_def, model_bin, weights)
    poses = np.zeros((len(items), 5, 14), np.float32)
    if dist is None:
        poses[mine] = local
    else:
        import torch
        cap = max(len(b) for b in bins)
        send = torch.zeros((cap, 5, 14), dtype=torch.float32, device="cuda")
        if len(mine):
            send[:len(mine)] = torch.from_numpy(local).cuda()
        recv = [torch.empty_like(send) for _ in range(world)]
        dist.all_gather(recv, send)                                   # the result gather: 280 B per item
        for r in range(world):
            got = recv[r].cpu().numpy()
            for slot, k in enumerate(sorted(bins[r])):
                poses[k] = got[slot]
    return best_of_scales(poses, items, len(images)), poses, items
