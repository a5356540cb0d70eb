"""Multi-GPU plumbing for batched inference: one process per GPU, torch.distributed (NCCL over
NVLink on the GPU box, gloo in the CPU tests).  The path shards per image -- BatchNorm uses stored
statistics (use_global_stats: true, ResNet-152.prototxt:35-37), so there is no cross-image reduction
and no data-path collective; the only exchange is the batch scatter / result gather the north_star
names, when one rank owns the host batch.  The reference has no inference multi-GPU at all
(python/pose/pose_demo.py:71-74 takes one --gpu).
"""
import numpy as np


def shard_counts(n_items, world):
    """Contiguous balanced split: the first n_items % world ranks get one extra item."""
    base, extra = divmod(int(n_items), int(world))
    return [base + (1 if r < extra else 0) for r in range(world)]


def shard_slices(n_items, world):
    counts = shard_counts(n_items, world)
    out, start = [], 0
    for c in counts:
        out.append(slice(start, start + c))
        start += c
    return out


def lpt_assign(costs, world):
    """Longest-processing-time-first assignment of work items (e.g. the pyramid scales of
    BASELINE configs[4], cost ~ pixels) to ranks; returns per-rank lists of item indices."""
    order = sorted(range(len(costs)), key=lambda i: -costs[i])
    load = [0.0] * world
    bins = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: load[k])
        bins[r].append(i)
        load[r] += costs[i]
    return bins


def scatter_batch(dist, rank, world, full, out, src=0):
    """out (per-rank tensor [cap, ...]) <- rows of `full` (src rank only, [n, ...]); returns the number of
    valid rows this rank received.  Uneven shards are padded to the largest shard (dist.scatter needs
    equal sizes)."""
    import torch
    n = torch.zeros(1, dtype=torch.int64, device=out.device)
    if rank == src:
        n[0] = full.shape[0]
    dist.broadcast(n, src)
    slices = shard_slices(int(n[0]), world)
    cap = out.shape[0]
    assert max(s.stop - s.start for s in slices) <= cap, "per-rank buffer smaller than the largest shard"
    chunks = None
    if rank == src:
        chunks = []
        for s in slices:
            c = torch.zeros_like(out)
            c[: s.stop - s.start] = full[s]
            chunks.append(c)
    dist.scatter(out, chunks, src=src)
    return slices[rank].stop - slices[rank].start


def gather_batch(dist, rank, world, local, n_items, dst=0):
    """Inverse of scatter_batch: rank dst gets the concatenation of every rank's valid rows."""
    import torch
    slices = shard_slices(n_items, world)
    bufs = [torch.zeros_like(local) for _ in range(world)] if rank == dst else None
    dist.gather(local, bufs, dst=dst)
    if rank != dst:
        return None
    return torch.cat([bufs[r][: slices[r].stop - slices[r].start] for r in range(world)], dim=0)


class _DevArray(object):
    """Zero-copy view of a Caffe blob's device memory for torch (__cuda_array_interface__)."""
    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"shape": tuple(int(s) for s in shape), "typestr": "<f4", "data": (int(ptr), False),
                                         "version": 2, "strides": None}


def blob_tensor(blob, mutable=True):
    import torch
    from caffe._caffe import lib, check_ptr
    ptr = check_ptr(lib.caffe_blob_mutable_gpu_data(blob._h) if mutable else lib.caffe_blob_gpu_data(blob._h))
    return torch.as_tensor(_DevArray(ptr, blob.shape), device="cuda")


class BatchExchange(object):
    """bench.py's N > 1 end-to-end step: rank 0 holds the global host batch (pinned), uploads it,
    NCCL-scatters one shard per rank straight into each rank's `data` blob, and after the forward
    NCCL-gathers the requested output blobs back to rank 0 and reads them to the host."""

    def __init__(self, dist, rank, world, shard_shape, out_shapes, x_rank0):
        import torch
        self.dist, self.rank, self.world = dist, rank, world
        self.n_items = shard_shape[0] * world
        self.host = None
        self.full = None
        if rank == 0:
            reps = [np.roll(x_rank0, r, axis=0) for r in range(world)]       # distinct images per shard
            self.host = torch.from_numpy(np.concatenate(reps, axis=0)).pin_memory()
            self.full = torch.empty_like(self.host, device="cuda")
            self.out_host = {k: torch.empty((s[0] * world,) + tuple(s[1:]), dtype=torch.float32).pin_memory() for k, s in out_shapes.items()}
        self.h2d_bytes = int(np.prod(shard_shape)) * 4 * world
        self.d2h_bytes = sum(int(np.prod(s)) * 4 * world for s in out_shapes.values())

    def scatter_into(self, data_blob, L, stream):
        import torch
        import caffe
        if self.rank == 0:
            self.full.copy_(self.host, non_blocking=True)
        scatter_batch(self.dist, self.rank, self.world, self.full, blob_tensor(data_blob))
        torch.cuda.current_stream().synchronize()      # NCCL ran on torch's stream; the forward runs on Caffe's

    def gather_from(self, blobs, L, stream):
        import torch
        import caffe
        caffe.sync()
        for k, b in blobs.items():
            g = gather_batch(self.dist, self.rank, self.world, blob_tensor(b, mutable=False), self.n_items)
            if self.rank == 0:
                self.out_host[k].copy_(g, non_blocking=True)
        torch.cuda.current_stream().synchronize()


class PipelinedExchange(object):
    """The north_star's multi-GPU data path: rank 0 owns the whole host batch; every step it is uploaded once, NCCL-scattered
    over NVLink into every rank, run through Net::Forward there, and the outputs the caller reads (`prob`, `loc_pred`) are
    NCCL-gathered back to rank 0 and copied to its pinned host memory.  (The reference has no multi-GPU inference at all:
    python/pose/pose_demo.py:71-74 takes one --gpu.)

    What travels is the decoded uint8 image (3 B/pixel; 8 x 16 x 720p = 354 MB per step through rank 0's one PCIe link instead
    of 1.4 GB of float input); each rank turns its shard into the float `data` blob on the device (dc_images_u8_to_blob).

    Two streams per rank: the exchange stream (torch side stream; NCCL ops, rank 0's H2D / D2H) and Caffe's forward stream.
    Step k+1's upload + scatter are queued BEFORE step k's forward, so they overlap it; step k's gather waits for its forward
    through an event and overlaps step k+1's.  Buffers are double-buffered by step parity; the collectives are issued in the
    same order on every rank (scatter k+1, gather k).
    """

    def __init__(self, dist, rank, world, net, libdc_mod, shard_n, h, w, out_names, host_u8=None, mean3=(104.0, 117.0, 123.0)):
        import ctypes as C
        import torch
        from caffe._caffe import lib as clib
        self.dist, self.rank, self.world, self.net = dist, rank, world, net
        self.n, self.h, self.w = shard_n, h, w
        self.C, self.L, self.check = C, libdc_mod.lib(), libdc_mod.check
        self.cs_ptr = clib.caffe_stream()
        self.cs = torch.cuda.ExternalStream(self.cs_ptr)
        self.xs = torch.cuda.Stream()
        self.mean = (C.c_float * 3)(*mean3)
        self.recv = [torch.empty((shard_n, h, w, 3), dtype=torch.uint8, device="cuda") for _ in range(2)]
        self.out_names = list(out_names)
        self.out_shapes = {k: tuple(net.blobs[k].shape) for k in self.out_names}
        self.stage = [{k: torch.empty(self.out_shapes[k], dtype=torch.float32, device="cuda") for k in self.out_names} for _ in range(2)]
        self.ev_scat = [torch.cuda.Event() for _ in range(2)]
        self.ev_fwd = [torch.cuda.Event() for _ in range(2)]
        self.ev_gath = [torch.cuda.Event() for _ in range(2)]
        self.full = self.host = self.gathered = self.out_host = None
        if rank == 0:
            assert host_u8 is not None and tuple(host_u8.shape) == (world * shard_n, h, w, 3) and host_u8.dtype == np.uint8
            self.host = torch.from_numpy(host_u8).pin_memory()
            self.full = [torch.empty_like(self.host, device="cuda") for _ in range(2)]
            self.gathered = {k: torch.empty((world,) + self.out_shapes[k], dtype=torch.float32, device="cuda") for k in self.out_names}
            self.out_host = {k: torch.empty((world * self.out_shapes[k][0],) + self.out_shapes[k][1:], dtype=torch.float32).pin_memory()
                             for k in self.out_names}
        self.h2d_bytes = world * shard_n * h * w * 3
        self.d2h_bytes = sum(int(np.prod(s)) * 4 * world for s in self.out_shapes.values())
        self.nvlink_bytes = (world - 1) * (shard_n * h * w * 3 + sum(int(np.prod(s)) * 4 for s in self.out_shapes.values()))

    def submit_scatter(self, k):
        import torch
        b = k & 1
        with torch.cuda.stream(self.xs):
            # recv[b] was read by step k-2's conversion kernel: its forward finished before gather k-2, queued earlier on xs
            chunks = None
            if self.rank == 0:
                self.full[b].copy_(self.host, non_blocking=True)
                chunks = list(self.full[b].chunk(self.world, dim=0))
            self.dist.scatter(self.recv[b], chunks, src=0)
            self.ev_scat[b].record(self.xs)

    def forward(self, k):
        import torch
        b = k & 1
        self.cs.wait_event(self.ev_scat[b])
        self.cs.wait_event(self.ev_gath[b])                 # stage[b] free again: step k-2's gather has read it
        data = self.net.blobs["data"]
        self.check(self.L.dc_images_u8_to_blob(self.C.c_void_p(self.recv[b].data_ptr()), self.n, self.h, self.w, self.mean,
                                                self.C.c_void_p(data.overwrite_gpu_data_ptr()), self.C.c_void_p(self.cs_ptr)))
        self.net.forward()
        with torch.cuda.stream(self.cs):
            for name in self.out_names:
                self.stage[b][name].copy_(blob_tensor(self.net.blobs[name], mutable=False), non_blocking=True)
            self.ev_fwd[b].record(self.cs)

    def submit_gather(self, k):
        import torch
        b = k & 1
        with torch.cuda.stream(self.xs):
            self.xs.wait_event(self.ev_fwd[b])
            for name in self.out_names:
                dst = list(self.gathered[name].unbind(0)) if self.rank == 0 else None
                self.dist.gather(self.stage[b][name], dst, dst=0)
                if self.rank == 0:
                    self.out_host[name].copy_(self.gathered[name].view(self.out_host[name].shape), non_blocking=True)
            self.ev_gath[b].record(self.xs)

    def run(self, steps):
        """`steps` pipelined steps; returns when rank 0's pinned host buffers hold the last step's outputs."""
        self.submit_scatter(0)
        for k in range(steps):
            if k + 1 < steps:
                self.submit_scatter(k + 1)
            self.forward(k)
            self.submit_gather(k)
        self.xs.synchronize()
        self.cs.synchronize()
