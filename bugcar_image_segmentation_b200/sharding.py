"""Frame-batch sharding across GPUs (one process per GPU).  The path is a pure function of
one frame plus constants (models.py:42-69, bev.py:166-246), so ranks share nothing: rank r
of N takes frames [r*B, (r+1)*B); weights and calibration are replicated; the only
exchange is the gather of the int8 grids to rank 0."""
import torch
import torch.distributed as dist


def frame_range(rank, world, per_rank):
    return rank * per_rank, (rank + 1) * per_rank


def frame_seeds(rank, per_rank, seed0=1234):
    """frame i of rank r is seeded seed0 + r*B + i."""
    return range(seed0 + rank * per_rank, seed0 + (rank + 1) * per_rank)


def gather_grids(local, rank, world, backend_device=None):
    """local: int8 (B, ...) grids of this rank -> on rank 0 the (world*B, ...) tensor in
    frame order, None elsewhere.  NCCL on GPUs (gloo in the CPU tests)."""
    if world == 1:
        return local
    out = None
    if rank == 0:
        out = torch.empty((world,) + tuple(local.shape), dtype=local.dtype, device=local.device)
    if dist.get_backend() == "nccl":
        # NCCL has no int8 gather restriction, but gather needs a list on the root
        lst = list(out.unbind(0)) if rank == 0 else None
        dist.gather(local, lst, dst=0)
    else:
        lst = [torch.empty_like(local) for _ in range(world)] if rank == 0 else None
        dist.gather(local, lst, dst=0)
        if rank == 0:
            out = torch.stack(lst)
    return out.reshape((-1,) + tuple(local.shape[1:])) if rank == 0 else None


class PeerGather:
    """Grids gathered WITHOUT a data collective: rank 0 owns one (world*B, Hc, Wc) int8 buffer, every
    other rank maps it into its own address space (CUDA IPC -> peer access over NVLink / NVSwitch) and
    hands the mapping to the library (``bc_gather_setup``), so the occupancy-grid kernel's own stores
    land in rank 0's memory at ``rank * B`` -- the 'gather' is the tail of K9.  What remains is one
    barrier per step (``ready()``) so that rank 0 knows every peer's kernel has finished.

    ``slots`` buffers alternate (``use(i)``) so that step i+1 may write while rank 0 still copies
    step i to the host."""

    def __init__(self, ctx, rank, world, B, grid_shape, device, slots=2):
        import torch
        self.torch, self.ctx, self.rank, self.world, self.B = torch, ctx, rank, world, B
        shape = (world * B,) + tuple(grid_shape)
        self.bufs = []
        for _ in range(slots):
            handle = [None]
            if rank == 0:
                buf = torch.empty(shape, dtype=torch.int8, device=f"cuda:{device}")
                handle = [buf.untyped_storage()._share_cuda_()]
            dist.broadcast_object_list(handle, src=0)
            if rank != 0:
                h = list(handle[0])
                h[0] = device                                   # open the mapping on THIS rank's device
                storage = torch.UntypedStorage._new_shared_cuda(*h)
                buf = torch.empty(0, dtype=torch.int8, device=f"cuda:{device}").set_(storage, 0, shape)
            self.bufs.append(buf)
        self.flag = torch.zeros(1, dtype=torch.int32, device=f"cuda:{device}")
        self.cur = 0
        self.use(0)

    def use(self, i):
        """select the buffer the next pipeline call writes to (d_grids = NULL in that call)"""
        self.cur = i % len(self.bufs)
        self.ctx.gather_setup(self.bufs[self.cur].data_ptr(), self.rank, self.world)

    def ready(self):
        """enqueue the barrier on the current stream; on rank 0 returns the full (world*B, ...) tensor,
        valid for work enqueued on this stream after the call"""
        # a one-element all-reduce is the barrier: stream ordered, it does not block the host
        # (dist.barrier() does), so the next step can be enqueued behind it
        dist.all_reduce(self.flag)
        return self.bufs[self.cur] if self.rank == 0 else None

    def close(self):
        self.ctx.gather_setup(None, 0, 1)


def _share(t):
    return t.untyped_storage()._share_cuda_()


def _open(handle, device, shape, dtype):
    h = list(handle)
    h[0] = device                                   # open the mapping on THIS rank's device
    storage = torch.UntypedStorage._new_shared_cuda(*h)
    return torch.empty(0, dtype=dtype, device=f"cuda:{device}").set_(storage, 0, shape)


class StreamingGather:
    """The multi-GPU form of ``bc_pipeline_host_submit`` / ``_wait`` (``bc_gather_stream_setup``): every rank
    submits ITS frame batches from pinned host memory; the occupancy-grid kernels store into rank 0's gather
    buffer over NVLink, the ranks hand over through flags in peer-mapped device memory, and rank 0's library
    context copies each complete (world*B, Hc, Wc) step to rank 0's host buffer.  torch.distributed is used
    once, here, to exchange the CUDA IPC handles; no collective runs per step.

        sg = StreamingGather(ctx, rank, world, B, (Hc, Wc), device)
        ctx.pipeline_host_submit(pinned_frames, h, w, B, lut, ..., host_grids if rank == 0 else None, stream)
        ctx.pipeline_host_wait(1)
    """

    def __init__(self, ctx, rank, world, B, grid_shape, device):
        self.ctx, self.rank, self.world = ctx, rank, world
        dev = f"cuda:{device}"
        shape = (world * B,) + tuple(grid_shape)
        handles = [None]
        if rank == 0:
            self.gather = [torch.zeros(shape, dtype=torch.int8, device=dev) for _ in range(2)]
            self.arrive = torch.zeros((2, world), dtype=torch.int32, device=dev)
            handles = [[_share(self.gather[0]), _share(self.gather[1]), _share(self.arrive)]]
        dist.broadcast_object_list(handles, src=0)
        if rank != 0:
            g0, g1, ar = handles[0]
            self.gather = [_open(g0, device, shape, torch.int8), _open(g1, device, shape, torch.int8)]
            self.arrive = _open(ar, device, (2, world), torch.int32)
        self.release = torch.zeros(2, dtype=torch.int32, device=dev)          # this rank's own flags
        torch.cuda.synchronize()
        rel = [None] * world
        dist.all_gather_object(rel, _share(self.release) if rank != 0 else None)
        peers = None
        if rank == 0:
            self.release_peers = [self.release] + [_open(rel[r], device, (2,), torch.int32) for r in range(1, world)]
            peers = [t.data_ptr() for t in self.release_peers]
        dist.barrier()
        ctx.gather_stream_setup([g.data_ptr() for g in self.gather], self.arrive.data_ptr(), self.release.data_ptr(),
                                peers, rank, world)

    def close(self):
        torch.cuda.synchronize()
        dist.barrier()
        self.ctx.gather_stream_setup(None, None, None, None, 0, 1)
