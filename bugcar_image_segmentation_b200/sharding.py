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
