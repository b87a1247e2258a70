"""Host-side plumbing shared by the drop-in modules: torch for device memory and
streams, ctypes for the C ABI.  PyTorch is plumbing here, not the product."""
import os

import numpy as np

from . import _lib

DEFAULT_MAX_BATCH = int(os.environ.get("BUGCAR_MAX_BATCH", "256"))
PACKAGE_DIR = os.path.dirname(os.path.abspath(__file__))
REPO_DIR = os.path.dirname(PACKAGE_DIR)
SYNTHETIC_WEIGHTS = os.path.join(REPO_DIR, "pretrained_models", "enet_synthetic_seed42.bcw")


def torch_cuda(device=None):
    """import torch and make sure a CUDA device exists (fail loudly otherwise)."""
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("bugcar_image_segmentation_b200 needs a B200 GPU: torch.cuda.is_available() is "
                           "False and there is no CPU fallback")
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0")) if "LOCAL_RANK" in os.environ else torch.cuda.current_device()
    return torch, int(device)


def stream_handle(torch, device):
    return torch.cuda.current_stream(device).cuda_stream


def to_device_u8(torch, device, a):
    """numpy/torch array -> contiguous uint8 CUDA tensor (no copy when already there)"""
    if isinstance(a, np.ndarray):
        return torch.from_numpy(np.ascontiguousarray(a, dtype=np.uint8)).to(f"cuda:{device}", non_blocking=False)
    if a.dtype != torch.uint8:
        a = a.to(torch.uint8)
    return a.to(f"cuda:{device}").contiguous()


def new_context(device, max_batch=None):
    return _lib.Context(device, DEFAULT_MAX_BATCH if max_batch is None else max_batch)


def gpu_numa_node(device):
    """NUMA node the GPU's PCIe root hangs off (sysfs), or None when the platform does not say."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(device)
        pci = f"{bus.pci_domain_id:04x}:{bus.pci_bus_id:02x}:{bus.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{pci}/numa_node") as f:
            node = int(f.read().strip())
        return node if node >= 0 else None
    except (OSError, ValueError, AttributeError, RuntimeError):
        return None


def bind_host_to_gpu(device):
    """Pin the calling process to the CPUs of the GPU's NUMA node, so that the pinned staging
    buffers allocated AFTERWARDS (first touch) and the submitting thread are local to the GPU: the
    H2D stream of a frame batch then never crosses the socket interconnect.  Returns a dict that
    says what was done ({"node": n, "cpus": k} or {"node": None, ...}); never raises."""
    node = gpu_numa_node(device)
    info = {"node": node, "cpus": None, "bound": False}
    if node is None:
        return info
    try:
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if cpus:
            os.sched_setaffinity(0, cpus)
            info.update(cpus=len(cpus), bound=True)
    except (OSError, ValueError):
        pass
    return info


def device_for_rank(local_rank, world, n_visible=None):
    """Which GPU a rank of an N-rank job on one node should take when the node has MORE GPUs than ranks.
    On the 8 x B200 boxes of this pool GPUs 0-3 and GPUs 4-7 hang off two host domains that each deliver ~ 116 GB/s
    of pinned-host -> device copies (one GPU alone: 55 GB/s; measured by bench.py's h2d ceiling): four ranks on GPUs
    0-3 share one domain (116 GB/s), four ranks on GPUs 0, 4, 1, 5 get 218 GB/s.  The VM exposes neither NUMA nor
    PCIe locality, so the rule is positional: ranks alternate between the two halves of the visible devices.
    BUGCAR_DEVICE_MAP="0,4,1,5" overrides, BUGCAR_DEVICE_MAP="identity" switches the spreading off."""
    import torch
    n = torch.cuda.device_count() if n_visible is None else n_visible
    env = os.environ.get("BUGCAR_DEVICE_MAP", "")
    if env and env != "identity":
        m = [int(v) for v in env.split(",")]
        return m[local_rank % len(m)]
    if env == "identity" or world >= n or n < 2 or n % 2:
        return local_rank
    return (local_rank // 2) + (local_rank % 2) * (n // 2)
