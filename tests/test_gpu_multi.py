"""Two ranks on two GPUs (skipped on a 1-GPU box): frame-batch sharding with the grids stored by the
occupancy-grid kernel straight into rank 0's peer-mapped buffer (sharding.PeerGather) gives the same
bytes as the NCCL gather and as a 1-rank run over the concatenated batch."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
sys.path.insert(0, %(root)r)
import numpy as np, torch, torch.distributed as dist
from bugcar_image_segmentation_b200 import synth, sharding
from bugcar_image_segmentation_b200.models import ENET
from bugcar_image_segmentation_b200.bev import bev_transform_tools
from bugcar_image_segmentation_b200.pipeline import FramePipeline
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
B = 4
model = ENET(os.path.join(%(root)r, "pretrained_models", "enet_synthetic_trained.bcw"), device=local, max_batch=world * B)
c = synth.calibration("A")
bev = bev_transform_tools(c["input image size"], c["output image size"], c["distance to target"], c["tile_length"],
                          c["cm_per_px"], c["yaw"], c["is_laserscan"])
bev._bev_matrix = np.asarray(c["bev matrix"]).reshape(3, 3)
pipe = FramePipeline(model, bev, 10.0, 10.0, 0.1)
frames = np.stack([synth.region_frame(s)[0] for s in sharding.frame_seeds(rank, B)])
d = torch.from_numpy(frames).cuda()
via_nccl = sharding.gather_grids(pipe.run_device(d), rank, world)
peer = sharding.PeerGather(model.ctx, rank, world, B, (pipe.Hc, pipe.Wc), local)
for i in range(3):                       # alternate the two buffers
    peer.use(i)
    pipe.run_device(d, to_gather=True)
    via_peer = peer.ready()
    torch.cuda.synchronize()
    if rank == 0:
        assert torch.equal(via_peer, via_nccl), i
peer.close()
# the library's own multi-GPU entry point: bc_pipeline_host_submit / _wait with bc_gather_stream_setup
# (pinned host frames in on every rank, gathered grids on rank 0's host, flags instead of a collective)
sg = sharding.StreamingGather(model.ctx, rank, world, B, (pipe.Hc, pipe.Wc), local)
pin = torch.from_numpy(frames).pin_memory()
outs = [torch.zeros((world * B, pipe.Hc, pipe.Wc), dtype=torch.int8).pin_memory() for _ in range(5)]
for i in range(5):                       # both slots, several generations
    model.ctx.pipeline_host_submit(pin, 256, 512, B, pipe.lut, 10.0, 10.0, 0.1, 0, 0, outs[i] if rank == 0 else None, None)
    model.ctx.pipeline_host_wait(1)
model.ctx.pipeline_host_wait(0)
if rank == 0:
    for i in range(5):
        assert torch.equal(outs[i], via_nccl.cpu()), i
sg.close()
if rank == 0:
    allf = np.stack([synth.region_frame(1234 + i)[0] for i in range(world * B)])
    one = pipe.run_device(torch.from_numpy(allf).cuda())
    assert torch.equal(one, via_nccl)
    print("peer gather ok", tuple(via_nccl.shape))
dist.barrier()
dist.destroy_process_group()
'''


def test_peer_gather_two_gpus(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    script = tmp_path / "w.py"
    script.write_text(WORKER % {"root": ROOT})
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29577", str(script)],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "peer gather ok" in r.stdout


TWO_CTX = r'''
import os, sys
sys.path.insert(0, %(root)r)
import numpy as np, torch
from bugcar_image_segmentation_b200 import synth
from bugcar_image_segmentation_b200.models import ENET
from bugcar_image_segmentation_b200.bev import bev_transform_tools
from bugcar_image_segmentation_b200.pipeline import FramePipeline
w = os.path.join(%(root)r, "pretrained_models", "enet_synthetic_trained.bcw")
c = synth.calibration("A")
frames = np.stack([synth.region_frame(60 + i)[0] for i in range(3)])
res = []
for dev in (0, 1, 0):                       # a context per GPU in ONE process, used alternately
    torch.cuda.set_device(dev)
    m = ENET(w, device=dev, max_batch=4)
    bev = bev_transform_tools(c["input image size"], c["output image size"], c["distance to target"], c["tile_length"],
                              c["cm_per_px"], c["yaw"], c["is_laserscan"])
    bev._bev_matrix = np.asarray(c["bev matrix"]).reshape(3, 3)
    res.append((m, FramePipeline(m, bev, 10.0, 10.0, 0.1)))
out = []
for rep in range(2):
    for (m, p), dev in zip(res, (0, 1, 0)):
        torch.cuda.set_device(dev)
        out.append(p(frames))
for o in out[1:]:
    assert np.array_equal(o, out[0])
print("two devices ok")
'''


def test_two_contexts_on_two_gpus_in_one_process(tmp_path):
    """include/bugcar_b200.h allows one context per GPU in the same process: the > 48 KB dynamic shared-memory
    opt-ins are per-device function attributes, set by bc_create for its own GPU (a process-wide flag would leave
    the second GPU's launches failing)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    script = tmp_path / "two.py"
    script.write_text(TWO_CTX % {"root": ROOT})
    r = subprocess.run([sys.executable, str(script)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "two devices ok" in r.stdout
