"""Small driver for ncu: a few pipeline steps at a given batch (default one 32-frame chunk)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bugcar_image_segmentation_b200 import synth
from bugcar_image_segmentation_b200.models import ENET
from bugcar_image_segmentation_b200.bev import bev_transform_tools
from bugcar_image_segmentation_b200.pipeline import FramePipeline

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
chunk = int(sys.argv[3]) if len(sys.argv) > 3 else 0
model = ENET(os.path.join(ROOT, "pretrained_models", "enet_synthetic_trained.bcw"), device=0, max_batch=B)
model.ctx.set_graphs(0)
if chunk:
    model.ctx.set_chunk(chunk)
cal = synth.calibration("A")
bev = bev_transform_tools(cal["input image size"], cal["output image size"], cal["distance to target"],
                          cal["tile_length"], cal["cm_per_px"], cal["yaw"], cal["is_laserscan"])
bev._bev_matrix = np.asarray(cal["bev matrix"]).reshape(3, 3)
pipe = FramePipeline(model, bev, 10.0, 10.0, 0.1)
frames = torch.from_numpy(np.tile(np.stack([synth.region_frame(1234 + i)[0] for i in range(8)]), (B // 8, 1, 1, 1))).cuda()
for _ in range(steps):
    g = pipe.run_device(frames)
torch.cuda.synchronize()
print("ok", g.shape, model.ctx.launch_count())
