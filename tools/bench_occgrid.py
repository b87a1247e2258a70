"""Time bc_occgrid alone at bs 256 on the labels the benchmark network produces (CUDA events)."""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from bugcar_image_segmentation_b200 import synth, _lib

B = 256
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ctx = _lib.Context(0, B)
ctx.load_enet(open(os.path.join(root, "pretrained_models", "enet_synthetic_trained.bcw"), "rb").read())
cal = synth.calibration("A")
ww, wh = cal["output image size"]
ctx.set_bev(cal["bev matrix"], 256, 512, ww, wh, cal["cm_per_px"])
lut = np.full(256, 2, np.uint8); lut[[2, 9]] = 0; lut[[0, 1]] = 1
frames = torch.from_numpy(np.stack([synth.region_frame(1234 + i)[0] for i in range(B)])).cuda()
labels = torch.empty((B, 256, 512), dtype=torch.uint8, device="cuda")
grids = torch.empty((B, 100, 100), dtype=torch.int8, device="cuda")
ctx.pipeline(frames, 256, 512, B, lut, 10.0, 10.0, 0.1, 0, 0, labels, grids)
torch.cuda.synchronize()
ref = grids.clone()
res = {}
for binary in (0, 1):
    for _ in range(3):
        ctx.occgrid(labels, B, 10.0, 10.0, 0.1, binary, 0, grids)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        ctx.occgrid(labels, B, 10.0, 10.0, 0.1, binary, 0, grids)
    e1.record()
    torch.cuda.synchronize()
    res["binary" if binary else "3way"] = e0.elapsed_time(e1) / 20 * 1e3
ctx.occgrid(labels, B, 10.0, 10.0, 0.1, 0, 0, grids)
torch.cuda.synchronize()
res["same_as_pipeline"] = bool(torch.equal(ref, grids))
res["occupied_frac"] = float((grids == 100).float().mean())
print(json.dumps(res))
