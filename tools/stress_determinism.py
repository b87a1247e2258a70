"""Race hunting: the same 256-frame batch through bc_pipeline N times; labels and grids must be bit-identical
every time (a missed barrier or an overtaken ring slot shows up as run-to-run differences)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bugcar_image_segmentation_b200 import synth, _lib

N = int(sys.argv[1]) if len(sys.argv) > 1 else 40
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
ref_l = ref_g = None
bad = 0
for it in range(N):
    if it % 2:                       # alternate graph replay and plain launches
        ctx.set_graphs(it % 4 == 1)
    ctx.pipeline(frames, 256, 512, B, lut, 10.0, 10.0, 0.1, 0, 0, labels, grids)
    torch.cuda.synchronize()
    if ref_l is None:
        ref_l, ref_g = labels.clone(), grids.clone()
    else:
        dl, dg = int((labels != ref_l).sum()), int((grids != ref_g).sum())
        if dl or dg:
            bad += 1
            print("iteration", it, "differs:", dl, "label pixels,", dg, "grid cells")
print("stress:", N, "iterations,", bad, "differing")
sys.exit(1 if bad else 0)
