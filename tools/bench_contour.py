"""Time bc_contour_noise_removal at bs 256 (256x512 masks) with CUDA events; prints one JSON line."""
import json
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from bugcar_image_segmentation_b200 import synth, _lib

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
kinds = {"road": [3, 4, 5, 10, 11, 12, 17, 18], "noise": [0, 7, 14, 21], "blocks": [1, 8, 15, 22]}
ctx = _lib.Context(0, 1)
res = {}
for name, seeds in kinds.items():
    base = np.stack([synth.road_mask(s) for s in seeds])
    masks = torch.from_numpy(np.concatenate([base] * (B // len(seeds)))).cuda()
    out = torch.empty_like(masks)
    n = masks.shape[0]
    for _ in range(3):
        ctx.contour_noise_removal(masks, 256, 512, n, out)
    torch.cuda.synchronize()
    ctx.set_profile(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ctx.contour_noise_removal(masks, 256, 512, n, out)
    e1.record()
    torch.cuda.synchronize()
    ctx.set_profile(False)
    ms = e0.elapsed_time(e1) / 10
    res[name] = {"frames": n, "ms": ms, "frames_per_s": n / ms * 1e3, "kept_frac": float(out.float().mean())}
print(json.dumps(res))
