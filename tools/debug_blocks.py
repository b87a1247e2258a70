"""GPU-side debugging aid: per-block activations of the CUDA path against the oracle."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bugcar_image_segmentation_b200 import synth, weights as W, _lib
from bugcar_image_segmentation_b200.weights import ENET_BLOCKS
from oracle import pre_oracle, enet_oracle

prec = sys.argv[1] if len(sys.argv) > 1 else "fp32"
tc = int(sys.argv[2]) if len(sys.argv) > 2 else 1
wpath = os.path.join(ROOT, "pretrained_models", "enet_synthetic_seed42.bcw")
blob = open(wpath, "rb").read()
w, nc, eps = W.unpack_flat(blob)
frames = synth.frames(2, 1234)
x = np.ascontiguousarray(np.concatenate([pre_oracle.preprocess(f) for f in frames]), dtype=np.float32)
lg, inter = enet_oracle.forward(w, x, eps, emulate=None if prec == "fp32" else "bf16", return_intermediates=True)
ctx = _lib.Context(0, 4)
ctx.load_enet(blob)
ctx.set_precision(_lib.BC_PREC_FP32 if prec == "fp32" else _lib.BC_PREC_BF16)
ctx.set_tensor_cores(tc)
dx = torch.from_numpy(x).cuda()
names = ["initial_block"] + [b[0] for b in ENET_BLOCKS]
for i, name in enumerate(names):
    want = inter[name]
    out = torch.empty(want.shape, dtype=torch.float32, device="cuda")
    ctx.enet_block_output(dx, _lib.BC_IN_NCHW_F32, 2, i - 1, out)
    got = out.cpu().numpy()
    d = np.abs(got - want)
    print(f"{name:16s} shape {want.shape} max|ref| {np.abs(want).max():8.3f} maxerr {d.max():9.5f} meanerr {d.mean():10.7f} "
          f"frac>1e-3 {(d > 1e-3 * np.abs(want).max()).mean():.5f}")
