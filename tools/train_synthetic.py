"""Briefly TRAIN the seeded ENet on a synthetic segmentation task (container-side tool).

Why: the reference's trained blobs (pretrained_models/model.h5, enet.pb) are absent
(.MISSING_LARGE_BLOBS).  A random-weight ENet is numerically chaotic -- bf16 rounding
flips max-pool indices on white-noise features and per-pixel argmax margins are tiny --
so "argmax agreement" measured on it says little about a deployed network.  A few
hundred Adam steps on a colour-region task give weights with trained-like statistics
(smooth features, confident interiors, ambiguous boundaries only).

Task: images of random rectangles, each filled with one of 15 palette colours
(+ brightness jitter + N(0,8) pixel noise); the label of a pixel is its palette index,
i.e. the class ids of note_label:1-15 stand for colours.  Output:
pretrained_models/enet_synthetic_trained.bcw (committed; deterministic given torch CPU).

usage: python tools/train_synthetic.py [steps] [batch]
"""
import os
import sys
import time

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bugcar_image_segmentation_b200 import weights as W        # noqa: E402
from bugcar_image_segmentation_b200.synth import PALETTE, region_frame   # noqa: E402
from oracle import enet_oracle, pre_oracle                      # noqa: E402


class TrainNet(enet_oracle._Net):
    """the oracle network with trainable tensors and training-mode batch norm"""

    def __init__(self, weights):
        super().__init__(weights)
        self.params = []
        for k, v in self.w.items():
            if "running_" not in k:
                v.requires_grad_(True)
                self.params.append(v)
        self.training = True

    def bn(self, x, p):
        w = self.w
        return F.batch_norm(x, w[p + ".running_mean"], w[p + ".running_var"], w[p + ".weight"], w[p + ".bias"],
                            self.training, 0.1, self.eps)


def batch(rng, n, h, w):
    xs, ys = [], []
    lut = pre_oracle.normalise_lut().astype(np.float32)
    for _ in range(n):
        img, lab = region_frame(int(rng.integers(1 << 30)), h, w)
        rgb = img[:, :, ::-1]
        xs.append(np.stack([lut[rgb[:, :, c], c] for c in range(3)]))
        ys.append(lab)
    return torch.from_numpy(np.stack(xs)), torch.from_numpy(np.stack(ys).astype(np.int64))


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 400
    bs = int(sys.argv[2]) if len(sys.argv) > 2 else 6
    torch.manual_seed(0)
    torch.set_num_threads(int(os.environ.get("TRAIN_THREADS", "6")))
    rng = np.random.default_rng(2024)
    net = TrainNet(W.synthetic_weights(42, num_classes=15))
    opt = torch.optim.Adam(net.params, lr=2e-3)
    sched = torch.optim.lr_scheduler.OneCycleLR(opt, max_lr=3e-3, total_steps=steps)
    t0 = time.time()
    for it in range(steps):
        h, w = (128, 256) if it < steps * 3 // 4 else (256, 512)     # finish at the deployment resolution
        x, y = batch(rng, bs if h == 128 else max(2, bs // 3), h, w)
        loss = F.cross_entropy(net.forward(x), y)
        opt.zero_grad()
        loss.backward()
        torch.nn.utils.clip_grad_norm_(net.params, 5.0)
        opt.step()
        sched.step()
        if it % 20 == 0 or it == steps - 1:
            print(f"step {it} loss {loss.item():.4f} {time.time() - t0:.0f}s", flush=True)
    # evaluation + export
    net.training = False
    with torch.no_grad():
        x, y = batch(rng, 2, 256, 512)
        acc = (net.forward(x).argmax(1) == y).float().mean().item()
    print("eval pixel accuracy", acc)
    out = {k: v.detach().numpy().astype(np.float32) for k, v in net.w.items()}
    ordered = {n: out[n] for n, _, _ in W.enet_param_spec(15)}
    path = os.path.join(ROOT, "pretrained_models", "enet_synthetic_trained.bcw")
    with open(path, "wb") as f:
        f.write(W.pack_flat(ordered))
    print("wrote", path)


if __name__ == "__main__":
    main()
