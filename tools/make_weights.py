"""Generate pretrained_models/enet_synthetic_seed42.bcw (container-side tool).

The reference's pretrained_models/{model.h5,enet.pb} are absent
(.MISSING_LARGE_BLOBS), so benchmarks and parity tests use this seeded stand-in:
random weights per SURVEY.md 8d (seed 42) whose batch-norm running statistics
are then calibrated on two seeded frames so activations stay O(1), as in a
trained network.  Deterministic given torch's CPU conv kernels; the produced
file is committed so every machine uses identical bytes.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bugcar_image_segmentation_b200 import weights as W   # noqa: E402
from bugcar_image_segmentation_b200 import synth          # noqa: E402
from oracle import enet_oracle, pre_oracle                # noqa: E402


def main():
    w = W.synthetic_weights(42, num_classes=15)
    frames = np.stack([synth.noise_frame(9001), synth.blocky_frame(9002),
                       synth.blocky_frame(9003), synth.noise_frame(9004)])
    x = np.concatenate([pre_oracle.preprocess(f) for f in frames])
    w = enet_oracle.calibrate_bn(w, x)
    out = os.path.join(ROOT, "pretrained_models", "enet_synthetic_seed42.bcw")
    with open(out, "wb") as f:
        f.write(W.pack_flat(w))
    print("wrote", out, os.path.getsize(out))


if __name__ == "__main__":
    main()
