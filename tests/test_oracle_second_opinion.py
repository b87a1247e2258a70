"""The ENet oracle (torch) against an independent engine: the same network exported to ONNX and executed by
OpenCV's dnn module (oracle/enet_onnx.py).  The reference's own network cannot run here (models.py:21-31: the
GraphDef and TensorFlow are absent), so this is the pin that exists: two unrelated CPU implementations of
every layer type agree on the logits to fp32 round-off."""
import os

import numpy as np
import pytest

from bugcar_image_segmentation_b200 import synth, weights as W
from oracle import enet_onnx, enet_oracle, pre_oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _weights(which):
    with open(os.path.join(ROOT, "pretrained_models", f"enet_synthetic_{which}.bcw"), "rb") as f:
        return W.unpack_flat(f.read())


@pytest.mark.parametrize("which", ["seed42", "trained"])
def test_torch_oracle_equals_opencv_dnn(which):
    w, nc, eps = _weights(which)
    frame = synth.region_frame(777)[0] if which == "trained" else synth.noise_frame(1234)
    x = pre_oracle.preprocess(frame).astype(np.float32)
    want = enet_oracle.forward(w, x, eps)
    # the exportable graph (mask-based unpool) is the same function as the oracle (index-based unpool)
    assert np.array_equal(enet_onnx.forward_torch_exportable(w, x, eps), want)
    got = enet_onnx.forward_cv2_dnn(w, x, eps)
    assert got.shape == want.shape == (1, nc, 256, 512)
    d = np.abs(got - want) / np.abs(want).max()
    agree = (got.argmax(1) == want.argmax(1)).mean()
    print(f"[{which}] cv2.dnn vs torch oracle: max |d|/max|logit| {d.max():.2e}, p99.9 {np.percentile(d, 99.9):.2e}, argmax agreement {agree:.6f}")
    # two engines, different summation orders: fp32 round-off, plus the rare max-pool near-tie flip downstream
    assert np.percentile(d, 99.9) <= 1e-4, np.percentile(d, 99.9)
    assert agree >= 0.9995, agree
