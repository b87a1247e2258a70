"""Second opinion on the ENet oracle: the same network executed by OpenCV's dnn engine.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  The reference's own ENet (a frozen
TensorFlow graph, models.py:21-31,43-44) cannot run here, so oracle/enet_oracle.py -- a torch
restatement -- is what the CUDA path is compared with.  To make sure that restatement does not
merely agree with itself, this module exports it to ONNX (legacy TorchScript exporter; the
``onnx`` Python package is not installed, and is only needed for a hook that does nothing for this
model) and runs the file through ``cv2.dnn`` (OpenCV 4.13: its own convolution, pooling,
batch-norm, PReLU and resize kernels, none of torch's).  tests/test_oracle_second_opinion.py asserts
that the two engines produce the same logits to fp32 round-off.

Max-unpool has no ONNX exporter mapping, so the exported graph uses an equivalent formulation without
indices: a 0/1 mask that marks, per 2x2 window, the first position (row-major) holding the window maximum,
times the nearest-upsampled values (`unpool_by_mask`); the test checks that torch evaluates this graph to
exactly the oracle's logits before handing it to OpenCV.
"""
import os
import tempfile

import numpy as np
import torch
import torch.nn.functional as F

from . import enet_oracle
from bugcar_image_segmentation_b200.weights import BN_EPS


def _up2(t):
    return F.interpolate(t, scale_factor=2, mode="nearest")


def unpool_by_mask(v, pre):
    """max_unpool2d(v, argmax indices of max_pool2d(pre, 2, 2)) without indices: the value goes to the FIRST
    window position (row-major) that holds the maximum, as torch's / the CUDA kernels' strict '>' scan does
    (the pooled image channels of the initial block are byte-valued, so equal maxima do occur)."""
    m = F.max_pool2d(pre, 2, stride=2)
    H, W = int(pre.shape[2]), int(pre.shape[3])             # static shapes: plain ints even while tracing
    yy, xx = np.mgrid[0:H, 0:W]                       # NumPy: the position planes enter the graph as constants
    taken = torch.zeros_like(m)
    mask = torch.zeros_like(pre)
    for dy in (0, 1):
        for dx in (0, 1):
            hit = (pre[:, :, dy::2, dx::2] == m).to(v.dtype) * (1.0 - taken)      # first maximum only
            taken = taken + hit
            where = torch.from_numpy(((yy % 2 == dy) & (xx % 2 == dx)).astype(np.float32)).view(1, 1, H, W)
            mask = mask + _up2(hit) * where
    return _up2(v) * mask


class _ExportNet(enet_oracle._Net):
    """enet_oracle._Net with the two index-driven ops in exportable form: `down` hands the
    pre-pool tensor to `up` instead of the argmax indices."""

    def down(self, x, n):
        main = F.max_pool2d(x, 2, stride=2)
        e = self.act(self.conv_bn(x, n + ".ext_conv1.0", n + ".ext_conv1.1", stride=2), n + ".ext_conv1.2")
        e = self.act(self.conv_bn(e, n + ".ext_conv2.0", n + ".ext_conv2.1", padding=1), n + ".ext_conv2.2")
        e = self.act(self.conv_bn(e, n + ".ext_conv3.0", n + ".ext_conv3.1"), n + ".ext_conv3.2")
        pad = torch.zeros(main.shape[0], e.shape[1] - main.shape[1], main.shape[2], main.shape[3])
        out = torch.cat((main, pad), 1) + e
        return self.act(out, n + ".out_activation"), x

    def up(self, x, n, pre, out_hw):
        main = self.conv_bn(x, n + ".main_conv1.0", n + ".main_conv1.1")
        main = unpool_by_mask(main, pre[:, :main.shape[1]] if pre.shape[1] != main.shape[1] else pre)
        e = self.act(self.conv_bn(x, n + ".ext_conv1.0", n + ".ext_conv1.1"), n + ".ext_conv1.2")
        e = self.act(self.conv_bn(e, n + ".ext_tconv1", n + ".ext_tconv1_bnorm", transposed=True, stride=2),
                     n + ".ext_tconv1_activation")
        e = self.conv_bn(e, n + ".ext_conv2.0", n + ".ext_conv2.1")
        return self.act(main + e, n + ".out_activation")


class _Module(torch.nn.Module):
    def __init__(self, weights, bn_eps):
        super().__init__()
        self.net = _ExportNet(weights, bn_eps)

    def forward(self, x):
        return self.net.forward(x)


@torch.no_grad()
def forward_torch_exportable(weights, x, bn_eps=BN_EPS):
    """the exportable formulation evaluated by torch (must equal enet_oracle.forward exactly)"""
    return _Module(weights, bn_eps)(torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))).numpy()


def export_onnx(weights, path, shape=(1, 3, 256, 512), bn_eps=BN_EPS):
    # the exporter's last step inserts onnx-script custom functions and imports `onnx` for it; this model has
    # none, so the step is the identity
    from torch.onnx._internal.torchscript_exporter import onnx_proto_utils
    onnx_proto_utils._add_onnxscript_fn = lambda model_bytes, custom_opsets: model_bytes
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        torch.onnx.export(_Module(weights, bn_eps).eval(), (torch.zeros(shape),), path, opset_version=13, dynamo=False)


def forward_cv2_dnn(weights, x, bn_eps=BN_EPS):
    """fp32 logits (B,C,H,W) computed by OpenCV's dnn engine from the exported ONNX file"""
    import cv2
    x = np.ascontiguousarray(np.asarray(x), dtype=np.float32)
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "enet.onnx")
        export_onnx(weights, path, (1,) + x.shape[1:], bn_eps)
        net = cv2.dnn.readNetFromONNX(path)
        net.setPreferableBackend(cv2.dnn.DNN_BACKEND_OPENCV)
        net.setPreferableTarget(cv2.dnn.DNN_TARGET_CPU)
        outs = []
        for i in range(x.shape[0]):
            net.setInput(x[i:i + 1])
            outs.append(net.forward().copy())
    return np.concatenate(outs)
