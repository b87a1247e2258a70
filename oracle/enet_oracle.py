"""torch-CPU fp32 restatement of the ENet forward pass.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  PARITY UNPINNED: the reference
executes a frozen TensorFlow graph (models.py:43-44; tensor names models.py:15-16)
whose definition and weights are absent (.MISSING_LARGE_BLOBS:1-3).  From the
reference we take only the interface: fp32 NCHW input (B,3,256,512) normalised as in
models.py:84-95, NCHW logits (B,C,256,512) (models.py:52).  The structure is the
canonical ENet (Paszke et al. 2016, arXiv 1606.02147 table 1) in its PyTorch form,
from which the Keras model.h5 / enet.pb were evidently converted (pytorch2keras
short names, utils.py:49-83), tabulated in SURVEY.md 8a row 4.

``forward(weights, x)`` is the plain fp32 network with un-folded batch norm.
``forward(..., emulate="bf16")`` additionally folds BN and rounds weights and the
activations the CUDA kernels keep in bf16, to predict the bf16-storage error
budget on CPU; it is a development aid, not the parity reference.
"""
import numpy as np
import torch
import torch.nn.functional as F

from bugcar_image_segmentation_b200.weights import ENET_BLOCKS, BN_EPS, block_name


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))


class _Net:
    def __init__(self, weights, bn_eps=BN_EPS, emulate=None, calibrate=False):
        self.w = {k: _t(v) for k, v in weights.items()}
        # graph and variant switches travel with the weights (weights.py: __graph__ / __spec__)
        spec = weights.get("__spec__")
        self.initial_pool = int(spec[0]) if spec is not None else 3
        self.head_kernel = int(spec[1]) if spec is not None else 3
        g = weights.get("__graph__")
        if g is None:
            self.blocks = [(n, k, (a[1] if k == "reg" else 1)) for n, k, a in ENET_BLOCKS]
        else:
            kinds = {0: "down", 1: "reg", 2: "asym", 3: "up"}
            self.blocks = [(block_name(r[0], r[1], r[2], r[6]), kinds[int(r[0])], int(r[6])) for r in np.asarray(g)]
        self.eps = bn_eps
        self.emulate = emulate            # None | "bf16" | "fp16": the 16-bit storage type to emulate
        self.qdtype = {None: None, "bf16": torch.bfloat16, "fp16": torch.float16}[emulate]
        self.calibrate = calibrate   # tools/make_weights.py: set BN running stats from data

    # ---- helpers
    def q(self, x):
        """activation storage rounding"""
        if self.emulate:
            return x.to(self.qdtype).to(torch.float32)
        return x

    def act(self, x, name):
        a = self.w.get(name + ".weight")
        if a is None:
            return F.relu(x)
        return F.prelu(x, a)

    def bn(self, x, p):
        w = self.w
        if self.calibrate:
            w[p + ".running_mean"] = x.mean(dim=(0, 2, 3)).contiguous()
            w[p + ".running_var"] = x.var(dim=(0, 2, 3), unbiased=False).contiguous()
        return F.batch_norm(x, w[p + ".running_mean"], w[p + ".running_var"],
                            w[p + ".weight"], w[p + ".bias"], False, 0.0, self.eps)

    def conv_bn(self, x, conv, bn, transposed=False, **kw):
        """conv (no bias) followed by eval-mode BN.  In bf16 emulation the BN is
        folded into bf16 weights + fp32 bias, as the CUDA loader does."""
        W = self.w[conv + ".weight"]
        cb = self.w.get(conv + ".bias")                  # optional convolution bias (checkpoints built with bias=True)
        if self.emulate:
            g = self.w[bn + ".weight"] / torch.sqrt(self.w[bn + ".running_var"] + self.eps)
            b = self.w[bn + ".bias"] - self.w[bn + ".running_mean"] * g
            if cb is not None:
                b = b + cb * g
            if transposed:
                Wf = (W * g.view(1, -1, 1, 1)).to(self.qdtype).to(torch.float32)
                return F.conv_transpose2d(x, Wf, b, **kw)
            Wf = (W * g.view(-1, 1, 1, 1)).to(self.qdtype).to(torch.float32)
            return F.conv2d(x, Wf, b, **kw)
        y = F.conv_transpose2d(x, W, cb, **kw) if transposed else F.conv2d(x, W, cb, **kw)
        return self.bn(y, bn)

    # ---- blocks
    def initial(self, x):
        main = F.conv2d(x, self.w["initial_block.main_branch.weight"], self.w.get("initial_block.main_branch.bias"),
                        stride=2, padding=1)
        ext = F.max_pool2d(x, 3, stride=2, padding=1) if self.initial_pool == 3 else F.max_pool2d(x, 2, stride=2)
        out = self.bn(torch.cat((main, ext), 1), "initial_block.batch_norm")
        return self.q(self.act(out, "initial_block.out_activation"))

    def down(self, x, n):
        main, idx = F.max_pool2d(x, 2, stride=2, return_indices=True)
        e = self.q(self.act(self.conv_bn(x, n + ".ext_conv1.0", n + ".ext_conv1.1", stride=2), n + ".ext_conv1.2"))
        e = self.q(self.act(self.conv_bn(e, n + ".ext_conv2.0", n + ".ext_conv2.1", padding=1), n + ".ext_conv2.2"))
        e = self.act(self.conv_bn(e, n + ".ext_conv3.0", n + ".ext_conv3.1"), n + ".ext_conv3.2")
        pad = torch.zeros(main.shape[0], e.shape[1] - main.shape[1], *main.shape[2:])
        out = torch.cat((main, pad), 1) + e
        return self.q(self.act(out, n + ".out_activation")), idx

    def regular(self, x, n, dilation=1, asym=False):
        e = self.q(self.act(self.conv_bn(x, n + ".ext_conv1.0", n + ".ext_conv1.1"), n + ".ext_conv1.2"))
        if asym:
            e = self.q(self.act(self.conv_bn(e, n + ".ext_conv2.0", n + ".ext_conv2.1", padding=(2, 0)), n + ".ext_conv2.2"))
            e = self.q(self.act(self.conv_bn(e, n + ".ext_conv2.3", n + ".ext_conv2.4", padding=(0, 2)), n + ".ext_conv2.5"))
        else:
            e = self.q(self.act(self.conv_bn(e, n + ".ext_conv2.0", n + ".ext_conv2.1",
                                             padding=dilation, dilation=dilation), n + ".ext_conv2.2"))
        e = self.act(self.conv_bn(e, n + ".ext_conv3.0", n + ".ext_conv3.1"), n + ".ext_conv3.2")
        return self.q(self.act(x + e, n + ".out_activation"))

    def up(self, x, n, idx, out_hw):
        main = self.conv_bn(x, n + ".main_conv1.0", n + ".main_conv1.1")
        main = F.max_unpool2d(main, idx, 2, output_size=out_hw)
        e = self.q(self.act(self.conv_bn(x, n + ".ext_conv1.0", n + ".ext_conv1.1"), n + ".ext_conv1.2"))
        e = self.q(self.act(self.conv_bn(e, n + ".ext_tconv1", n + ".ext_tconv1_bnorm", transposed=True, stride=2),
                            n + ".ext_tconv1_activation"))
        e = self.conv_bn(e, n + ".ext_conv2.0", n + ".ext_conv2.1")
        return self.q(self.act(main + e, n + ".out_activation"))

    def forward(self, x, return_intermediates=False):
        inter = {}
        x = self.initial(x)
        inter["initial_block"] = x
        pools = []                                       # (indices, pre-pool size) of the down-sampling blocks, innermost last
        for name, kind, dil in self.blocks:
            if kind == "down":
                size = x.shape[2:]
                x, idx = self.down(x, name)
                pools.append((idx, size))
            elif kind == "reg":
                x = self.regular(x, name, dilation=dil)
            elif kind == "asym":
                x = self.regular(x, name, asym=True)
            elif kind == "up":                           # pairs with the innermost open down-sampling block
                idx, size = pools.pop()
                x = self.up(x, name, idx, size)
            inter[name] = x
        W = self.w["transposed_conv.weight"]
        if self.emulate:
            W = W.to(self.qdtype).to(torch.float32)
        if self.head_kernel == 3:
            logits = F.conv_transpose2d(x, W, None, stride=2, padding=1, output_padding=1)
        else:
            logits = F.conv_transpose2d(x, W, None, stride=2)
        if return_intermediates:
            return logits, inter
        return logits


@torch.no_grad()
def forward(weights, x, bn_eps=BN_EPS, emulate=None, return_intermediates=False):
    """weights: dict name -> float32 array (bugcar_image_segmentation_b200.weights);
    x: float array (B,3,256,512) as returned by ENET.preprocess (any float dtype,
    cast to fp32 exactly as TensorFlow's feed does).  Returns fp32 logits
    (B,C,256,512) as a NumPy array (and the per-block activations on request)."""
    xt = torch.from_numpy(np.ascontiguousarray(np.asarray(x), dtype=np.float32))
    net = _Net(weights, bn_eps, emulate)
    r = net.forward(xt, return_intermediates)
    if return_intermediates:
        return r[0].numpy(), {k: v.numpy() for k, v in r[1].items()}
    return r.numpy()


@torch.no_grad()
def calibrate_bn(weights, x, bn_eps=BN_EPS):
    """Return a copy of ``weights`` whose BN running statistics are the batch
    statistics seen on ``x`` (what training would have left there), so that the
    synthetic network has O(1) activations like a trained one.  Used only by
    tools/make_weights.py to produce pretrained_models/enet_synthetic_seed42.bcw."""
    xt = torch.from_numpy(np.ascontiguousarray(np.asarray(x), dtype=np.float32))
    net = _Net(weights, bn_eps, None, calibrate=True)
    net.forward(xt)
    return {k: v.numpy().astype(np.float32) for k, v in net.w.items()}
