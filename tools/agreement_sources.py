"""Where does the argmax disagreement of 16-bit storage come from?  CPU-only experiment on the storage-emulating
oracle (oracle/enet_oracle.py, test infrastructure): the network is evaluated with fp16 rounding at (a) every
storage point, (b) the block outputs only, (c) the block-internal tensors only, (d) everywhere but with the
max-unpool indices of the fp32 network, (e) everywhere but with pooling indices taken from the producing block's
fp32 values; each is compared with the fp32 network's per-pixel argmax.  Result (trained-like weights, 6 scene
frames): ~85 % of the disagreement is max-unpool index flips -- two values of a 2x2 pooling
window that differ by less than an fp16 ulp round to the same number, the first one wins, the fp32 network picks the
other, and the decoder places the feature one pixel away.      usage: python tools/agreement_sources.py [n_frames]"""
import os, sys
import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bugcar_image_segmentation_b200 import synth, weights as W   # noqa: E402
from oracle import pre_oracle, enet_oracle                        # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 6
w, nc, eps = W.unpack_flat(open(os.path.join(ROOT, "pretrained_models", "enet_synthetic_trained.bcw"), "rb").read())
frames = np.stack([synth.region_frame(500 + i)[0] for i in range(n)])
xt = torch.from_numpy(np.concatenate([pre_oracle.preprocess(f) for f in frames]).astype(np.float32))
Net = enet_oracle._Net
ref_idx = {}


class RefNet(Net):                       # fp32 network, remembers its pooling indices
    def down(self, x, name):
        o, i = Net.down(self, x, name)
        ref_idx[name] = i
        return o, i


with torch.no_grad():
    ref = RefNet(w, eps, None).forward(xt).numpy().argmax(1)


def agree(net, label):
    with torch.no_grad():
        a = (net.forward(xt).numpy().argmax(1) == ref).reshape(n, -1).mean(1)
    print(f"{label:58s} mean {a.mean():.5f}  min {a.min():.5f}  max {a.max():.5f}")


class OutputsOnly(Net):                  # (b) only the residual stream is rounded
    def q(self, t):
        return t

    def _r(self, t):
        return Net.q(self, t)

    def initial(self, x):
        return self._r(Net.initial(self, x))

    def regular(self, x, name, dilation=1, asym=False):
        return self._r(Net.regular(self, x, name, dilation, asym))

    def down(self, x, name):
        o, i = Net.down(self, x, name)
        return self._r(o), i

    def up(self, x, name, idx, hw):
        return self._r(Net.up(self, x, name, idx, hw))


class InternalsOnly(Net):                # (c) block outputs stay fp32: the last q() of every block method is undone
    def _wrap(self, method, *a):
        seen, q = [], Net.q
        self.q = lambda t: (seen.append(t), q(self, t))[1]
        r = method(self, *a)
        del self.q
        return (seen[-1],) + r[1:] if isinstance(r, tuple) else seen[-1]

    def initial(self, x):
        return self._wrap(Net.initial, x)

    def regular(self, x, name, dilation=1, asym=False):
        return self._wrap(Net.regular, x, name, dilation, asym)

    def down(self, x, name):
        return self._wrap(Net.down, x, name)

    def up(self, x, name, idx, hw):
        return self._wrap(Net.up, x, name, idx, hw)


def with_ref_indices(only=None):
    class N(Net):                        # (d) fp16 everywhere, unpool positions of the fp32 network
        def down(self, x, name):
            o, i = Net.down(self, x, name)
            return o, (ref_idx[name] if only in (None, name) else i)
    return N


class ProducerIndices(Net):              # (e) pooling indices from the producing block's values before the storage rounding
    def _keep(self, method, *a):
        seen, q = [], Net.q
        self.q = lambda t: (seen.append(t), q(self, t))[1]
        r = method(self, *a)
        del self.q
        self.pre = seen[-1]
        return r

    def initial(self, x):
        return self._keep(Net.initial, x)

    def regular(self, x, name, dilation=1, asym=False):
        return self._keep(Net.regular, x, name, dilation, asym)

    def down(self, x, name):
        o, _ = Net.down(self, x, name)
        return o, F.max_pool2d(self.pre, 2, stride=2, return_indices=True)[1]


agree(Net(w, eps, "fp16"), "(a) fp16 at every storage point")
agree(Net(w, eps, "bf16"), "    bf16 at every storage point")
agree(OutputsOnly(w, eps, "fp16"), "(b) fp16, block outputs only")
agree(InternalsOnly(w, eps, "fp16"), "(c) fp16, block-internal tensors only")
agree(with_ref_indices()(w, eps, "fp16"), "(d) fp16 everywhere, max-unpool indices of the fp32 network")
agree(with_ref_indices("downsample1_0")(w, eps, "fp16"), "    ... for downsample1_0 only")
agree(with_ref_indices("downsample2_0")(w, eps, "fp16"), "    ... for downsample2_0 only")
agree(ProducerIndices(w, eps, "fp16"), "(e) fp16 everywhere, indices from the producer's fp32 values")
