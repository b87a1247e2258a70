"""ENet variants (hard part 1): the real blob is absent, so the loader takes the graph and the
variant switches from the weight container ("__graph__" / "__spec__", weights.py) instead of hard-coding one
network.  For every variant: fp32 mode against the torch oracle (logits to round-off), fp16 + tcgen05 against the
fp16-emulating oracle, fused labels == LUT(argmax of the same mode's logits)."""
import os

import numpy as np
import pytest

from bugcar_image_segmentation_b200 import synth, weights as W

pytestmark = pytest.mark.gpu

SHORT = [b for b in W.ENET_BLOCKS if not b[0].startswith(("regular3", "dilated3", "asymmetric3", "dilated2_6", "asymmetric2_7",
                                                         "dilated2_8", "regular1_3", "regular1_4"))]
VARIANTS = {
    "per_channel_prelu_eps1e-3_c19": dict(num_classes=19, gen=dict(prelu_per_channel=True), eps=1e-3),
    "conv_bias_head2x2_c12": dict(num_classes=12, gen=dict(conv_bias=True, head_kernel=2), head_kernel=2),
    "pool2x2_relu_c2": dict(num_classes=2, gen=dict(), encoder_relu=True, initial_pool=2),
    "short_graph_c15": dict(num_classes=15, gen=dict(blocks=SHORT), blocks=SHORT),
}


def _make(name, tmp_path):
    from oracle import enet_oracle, pre_oracle
    v = VARIANTS[name]
    w = W.synthetic_weights(7, v["num_classes"], encoder_relu=v.get("encoder_relu", False), **v["gen"])
    w["__graph__"] = W.graph_rows(v.get("blocks"))
    w["__spec__"] = np.asarray([v.get("initial_pool", 3), v.get("head_kernel", 3), 0, 0], np.float32)
    eps = v.get("eps", W.BN_EPS)
    frames = synth.frames(2, 4321)
    x = np.concatenate([pre_oracle.preprocess(f) for f in frames])
    w = enet_oracle.calibrate_bn(w, x, eps)                       # O(1) activations, like a trained net
    path = tmp_path / f"{name}.bcw"
    path.write_bytes(W.pack_flat(w, v["num_classes"], eps))
    return w, eps, frames, x, str(path)


@pytest.mark.parametrize("name", sorted(VARIANTS))
def test_variant_parity(name, tmp_path):
    from bugcar_image_segmentation_b200 import _lib
    from bugcar_image_segmentation_b200.models import ENET
    from oracle import enet_oracle, pre_oracle
    w, eps, frames, x, path = _make(name, tmp_path)
    want = enet_oracle.forward(w, x, eps)
    m = ENET(path, device=0, max_batch=2, precision="fp32")
    assert m.num_classes == VARIANTS[name]["num_classes"]
    for inp in (x, frames):                                       # float NCHW and uint8 frames
        got = m.logits(inp)
        d = np.abs(got - want) / np.abs(want).max()
        print(f"[{name}] fp32 logits p99 {np.percentile(d, 99):.2e} max {d.max():.2e}")
        assert got.shape == want.shape and np.percentile(d, 99) <= 1e-4, np.percentile(d, 99)
    m.ctx.set_precision(_lib.BC_PREC_FP16)
    want16 = enet_oracle.forward(w, x, eps, emulate="fp16")
    for tc in (0, 1):
        m.ctx.set_tensor_cores(tc)
        got = m.logits(frames)
        scale = np.abs(want16).max()
        tol = 2.0 ** -6 * scale
        s = np.sort(want16, axis=1)
        big = (s[:, -1] - s[:, -2]) > 2 * tol
        conf = (got.argmax(1) == want16.argmax(1))[big].mean()
        frac = (np.abs(got - want16) <= tol).mean()
        print(f"[{name} fp16 tc={tc}] within 2^-6: {frac:.5f}; argmax on confident px ({big.mean():.3f}): {conf:.5f}")
        assert frac >= 0.99 and conf >= 0.995, (frac, conf)
        lab = m.predict(frames)
        ref = pre_oracle.labels_from_logits(got, pre_oracle.LUT_3WAY)
        assert (lab != ref).mean() <= 1e-4


def test_unsupported_variant_is_rejected_with_a_reason(tmp_path):
    """the paper's down-sampling bottleneck with internal width out/4 is not implemented: the loader must say so
    instead of running a different network"""
    from bugcar_image_segmentation_b200 import _lib
    w = W.synthetic_weights(7, 15)
    w["__graph__"] = W.graph_rows(None, down_internal="out/4")
    path = tmp_path / "out4.bcw"
    path.write_bytes(W.pack_flat(w, 15))
    ctx = _lib.Context(0, 1)
    with pytest.raises(_lib.BugcarError) as e:
        ctx.load_enet(path.read_bytes())
    assert "internal width" in str(e.value)
    w["__graph__"] = W.graph_rows(W.ENET_BLOCKS[:-2])            # decoder does not return to 16 channels
    with pytest.raises(_lib.BugcarError):
        ctx.load_enet(W.pack_flat(w, 15))
