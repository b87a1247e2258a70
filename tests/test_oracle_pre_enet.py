"""Oracle checks for ENET.preprocess, the argmax+LUT tail and the ENet restatement."""
import hashlib

import numpy as np
import pytest

from conftest import golden
from bugcar_image_segmentation_b200 import synth, weights as W
from oracle import pre_oracle, enet_oracle, refstub


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize("case", ["native", "p720", "odd", "up", "x2"])
def test_preprocess_matches_reference_golden(case):
    g = golden("pre.npz")
    h, w, seed = (int(v) for v in g[case + "_hw_seed"])
    frame = synth.blocky_frame(seed, h, w) if seed % 2 else synth.noise_frame(seed, h, w)
    out = pre_oracle.preprocess(frame)
    assert out.dtype == np.float64 and out.shape == (1, 3, 256, 512)
    assert np.array_equal(out.reshape(-1)[::997], g[case + "_sample"])
    assert _sha(out) == str(g[case + "_sha"])


def test_normalise_lut_is_exact():
    # (rgb/256 - mean)/std through a 256x3 fp64 table is bit-identical to the array expression
    lut = pre_oracle.normalise_lut()
    frame = synth.noise_frame(3)
    out = pre_oracle.preprocess(frame)
    rgb = frame[:, :, ::-1]
    via = np.stack([lut[rgb[:, :, c], c] for c in range(3)])[None]
    assert np.array_equal(out, via)


@pytest.mark.skipif(not refstub.available(), reason="/root/reference only exists in the build container")
def test_preprocess_live_reference():
    ref = refstub.load()
    for h, w, seed in ((256, 512, 31), (480, 640, 32), (1080, 1920, 33)):
        frame = synth.noise_frame(seed, h, w)
        assert np.array_equal(pre_oracle.preprocess(frame), ref.models.ENET.preprocess(frame))
    assert np.array_equal(pre_oracle.IMAGE_MEAN, ref.models.ENET.IMAGE_MEAN)
    assert (ref.models.ENET.INPUT_WIDTH, ref.models.ENET.INPUT_HEIGHT) == (512, 256)


def test_argmax_lut_golden_and_ties():
    g = golden("argmax.npz")
    rng = np.random.default_rng(int(g["seed"]))
    logits = rng.integers(-3, 4, (2, 15, 32, 64)).astype(np.float32)
    assert np.array_equal(pre_oracle.labels_from_logits(logits, pre_oracle.LUT_3WAY), g["labels3"])
    assert np.array_equal(pre_oracle.labels_from_logits(logits, pre_oracle.LUT_BINARY), g["labels2"])
    # tie -> lowest class index (tf.math.argmax / np.argmax): classes 1 and 2 tie => class 1 => label 1
    t = np.zeros((1, 15, 1, 1), np.float32)
    t[0, 1] = t[0, 2] = 5.0
    assert pre_oracle.labels_from_logits(t, pre_oracle.LUT_3WAY)[0, 0, 0] == 1
    assert set(np.unique(pre_oracle.LUT_3WAY)) == {0, 1, 2}
    assert pre_oracle.LUT_3WAY[2] == 0 and pre_oracle.LUT_3WAY[9] == 0 and pre_oracle.LUT_3WAY[0] == 1


def test_param_spec_and_container_roundtrip(synthetic_weights):
    blob, w, nc, eps = synthetic_weights
    assert nc == 15 and abs(eps - 1e-5) < 1e-12
    spec = W.enet_param_spec(15)
    assert [n for n, _, _ in spec] == list(w.keys())
    for n, shape, _ in spec:
        assert w[n].shape == tuple(shape), n
    assert W.pack_flat(w, nc, eps) == blob
    conv_params = sum(v.size for k, v in w.items() if v.ndim == 4)
    assert 330_000 < conv_params < 360_000          # ~342 k conv parameters


def test_enet_oracle_golden_sample(synthetic_weights):
    _, w, nc, eps = synthetic_weights
    g = golden("enet.npz")
    frames = synth.frames(int(g["n"]), int(g["seed0"]))
    x = np.concatenate([pre_oracle.preprocess(f) for f in frames])
    lg = enet_oracle.forward(w, x, eps)
    assert lg.shape == (2, 15, 256, 512) and lg.dtype == np.float32
    s = lg.reshape(-1)[::4099]
    assert np.allclose(s, g["logits_sample"], rtol=1e-3, atol=1e-3 * np.abs(g["logits_sample"]).max())
    hist = np.bincount(lg.argmax(1).reshape(-1), minlength=nc)
    assert (hist > 0).sum() >= 8                     # the synthetic net uses most classes


def test_enet_oracle_structure(synthetic_weights):
    _, w, nc, eps = synthetic_weights
    x = pre_oracle.preprocess(synth.blocky_frame(5))
    lg, inter = enet_oracle.forward(w, x, eps, return_intermediates=True)
    assert inter["initial_block"].shape == (1, 16, 128, 256)
    assert inter["downsample1_0"].shape == (1, 64, 64, 128)
    assert inter["dilated3_7"].shape == (1, 128, 32, 64)
    assert inter["upsample4_0"].shape == (1, 64, 64, 128)
    assert inter["regular5_1"].shape == (1, 16, 128, 256)
    # bf16 emulation stays close to fp32 (error budget of the bf16 storage mode)
    lb = enet_oracle.forward(w, x, eps, emulate="bf16")
    # (the max is dominated by a few max-unpool index flips, so bound the 99th percentile)
    d = np.abs(lb - lg) / np.abs(lg).max()
    assert np.median(d) < 0.02 and np.percentile(d, 99) < 0.1, (np.median(d), np.percentile(d, 99))
