"""GPU parity, pre/post-processing: CUDA kernels (through the C ABI) against the committed
golden vectors produced by the reference's own code, and against the NumPy oracle on
seeded inputs.  Bar: bit-exact (integer / byte / fp64-table work)."""
import hashlib
import json

import numpy as np
import pytest

from conftest import golden
from bugcar_image_segmentation_b200 import synth

pytestmark = pytest.mark.gpu


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def ctx():
    import torch
    from bugcar_image_segmentation_b200 import _lib
    assert torch.cuda.is_available()
    c = _lib.Context(0, 8)
    yield c
    c.close()


def _bev(cal_name):
    from bugcar_image_segmentation_b200.bev import bev_transform_tools
    rows, cols = (720, 1280) if cal_name == "E" else (256, 512)
    c = synth.calibration(cal_name, rows, cols)
    bev = bev_transform_tools(c["input image size"], c["output image size"], c["distance to target"],
                              c["tile_length"], c["cm_per_px"], c["yaw"], c["is_laserscan"])
    bev._bev_matrix = np.asarray(c["bev matrix"]).reshape(3, 3)
    return bev, c, rows, cols


# ------------------------------------------------------------------ ENET.preprocess
@pytest.mark.parametrize("case", ["native", "p720", "odd", "up", "x2"])
def test_preprocess_bit_exact_vs_reference_golden(case):
    from bugcar_image_segmentation_b200.models import ENET
    g = golden("pre.npz")
    h, w, seed = (int(v) for v in g[case + "_hw_seed"])
    frame = synth.blocky_frame(seed, h, w) if seed % 2 else synth.noise_frame(seed, h, w)
    out = ENET.preprocess(frame)
    assert out.dtype == np.float64 and out.shape == (1, 3, 256, 512)
    assert np.array_equal(out.reshape(-1)[::997], g[case + "_sample"])
    assert _sha(out) == str(g[case + "_sha"])


@pytest.mark.parametrize("hw", [(720, 1280), (375, 621), (120, 160), (512, 1024), (1080, 1920), (255, 511), (256, 512)])
def test_resize_and_preprocess_vs_oracle(ctx, hw):
    import torch
    from oracle import cv_ops, pre_oracle
    h, w = hw
    frames = np.stack([synth.noise_frame(40 + i, h, w) for i in range(3)])
    d = torch.from_numpy(frames).cuda()
    r = torch.empty((3, 256, 512, 3), dtype=torch.uint8, device="cuda")
    ctx.resize_bgr(d, h, w, 3, r)
    r = r.cpu().numpy()
    for i in range(3):
        assert np.array_equal(r[i], cv_ops.resize_bilinear_u8(frames[i], (512, 256))), i
    o32 = torch.empty((3, 3, 256, 512), dtype=torch.float32, device="cuda")
    ctx.preprocess(d, h, w, 3, o32, 0)
    want = np.concatenate([pre_oracle.preprocess(f) for f in frames])
    assert np.array_equal(o32.cpu().numpy(), want.astype(np.float32))       # TensorFlow's feed cast


# --------------------------------------------------------------------- argmax + LUT
def test_argmax_lut_golden_ties(ctx):
    import torch
    from bugcar_image_segmentation_b200.models import ENET
    g = golden("argmax.npz")
    rng = np.random.default_rng(int(g["seed"]))
    logits = rng.integers(-3, 4, (2, 15, 32, 64)).astype(np.float32)
    d = torch.from_numpy(logits).cuda()
    out = torch.empty((2, 32, 64), dtype=torch.uint8, device="cuda")
    ctx.argmax_lut(d, 2, 15, 32, 64, ENET.LUT_3WAY, out)
    assert np.array_equal(out.cpu().numpy(), g["labels3"])
    ctx.argmax_lut(d, 2, 15, 32, 64, ENET.LUT_BINARY, out)
    assert np.array_equal(out.cpu().numpy(), g["labels2"])


# ------------------------------------------------------------- label map -> grid (K9)
@pytest.mark.parametrize("cal", ["A", "B", "C", "D", "E"])
def test_grid_bit_exact_vs_reference_golden(cal):
    bev, c, rows, cols = _bev(cal)
    g = golden(f"bev_{cal}.npz")
    args = tuple(float(v) for v in g["grid_args"])
    for s in g["label_seeds"]:
        lab3 = synth.label_map(int(s), 3, rows, cols)
        lab2 = synth.label_map(100 + int(s), 2, rows, cols)
        o3 = bev.create_occupancy_grid(lab3, *args)
        o2 = bev.create_occupancy_grid_binary(lab2, *args)
        assert o3.dtype == np.int8 and o3.shape == g[f"grid3_{s}"].shape
        assert np.array_equal(o3, g[f"grid3_{s}"])
        assert np.array_equal(o2, g[f"grid2_{s}"])
    fine = synth.label_map(7, 3, rows, cols, block=2)
    assert np.array_equal(bev.create_occupancy_grid(fine, *args), g["grid3_fine"])
    assert np.array_equal(bev.create_occupancy_grid_binary((fine == 1).astype(np.uint8), *args), g["grid2_fine"])
    alt = tuple(float(v) for v in g["grid3_alt_args"])
    assert np.array_equal(bev.create_occupancy_grid(synth.label_map(0, 3, rows, cols), *alt), g["grid3_alt"])


def test_grid_fuzz_vs_oracle_batched_and_ros(ctx):
    """random homographies, grid requests and label maps (incl. labels up to 255, which wrap
    in np.add(segmap, 1), bev.py:177); batch of frames; ROS layout."""
    import torch
    from oracle import bev_oracle
    rng = np.random.default_rng(17)
    for it in range(10):
        rows, cols = int(rng.integers(40, 300)), int(rng.integers(40, 520))
        ww, wh = int(rng.integers(60, 700)), int(rng.integers(60, 700))
        quad = np.array([[0.3 * cols, 0.5 * rows], [0.7 * cols, 0.5 * rows], [cols, rows], [0, rows]], np.float64)
        quad += rng.normal(0, 4, (4, 2))
        dst = np.array([[.3 * ww, 0], [.7 * ww, 0], [.7 * ww, wh], [.3 * ww, wh]]) + rng.normal(0, 3, (4, 2))
        M = synth.perspective_transform(quad, dst)
        cm = float(rng.choice([1, 2, 2.5, 3, 4]))
        w_m, h_m = float(rng.uniform(3, 14)), float(rng.uniform(3, 14))
        cell = float(rng.choice([0.05, 0.1, 0.15, 0.2, 0.33]))
        B = 3
        hi = 256 if it % 3 == 0 else 3
        labs = np.stack([synth.label_map(int(rng.integers(1 << 20)), hi, rows, cols, block=int(rng.integers(1, 12)))
                         for _ in range(B)])
        ctx.set_bev(M.reshape(-1), rows, cols, ww, wh, cm)
        hc, wc = ctx.occgrid_shape(w_m, h_m, cell)
        d = torch.from_numpy(labs).cuda()
        for binary in (0, 1):
            out = torch.empty((B, hc, wc), dtype=torch.int8, device="cuda")
            ctx.occgrid(d, B, w_m, h_m, cell, binary, 0, out)
            ros = torch.empty((B, wc, hc), dtype=torch.int8, device="cuda")
            ctx.occgrid(d, B, w_m, h_m, cell, binary, 1, ros)
            out, ros = out.cpu().numpy(), ros.cpu().numpy()
            for i in range(B):
                want = bev_oracle.occupancy_grid(labs[i], M, ww, wh, cm, w_m, h_m, cell, binary=bool(binary))
                assert want.shape == (hc, wc)
                assert np.array_equal(out[i], want), (it, binary, i)
                assert np.array_equal(ros[i].reshape(-1), bev_oracle.ros_layout(want)), (it, binary, i)


def test_grid_tiny_label_maps_every_footprint_on_a_border(ctx):
    """K9 samples fixed 2 x 2 blocks of label pixels with the border taps folded into the weights (prepost.cu
    k_occ_table): label maps of 2..9 pixels per side under shifted / scaled homographies put every case of that
    folding (footprint one pixel before the first or on the last row / column, fully outside) into play."""
    import torch
    from oracle import bev_oracle
    rng = np.random.default_rng(5)
    for it in range(24):
        rows, cols = int(rng.integers(2, 10)), int(rng.integers(2, 10))
        ww, wh = int(rng.integers(20, 60)), int(rng.integers(20, 60))
        # source rectangle a little larger than the image (so the warp reads past all four borders), jittered
        quad = np.array([[-1.5, -1.5], [cols + 0.5, -1.5], [cols + 0.5, rows + 0.5], [-1.5, rows + 0.5]], np.float64)
        quad += rng.normal(0, 0.4, (4, 2))
        dst = np.array([[0, 0], [ww, 0], [ww, wh], [0, wh]], np.float64)
        M = synth.perspective_transform(quad, dst)
        cm = float(rng.choice([2, 4, 5]))
        cell = float(rng.choice([0.05, 0.1, 0.2]))
        w_m, h_m = ww * cm / 100.0 * float(rng.uniform(0.6, 1.0)), wh * cm / 100.0 * float(rng.uniform(0.6, 1.0))
        labs = rng.integers(0, 4, (2, rows, cols)).astype(np.uint8)
        ctx.set_bev(M.reshape(-1), rows, cols, ww, wh, cm)
        try:
            hc, wc = ctx.occgrid_shape(w_m, h_m, cell)
        except Exception:
            continue                                        # empty grid request
        d = torch.from_numpy(labs).cuda()
        for binary in (0, 1):
            out = torch.empty((2, hc, wc), dtype=torch.int8, device="cuda")
            ctx.occgrid(d, 2, w_m, h_m, cell, binary, 0, out)
            for i in range(2):
                want = bev_oracle.occupancy_grid(labs[i], M, ww, wh, cm, w_m, h_m, cell, binary=bool(binary))
                assert np.array_equal(out[i].cpu().numpy(), want), (it, rows, cols, binary, i)


def test_grid_error_behaviour(tmp_path):
    from bugcar_image_segmentation_b200 import _lib
    bev, c, rows, cols = _bev("A")
    with pytest.raises(AssertionError):                     # bev.py:169-170
        bev.create_occupancy_grid(np.zeros((100, 100), np.uint8), 10.0, 10.0, 0.1)
    bev.laserscan_like_occupancy_grid = True
    with pytest.raises(NotImplementedError):                # the ROS layout exists for ordinary grids only
        bev.create_occupancy_grid_ros(np.zeros((256, 512), np.uint8), 10.0, 10.0, 0.1)
    bev.laserscan_like_occupancy_grid = False
    raw = _lib.Context(0, 1)
    with pytest.raises(_lib.BugcarError) as e:              # calibration not set
        raw.occgrid_shape(10.0, 10.0, 0.1)
    assert e.value.code == _lib.BC_ERR_STATE
    with pytest.raises(_lib.BugcarError) as e:              # a label map needs two pixels per side (include/bugcar_b200.h)
        raw.set_bev(np.eye(3).reshape(-1), 1, 512, 500, 500, 2.0)
    assert e.value.code == _lib.BC_ERR_ARG
    with pytest.raises(_lib.BugcarError) as e:              # weights not loaded
        raw.enet_labels(1, 0, 1, np.zeros(256, np.uint8), 1)
    assert e.value.code == _lib.BC_ERR_STATE
    with pytest.raises(_lib.BugcarError) as e:
        raw.load_enet(b"not a container")
    assert e.value.code == _lib.BC_ERR_FORMAT
    raw.close()
