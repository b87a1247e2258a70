"""GPU parity of bc_contour_noise_removal (contour.cu) with the reference's
contour_noise_removal (image_processing_utils.py:4-44): committed reference outputs
(tests/golden/contour.npz), the CPU oracle on seeded masks of several shapes, and
size-independent properties at the full batch size.  Bit-exact."""
import numpy as np
import pytest

from conftest import golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cn():
    import torch
    from bugcar_image_segmentation_b200 import image_processing_utils as ipu
    assert torch.cuda.is_available()
    return ipu.contour_noise_removal


def test_reference_golden(cn):
    from bugcar_image_segmentation_b200 import synth
    g = golden("contour.npz")
    kept = 0
    for s in g["seeds"]:
        h, w = (int(v) for v in g["shapes"][int(s) % 4])
        m = synth.road_mask(int(s), h, w)
        want = np.unpackbits(g[f"out_{s}"])[:h * w].reshape(h, w)
        got = cn(m)
        assert got.dtype == np.uint8 and got.shape == (h, w)
        assert np.array_equal(got, want), f"seed {s}: {(got != want).sum()} pixels differ"
        kept += int(want.any())
    assert kept >= 10


@pytest.mark.parametrize("shape", [(256, 512), (120, 200), (50, 50), (300, 300), (360, 640), (97, 333), (720, 1280)])
def test_oracle_shapes(cn, shape):
    from bugcar_image_segmentation_b200 import synth
    from oracle import contour_oracle
    h, w = shape
    n = 4 if h * w > 500_000 else 14
    masks = np.stack([synth.road_mask(1000 + i, h, w) for i in range(n)])
    got = cn(masks)                                   # one batched call
    for i in range(n):
        want = contour_oracle.contour_noise_removal(masks[i])
        assert np.array_equal(got[i], want), f"shape {shape} seed {1000 + i}: {(got[i] != want).sum()} differ"


def test_opencv_formulation(cn):
    """The contour-based formulation (the OpenCV calls the reference makes) on seeds the golden
    file does not hold; OpenCV is in the image on the GPU box too."""
    pytest.importorskip("cv2")
    from bugcar_image_segmentation_b200 import synth
    from oracle import contour_oracle
    masks = np.stack([synth.road_mask(3000 + i) for i in range(28)])
    got = cn(masks)
    for i in range(len(masks)):
        assert np.array_equal(got[i], contour_oracle.contour_noise_removal_cv2(masks[i])), f"seed {3000 + i}"


def test_kept_hole_and_island(cn):
    from oracle import contour_oracle
    m = np.zeros((256, 512), np.uint8)
    m[100:256, :] = 1
    m[150:254, 20:500] = 0
    m[160:246, 40:480] = 1
    m[200:240, 150:350] = 0
    out = cn(m)
    assert np.array_equal(out, contour_oracle.contour_noise_removal(m))
    assert out[155, 30] == 0 and out[180, 100] == 1 and out[220, 200] == 1 and out[149, 100] == 1


def test_edges(cn):
    z = np.zeros((256, 512), np.uint8)
    assert not cn(z).any()
    assert cn(z + 1).all()
    assert cn(z + 255).all()                        # any non-zero value is road
    from bugcar_image_segmentation_b200._lib import BugcarError
    with pytest.raises(BugcarError):
        cn(np.zeros((40, 80), np.uint8))            # k = int(40/50) = 0: the reference's kernel is empty
    with pytest.raises(TypeError):
        cn(np.zeros((256, 512), np.float32))


def test_full_batch_properties(cn):
    """bs 256 at network resolution: the batched launch equals per-frame launches (frame i and its copy
    at i + 240 agree, and both agree with a single-frame call), and the output is always {0,1}."""
    import torch
    from bugcar_image_segmentation_b200 import synth
    base = np.stack([synth.road_mask(2000 + i) for i in range(16)])
    masks = np.concatenate([base] * 16)               # (256,256,512)
    d = torch.from_numpy(masks).cuda()
    out = cn(d)
    assert out.is_cuda and out.dtype == torch.uint8 and int(out.max()) <= 1
    out = out.cpu().numpy()
    for i in range(16):
        assert np.array_equal(out[i], out[i + 240])
        assert np.array_equal(out[i], cn(base[i]))


def test_pipeline_with_contour_filter(synthetic_weights):
    """bc_set_contour_filter(1): binary pipeline == predict_binary -> contour_noise_removal -> grid."""
    import torch
    from bugcar_image_segmentation_b200 import synth, _lib
    blob = synthetic_weights[0]
    ctx = _lib.Context(0, 4)
    ctx.load_enet(blob)
    cal = synth.calibration("A")
    ww, wh = cal["output image size"]
    ctx.set_bev(cal["bev matrix"], 256, 512, ww, wh, cal["cm_per_px"])
    lut = np.zeros(256, np.uint8)
    lut[[0, 1]] = 1
    frames = torch.from_numpy(np.stack([synth.region_frame(50 + i)[0] for i in range(4)])).cuda()
    hc, wc = ctx.occgrid_shape(10.0, 10.0, 0.1)
    labels = torch.empty((4, 256, 512), dtype=torch.uint8, device="cuda")
    g_plain = torch.empty((4, hc, wc), dtype=torch.int8, device="cuda")
    g_filt = torch.empty_like(g_plain)
    g_want = torch.empty_like(g_plain)
    ctx.pipeline(frames, 256, 512, 4, lut, 10.0, 10.0, 0.1, 1, 0, labels, g_plain)
    ctx.set_contour_filter(True)
    ctx.pipeline(frames, 256, 512, 4, lut, 10.0, 10.0, 0.1, 1, 0, labels, g_filt)
    filt = torch.empty_like(labels)
    ctx.contour_noise_removal(labels, 256, 512, 4, filt)
    ctx.occgrid(filt, 4, 10.0, 10.0, 0.1, 1, 0, g_want)
    torch.cuda.synchronize()
    assert torch.equal(g_filt, g_want)
    ctx.set_contour_filter(False)
    g_again = torch.empty_like(g_plain)
    ctx.pipeline(frames, 256, 512, 4, lut, 10.0, 10.0, 0.1, 1, 0, labels, g_again)
    torch.cuda.synchronize()
    assert torch.equal(g_again, g_plain)
    ctx.close()
