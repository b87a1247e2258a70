"""GPU parity of bc_occgrid_laserscan (laser.cu) with the reference's laserscan branch
(bev.py:145-164, 216-240) made deterministic (WARP_FILL_OUTLIERS): committed outputs of the
patched reference (tests/golden/laser.npz) and the CPU oracle on further seeds.  Bit-exact."""
import numpy as np
import pytest

from conftest import golden

pytestmark = pytest.mark.gpu


def _bev(cal_name):
    from bugcar_image_segmentation_b200 import synth
    from bugcar_image_segmentation_b200.bev import bev_transform_tools
    cal = dict(synth.calibration(cal_name), is_laserscan=True)
    bev = bev_transform_tools(cal["input image size"], cal["output image size"], cal["distance to target"],
                              cal["tile_length"], cal["cm_per_px"], cal["yaw"], True)
    bev._bev_matrix = np.asarray(cal["bev matrix"]).reshape(3, 3)
    return bev, cal


def test_patched_reference_golden():
    from bugcar_image_segmentation_b200 import synth
    g = golden("laser.npz")
    for name in g["cals"]:
        bev, _ = _bev(str(name))
        for ai, args in enumerate(g["grid_args"]):
            args = tuple(float(v) for v in args)
            for s in g["seeds"]:
                s = int(s)
                lab3 = synth.label_map(300 + s, 3, block=16 if s else 32)
                lab2 = synth.label_map(400 + s, 2, block=16 if s else 32)
                key = f"{name}_{ai}_{s}"
                g3 = bev.create_occupancy_grid(lab3, *args)
                assert g3.dtype == np.int8 and np.array_equal(g3, g["g3_" + key]), key
                plain, laser = bev.create_occupancy_grid_binary(lab2, *args)
                assert np.array_equal(plain, g["g2p_" + key]) and np.array_equal(laser, g["g2l_" + key]), key


def test_batched_vs_oracle():
    import torch
    from bugcar_image_segmentation_b200 import synth
    from oracle import bev_oracle, laser_oracle
    bev, cal = _bev("C")
    ctx = bev._context()
    ww, wh = cal["output image size"]
    args = (9.0, 7.0, 0.1)
    B = 6
    lab3 = np.stack([synth.label_map(800 + i, 3, block=8 + 4 * i) for i in range(B)])
    hc, wc = ctx.occgrid_shape(*args)
    for binary in (0, 1):
        labs = lab3 if not binary else (lab3 == 1).astype(np.uint8)
        d_lab = torch.from_numpy(labs).cuda()
        d_plain = torch.empty((B, hc, wc), dtype=torch.int8, device="cuda")
        d_laser = torch.empty_like(d_plain)
        ctx.occgrid_laserscan(d_lab, B, *args, binary, d_plain if binary else None, d_laser)
        torch.cuda.synchronize()
        for i in range(B):
            plain, templ = bev_oracle.occupancy_grid(labs[i], cal["bev matrix"], ww, wh, cal["cm_per_px"], *args,
                                                     binary=bool(binary), return_template=True)
            if binary:
                _, want = laser_oracle.laserscan_binary(plain)
                assert np.array_equal(d_plain[i].cpu().numpy(), plain)
            else:
                want = laser_oracle.laserscan_3way(templ)
            assert np.array_equal(d_laser[i].cpu().numpy(), want), (binary, i)


def test_no_obstacle():
    bev, _ = _bev("A")
    road = np.ones((256, 512), np.uint8)                       # everything road: nothing to hit
    g3 = bev.create_occupancy_grid(road, 10.0, 10.0, 0.1)
    assert not (g3 == 100).any()


def test_fused_pipeline_laserscan():
    """FramePipeline on a laserscan calibration == its own labels -> create_occupancy_grid[_binary] on that calibration."""
    import os
    import torch
    from bugcar_image_segmentation_b200 import synth
    from bugcar_image_segmentation_b200.models import ENET
    from bugcar_image_segmentation_b200.pipeline import FramePipeline
    from conftest import ROOT
    bev, _ = _bev("A")
    model = ENET(os.path.join(ROOT, "pretrained_models", "enet_synthetic_trained.bcw"), device=0, max_batch=2)
    frames = np.stack([synth.region_frame(70 + i)[0] for i in range(3)])
    for binary in (False, True):
        pipe = FramePipeline(model, bev, 10.0, 10.0, 0.1, binary=binary)
        d_lab = torch.empty((3, 256, 512), dtype=torch.uint8, device="cuda")
        got = pipe.run_device(torch.from_numpy(frames).cuda(), d_labels=d_lab).cpu().numpy()   # 3 frames, max_batch 2
        lab = d_lab.cpu().numpy()
        assert np.array_equal(got, pipe(frames))                    # host entry point, same grids
        for i in range(3):
            if binary:
                want = bev.create_occupancy_grid_binary(lab[i], 10.0, 10.0, 0.1)[1]
            else:
                want = bev.create_occupancy_grid(lab[i], 10.0, 10.0, 0.1)
            assert np.array_equal(got[i], want), (binary, i)
