"""Pins oracle/laser_oracle.py (laserscan-like grids, bev.py:145-164 / 216-240 with outliers defined
as 0): its two coordinate maps against cv2.warpPolar, the whole branch against outputs of the
reference's own code run with WARP_FILL_OUTLIERS patched in (tests/golden/laser.npz), and live in the
build container."""
import numpy as np
import pytest

from conftest import golden
from bugcar_image_segmentation_b200 import synth
from oracle import bev_oracle, laser_oracle, refstub


def _grids(cal_name, args, seed):
    cal = synth.calibration(cal_name)
    ww, wh = cal["output image size"]
    lab3 = synth.label_map(300 + seed, 3, block=16 if seed else 32)
    lab2 = synth.label_map(400 + seed, 2, block=16 if seed else 32)
    _, templ = bev_oracle.occupancy_grid(lab3, cal["bev matrix"], ww, wh, cal["cm_per_px"], *args, return_template=True)
    plain2 = bev_oracle.occupancy_grid(lab2, cal["bev matrix"], ww, wh, cal["cm_per_px"], *args, binary=True)
    return laser_oracle.laserscan_3way(templ), laser_oracle.laserscan_binary(plain2)


def test_matches_patched_reference_golden():
    g = golden("laser.npz")
    marked = 0
    for name in g["cals"]:
        for ai, args in enumerate(g["grid_args"]):
            for s in g["seeds"]:
                g3, (plain, laser) = _grids(str(name), tuple(float(v) for v in args), int(s))
                key = f"{name}_{ai}_{s}"
                assert g3.dtype == np.int8 and np.array_equal(g3, g["g3_" + key]), key
                assert np.array_equal(plain, g["g2p_" + key]) and np.array_equal(laser, g["g2l_" + key]), key
                marked += int((laser == 100).sum()) + int((g3 == 100).sum())
    assert marked > 500


@pytest.mark.parametrize("shape", [(100, 100), (80, 60), (101, 77), (40, 100), (32, 24)])
def test_maps_match_warp_polar(shape):
    cv2 = pytest.importorskip("cv2")
    wc, hc = shape
    centre, radius = (wc / 2 - 1, hc), max(shape)
    idx = (np.arange(hc * wc).reshape(hc, wc) + 1).astype(np.float32)
    for dsize in (shape, (-1, -1)):
        pol = cv2.warpPolar(idx, dsize, centre, radius, cv2.WARP_POLAR_LINEAR | cv2.WARP_FILL_OUTLIERS)
        ph, pw = pol.shape
        if dsize == (-1, -1):
            assert (pw, ph) == laser_oracle.polar_dsize(radius)
        assert np.array_equal(pol.astype(np.int64) - 1, laser_oracle.forward_map(pw, ph, centre[0], centre[1], radius, wc, hc))
        pidx = (np.arange(ph * pw).reshape(ph, pw) + 1).astype(np.float64)
        back = cv2.warpPolar(pidx, shape, centre, radius, cv2.WARP_INVERSE_MAP | cv2.WARP_FILL_OUTLIERS)
        assert np.array_equal(back.astype(np.int64) - 1, laser_oracle.inverse_map(wc, hc, centre[0], centre[1], radius, pw, ph))


def test_no_obstacle_and_values():
    t = np.full((50, 60), 2, np.uint8)                       # all free
    assert (laser_oracle.laserscan_3way(t) == 0).all()
    t[10:20, 20:40] = 3
    out = laser_oracle.laserscan_3way(t)
    assert set(np.unique(out)) <= {-1, 0, 100} and (out == 100).any() and (out[10:20, 20:40] != 0).all()


@pytest.mark.skipif(not refstub.available(), reason="/root/reference only exists in the build container")
def test_live_patched_reference(tmp_path):
    import contextlib, io, json
    ref = refstub.load()
    undo = refstub.deterministic_laserscan(ref)
    try:
        cal = dict(synth.calibration("D"), is_laserscan=True)
        p = tmp_path / "cal.json"
        p.write_text(json.dumps(cal))
        with contextlib.redirect_stdout(io.StringIO()):
            bev = ref.bev.bev_transform_tools.fromJSON(str(p))
        ww, wh = cal["output image size"]
        for seed in range(3):
            lab3 = synth.label_map(600 + seed, 3)
            lab2 = synth.label_map(700 + seed, 2)
            for args in ((10.0, 10.0, 0.1), (5.0, 7.0, 0.1)):
                _, templ = bev_oracle.occupancy_grid(lab3, cal["bev matrix"], ww, wh, cal["cm_per_px"], *args, return_template=True)
                assert np.array_equal(laser_oracle.laserscan_3way(templ), bev.create_occupancy_grid(lab3, *args))
                plain = bev_oracle.occupancy_grid(lab2, cal["bev matrix"], ww, wh, cal["cm_per_px"], *args, binary=True)
                want_plain, want_laser = bev.create_occupancy_grid_binary(lab2, *args)
                got_plain, got_laser = laser_oracle.laserscan_binary(plain)
                assert np.array_equal(got_plain, want_plain) and np.array_equal(got_laser, want_laser)
    finally:
        undo()
