"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference, under
the import stubs of oracle/refstub.py) and OpenCV 4.13 on seeded inputs.

Container-side tool: /root/reference does not exist on the GPU box, so the vectors are
committed.  Inputs are not stored, only their seeds (bugcar_image_segmentation_b200.synth
regenerates them); large outputs are stored as sha256 + a strided sample.

  bev_<cal>.npz    reference create_occupancy_grid / _binary on label maps, calibrations A-E
  warp_<cal>.npz   cv2.warpPerspective of (labels+1)                    (bev.py:182)
  pre.npz          reference ENET.preprocess on native / 720p / odd-size frames (models.py:84-95)
  resize.npz       cv2.resize(bgr,(512,256)) hashes                     (models.py:87)
  argmax.npz       logits with ties -> labels (np.argmax == tf.math.argmax tie-break; models.py:55-58)
  contour.npz      reference contour_noise_removal on synth.road_mask seeds (image_processing_utils.py:4-44)
  laser.npz        reference laserscan-like grids (bev.py:145-164, 216-240) with WARP_FILL_OUTLIERS patched
                   into its two warpPolar calls (the unpatched branch reads uninitialised memory)
  enet.npz         oracle (torch fp32) logits sample for the synthetic weights: PARITY UNPINNED,
                   guards the oracle against drift only
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bugcar_image_segmentation_b200 import synth, weights as W   # noqa: E402
from oracle import refstub, enet_oracle, pre_oracle             # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
GRID_ARGS = (10.0, 10.0, 0.1)
LABEL_SEEDS = (0, 1, 2)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def ref_bev(ref, cal):
    return _ref_bev(ref, cal)


def _ref_bev(ref, cal):
    import json, tempfile
    with tempfile.NamedTemporaryFile("w", suffix=".json", delete=False) as f:
        json.dump(cal, f)
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        bev = ref.bev.bev_transform_tools.fromJSON(f.name)
    os.unlink(f.name)
    return bev


def main():
    import cv2
    ref = refstub.load()
    os.makedirs(OUT, exist_ok=True)

    # ---- BEV grids + warps
    for name in synth.CALIBRATIONS:
        rows, cols = (720, 1280) if name == "E" else (256, 512)
        cal = synth.calibration(name, rows, cols)
        bev = _ref_bev(ref, cal)
        M = np.asarray(cal["bev matrix"], np.float64).reshape(3, 3)
        ww, wh = cal["output image size"]
        d = {"grid_args": np.array(GRID_ARGS), "label_seeds": np.array(LABEL_SEEDS), "rows": rows, "cols": cols}
        wd = {"label_seeds": np.array(LABEL_SEEDS)}
        for s in LABEL_SEEDS:
            lab3 = synth.label_map(s, 3, rows, cols)
            lab2 = synth.label_map(100 + s, 2, rows, cols)
            d[f"grid3_{s}"] = bev.create_occupancy_grid(lab3, *GRID_ARGS)
            d[f"grid2_{s}"] = bev.create_occupancy_grid_binary(lab2, *GRID_ARGS)
            assert d[f"grid3_{s}"].dtype == np.int8 and d[f"grid2_{s}"].dtype == np.int8
            wd[f"warp3_{s}"] = cv2.warpPerspective(np.add(lab3, 1), M, (ww, wh))
        # a fine-grained label map (2-px blocks): many specks for the 3x3 opening
        fine = synth.label_map(7, 3, rows, cols, block=2)
        d["grid3_fine"] = bev.create_occupancy_grid(fine, *GRID_ARGS)
        d["grid2_fine"] = bev.create_occupancy_grid_binary((fine == 1).astype(np.uint8), *GRID_ARGS)
        # other grid requests on the same calibration (non-square, coarse cells)
        d["grid3_alt_args"] = np.array([8.0, 6.0, 0.25])
        d["grid3_alt"] = bev.create_occupancy_grid(synth.label_map(0, 3, rows, cols), 8.0, 6.0, 0.25)
        np.savez_compressed(os.path.join(OUT, f"bev_{name}.npz"), **d)
        if name != "E":
            np.savez_compressed(os.path.join(OUT, f"warp_{name}.npz"), **wd)

    # ---- preprocess / resize
    pre = {}
    rs = {}
    cases = {"native": (256, 512, 11), "p720": (720, 1280, 12), "odd": (375, 621, 13), "up": (120, 160, 14),
             "x2": (512, 1024, 15)}
    for k, (h, w, seed) in cases.items():
        frame = synth.blocky_frame(seed, h, w) if seed % 2 else synth.noise_frame(seed, h, w)
        out = ref.models.ENET.preprocess(frame)
        assert out.dtype == np.float64 and out.shape == (1, 3, 256, 512)
        pre[k + "_hw_seed"] = np.array([h, w, seed])
        pre[k + "_sha"] = np.array(sha(out))
        pre[k + "_sample"] = out.reshape(-1)[::997].copy()
        r = cv2.resize(frame, (512, 256))
        rs[k + "_sha"] = np.array(sha(r))
        rs[k + "_sample"] = r.reshape(-1)[::499].copy()
    np.savez_compressed(os.path.join(OUT, "pre.npz"), **pre)
    np.savez_compressed(os.path.join(OUT, "resize.npz"), **rs)

    # ---- argmax + LUT with ties
    rng = np.random.default_rng(77)
    logits = rng.integers(-3, 4, (2, 15, 32, 64)).astype(np.float32)   # coarse values: many exact ties
    np.savez_compressed(os.path.join(OUT, "argmax.npz"), seed=77,
                        labels3=pre_oracle.labels_from_logits(logits, pre_oracle.LUT_3WAY),
                        labels2=pre_oracle.labels_from_logits(logits, pre_oracle.LUT_BINARY))

    # ---- contour_noise_removal (8f-2)
    cn = {"seeds": np.arange(28), "shapes": np.array([[256, 512], [256, 512], [120, 200], [360, 640]])}
    for s_ in cn["seeds"]:
        h, w = cn["shapes"][s_ % 4]
        m = synth.road_mask(int(s_), int(h), int(w))
        out = ref.image_processing_utils.contour_noise_removal(m)
        assert out.dtype == np.uint8 and out.shape == m.shape and out.max() <= 1
        k = int(min(h, w) / 50)
        cn[f"out_{s_}"] = np.packbits(out)
        cn[f"closed_sha_{s_}"] = np.array(sha(cv2.morphologyEx(m, cv2.MORPH_CLOSE, np.ones((k, k), np.uint8))))
    np.savez_compressed(os.path.join(OUT, "contour.npz"), **cn)

    # ---- laserscan-like grids (8f-3), reference made deterministic by refstub.deterministic_laserscan
    undo = refstub.deterministic_laserscan(ref)
    ls = {"cals": np.array(["A", "B"]), "grid_args": np.array([[10.0, 10.0, 0.1], [8.0, 6.0, 0.25], [6.0, 9.0, 0.2]]),
          "seeds": np.arange(3)}
    for name in ls["cals"]:
        cal = dict(synth.calibration(str(name)), is_laserscan=True)
        bev = _ref_bev(ref, cal)
        for ai, args_ in enumerate(ls["grid_args"]):
            for s_ in ls["seeds"]:
                lab3 = synth.label_map(300 + int(s_), 3, block=16 if s_ else 32)
                lab2 = synth.label_map(400 + int(s_), 2, block=16 if s_ else 32)
                ls[f"g3_{name}_{ai}_{s_}"] = bev.create_occupancy_grid(lab3, *args_)
                plain, laser = bev.create_occupancy_grid_binary(lab2, *args_)
                ls[f"g2p_{name}_{ai}_{s_}"] = plain
                ls[f"g2l_{name}_{ai}_{s_}"] = laser
    undo()
    np.savez_compressed(os.path.join(OUT, "laser.npz"), **ls)

    # ---- ENet oracle self-pin
    with open(os.path.join(ROOT, "pretrained_models", "enet_synthetic_seed42.bcw"), "rb") as f:
        w, nc, eps = W.unpack_flat(f.read())
    frames = synth.frames(2, 1234)
    x = np.concatenate([pre_oracle.preprocess(f) for f in frames])
    lg = enet_oracle.forward(w, x, eps)
    np.savez_compressed(os.path.join(OUT, "enet.npz"), seed0=1234, n=2,
                        logits_sample=lg.reshape(-1)[::4099].copy(),
                        labels3_sha=np.array(sha(pre_oracle.labels_from_logits(lg, pre_oracle.LUT_3WAY))),
                        class_hist=np.bincount(lg.argmax(1).reshape(-1), minlength=nc))
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
