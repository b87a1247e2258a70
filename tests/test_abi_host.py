"""CPU-side checks: the C-ABI library loads and exports every symbol include/bugcar_b200.h
declares, host logic of the drop-in modules, frame sharding across ranks (gloo)."""
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "bugcar_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bc_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from bugcar_image_segmentation_b200 import _lib
    if not os.path.isfile(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    lib = _lib.load()
    syms = _header_symbols()
    assert len(syms) >= 20
    assert set(syms) == set(_lib.SIGNATURES), set(syms) ^ set(_lib.SIGNATURES)
    for s in syms:
        assert hasattr(lib, s), s
    assert lib.bc_abi_version() == 1


def test_no_cpu_fallback():
    """Without a GPU the product path must fail loudly, not fall back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from bugcar_image_segmentation_b200 import _lib
    from bugcar_image_segmentation_b200.models import ENET
    with pytest.raises(_lib.BugcarError) as e:
        _lib.Context(0, 1)
    assert "no CPU fallback" in str(e.value)
    with pytest.raises(RuntimeError):
        ENET()
    with pytest.raises(RuntimeError):
        ENET.preprocess(np.zeros((256, 512, 3), np.uint8))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "bugcar_image_segmentation_b200")
    for f in os.listdir(pkg):
        if f.endswith(".py"):
            src = open(os.path.join(pkg, f)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f


def test_bev_json_roundtrip_and_errors(tmp_path):
    from bugcar_image_segmentation_b200 import synth
    from bugcar_image_segmentation_b200.bev import bev_transform_tools
    cal = synth.calibration("B")
    p = tmp_path / "c.json"
    p.write_text(json.dumps(cal))
    bev = bev_transform_tools.fromJSON(str(p))
    assert (bev.input_width, bev.input_height) == (256, 512)           # (rows, cols) quirk, bev.py:169
    assert (bev.after_warp_width, bev.after_warp_height) == (600, 400)
    assert bev._bev_matrix.shape == (3, 3) and bev.laserscan_like_occupancy_grid is False
    q = tmp_path / "d.json"
    bev.save_to_JSON(str(q))
    again = bev_transform_tools.fromJSON(str(q))
    assert np.array_equal(again._bev_matrix, bev._bev_matrix) and again.cm_per_px == bev.cm_per_px
    bad = dict(cal)
    del bad["cm_per_px"]
    p.write_text(json.dumps(bad))
    with pytest.raises(KeyError):                                        # bev.py:35
        bev_transform_tools.fromJSON(str(p))
    with pytest.raises(AssertionError):                                  # bev.py:169-170 (before any GPU work)
        bev.create_occupancy_grid(np.zeros((512, 256), np.uint8), 10.0, 10.0, 0.1)


def test_calibration_matrix_maps_tile_to_square():
    from bugcar_image_segmentation_b200.bev import bev_transform_tools
    bev = bev_transform_tools([256, 512], [500, 500], (0, 100), 60, 2, 0.0)
    tile = np.array([[200, 150], [312, 150], [340, 200], [172, 200]], np.float64)
    M = bev.calculate_transform_matrix(tile)
    pts = np.concatenate([tile, np.ones((4, 1))], 1) @ M.T
    pts = pts[:, :2] / pts[:, 2:]
    # a 60 cm tile at 2 cm/px is a 30 px square centred 50 px above the bottom centre
    assert np.allclose(sorted(pts[:, 0]), [235, 235, 265, 265], atol=1e-6)
    assert np.allclose(sorted(pts[:, 1]), [435, 435, 465, 465], atol=1e-6)
    cv2 = pytest.importorskip("cv2")
    order = np.argsort(pts[:, 0] * 1000 + pts[:, 1])
    dst = pts[order].astype(np.float32)
    assert np.allclose(cv2.getPerspectiveTransform(tile[order].astype(np.float32), dst), M, rtol=1e-6, atol=1e-6)


def test_small_host_helpers():
    from bugcar_image_segmentation_b200.image_processing_utils import find_intersection_line
    from bugcar_image_segmentation_b200.occgrid_to_ros import ros_cell_order
    from bugcar_image_segmentation_b200.models import ENET
    p = find_intersection_line([(0, 0), (2, 2)], [(0, 2), (2, 0)])
    assert np.allclose(p, [1, 1])
    assert find_intersection_line([(0, 0), (1, 1)], [(0, 1), (1, 2)]) is None
    g = np.arange(6, dtype=np.int8).reshape(2, 3)
    assert ros_cell_order(g).tolist() == [5, 2, 4, 1, 3, 0]
    assert ENET.LUT_3WAY[[0, 1, 2, 9, 5]].tolist() == [1, 1, 0, 0, 2] and ENET.LUT_BINARY.sum() == 2
    assert (ENET.INPUT_WIDTH, ENET.INPUT_HEIGHT) == (512, 256)


def test_frame_sharding_two_ranks_gloo(tmp_path):
    """N-rank run == 1-rank run on the concatenated batch: host logic of bench.py's
    sharding + gather, world_size 2 over gloo on CPU (no kernels involved)."""
    script = tmp_path / "w.py"
    script.write_text(f"""
import os, sys
sys.path.insert(0, {ROOT!r})
import numpy as np, torch, torch.distributed as dist
from bugcar_image_segmentation_b200 import sharding
dist.init_process_group('gloo')
r, n = dist.get_rank(), dist.get_world_size()
B = 6
lo, hi = sharding.frame_range(r, n, B)
seeds = sharding.frame_seeds(r, B)
assert list(seeds) == [1234 + r * B + i for i in range(B)]
local = torch.tensor([[s % 101] * 4 for s in seeds], dtype=torch.int8)        # stand-in "grids"
out = sharding.gather_grids(local, r, n, backend_device='cpu')
if r == 0:
    want = np.array([[(1234 + i) % 101] * 4 for i in range(n * B)], dtype=np.int8)
    assert np.array_equal(out.numpy(), want), (out, want)
    print('gather ok', out.shape)
dist.destroy_process_group()
""")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29571", str(script)],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "gather ok" in r.stdout


def test_state_dict_converter_roundtrip(tmp_path, synthetic_weights):
    """tools/convert_weights.py: a PyTorch state_dict of the canonical ENet -> the same container"""
    import torch
    blob, w, nc, eps = synthetic_weights
    sd = {("module." + k): torch.from_numpy(v) for k, v in w.items()}
    sd["module.initial_block.batch_norm.num_batches_tracked"] = torch.tensor(3)
    src, dst = tmp_path / "enet.pth", tmp_path / "enet.bcw"
    torch.save({"state_dict": sd, "epoch": 1}, src)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "convert_weights.py"), str(src), str(dst),
                        "--classes", str(nc)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    from bugcar_image_segmentation_b200 import weights as W
    got, nc2, eps2 = W.unpack_flat(dst.read_bytes())
    assert nc2 == nc and eps2 == eps
    # the converter also writes the graph and the variant switches it derived (canonical graph, 3x3 pool, 3x3 head)
    assert np.array_equal(got.pop("__graph__"), W.graph_rows()) and list(got.pop("__spec__")[:2]) == [3, 3]
    assert list(got) == list(w) and all(np.array_equal(got[k], w[k]) for k in w)
    assert "not part of the ENet graph" not in r.stderr           # num_batches_tracked is known noise

    # bias=True checkpoints: conv biases are carried over (the loader folds them), unknown keys are reported,
    # a non-zero bias on the class head is refused
    sd2 = dict(sd)
    sd2["module.regular2_1.ext_conv1.0.bias"] = torch.full((32,), 0.5)
    sd2["module.some_aux_head.weight"] = torch.zeros(3)
    torch.save(sd2, src)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "convert_weights.py"), str(src), str(dst),
                        "--classes", str(nc)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    got, _, _ = W.unpack_flat(dst.read_bytes())
    assert np.all(got["regular2_1.ext_conv1.0.bias"] == 0.5) and "some_aux_head.weight" in r.stderr
    sd2["module.transposed_conv.bias"] = torch.ones(nc)
    torch.save(sd2, src)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "convert_weights.py"), str(src), str(dst),
                        "--classes", str(nc)], capture_output=True, text=True)
    assert r.returncode != 0 and "class head" in r.stderr


def test_laser_tables_match_oracle():
    """bc_laser_tables is host-only: the C++ restatement of cv::warpPolar's coordinate maps equals the
    oracle's (which tests/test_oracle_laser.py pins against cv2.warpPolar itself)."""
    from bugcar_image_segmentation_b200 import _lib
    from oracle import laser_oracle
    for wc, hc in [(100, 100), (80, 60), (101, 77), (40, 100), (32, 24)]:
        for binary in (0, 1):
            fwd, inv = _lib.laser_tables(wc, hc, binary)
            ph, pw = fwd.shape
            centre, radius = (wc / 2 - 1, hc), max(wc, hc)
            assert binary or (pw, ph) == laser_oracle.polar_dsize(radius)
            assert np.array_equal(fwd, laser_oracle.forward_map(pw, ph, centre[0], centre[1], radius, wc, hc))
            assert np.array_equal(inv, laser_oracle.inverse_map(wc, hc, centre[0], centre[1], radius, pw, ph))


def test_rank_placement_spreads_over_both_host_domains(monkeypatch):
    """runtime.device_for_rank: identity when the job fills the node (or the node shows exactly N GPUs), alternate halves
    otherwise; BUGCAR_DEVICE_MAP overrides"""
    from bugcar_image_segmentation_b200 import runtime
    monkeypatch.delenv("BUGCAR_DEVICE_MAP", raising=False)
    assert [runtime.device_for_rank(r, 8, 8) for r in range(8)] == list(range(8))
    assert [runtime.device_for_rank(r, 4, 4) for r in range(4)] == [0, 1, 2, 3]
    assert [runtime.device_for_rank(r, 4, 8) for r in range(4)] == [0, 4, 1, 5]
    assert [runtime.device_for_rank(r, 2, 8) for r in range(2)] == [0, 4]
    assert runtime.device_for_rank(0, 1, 8) == 0 and runtime.device_for_rank(0, 1, 1) == 0
    monkeypatch.setenv("BUGCAR_DEVICE_MAP", "identity")
    assert [runtime.device_for_rank(r, 4, 8) for r in range(4)] == [0, 1, 2, 3]
    monkeypatch.setenv("BUGCAR_DEVICE_MAP", "3,2,1,0")
    assert [runtime.device_for_rank(r, 4, 8) for r in range(4)] == [3, 2, 1, 0]
