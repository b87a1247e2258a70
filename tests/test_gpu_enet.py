"""GPU parity, ENet forward + whole path, through the drop-in classes and the C ABI.

Oracle: oracle/enet_oracle.py (torch fp32 on CPU).  PARITY UNPINNED against the reference's
TensorFlow graph (blob absent, see oracle/__init__.py); tolerances:
  * fp32 mode  vs fp32 oracle:  encoder output (dilated3_7) max|d| <= 1e-4 * max|act|; logits
    |d| <= 1e-4 * max|logit| on >= 99 % of elements (the rest are max-unpool index flips: two
    window values within 1 fp32 ulp order differently under a different summation order, and
    the moved activation spreads through the decoder's 3x3 convs); argmax agreement >= 99.9 %
  * fp16 mode (production, tcgen05) vs the FP32 oracle: RAW per-pixel argmax agreement >= 99.9 %
    for the trained-like weights, no margin filter (north_star's bar); logits vs the
    fp16-emulating oracle (same rounding points) |d| <= 2^-6 * max|logit| on 99.9 % and <= 2^-9 * max|logit|
    on 97 % of logits
  * bf16 mode  vs bf16-emulating oracle: max|d| <= 2^-6 * max|logit| on 99.9 % of logits, argmax
    agreement >= 99.9 % of pixels whose top-2 margin exceeds the error bound; its raw agreement
    with the fp32 oracle is 99.7 % (reported; the reason fp16 is the production storage type)
  * fused labels == argmax+LUT of the same mode's logits, bit-exact
  * grids bit-exact given the label map
"""
import os

import numpy as np
import pytest

from bugcar_image_segmentation_b200 import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WEIGHTS = {"seed42": os.path.join(ROOT, "pretrained_models", "enet_synthetic_seed42.bcw"),
           "trained": os.path.join(ROOT, "pretrained_models", "enet_synthetic_trained.bcw")}


def _load(which):
    from bugcar_image_segmentation_b200 import weights as W
    with open(WEIGHTS[which], "rb") as f:
        return W.unpack_flat(f.read())


def _frames(which, n):
    if which == "trained":
        return np.stack([synth.region_frame(500 + i)[0] for i in range(n)])
    return synth.frames(n, 1234)


@pytest.fixture(scope="module", params=["seed42", "trained"])
def setup(request):
    from bugcar_image_segmentation_b200.models import ENET
    from oracle import pre_oracle, enet_oracle
    which = request.param
    w, nc, eps = _load(which)
    frames = _frames(which, 3)
    x = np.concatenate([pre_oracle.preprocess(f) for f in frames])
    want32 = enet_oracle.forward(w, x, eps)
    want16 = {e: enet_oracle.forward(w, x, eps, emulate=e) for e in ("bf16", "fp16")}
    model = ENET(WEIGHTS[which], device=0, max_batch=8)
    return dict(which=which, model=model, frames=frames, x=x, want32=want32, want16=want16)


def _margin(lg):
    s = np.sort(lg, axis=1)
    return s[:, -1] - s[:, -2]


def test_fp32_logits_and_argmax(setup):
    from bugcar_image_segmentation_b200 import _lib
    from oracle import pre_oracle
    m = setup["model"]
    m.ctx.set_precision(_lib.BC_PREC_FP32)
    want = setup["want32"]
    for inp in (setup["x"], setup["x"].astype(np.float32), setup["frames"]):     # f64 / f32 NCHW, u8 BGR frames
        got = m.logits(inp)
        assert got.shape == want.shape and got.dtype == np.float32
        d = np.abs(got - want) / np.abs(want).max()
        print(f"[{setup['which']}] fp32 logits: median {np.median(d):.2e} p99 {np.percentile(d, 99):.2e} max {d.max():.2e}")
        assert np.percentile(d, 99) <= 1e-4, np.percentile(d, 99)
    lab = m.predict(setup["x"])
    ref = pre_oracle.labels_from_logits(want, pre_oracle.LUT_3WAY)
    assert lab.dtype == np.uint8 and lab.shape == (3, 256, 512)
    assert (lab == ref).mean() >= 0.999
    labb = m.predict_binary(setup["x"])
    assert (labb == pre_oracle.labels_from_logits(want, pre_oracle.LUT_BINARY)).mean() >= 0.999
    assert set(np.unique(labb)) <= {0, 1}


def test_fp32_encoder_blocks_exact(setup):
    """per-block activations (bc_enet_block_output) against the oracle's intermediates: the
    encoder has no index-driven op, so it must agree to fp32 round-off everywhere."""
    import torch
    from bugcar_image_segmentation_b200 import _lib
    from bugcar_image_segmentation_b200.weights import ENET_BLOCKS
    from oracle import enet_oracle
    if setup["which"] != "seed42":
        pytest.skip("one weight set is enough")
    m = setup["model"]
    m.ctx.set_precision(_lib.BC_PREC_FP32)
    w, nc, eps = _load("seed42")
    x = np.ascontiguousarray(setup["x"][:2], dtype=np.float32)
    _, inter = enet_oracle.forward(w, x, eps, return_intermediates=True)
    dx = torch.from_numpy(x).cuda()
    names = ["initial_block"] + [b[0] for b in ENET_BLOCKS]
    for i, name in enumerate(names):
        want = inter[name]
        out = torch.empty(want.shape, dtype=torch.float32, device="cuda")
        m.ctx.enet_block_output(dx, _lib.BC_IN_NCHW_F32, 2, i - 1, out)
        d = np.abs(out.cpu().numpy() - want) / np.abs(want).max()
        if i <= 22:                                  # up to dilated3_7
            assert d.max() <= 1e-4, (name, d.max())
        else:                                        # decoder: a few unpool index flips allowed
            assert np.percentile(d, 99) <= 1e-4, (name, np.percentile(d, 99))


@pytest.mark.parametrize("mode", ["fp16", "bf16"])
def test_16bit_logits_and_argmax(setup, mode):
    from bugcar_image_segmentation_b200 import _lib
    from oracle import pre_oracle
    m = setup["model"]
    m.ctx.set_precision(_lib.PRECISIONS[mode])
    # logits against the oracle that rounds at the same points: |d| <= 2^-6 max|logit| (4 bf16 ulps of the largest
    # logit) on 99.9 % of logits in both modes; fp16 additionally <= 2^-9 max|logit| (4 fp16 ulps) on 97 %.  (What
    # is left is not rounding: a one-ulp difference between two summation orders moves a max-pool index, and
    # the moved activation spreads through the decoder -- finer ulps make such flips smaller but MORE frequent.)
    for tc in (0, 1):
        m.ctx.set_tensor_cores(tc)
        got = m.logits(setup["x"])
        want = setup["want16"][mode]
        scale = np.abs(want).max()
        d = np.abs(got - want)
        tol = 2.0 ** -6 * scale
        frac_ok = (d <= tol).mean()
        frac_fine = (d <= 2.0 ** -9 * scale).mean()
        a, b = got.argmax(1), want.argmax(1)
        raw = (a == b).mean()
        big = _margin(want) > 2 * tol
        conf = (a == b)[big].mean()
        raw32 = (a == setup["want32"].argmax(1)).mean()
        print(f"[{setup['which']} {mode} tc={tc}] vs emulated oracle: |d|/max p50 {np.median(d) / scale:.2e} p99 "
              f"{np.percentile(d, 99) / scale:.2e} p99.9 {np.percentile(d, 99.9) / scale:.2e}; within 2^-6 {frac_ok:.5f}, "
              f"within 2^-9 {frac_fine:.5f}; argmax raw {raw:.5f}, confident ({big.mean():.3f} of px) {conf:.5f}; "
              f"vs FP32 oracle raw {raw32:.5f}")
        # the random-weight net is chaotic (max-unpool flips on white-noise features): 99 % there
        assert frac_ok >= (0.999 if setup["which"] == "trained" else 0.99), frac_ok
        if mode == "fp16" and setup["which"] == "trained":
            assert frac_fine >= 0.97, frac_fine      # measured 0.9945 (CUDA cores) / 0.982 (tcgen05)
        assert conf >= (0.999 if setup["which"] == "trained" else 0.995), conf
        # fused head (argmax + LUT inside the kernel) == argmax + LUT of this mode's logits
        lab = m.predict(setup["x"])
        ref_lab = pre_oracle.labels_from_logits(got, pre_oracle.LUT_3WAY)
        if tc == 0:
            assert np.array_equal(lab, ref_lab)       # same kernel, same arithmetic
        else:
            # the tcgen05 head accumulates in a different order than the CUDA-core logits kernel:
            # labels may differ only where the top-2 logits are within fp32 round-off of each other
            diff = lab != ref_lab
            assert diff.mean() <= 1e-4, diff.mean()
            assert (_margin(got)[diff] <= 1e-4 * scale).all()
        if setup["which"] == "trained":
            # north_star: >= 99.9 % per-pixel argmax agreement with the fp32 network -- RAW, every pixel counted,
            # tensor cores on and off.  bf16 storage cannot reach it (99.7 %), which is why fp16 is the default.
            assert raw32 >= (0.999 if mode == "fp16" else 0.99), (mode, tc, raw32)
            lab32 = pre_oracle.labels_from_logits(setup["want32"], pre_oracle.LUT_3WAY)
            assert (lab == lab32).mean() >= (0.999 if mode == "fp16" else 0.99)
    m.ctx.set_precision(_lib.BC_PREC_FP16)


def test_labels_same_for_all_input_kinds_and_chunks(setup):
    from bugcar_image_segmentation_b200 import _lib
    m = setup["model"]
    m.ctx.set_precision(_lib.BC_PREC_FP16)
    base = m.predict(setup["frames"])
    # fp64 / fp32 NCHW inputs take the same kernel with the same fp32 operands: identical labels
    from_f64 = m.predict(setup["x"])
    assert np.array_equal(m.predict(setup["x"].astype(np.float32)), from_f64)
    # uint8 frames fold the normalisation into the weights (s*u + t, exact bytes as operands, two fp16 weight
    # limbs = 22 significant bits): equal to the float path far below the 16-bit rounding of the block's output,
    # i.e. labels may differ only at 16-bit rounding ties
    # (the random-weight net is chaotic -- a one-ulp 16-bit flip in the first block moves max-pool indices
    # downstream -- so only the trained-like weights give a meaningful bound)
    assert (from_f64 == base).mean() >= (0.9995 if setup["which"] == "trained" else 0.9), (from_f64 == base).mean()
    for chunk in (1, 2, 0):
        m.ctx.set_chunk(chunk)
        assert np.array_equal(m.predict(setup["frames"]), base), chunk
    # frame independence: permuting the batch permutes the output
    perm = [2, 0, 1]
    assert np.array_equal(m.predict(setup["frames"][perm]), base[perm])


@pytest.mark.parametrize("cal", ["A", "C"])
def test_pipeline_equals_staged_calls(setup, cal):
    import torch
    from bugcar_image_segmentation_b200 import _lib
    from bugcar_image_segmentation_b200.bev import bev_transform_tools
    from bugcar_image_segmentation_b200.pipeline import FramePipeline
    from oracle import bev_oracle
    m = setup["model"]
    m.ctx.set_precision(_lib.BC_PREC_FP16)
    c = synth.calibration(cal)
    bev = bev_transform_tools(c["input image size"], c["output image size"], c["distance to target"],
                              c["tile_length"], c["cm_per_px"], c["yaw"], c["is_laserscan"])
    bev._bev_matrix = np.asarray(c["bev matrix"]).reshape(3, 3)
    ww, wh = c["output image size"]
    frames = setup["frames"]
    labels = m.predict(frames)
    for binary in (False, True):
        lab = m.predict_binary(frames) if binary else labels
        staged = np.stack([(bev.create_occupancy_grid_binary if binary else bev.create_occupancy_grid)(l, 10.0, 10.0, 0.1)
                           for l in lab])
        for i in range(len(lab)):      # grids bit-exact given the identical label map
            assert np.array_equal(staged[i], bev_oracle.occupancy_grid(lab[i], c["bev matrix"], ww, wh, c["cm_per_px"],
                                                                        10.0, 10.0, 0.1, binary=binary))
        pipe = FramePipeline(m, bev, 10.0, 10.0, 0.1, binary=binary)
        for graphs in (1, 0):
            m.ctx.set_graphs(graphs)
            assert np.array_equal(pipe(frames), staged)                 # host entry point (H2D + graph + D2H)
            assert np.array_equal(pipe(frames[0]), staged[0])           # single frame
            d_lab = torch.empty((3, 256, 512), dtype=torch.uint8, device="cuda")
            got = pipe.run_device(torch.from_numpy(frames).cuda(), d_labels=d_lab)
            assert np.array_equal(got.cpu().numpy(), staged)
            assert np.array_equal(d_lab.cpu().numpy(), lab)
        m.ctx.set_graphs(1)
        ros = FramePipeline(m, bev, 10.0, 10.0, 0.1, binary=binary, ros_layout=True)(frames)
        for i in range(3):
            assert np.array_equal(ros[i].reshape(-1), bev_oracle.ros_layout(staged[i]))


def test_pipeline_resizes_camera_frames(setup):
    """720p frames: the fused path == ENET.preprocess + predict + create_occupancy_grid."""
    from bugcar_image_segmentation_b200 import _lib
    from bugcar_image_segmentation_b200.models import ENET
    from bugcar_image_segmentation_b200.bev import bev_transform_tools
    from bugcar_image_segmentation_b200.pipeline import FramePipeline
    m = setup["model"]
    m.ctx.set_precision(_lib.BC_PREC_FP16)
    c = synth.calibration("B")
    bev = bev_transform_tools(c["input image size"], c["output image size"], c["distance to target"],
                              c["tile_length"], c["cm_per_px"], c["yaw"], c["is_laserscan"])
    bev._bev_matrix = np.asarray(c["bev matrix"]).reshape(3, 3)
    frames = np.stack([synth.blocky_frame(70 + i, 720, 1280) for i in range(2)])
    x = np.concatenate([ENET.preprocess(f) for f in frames])            # reference-style per-frame calls
    seg = m.predict(x)
    staged = np.stack([bev.create_occupancy_grid(s, 10.0, 10.0, 0.1) for s in seg])
    fused = FramePipeline(m, bev, 10.0, 10.0, 0.1)(frames)
    # the staged calls feed ENet the fp64 tensor of ENET.preprocess, the fused path folds the normalisation into
    # the first conv (raw bytes as operands): same numbers to about 2^-22, so labels / cells can differ only
    # at bf16 rounding ties (chaotic random-weight net: loose bound, see the input-kinds test)
    same = (fused == staged).mean()
    # (trained-like weights on these out-of-distribution blocky frames: 99.6 % of the cells identical, the same
    # order as the bf16-vs-fp32 label noise itself)
    assert fused.shape == staged.shape and same >= (0.99 if setup["which"] == "trained" else 0.9), same


def test_full_batch_properties(setup):
    """bs=256 (BASELINE config 2 size): size-independent checks -- every frame of a batch of
    repeated frames gives the identical grid, and equals the small-batch result."""
    import torch
    from bugcar_image_segmentation_b200 import _lib
    from bugcar_image_segmentation_b200.models import ENET
    from bugcar_image_segmentation_b200.bev import bev_transform_tools
    from bugcar_image_segmentation_b200.pipeline import FramePipeline
    if setup["which"] != "seed42":
        pytest.skip("one weight set is enough")
    big = ENET(WEIGHTS["seed42"], device=0, max_batch=256)
    c = synth.calibration("A")
    bev = bev_transform_tools(c["input image size"], c["output image size"], c["distance to target"],
                              c["tile_length"], c["cm_per_px"], c["yaw"], c["is_laserscan"])
    bev._bev_matrix = np.asarray(c["bev matrix"]).reshape(3, 3)
    pipe = FramePipeline(big, bev, 10.0, 10.0, 0.1)
    frames = setup["frames"]
    small = pipe(frames)
    idx = np.arange(256) % 3
    d = torch.from_numpy(frames[idx]).cuda()
    grids = pipe.run_device(d).cpu().numpy()
    assert grids.shape == (256, 100, 100)
    assert np.array_equal(grids, small[idx])
    assert big.ctx.launch_count() > 0


def test_bs256_distinct_frames_vs_oracle():
    """BASELINE config 2 as the bench runs it: ONE batch of 256 DISTINCT frames (128 colour-region scenes, 64
    blocky, 64 white noise) through the production mode (fp16 storage, tcgen05, bs-256 launch geometry:
    every CTA walks several tiles, all ring / parity wrap-arounds happen), graphs on and off.  The oracle
    (torch fp32 on CPU) is evaluated on 10 sampled frames including the first and the last of the batch
    (= the first and last tiles of every kernel).  Trained-like weights: RAW argmax agreement >= 99.9 % on
    the scene frames, no margin filter."""
    import torch
    from bugcar_image_segmentation_b200.models import ENET
    from bugcar_image_segmentation_b200.bev import bev_transform_tools
    from bugcar_image_segmentation_b200.pipeline import FramePipeline
    from oracle import pre_oracle, enet_oracle, bev_oracle
    w, nc, eps = _load("trained")
    frames = np.empty((256, 256, 512, 3), np.uint8)
    for i in range(256):
        frames[i] = (synth.region_frame(3000 + i)[0] if i % 2 == 0 else
                     synth.blocky_frame(3000 + i) if i % 4 == 1 else synth.noise_frame(3000 + i))
    sample = [0, 1, 2, 3, 126, 128, 191, 253, 254, 255]
    x = np.concatenate([pre_oracle.preprocess(frames[i]) for i in sample])
    want = enet_oracle.forward(w, x, eps)
    want_cls = want.argmax(1)
    big = ENET(WEIGHTS["trained"], device=0, max_batch=256)              # default precision: fp16
    d = torch.from_numpy(frames).cuda()
    ident = np.arange(256, dtype=np.uint8)
    cls = big.predict_device(d, lut=ident).cpu().numpy()                 # raw class map, fused head
    agree = np.array([(cls[i] == want_cls[k]).mean() for k, i in enumerate(sample)])
    scene = np.array([i % 2 == 0 for i in sample])
    print("bs256 raw argmax agreement vs fp32 oracle per sampled frame:", dict(zip(sample, np.round(agree, 5))))
    assert agree[scene].mean() >= 0.999, agree
    assert agree[scene].min() >= 0.998, agree
    # out-of-distribution inputs (blocky / white noise) give the trained net many near-ties: looser, but not chaos
    assert agree[~scene].min() >= 0.97, agree
    # same frames in another batch position and at batch 10 give the same classes (frame independence)
    small = big.predict_device(torch.from_numpy(frames[sample]).cuda(), lut=ident).cpu().numpy()
    assert np.array_equal(small, cls[sample])
    # whole pipeline at bs 256, graphs on and off: labels == the 3-way LUT of those classes, grids bit-exact given them
    c = synth.calibration("A")
    bev = bev_transform_tools(c["input image size"], c["output image size"], c["distance to target"],
                              c["tile_length"], c["cm_per_px"], c["yaw"], c["is_laserscan"])
    bev._bev_matrix = np.asarray(c["bev matrix"]).reshape(3, 3)
    pipe = FramePipeline(big, bev, 10.0, 10.0, 0.1)
    ww, wh = c["output image size"]
    for graphs in (1, 0):
        big.ctx.set_graphs(graphs)
        d_lab = torch.zeros((256, 256, 512), dtype=torch.uint8, device="cuda")
        grids = pipe.run_device(d, d_labels=d_lab).cpu().numpy()
        lab = d_lab.cpu().numpy()
        assert np.array_equal(lab, ENET.LUT_3WAY[cls]), graphs
        for i in sample:
            ref = bev_oracle.occupancy_grid(lab[i], c["bev matrix"], ww, wh, c["cm_per_px"], 10.0, 10.0, 0.1)
            assert np.array_equal(grids[i], ref), (graphs, i)
    big.ctx.set_graphs(1)
    # the blocking host call with sub-batch copy/compute overlap (B >= 32) AFTER the streaming call on the same
    # context: both share the context's copy stream and events (a call order that used to hit a null event)
    pin = torch.from_numpy(frames[:64].copy()).pin_memory()
    out = torch.zeros((64, pipe.Hc, pipe.Wc), dtype=torch.int8).pin_memory()
    big.ctx.pipeline_host_submit(pin, 256, 512, 64, pipe.lut, 10.0, 10.0, 0.1, 0, 0, out, None)
    big.ctx.pipeline_host_wait(0)
    blocking = pipe(frames[:64])
    assert np.array_equal(out.numpy(), grids[:64]) and np.array_equal(blocking, grids[:64])


@pytest.mark.parametrize("B", [193, 199, 211, 233, 255])
def test_odd_batch_sizes_in_the_race_regime(B):
    """Batch sizes 193..255 (where round 1's x-ring race showed up): the batch result equals the same frames
    pushed through in chunks of 7, frame by frame."""
    import torch
    from bugcar_image_segmentation_b200.models import ENET
    m = ENET(WEIGHTS["trained"], device=0, max_batch=256)
    base = np.stack([synth.region_frame(4000 + i)[0] for i in range(8)] + [synth.noise_frame(4100 + i) for i in range(4)])
    idx = (np.arange(B) * 7 + B) % len(base)
    d = torch.from_numpy(base[idx]).cuda()
    ident = np.arange(256, dtype=np.uint8)
    got = m.predict_device(d, lut=ident).cpu().numpy()
    m.ctx.set_chunk(7)
    ref = m.predict_device(torch.from_numpy(base).cuda(), lut=ident).cpu().numpy()
    assert np.array_equal(got, ref[idx])


def test_streaming_host_entry_point(setup):
    """bc_pipeline_host_submit / _wait (two staging slots, copy/compute overlap across steps)
    returns the same grids as the blocking bc_pipeline_host, step after step."""
    import torch
    from bugcar_image_segmentation_b200 import _lib
    from bugcar_image_segmentation_b200.bev import bev_transform_tools
    from bugcar_image_segmentation_b200.pipeline import FramePipeline
    m = setup["model"]
    m.ctx.set_precision(_lib.BC_PREC_FP16)
    c = synth.calibration("A")
    bev = bev_transform_tools(c["input image size"], c["output image size"], c["distance to target"],
                              c["tile_length"], c["cm_per_px"], c["yaw"], c["is_laserscan"])
    bev._bev_matrix = np.asarray(c["bev matrix"]).reshape(3, 3)
    pipe = FramePipeline(m, bev, 10.0, 10.0, 0.1)
    frames = setup["frames"]
    want = [pipe(frames[[i % 3, (i + 1) % 3]]) for i in range(5)]
    ins = [torch.from_numpy(frames[[i % 3, (i + 1) % 3]].copy()).pin_memory() for i in range(5)]
    outs = [torch.zeros((2, pipe.Hc, pipe.Wc), dtype=torch.int8).pin_memory() for _ in range(5)]
    for i in range(5):
        m.ctx.pipeline_host_submit(ins[i], 256, 512, 2, pipe.lut, 10.0, 10.0, 0.1, 0, 0, outs[i], None)
        m.ctx.pipeline_host_wait(1)
        if i >= 1:
            assert np.array_equal(outs[i - 1].numpy(), want[i - 1]), i
    m.ctx.pipeline_host_wait(0)
    assert np.array_equal(outs[4].numpy(), want[4])
