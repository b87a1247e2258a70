#!/usr/bin/env python
"""Headline benchmark: camera frame -> ENet -> class argmax/LUT -> BEV warp -> occupancy grid,
frames/s (BASELINE.json metric), bs 256 per GPU, fp16 storage / fp32 accumulate on tcgen05 (--precision
bf16 for the bf16 variant), synthetic scene frames, weights = the ENet architecture briefly trained on a
synthetic colour-region task (pretrained_models/enet_synthetic_trained.bcw; --weights seed42 for random init).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One process per GPU (torchrun for N > 1).  A step = one pass of the hot path over one batch
of 256 frames per GPU.  `value` = frames/s with inputs resident in HBM (bc_pipeline);
`e2e` = the same through the host entry point (pinned host frames -> H2D -> graph -> D2H of
the grids).  `roofline` = the dominant kernel, timed with CUDA events by the library's
per-kernel profiler in a separate pass right after the timed region (same workload).
`cpu_baseline` / `--impl reference` time the CPU restatement of the reference path
(torch-fp32 ENet + OpenCV pre/post, what the reference's Python does minus TensorFlow,
which is not installable here) on the box's host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BATCH = 256
GRID = (10.0, 10.0, 0.1)
CAL = "A"
METRIC = "enet_frame_to_occupancy_grid_frames_per_s"
N_INPUT_SETS = 4          # 4 x 100 MB of frames rotate through the steps (> 126 MB L2)
# Weights: the reference's trained blobs are absent, so the default is the ENet architecture
# briefly trained on a synthetic colour-region task (trained-like statistics), fed frames of the
# same kind ("scene"); --weights seed42 --frames noise gives random init on white noise.
WEIGHT_FILES = {"trained": "enet_synthetic_trained.bcw", "seed42": "enet_synthetic_seed42.bcw"}
WEIGHTS, FRAMES = "trained", "scene"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), "measured"
    return 6650.0, 1400.0, "fallback"


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed regions: NVML from a thread every
    5 ms (nvidia-smi -lms as the fallback, its start-up alone outlasts a 100 ms region)."""
    BAD = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, index):
        self.index, self.sm, self.reasons, self.max_sm = index, [], set(), None
        self._stop = threading.Event()
        self.t = None

    def _nvml_loop(self):
        import pynvml as nv
        nv.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = int(vis.split(",")[self.index]) if vis and vis.split(",")[self.index].isdigit() else self.index
        h = nv.nvmlDeviceGetHandleByIndex(idx)
        self.max_sm = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
        while not self._stop.is_set():
            self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
            r = nv.nvmlDeviceGetCurrentClocksEventReasons(h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                else nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
            for name, bit in self.BAD.items():
                if r & bit:
                    self.reasons.add(name)
            time.sleep(0.005)

    def _smi_loop(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout
            except (OSError, subprocess.SubprocessError):
                return
            c = [x.strip() for x in out.strip().split(",")]
            if len(c) >= 6 and c[0].replace(".", "").isdigit():
                self.sm.append(float(c[0]))
                self.max_sm = float(c[1])
                self.reasons |= {names[i] for i in range(4) if c[2 + i] == "Active"}

    def _run(self):
        try:
            self._nvml_loop()
        except Exception:
            self._smi_loop()

    def start(self):
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()

    def stop(self):
        self._stop.set()
        if self.t:
            self.t.join(timeout=6)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_sm,
                "reasons": sorted(self.reasons), "samples": len(self.sm)}


def base_frames(rank, k):
    """the 32 distinct frames of rank `rank`'s input set k; frame i of the batch is base[i % 32]"""
    from bugcar_image_segmentation_b200 import synth
    s0 = 1234 + k * 100000 + rank * BATCH
    if FRAMES == "scene":
        return np.stack([synth.region_frame(s0 + i)[0] for i in range(32)])
    return synth.frames(32, s0)


def make_frames(rank, n_sets):
    """frame i of rank r in set k: seed 1234 + k*100000 + r*BATCH + i.
    Generating 1024 seeded frames with NumPy takes a while; build 32 per set and tile."""
    return [np.ascontiguousarray(np.tile(base_frames(rank, k), (BATCH // 32, 1, 1, 1))) for k in range(n_sets)]


# SURVEY.md 8(a) block-I/O model: bytes per frame a launch must move (block input once + block output once, 16-bit
# activations, 1-byte pool indices); an asymmetric block's I/O is booked on its second launch.  Sum = 42.3 MB/frame.
BLOCKIO_BYTES_PER_FRAME = {
    "umma_initial": 393216 + 1048576, "umma_pool_conv16": 0, "umma_down64": 1048576 + 1048576 + 131072,
    "umma_bottleneck64": 2097152, "umma_pool_conv64": 0, "umma_down128": 1048576 + 524288 + 131072,
    "umma_bottleneck128": 1048576, "umma_conv5x1": 0, "umma_bottleneck128_asym": 1048576, "umma_asym_fused": 1048576,
    "umma_up4": 524288 + 1048576 + 131072, "umma_up5": 1048576 + 1048576 + 131072, "stage5_bottleneck": 2097152,
    "umma_head_argmax_lut": 1048576 + 131072, "occgrid": 131072 + 10000,
}


def workload_config(B):
    return {"workload": f"full pipeline frame->ENet->argmax/LUT->BEV grid, bs {B} per GPU, 256x512 BGR frames, "
                        f"15 classes, calibration {CAL} (500x500 warp), grid 10x10 m @ 0.1 m",
            "weights": WEIGHT_FILES[WEIGHTS] + (" (ENet architecture, briefly trained on synthetic colour-region scenes)"
                                                if WEIGHTS == "trained" else " (random init, BN calibrated)"),
            "frames": "synthetic colour-region scenes + noise" if FRAMES == "scene" else "uniform noise / blocky"}


def cpu_reference_step(w, eps, frames, cal, torch_threads, classes=None):
    """the reference's per-frame path on the CPU: preprocess (models.py:84-95), ENet forward
    (torch fp32 stand-in for sess.run, models.py:43-44), argmax + LUT (models.py:55-67),
    create_occupancy_grid (bev.py:166-246, OpenCV back end)."""
    from oracle import pre_oracle, enet_oracle, bev_oracle
    ww, wh = cal["output image size"]
    out = []
    for f in frames:
        x = pre_oracle.preprocess(f, backend="cv2")
        lg = enet_oracle.forward(w, x, eps)
        if classes is not None:
            classes.append(lg[0].argmax(0).astype(np.uint8))        # models.py:55, for the agreement check
        lab = pre_oracle.labels_from_logits(lg, pre_oracle.LUT_3WAY)
        out.append(bev_oracle.occupancy_grid(lab[0], cal["bev matrix"], ww, wh, cal["cm_per_px"], *GRID, backend="cv2"))
    return out


def run_reference(args):
    """--impl reference: the CPU restatement on all host threads; rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from bugcar_image_segmentation_b200 import synth, weights as W
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    with open(os.path.join(ROOT, "pretrained_models", WEIGHT_FILES[WEIGHTS]), "rb") as f:
        w, nc, eps = W.unpack_flat(f.read())
    cal = synth.calibration(CAL)
    per_step = 16                                  # bounded sample of the 256-frame batch
    frames = make_frames(0, 1)[0][:per_step]
    for _ in range(args.warmup):
        cpu_reference_step(w, eps, frames[:1], cal, threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_reference_step(w, eps, frames, cal, threads)
    dt = time.perf_counter() - t0
    fps = per_step * args.steps / dt
    sample = f"{per_step} frames/step of the {BATCH}-frame batch, bs 1 per call as the reference's loop does"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": dict(workload_config(BATCH),
                       reference_arm="CPU restatement of the reference path (reference preprocess semantics, torch fp32 "
                                     "ENet, NumPy argmax/LUT, OpenCV create_occupancy_grid); TensorFlow/Keras and the "
                                     "reference's model blobs are not available here"),
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    global WEIGHTS, FRAMES
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--chunk", type=int, default=0)
    ap.add_argument("--no-tc", action="store_true")
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-latency", action="store_true")
    ap.add_argument("--skip-contour", action="store_true")
    ap.add_argument("--skip-config5", action="store_true")
    ap.add_argument("--nccl-gather", action="store_true")
    ap.add_argument("--weights", default=WEIGHTS, choices=sorted(WEIGHT_FILES))
    ap.add_argument("--frames", default=FRAMES, choices=["scene", "noise"])
    ap.add_argument("--precision", default="fp16", choices=["fp16", "bf16"])
    ap.add_argument("--wc-staging", action="store_true", help="frame staging buffers in write-combined pinned memory (bc_host_alloc)")
    args = ap.parse_args()
    WEIGHTS, FRAMES = args.weights, args.frames
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    from bugcar_image_segmentation_b200 import synth, sharding, weights as W
    from bugcar_image_segmentation_b200.models import ENET
    from bugcar_image_segmentation_b200.bev import bev_transform_tools
    from bugcar_image_segmentation_b200.pipeline import FramePipeline

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    from bugcar_image_segmentation_b200 import runtime
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # which GPU this rank takes: the local rank, unless the node has more GPUs than the job has ranks -- then the ranks
    # are spread over the node's two host domains (runtime.device_for_rank: 4 ranks on GPUs 0-3 share 116 GB/s of
    # H2D bandwidth, on GPUs 0, 4, 1, 5 they get 218 GB/s)
    local = runtime.device_for_rank(local_rank, world)
    torch.cuda.set_device(local)
    host_bind = runtime.bind_host_to_gpu(local)     # before any pinned allocation: staging buffers local to the GPU's NUMA node
    # stdout carries exactly one JSON line: libraries that write to fd 1 (NCCL prints its version banner there
    # when NCCL_DEBUG is set) go to stderr until the result is printed
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    B = args.batch

    wpath = os.path.join(ROOT, "pretrained_models", WEIGHT_FILES[WEIGHTS])
    model = ENET(wpath, device=local, max_batch=B, precision=args.precision)
    if args.chunk:
        model.ctx.set_chunk(args.chunk)
    if args.no_tc:
        model.ctx.set_tensor_cores(0)
    cal = synth.calibration(CAL)
    bev = bev_transform_tools(cal["input image size"], cal["output image size"], cal["distance to target"],
                              cal["tile_length"], cal["cm_per_px"], cal["yaw"], cal["is_laserscan"])
    bev._bev_matrix = np.asarray(cal["bev matrix"]).reshape(3, 3)
    pipe = FramePipeline(model, bev, *GRID)
    Hc, Wc = pipe.Hc, pipe.Wc

    host_sets = make_frames(rank, N_INPUT_SETS)
    if B != BATCH:
        host_sets = [s[:B] for s in host_sets]
    if args.wc_staging:
        from bugcar_image_segmentation_b200 import _lib
        wc_bufs = [_lib.HostBuffer(s.shape, np.uint8, write_combined=True) for s in host_sets]
        for b, s in zip(wc_bufs, host_sets):
            b.array[...] = s
        pinned = wc_bufs                                   # data_ptr() -> the C ABI takes them as they are
        dev_sets = [torch.from_numpy(s).cuda() for s in host_sets]
    else:
        pinned = [torch.from_numpy(s).pin_memory() for s in host_sets]
        dev_sets = [p.cuda(non_blocking=True) for p in pinned]
    d_grids = torch.empty((B, Hc, Wc), dtype=torch.int8, device="cuda")
    pinned_out = torch.empty((world * B if rank == 0 else B, Hc, Wc), dtype=torch.int8).pin_memory()
    stream = torch.cuda.current_stream()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # N > 1: the grids of every rank land in rank 0's buffer through K9's own peer stores
    # (sharding.PeerGather); --nccl-gather switches back to one NCCL gather per step
    peer = None
    if world > 1 and not args.nccl_gather:
        peer = sharding.PeerGather(model.ctx, rank, world, B, (Hc, Wc), local)

    def step_device(i):
        if peer is not None:
            peer.use(i)
            pipe.run_device(dev_sets[i % N_INPUT_SETS], to_gather=True)
            return peer.ready()
        pipe.run_device(dev_sets[i % N_INPUT_SETS], d_grids)
        if world > 1:
            return sharding.gather_grids(d_grids, rank, world)
        return d_grids

    pinned_out2 = [pinned_out, torch.empty_like(pinned_out).pin_memory()]

    def step_e2e(i):
        # the library's streaming host entry point at every N: step i's H2D (internal copy stream) overlaps step
        # i-1's kernels; returns when step i-1's grids are on the host (one step in flight).  N > 1: the same two
        # calls per rank -- K9 stores into rank 0's gather buffer, flags in peer-mapped memory order the ranks,
        # rank 0's context copies each complete step to ITS host buffer (sharding.StreamingGather)
        t0 = time.perf_counter()
        model.ctx.pipeline_host_submit(pinned[i % N_INPUT_SETS], 256, 512, B, pipe.lut, *GRID, 0, 0,
                                       pinned_out2[i & 1] if rank == 0 else None, stream.cuda_stream)
        t1 = time.perf_counter()
        model.ctx.pipeline_host_wait(1)
        trace.append((t1 - t0, time.perf_counter() - t1))

    trace = []      # per step: host seconds inside submit / inside wait (BC_E2E_TRACE=1 prints a per-rank summary)

    def timed(fn, steps, finish=None):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(steps):
            fn(i)
        if finish is not None:
            finish()                # e.g. the last step's grids must be on the host before the clock stops
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    # ---- device-resident throughput
    for i in range(args.warmup):
        step_device(i)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = model.ctx.launch_count()
    ms = timed(step_device, args.steps)
    launches = model.ctx.launch_count() - l0
    value = world * B * args.steps / (ms / 1e3)

    # ---- end to end through the host entry point
    if world == 1:
        for i in range(args.warmup):
            step_e2e(i)
    streaming = None
    if world > 1:
        if peer is not None:
            peer.close()
        streaming = sharding.StreamingGather(model.ctx, rank, world, B, (Hc, Wc), local)
        for i in range(args.warmup):
            step_e2e(i)
        model.ctx.pipeline_host_wait(0)
    ms_e2e = timed(step_e2e, args.steps, finish=lambda: model.ctx.pipeline_host_wait(0))
    if os.environ.get("BC_E2E_TRACE"):
        tr = np.array(trace[-args.steps:])
        print(f"[rank {rank}] e2e host time per step: submit {tr[:, 0].mean() * 1e3:.3f} ms (max {tr[:, 0].max() * 1e3:.3f}), "
              f"wait {tr[:, 1].mean() * 1e3:.3f} ms (max {tr[:, 1].max() * 1e3:.3f})", file=sys.stderr, flush=True)
    clocks = sampler.stop() if rank == 0 else None      # sampled over both timed regions
    e2e = world * B * args.steps / (ms_e2e / 1e3)
    # the ceiling the end-to-end number lives under: every rank's bare pinned-host -> device copies of the same
    # frame batches, all ranks at once, nothing else running (bytes/s summed over the ranks)
    h2d_stage = torch.empty_like(dev_sets[0])
    if args.wc_staging:
        import ctypes
        _rt = ctypes.CDLL("libcudart.so.12")
        h2d_only = lambda i: _rt.cudaMemcpyAsync(ctypes.c_void_p(h2d_stage.data_ptr()), ctypes.c_void_p(pinned[i % N_INPUT_SETS].data_ptr()),
                                                 ctypes.c_size_t(B * 393216), 1, ctypes.c_void_p(stream.cuda_stream))
    else:
        h2d_only = lambda i: h2d_stage.copy_(pinned[i % N_INPUT_SETS], non_blocking=True)
    for i in range(3):
        h2d_only(i)
    ms_h2d = timed(h2d_only, args.steps)
    h2d_ceiling_gbs = world * B * 393216 * args.steps / (ms_h2d * 1e-3) / 1e9
    del h2d_stage
    barrier()
    step_e2e(0)                                          # untimed: grids of input set 0 for the cross-checks
    model.ctx.pipeline_host_wait(0)
    torch.cuda.synchronize()
    e2e_out = pinned_out2[0]                             # step_e2e(0) wrote slot 0
    grids_check = e2e_out.numpy()[:B].copy()
    if streaming is not None:
        streaming.close()

    # ---- N > 1: what rank 0 holds after the gather == what one GPU computes for the same frames.  Rank 0
    # regenerates 8 sampled frames of EVERY rank's input set 0 (seeds are a function of the rank), pushes them
    # through its own pipeline and compares with the rows of the gathered host buffer.
    gather_check = None
    if world > 1:
        if rank == 0:
            gathered_host = e2e_out.numpy()
            idx = np.array([0, 31, 32, 63, 128, 200, 254, 255]) % B
            ok = True
            for r in range(world):
                fr = base_frames(r, 0)[idx % 32]
                want = pipe.run_device(torch.from_numpy(np.ascontiguousarray(fr)).cuda()).cpu().numpy()
                ok = ok and bool(np.array_equal(gathered_host[r * B + idx], want))
            gather_check = ok
        barrier()

    # ---- per-kernel profile (events around every launch), rank 0 at any N
    roofline, kernels = None, None
    if rank == 0:
        hbm, tf, which = peaks()
        model.ctx.set_profile(True)
        for i in range(max(2, min(args.steps, 5))):
            pipe.run_device(dev_sets[i % N_INPUT_SETS], d_grids)
        kernels = model.ctx.profile()
        model.ctx.set_profile(False)
        tot = sum(k["ms"] for k in kernels)
        for k in kernels:
            k["share"] = k["ms"] / tot
            k["gbs"] = k["bytes"] / (k["ms"] * 1e-3) / 1e9 if k["ms"] > 0 else 0.0
            k["tflops"] = k["flops"] / (k["ms"] * 1e-3) / 1e12 if k["ms"] > 0 else 0.0
        kernels.sort(key=lambda k: -k["ms"])
        top = kernels[0]
        traffic = None      # dram bytes per launch from the committed ncu --set full capture (profiles/traffic.json)
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.isfile(tp):
            tj = json.load(open(tp))
            for k in kernels:
                k["dram_bytes_per_launch_ncu"] = tj.get(k["kernel"], {}).get("dram_bytes_per_launch")
            traffic = top.get("dram_bytes_per_launch_ncu")
        for k in kernels:
            k["blockio_bytes_per_launch"] = BLOCKIO_BYTES_PER_FRAME.get(k["kernel"], 0) * B
        top_us = top["ms"] * 1e3 / top["launches"]
        blockio_gbs = top["blockio_bytes_per_launch"] / (top_us * 1e-6) / 1e9
        step_s = ms / args.steps * 1e-3                 # the TIMED device-resident step, not the profiling pass
        n_prof_steps = max(2, min(args.steps, 5))
        own_bytes_per_step = sum(k["bytes"] for k in kernels) / n_prof_steps
        blockio_per_step = sum(k["blockio_bytes_per_launch"] * k["launches"] for k in kernels) / n_prof_steps
        # `achieved` / `frac`: SURVEY 8(a)/(d) block-I/O bytes (x in + y out) of one launch / its average duration.
        # `frac_own_model` also counts the quarter-width e1 in / e1' out that exist because the block spans two
        # kernels; `dram_frac` uses the DRAM bytes ncu measured for this kernel (profiles/traffic.json).
        roofline = {"kernel": top["kernel"], "bound": "hbm", "achieved": blockio_gbs, "peak": hbm, "unit": "GB/s",
                    "frac": blockio_gbs / hbm, "frac_blockio": blockio_gbs / hbm, "frac_own_model": top["gbs"] / hbm,
                    "dram_frac": (traffic / (top_us * 1e-6) / 1e9 / hbm) if traffic else None,
                    "traffic": traffic, "peak_source": which,
                    "share_of_step": top["share"], "avg_launch_us": top_us,
                    "algorithmic_bytes_per_launch": top["blockio_bytes_per_launch"],
                    "own_model_bytes_per_launch": top["bytes"] / top["launches"],
                    "whole_path": {"timed_step_ms": ms / args.steps,
                                   "blockio_mb_per_frame": blockio_per_step / B / 1e6,
                                   "blockio_gbs": blockio_per_step / step_s / 1e9,
                                   "frac_blockio": blockio_per_step / step_s / 1e9 / hbm,
                                   "own_model_mb_per_frame": own_bytes_per_step / B / 1e6,
                                   "frac_own_model": own_bytes_per_step / step_s / 1e9 / hbm,
                                   "tflops": sum(k["flops"] for k in kernels) / n_prof_steps / step_s / 1e12}}

    # ---- batch-1 streaming latency (BASELINE config 3) on rank 0 of a 1-GPU run
    latency = None
    if rank == 0 and world == 1 and not args.skip_latency:
        one_in = torch.from_numpy(host_sets[0][:1].copy()).pin_memory()
        one_out = torch.empty((1, Hc, Wc), dtype=torch.int8).pin_memory()
        ts = []
        for i in range(220):
            one_in.numpy()[0, 0, 0, 0] = i % 251
            t0 = time.perf_counter()
            model.ctx.pipeline_host(one_in, 256, 512, 1, pipe.lut, *GRID, 0, 0, one_out, stream.cuda_stream)
            ts.append((time.perf_counter() - t0) * 1e3)
        ts = np.array(ts[20:])
        latency = {"p50_ms": float(np.percentile(ts, 50)), "p99_ms": float(np.percentile(ts, 99)),
                   "mean_ms": float(ts.mean()), "frames": int(ts.size),
                   "how": "bc_pipeline_host, bs 1, pinned host frame -> grid on host, wall clock"}

    # ---- predict_binary -> contour_noise_removal -> grid (binary pipeline with the filter on)
    contour = None
    if rank == 0 and world == 1 and not args.skip_contour:
        from oracle import contour_oracle
        lut2 = np.zeros(256, np.uint8)
        lut2[[0, 1]] = 1                                   # models.py:79-80
        d_lab = torch.empty((B, 256, 512), dtype=torch.uint8, device="cuda")
        d_filt = torch.empty_like(d_lab)
        res = {}
        for on in (0, 1):
            model.ctx.set_contour_filter(on)
            run = lambda i: model.ctx.pipeline(dev_sets[i % N_INPUT_SETS], 256, 512, B, lut2, *GRID, 1, 0, d_lab, d_grids,
                                               stream.cuda_stream)
            for i in range(3):
                run(i)
            res[on] = B * 10 / (timed(run, 10) / 1e3)
        model.ctx.set_contour_filter(0)
        run_f = lambda i: model.ctx.contour_noise_removal(d_lab, 256, 512, B, d_filt, stream.cuda_stream)
        for i in range(3):
            run_f(i)
        ms_f = timed(run_f, 10) / 10
        masks = d_lab[:8].cpu().numpy()
        t0 = time.perf_counter()
        want = [contour_oracle.contour_noise_removal_cv2(m) for m in masks]
        cpu_ms = (time.perf_counter() - t0) / len(masks) * 1e3
        exact = bool(np.array_equal(np.stack(want), d_filt[:8].cpu().numpy()))
        contour = {"binary_pipeline_frames_per_s": res[0], "binary_pipeline_with_filter_frames_per_s": res[1],
                   "filter_alone_ms_per_batch": ms_f, "filter_alone_frames_per_s": B / ms_f * 1e3,
                   "filter_algorithmic_gbs": 2 * B * 131072 / (ms_f * 1e-3) / 1e9,
                   "cpu_opencv_ms_per_frame": cpu_ms, "bit_exact_vs_opencv_on_8_frames": exact,
                   "masks": "predict_binary labels of the benchmark batch"}

    # ---- laserscan-like grids from the benchmark's labels (deterministic re-specification)
    laser = None
    if rank == 0 and world == 1 and not args.skip_contour:
        from oracle import bev_oracle, laser_oracle
        d_lab3 = torch.empty((B, 256, 512), dtype=torch.uint8, device="cuda")
        pipe.run_device(dev_sets[0], d_grids, d_labels=d_lab3)
        d_laser = torch.empty_like(d_grids)
        run_l = lambda i: model.ctx.occgrid_laserscan(d_lab3, B, *GRID, 0, None, d_laser, stream.cuda_stream)
        for i in range(3):
            run_l(i)
        ms_l = timed(run_l, 10) / 10
        labs = d_lab3[:4].cpu().numpy()
        t0 = time.perf_counter()
        want_l = []
        for m in labs:
            _, templ = bev_oracle.occupancy_grid(m, cal["bev matrix"], *cal["output image size"], cal["cm_per_px"], *GRID,
                                                 backend="cv2", return_template=True)
            want_l.append(laser_oracle.laserscan_3way(templ))
        cpu_l = (time.perf_counter() - t0) / len(labs) * 1e3
        laser = {"what": "labels -> laserscan-like grid (bc_occgrid_laserscan, three-way), bs %d" % B,
                 "frames_per_s": B / ms_l * 1e3, "ms_per_batch": ms_l, "cpu_port_ms_per_frame": cpu_l,
                 "bit_exact_vs_cpu_on_4_frames": bool(np.array_equal(np.stack(want_l), d_laser[:4].cpu().numpy()))}

    # ---- BASELINE config 5 (DeepLab 720p, bs 32): the forward is blocked (no graph, no weights), so only the
    # post-processing is measured: label maps (32,720,1280) -> grids with a 720p calibration
    config5 = None
    if rank == 0 and world == 1 and not args.skip_config5:
        from bugcar_image_segmentation_b200 import _lib
        from oracle import bev_oracle
        c5 = _lib.Context(local, 32)
        calE = synth.calibration("E", 720, 1280)
        wwE, whE = calE["output image size"]
        c5.set_bev(calE["bev matrix"], 720, 1280, wwE, whE, calE["cm_per_px"])
        hE, wE = c5.occgrid_shape(*GRID)
        lab5 = np.stack([synth.label_map(500 + i, 3, 720, 1280) for i in range(32)])
        d_lab5 = torch.from_numpy(lab5).cuda()
        d_g5 = torch.empty((32, hE, wE), dtype=torch.int8, device="cuda")
        run5 = lambda i: c5.occgrid(d_lab5, 32, *GRID, 0, 0, d_g5, stream.cuda_stream)
        for i in range(3):
            run5(i)
        ms5 = timed(run5, 20) / 20
        t0 = time.perf_counter()
        want5 = [bev_oracle.occupancy_grid(lab5[i], calE["bev matrix"], wwE, whE, calE["cm_per_px"], *GRID, backend="cv2")
                 for i in range(4)]
        cpu5 = (time.perf_counter() - t0) / 4 * 1e3
        config5 = {"what": "post-processing only (K9) on synthetic label maps (32,720,1280), calibration E",
                   "frames_per_s": 32 / ms5 * 1e3, "ms_per_batch": ms5, "cpu_opencv_ms_per_frame": cpu5,
                   "bit_exact_vs_cpu_on_4_frames": bool(np.array_equal(np.stack(want5), d_g5[:4].cpu().numpy()))}
        c5.close()

    # ---- CPU baseline on a bounded sample (rank 0, N = 1)
    cpu = None
    if rank == 0 and world == 1 and not args.skip_cpu:
        threads = len(os.sched_getaffinity(0)) or 1       # the cores this process may use (NUMA binding above)
        torch.set_num_threads(threads)
        with open(wpath, "rb") as f:
            w, nc, eps = W.unpack_flat(f.read())
        cpu_reference_step(w, eps, host_sets[0][:1], cal, threads)
        ref_grids, ref_cls, t0, budget_s = [], [], time.perf_counter(), 12.0
        for i in range(B):                           # bounded sample: frames of the batch until ~12 s of CPU work
            ref_grids += cpu_reference_step(w, eps, host_sets[0][i:i + 1], cal, threads, ref_cls)
            if time.perf_counter() - t0 > budget_s:
                break
        dt = time.perf_counter() - t0
        nref = len(ref_grids)
        same = float(np.mean([np.mean(ref_grids[i] == grids_check[i]) for i in range(nref)]))
        # RAW per-pixel class agreement of the benchmarked path (fused head, identity LUT) with the fp32 CPU network
        gpu_cls = model.predict_device(dev_sets[0][:nref], lut=np.arange(256, dtype=np.uint8)).cpu().numpy()
        argmax_agree = float(np.mean(gpu_cls == np.stack(ref_cls)))
        cpu = {"value": nref / dt, "unit": "frames/s", "cores": threads, "kind": "port",
               "sample": f"first {nref} frames of the batch ({dt:.1f} s), per-frame calls as the reference's loop makes them "
                         "(torch fp32 ENet + OpenCV pre/post)",
               "argmax_agreement_vs_fp32": argmax_agree, "grid_cell_agreement_vs_cpu_fp32": same,
               "agreement_note": f"{args.precision} storage + tcgen05 on the GPU vs torch fp32 on the CPU, every pixel / cell "
                                 "of the sampled frames counted (no margin filter)"}

    sys.stdout.flush()
    os.dup2(real_stdout, 1)
    os.close(real_stdout)
    if rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
            "argmax_agreement_vs_fp32": cpu["argmax_agreement_vs_fp32"] if cpu else None,
            "config": dict(workload_config(B),
                       l2=f"{N_INPUT_SETS} input sets x {B * 393216 / 1e6:.0f} MB rotate (> 126 MB L2)",
                       parallelism=f"frame-sharded dp{world}" + ("" if world == 1 else ", grids gathered to rank 0 (NCCL gather)"
                                                                 if peer is None else ", grids stored by K9 into rank 0's "
                                                                 "peer-mapped buffer (NVLink), one barrier per step"),
                       chunk=args.chunk, tensor_cores=not args.no_tc,
                       devices=[runtime.device_for_rank(r, world) for r in range(world)]),
            "e2e": {"value": e2e, "unit": "frames/s", "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": B * 393216, "d2h_bytes_per_step": (world if world > 1 else 1) * B * Hc * Wc,
                    "api": "bc_pipeline_host_submit / bc_pipeline_host_wait on every rank" +
                           ("" if world == 1 else " + bc_gather_stream_setup (grids peer-stored into rank 0's buffer, flag hand-over, "
                                                  "D2H by rank 0's context)"),
                    "h2d_ceiling_gbs": h2d_ceiling_gbs, "h2d_needed_gbs": e2e * 393216 / 1e9,
                    "e2e_frac_of_h2d_ceiling": e2e * 393216 / 1e9 / h2d_ceiling_gbs,
                    "host_numa": host_bind, "staging": "write-combined pinned" if args.wc_staging else "pinned"},
            "gpu_launches": int(launches), "gather_check": gather_check, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
            "latency_bs1": latency, "contour_filter": contour, "laserscan": laser, "config5_postprocessing": config5,
            "kernels": kernels,
        }))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
