"""The per-frame loop of the reference's (missing) ``inference_video.py``
(README.md:18-20) as one call: frames -> ENet -> labels -> occupancy grids, all on the
GPU through ``bc_pipeline`` / ``bc_pipeline_host`` (one CUDA graph per argument set).

    model = ENET(weights)                       # models.py drop-in
    bev = bev_transform_tools.fromJSON(path)    # bev.py drop-in
    pipe = FramePipeline(model, bev, 10.0, 10.0, 0.1)
    grids = pipe(frames_u8)                     # (B,h,w,3) BGR uint8 -> int8 (B,Hc,Wc)

is the fused equivalent of
    seg = model.predict(np.concatenate([ENET.preprocess(f) for f in frames]))
    grids = [bev.create_occupancy_grid(s, 10.0, 10.0, 0.1) for s in seg]
and, with ``binary=True, contour_filter=True``, of
    seg = model.predict_binary(...)
    grids = [bev.create_occupancy_grid_binary(contour_noise_removal(s), 10.0, 10.0, 0.1) for s in seg]
"""
import numpy as np

from . import runtime


class FramePipeline:
    def __init__(self, model, bev, grid_width_in_m, grid_height_in_m, cell_size_in_m, binary=False,
                 ros_layout=False, contour_filter=False):
        self.laserscan = bool(bev.laserscan_like_occupancy_grid)      # bev.py:145-164 / 216-240, outliers = 0
        if self.laserscan and ros_layout:
            raise NotImplementedError("ros_layout is not available for laserscan-like grids")
        self.model, self.bev = model, bev
        self.ctx = bev._context(model.ctx)              # the model's context now carries the calibration
        self.w_m, self.h_m, self.cell_m = float(grid_width_in_m), float(grid_height_in_m), float(cell_size_in_m)
        self.binary, self.ros_layout = int(bool(binary)), int(bool(ros_layout))
        self.lut = model.LUT_BINARY if binary else model.LUT_3WAY
        if contour_filter and not binary:
            raise ValueError("contour_noise_removal works on the binary road mask (image_processing_utils.py:4): "
                             "use binary=True")
        self.ctx.set_contour_filter(bool(contour_filter))      # a property of the context's binary pipeline
        self.Hc, self.Wc = self.ctx.occgrid_shape(self.w_m, self.h_m, self.cell_m)
        self._torch, self.device = model._torch, model.device
        self._pinned_in = None
        self._pinned_out = None

    @property
    def grid_shape(self):
        return (self.Wc, self.Hc) if self.ros_layout else (self.Hc, self.Wc)

    def run_device(self, d_frames, d_grids=None, d_labels=None, to_gather=False):
        """d_frames: CUDA uint8 (B,h,w,3).  Returns CUDA int8 (B,*grid_shape); asynchronous
        on the current stream.  ``to_gather``: write the grids straight into the peer-mapped
        gather buffer selected with ``sharding.PeerGather.use`` (returns None)."""
        torch = self._torch
        B, h, w, _ = d_frames.shape
        if self.laserscan:
            return self._run_laserscan(d_frames, d_grids, d_labels, to_gather)
        if to_gather:
            if B > self.ctx.max_batch:
                raise ValueError("to_gather needs the batch in one call (B <= max_batch)")
            self.ctx.pipeline(d_frames, h, w, B, self.lut, self.w_m, self.h_m, self.cell_m, self.binary,
                              self.ros_layout, d_labels, None, runtime.stream_handle(torch, self.device))
            return None
        if d_grids is None:
            d_grids = torch.empty((B,) + self.grid_shape, dtype=torch.int8, device=d_frames.device)
        s = runtime.stream_handle(torch, self.device)
        step = self.ctx.max_batch
        for b0 in range(0, B, step):
            n = min(step, B - b0)
            self.ctx.pipeline(d_frames[b0:b0 + n], h, w, n, self.lut, self.w_m, self.h_m, self.cell_m, self.binary,
                              self.ros_layout, None if d_labels is None else d_labels[b0:b0 + n],
                              d_grids[b0:b0 + n], s)
        return d_grids

    def _run_laserscan(self, d_frames, d_grids, d_labels, to_gather):
        """frames -> labels (the fused pipeline, whose ordinary grid is discarded) -> laserscan-like grid.
        With binary=True the result is the laserscan grid, the second element of the reference's tuple."""
        torch = self._torch
        if to_gather:
            raise NotImplementedError("peer-gathered grids are ordinary grids")
        B, h, w, _ = d_frames.shape
        s = runtime.stream_handle(torch, self.device)
        if d_grids is None:
            d_grids = torch.empty((B, self.Hc, self.Wc), dtype=torch.int8, device=d_frames.device)
        if d_labels is None:
            d_labels = torch.empty((B, 256, 512), dtype=torch.uint8, device=d_frames.device)
        step = self.ctx.max_batch
        for b0 in range(0, B, step):
            n = min(step, B - b0)
            self.ctx.pipeline(d_frames[b0:b0 + n], h, w, n, self.lut, self.w_m, self.h_m, self.cell_m, self.binary, 0,
                              d_labels[b0:b0 + n], d_grids[b0:b0 + n], s)
            self.ctx.occgrid_laserscan(d_labels[b0:b0 + n], n, self.w_m, self.h_m, self.cell_m, self.binary, None,
                                       d_grids[b0:b0 + n], s)
        return d_grids

    def __call__(self, frames):
        """frames: uint8 (B,h,w,3) or (h,w,3) BGR host array -> int8 (B,*grid_shape) host
        array.  Stages through pinned buffers: H2D, the graph, D2H, one sync."""
        torch = self._torch
        frames = np.asarray(frames, dtype=np.uint8)
        single = frames.ndim == 3
        if single:
            frames = frames[None]
        B, h, w, _ = frames.shape
        if self._pinned_in is None or self._pinned_in.shape != frames.shape:
            self._pinned_in = torch.empty(frames.shape, dtype=torch.uint8, pin_memory=True)
            self._pinned_out = torch.empty((B,) + self.grid_shape, dtype=torch.int8, pin_memory=True)
        self._pinned_in.numpy()[...] = frames
        if self.laserscan:
            d = self._run_laserscan(self._pinned_in.to(f"cuda:{self.device}", non_blocking=True), None, None, False)
            out = d.cpu().numpy()
            return out[0] if single else out
        s = runtime.stream_handle(torch, self.device)
        step = self.ctx.max_batch
        for b0 in range(0, B, step):
            n = min(step, B - b0)
            self.ctx.pipeline_host(self._pinned_in[b0:b0 + n], h, w, n, self.lut, self.w_m, self.h_m, self.cell_m,
                                   self.binary, self.ros_layout, self._pinned_out[b0:b0 + n], s)
        out = self._pinned_out.numpy().copy()
        return out[0] if single else out
