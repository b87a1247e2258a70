"""ctypes binding of ``libbugcar_b200.so`` (the C ABI in ``include/bugcar_b200.h``).

There is no CPU fallback: if the shared library is missing or no B200 is visible the
product path raises, loudly.  Loading the library itself does not need a GPU (the
``not gpu`` tests check the exported symbols); creating a context does.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BUGCAR_B200_LIB") or os.path.join(_HERE, "libbugcar_b200.so")     # override: A/B builds

BC_OK, BC_ERR_ARG, BC_ERR_STATE, BC_ERR_CUDA, BC_ERR_FORMAT, BC_ERR_NOMEM = 0, -1, -2, -3, -4, -5
BC_IN_BGR_U8, BC_IN_NCHW_F32, BC_IN_NCHW_F64 = 0, 1, 2
BC_PREC_BF16, BC_PREC_FP32, BC_PREC_FP16 = 0, 1, 2
PRECISIONS = {"fp16": BC_PREC_FP16, "float16": BC_PREC_FP16, "half": BC_PREC_FP16, "bf16": BC_PREC_BF16,
              "bfloat16": BC_PREC_BF16, "fp32": BC_PREC_FP32, "float32": BC_PREC_FP32}
NET_H, NET_W = 256, 512          # models.py:19

_vp, _i, _d, _sz = C.c_void_p, C.c_int, C.c_double, C.c_size_t

# name -> (restype, argtypes); exactly the entry points declared in include/bugcar_b200.h
SIGNATURES = {
    "bc_create": (_i, [C.POINTER(_vp), _i, _i]),
    "bc_destroy": (None, [_vp]),
    "bc_last_error": (C.c_char_p, [_vp]),
    "bc_abi_version": (_i, []),
    "bc_load_enet": (_i, [_vp, _vp, _sz]),
    "bc_num_classes": (_i, [_vp]),
    "bc_set_precision": (_i, [_vp, _i]),
    "bc_set_chunk": (_i, [_vp, _i]),
    "bc_set_tensor_cores": (_i, [_vp, _i]),
    "bc_set_graphs": (_i, [_vp, _i]),
    "bc_set_host_overlap": (_i, [_vp, _i]),
    "bc_set_bev": (_i, [_vp, C.POINTER(_d), _i, _i, _i, _i, _d]),
    "bc_resize_bgr": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp]),
    "bc_preprocess": (_i, [_vp, _vp, _i, _i, _i, _vp, _i, _vp]),
    "bc_enet_logits": (_i, [_vp, _vp, _i, _i, _vp, _vp]),
    "bc_enet_block_output": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp]),
    "bc_enet_labels": (_i, [_vp, _vp, _i, _i, _vp, _vp, _vp]),
    "bc_argmax_lut": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp]),
    "bc_occgrid_laserscan": (_i, [_vp, _vp, _i, _d, _d, _d, _i, _vp, _vp, _vp]),
    "bc_laser_tables": (_i, [_i, _i, _i, C.POINTER(_i), C.POINTER(_i), _vp, _vp]),
    "bc_contour_noise_removal": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp]),
    "bc_set_contour_filter": (_i, [_vp, _i]),
    "bc_occgrid_shape": (_i, [_vp, _d, _d, _d, C.POINTER(_i), C.POINTER(_i)]),
    "bc_occgrid": (_i, [_vp, _vp, _i, _d, _d, _d, _i, _i, _vp, _vp]),
    "bc_pipeline": (_i, [_vp, _vp, _i, _i, _i, _vp, _d, _d, _d, _i, _i, _vp, _vp, _vp]),
    "bc_pipeline_host": (_i, [_vp, _vp, _i, _i, _i, _vp, _d, _d, _d, _i, _i, _vp, _vp]),
    "bc_pipeline_host_submit": (_i, [_vp, _vp, _i, _i, _i, _vp, _d, _d, _d, _i, _i, _vp, _vp]),
    "bc_pipeline_host_wait": (_i, [_vp, _i]),
    "bc_gather_setup": (_i, [_vp, _vp, _i, _i]),
    "bc_gather_stream_setup": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i]),
    "bc_host_alloc": (_i, [C.POINTER(_vp), _sz, _i]),
    "bc_host_free": (_i, [_vp]),
    "bc_launch_count": (C.c_longlong, [_vp]),
    "bc_set_profile": (_i, [_vp, _i]),
    "bc_profile_json": (C.c_char_p, [_vp]),
}

_lib = None


class BugcarError(RuntimeError):
    """A C-ABI call returned a negative bc_status."""

    def __init__(self, code, msg):
        super().__init__(f"libbugcar_b200: {msg} (status {code})")
        self.code = code


def load():
    """dlopen the library once and attach the prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C bugcar_image_segmentation_b200/csrc`).  There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)       # AttributeError if the .so does not export it
        fn.restype = res
        fn.argtypes = args
    if lib.bc_abi_version() != 1:
        raise ImportError("libbugcar_b200.so: ABI version mismatch")
    _lib = lib
    return lib


def _ptr(x):
    """device/host address of a torch tensor, numpy array, int or None"""
    if x is None:
        return None
    if isinstance(x, int):
        return x
    if hasattr(x, "data_ptr"):
        return x.data_ptr()
    if hasattr(x, "ctypes"):
        return x.ctypes.data
    raise TypeError(type(x))


class Context:
    """One ``bc_ctx``: one GPU, one thread.  Thin 1:1 wrapper; arguments are raw
    addresses (``tensor.data_ptr()`` / ``ndarray.ctypes.data``) or the objects
    themselves."""

    def __init__(self, device=0, max_batch=1):
        self.lib = load()
        h = _vp()
        rc = self.lib.bc_create(C.byref(h), int(device), int(max_batch))
        if rc != BC_OK:
            raise BugcarError(rc, self.lib.bc_last_error(None).decode())
        self.h = h
        self.device = int(device)
        self.max_batch = int(max_batch)

    def close(self):
        if getattr(self, "h", None):
            self.lib.bc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != BC_OK:
            raise BugcarError(rc, self.lib.bc_last_error(self.h).decode())

    # ---- model + calibration
    def load_enet(self, blob: bytes):
        buf = (C.c_char * len(blob)).from_buffer_copy(blob)
        self._ck(self.lib.bc_load_enet(self.h, C.addressof(buf), len(blob)))

    def num_classes(self):
        n = self.lib.bc_num_classes(self.h)
        if n < 0:
            raise BugcarError(n, "weights not loaded")
        return n

    def set_precision(self, p):
        self._ck(self.lib.bc_set_precision(self.h, int(p)))

    def set_chunk(self, n):
        self._ck(self.lib.bc_set_chunk(self.h, int(n)))

    def set_tensor_cores(self, on):
        self._ck(self.lib.bc_set_tensor_cores(self.h, int(bool(on))))

    def set_graphs(self, on):
        self._ck(self.lib.bc_set_graphs(self.h, int(bool(on))))

    def set_host_overlap(self, on):
        self._ck(self.lib.bc_set_host_overlap(self.h, int(bool(on))))

    def set_bev(self, M, in_rows, in_cols, warp_w, warp_h, cm_per_px):
        m = (_d * 9)(*[float(v) for v in M])
        self._ck(self.lib.bc_set_bev(self.h, m, int(in_rows), int(in_cols), int(warp_w), int(warp_h),
                                     float(cm_per_px)))

    # ---- stages
    def resize_bgr(self, d_src, h, w, B, d_dst, stream=None):
        self._ck(self.lib.bc_resize_bgr(self.h, _ptr(d_src), h, w, B, _ptr(d_dst), _ptr(stream)))

    def preprocess(self, d_bgr, h, w, B, d_out, out_f64, stream=None):
        self._ck(self.lib.bc_preprocess(self.h, _ptr(d_bgr), h, w, B, _ptr(d_out), int(out_f64), _ptr(stream)))

    def enet_logits(self, d_x, kind, B, d_logits, stream=None):
        self._ck(self.lib.bc_enet_logits(self.h, _ptr(d_x), kind, B, _ptr(d_logits), _ptr(stream)))

    def enet_block_output(self, d_x, kind, B, block, d_out, stream=None):
        self._ck(self.lib.bc_enet_block_output(self.h, _ptr(d_x), kind, B, int(block), _ptr(d_out), _ptr(stream)))

    def enet_labels(self, d_x, kind, B, lut, d_labels, stream=None):
        self._ck(self.lib.bc_enet_labels(self.h, _ptr(d_x), kind, B, _lut(lut), _ptr(d_labels), _ptr(stream)))

    def argmax_lut(self, d_logits, B, Cn, H, W, lut, d_labels, stream=None):
        self._ck(self.lib.bc_argmax_lut(self.h, _ptr(d_logits), B, Cn, H, W, _lut(lut), _ptr(d_labels),
                                        _ptr(stream)))

    def occgrid_laserscan(self, d_labels, B, w_m, h_m, cell_m, binary, d_grid_plain, d_grid_laser, stream=None):
        self._ck(self.lib.bc_occgrid_laserscan(self.h, _ptr(d_labels), B, float(w_m), float(h_m), float(cell_m),
                                               int(binary), _ptr(d_grid_plain), _ptr(d_grid_laser), _ptr(stream)))

    def contour_noise_removal(self, d_seg, H, W, B, d_out, stream=None):
        self._ck(self.lib.bc_contour_noise_removal(self.h, _ptr(d_seg), int(H), int(W), int(B), _ptr(d_out),
                                                   _ptr(stream)))

    def set_contour_filter(self, on):
        self._ck(self.lib.bc_set_contour_filter(self.h, int(bool(on))))

    def occgrid_shape(self, w_m, h_m, cell_m):
        hc, wc = _i(), _i()
        self._ck(self.lib.bc_occgrid_shape(self.h, float(w_m), float(h_m), float(cell_m), C.byref(hc), C.byref(wc)))
        return hc.value, wc.value

    def occgrid(self, d_labels, B, w_m, h_m, cell_m, binary, ros_layout, d_grids, stream=None):
        self._ck(self.lib.bc_occgrid(self.h, _ptr(d_labels), B, float(w_m), float(h_m), float(cell_m),
                                     int(binary), int(ros_layout), _ptr(d_grids), _ptr(stream)))

    def pipeline(self, d_bgr, h, w, B, lut, w_m, h_m, cell_m, binary, ros_layout, d_labels_out, d_grids,
                 stream=None):
        self._ck(self.lib.bc_pipeline(self.h, _ptr(d_bgr), h, w, B, _lut(lut), float(w_m), float(h_m),
                                      float(cell_m), int(binary), int(ros_layout), _ptr(d_labels_out),
                                      _ptr(d_grids), _ptr(stream)))

    def pipeline_host(self, h_bgr, h, w, B, lut, w_m, h_m, cell_m, binary, ros_layout, h_grids, stream=None):
        self._ck(self.lib.bc_pipeline_host(self.h, _ptr(h_bgr), h, w, B, _lut(lut), float(w_m), float(h_m),
                                           float(cell_m), int(binary), int(ros_layout), _ptr(h_grids),
                                           _ptr(stream)))

    def pipeline_host_submit(self, h_bgr, h, w, B, lut, w_m, h_m, cell_m, binary, ros_layout, h_grids, stream=None):
        self._ck(self.lib.bc_pipeline_host_submit(self.h, _ptr(h_bgr), h, w, B, _lut(lut), float(w_m), float(h_m),
                                                  float(cell_m), int(binary), int(ros_layout), _ptr(h_grids),
                                                  _ptr(stream)))

    def pipeline_host_wait(self, keep_in_flight=0):
        self._ck(self.lib.bc_pipeline_host_wait(self.h, int(keep_in_flight)))

    def gather_setup(self, d_base, rank, world):
        self._ck(self.lib.bc_gather_setup(self.h, _ptr(d_base), int(rank), int(world)))

    def gather_stream_setup(self, d_gather, d_arrive, d_release_mine, d_release_peers, rank, world):
        """d_gather: two device pointers (or None to switch the mode off); d_release_peers: list of `world`
        device pointers on rank 0, None elsewhere (see include/bugcar_b200.h)"""
        if d_gather is None:
            self._ck(self.lib.bc_gather_stream_setup(self.h, None, None, None, None, 0, 1))
            return
        ga = (_vp * 2)(*[_ptr(p) for p in d_gather])
        peers = None
        if d_release_peers is not None:
            peers = (_vp * len(d_release_peers))(*[_ptr(p) for p in d_release_peers])
        self._ck(self.lib.bc_gather_stream_setup(self.h, ga, _ptr(d_arrive), _ptr(d_release_mine), peers,
                                                 int(rank), int(world)))

    def launch_count(self):
        return int(self.lib.bc_launch_count(self.h))

    def set_profile(self, on):
        self._ck(self.lib.bc_set_profile(self.h, int(bool(on))))

    def profile(self):
        """list of {"kernel", "launches", "ms", "bytes", "flops"} since set_profile(True)"""
        import json
        return json.loads(self.lib.bc_profile_json(self.h).decode())


class HostBuffer:
    """Page-locked host memory from bc_host_alloc as a NumPy array (write_combined: frame staging that the CPU
    only writes and the GPU only reads)."""

    def __init__(self, shape, dtype, write_combined=False):
        import numpy as np
        self.lib = load()
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = _vp()
        rc = self.lib.bc_host_alloc(C.byref(p), n, int(bool(write_combined)))
        if rc != BC_OK:
            raise BugcarError(rc, "bc_host_alloc failed")
        self.ptr = p.value
        self.array = np.ctypeslib.as_array((C.c_uint8 * n).from_address(self.ptr)).view(dtype).reshape(shape)

    def data_ptr(self):
        return self.ptr

    def close(self):
        if self.ptr:
            self.array = None
            self.lib.bc_host_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def laser_tables(Wc, Hc, binary):
    """(fwd (pol_h, pol_w), inv (Hc, Wc)) int32 gather tables of the laserscan branch; host only."""
    import numpy as np
    lib = load()
    pw, ph = _i(), _i()
    rc = lib.bc_laser_tables(int(Wc), int(Hc), int(binary), C.byref(pw), C.byref(ph), None, None)
    if rc != BC_OK:
        raise BugcarError(rc, "bad grid shape")
    fwd = np.empty((ph.value, pw.value), np.int32)
    inv = np.empty((int(Hc), int(Wc)), np.int32)
    lib.bc_laser_tables(int(Wc), int(Hc), int(binary), C.byref(pw), C.byref(ph), fwd.ctypes.data, inv.ctypes.data)
    return fwd, inv


def _lut(lut):
    """256-byte class LUT as a ctypes buffer address (kept alive by the caller's frame)."""
    import numpy as np
    a = np.ascontiguousarray(lut, dtype=np.uint8)
    if a.size != 256:
        raise ValueError("class LUT must have 256 entries")
    _lut.keep = a                     # the C side copies it before returning
    return a.ctypes.data
