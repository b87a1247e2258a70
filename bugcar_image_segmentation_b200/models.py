"""Drop-in for the reference's ``models.py`` (ENet side): same class, method names,
argument meaning and return types; the bodies call the CUDA library.

Reference behaviour mirrored (``models.py`` of tranqkhue/bugcar_image_segmentation):
  * class constants                        models.py:15-19
  * ``ENET(GRAPH_PB_PATH=None)``           models.py:21-31   (weights file instead of a GraphDef)
  * ``ENET.predict``  -> uint8 (B,256,512) in {0,1,2}      models.py:42-69
  * ``ENET.predict_binary`` -> uint8 (B,256,512) in {0,1}  models.py:70-82
  * ``ENET.preprocess`` -> float64 (1,3,256,512)           models.py:84-95
Deliberate deviations: no per-call ``print`` (models.py:53); the model file is a
BCENETW1 container (``weights.py``) because the reference's frozen graph
``pretrained_models/enet.pb`` / ``model.h5`` is not part of its source tree.
``DeepLabV3`` (models.py:98-136) is not provided: its graph is a missing blob and the
reference wrapper itself cannot run.
"""
import os
import warnings
from abc import ABC

import numpy as np

from . import _lib, runtime


class InferenceModel(ABC):                       # models.py:8-13
    def predict(self, preprocessed_image):
        pass

    @classmethod
    def preprocess(rgb_image):
        pass


def _lut_3way():
    lut = np.full(256, 2, np.uint8)              # models.py:56
    lut[[2, 9]] = 0                              # models.py:57  pavement, vegetation
    lut[[0, 1]] = 1                              # models.py:58  road, lane marking
    return lut


def _lut_binary():
    lut = np.zeros(256, np.uint8)
    lut[[0, 1]] = 1                              # models.py:79-80
    return lut


class ENET(InferenceModel):
    INPUT_TENSOR_NAME = "input0:0"               # kept for source compatibility (models.py:15-16)
    OUTPUT_TENSOR_NAME = "CATkrIDy/concat:0"
    IMAGE_MEAN = np.array([0.485, 0.456, 0.406])
    IMAGE_STD = np.array([0.229, 0.224, 0.225])
    INPUT_WIDTH, INPUT_HEIGHT = (512, 256)

    LUT_3WAY = _lut_3way()
    LUT_BINARY = _lut_binary()
    DEFAULT_WEIGHTS = "./pretrained_models/enet.bcw"

    _shared_pre = {}                             # device -> Context used by the classmethod preprocess

    def __init__(self, GRAPH_PB_PATH=None, device=None, max_batch=None, precision="fp16"):
        torch, dev = runtime.torch_cuda(device)
        self._torch, self.device = torch, dev
        if GRAPH_PB_PATH is None:
            GRAPH_PB_PATH = self.DEFAULT_WEIGHTS
            if not os.path.isfile(GRAPH_PB_PATH):
                warnings.warn("no ./pretrained_models/enet.bcw; loading the seeded synthetic stand-in "
                              "(the reference's trained blobs are not part of its source tree)")
                GRAPH_PB_PATH = runtime.SYNTHETIC_WEIGHTS
        if str(GRAPH_PB_PATH).endswith((".pb", ".h5")):
            raise ValueError("TensorFlow/Keras blobs are not read directly: convert with tools/convert_weights.py "
                             "to a .bcw container first")
        with open(GRAPH_PB_PATH, "rb") as f:     # models.py:25-26
            blob = f.read()
        if precision not in _lib.PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(_lib.PRECISIONS)}")
        self.ctx = runtime.new_context(dev, max_batch)
        self.ctx.set_precision(_lib.PRECISIONS[precision])      # before the load: weights are packed once
        self.ctx.load_enet(blob)
        self.num_classes = self.ctx.num_classes()
        self.test = None                         # models.py:31

    # ------------------------------------------------------------------ inference
    def _labels(self, preprocessed_imgs, lut):
        torch, dev = self._torch, self.device
        x = preprocessed_imgs
        if isinstance(x, (list, tuple)):
            x = np.asarray(x)
        if isinstance(x, np.ndarray):
            if x.dtype == np.uint8:
                kind = _lib.BC_IN_BGR_U8
            elif x.dtype == np.float64:
                kind = _lib.BC_IN_NCHW_F64
            else:
                x = x.astype(np.float32, copy=False)
                kind = _lib.BC_IN_NCHW_F32
            x = torch.from_numpy(np.ascontiguousarray(x)).to(f"cuda:{dev}")
        else:
            x = x.to(f"cuda:{dev}").contiguous()
            kind = {torch.uint8: _lib.BC_IN_BGR_U8, torch.float64: _lib.BC_IN_NCHW_F64}.get(x.dtype)
            if kind is None:
                x = x.float()
                kind = _lib.BC_IN_NCHW_F32
        if kind == _lib.BC_IN_BGR_U8:
            ok = x.dim() == 4 and tuple(x.shape[1:]) == (self.INPUT_HEIGHT, self.INPUT_WIDTH, 3)
        else:
            ok = x.dim() == 4 and tuple(x.shape[1:]) == (3, self.INPUT_HEIGHT, self.INPUT_WIDTH)
        if not ok:
            raise ValueError(f"expected (B,3,{self.INPUT_HEIGHT},{self.INPUT_WIDTH}) floats "
                             f"(or (B,{self.INPUT_HEIGHT},{self.INPUT_WIDTH},3) uint8 BGR), got {tuple(x.shape)}")
        B = x.shape[0]
        out = torch.empty((B, self.INPUT_HEIGHT, self.INPUT_WIDTH), dtype=torch.uint8, device=x.device)
        s = runtime.stream_handle(torch, dev)
        step = self.ctx.max_batch
        for b0 in range(0, B, step):
            n = min(step, B - b0)
            self.ctx.enet_labels(x[b0:b0 + n], kind, n, lut, out[b0:b0 + n], s)
        return out

    def predict(self, preprocessed_imgs):
        """(B,3,256,512) float -> uint8 (B,256,512): 1 road, 0 flat non-road, 2 the rest."""
        return self._labels(preprocessed_imgs, self.LUT_3WAY).cpu().numpy()

    def predict_binary(self, preprocessed_imgs):
        """(B,3,256,512) float -> uint8 (B,256,512): 1 road / lane marking, 0 the rest."""
        return self._labels(preprocessed_imgs, self.LUT_BINARY).cpu().numpy()

    def predict_device(self, x, lut=None):
        """Extension: same as predict but the result stays on the GPU (torch uint8)."""
        return self._labels(x, self.LUT_3WAY if lut is None else lut)

    def logits(self, preprocessed_imgs):
        """Extension (parity/debug): what ``sess.run`` returns, fp32 NCHW (B,C,256,512)."""
        torch, dev = self._torch, self.device
        x = preprocessed_imgs
        if isinstance(x, np.ndarray):
            if x.dtype == np.uint8:
                kind = _lib.BC_IN_BGR_U8
            elif x.dtype == np.float64:
                kind = _lib.BC_IN_NCHW_F64
            else:
                x = x.astype(np.float32, copy=False)
                kind = _lib.BC_IN_NCHW_F32
            x = torch.from_numpy(np.ascontiguousarray(x)).to(f"cuda:{dev}")
        else:
            kind = {torch.uint8: _lib.BC_IN_BGR_U8, torch.float64: _lib.BC_IN_NCHW_F64}.get(x.dtype, _lib.BC_IN_NCHW_F32)
            x = x.to(f"cuda:{dev}").contiguous()
        B = x.shape[0]
        out = torch.empty((B, self.num_classes, self.INPUT_HEIGHT, self.INPUT_WIDTH), dtype=torch.float32,
                          device=x.device)
        step = self.ctx.max_batch
        for b0 in range(0, B, step):
            n = min(step, B - b0)
            self.ctx.enet_logits(x[b0:b0 + n], kind, n, out[b0:b0 + n], runtime.stream_handle(torch, dev))
        return out.cpu().numpy()

    # ---------------------------------------------------------------- preprocessing
    @classmethod
    def preprocess(cls, bgr_frame, device=None):
        """uint8 BGR (h,w,3) -> float64 (1,3,256,512); resize, BGR->RGB,
        (rgb/256 - mean)/std, HWC->CHW on the GPU, bit-exact with models.py:84-95."""
        torch, dev = runtime.torch_cuda(device)
        ctx = cls._shared_pre.get(dev)
        if ctx is None:
            ctx = cls._shared_pre[dev] = runtime.new_context(dev, 8)
        frame = np.ascontiguousarray(bgr_frame, dtype=np.uint8)
        if frame.ndim != 3 or frame.shape[2] != 3:
            raise ValueError("expected a (h, w, 3) BGR frame")
        h, w = frame.shape[:2]
        d_in = torch.from_numpy(frame).to(f"cuda:{dev}")
        d_out = torch.empty((1, 3, cls.INPUT_HEIGHT, cls.INPUT_WIDTH), dtype=torch.float64, device=d_in.device)
        ctx.preprocess(d_in, h, w, 1, d_out, 1, runtime.stream_handle(torch, dev))
        return d_out.cpu().numpy()
