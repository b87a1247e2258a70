"""CPU restatement of ENET.preprocess and the argmax + class LUT tail.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
  * preprocess           models.py:84-95  (pinned: reference code run under stubs)
  * labels_3way          models.py:55-58,67
  * labels_binary        models.py:78-81
The reference calls tf.math.argmax (TensorFlow == 2.2, requirements.txt:2, absent
here); its documented tie-break (lowest index) equals np.argmax.
"""
import numpy as np

from . import cv_ops

IMAGE_MEAN = np.array([0.485, 0.456, 0.406])   # models.py:17
IMAGE_STD = np.array([0.229, 0.224, 0.225])    # models.py:18
INPUT_WIDTH, INPUT_HEIGHT = 512, 256           # models.py:19


def preprocess(bgr, backend="numpy"):
    """uint8 (h,w,3) BGR -> float64 (1,3,256,512); models.py:85-95."""
    if backend == "cv2":
        import cv2
        resized = cv2.resize(bgr, (INPUT_WIDTH, INPUT_HEIGHT))
    else:
        resized = cv_ops.resize_bilinear_u8(bgr, (INPUT_WIDTH, INPUT_HEIGHT))
    rgb = resized[:, :, ::-1]                                  # models.py:89
    normalized = (rgb / 256.0 - IMAGE_MEAN) / IMAGE_STD        # models.py:91
    return np.expand_dims(np.moveaxis(normalized, -1, 0), 0)   # models.py:92-94


def normalise_lut():
    """(256,3) fp64 table, RGB order: ((u/256.0) - mean) / std."""
    u = np.arange(256, dtype=np.float64)[:, None]
    return (u / 256.0 - IMAGE_MEAN[None, :]) / IMAGE_STD[None, :]


LUT_3WAY = np.full(256, 2, np.uint8)
LUT_3WAY[[2, 9]] = 0       # models.py:57  flat non-road: pavement, vegetation
LUT_3WAY[[0, 1]] = 1       # models.py:58  road, lane marking
LUT_BINARY = np.zeros(256, np.uint8)
LUT_BINARY[[0, 1]] = 1     # models.py:79-80


def labels_from_logits(logits, lut):
    """logits (B,C,H,W) float -> uint8 (B,H,W): argmax over axis 1 (first max wins)
    then class LUT; models.py:55-58,67 (3-way) / 78-81 (binary)."""
    cls = np.argmax(logits, axis=1)
    return lut[cls].astype(np.uint8)
