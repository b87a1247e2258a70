"""Drop-in module name for the reference's ``image_processing_utils.py``.

None of these helpers is called on the hot path in the reference snapshot
(``contour_noise_removal`` is imported by models.py:6 next to ``predict_binary`` and never
called; ``create_skeleton`` calls ``create_occupancy_grid`` with a stale signature and cannot
run).  ``contour_noise_removal`` runs on the GPU through
``bc_contour_noise_removal``; there is no CPU fallback in this package.
"""
import numpy as np

from . import runtime

_ctx = {}


def _context(device):
    if device not in _ctx:
        _ctx[device] = runtime.new_context(device, 1)
    return _ctx[device]


def find_intersection_line(line1, line2):
    """image_processing_utils.py:63-91: intersection of two lines, each given by two
    points; ``None`` for parallel lines."""
    def coeffs(line):
        (x1, y1), (x2, y2) = line
        if x2 - x1 == 0:
            return 1.0, 0.0, x1
        a = (y2 - y1) / (x2 - x1)
        return a, -1.0, (x1 * y2 - x2 * y1) / (x2 - x1)

    a1, b1, c1 = coeffs(line1)
    a2, b2, c2 = coeffs(line2)
    if a1 == a2:
        return None
    return np.linalg.solve(np.array([[a1, b1], [a2, b2]]), np.array([c1, c2]))


def contour_noise_removal(segmap, device=None):
    """image_processing_utils.py:4-44: close the road mask with a k x k box
    (k = int(min(h, w) / 50)), keep the contours that cover more than 40 % of the bottom tenth
    of the image and fill them.  ``segmap``: uint8 (h, w) -- or (B, h, w) for a batch, or a CUDA
    uint8 tensor, in which case a CUDA tensor comes back.  Returns uint8 {0, 1} of the same shape."""
    torch, device = runtime.torch_cuda(device)
    is_np = isinstance(segmap, np.ndarray)
    if is_np and segmap.dtype != np.uint8:
        raise TypeError("contour_noise_removal expects a uint8 mask (cv2.morphologyEx / findContours do)")
    if segmap.ndim not in (2, 3):
        raise ValueError("expected a (h, w) mask or a (B, h, w) batch of masks")
    d_in = runtime.to_device_u8(torch, device, segmap)
    h, w = int(d_in.shape[-2]), int(d_in.shape[-1])
    B = int(d_in.shape[0]) if d_in.ndim == 3 else 1
    d_out = torch.empty_like(d_in)
    _context(device).contour_noise_removal(d_in, h, w, B, d_out, runtime.stream_handle(torch, device))
    return d_out.cpu().numpy() if is_np else d_out


def clahe(img):
    raise NotImplementedError("clahe (image_processing_utils.py:46-61) is an optional camera pre-filter outside "
                              "the accelerated path")
