"""Drop-in for the reference's ``bev.py``: ``bev_transform_tools`` with the same
constructor, JSON schema and method names; the label-map -> occupancy-grid chain
(warpPerspective + crop/paste + 3x3 opening + nearest resize + int8 map) runs as one
CUDA kernel through ``bc_occgrid``.

Reference behaviour mirrored (``bev.py`` of tranqkhue/bugcar_image_segmentation):
  * ``__init__`` / ``fromJSON`` / ``save_to_JSON``          bev.py:13-56
  * ``calculate_transform_matrix``                         bev.py:58-92
  * ``create_occupancy_grid``        -> int8 (Hc, Wc)      bev.py:166-246
  * ``create_occupancy_grid_binary`` -> int8 (Hc, Wc)      bev.py:97-165
  * shape precondition -> AssertionError                   bev.py:169-170
  * missing JSON key   -> KeyError                         bev.py:29-37
Deliberate deviations: no ``cv2.imshow`` inside the path (bev.py:132,213) and no
``print`` (bev.py:38,80,86); ``save_to_JSON`` also writes ``is_laserscan`` so that its
output can be re-loaded (the reference's cannot, bev.py:47-55 vs bev.py:37);
laserscan mode (bev.py:145-164, 216-240) reads uninitialised ``warpPolar`` memory in
the reference (its output is not a function of its input); here those outlier pixels are
0, i.e. the result equals the reference's with ``WARP_FILL_OUTLIERS`` set in its two
``cv2.warpPolar`` calls (``bc_occgrid_laserscan``).
"""
import json

import numpy as np

from . import runtime
from .utils import order_points_counter_clockwise


class bev_transform_tools:

    # dist2target: distance from camera to the target, (x, y) in cm
    def __init__(self, input_image_shape, desired_image_shape, dist2target, tile_length, cm_per_px, yaw,
                 make_laserscan_like=False):
        self.input_width = input_image_shape[0]          # sic: compared with segmap.shape[0] (rows), bev.py:169
        self.input_height = input_image_shape[1]
        self.after_warp_width = desired_image_shape[0]
        self.after_warp_height = desired_image_shape[1]
        self.dist2target = dist2target
        self.tile_length = tile_length                   # in cm
        self.cm_per_px = cm_per_px
        self.yaw = yaw
        self.laserscan_like_occupancy_grid = make_laserscan_like
        self._ctx = None
        self._ctx_key = None

    @classmethod
    def fromJSON(cls, filepath):
        with open(filepath, mode="r") as f:
            data = json.load(f)
        shape = data["output image size"]
        input_shape = data["input image size"]
        bev_matrix = np.reshape(np.array(data["bev matrix"]), (3, 3))
        dist2target = data["distance to target"]
        tile_length = data["tile_length"]
        cm_per_px = data["cm_per_px"]
        yaw = data["yaw"]
        is_laserscan = data["is_laserscan"]
        bev = cls(input_shape, shape, dist2target, tile_length, cm_per_px, yaw, is_laserscan)
        bev._bev_matrix = bev_matrix
        return bev

    def save_to_JSON(self, file_path):
        data = {
            "input image size": (self.input_width, self.input_height),
            "output image size": (self.after_warp_width, self.after_warp_height),
            "bev matrix": np.asarray(self._bev_matrix).tolist(),
            "distance to target": self.dist2target,
            "tile_length": self.tile_length,
            "cm_per_px": self.cm_per_px,
            "yaw": self.yaw,
            "is_laserscan": bool(self.laserscan_like_occupancy_grid),
        }
        with open(file_path, mode="w") as f:
            json.dump(data, f)

    # ---------------------------------------------------------------- calibration (host, offline)
    def calculate_transform_matrix(self, tile_coords):
        """bev.py:58-92: homography taking the four fiducial-tile corners (image px) onto
        the yaw-rotated square placed ``dist2target`` ahead of the camera."""
        cm_per_px = self.cm_per_px
        yaw = self.yaw
        dist_px = (self.dist2target[0] / cm_per_px, self.dist2target[1] / cm_per_px)
        half = self.tile_length / cm_per_px / 2
        square = np.array([[half, half], [half, -half], [-half, -half], [-half, half]])
        rot = np.array([[np.cos(yaw), -np.sin(yaw)], [np.sin(yaw), np.cos(yaw)]])
        target = np.array([self.after_warp_width / 2 + dist_px[0], self.after_warp_height - dist_px[1]])
        axis = np.stack([target, rot @ np.array([100.0, 0.0]) + target], axis=0)
        corners = (rot @ square.T).T + target
        corners = order_points_counter_clockwise(corners, axis)
        M = _perspective_transform(np.asarray(tile_coords, np.float32), corners.astype(np.float32))
        self._bev_matrix = M
        return M

    # ---------------------------------------------------------------- hot path
    def _context(self, ctx=None):
        """C-ABI context carrying this calibration (own one unless a model's is given)."""
        M = np.asarray(self._bev_matrix, np.float64).reshape(9)
        key = (M.tobytes(), int(self.input_width), int(self.input_height), int(self.after_warp_width),
               int(self.after_warp_height), float(self.cm_per_px))
        if ctx is not None:
            ctx.set_bev(M, key[1], key[2], key[3], key[4], key[5])
            return ctx
        if self._ctx is None:
            _, dev = runtime.torch_cuda(None)
            self._ctx = runtime.new_context(dev, 1)
            self._ctx_key = None
        if self._ctx_key != key:
            self._ctx.set_bev(M, key[1], key[2], key[3], key[4], key[5])
            self._ctx_key = key
        return self._ctx

    def _grid(self, segmap, w_m, h_m, cell_m, binary, ros_layout=False):
        shape = tuple(segmap.shape)
        assert shape == (self.input_width, self.input_height), \
            "current segmap size: {},the segmap's original size must be the same as the required input shape, which is {}" \
            .format(shape, (self.input_width, self.input_height))
        torch, dev = runtime.torch_cuda(None if self._ctx is None else self._ctx.device)
        ctx = self._context()
        d_lab = runtime.to_device_u8(torch, ctx.device, segmap)
        hc, wc = ctx.occgrid_shape(w_m, h_m, cell_m)
        if self.laserscan_like_occupancy_grid:
            # bev.py:145-164 / 216-240 with the warpPolar outliers defined as 0 (the reference leaves them
            # uninitialised); the binary variant returns the reference's 2-tuple (plain grid, laserscan grid)
            if ros_layout:
                raise NotImplementedError("ros_layout is not available for laserscan-like grids")
            d_laser = torch.empty((hc, wc), dtype=torch.int8, device=d_lab.device)
            d_plain = torch.empty((hc, wc), dtype=torch.int8, device=d_lab.device) if binary else None
            ctx.occgrid_laserscan(d_lab, 1, w_m, h_m, cell_m, binary, d_plain, d_laser,
                                  runtime.stream_handle(torch, ctx.device))
            return (d_plain.cpu().numpy(), d_laser.cpu().numpy()) if binary else d_laser.cpu().numpy()
        out_shape = (wc, hc) if ros_layout else (hc, wc)
        d_grid = torch.empty(out_shape, dtype=torch.int8, device=d_lab.device)
        ctx.occgrid(d_lab, 1, w_m, h_m, cell_m, binary, ros_layout, d_grid, runtime.stream_handle(torch, ctx.device))
        return d_grid.cpu().numpy()

    def create_occupancy_grid(self, segmap, occupancy_grid_width_in_m, occupancy_grid_height_in_m, cell_size_in_m):
        """labels {0,1,2} from ENET.predict -> int8 grid: -1 unknown, 0 free, 100 occupied."""
        return self._grid(segmap, occupancy_grid_width_in_m, occupancy_grid_height_in_m, cell_size_in_m, 0)

    def create_occupancy_grid_binary(self, segmap, occupancy_grid_width_in_m, occupancy_grid_height_in_m,
                                     cell_size_in_m):
        """road mask {0,1} from ENET.predict_binary -> int8 grid."""
        return self._grid(segmap, occupancy_grid_width_in_m, occupancy_grid_height_in_m, cell_size_in_m, 1)

    def create_occupancy_grid_ros(self, segmap, occupancy_grid_width_in_m, occupancy_grid_height_in_m,
                                  cell_size_in_m, binary=False):
        """Extension: the grid already flipped + rotated 90 deg CCW as
        occgrid_to_ros.py:18-21 does, shape (Wc, Hc), ready for OccupancyGrid.data."""
        return self._grid(segmap, occupancy_grid_width_in_m, occupancy_grid_height_in_m, cell_size_in_m,
                          int(binary), True)


def _perspective_transform(src, dst):
    """3x3 homography from 4 point pairs: the linear system of cv2.getPerspectiveTransform
    (bev.py:88), solved in fp64 from the float32 points."""
    src = np.asarray(src, np.float32).astype(np.float64)
    dst = np.asarray(dst, np.float32).astype(np.float64)
    A = np.zeros((8, 8))
    b = np.zeros(8)
    for i in range(4):
        x, y = src[i]
        u, v = dst[i]
        A[i] = [x, y, 1, 0, 0, 0, -x * u, -y * u]
        A[i + 4] = [0, 0, 0, x, y, 1, -x * v, -y * v]
        b[i], b[i + 4] = u, v
    return np.append(np.linalg.solve(A, b), 1.0).reshape(3, 3)
