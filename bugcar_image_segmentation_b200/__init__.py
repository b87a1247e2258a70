"""B200-native perception hot path of tranqkhue/bugcar_image_segmentation.

Camera frame -> ENet segmentation -> class argmax/LUT -> bird's-eye-view warp ->
occupancy grid, as hand-written sm_100a CUDA behind the reference's own Python call
surface.  Module names mirror the reference package (``models``, ``bev``,
``image_processing_utils``, ``occgrid_to_ros``, ``utils``) so that
``from bugcar_image_segmentation_b200.models import ENET`` replaces
``from <reference>.models import ENET``.  All compute goes through the C ABI in
``include/bugcar_b200.h`` (``libbugcar_b200.so``); there is no CPU fallback.
"""
__version__ = "0.1.0"
