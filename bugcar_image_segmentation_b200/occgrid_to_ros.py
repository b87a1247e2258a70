"""Drop-in for the reference's ``occgrid_to_ros.py``.  ROS publishing stays on the host
(north_star); what the GPU path contributes is the data layout: ``bc_occgrid`` with
``ros_layout=1`` emits the grid already flipped and rotated (occgrid_to_ros.py:18-21),
so the message fill is one flat copy.

``convert_to_occupancy_grid_msg`` keeps the reference signature (occgrid_to_ros.py:13)
and needs rospy / nav_msgs / geometry_msgs / scipy at call time, exactly like the
reference; importing this module does not.
"""
import numpy as np


def ros_cell_order(occ_grid):
    """occgrid_to_ros.py:18-24: cv2.flip(g, 0) then rotate 90 deg CCW, flattened
    (== g[::-1, ::-1].T).  Host NumPy; used when the grid was not produced with
    ``ros_layout=1``."""
    g = np.asarray(occ_grid)
    return np.ascontiguousarray(g[::-1, ::-1].T).reshape(-1)


def convert_to_occupancy_grid_msg(occ_grid, map_resolution, map_width, map_height, time_stamp, frame_id, pose,
                                  already_ros_layout=False):
    import rospy
    from std_msgs.msg import Header
    from nav_msgs.msg import OccupancyGrid, MapMetaData
    from geometry_msgs.msg import Pose, Point, Quaternion
    from scipy.spatial.transform import Rotation as R

    data = np.asarray(occ_grid).reshape(-1) if already_ros_layout else ros_cell_order(occ_grid)
    rot = R.from_euler("xyz", pose[3:])                                   # occgrid_to_ros.py:27-30
    quat = rot.as_quat()
    first_cell = rot.as_matrix() @ (np.array([0, -map_width / 2, 0]) + pose[:3])
    msg = OccupancyGrid()
    msg.header = Header()
    msg.header.frame_id = frame_id
    msg.header.stamp = time_stamp
    msg.info = MapMetaData()
    msg.info.height = int(map_width / map_resolution)                      # occgrid_to_ros.py:39
    msg.info.width = int(map_height / map_resolution)                      # occgrid_to_ros.py:41
    msg.info.resolution = map_resolution
    msg.info.origin = Pose()
    msg.info.origin.position = Point()
    msg.info.origin.position.x, msg.info.origin.position.y, msg.info.origin.position.z = first_cell
    msg.info.origin.orientation = Quaternion()
    (msg.info.origin.orientation.x, msg.info.origin.orientation.y, msg.info.origin.orientation.z,
     msg.info.origin.orientation.w) = quat
    msg.data.extend(data.tolist())
    msg.info.map_load_time = rospy.Time.now()
    return msg
