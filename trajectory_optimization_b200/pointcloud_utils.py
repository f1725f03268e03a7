"""Device-side PointCloud2 codec with the names of the reference's `src/pointcloud_utils.py`.

The reference parses every message on the host (`np.fromstring` + structured-dtype field stripping + NaN filter,
src/pointcloud_utils.py:58-80,180-198) and only then moves the points to the GPU.  Here the message payload is copied
to the device as raw bytes and one pass of libcovb200.so extracts x, y, z and drops non-finite points in order
(`cov_pc2_to_xyz`); the inverse (`cov_xyz_to_pc2`) builds the payload of an outgoing message from device tensors.

`pointcloud2_to_xyz_tensor` returns the (M,3) fp32 CUDA tensor the models take; `pointcloud2_to_xyz_array` keeps the
reference's return type (numpy float64, M x 3) for callers that want it.  CUDA only: no CPU fallback.
"""
import ctypes

import numpy as np
import torch

from . import _lib

FLOAT32, FLOAT64 = 7, 8  # sensor_msgs/PointField datatypes


def _xyz_layout(cloud_msg):
    off, dt = {}, {}
    for f in cloud_msg.fields:
        if f.name in ("x", "y", "z"):
            off[f.name], dt[f.name] = int(f.offset), int(f.datatype)
    if set(off) != {"x", "y", "z"}:
        raise ValueError("PointCloud2 message has no x/y/z fields")
    if len(set(dt.values())) != 1 or dt["x"] not in (FLOAT32, FLOAT64):
        raise ValueError("x/y/z must share one datatype, FLOAT32 or FLOAT64")
    if getattr(cloud_msg, "is_bigendian", False):
        raise ValueError("big-endian PointCloud2 payloads are not supported (the reference assumes little endian too)")
    return off["x"], off["y"], off["z"], dt["x"]


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


@torch.no_grad()
def payload_to_xyz(data_dev, n_points, point_step, off_x, off_y, off_z, datatype=FLOAT32, remove_nans=True):
    """Raw payload bytes on the device (uint8 tensor) -> (M,3) fp32 tensor of the finite points, in order."""
    if not data_dev.is_cuda or data_dev.dtype != torch.uint8:
        raise RuntimeError("payload must be a CUDA uint8 tensor (there is no CPU fallback)")
    L = _lib.lib()
    n = int(n_points)
    xyz = torch.empty(max(n, 1), 3, dtype=torch.float32, device=data_dev.device)
    cnt = torch.zeros(1, dtype=torch.int64, device=data_dev.device)
    ws_bytes = L.cov_pc2_workspace_bytes(n)
    ws = torch.empty(max(ws_bytes, 4), dtype=torch.uint8, device=data_dev.device)
    _lib.check(L.cov_pc2_to_xyz(data_dev.data_ptr(), n, int(point_step), int(off_x), int(off_y), int(off_z), int(datatype),
                                1 if remove_nans else 0, xyz.data_ptr(), cnt.data_ptr(), ws.data_ptr(), ws_bytes, _stream()),
               "cov_pc2_to_xyz")
    return xyz[:int(cnt.item())]


def pointcloud2_to_xyz_tensor(cloud_msg, remove_nans=True, device=torch.device("cuda:0")):
    """src/pointcloud_utils.py:197-198 on the device: message -> (M,3) fp32 CUDA tensor."""
    ox, oy, oz, dt = _xyz_layout(cloud_msg)
    n = int(cloud_msg.width) * int(cloud_msg.height)
    raw = np.frombuffer(bytes(cloud_msg.data) if not isinstance(cloud_msg.data, (bytes, bytearray, memoryview)) else cloud_msg.data,
                        dtype=np.uint8, count=n * int(cloud_msg.point_step))
    data_dev = torch.from_numpy(raw.copy()).to(device, non_blocking=True)
    return payload_to_xyz(data_dev, n, cloud_msg.point_step, ox, oy, oz, dt, remove_nans)


def pointcloud2_to_xyz_array(cloud_msg, remove_nans=True, device=torch.device("cuda:0")):
    """Same return type as the reference (numpy float64, M x 3)."""
    return pointcloud2_to_xyz_tensor(cloud_msg, remove_nans, device).cpu().numpy().astype(np.float64)


@torch.no_grad()
def xyz_to_payload(points, extra=None):
    """(N,3) CUDA tensor [+ (N,) fourth field] -> (payload uint8 CUDA tensor of N*12 [N*16] bytes, is_dense bool)."""
    L = _lib.lib()
    if not points.is_cuda:
        raise RuntimeError("points must be a CUDA tensor (there is no CPU fallback)")
    pts = points.detach().float().contiguous()
    ex = None if extra is None else extra.detach().float().contiguous().reshape(-1)
    n = pts.shape[0]
    out = torch.empty(max(n, 1) * (12 if ex is None else 16), dtype=torch.uint8, device=pts.device)
    dense = torch.zeros(1, dtype=torch.int32, device=pts.device)
    _lib.check(L.cov_xyz_to_pc2(pts.data_ptr(), 0 if ex is None else ex.data_ptr(), n, out.data_ptr(), dense.data_ptr(),
                                _stream()), "cov_xyz_to_pc2")
    return out[:n * (12 if ex is None else 16)], bool(dense.item() != 0)


def _fill_msg(points, extra, stamp, frame_id):
    from sensor_msgs.msg import PointCloud2, PointField  # lazy: the module loads without ROS
    payload, dense = xyz_to_payload(points, extra)
    msg = PointCloud2()
    if stamp:
        msg.header.stamp = stamp
    if frame_id:
        msg.header.frame_id = frame_id
    msg.height = 1
    msg.width = int(points.shape[0])
    names = ["x", "y", "z"] + ([] if extra is None else ["i"])
    msg.fields = [PointField(nm, 4 * k, PointField.FLOAT32, 1) for k, nm in enumerate(names)]
    msg.is_bigendian = False
    msg.point_step = 4 * len(names)
    msg.row_step = int(points.shape[0])   # as the reference writes it (src/pointcloud_utils.py:309)
    msg.is_dense = int(dense)
    msg.data = payload.cpu().numpy().tobytes()
    return msg


def xyz_array_to_pointcloud2(points, stamp=None, frame_id=None):
    """src/pointcloud_utils.py:290-313 from a device tensor."""
    return _fill_msg(torch.as_tensor(points), None, stamp, frame_id)


def xyzi_array_to_pointcloud2(points, stamp=None, frame_id=None):
    """src/pointcloud_utils.py:315-338: (N,4) x, y, z, i."""
    p = torch.as_tensor(points)
    return _fill_msg(p[:, :3], p[:, 3], stamp, frame_id)
