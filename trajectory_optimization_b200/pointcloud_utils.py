"""Device-side PointCloud2 codec with the names of the reference's `src/pointcloud_utils.py`.

The reference parses every message on the host (`np.fromstring` + structured-dtype field stripping + NaN filter,
src/pointcloud_utils.py:58-80,180-198) and only then moves the points to the GPU.  Here the message payload goes to
the device as raw bytes through a pinned staging buffer (one host memcpy out of the message's `bytes`, one async DMA)
and one pass of libcovb200.so extracts x, y, z and drops non-finite points in order (`cov_pc2_to_xyz`); the inverse
(`cov_xyz_to_pc2`) builds the payload of an outgoing message.

`pointcloud2_to_xyz_tensor` returns the (M,3) fp32 CUDA tensor the models take; `pointcloud2_to_xyz_array` keeps the
reference's return type (numpy float64, M x 3).  `xyz_array_to_pointcloud2` / `xyzi_array_to_pointcloud2` take what the
reference's callers pass — numpy arrays (src/tools.py:224-231 from src/pose_optimization.py:108-112,
src/trajectory_optimization.py:147-157, src/pc_processor.py:130,173,183), CPU tensors or CUDA tensors; host inputs are
moved to the device, the arithmetic (fp32 conversion, finiteness test, record packing) is the kernel's.  No CPU
fallback: without a CUDA device and the library these functions raise.
"""
import threading

import numpy as np
import torch

from . import ops

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


class CudaCodec:
    """The two codec kernels + the staging they need.  `pointcloud_utils._CODEC` is the only instance the product uses;
    the GPU-less node test swaps in a numpy stand-in with the same methods (tests/_standins.py)."""

    def __init__(self):
        self._pinned = None
        self._lock = threading.Lock()

    def default_device(self):
        if not torch.cuda.is_available():
            raise RuntimeError("the PointCloud2 codec is CUDA-only (there is no CPU fallback)")
        return torch.device("cuda", torch.cuda.current_device())

    def stage_bytes(self, data, nbytes, device):
        """message bytes -> uint8 device tensor via a grow-only pinned buffer (async H2D on the current stream)."""
        mv = memoryview(data) if isinstance(data, (bytes, bytearray, memoryview)) else memoryview(bytes(data))
        if mv.nbytes < nbytes:
            raise ValueError(f"PointCloud2 payload has {mv.nbytes} bytes, header promises {nbytes}")
        dev = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=device)
        if nbytes == 0:
            return dev[:0]
        with self._lock:
            if self._pinned is None or self._pinned.numel() < nbytes:
                self._pinned = torch.empty(max(nbytes, 2 * (0 if self._pinned is None else self._pinned.numel())),
                                           dtype=torch.uint8, pin_memory=True)
            host = self._pinned[:nbytes]
            # the previous message's DMA out of this buffer must be over before it is overwritten
            torch.cuda.current_stream(device).synchronize()
            host.numpy()[:] = np.frombuffer(mv, dtype=np.uint8, count=nbytes)
            dev.copy_(host, non_blocking=True)
        return dev[:nbytes]

    def to_device(self, points, device):
        """numpy / CPU tensor / CUDA tensor -> tensor on `device` in its own dtype."""
        t = points if isinstance(points, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(points))
        return t.detach().to(device)

    def pc2_to_xyz(self, data_dev, n, point_step, off_x, off_y, off_z, datatype, remove_nans):
        if not data_dev.is_cuda or data_dev.dtype != torch.uint8:
            raise RuntimeError("payload must be a CUDA uint8 tensor (there is no CPU fallback)")
        L = ops._lib.lib()
        xyz = torch.empty(max(n, 1), 3, dtype=torch.float32, device=data_dev.device)
        cnt = torch.zeros(1, dtype=torch.int64, device=data_dev.device)
        ws_bytes = L.cov_pc2_workspace_bytes(n)
        ws = torch.empty(max(ws_bytes, 4), dtype=torch.uint8, device=data_dev.device)
        ops._call("cov_pc2_to_xyz", data_dev, data_dev.data_ptr(), n, int(point_step), int(off_x), int(off_y), int(off_z),
                  int(datatype), 1 if remove_nans else 0, xyz.data_ptr(), cnt.data_ptr(), ws.data_ptr(), ws_bytes)
        return xyz[:int(cnt.item())]

    def xyz_to_pc2(self, pts, extra):
        if not pts.is_cuda:
            raise RuntimeError("points must be a CUDA tensor (there is no CPU fallback)")
        n = pts.shape[0]
        rec = 12 if extra is None else 16
        out = torch.empty(max(n, 1) * rec, dtype=torch.uint8, device=pts.device)
        dense = torch.zeros(1, dtype=torch.int32, device=pts.device)
        ops._call("cov_xyz_to_pc2", pts, pts.data_ptr(), 0 if extra is None else extra.data_ptr(), n, out.data_ptr(),
                  dense.data_ptr())
        return out[:n * rec], bool(dense.item() != 0)


_CODEC = CudaCodec()


@torch.no_grad()
def payload_to_xyz(data_dev, n_points, point_step, off_x, off_y, off_z, datatype=FLOAT32, remove_nans=True):
    """Raw payload bytes on the device (uint8 tensor) -> (M,3) fp32 tensor of the finite points, in order."""
    return _CODEC.pc2_to_xyz(data_dev, int(n_points), point_step, off_x, off_y, off_z, datatype, remove_nans)


def pointcloud2_to_xyz_tensor(cloud_msg, remove_nans=True, device=None):
    """src/pointcloud_utils.py:197-198 on the device: message -> (M,3) fp32 CUDA tensor."""
    ox, oy, oz, dt = _xyz_layout(cloud_msg)
    n = int(cloud_msg.width) * int(cloud_msg.height)
    device = _CODEC.default_device() if device is None else torch.device(device)
    data_dev = _CODEC.stage_bytes(cloud_msg.data, n * int(cloud_msg.point_step), device)
    return payload_to_xyz(data_dev, n, cloud_msg.point_step, ox, oy, oz, dt, remove_nans)


def pointcloud2_to_xyz_array(cloud_msg, remove_nans=True, device=None):
    """Same return type as the reference (numpy float64, M x 3)."""
    return pointcloud2_to_xyz_tensor(cloud_msg, remove_nans, device).cpu().numpy().astype(np.float64)


@torch.no_grad()
def xyz_to_payload(points, extra=None, device=None):
    """(N,3) points [+ (N,) fourth field] -> (payload uint8 device tensor of N*12 [N*16] bytes, is_dense bool).
    `points` may be a numpy array or a tensor on any device; `is_dense` follows the reference
    (`np.isfinite(points).all()` on the values as given, src/pointcloud_utils.py:310,335)."""
    if device is None:
        device = points.device if isinstance(points, torch.Tensor) and points.is_cuda else _CODEC.default_device()
    p = _CODEC.to_device(points, device)
    e = None if extra is None else _CODEC.to_device(extra, device).reshape(-1)
    wide = p.dtype == torch.float64 or (e is not None and e.dtype == torch.float64)
    # finite fp64 values beyond the fp32 range become inf in the payload but the reference's flag looks at the input
    wide_dense = None
    if wide:
        wide_dense = bool(torch.isfinite(p).all().item()) and (e is None or bool(torch.isfinite(e).all().item()))
    p = p.float().contiguous()
    e = None if e is None else e.float().contiguous()
    payload, dense = _CODEC.xyz_to_pc2(p, e)
    return payload, (dense if wide_dense is None else wide_dense)


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
    """src/pointcloud_utils.py:290-313; `points` (N,3): numpy array (what the reference passes) or tensor."""
    return _fill_msg(points, None, stamp, frame_id)


def xyzi_array_to_pointcloud2(points, stamp=None, frame_id=None):
    """src/pointcloud_utils.py:315-338: (N,4) x, y, z, i; numpy array (what the reference passes) or tensor."""
    return _fill_msg(points[:, :3], points[:, 3], stamp, frame_id)
