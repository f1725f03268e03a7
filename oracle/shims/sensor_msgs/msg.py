class _Header:
    def __init__(self):
        self.stamp, self.frame_id, self.seq = None, "", 0


class PointCloud2:
    def __init__(self):
        self.header = _Header()
        self.height = self.width = 0
        self.fields = []
        self.is_bigendian = False
        self.point_step = self.row_step = 0
        self.data = b""
        self.is_dense = False


class CameraInfo:
    def __init__(self):
        self.header = _Header()


class Image:
    pass


class CompressedImage:
    pass


class PointField:
    INT8, UINT8, INT16, UINT16, INT32, UINT32, FLOAT32, FLOAT64 = range(1, 9)

    def __init__(self, name="", offset=0, datatype=0, count=1):
        self.name, self.offset, self.datatype, self.count = name, offset, datatype, count
