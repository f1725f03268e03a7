class PointCloud2:
    pass


class CameraInfo:
    pass


class Image:
    pass


class CompressedImage:
    pass


class PointField:
    INT8, UINT8, INT16, UINT16, INT32, UINT32, FLOAT32, FLOAT64 = range(1, 9)
