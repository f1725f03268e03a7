class Odometry:
    pass


class Path:
    pass
