class TransformStamped:
    pass


class PoseStamped:
    pass
