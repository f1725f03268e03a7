from _msgbag import Bag, Header


class TransformStamped(Bag):
    def __init__(self):
        self.header = Header()
        self.child_frame_id = ""


class PoseStamped(Bag):
    def __init__(self):
        self.header = Header()
