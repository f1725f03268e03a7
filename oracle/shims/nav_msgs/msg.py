from _msgbag import Bag, Header


class Odometry(Bag):
    def __init__(self):
        self.header = Header()


class Path(Bag):
    def __init__(self):
        self.header = Header()
        self.poses = []
