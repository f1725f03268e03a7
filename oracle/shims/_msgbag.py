"""Attribute bags for the message stand-ins (oracle/shims/README.md)."""


class Bag:
    """`msg.pose.pose.position.x = 1.0` works without declaring the tree: unknown attributes become nested bags."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        v = Bag()
        object.__setattr__(self, name, v)
        return v


class Header(Bag):
    def __init__(self):
        self.stamp, self.frame_id, self.seq = None, "", 0
