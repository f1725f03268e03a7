"""Stand-in for rospy (oracle/shims/README.md): the names the reference's modules touch at import time, plus what the
node classes use when a test drives their callbacks — parameters, a rate, the clock and publishers that RECORD what is
published (`PUBLISHED`: list of (topic, message))."""
PARAMS = {}       # tests set launch-file parameters here (rospy.get_param reads it)
PUBLISHED = []    # (topic, message) in publication order


class Time:
    @staticmethod
    def now():
        return 0.0


class Rate:
    def __init__(self, hz):
        self.hz = hz

    def sleep(self):
        pass


class Publisher:
    def __init__(self, topic, msg_type, queue_size=1):
        self.topic, self.msg_type = topic, msg_type

    def publish(self, msg):
        PUBLISHED.append((self.topic, msg))


def get_param(name, default=None):
    return PARAMS.get(name, default)


def init_node(*a, **k):
    pass


def spin():
    pass


def is_shutdown():
    return True


def loginfo(*a, **k):
    pass
