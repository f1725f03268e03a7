"""Stand-in for message_filters (oracle/shims/README.md): subscriptions are recorded, nothing is delivered; a test calls
the registered callback itself."""


class Subscriber:
    def __init__(self, topic, msg_type):
        self.topic, self.msg_type = topic, msg_type


class ApproximateTimeSynchronizer:
    def __init__(self, subs, queue_size, slop=0.1):
        self.subs, self.callbacks = subs, []

    def registerCallback(self, cb):
        self.callbacks.append(cb)
