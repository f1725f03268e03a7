"""Empty stand-in (oracle/shims/README.md)."""


class Time:
    @staticmethod
    def now():
        return 0.0


def loginfo(*a, **k):
    pass
