"""Empty stand-in (oracle/shims/README.md)."""


class CvBridge:
    pass
