"""Stand-in for rospkg (oracle/shims/README.md): the package path is the reference checkout."""
import os


class RosPack:
    def get_path(self, name):
        return os.environ.get("COV_REFERENCE_ROOT", "/root/reference")
