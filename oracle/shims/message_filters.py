"""Empty stand-in (oracle/shims/README.md)."""
