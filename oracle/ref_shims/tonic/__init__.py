"""Stand-in for the third-party `tonic` package (absent offline; unpinned, README.md:29).
TEST INFRASTRUCTURE ONLY. See transforms.py."""
from . import transforms  # noqa: F401
