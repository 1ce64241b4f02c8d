"""Host-side mirror of the reference's io_tools package: only what sits directly upstream of the spot-finding path."""
from . import load  # noqa: F401
