"""Host-side mirror of the reference's correction_tools package: only the drift estimation that consumes the spot finder."""
from . import alignment, chromatic  # noqa: F401
