"""Mirror of the reference's ``External`` sub-package (External/__init__.py:3)."""
from .. import _sigma_zxy
