"""Mirror of the reference's ``spot_tools`` sub-package constants (spot_tools/__init__.py:2-8)."""
from .. import _distance_zxy, _sigma_zxy, _allowed_colors
_seed_th = {'750': 400, '647': 600, '561': 400}
