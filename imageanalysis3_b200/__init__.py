"""imageanalysis3_b200 -- B200-native (sm_100a) spot finding behind ImageAnalysis3's entry points.

Package constants mirror the reference's ``__init__.py:4-19``.
"""
_distance_zxy = [200, 108, 108]
_sigma_zxy = [1.35, 1.9, 1.9]
_image_size = [30, 2048, 2048]
_allowed_colors = ['750', '647', '561', '488', '405']
_corr_channels = ['750', '647', '561']
_correction_folder = r'\\10.245.74.158\Chromatin_NAS_0\Corrections\Corrections_202008'
_image_dtype = 'uint16'
__version__ = '0.1.0'
