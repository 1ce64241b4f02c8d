"""Spot-coordinate translation functions (reference correction_tools/chromatic.py:41-114, 415-438): what
``correct_fov_image(warp_image=False)`` returns instead of warping the images -- the fitted polynomial chromatic shift and
the drift applied to spot coordinates.  Host code on (n, 3) or (n, 11) spot tables."""
import itertools
import pickle

import numpy as np


def generate_polynomial_data(coords, max_order):
    """design matrix of all monomials of the coordinate columns up to ``max_order`` (n_points x n_terms), in the
    order of itertools.combinations_with_replacement"""
    cols = []
    for order in range(int(max_order) + 1):
        for combo in itertools.combinations_with_replacement(coords.transpose(), order):
            term = np.ones(np.shape(coords)[0])
            for v in combo:
                term *= v
            cols.append(term)
    return np.array(cols).transpose()


def generate_chromatic_function(chromatic_const_file, drift=None):
    """f(coords) -> coords - polynomial chromatic shift + drift, from a constants dict / its .pkl file; with None: the drift
    alone (or the identity when there is no drift either)"""
    if isinstance(chromatic_const_file, dict):
        info = dict(chromatic_const_file)
    elif isinstance(chromatic_const_file, str):
        with open(chromatic_const_file, 'rb') as fh:
            info = pickle.load(fh)
    elif chromatic_const_file is None:
        if drift is None:
            return lambda _coords, _drift=None: _coords
        info = {'constants': [np.array([0]) for _ in drift], 'fitting_orders': np.zeros(len(drift), dtype=int),
                'ref_center': np.zeros(len(drift))}
    else:
        raise TypeError("Wrong input chromatic_const_file")
    consts, orders, centre = info['constants'], info['fitting_orders'], info['ref_center']
    shift0 = np.zeros(len(centre)) if drift is None else drift[:len(centre)]

    def _shift_function(_coords, _drift=shift0, _consts=consts, _fitting_orders=orders, _ref_center=centre):
        if len(_coords) == 0:
            return _coords
        _coords = np.array(_coords)
        ndim = len(_ref_center)
        if np.shape(_coords)[1] == ndim:
            pts = _coords.copy()
        elif np.shape(_coords)[1] == 11:          # a fit_fov_image spot table
            pts = _coords.copy()[:, 1:1 + ndim]
        else:
            raise ValueError("Wrong input coords")
        shifts = np.array([np.dot(generate_polynomial_data(pts - _ref_center[np.newaxis, :], order), const)
                           for const, order in zip(_consts, _fitting_orders)]).transpose()
        moved = pts - shifts + _drift
        if np.shape(_coords)[1] == ndim:
            return moved
        out = _coords.copy()
        out[:, 1:1 + ndim] = moved
        return out

    return _shift_function
