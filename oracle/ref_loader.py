"""TEST INFRASTRUCTURE ONLY -- loads the *real* reference modules under import shims.

This file is part of ``oracle/`` and must never be imported by the product package
(``imageanalysis3_b200``).  It only works where ``/root/reference`` exists (the build
container); the GPU box does not have it.  It is used to
  * pin the restated oracle (``oracle/seed_oracle.py``, ``oracle/fit_oracle.py``) against the
    reference's own code, and
  * generate the committed fixtures under ``tests/golden/`` (``oracle/make_golden.py``).

Shims: the reference was written against numpy 1.x / old scipy and
imports GUI / FFT packages that are not installed here:
  np.int / np.float aliases               spot_tools/fitting.py:60,63,150
  scipy.signal.gaussian                   External/Fitting_v4.py:13
  pyfftw.interfaces.numpy_fft             External/Fitting_v4.py:5-6
  matplotlib.pyplot                       External/Fitting_v4.py:753
  parent package constants                __init__.py:4-19, spot_tools/__init__.py:4-8
"""
import ast
import importlib.util
import os
import sys
import types
import warnings

import numpy as np

REF_ROOT = os.environ.get("IA3_REFERENCE_ROOT", "/root/reference")
_PKG = "IA3REF"
_cache = {}


def available():
    return os.path.isfile(os.path.join(REF_ROOT, "spot_tools", "fitting.py"))


def _install_shims():
    import scipy.fft
    import scipy.signal
    import scipy.signal.windows

    if not hasattr(np, "int"):
        np.int = int
    if not hasattr(np, "float"):
        np.float = float
    if not hasattr(scipy.signal, "gaussian"):
        scipy.signal.gaussian = scipy.signal.windows.gaussian
    if "pyfftw" not in sys.modules:
        m0 = types.ModuleType("pyfftw")
        m1 = types.ModuleType("pyfftw.interfaces")
        m2 = types.ModuleType("pyfftw.interfaces.numpy_fft")
        m2.rfftn = scipy.fft.rfftn
        m2.irfftn = scipy.fft.irfftn
        m0.interfaces = m1
        m1.numpy_fft = m2
        sys.modules.update({"pyfftw": m0, "pyfftw.interfaces": m1, "pyfftw.interfaces.numpy_fft": m2})
    if "matplotlib" not in sys.modules:
        try:
            import matplotlib  # noqa: F401
        except Exception:
            mm = types.ModuleType("matplotlib")
            mp = types.ModuleType("matplotlib.pyplot")
            mm.pyplot = mp
            sys.modules.update({"matplotlib": mm, "matplotlib.pyplot": mp})


def _load(name, relpath):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF_ROOT, relpath))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        spec.loader.exec_module(mod)
    return mod


def load():
    """Return a namespace with .Fitting_v3, .Fitting_v4, .fitting, .visual (lifted functions)."""
    if "ns" in _cache:
        return _cache["ns"]
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    _install_shims()
    sigma_zxy = [1.35, 1.9, 1.9]
    root = types.ModuleType(_PKG)
    root.__path__ = []
    root._correction_folder = ""
    root._temp_folder = ""
    root._distance_zxy = [200, 108, 108]
    root._sigma_zxy = sigma_zxy
    root._allowed_colors = ["750", "647", "561", "488", "405"]
    root._image_size = [30, 2048, 2048]
    sys.modules[_PKG] = root

    ext = types.ModuleType(_PKG + ".External")
    ext.__path__ = [os.path.join(REF_ROOT, "External")]
    ext._sigma_zxy = sigma_zxy
    sys.modules[_PKG + ".External"] = ext
    root.External = ext

    st = types.ModuleType(_PKG + ".spot_tools")
    st.__path__ = []
    st._seed_th = {"750": 400, "647": 600, "561": 400}
    sys.modules[_PKG + ".spot_tools"] = st
    root.spot_tools = st

    vt = types.ModuleType(_PKG + ".visual_tools")
    vt.get_seed_points_base = lambda *a, **k: None  # imported but unused by spot_tools/fitting.py:12
    sys.modules[_PKG + ".visual_tools"] = vt
    root.visual_tools = vt

    _lift_io(root)
    v3 = _load(_PKG + ".External.Fitting_v3", "External/Fitting_v3.py")
    v4 = _load(_PKG + ".External.Fitting_v4", "External/Fitting_v4.py")
    ext.Fitting_v3 = v3
    ext.Fitting_v4 = v4
    fitting = _load(_PKG + ".spot_tools.fitting", "spot_tools/fitting.py")

    ns = types.SimpleNamespace(Fitting_v3=v3, Fitting_v4=v4, fitting=fitting, visual=_lift_visual(v3, sigma_zxy))
    _cache["ns"] = ns
    return ns


def _lift_nodes(relpath, names, env, kinds=(ast.FunctionDef, ast.ClassDef)):
    with open(os.path.join(REF_ROOT, relpath), "r", encoding="utf-8", errors="replace") as fh:
        tree = ast.parse(fh.read())
    nodes = [n for n in tree.body if isinstance(n, kinds) and n.name in names]
    code = compile(ast.Module(body=nodes, type_ignores=[]), f"<reference {relpath} (lifted)>", "exec")
    exec(code, env)
    return env


def _lift_io(root):
    """fit_fov_image(normalize_local=True) imports io_tools.load.find_image_background
    (io_tools/load.py:642-686) and io_tools.crop.generate_neighboring_crop (io_tools/crop.py:59-88, which
    needs classes.preprocess.ImageCrop, classes/preprocess.py:17-93).  Those packages do not import here
    (h5py, skimage, ...): lift the three definitions by AST into stand-in modules."""
    import scipy
    import scipy.signal
    pre = types.ModuleType(_PKG + ".classes.preprocess")
    _lift_nodes("classes/preprocess.py", {"ImageCrop"}, pre.__dict__.update(np=np, _image_size=root._image_size) or pre.__dict__)
    classes = types.ModuleType(_PKG + ".classes")
    classes.__path__ = []
    classes.preprocess = pre
    load_m = types.ModuleType(_PKG + ".io_tools.load")
    load_m.__dict__.update(np=np, scipy=scipy, _image_dtype="uint16")
    _lift_nodes("io_tools/load.py", {"find_image_background"}, load_m.__dict__)
    crop_m = types.ModuleType(_PKG + ".io_tools.crop")
    crop_m.__package__ = _PKG + ".io_tools"
    crop_m.__dict__.update(np=np, _image_size=root._image_size)
    _lift_nodes("io_tools/crop.py", {"generate_neighboring_crop"}, crop_m.__dict__)
    io = types.ModuleType(_PKG + ".io_tools")
    io.__path__ = []
    io.load, io.crop = load_m, crop_m
    sys.modules.update({_PKG + ".classes": classes, _PKG + ".classes.preprocess": pre, _PKG + ".io_tools": io,
                        _PKG + ".io_tools.load": load_m, _PKG + ".io_tools.crop": crop_m})
    root.classes, root.io_tools = classes, io


def _lift_visual(v3, sigma_zxy):
    """visual_tools.py does not import here (matplotlib/skimage/...); lift the four seeding
    functions (visual_tools.py:260-381, 1775-1870, 3081-3139) out of the file by AST."""
    from scipy.ndimage import gaussian_filter, maximum_filter, minimum_filter, median_filter

    wanted = {"get_STD_centers", "get_seed_points_base", "get_seed_in_distance", "find_matched_seeds",
              "select_sparse_centers"}
    with open(os.path.join(REF_ROOT, "visual_tools.py"), "r", encoding="utf-8", errors="replace") as fh:
        tree = ast.parse(fh.read())
    nodes = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in wanted]
    env = dict(np=np, gaussian_filter=gaussian_filter, maximum_filter=maximum_filter,
               minimum_filter=minimum_filter, median_filter=median_filter,
               Fitting_v3=v3, _sigma_zxy=sigma_zxy, os=os, sys=sys)
    code = compile(ast.Module(body=nodes, type_ignores=[]), "<reference visual_tools.py (lifted)>", "exec")
    exec(code, env)
    return types.SimpleNamespace(**{k: env[k] for k in wanted if k in env})


def load_corrections():
    """The reference's pre-processing step: io_tools/load.py correct_fov_image (:166-521) with the helpers it calls
    (get_num_frame :17-45, split_im_by_channels :524-550, load_correction_profile :553-640), corrections.py
    Remove_Hot_Pixels (:490-510) and the .dax reader (visual_tools.py:906-1089), lifted by AST because neither
    package imports here.  -> namespace with .correct_fov_image, .Remove_Hot_Pixels, .split_im_by_channels"""
    if "corr" in _cache:
        return _cache["corr"]
    load()
    from scipy.ndimage import map_coordinates, shift
    import scipy
    import time as _time
    vt_env = dict(np=np, os=os, sys=sys, _correction_folder="")
    _lift_nodes("visual_tools.py", {"Reader", "DaxReader"}, vt_env)
    vt = sys.modules[_PKG + ".visual_tools"]
    vt.DaxReader = vt_env["DaxReader"]
    cor_env = dict(np=np, os=os)
    _lift_nodes("corrections.py", {"Remove_Hot_Pixels", "Z_Shift_Correction"}, cor_env)
    corrections = types.SimpleNamespace(Remove_Hot_Pixels=cor_env["Remove_Hot_Pixels"], Z_Shift_Correction=cor_env["Z_Shift_Correction"])
    load_m = sys.modules[_PKG + ".io_tools.load"]
    env = load_m.__dict__
    env.update(np=np, os=os, sys=sys, time=_time, scipy=scipy, map_coordinates=map_coordinates, shift=shift, corrections=corrections,
               _image_size=[30, 2048, 2048], _allowed_colors=["750", "647", "561", "488", "405"], _corr_channels=["750", "647", "561"],
               _correction_folder="", _num_buffer_frames=10, _num_empty_frames=0, __package__=_PKG + ".io_tools",
               __name__=_PKG + ".io_tools.load")
    _lift_nodes("io_tools/load.py", {"get_num_frame", "correct_fov_image", "split_im_by_channels", "load_correction_profile"}, env)
    ns = types.SimpleNamespace(correct_fov_image=env["correct_fov_image"], Remove_Hot_Pixels=cor_env["Remove_Hot_Pixels"],
                               split_im_by_channels=env["split_im_by_channels"], load_correction_profile=env["load_correction_profile"])
    _cache["corr"] = ns
    return ns


def _lift_assigns(relpath, names, env):
    """module-level ``name = literal`` statements of the reference file, evaluated in env"""
    with open(os.path.join(REF_ROOT, relpath), "r", encoding="utf-8", errors="replace") as fh:
        tree = ast.parse(fh.read())
    nodes = [n for n in tree.body if isinstance(n, ast.Assign) and len(n.targets) == 1 and isinstance(n.targets[0], ast.Name)
             and n.targets[0].id in names]
    exec(compile(ast.Module(body=nodes, type_ignores=[]), f"<reference {relpath} (lifted assignments)>", "exec"), env)
    return env


def load_alignment():
    """The reference's drift estimation: correction_tools/alignment.py align_image (:527-696), align_beads (:139-217),
    generate_drift_crops (:87-136) with alignment_tools.py fft3d_from2d / fftalign_2d (:286-353) and spot_tools/matching.py
    find_paired_centers / check_paired_centers (:148-287), lifted by AST.  skimage is not installed here, so
    ``phase_cross_correlation`` is a stub that raises: only the bead-fitting mode (use_autocorr=False) can run.
    -> namespace with .align_image, .align_beads, .generate_drift_crops, .fft3d_from2d, .find_paired_centers, .check_paired_centers"""
    if "align" in _cache:
        return _cache["align"]
    load()
    load_corrections()
    import time as _time
    root = sys.modules[_PKG]
    if "skimage" not in sys.modules:
        try:
            import skimage.registration  # noqa: F401
        except Exception:
            sk, reg = types.ModuleType("skimage"), types.ModuleType("skimage.registration")

            def phase_cross_correlation(*a, **k):
                raise RuntimeError("skimage is not installed: the phase-correlation mode of align_image cannot run here")
            reg.phase_cross_correlation = phase_cross_correlation
            sk.registration = reg
            sys.modules.update({"skimage": sk, "skimage.registration": reg})
    at = types.ModuleType(_PKG + ".alignment_tools")
    at.__dict__.update(np=np, os=os, sys=sys)
    _lift_nodes("alignment_tools.py", {"fft3d_from2d", "fftalign_2d", "blurnorm2d", "translation_align_pts"}, at.__dict__)
    sys.modules[_PKG + ".alignment_tools"] = at
    root.alignment_tools = at
    mt = types.ModuleType(_PKG + ".spot_tools.matching")
    mt.__dict__.update(np=np, os=os, sys=sys)
    _lift_nodes("spot_tools/matching.py", {"find_paired_centers", "check_paired_centers"}, mt.__dict__)
    sys.modules[_PKG + ".spot_tools.matching"] = mt
    sys.modules[_PKG + ".spot_tools"].matching = mt
    ct = _correction_tools_pkg()
    al = types.ModuleType(_PKG + ".correction_tools.alignment")
    al.__dict__.update(np=np, os=os, time=_time, _allowed_colors=root._allowed_colors, _image_size=root._image_size, _num_buffer_frames=10,
                       _num_empty_frames=0, _image_dtype="uint16", _correction_folder="", __package__=_PKG + ".correction_tools",
                       __name__=_PKG + ".correction_tools.alignment")
    _lift_assigns("correction_tools/alignment.py", {"_default_align_corr_args", "_default_align_fitting_args"}, al.__dict__)
    _lift_nodes("correction_tools/alignment.py", {"_find_boundary", "generate_drift_crops", "align_beads", "align_image"}, al.__dict__)
    ct.alignment = al
    sys.modules[_PKG + ".correction_tools.alignment"] = al
    ns = types.SimpleNamespace(align_image=al.align_image, align_beads=al.align_beads, generate_drift_crops=al.generate_drift_crops,
                               fft3d_from2d=at.fft3d_from2d, fftalign_2d=at.fftalign_2d, find_paired_centers=mt.find_paired_centers,
                               check_paired_centers=mt.check_paired_centers, defaults=(al._default_align_corr_args, al._default_align_fitting_args))
    _cache["align"] = ns
    return ns


def _correction_tools_pkg():
    name = _PKG + ".correction_tools"
    if name not in sys.modules:
        ct = types.ModuleType(name)
        ct.__path__ = []
        sys.modules[name] = ct
        sys.modules[_PKG].correction_tools = ct
    return sys.modules[name]


def load_chromatic():
    """correction_tools/chromatic.py generate_chromatic_function (:41-114) and generate_polynomial_data (:415-438), lifted:
    what the reference's correct_fov_image(warp_image=False) imports, and correction_tools/filter.py gaussian_high_pass_filter
    (gaussian_highpass=True).  -> namespace with the three functions"""
    if "chrom" in _cache:
        return _cache["chrom"]
    load()
    import pickle
    ct = _correction_tools_pkg()
    ch = types.ModuleType(_PKG + ".correction_tools.chromatic")
    ch.__dict__.update(np=np, os=os, pickle=pickle)
    _lift_nodes("correction_tools/chromatic.py", {"generate_chromatic_function", "generate_polynomial_data"}, ch.__dict__)
    ct.chromatic = ch
    sys.modules[_PKG + ".correction_tools.chromatic"] = ch
    # correction_tools/filter.py:14-19 gaussian_high_pass_filter (its module imports the removed scipy.ndimage.filters namespace)
    from scipy.ndimage import gaussian_filter
    fl = types.ModuleType(_PKG + ".correction_tools.filter")
    fl.__dict__.update(np=np, gaussian_filter=gaussian_filter)
    _lift_nodes("correction_tools/filter.py", {"gaussian_high_pass_filter"}, fl.__dict__)
    ct.filter = fl
    sys.modules[_PKG + ".correction_tools.filter"] = fl
    ns = types.SimpleNamespace(generate_chromatic_function=ch.generate_chromatic_function, generate_polynomial_data=ch.generate_polynomial_data,
                               gaussian_high_pass_filter=fl.gaussian_high_pass_filter)
    _cache["chrom"] = ns
    return ns
