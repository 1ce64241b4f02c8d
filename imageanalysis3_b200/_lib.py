"""ctypes binding of libia3b200.so (C ABI in include/ia3b200.h).

There is deliberately no CPU fallback: if the library is missing or no B200 is visible the
calls raise.  CUDA is initialised lazily on first use inside the calling process, so the module
can be imported (and forked, e.g. by multiprocessing.Pool as the reference's callers do,
classes/field_of_view.py:1129) before any GPU work happens.
"""
import ctypes as C
import os
import threading
import weakref

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("IA3_LIB") or os.path.join(_HERE, "libia3b200.so")   # IA3_LIB: developer override (kernel variants)

DTYPE_U16, DTYPE_F32, DTYPE_F64 = 0, 1, 2
_NP2DT = {np.dtype(np.uint16): DTYPE_U16, np.dtype(np.float32): DTYPE_F32, np.dtype(np.float64): DTYPE_F64}


class IA3Error(RuntimeError):
    pass


class SeedCfg(C.Structure):
    _fields_ = [("w_fg", C.POINTER(C.c_double)), ("r_fg", C.c_int),
                ("w_bg", C.POINTER(C.c_double)), ("r_bg", C.c_int),
                ("filt_size", C.c_int), ("variant", C.c_int),
                ("edge", C.c_double), ("h_min", C.c_double), ("two_d", C.c_int)]


class SeedTiming(C.Structure):
    _fields_ = [("ms_gauss_fg", C.c_float), ("ms_gauss_bg", C.c_float), ("ms_rank", C.c_float),
                ("ms_compact", C.c_float), ("ms_total", C.c_float)]


class MomentCfg(C.Structure):
    _fields_ = [("radius", C.c_int), ("avoid_neighbors", C.c_int), ("recenter", C.c_int), ("bk_f", C.c_double)]


class FitCfg(C.Structure):
    _fields_ = [("personality", C.c_int), ("radius", C.c_int),
                ("min_w", C.c_double), ("max_w", C.c_double),
                ("init_w", C.c_double * 3), ("weight_sigma", C.c_double),
                ("maxfev", C.c_int), ("eval_fp32", C.c_int)]      # eval_fp32: reserved, 0


class FitOut(C.Structure):
    _fields_ = [("ps", C.c_void_p), ("p_raw", C.c_void_p), ("success", C.c_void_p), ("nfev", C.c_void_p), ("info", C.c_void_p),
                ("converged", C.c_void_p), ("dists", C.c_void_p), ("n_visits", C.c_void_p),
                ("success_old", C.c_void_p), ("centers_old", C.c_void_p)]


EXPORTS = [
    "ia3_init", "ia3_last_error", "ia3_version", "ia3_device_sm_count", "ia3_launch_count", "ia3_debug_stats",
    "ia3_timer_start", "ia3_timer_stop",
    "ia3_stack_create", "ia3_stack_wrap_device", "ia3_stack_destroy", "ia3_stack_trim",
    "ia3_stack_alloc", "ia3_stack_fetch", "ia3_corr_hot_pixels", "ia3_corr_zshift", "ia3_corr_highpass", "ia3_corr_mix", "ia3_corr_warp", "ia3_device_upload", "ia3_device_free",
    "ia3_seed_run", "ia3_seed_fetch", "ia3_seed_fetch_volume", "ia3_seed_gather_volume", "ia3_box_background",
    "ia3_seed_v2", "ia3_fft_gaussian", "ia3_seed_logratio", "ia3_stack_histogram",
    "ia3_fit_create", "ia3_fit_destroy", "ia3_fit_first_prepare", "ia3_fit_first_ties",
    "ia3_fit_first_resolve", "ia3_fit_run", "ia3_fit_first_run", "ia3_fit_repeat_sweep", "ia3_fit_engine_stats", "ia3_fit_get_volume",
    "ia3_fit_get_rec", "ia3_fit_num_levels", "ia3_fit_last_ms", "ia3_gaussfit_batch", "ia3_gauss_eval", "ia3_moment_fit",
]

_lib = None
# bytes this process copied through the C ABI (host->device, device->host); bench.py's e2e record
COPIED = {"h2d": 0, "d2h": 0}     # bytes moved over PCIe by this process (bench.py's e2e accounting)
_COPIED_LOCK = threading.Lock()


def _count(kind, nbytes):
    with _COPIED_LOCK:               # stacks are driven by many host threads; `+=` on a dict entry is not atomic
        COPIED[kind] += int(nbytes)


def load():
    """Load the shared library (no CUDA call is made here)."""
    global _lib
    if _lib is not None:
        return _lib
    # stacks processed by concurrent host threads each own a CUDA stream; give them separate hardware
    # queues (effective only if this process has not created its CUDA context yet)
    os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
    if not os.path.exists(LIB_PATH):
        raise IA3Error(f"{LIB_PATH} not found: build it with `python -m imageanalysis3_b200.build` "
                       "(there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, dbl = C.c_void_p, C.c_int, C.c_int64, C.c_double
    P = C.POINTER
    lib.ia3_last_error.restype = C.c_char_p
    lib.ia3_launch_count.restype = i64
    lib.ia3_fit_last_ms.restype = C.c_float
    lib.ia3_init.argtypes = [i32]
    lib.ia3_stack_create.argtypes = [vp, i32, i32, i32, i32, P(vp)]
    lib.ia3_stack_wrap_device.argtypes = [vp, i32, i32, i32, i32, P(vp)]
    lib.ia3_stack_destroy.argtypes = [vp]
    lib.ia3_stack_trim.argtypes = [vp, i32]
    lib.ia3_stack_alloc.argtypes = [i32, i32, i32, i32, P(vp)]
    lib.ia3_stack_fetch.argtypes = [vp, vp]
    lib.ia3_device_upload.argtypes = [vp, C.c_size_t, P(vp)]
    lib.ia3_device_free.argtypes = [vp]
    lib.ia3_corr_hot_pixels.argtypes = [vp, dbl, dbl, P(i64)]
    lib.ia3_corr_zshift.argtypes = [vp]
    lib.ia3_corr_highpass.argtypes = [vp, vp, i32]
    lib.ia3_corr_mix.argtypes = [P(vp), i32, vp, vp, i32, vp]
    lib.ia3_corr_warp.argtypes = [vp, vp, vp, i32, i32, vp]
    lib.ia3_seed_run.argtypes = [vp, P(SeedCfg), P(i64), P(SeedTiming)]
    lib.ia3_seed_fetch.argtypes = [vp, vp, vp, i64]
    lib.ia3_seed_fetch_volume.argtypes = [vp, i32, vp]
    lib.ia3_seed_gather_volume.argtypes = [vp, i32, vp, i64, vp]
    lib.ia3_box_background.argtypes = [vp, vp, i64, i32, i32, i32, i32, vp]
    lib.ia3_stack_histogram.argtypes = [vp, vp]
    lib.ia3_seed_v2.argtypes = [vp, i32, i32, dbl, P(dbl), vp, vp, i64, P(i64)]
    lib.ia3_fft_gaussian.argtypes = [vp, vp, i32, vp]
    lib.ia3_seed_logratio.argtypes = [vp, dbl, i32, dbl, P(dbl), vp, vp, i64, P(i64)]
    lib.ia3_fit_create.argtypes = [vp, vp, i64, P(FitCfg), P(vp)]
    lib.ia3_fit_destroy.argtypes = [vp]
    lib.ia3_fit_first_prepare.argtypes = [vp, P(i64)]
    lib.ia3_fit_first_ties.argtypes = [vp, vp, vp, i64]
    lib.ia3_fit_first_resolve.argtypes = [vp, vp, i64]
    lib.ia3_fit_run.argtypes = [vp, i32, dbl, dbl, dbl, i32, P(FitOut)]
    lib.ia3_fit_engine_stats.argtypes = [vp, vp, i32]
    lib.ia3_fit_first_run.argtypes = [vp, dbl, vp, vp, vp, vp, vp]
    lib.ia3_fit_repeat_sweep.argtypes = [vp, dbl, vp, vp, vp, vp, vp, vp]
    lib.ia3_fit_get_volume.argtypes = [vp, i32, vp]
    lib.ia3_fit_get_rec.argtypes = [vp, i64, vp, vp, P(i32)]
    lib.ia3_fit_num_levels.argtypes = [vp]
    lib.ia3_fit_last_ms.argtypes = [vp]
    lib.ia3_gaussfit_batch.argtypes = [P(FitCfg), dbl, i64, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.ia3_gauss_eval.argtypes = [P(FitCfg), dbl, vp, vp, vp, i64, vp]
    lib.ia3_moment_fit.argtypes = [vp, vp, i64, P(MomentCfg), vp]
    _lib = lib
    return lib


def _check(rc):
    if rc != 0:
        raise IA3Error(load().ia3_last_error().decode("utf-8", "replace"))


def init(device=-1):
    _check(load().ia3_init(int(device)))


def timer_start():
    _check(load().ia3_timer_start())


def timer_stop():
    ms = C.c_float(0)
    _check(load().ia3_timer_stop(C.byref(ms)))
    return float(ms.value)


def debug_stats(reset=False):
    if reset:
        load().ia3_debug_stats(None, 0)
        return ""
    buf = C.create_string_buffer(8192)
    load().ia3_debug_stats(buf, 8192)
    return buf.value.decode()


def launch_count():
    return int(load().ia3_launch_count())


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class DeviceArray:
    """A constant array (a correction profile) uploaded once and kept in device memory; rows of it are addressed
    by byte offset (``ptr(i)``)."""

    def __init__(self, a):
        a = np.ascontiguousarray(a)
        self.shape, self.dtype, self.nbytes = a.shape, a.dtype, a.nbytes
        self._d = C.c_void_p()
        _check(load().ia3_device_upload(_ptr(a), a.nbytes, C.byref(self._d)))
        _count("h2d", a.nbytes)
        self._fin = weakref.finalize(self, load().ia3_device_free, self._d)

    def ptr(self, row=None):
        if row is None:
            return C.c_void_p(self._d.value)
        return C.c_void_p(self._d.value + int(row) * (self.nbytes // self.shape[0]))

    def close(self):
        self._fin()


def _profile_arg(a, row=None):
    """(pointer, dtype, shape of the addressed part, host bytes copied by the call) of a numpy array or DeviceArray"""
    if a is None:
        return None, None, None, 0
    if isinstance(a, DeviceArray):
        return a.ptr(row), a.dtype, (a.shape if row is None else a.shape[1:]), 0
    a = np.ascontiguousarray(a if row is None else a[row])
    return _ptr(a), a.dtype, a.shape, a.nbytes


class Stack:
    """An image stack resident in HBM (one H2D copy shared by the seed and fit stages)."""

    def __init__(self, im=None, device_ptr=None, shape=None, dtype=None):
        lib = load()
        self._h = C.c_void_p()
        self._keep = None
        if im is None and device_ptr is None:        # an owned, uninitialised stack (destination of the corrections)
            self.shape = tuple(int(s) for s in shape)
            self.dtype = np.dtype(dtype)
            _check(lib.ia3_stack_alloc(_NP2DT[self.dtype], *self.shape, C.byref(self._h)))
        elif device_ptr is not None:
            self.shape = tuple(int(s) for s in shape)
            self.dtype = np.dtype(dtype)
            _check(lib.ia3_stack_wrap_device(C.c_void_p(int(device_ptr)), _NP2DT[self.dtype], *self.shape, C.byref(self._h)))
        else:
            if not isinstance(im, np.ndarray):
                raise TypeError(f"image given should be a numpy.ndarray, but {type(im)} is given.")
            if im.ndim != 3:
                raise ValueError("the device path needs a 3D (Z, X, Y) stack")
            if im.dtype not in _NP2DT:
                if im.dtype.kind in "ui" and im.size and im.min() >= 0 and im.max() <= 65535:
                    im = im.astype(np.uint16)
                elif im.dtype.kind in "uib":
                    im = im.astype(np.float64)
                else:
                    im = im.astype(np.float64)
            im = np.ascontiguousarray(im)
            self.shape = im.shape
            self.dtype = im.dtype
            self._keep = im
            _check(lib.ia3_stack_create(_ptr(im), _NP2DT[im.dtype], *im.shape, C.byref(self._h)))
            _count("h2d", im.nbytes)
            self._keep = None
        self._fin = weakref.finalize(self, lib.ia3_stack_destroy, self._h)

    @property
    def handle(self):
        return self._h

    def close(self):
        self._fin()

    def trim(self, what):
        """release device memory early: 1 = seed-stage work volumes, 2 = the library's copy of the image"""
        _check(load().ia3_stack_trim(self._h, int(what)))

    # ---- seed stage --------------------------------------------------------------------------
    def seed_candidates(self, w_fg, w_bg, filt_size, variant, edge, h_min, two_d=False):
        """Runs the device seed stage; returns (zxy int32 (n,3) in C order, h float32 (n,), timing)."""
        lib = load()
        cfg = SeedCfg()
        keep = []

        def half(w):
            if w is None:
                return None, -1
            w = np.ascontiguousarray(w, dtype=np.float64)
            keep.append(w)
            return w.ctypes.data_as(C.POINTER(C.c_double)), len(w) - 1

        cfg.w_fg, cfg.r_fg = half(w_fg)
        cfg.w_bg, cfg.r_bg = half(w_bg)
        cfg.filt_size = int(filt_size)
        cfg.variant = int(variant)
        cfg.edge = float(edge)
        cfg.h_min = float(h_min)
        cfg.two_d = int(bool(two_d))
        n = C.c_int64(0)
        t = SeedTiming()
        _check(lib.ia3_seed_run(self._h, C.byref(cfg), C.byref(n), C.byref(t)))
        n = n.value
        zxy = np.empty((n, 3), dtype=np.int32)
        h = np.empty((n,), dtype=np.float32)
        if n:
            _check(lib.ia3_seed_fetch(self._h, _ptr(zxy), _ptr(h), n))
        _count("d2h", zxy.nbytes + h.nbytes + 8)
        return zxy, h, t

    def box_background(self, boxes, first, last, bin_size, max_iter):
        """find_image_background (io_tools/load.py:642-686) of n boxes (n, 6) int32 [z0, z1, x0, x1, y0, y1)"""
        boxes = np.ascontiguousarray(boxes, dtype=np.int32).reshape(-1, 6)
        out = np.empty(len(boxes), dtype=np.float64)
        _check(load().ia3_box_background(self._h, _ptr(boxes), len(boxes), int(first), int(last), int(bin_size), int(max_iter), _ptr(out)))
        _count("h2d", boxes.nbytes)
        _count("d2h", out.nbytes)
        return out

    def moment_fit(self, centers_nx3, radius, avoid_neighbors=True, recenter=False, bk_f=0.1):
        """gfit_fast for every seed (Fitting_v4.py:433-447, 494-556) -> (n, 12) float64"""
        cen = np.ascontiguousarray(centers_nx3, dtype=np.float64).reshape(-1, 3)
        out = np.empty((len(cen), 12), dtype=np.float64)
        cfg = MomentCfg(int(radius), int(bool(avoid_neighbors)), int(bool(recenter)), float(bk_f))
        _check(load().ia3_moment_fit(self._h, _ptr(cen), len(cen), C.byref(cfg), _ptr(out)))
        _count("h2d", cen.nbytes)
        _count("d2h", out.nbytes)
        return out

    # ---- pre-processing (correct_fov_image's compute core) ----------------------------------------
    def fetch(self):
        """the resident image as a numpy array"""
        out = np.empty(self.shape, dtype=self.dtype)
        _check(load().ia3_stack_fetch(self._h, _ptr(out)))
        _count("d2h", out.nbytes)
        return out

    def remove_hot_pixels(self, hot_th=4, hot_pix_th=0.5):
        """corrections.Remove_Hot_Pixels in place -> number of hot columns found"""
        n = C.c_int64(0)
        _check(load().ia3_corr_hot_pixels(self._h, float(hot_th), float(hot_pix_th), C.byref(n)))
        return int(n.value)

    def z_shift_correct(self):
        """corrections.Z_Shift_Correction in place: planes scaled to the stack's median"""
        _check(load().ia3_corr_zshift(self._h))

    def gaussian_highpass(self, sigma=5, truncate=2):
        """correction_tools.filter.gaussian_high_pass_filter in place"""
        sd = float(sigma)
        r = int(truncate * sd + 0.5)                               # scipy.ndimage._filters.gaussian_filter1d
        x = np.arange(-r, r + 1)
        phi = np.exp(-0.5 / (sd * sd) * x ** 2)
        w = np.ascontiguousarray((phi / phi.sum())[r:])
        _check(load().ia3_corr_highpass(self._h, _ptr(w), r))

    @staticmethod
    def mix(ins, bleed=None, illum=None, out=None, bleed_row=None):
        """out = illumination(bleed-through(ins)); bleed (len(ins), X, Y) -- or row ``bleed_row`` of an (n, n, X, Y) array --,
        illum (X, Y); numpy arrays or DeviceArrays, all of one floating dtype"""
        ins = list(ins)
        out = Stack(shape=ins[0].shape, dtype=np.uint16) if out is None else out
        keep = []
        pb, dtb, shb, nb = _profile_arg(bleed, bleed_row)
        pi, dti, shi, ni = _profile_arg(illum)
        dts = {np.dtype(d) for d in (dtb, dti) if d is not None}
        if len(dts) > 1 or (dts and next(iter(dts)) not in (np.dtype(np.float32), np.dtype(np.float64))):
            raise TypeError("bleed-through and illumination profiles of one call are both float32 or both float64")
        f64 = bool(dts) and next(iter(dts)) == np.dtype(np.float64)
        if shb is not None and tuple(shb) != (len(ins),) + tuple(ins[0].shape[1:]):
            raise ValueError(f"bleed-through rows of shape {tuple(shb)} do not match {len(ins)} stacks of {ins[0].shape}")
        if shi is not None and tuple(shi) != tuple(ins[0].shape[1:]):
            raise ValueError(f"illumination profile of shape {tuple(shi)} does not match stacks of {ins[0].shape}")
        arr = (C.c_void_p * len(ins))(*[s._h for s in ins])
        _check(load().ia3_corr_mix(arr, len(ins), pb, pi, int(f64), out._h))
        _count("h2d", nb + ni)
        del keep
        return out

    def warp(self, drift=None, chroma=None, out=None):
        """map_coordinates(im, grid + chroma - drift, order 3, mode 'nearest') -> a new uint16 stack; chroma: numpy
        array or DeviceArray of shape (3, 1 or Z, X, Y)"""
        out = Stack(shape=self.shape, dtype=np.uint16) if out is None else out
        # float32 drifts (what correct_fov_image makes of a given drift) widen exactly; align_image's float64 result passes through
        d = None if drift is None else np.ascontiguousarray(np.asarray(drift), dtype=np.float64)
        cz, f64, pc, nc = 0, 0, None, 0
        if chroma is not None:
            if not isinstance(chroma, DeviceArray):
                chroma = np.ascontiguousarray(chroma)
                if chroma.dtype not in (np.dtype(np.float32), np.dtype(np.float64)):
                    chroma = chroma.astype(np.float64)                       # what numpy promotes int64 + profile to
            if len(chroma.shape) != 4 or chroma.shape[0] != 3 or tuple(chroma.shape[2:]) != tuple(self.shape[1:]) or chroma.shape[1] not in (1, self.shape[0]):
                raise ValueError(f"chromatic profile of shape {tuple(chroma.shape)} does not match a stack of {self.shape}")
            if chroma.dtype not in (np.dtype(np.float32), np.dtype(np.float64)):
                raise TypeError("chromatic profile: float32 or float64")
            cz, f64 = chroma.shape[1], int(chroma.dtype == np.float64)
            pc, _, _, nc = _profile_arg(chroma)
        _check(load().ia3_corr_warp(self._h, _ptr(d), pc, f64, int(cz), out._h))
        _count("h2d", nc)
        return out

    def histogram(self):
        """counts of every uint16 value (65536,) uint64"""
        out = np.zeros(65536, dtype=np.uint64)
        _check(load().ia3_stack_histogram(self._h, _ptr(out)))
        _count("d2h", out.nbytes)
        return out

    def seed_v2(self, gfilt_size, filt_size, th_seed):
        """Fitting_v4.get_seed_points_base_v2 on the device -> (flat C-order indices, im_norm values, std), in C order"""
        cap = 1 << 20
        while True:
            idx = np.empty(cap, dtype=np.int64)
            h = np.empty(cap, dtype=np.float32)
            std, n = C.c_double(0), C.c_int64(0)
            rc = load().ia3_seed_v2(self._h, int(gfilt_size), int(filt_size), float(th_seed), C.byref(std), _ptr(idx), _ptr(h), cap, C.byref(n))
            if rc == -2 and cap < (1 << 24):
                cap = 1 << 24
                continue
            _check(rc)
            break
        order = np.argsort(idx[:n.value], kind="stable")
        _count("d2h", 12 * n.value)
        return idx[:n.value][order], h[:n.value][order], np.float32(std.value)

    def seed_logratio(self, gfilt_size, filt_size, th_seed):
        """Fitting_v4.get_seed_points_base on the device -> (flat indices, im_diff values, std), in C order"""
        cap = 1 << 20
        while True:
            idx = np.empty(cap, dtype=np.int64)
            h = np.empty(cap, dtype=np.float64)
            std, n = C.c_double(0), C.c_int64(0)
            rc = load().ia3_seed_logratio(self._h, float(gfilt_size), int(filt_size), float(th_seed), C.byref(std), _ptr(idx), _ptr(h), cap, C.byref(n))
            if rc == -2 and cap < (1 << 24):
                cap = 1 << 24
                continue
            _check(rc)
            break
        order = np.argsort(idx[:n.value], kind="stable")
        _count("d2h", 16 * n.value)
        return idx[:n.value][order], h[:n.value][order], float(std.value)

    def fft_gaussian(self, gaus, exp=8):
        g = np.ascontiguousarray(np.broadcast_to(np.asarray(gaus, dtype=np.float64), (3,)))
        out = np.empty(self.shape, dtype=np.float64)
        _check(load().ia3_fft_gaussian(self._h, _ptr(g), int(exp), _ptr(out)))
        _count("d2h", out.nbytes)
        return out

    def seed_volume(self, which):
        out = np.empty(self.shape, dtype=self.dtype)
        _check(load().ia3_seed_fetch_volume(self._h, int(which), _ptr(out)))
        return out

    def seed_volume_at(self, which, flat_idx):
        """values of the foreground (0) / background (1) blur at flat C-order voxel indices"""
        idx = np.ascontiguousarray(flat_idx, dtype=np.int64).ravel()
        out = np.empty(len(idx), dtype=self.dtype)
        _check(load().ia3_seed_gather_volume(self._h, int(which), _ptr(idx), len(idx), _ptr(out)))
        _count("h2d", idx.nbytes)
        _count("d2h", out.nbytes)
        return out


def make_fit_cfg(personality, radius, min_w, max_w, init_w, weight_sigma=0.0, maxfev=0):
    cfg = FitCfg()
    cfg.personality = int(personality)
    cfg.radius = int(radius)
    cfg.min_w = float(min_w)
    cfg.max_w = float(max_w)
    iw = np.broadcast_to(np.asarray(init_w, dtype=np.float64).ravel()[:3] if np.ndim(init_w) else float(init_w), (3,))
    for i in range(3):
        cfg.init_w[i] = float(iw[i])
    cfg.weight_sigma = float(weight_sigma)
    cfg.maxfev = int(maxfev)
    cfg.eval_fp32 = 0
    return cfg


class FitHandle:
    """Device state of one iter_fit_seed_points object."""

    def __init__(self, stack, centers_nx3, cfg):
        lib = load()
        self.stack = stack
        self.centers = np.ascontiguousarray(centers_nx3, dtype=np.float64).reshape(-1, 3)
        self.n = len(self.centers)
        self._h = C.c_void_p()
        _check(lib.ia3_fit_create(stack.handle, _ptr(self.centers), self.n, C.byref(cfg), C.byref(self._h)))
        self._fin = weakref.finalize(self, lib.ia3_fit_destroy, self._h)
        _count("h2d", self.centers.nbytes * 2)
        self.ps = np.full((self.n, 11), np.nan, dtype=np.float32)
        self.p_raw = np.full((self.n, 10), np.nan, dtype=np.float64)
        self.success = np.zeros(self.n, dtype=np.uint8)
        self.nfev = np.zeros(self.n, dtype=np.int32)
        self.info = np.zeros(self.n, dtype=np.int32)
        self.converged = np.zeros(self.n, dtype=np.uint8)
        self.dists = np.full(self.n, np.inf, dtype=np.float64)
        self.n_visits = np.zeros(self.n, dtype=np.int32)
        self.success_old = np.zeros(self.n, dtype=np.uint8)
        self.centers_old = np.full((self.n, 3), np.nan, dtype=np.float32)

    def close(self):
        self._fin()

    def first_prepare(self):
        n = C.c_int64(0)
        _check(load().ia3_fit_first_prepare(self._h, C.byref(n)))
        _count("d2h", 8 * min(int(n.value), 16384) + 256)
        return n.value

    def first_ties(self, n):
        spot = np.empty(n, dtype=np.int32)
        zxy = np.empty((n, 3), dtype=np.int32)
        if n:
            _check(load().ia3_fit_first_ties(self._h, _ptr(spot), _ptr(zxy), n))
        return spot, zxy

    def first_resolve(self, keep):
        keep = np.ascontiguousarray(keep, dtype=np.uint8)
        _check(load().ia3_fit_first_resolve(self._h, _ptr(keep), len(keep)))
        _count("h2d", keep.nbytes)

    def run(self, phases, min_delta_center, max_delta_center, max_dist_th2, n_max_iter):
        """firstfit (1), repeatfit (2) or both (3) on the device; fills every result array of the handle"""
        out = FitOut(*[a.ctypes.data for a in (self.ps, self.p_raw, self.success, self.nfev, self.info, self.converged,
                                               self.dists, self.n_visits, self.success_old, self.centers_old)])
        _check(load().ia3_fit_run(self._h, int(phases), float(min_delta_center), float(max_delta_center), float(max_dist_th2),
                                  int(n_max_iter), C.byref(out)))
        _count("d2h", self._result_bytes() + self.n * (1 + 8 + 4 + 1 + 12))

    def first_run(self, delta_center):
        _check(load().ia3_fit_first_run(self._h, float(delta_center), _ptr(self.ps), _ptr(self.p_raw),
                                        _ptr(self.success), _ptr(self.nfev), _ptr(self.info)))
        _count("d2h", self._result_bytes())

    def repeat_sweep(self, delta_center, active):
        active = np.ascontiguousarray(active, dtype=np.uint8)
        _check(load().ia3_fit_repeat_sweep(self._h, float(delta_center), _ptr(active), _ptr(self.ps), _ptr(self.p_raw),
                                           _ptr(self.success), _ptr(self.nfev), _ptr(self.info)))
        _count("d2h", self._result_bytes())
        _count("h2d", self.n)

    def engine_stats(self, trace=False):
        out = np.zeros(16 + 1024, dtype=np.int64)
        _check(load().ia3_fit_engine_stats(self._h, _ptr(out), len(out)))
        keys = ("rounds", "tasks", "lm_runs", "evals", "memo_hits", "spec_runs", "spec_hits", "parked", "team_tasks", "bricks")
        st = dict(zip(keys, (int(v) for v in out[:10])))
        if out[15] > 0:          # IA3_FIT_PROF build: cycles of the team kernel's phases per evaluation
            st["team_cycles_per_eval"] = {k: round(float(out[10 + i]) / float(out[15]), 1)
                                          for i, k in enumerate(("outer", "lmpar", "consts", "pass", "judge"))}
        if trace:
            nr = min(st["rounds"], 512)
            w = out[16:16 + 2 * nr:2]
            ns = out[17:17 + 2 * nr:2]
            st["trace"] = [(int(x >> 16), int(x & 0xffff), round(float(t - ns[0]) * 1e-6, 3)) for x, t in zip(w, ns)]
        return st

    def _result_bytes(self):
        return self.ps.nbytes + self.p_raw.nbytes + self.success.nbytes + self.nfev.nbytes + self.info.nbytes

    def volume(self, which):
        out = np.empty(self.stack.shape, dtype=np.float64)
        _check(load().ia3_fit_get_volume(self._h, int(which), _ptr(out)))
        return out

    def rec(self, i, K=8192):
        rec = np.empty(K, dtype=np.float64)
        zxy = np.empty((K, 3), dtype=np.int32)
        cnt = C.c_int32(0)
        _check(load().ia3_fit_get_rec(self._h, int(i), _ptr(rec), _ptr(zxy), C.byref(cnt)))
        return rec[:cnt.value].copy(), zxy[:cnt.value].copy()

    @property
    def num_levels(self):
        return int(load().ia3_fit_num_levels(self._h))

    @property
    def last_ms(self):
        return float(load().ia3_fit_last_ms(self._h))


def gaussfit_batch(cfg, delta_center, values_list, coords_list, centers, want_rec=False):
    """Batch of independent GaussianFit problems.  values_list[i]: (m_i,) float64; coords_list[i]: (3, m_i)."""
    lib = load()
    n = len(values_list)
    off = np.zeros(n + 1, dtype=np.int64)
    for i, v in enumerate(values_list):
        off[i + 1] = off[i] + len(v)
    values = np.ascontiguousarray(np.concatenate([np.asarray(v, dtype=np.float64).ravel() for v in values_list]) if n else np.zeros(0))
    coords = np.ascontiguousarray(np.concatenate([np.asarray(c, dtype=np.float32).T.reshape(-1, 3) for c in coords_list]) if n else np.zeros((0, 3), np.float32), dtype=np.float32)
    centers = np.ascontiguousarray(centers, dtype=np.float64).reshape(-1, 3)
    ps = np.empty((n, 11), np.float32)
    praw = np.empty((n, 10), np.float64)
    succ = np.zeros(n, np.uint8)
    nfev = np.zeros(n, np.int32)
    info = np.zeros(n, np.int32)
    rec = np.empty(len(values), np.float64) if want_rec else None
    _check(lib.ia3_gaussfit_batch(C.byref(cfg), float(delta_center), n, _ptr(off), _ptr(values), _ptr(coords),
                                  _ptr(centers), _ptr(ps), _ptr(praw), _ptr(succ), _ptr(nfev), _ptr(info), _ptr(rec)))
    recs = [rec[off[i]:off[i + 1]] for i in range(n)] if want_rec else None
    return ps, praw, succ.astype(bool), nfev, info, recs


def gauss_eval(cfg, delta_center, p_raw, center, coords_3xm):
    """GaussianFit.get_im(): Gaussian part of the model on (3, m) coordinates -> (m,) float64."""
    co = np.ascontiguousarray(np.asarray(coords_3xm, dtype=np.float32).T.reshape(-1, 3))
    p_raw = np.ascontiguousarray(p_raw, dtype=np.float64)
    center = np.ascontiguousarray(center, dtype=np.float64)
    out = np.empty(len(co), dtype=np.float64)
    _check(load().ia3_gauss_eval(C.byref(cfg), float(delta_center), _ptr(p_raw), _ptr(center), _ptr(co), len(co), _ptr(out)))
    return out
