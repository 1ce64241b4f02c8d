"""correct_fov_image on the device -- the step that hands fit_fov_image its stacks (reference io_tools/load.py:160-520).

The reference reads a multi-colour .dax movie, splits it by channel and runs, on the host with numpy / scipy.ndimage,
hot-pixel removal, bleed-through mixing, illumination flattening and a cubic-spline warp that undoes chromatic
aberration and stage drift.  Here the file is read on the host (it is I/O), each channel is uploaded once, and every
correction runs on the resident uint16 stacks through libia3b200 (ia3_corr_*); the corrected stacks can be handed
to spot_tools.fitting.fit_fov_image without leaving the device (``return_stacks=True``).

Same signature, argument meaning and error behaviour as the reference.  Branches of the reference that leave this
path raise NotImplementedError instead of computing something else: calculate_drift with use_autocorr=True (the phase
correlation of correction_tools/alignment.py align_image needs skimage; its bead-fitting mode, use_autocorr=False, is built),
normalization and a non-uint16 output_dtype.  ``warp_image=False`` returns, like the
reference, the unwarped images plus one spot-coordinate function per channel (correction_tools/chromatic.py).

One reference behaviour is kept on purpose because results must be the same: the warp code in the reference sits
inside its ``if verbose:`` block (io_tools/load.py:436-459), so images are only warped when ``verbose`` is true.
``force_warp=True`` (not a reference argument) warps regardless.
"""
import os
import re
import time

import numpy as np

from .. import _allowed_colors, _corr_channels, _correction_folder, _image_size
from .. import _lib


def get_num_frame(dax_filename, frame_per_color=_image_size[0], buffer_frame=10, empty_frame=0, verbose=False):
    """[frames, dx, dy] and the number of colours of a .dax movie, from its .inf file (reference io_tools/load.py:17-45)"""
    if '.dax' not in dax_filename:
        raise ValueError(f"Wrong input type, .dax file expected for {dax_filename}")
    if not os.path.isfile(dax_filename):
        raise IOError(f"input file:{dax_filename} doesn't exist!")
    n_frame = n_color = dx = dy = 0
    with open(dax_filename.replace('.dax', '.inf'), 'r') as fh:
        for line in fh:
            line = line.rstrip()
            if "number of frames" in line:
                n_frame = int(line.split('=')[1])
                colors = (n_frame - 2 * buffer_frame - empty_frame) / frame_per_color
                if colors != int(colors):
                    raise ValueError("Wrong num_color, should be integer!")
                n_color = int(colors)
            if "frame dimensions" in line:
                dims = line.split('=')[1].split('x')
                dx, dy = int(dims[0]), int(dims[1])
    return [n_frame, dx, dy], n_color


def read_dax(dax_filename):
    """the whole movie as (frames, width, height) uint16 in native byte order (visual_tools.py:976-1073 DaxReader.loadAll)"""
    inf = os.path.splitext(dax_filename)[0] + '.inf'
    height = width = frames = None
    big = False
    with open(inf, 'r') as fh:
        for line in fh:
            m = re.match(r'frame dimensions = ([\d]+) x ([\d]+)', line)
            if m:
                height, width = int(m.group(1)), int(m.group(2))
            m = re.match(r'number of frames = ([\d]+)', line)
            if m:
                frames = int(m.group(1))
            m = re.search(r' (big|little) endian', line)
            if m:
                big = m.group(1) == 'big'
    if not height:
        height = width = 256          # the reference's fallback
    data = np.fromfile(dax_filename, dtype='>u2' if big else '<u2')
    if frames is None:
        frames = data.size // (height * width)
    return data.reshape(frames, width, height).astype(np.uint16, copy=False)


def split_im_by_channels(im, sel_channels, all_channels, single_im_size=_image_size,
                         num_buffer_frames=10, num_empty_frames=0, skip_frame0=False):
    """one (Z, X, Y) stack per selected channel out of an interleaved movie (reference io_tools/load.py:523-550)"""
    if isinstance(sel_channels, (str, int)):
        sel_channels = [sel_channels]
    if isinstance(all_channels, (str, int)):
        all_channels = [all_channels]
    sel, every = [str(c) for c in sel_channels], [str(c) for c in all_channels]
    for ch in sel:
        if ch not in every:
            raise ValueError(f"Wrong input channel:{ch}, should be within {every}")
    n_col, lead = len(every), num_empty_frames + num_buffer_frames
    starts = [lead + (every.index(ch) - lead) % n_col for ch in sel]
    if skip_frame0:
        starts = [s + n_col if s == num_buffer_frames else s for s in starts]
    return [np.ascontiguousarray(im[s:s + single_im_size[0] * n_col:n_col]) for s in starts]


def load_correction_profile(corr_type, corr_channels=_corr_channels, correction_folder=_correction_folder,
                            all_channels=_allowed_colors, ref_channel='647', im_size=_image_size, verbose=False):
    """the saved correction profiles, by the reference's file naming (reference io_tools/load.py:553-637)"""
    kinds = ['chromatic', 'illumination', 'bleedthrough', 'chromatic_constants']
    kind = str(corr_type).lower()
    if kind not in kinds:
        raise ValueError(f"Wrong input corr_type, should be one of {kinds}")
    every, chans = [str(c) for c in all_channels], [str(c) for c in corr_channels]
    for ch in chans:
        if ch not in every:
            raise ValueError(f"Wrong input channel:{ch}, should be one of {every}")
    ref = str(ref_channel).lower()
    if ref not in every:
        raise ValueError(f"Wrong input ref_channel:{ref}, should be one of {every}")
    if verbose:
        print(f"-- loading {kind} correction profile from file")
    if kind == 'bleedthrough':
        name = 'bleedthrough_correction_' + '_'.join(sorted(chans, key=lambda v: -int(v))) + f'_{im_size[-2]}_{im_size[-1]}.npy'
        pf = np.load(os.path.join(correction_folder, name), allow_pickle=True)
        return pf.reshape(len(chans), len(chans), im_size[-2], im_size[-1])
    if kind == 'illumination':
        return {ch: np.load(os.path.join(correction_folder, f'illumination_correction_{ch}_{im_size[-2]}x{im_size[-1]}.npy'), allow_pickle=True)
                for ch in chans}
    pf = {}
    for ch in chans:
        if ch == ref:
            pf[ch] = None
            continue
        base = f'chromatic_correction_{ch}_{ref}' + ''.join(f'_{int(d)}' for d in im_size)
        if kind == 'chromatic':
            pf[ch] = np.load(os.path.join(correction_folder, base + '.npy'), allow_pickle=True)
        else:
            import pickle
            with open(os.path.join(correction_folder, base + '_const.pkl'), 'rb') as fh:
                pf[ch] = pickle.load(fh)
    return pf


def _float_profile(a, what):
    if isinstance(a, _lib.DeviceArray):
        return a
    a = np.asarray(a)
    if a.dtype not in (np.dtype(np.float32), np.dtype(np.float64)):
        raise NotImplementedError(f"{what} of dtype {a.dtype}: the device path takes float32 or float64 profiles")
    return a


# Correction profiles are per-dataset constants (a chromatic profile is 3 x Z x X x Y float64 = 3 GB at the reference's
# stack size): a profile array that comes back on a later call -- the same numpy object, as when a caller loads the
# profiles once and corrects many fields of view -- is uploaded once.  Entries die with their arrays.
_RESIDENT = {}
_RESIDENT_MAX_BYTES = int(os.environ.get("IA3_PROFILE_CACHE_BYTES", str(24 << 30)))


def resident_profile(a):
    """the DeviceArray holding numpy array ``a`` (uploaded on first use, reused while ``a`` is alive and unchanged in
    identity); arrays that are not C-contiguous, or a cache that is full, fall back to per-call copies"""
    if a is None or isinstance(a, _lib.DeviceArray):
        return a
    if not isinstance(a, np.ndarray) or not a.flags.c_contiguous or a.base is not None and not isinstance(a.base, np.ndarray):
        return a
    key = id(a)
    mark = float(a.ravel()[::max(1, a.size // 4096)].sum(dtype=np.float64))      # an array edited in place is uploaded again
    hit = _RESIDENT.get(key)
    if hit is not None and hit[0]() is a and hit[2] == mark:
        return hit[1]
    _RESIDENT.pop(key, None)
    if sum(v[1].nbytes for v in _RESIDENT.values()) + a.nbytes > _RESIDENT_MAX_BYTES:
        return a
    import weakref
    try:
        ref = weakref.ref(a, lambda _r, k=key: _RESIDENT.pop(k, None))
    except TypeError:
        return a
    dev = _lib.DeviceArray(a)
    _RESIDENT[key] = (ref, dev, mark)
    return dev


def correct_image_stacks(ims, load_channels, sel_channels, corr_channels, drift=None,
                         hot_pixel_corr=True, hot_pixel_th=4, z_shift_corr=False,
                         illumination_corr=True, illumination_profile=None,
                         bleed_corr=True, bleed_profile=None,
                         chromatic_ref_channel='647', chromatic_corr=True, chromatic_profile=None,
                         warp=True, return_stacks=False, drift_from=None,
                         gaussian_highpass=False, gauss_sigma=3, gauss_truncate=2):
    """The compute core of correct_fov_image on in-memory channel stacks (reference io_tools/load.py:318-459).

    ims: one (Z, X, Y) uint16 array (or resident _lib.Stack) per channel of load_channels -> the corrected stacks of
    sel_channels, as numpy arrays or, with return_stacks, as resident stacks.  ``drift_from``: called with {channel: stack}
    after the illumination step (where the reference estimates the drift, :383-419); its return value replaces ``drift``."""
    stacks = []
    for im in ims:
        if isinstance(im, _lib.Stack):
            stacks.append(im)
            continue
        if im.dtype != np.uint16:
            raise NotImplementedError(f"channel stacks of dtype {im.dtype}: the device corrections work on uint16 movies")
        stacks.append(_lib.Stack(np.ascontiguousarray(im)))
    if hot_pixel_corr:
        for s in stacks:
            s.remove_hot_pixels(hot_th=hot_pixel_th)
    if z_shift_corr:
        for s in stacks:
            s.z_shift_correct()
    overlap = [ch for ch in corr_channels if ch in sel_channels]
    illum = {}
    if illumination_corr:
        illum = {ch: resident_profile(_float_profile(illumination_profile[ch], "illumination profile")) for ch in load_channels}
    done = set()
    if overlap and bleed_corr:
        bleed_profile = _float_profile(bleed_profile, "bleed-through profile")
        if len(bleed_profile.shape) != 4:
            raise NotImplementedError("per-plane (n, n, Z, X, Y) bleed-through profiles are not supported on the device")
        bleed_dev = resident_profile(bleed_profile)
        bld = [stacks[load_channels.index(ch)] for ch in corr_channels]
        mixed = []
        for i, ch in enumerate(corr_channels):
            fuse = ch in illum and illum[ch].dtype == bleed_profile.dtype
            mixed.append(_lib.Stack.mix(bld, bleed=bleed_dev, bleed_row=i, illum=illum[ch] if fuse else None))
            if fuse:
                done.add(ch)
        for s, ch in zip(mixed, corr_channels):
            stacks[load_channels.index(ch)] = s
    for ch, pf in illum.items():
        if ch not in done:
            s = stacks[load_channels.index(ch)]
            _lib.Stack.mix([s], illum=pf, out=s)
    if drift_from is not None:
        drift = drift_from({ch: stacks[load_channels.index(ch)] for ch in load_channels})
    # (the reference warps with whatever align_image returned -- float64 -- and with float32 for a given drift)
    drift = np.zeros(3, dtype=np.float32) if drift is None else (np.asarray(drift) if drift_from is not None else np.array(drift, dtype=np.float32))
    chroma_channels = [ch for ch in corr_channels if ch in sel_channels and ch != chromatic_ref_channel]
    if warp:
        for ch in sel_channels:
            with_chroma = chromatic_corr and ch in chroma_channels
            if with_chroma or drift.any():
                pf = resident_profile(chromatic_profile[ch]) if with_chroma else None
                k = load_channels.index(ch)
                stacks[k] = stacks[k].warp(drift=drift if drift.any() else None, chroma=pf)
    out = [stacks[load_channels.index(ch)] for ch in sel_channels]
    if gaussian_highpass:
        for s in {id(s): s for s in out}.values():          # (the reference filters every loaded channel; only the selected ones are returned)
            s.gaussian_highpass(gauss_sigma, gauss_truncate)
    return out if return_stacks else [s.fetch() for s in out]


def correct_fov_image(dax_filename, sel_channels,
                      single_im_size=_image_size, all_channels=_allowed_colors,
                      num_buffer_frames=10, num_empty_frames=0,
                      drift=None, calculate_drift=False,
                      drift_channel='488', ref_filename=None,
                      use_autocorr=True, drift_args={},
                      corr_channels=_corr_channels, correction_folder=_correction_folder,
                      warp_image=True,
                      hot_pixel_corr=True, hot_pixel_th=4, z_shift_corr=False,
                      illumination_corr=True, illumination_profile=None,
                      bleed_corr=True, bleed_profile=None,
                      chromatic_ref_channel='647', chromatic_corr=True, chromatic_profile=None,
                      gaussian_highpass=False, gauss_sigma=3, gauss_truncate=2,
                      normalization=False, output_dtype=np.uint16,
                      return_drift=False, verbose=True, force_warp=False, return_stacks=False):
    """Correct one whole field of view (reference io_tools/load.py:160-520); returns (list of corrected stacks,)
    [+ drift, drift_flag with return_drift] like the reference."""
    if not os.path.isfile(dax_filename):
        raise IOError(f"Dax file: {dax_filename} is not a file, exit!")
    if not isinstance(dax_filename, str) or dax_filename[-4:] != '.dax':
        raise IOError(f"Dax file: {dax_filename} has wrong data type, exit!")
    t_total = time.time()
    if verbose:
        print(f"- correct the whole fov for image: {dax_filename}")
    sel_channels = [str(sel_channels)] if isinstance(sel_channels, (str, int)) else [str(ch) for ch in sel_channels]
    single_im_size = np.array(single_im_size, dtype=np.int64)
    all_channels = [str(ch) for ch in all_channels]
    num_buffer_frames, num_empty_frames = int(num_buffer_frames), int(num_empty_frames)
    drift = np.zeros(len(single_im_size), dtype=np.float32) if drift is None else np.array(drift, dtype=np.float32)
    if len(drift) != len(single_im_size):
        raise IndexError("drift should have the same dimension as single_im_size.")
    corr_channels = [str(ch) for ch in sorted(corr_channels, key=lambda v: -int(v)) if str(ch) in all_channels]
    overlap = [ch for ch in corr_channels if ch in sel_channels]
    load_channels = list(corr_channels) if (overlap and bleed_corr) else []
    load_channels += [ch for ch in sel_channels if ch not in load_channels]
    if str(drift_channel) not in all_channels:
        raise ValueError(f"Wrong input of drift_channel:{drift_channel}, should be among {all_channels}")
    # branches of the reference that leave the device path (see the module docstring)
    if calculate_drift and use_autocorr:
        raise NotImplementedError("calculate_drift=True with use_autocorr=True needs skimage's phase_cross_correlation (correction_tools."
                                  "alignment.align_image), not built; use_autocorr=False (bead fitting on the device) or pass drift=")
    if calculate_drift and str(drift_channel) not in load_channels:
        load_channels.append(str(drift_channel))
    if normalization:
        raise NotImplementedError("normalization=True (a float32 pipeline end to end) is not built on the device path")
    if np.dtype(output_dtype) != np.dtype(np.uint16):
        raise NotImplementedError("the device corrections produce uint16 stacks (the reference's default output_dtype)")
    if illumination_corr:
        if illumination_profile is None:
            illumination_profile = load_correction_profile('illumination', corr_channels=load_channels, correction_folder=correction_folder,
                                                           all_channels=all_channels, ref_channel=chromatic_ref_channel,
                                                           im_size=single_im_size, verbose=verbose)
        else:
            if not isinstance(illumination_profile, dict):
                raise TypeError("Wrong input type of illumination_profile, should be dict!")
            for ch in load_channels:
                if ch not in illumination_profile:
                    raise KeyError(f"channel:{ch} not given in illumination_profile")
    if bleed_corr and overlap:
        n = len(corr_channels)
        if bleed_profile is None:
            bleed_profile = load_correction_profile('bleedthrough', corr_channels=corr_channels, correction_folder=correction_folder,
                                                    all_channels=all_channels, ref_channel=chromatic_ref_channel,
                                                    im_size=single_im_size, verbose=verbose)
        else:
            bleed_profile = np.array(bleed_profile, dtype=np.float32)
            if bleed_profile.shape != (n, n, single_im_size[-2], single_im_size[-1]) and bleed_profile.shape != tuple([n, n] + list(single_im_size)):
                raise IndexError(f"Wrong input shape for bleed_profile: {bleed_profile.shape}, should be {(n, n, single_im_size[-2], single_im_size[-1])}")
    if chromatic_corr and overlap:
        if chromatic_profile is None:
            chromatic_profile = load_correction_profile('chromatic' if warp_image else 'chromatic_constants', corr_channels=corr_channels, correction_folder=correction_folder,
                                                        all_channels=all_channels, ref_channel=chromatic_ref_channel,
                                                        im_size=single_im_size, verbose=verbose)
        else:
            if not isinstance(chromatic_profile, dict):
                raise TypeError("Wrong input type of chromatic_profile, should be dict!")
            for ch in load_channels:
                if ch in corr_channels and ch not in chromatic_profile:
                    raise KeyError(f"channel:{ch} not given in chromatic_profile")
    t0 = time.time()
    raw = read_dax(dax_filename)
    _, n_color = get_num_frame(dax_filename, frame_per_color=single_im_size[0], buffer_frame=num_buffer_frames, empty_frame=num_empty_frames)
    ims = split_im_by_channels(raw, load_channels, all_channels[:n_color], single_im_size=single_im_size,
                               num_buffer_frames=num_buffer_frames, num_empty_frames=num_empty_frames, skip_frame0=False)
    del raw
    if verbose:
        print(f"-- loaded image from file:{dax_filename} in {time.time() - t0:.3f}s")
    found = {"drift": drift.copy(), "flag": 0}
    drift_from = None
    if calculate_drift:
        def drift_from(stacks):
            # io_tools/load.py:383-419: the bead channel after hot-pixel / bleed-through / illumination against the reference file
            from ..correction_tools.alignment import align_image
            args = dict(drift_args)
            args.update(all_channels=all_channels, ref_all_channels=all_channels, drift_channel=drift_channel)
            found["drift"], found["flag"] = align_image(
                stacks[str(drift_channel)].fetch(), ref_filename, use_autocorr=use_autocorr,
                correction_args={'single_im_size': single_im_size, 'num_buffer_frames': num_buffer_frames,
                                 'num_empty_frames': num_empty_frames, 'correction_folder': correction_folder},
                verbose=verbose, **args)
            if verbose:
                print(f"--- finish drift: {np.around(found['drift'], 2)}")
            return found["drift"]
    t0 = time.time()
    out = correct_image_stacks(ims, load_channels, sel_channels, corr_channels, drift=drift, drift_from=drift_from,
                               hot_pixel_corr=hot_pixel_corr, hot_pixel_th=hot_pixel_th, z_shift_corr=z_shift_corr,
                               illumination_corr=illumination_corr, illumination_profile=illumination_profile,
                               bleed_corr=bleed_corr, bleed_profile=bleed_profile,
                               chromatic_ref_channel=chromatic_ref_channel, chromatic_corr=chromatic_corr,
                               chromatic_profile=chromatic_profile, warp=bool(warp_image and (verbose or force_warp)), return_stacks=return_stacks,
                               gaussian_highpass=gaussian_highpass, gauss_sigma=gauss_sigma, gauss_truncate=gauss_truncate)
    warp_functions = None
    if not warp_image:
        # io_tools/load.py:461-485: the images stay where they are, the spots are moved instead
        from ..correction_tools.chromatic import generate_chromatic_function
        chroma_channels = [ch for ch in corr_channels if ch in sel_channels and ch != chromatic_ref_channel]
        final_drift = np.asarray(found["drift"])
        warp_functions = []
        for ch in sel_channels:
            if (chromatic_corr and ch in chroma_channels) or final_drift.any():
                warp_functions.append(generate_chromatic_function(chromatic_profile[ch] if (chromatic_corr and ch in chroma_channels) else None, final_drift))
            else:
                warp_functions.append(lambda _spots: _spots)
    if verbose:
        print(f"-- corrected channels {sel_channels} on the device in {time.time() - t0:.3f}s")
        print(f"-- finish correction in {time.time() - t_total:.3f}s")
    ret = [out]
    if not warp_image:
        ret.append(warp_functions)
    if return_drift:
        ret.extend([found["drift"], found["flag"]])
    return tuple(ret)
