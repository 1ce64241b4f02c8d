"""ORACLE (test infrastructure, CPU) -- fit stage of ImageAnalysis3 restated.

Not part of the product: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this.

Restates, in a functional style,
  GaussianFit (v4)          External/Fitting_v4.py:165-396
  GaussianFit (v3)          External/Fitting_v3.py:50-257   (to_center quirk :86, width prior :173-177,:216-221)
  iter_fit_seed_points      External/Fitting_v4.py:559-683 / External/Fitting_v3.py:312-421
  fit_fov_image / get_centers  spot_tools/fitting.py:169-334 (post-filters)
on top of scipy.optimize.leastsq (MINPACK lmder), scipy.spatial.cKDTree and numpy -- the
reference's own third-party layer (scipy 1.18.1 / numpy 2.3.5 in this image; unpinned upstream).
Precision follows what the reference does under numpy >= 2: data and voxel
coordinates are cast to float32, the model is evaluated in float64, the Jacobian is cast to
float32 before MINPACK sees it.

Pinned: yes -- tests/test_oracle_pinned.py compares with the unmodified reference executed under
the import shims (oracle/ref_loader.py) and with the committed fixtures in tests/golden/.
"""
import numpy as np
from scipy.optimize import leastsq
from scipy.spatial import cKDTree
from scipy.spatial.distance import cdist

from . import seed_oracle

SIGMA_ZXY = [1.35, 1.9, 1.9]
_LOGMAX = np.log(np.finfo(np.float64).max)


class _Problem:
    """One GaussianFit problem: data, coordinates, constraints."""

    def __init__(self, vals, X, center, version, delta, min_w, max_w, init_w, weight_sigma):
        self.version = version
        self.wlo, self.whi = min_w * min_w, max_w * max_w
        self.delta = delta
        self.ws = weight_sigma
        self.data = np.array(vals, dtype=np.float32)
        self.X = np.array(X, dtype=np.float32)
        order = np.argsort(vals)
        if center is None:
            center = np.median(np.asarray(X)[:, order][:, -10:], -1)
        self.c_est = center
        srt = np.asarray(vals)[order]
        tiny = np.exp(-10.)
        bk0 = np.log(np.max([np.mean(srt[:10]), tiny]))
        h0 = np.log(np.max([np.mean(srt[-10:]), tiny]))
        if version == 4:
            w2 = init_w ** 2
            g = np.log((self.whi - w2) / (w2 - self.wlo))
            wg = [g, g, g]
            self.prior = None
        else:
            iw = np.array(init_w[:3]).copy()
            for i in range(3):
                v = iw[i]
                if v ** 2 > max_w or v ** 2 < min_w:      # sic: squared value vs unsquared bound
                    iw[i] = 1.5 ** 2
                iw[i] = np.log((self.whi - iw[i] ** 2) / (iw[i] ** 2 - self.wlo))
            wg = list(iw)
            self.prior = iw
        self.p0 = np.array([bk0, h0, 0, 0, 0, wg[0], wg[1], wg[2], 0, 0], dtype=np.float32)

    # ---- constraint maps ------------------------------------------------------------------
    def _sig(self, a, num, off, lo, hi):
        if self.version == 4:
            if a >= _LOGMAX:
                return lo
            if a <= -_LOGMAX:
                return hi
        return num / (1. + np.exp(a)) + off

    def centre(self, a0, a1, a2):
        d, c = self.delta, self.c_est
        if self.version == 4:
            out = []
            for a, ce in zip((a0, a1, a2), c):
                if a >= _LOGMAX:
                    out.append(-d + ce)
                elif a <= -_LOGMAX:
                    out.append(d + ce)
                else:
                    out.append(2. * d / (1. + np.exp(a)) - d + ce)
            return out
        e0, e1, e2 = np.exp(-a0), np.exp(-a1), np.exp(-a2)
        return [2. * d * e0 / (1. + e0) - d + c[0],
                2. * d * e1 / (1. + e1) - d + c[1],
                2. * d * e1 / (1. + e2) - d + c[2]]      # sic (Fitting_v3.py:86)

    def shape_terms(self, q):
        bk, h, a0, a1, a2, w1, w2, w3, pp, tp = q
        t = self._sig(tp, 2., -1., -1, 1)
        p = self._sig(pp, 2., -1., -1, 1)
        dw = self.whi - self.wlo
        v = [self._sig(w, dw, self.wlo, self.wlo, dw + self.wlo) for w in (w1, w2, w3)]
        return t, p, v

    def quad(self, q, X=None):
        """returns (xt, yt, zt, coefficient tuple, t, p, s)"""
        X = self.X if X is None else X
        t, p, v = self.shape_terms(q)
        c = self.centre(q[2], q[3], q[4])
        xt, yt, zt = X[0] - c[0], X[1] - c[1], X[2] - c[2]
        p2, t2 = p * p, t * t
        tc2, pc2 = 1 - t2, 1 - p2
        tc, pc = np.sqrt(tc2), np.sqrt(pc2)
        s1, s2, s3 = 1. / v[0], 1. / v[1], 1. / v[2]
        A = pc2 * tc2 * s1 + t2 * s2 + p2 * tc2 * s3
        B = pc2 * t2 * s1 + tc2 * s2 + p2 * t2 * s3
        Cc = p2 * s1 + pc2 * s3
        D = 2 * tc * t * (pc2 * s1 - s2 + p2 * s3)
        E = 2 * p * pc * tc * (s3 - s1)
        F = 2 * p * pc * t * (s3 - s1)
        return xt, yt, zt, (A, B, Cc, D, E, F), (t, p, tc, pc, t2, p2, tc2, pc2), (s1, s2, s3)

    def gauss(self, q, X=None):
        xt, yt, zt, (A, B, Cc, D, E, F), _, _ = self.quad(q, X)
        arg = A * xt * xt + B * yt * yt + Cc * zt * zt + D * xt * yt + E * xt * zt + F * yt * zt
        return np.exp(q[1] - 0.5 * arg)

    def residual(self, q):
        bk = q[0]
        if self.version == 4:
            bk = np.clip(bk, -709.78, 709.78)
        r = (np.exp(bk) + self.gauss(q)) - self.data
        if self.version == 3 and self.ws > 0:
            r = r + self.ws * np.linalg.norm(self.prior - np.array([q[5], q[6], q[7]]))
        return r

    @staticmethod
    def _nw(w, lo, hi):
        if w > 0:
            e = np.exp(-w)
            return 0.5 * (hi - lo) * e / (hi * e + lo) ** 2
        e = np.exp(w)
        return 0.5 * (hi - lo) * e / (lo * e + hi) ** 2

    def jacobian(self, q):
        bk, h, a0, a1, a2, w1, w2, w3, pp, tp = q
        xt, yt, zt, (A, B, Cc, D, E, F), (t, p, tc, pc, t2, p2, tc2, pc2), (s1, s2, s3) = self.quad(q)
        xx, xy, xz, yy, yz, zz = xt * xt, xt * yt, xt * zt, yt * yt, yt * zt, zt * zt
        arg = A * xx + B * yy + Cc * zz + D * xy + E * xz + F * yz
        g = np.exp(h - 0.5 * arg)
        d, lo, hi = self.delta, self.wlo, self.whi
        cols = [np.exp(bk) + np.zeros(len(g)), g]
        for a, lin in ((a0, 2 * A * xt + D * yt + E * zt), (a1, xt * D + 2 * B * yt + F * zt), (a2, xt * E + yt * F + 2 * Cc * zt)):
            e = np.exp(-np.abs(a))
            cols.append((g * lin) * (-d * e / ((1 + e) * (1 + e))))
        c6 = (g * (-pc2 * tc2 * xx - 2 * pc2 * t * tc * xy - pc2 * t2 * yy + 2 * p * pc * tc * xz + 2 * p * pc * t * yz - p2 * zz)) * self._nw(w1, lo, hi)
        c7 = (g * (-t2 * xx + 2 * t * tc * xy - tc2 * yy)) * self._nw(w2, lo, hi)
        c8 = (g * (-p2 * tc2 * xx - 2 * p2 * t * tc * xy - p2 * t2 * yy - 2 * p * pc * tc * xz - 2 * p * pc * t * yz - pc2 * zz)) * self._nw(w3, lo, hi)
        if self.version == 3:
            pr, ws = self.prior, self.ws
            c6 = c6 + int(pr[0] > w1) * ws - int(pr[0] < w1) * ws
            c7 = c7 + int(pr[1] > w2) * ws - int(pr[1] < w2) * ws
            c8 = c8 + int(pr[2] > w3) * ws - int(pr[2] < w3) * ws
        cols += [c6, c7, c8]
        ep = np.exp(-np.abs(pp) / 2)
        cols.append(g * (s3 - s1) * ((2 * pc2 - 1.) * (tc * xz + t * yz) + p * pc * (tc2 * xx + 2 * t * tc * xy + t2 * yy - zz)) * (ep / (1 + ep * ep)))
        et = np.exp(-np.abs(tp) / 2)
        cols.append(g * ((pc2 * s1 - s2 + p2 * s3) * (t * tc * (yy - xx) - (t2 - tc2) * xy) + p * pc * (s1 - s3) * (t * xz - tc * yz)) * (et / (1 + et * et)))
        return np.array(cols, np.float32).T

    def natural(self, q):
        t, p, v = self.shape_terms(q)
        c = self.centre(q[2], q[3], q[4])
        eps = np.mean(np.abs(self.residual(q)))
        return np.array([np.exp(q[1]), c[0], c[1], c[2], np.exp(q[0]), np.sqrt(v[0]), np.sqrt(v[1]), np.sqrt(v[2]), t, p, eps],
                        dtype=np.float32)

    def solve(self):
        """-> (ok, natural parameters float32 (11), raw parameters float64 (10), nfev, ier)"""
        self.cond = np.inf
        self.last_step_px = self.last_step_rel = 0.0
        if len(self.p0) > len(self.data):
            return False, None, None, 0, 0
        with np.errstate(all='ignore'):
            kw = dict(maxfev=1000) if self.version == 4 else {}
            jac_at = []                                   # lmder evaluates the Jacobian at every accepted iterate

            def jac(x):
                jac_at.append(np.array(x, dtype=np.float64))
                return self.jacobian(x)
            q, _, info, _, ier = leastsq(self.residual, self.p0, Dfun=jac, full_output=True, **kw)
            self.cond = jacobian_condition(self.jacobian(q)) if ier in (1, 2, 3, 4) else np.inf
            # size of MINPACK's last accepted step in natural parameters: where its ftol test fires is
            # uncertain by one iteration, so this is the reference's own stopping uncertainty
            nat, prev = self.natural(q), self.natural(jac_at[-1])
            self.last_step_px = float(np.abs(nat[1:4].astype(np.float64) - prev[1:4]).max())
            cols = [0, 4, 5, 6, 7]
            self.last_step_rel = float((np.abs(nat[cols].astype(np.float64) - prev[cols]) / np.abs(nat[cols].astype(np.float64))).max())
            return True, self.natural(q), q, info['nfev'], ier


# ---- which reference fits are determined by their data ("well posed") --------------------------
# The reference accepts whatever leastsq returns (Fitting_v4.py:388-393).  For junk seeds (pure
# noise, windows clipped to a corner of the stack) that answer is not a function of the data to
# anything like 1e-3 px: the Jacobian at the returned point is close to rank deficient, so a whole
# valley of parameter vectors fits equally well and where MINPACK stops in it depends on the last
# bits of its sums; or MINPACK gives up at maxfev (ier = 5) somewhere along the way.  Such rows are
# excluded from tolerance comparisons (accept/reject is still compared).  The measure: the 2-norm
# condition number of the column-normalised final Jacobian over its NON-ZERO columns (an exactly
# zero column -- a width pinned at its bound, an angle that does not matter for a round spot -- is
# harmless: MINPACK freezes that parameter and every implementation agrees on it).  Real spots,
# also densely overlapping ones, have cond = 4.5 .. 25 (tests/test_oracle_pinned.py::
# test_well_posed_statistics); the cut-off is far above that and far below the 1e6 .. 1e8 of the
# junk seeds.  It is also where a normal-equations solver in FP64 (error ~ cond^2 * 1e-16) is still
# 1e-10 accurate, which is what the CUDA fit uses (csrc/lm_core.h).
COND_WELL_POSED = 1.0e3
# A third kind of ill-posed fit converges, on a well-conditioned Jacobian, but *slowly*: a width pinned
# at its bound (the sigmoid's tail) lets MINPACK crawl for 60 .. 150 evaluations, and the iterate at
# which its ftol test (1.49e-8) first fires moves with the last bits of the data.  Measured on such a
# fit: raising 1 % of the float32 voxel values by ONE ulp changes scipy's own answer by 4e-4 px and
# 4e-4 relative (nfev 136 -> 81).  So every seed whose slowest fit needed more than NFEV_PROBE
# evaluations (real spots: 6 .. 30), or ended with a width on one of its bounds, is probed exactly like
# that, and is ill-posed if the reference moves by more than half the parity tolerance under the
# perturbation.
NFEV_PROBE = 20
# ... and, cheaper and sharper: MINPACK stops at the first iterate whose actual and predicted relative
# reductions are both below ftol; a last-bit difference in those two numbers moves the stop by one
# iteration.  If the reference's own LAST accepted step is larger than the tolerance in natural
# parameters, an implementation that follows the same trajectory to 1e-13 can still land one step away,
# outside the tolerance.  Real spots converge quadratically (last step ~1e-6 px); crawls do not.
PROBE_TOL_PX, PROBE_TOL_REL = 5.0e-4, 5.0e-5
LAST_STEP_PX, LAST_STEP_REL = 1.0e-3, 1.0e-4       # the parity tolerance itself


def _on_bound(nat, min_w, max_w):
    sig = np.asarray(nat[5:8], dtype=np.float64)
    return bool((np.abs(sig - max_w) < 1e-3 * max_w).any() or (np.abs(sig - min_w) < 1e-3 * max_w).any())


def reference_is_unstable(problem, base_nat):
    """re-solve ``problem`` twice with a few float32 voxel values moved by ONE ulp (every 97th voxel up; every 5th
    voxel down); True if the natural parameters move by more than (PROBE_TOL_PX, PROBE_TOL_REL) in either"""
    import copy
    for stride, direction in ((97, np.inf), (5, -np.inf)):
        q = copy.copy(problem)
        d = problem.data.copy()
        idx = np.arange(0, len(d), stride)
        d[idx] = np.nextafter(d[idx], np.float32(direction))
        q.data = d
        ok, nat, _, _, _ = q.solve()
        if not ok:
            return True
        with np.errstate(all='ignore'):
            dc = np.abs(nat[1:4].astype(np.float64) - base_nat[1:4]).max()
            cols = [0, 4, 5, 6, 7]
            rel = (np.abs(nat[cols].astype(np.float64) - base_nat[cols]) / np.abs(base_nat[cols].astype(np.float64))).max()
        if not np.isfinite(dc) or not np.isfinite(rel) or dc > PROBE_TOL_PX or rel > PROBE_TOL_REL:
            return True
    return False


def jacobian_condition(J):
    """2-norm condition number of the column-normalised Jacobian over its non-zero columns."""
    J = np.asarray(J, dtype=np.float64)
    if not np.isfinite(J).all():
        return np.inf
    nrm = np.linalg.norm(J, axis=0)
    keep = nrm > 0
    if not keep.any():
        return np.inf
    sv = np.linalg.svd(J[:, keep] / nrm[keep], compute_uv=False)
    return np.inf if sv[-1] <= 0 else float(sv[0] / sv[-1])


def comparable_mask(centers_nx3, well_posed, radius_fit=5, unstable=None):
    """Seeds whose reference result can be compared at the north-star tolerances.  A grossly ill-posed fit
    (``well_posed`` False: rank-deficient final Jacobian or maxfev; off by 1e-2 .. 1 px between runs) spoils its whole
    window-overlap component (the seeds coupled to it through im_subtr / im_add).  A slow crawl (``unstable``: the
    reference itself moves by 1e-4 .. 1e-2 px under a one-ulp probe) changes the data of the seeds whose windows
    overlap its own by up to a count, i.e. their answer by about the tolerance: it takes those direct neighbours
    with it, but no further (the effect on their neighbours is another three orders smaller).  Measured on the
    full-size dense config (46 182 seeds): 5 rows next to a crawl deviate by 1e-3 .. 1e-2 px, nothing beyond."""
    cen = np.asarray(centers_nx3, dtype=np.float64).reshape(-1, 3)
    ok = np.asarray(well_posed, dtype=bool).copy()
    n = len(cen)
    uns = np.zeros(n, dtype=bool) if unstable is None else np.asarray(unstable, dtype=bool)
    if n == 0 or (ok.all() and not uns.any()):
        return ok & ~uns
    ic = np.trunc(cen).astype(np.int64)
    tree = cKDTree(ic)

    def overlapping(i):
        return [j for j in tree.query_ball_point(ic[i], 2 * radius_fit + 1e-9) if np.abs(ic[j] - ic[i]).max() <= 2 * radius_fit - 1]
    bad = list(np.nonzero(~ok)[0])
    seen = set(bad)
    while bad:
        i = bad.pop()
        for j in overlapping(i):
            if j not in seen:
                seen.add(j)
                bad.append(j)
    ok[list(seen)] = False
    for i in np.nonzero(uns)[0]:
        ok[overlapping(i)] = False
    return ok


def gaussian_fit(vals, X, center=None, version=4, delta_center=3., min_w=0.5, max_w=4., init_w=None, weight_sigma=0):
    if init_w is None:
        init_w = 1.5 if version == 4 else SIGMA_ZXY
    pr = _Problem(vals, X, center, version, delta_center, min_w, max_w, init_w, weight_sigma)
    ok, nat, q, nfev, ier = pr.solve()
    return dict(success=ok, p=nat, p_raw=q, nfev=nfev, ier=ier, problem=pr)


def window(radius):
    g = np.reshape(np.indices([radius * 2] * 3) - radius, [3, -1])
    return g[:, (g * g).sum(0) <= radius ** 2]


def iter_fit(im, centers_3xn, version=4, radius_fit=5, min_delta_center=1., max_delta_center=2.5, n_max_iter=10,
             max_dist_th=0.1, min_w=0.5, max_w=4., init_w=None, weight_sigma=0, do_repeat=True,
             origin=None, full_shape=None, tree=None, global_index=None):
    """firstfit (+ repeatfit) -> dict(ps=list, success, converged, n_iter, dists, first_ps, nfev_first, im_add,
    nfev_max / cond_max = per seed, the largest number of MINPACK function evaluations / Jacobian condition
    number over its fits (inf if one of them ended with ier = 5), well_posed, comparable = see comparable_mask).

    origin / full_shape / tree / global_index (oracle/parallel.py): ``im`` is the crop of a larger stack starting at
    ``origin``; coordinates stay those of the full stack (no arithmetic changes), windows are clipped against
    ``full_shape``, firstfit's nearest-seed rule asks ``tree`` (built over ALL seeds of the stack) and compares with
    ``global_index[i]``.  The seeds given must be closed under "windows can interact"."""
    if init_w is None:
        init_w = 1.5 if version == 4 else SIGMA_ZXY
    cen = np.asarray(centers_3xn).T
    n = len(cen)
    if n == 0:
        if version == 3:
            raise ValueError(f"{n} points have been seeded, exit.")
        raise AttributeError("'iter_fit_seed_points' object has no attribute 'im_subtr'")
    off = window(radius_fit)
    org = np.zeros(3, dtype=np.int64) if origin is None else np.asarray(origin, dtype=np.int64)
    shape = np.array(im.shape if full_shape is None else full_shape)[:, None]
    work = np.array(im, dtype=float)
    if version == 4 and tree is None:
        tree = cKDTree(cen)
    gidx = np.arange(n) if global_index is None else np.asarray(global_index)
    if origin is not None and version != 4:
        raise ValueError("crops are supported for the v4 nearest-seed rule only")
    mk = lambda v, X, c, d: _Problem(v, X, c, version, d, min_w, max_w, init_w, weight_sigma)
    ps, cfit, ok_l, recs, nfevs, conds = [], [], [], [], [], []
    coarse = np.zeros(n, dtype=bool)               # a fit of this seed ended with a last step > the tolerance
    slowest = {}                                   # seed -> (problem, natural parameters) of its longest fit

    def ball(c):
        v = off + np.array([int(c[0]), int(c[1]), int(c[2])])[:, None]
        return v[:, ((v >= 0) & (v < shape)).all(0)]

    def at(a, v):                                  # a[v] with v in full-stack coordinates
        return a[v[0] - org[0], v[1] - org[1], v[2] - org[2]]

    for i, c in enumerate(cen):
        full = ball(c)
        if version == 4:
            _, nn = tree.query(full.T, distance_upper_bound=radius_fit * 2)
            mine = full[:, nn == gidx[i]]
        else:
            near = np.argmin(cdist(full.T, cen), axis=-1)
            me = np.argmin(cdist([c], cen)[0, :])
            mine = full[:, near == me]
        pr = mk(at(im, mine), mine, [c[0], c[1], c[2]], min_delta_center)
        ok, nat, q, nfev, _ = pr.solve()
        ok_l.append(ok)
        nfevs.append(nfev)
        conds.append(pr.cond if ok else 0.0)
        if ok and (nfev > NFEV_PROBE or _on_bound(nat, min_w, max_w)):
            slowest[i] = (pr, nat)
        if ok and (pr.last_step_px > LAST_STEP_PX or pr.last_step_rel > LAST_STEP_REL):
            coarse[i] = True
        if ok:
            rec = pr.gauss(q, full)
            work[full[0] - org[0], full[1] - org[1], full[2] - org[2]] -= rec
            ps.append(nat); cfit.append(nat[1:4]); recs.append(rec)
        else:
            ps.append([np.nan] * 11); cfit.append([np.nan] * 3); recs.append(np.nan)
    out = dict(first_ps=list(ps), first_success=list(ok_l), nfev_first=list(nfevs), im_subtr=work.copy())
    nfev_max = np.array(nfevs, dtype=np.int64)
    cond_max = np.array(conds, dtype=np.float64)
    n_iter, done = 0, np.zeros(n, dtype=bool)
    dists = np.zeros(n) + np.inf
    nfev_rep = []
    stop = not do_repeat
    while not stop:
        ok_old, cf_old = np.array(ok_l), np.array(cfit)
        for i, c in enumerate(cen):
            if done[i]:
                continue
            full = ball(c)
            vals = at(work, full)
            if ok_old[i]:
                vals = recs[i] + vals
            pr = mk(vals, full, [c[0], c[1], c[2]], max_delta_center)
            ok, nat, q, nfev, _ = pr.solve()
            ok_l[i] = ok
            nfev_rep.append(nfev)
            if ok and (nfev > NFEV_PROBE or _on_bound(nat, min_w, max_w)) and (nfev >= nfev_max[i] or i not in slowest):
                slowest[i] = (pr, nat)
            nfev_max[i] = max(nfev_max[i], nfev)
            if ok:
                cond_max[i] = max(cond_max[i], pr.cond)
                if pr.last_step_px > LAST_STEP_PX or pr.last_step_rel > LAST_STEP_REL:
                    coarse[i] = True
            if ok:
                rec = pr.gauss(q)
                ps[i], cfit[i], recs[i] = nat, nat[1:4], rec
                work[full[0] - org[0], full[1] - org[1], full[2] - org[2]] = vals - rec
        both = (np.array(ok_l) & ok_old) > 0
        dists[~both] = 0
        dists[both] = np.sum((cf_old[both] - np.array(cfit)[both]) ** 2, axis=-1)
        done = dists < max_dist_th ** 2
        n_iter += 1
        stop = np.all(done) or (n_iter > n_max_iter)
    well = cond_max <= COND_WELL_POSED       # (a fit that ran into maxfev has cond = inf)
    unstable = coarse.copy()
    for i, (pr, nat) in slowest.items():
        if well[i] and not unstable[i]:
            unstable[i] = reference_is_unstable(pr, nat)
    gross = ~well                                # cond / maxfev: the answer is off by 1e-2 .. 1 px between runs
    well &= ~unstable
    out.update(ps=ps, success=ok_l, converged=done, n_iter=n_iter, dists=dists, im_add=work, nfev_repeat=nfev_rep,
               nfev_max=nfev_max, cond_max=cond_max, well_posed=well, unstable=unstable,
               # grossly irreproducible fits spoil the data of every seed coupled to them; a slow crawl (moves by
               # 1e-4 .. 1e-2 px) takes its direct window neighbours with it (comparable_mask)
               comparable=comparable_mask(cen, ~gross, radius_fit, unstable=unstable))
    return out


def image_background(im, dtype='uint16', bin_size=10, max_iter=10):
    """io_tools/load.py:642-686 find_image_background: mode of the intensity histogram."""
    from scipy.signal import find_peaks
    if dtype is None:
        dtype = im.dtype
    cts, bins = np.histogram(im, bins=np.arange(np.iinfo(dtype).min, np.iinfo(dtype).max, bin_size))
    peaks, height, it = [], np.size(im) / 50, 0
    while len(peaks) == 0:
        height = height / 2
        peaks, props = find_peaks(cts, height=height)
        it += 1
        if it > max_iter:
            break
    if it > max_iter:
        return np.nanmedian(im)
    p = peaks[np.argmax(props['peak_heights'])]
    return (bins[p] + bins[p + 1]) / 2


def local_backgrounds(im, spots, fit_radius=5, **background_args):
    """spot_tools/fitting.py:246-258 + io_tools/crop.py:59-88: background of the crop around every spot."""
    out = []
    shape = np.array(np.shape(im), dtype=np.int32)
    for pt in spots:
        c = np.array(pt[1:4])[:3]
        size = np.ones(3, dtype=np.int32) * (fit_radius * 2)
        lo = np.max([np.round(c - size), np.zeros(3)], axis=0)
        hi = np.min([np.round(c + size + 1), shape], axis=0)
        lim = np.array(np.array([lo, hi]).transpose(), dtype=np.int32)
        out.append(image_background(im[tuple(slice(a, b) for a, b in lim)], **background_args))
    return np.array(out)


def fit_fov_image_oracle(im, th_seed=300, max_num_seeds=500, fit_radius=5, remove_boundary_points=True,
                         seed_backend="scipy", seeds=None, normalize_background=False, normalize_local=False,
                         background_args={}, procs=1, **seed_kw):
    """spot_tools/fitting.py:169-262.  procs > 1: the fits of independent groups of seeds run in a process
    pool (oracle/parallel.py; same results, the reference's loop is sequential only where windows interact)."""
    if seeds is None:
        seeds = seed_oracle.get_seeds_oracle(im, max_num_seeds=max_num_seeds, th_seed=float(th_seed),
                                             backend=seed_backend, **seed_kw)
    if len(seeds) == 0:
        return np.array([]), seeds
    if procs > 1:
        from . import parallel
        res = parallel.iter_fit_parallel(im, seeds.T, procs=procs, radius_fit=fit_radius)
    else:
        res = iter_fit(im, seeds.T, version=4, radius_fit=fit_radius)
    fit_fov_image_oracle.last_result = res
    spots = np.array(res['ps'])
    ok = np.sum(np.isnan(spots), axis=1) == 0
    spots, cmp_ok = spots[ok], res['comparable'][ok]
    if remove_boundary_points:
        inside = (spots[:, 1:4] > np.zeros(3)).all(1) * (spots[:, 1:4] < np.array(im.shape)).all(1)
        spots, cmp_ok = spots[np.where(inside)[0]], cmp_ok[np.where(inside)[0]]
    fit_fov_image_oracle.last_comparable = cmp_ok
    if normalize_background and not normalize_local:
        spots[:, 0] = spots[:, 0] / image_background(im, **background_args)
    elif normalize_local:
        spots[:, 0] = spots[:, 0] / local_backgrounds(im, spots, fit_radius, **background_args)
    return spots, seeds


def fast_fit_big_image_oracle(im, centers_zxy, radius_fit=4, avoid_neigbors=True, recenter=False, better_fit=False):
    """External/Fitting_v4.py:433-447 (gfit_fast, reconstruct=False) + :494-556 (fast_fit_big_image)."""
    ps = []
    centers_zxy = np.asarray(centers_zxy)
    if len(centers_zxy) > 0:
        if avoid_neigbors:
            tree = cKDTree(centers_zxy)
            inters = tree.query_ball_tree(tree, radius_fit * 2)
        g = window(radius_fit)
        zb, xb, yb = g
        X_c = g.T
        sz, sx, sy = im.shape

        def inside(z, x, y):
            k = (z >= 0) & (z < sz) & (x >= 0) & (x < sx) & (y >= 0) & (y < sy)
            return z[k], x[k], y[k]
        for ic, (zc, xc, yc) in enumerate(centers_zxy):
            if avoid_neigbors:
                common = inters[ic]
                rel = centers_zxy[common] - [zc, xc, yc]
                zb_, xb_, yb_ = X_c[np.argmin(cdist(rel, X_c), 0) == common.index(ic)].T
            else:
                zb_, xb_, yb_ = zb, xb, yb
            z, x, y = inside(int(zc) + zb_, int(xc) + xb_, int(yc) + yb_)
            X_ = np.array([z, x, y]).T
            im_ = im[z, x, y]
            if recenter and len(im_) > 0:
                k = np.argmax(im_)
                zc, xc, yc = z[k], x[k], y[k]
                z, x, y = inside(int(zc) + zb_, int(xc) + xb_, int(yc) + yb_)
                X_ = np.array([z, x, y]).T
                im_ = im[z, x, y]
            if not better_fit:
                p = np.array([np.nan] * 12)
                if len(im_) > 0:
                    X = X_.T
                    bk = np.sort(im_)[int(len(im_) * 0.1)]
                    w = im_ - bk
                    w[w < 0] = 0
                    h = np.max(w)
                    w = w / np.sum(w)
                    mu = np.sum(X * w, -1)
                    Xc = X.T - mu
                    cov = np.sum(np.array([[Xc[:, i] * Xc[:, j] for i in range(3)] for j in range(3)]) * w, -1)
                    [[a, d, e], [d, b, f], [e, f, c]] = cov
                    p = np.array([h, mu[0], mu[1], mu[2], bk, a, b, c, d, e, f, np.nan])
            else:
                p = np.array([np.nan] * 11)
                if len(im_) > 0:
                    p = gaussian_fit(im_, X_.T, center=X_[np.argmax(im_)], version=4, delta_center=2.5)["p"]
            ps.append(p)
    return np.array(ps)
