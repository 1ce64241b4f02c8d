"""Synthetic 3D FISH stacks (SURVEY.md 8(d) / Appendix E recipe).

uint16, C-contiguous (Z, X, Y): flat background + planted anisotropic Gaussian spots,
Poisson shot noise + Gaussian read noise.  Deterministic for a given (shape, n, seed).
``synth`` is the numpy generator used for parity tests and fixtures; ``synth_torch`` builds
bench-scale stacks on a torch device (same statistical model, different random stream).
"""
import numpy as np

SIGMA_ZXY = (1.35, 1.9, 1.9)


def planted_spots(shape, n, seed, h_range=(600.0, 4000.0), margin=8):
    rng = np.random.default_rng(seed)
    shape = np.asarray(shape)
    lo = np.minimum(margin, (shape - 1) / 2.0)
    centers = rng.uniform(lo, shape - 1 - lo, size=(n, 3))
    heights = rng.uniform(h_range[0], h_range[1], size=n)
    return centers, heights, rng


def _add_spots(signal, centers, heights, sigma):
    shape = signal.shape
    half = [int(4 * s + 1) for s in sigma]
    for (cz, cx, cy), h in zip(centers, heights):
        iz, ix, iy = int(round(cz)), int(round(cx)), int(round(cy))
        z0, z1 = max(0, iz - half[0]), min(shape[0], iz + half[0] + 1)
        x0, x1 = max(0, ix - half[1]), min(shape[1], ix + half[1] + 1)
        y0, y1 = max(0, iy - half[2]), min(shape[2], iy + half[2] + 1)
        gz = np.exp(-0.5 * ((np.arange(z0, z1) - cz) / sigma[0]) ** 2)
        gx = np.exp(-0.5 * ((np.arange(x0, x1) - cx) / sigma[1]) ** 2)
        gy = np.exp(-0.5 * ((np.arange(y0, y1) - cy) / sigma[2]) ** 2)
        signal[z0:z1, x0:x1, y0:y1] += (h * gz[:, None, None] * gx[None, :, None] * gy[None, None, :]).astype(np.float32)


def synth(shape, n, seed, bg=300.0, h_range=(600.0, 4000.0), sigma=SIGMA_ZXY, margin=8,
          read_noise=3.0, return_truth=False):
    """uint16 stack of ``shape`` with ``n`` planted spots."""
    centers, heights, rng = planted_spots(shape, n, seed, h_range, margin)
    signal = np.full(tuple(shape), bg, dtype=np.float32)
    _add_spots(signal, centers, heights, sigma)
    noisy = rng.poisson(signal).astype(np.float32)
    noisy += rng.normal(0.0, read_noise, size=signal.shape).astype(np.float32)
    im = np.clip(np.rint(noisy), 0, 65535).astype(np.uint16)
    if return_truth:
        return im, centers, heights
    return im


def bead_pair(shape, n, drift, seed, bg=300.0, h_range=(1500.0, 4000.0), sigma=SIGMA_ZXY, margin=6, read_noise=3.0):
    """(reference image, drifted image, centres): the same ``n`` beads, in the second image displaced by ``drift`` (z, x, y),
    independent noise -- the input of correction_tools.alignment.align_image"""
    centers, heights, rng = planted_spots(shape, n, seed, h_range, margin)
    out = []
    for shift in (np.zeros(3), np.asarray(drift, dtype=float)):
        signal = np.full(tuple(shape), bg, dtype=np.float32)
        _add_spots(signal, centers + shift, heights, sigma)
        noisy = rng.poisson(signal).astype(np.float32)
        noisy += rng.normal(0.0, read_noise, size=signal.shape).astype(np.float32)
        out.append(np.clip(np.rint(noisy), 0, 65535).astype(np.uint16))
    return out[0], out[1], centers


def synth_torch(shape, n, seed, device, bg=300.0, h_range=(600.0, 4000.0), sigma=SIGMA_ZXY,
                margin=8, read_noise=3.0):
    """Bench-scale generator on a torch device (returns a uint16 torch tensor on ``device``).

    torch has no uint16 arithmetic, so the stack is produced as int32, clipped, and bit-cast via
    int16 -> uint16 view on the host side by the caller (``.cpu().numpy().view(np.uint16)``).
    """
    import torch

    centers, heights, _ = planted_spots(shape, n, seed, h_range, margin)
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    Z, X, Y = (int(s) for s in shape)
    out = torch.empty((Z, X, Y), dtype=torch.int16, device=device)
    # planted signal is sparse: rasterise it on host into a float32 delta volume per z-slab is too
    # slow at 2048^2; instead scatter each spot's box on the device.
    sig = torch.full((Z, X, Y), float(bg), dtype=torch.float32, device=device)
    half = [int(4 * s + 1) for s in sigma]
    cen = torch.as_tensor(centers, dtype=torch.float32, device=device)
    hts = torch.as_tensor(heights, dtype=torch.float32, device=device)
    dz = torch.arange(-half[0], half[0] + 1, device=device)
    dx = torch.arange(-half[1], half[1] + 1, device=device)
    dy = torch.arange(-half[2], half[2] + 1, device=device)
    B = 2048
    for s in range(0, n, B):
        c = cen[s:s + B]
        ic = torch.round(c).to(torch.int64)
        zz = ic[:, 0, None] + dz[None, :]
        xx = ic[:, 1, None] + dx[None, :]
        yy = ic[:, 2, None] + dy[None, :]
        gz = torch.exp(-0.5 * ((zz - c[:, 0, None]) / sigma[0]) ** 2) * ((zz >= 0) & (zz < Z))
        gx = torch.exp(-0.5 * ((xx - c[:, 1, None]) / sigma[1]) ** 2) * ((xx >= 0) & (xx < X))
        gy = torch.exp(-0.5 * ((yy - c[:, 2, None]) / sigma[2]) ** 2) * ((yy >= 0) & (yy < Y))
        val = hts[s:s + B, None, None, None] * gz[:, :, None, None] * gx[:, None, :, None] * gy[:, None, None, :]
        lin = (zz.clamp(0, Z - 1)[:, :, None, None] * X + xx.clamp(0, X - 1)[:, None, :, None]) * Y \
            + yy.clamp(0, Y - 1)[:, None, None, :]
        sig.view(-1).index_add_(0, lin.reshape(-1), val.reshape(-1))
    for z in range(Z):  # slab-wise to bound temporaries
        lam = sig[z]
        noisy = torch.poisson(lam, generator=g) + read_noise * torch.randn(lam.shape, device=device, generator=g)
        q = torch.clamp(torch.round(noisy), 0, 65535).to(torch.int32)
        out[z] = torch.where(q > 32767, q - 65536, q).to(torch.int16)
    return out
