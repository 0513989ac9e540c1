"""Synthetic stamp stacks of the BASELINE.json shapes.

Pure numpy data generator (not a model implementation used by the fits): narrow PSF = elliptical
Moffat + 3 Gaussian blobs of 2-5 % amplitude, stars at sub-pixel offsets, noise law of
lightcurver/processes/cutout_making.py:45  sigma = sqrt(sigma_sky^2 + |d|), 1 % of the stamps get a
3-pixel cosmic flagged in the mask.  Seeds: cfg2 20260102, cfg3 20260103, cfg4 20260104,
cfg5 20260105.
"""
import math

import numpy as np

SEEDS = {'cfg2': 20260102, 'cfg3': 20260103, 'cfg4': 20260104, 'cfg5': 20260105}
_SIG = 2.0 / (2.0 * math.sqrt(2.0 * math.log(2.0)))


def _shift_matrix(c, n, k, mean=True):
    """(..., n, nu) banded matrix: untruncated Gaussian shift + k-decimation (generator only)."""
    nu = n * k
    u = np.arange(nu, dtype=np.float64)
    t = u[:, None] - u[None, :]
    arg = t - np.asarray(c, dtype=np.float64)[..., None, None]
    g = np.exp(-arg * arg / (2 * _SIG * _SIG)) / (math.sqrt(2 * math.pi) * _SIG)
    g[np.abs(arg) > 8.0] = 0.0
    g = g.reshape(*g.shape[:-2], n, k, nu)
    return g.mean(-2) if mean else g.sum(-2)


def true_narrow_psf(rng, F, n, k):
    """(F, nu, nu) unit-sum PSFs and their Moffat parameters."""
    nu = n * k
    fwhm_x = rng.uniform(2.5, 4.5, F)
    fwhm_y = fwhm_x * rng.uniform(0.9, 1.1, F)
    phi = rng.uniform(0, np.pi, F)
    beta = rng.uniform(2.0, 4.0, F)
    ax = np.arange(nu) - (nu - 1) / 2.0
    y, x = np.meshgrid(ax, ax, indexing='ij')
    cp, sp = np.cos(phi)[:, None, None], np.sin(phi)[:, None, None]
    xr = x[None] * cp + y[None] * sp
    yr = -x[None] * sp + y[None] * cp
    fac = 2.0 * np.sqrt(2.0 ** (1.0 / beta) - 1.0)
    rx = (fwhm_x * k / fac)[:, None, None]
    ry = (fwhm_y * k / fac)[:, None, None]
    s = (1.0 + (xr / rx) ** 2 + (yr / ry) ** 2) ** (-beta[:, None, None])
    s /= s.sum((-1, -2), keepdims=True)
    peak = s.max((-1, -2))
    for _ in range(3):
        amp = rng.uniform(0.02, 0.05, F) * peak
        bx = rng.uniform(-1, 1, F) * fwhm_x * k
        by = rng.uniform(-1, 1, F) * fwhm_y * k
        bs = rng.uniform(1.0, 2.0, F) * k
        s += amp[:, None, None] * np.exp(-((x[None] - bx[:, None, None]) ** 2 + (y[None] - by[:, None, None]) ** 2)
                                         / (2 * bs[:, None, None] ** 2))
    s /= s.sum((-1, -2), keepdims=True)
    return s, dict(fwhm_x=fwhm_x, fwhm_y=fwhm_y, phi=phi, beta=beta)


def _render(s, flux, x0, y0, n, k):
    """stamps (F,N,n,n) with sum == flux: flux * k^2 * mean-downsample[s (*) g]."""
    Ay = _shift_matrix(k * y0, n, k)
    Ax = _shift_matrix(k * x0, n, k)
    core = Ay @ s[:, None] @ np.swapaxes(Ax, -1, -2)
    return (flux * k * k)[..., None, None] * core


def _noise_and_masks(rng, clean, sigma_sky):
    sig = np.sqrt(sigma_sky ** 2 + np.abs(clean))
    data = clean + sig * rng.standard_normal(clean.shape)
    noisemap = np.sqrt(sigma_sky ** 2 + np.abs(data))
    masks = np.ones(clean.shape, dtype=bool)
    flat = masks.reshape(-1, *clean.shape[-2:])
    dflat = data.reshape(-1, *clean.shape[-2:])
    n = clean.shape[-1]
    hit = np.nonzero(rng.uniform(size=flat.shape[0]) < 0.01)[0]
    for i in hit:
        yy, xx = rng.integers(1, n - 1, 2)
        for d in range(3):
            xq = min(xx + d, n - 1)
            dflat[i, yy, xq] += 50.0 * float(sig.reshape(-1, n, n)[i, yy, xq])
            flat[i, yy, xq] = False
    return data.astype(np.float32), noisemap.astype(np.float32), masks


def make_psf_frames(F, N, n, k, seed=SEEDS['cfg2'], chunk=256):
    """cfg2/cfg5 PSF-fit input: data, noisemap (F,N,n,n) f32, masks bool, truth."""
    rng = np.random.default_rng(seed)
    out = dict(data=np.empty((F, N, n, n), np.float32), noisemap=np.empty((F, N, n, n), np.float32),
               masks=np.empty((F, N, n, n), bool), psf=np.empty((F, n * k, n * k), np.float32),
               flux=np.empty((F, N)), x0=np.empty((F, N)), y0=np.empty((F, N)), fwhm=np.empty(F))
    for f0 in range(0, F, chunk):
        f1 = min(F, f0 + chunk)
        nf = f1 - f0
        s, mo = true_narrow_psf(rng, nf, n, k)
        x0 = rng.uniform(-0.5, 0.5, (nf, N))
        y0 = rng.uniform(-0.5, 0.5, (nf, N))
        flux = 10.0 ** rng.uniform(3.5, 5.0, (nf, N))
        sky = rng.uniform(5.0, 15.0, (nf, 1, 1, 1))
        clean = _render(s, flux, x0, y0, n, k)
        d, nm, mk = _noise_and_masks(rng, clean, sky)
        out['data'][f0:f1], out['noisemap'][f0:f1], out['masks'][f0:f1] = d, nm, mk
        out['psf'][f0:f1] = s
        out['flux'][f0:f1], out['x0'][f0:f1], out['y0'][f0:f1] = flux, x0, y0
        out['fwhm'][f0:f1] = 0.5 * (mo['fwhm_x'] + mo['fwhm_y'])
    return out


def make_phot_frames(F, S, n, k, seed=SEEDS['cfg3'], chunk=256):
    """cfg3 photometry input: per frame one narrow PSF (the true one) and S star stamps with a
    per-frame transparency c_f ~ LogNormal(0, 0.1) on all fluxes."""
    rng = np.random.default_rng(seed)
    star_flux = 10.0 ** rng.uniform(3.5, 5.0, S)
    out = dict(data=np.empty((F, S, n, n), np.float32), noisemap=np.empty((F, S, n, n), np.float32),
               psf=np.empty((F, n * k, n * k), np.float32), transparency=np.empty(F),
               star_flux=star_flux, x0=np.empty((F, S)), y0=np.empty((F, S)))
    for f0 in range(0, F, chunk):
        f1 = min(F, f0 + chunk)
        nf = f1 - f0
        s, _ = true_narrow_psf(rng, nf, n, k)
        x0 = rng.uniform(-0.5, 0.5, (nf, S))
        y0 = rng.uniform(-0.5, 0.5, (nf, S))
        c = rng.lognormal(0.0, 0.1, nf)
        flux = c[:, None] * star_flux[None]
        sky = rng.uniform(5.0, 15.0, (nf, 1, 1, 1))
        clean = _render(s, flux, x0, y0, n, k)
        sig = np.sqrt(sky ** 2 + np.abs(clean))
        d = clean + sig * rng.standard_normal(clean.shape)
        out['data'][f0:f1] = d
        out['noisemap'][f0:f1] = np.sqrt(sky ** 2 + np.abs(d))
        out['psf'][f0:f1] = s
        out['transparency'][f0:f1] = c
        out['x0'][f0:f1], out['y0'][f0:f1] = x0, y0
    return out


def make_deconv_epochs(E, n, k, M=4, n_psf=32, seed=SEEDS['cfg4']):
    """cfg4 joint-deconvolution input: E epochs of an n x n ROI with M point sources (quad-lens
    geometry), a smooth extended background on the nu x nu grid, narrow PSFs of side k*n_psf."""
    rng = np.random.default_rng(seed)
    nu, P = n * k, n_psf * k
    s, _ = true_narrow_psf(rng, E, n_psf, k)
    quad = np.array([[-3.1, 2.4], [2.8, 3.0], [3.3, -2.2], [-2.6, -2.9]])[:M] * (n / 64.0) * 2.0
    c_x, c_y = quad[:, 0], quad[:, 1]
    # fluxes are pixel sums (block-sum convention, amplitude == flux): 2e4 ... 6e3 e-/s per source
    base_flux = np.array([8e4, 6e4, 4e4, 2.5e4])[:M] / (k * k)
    a = base_flux[None] * rng.lognormal(0, 0.05, (E, M))
    dx = rng.uniform(-1, 1, E)
    dy = rng.uniform(-1, 1, E)
    ax = np.arange(nu) - (nu - 1) / 2.0
    y, x = np.meshgrid(ax, ax, indexing='ij')
    h = 30.0 * np.exp(-(x ** 2 + (y * 1.3) ** 2) / (2 * (3.0 * k) ** 2)) + 8.0 * np.exp(
        -((x - 4 * k) ** 2 + (y + 2 * k) ** 2) / (2 * (6.0 * k) ** 2))
    h /= (k * k) ** 2        # per upsampled pixel: a data pixel sums k^2 of them -> peak surface brightness 30 / k^2 per data pixel
    return dict(psf=s.astype(np.float32), c_x=c_x, c_y=c_y, a=a, dx=dx, dy=dy, h=h, P=P,
                sky=rng.uniform(5.0, 15.0, E), rng=rng)
