"""CPU restatement of the STARRED forward models, losses and optimisers used by lightcurver.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  PyTorch on CPU, dtype-generic: float64 is the
gradient truth, float32 is the stand-in for "STARRED/JAX run in float32".  Gradients come from
``torch.autograd`` (standing in for ``jax.grad``); nothing here shares code with the CUDA kernels,
whose adjoints are hand-derived.

Reference call sites restated (paths relative to /root/reference):
  * PSF model + loss:            lightcurver/processes/psf_modelling.py:164-171 -> starred build_psf
  * fixed-PSF photometry model:  lightcurver/processes/star_photometry.py:66-128
  * joint deconvolution:         lightcurver/processes/roi_modelling.py:213-335
  * flux uncertainties:          lightcurver/utilities/starred_utilities.py:10-39
Semantics: SURVEY.md Appendix A (A.1 PSF model, A.2 loss, A.3 starlet, A.5 optimiser, A.6/A.7
deconvolution), constants in ``oracle.conventions``.

Coordinate conventions (A.1, A.6; roi_modelling.py:207-210): positions are in DATA pixels with the
origin at the stamp centre (n-1)/2; upsampled grid nu = n*k, centre (nu-1)/2.
"""
import math

import numpy as np
import torch

from .conventions import Conventions, DEFAULT


# ----------------------------------------------------------------------------------------------
# Gaussian "target resolution" kernel (A.1): separable, FWHM = gauss_fwhm_up upsampled pixels
# ----------------------------------------------------------------------------------------------

def gauss_sigma(cv: Conventions = DEFAULT) -> float:
    return cv.gauss_fwhm_up / (2.0 * math.sqrt(2.0 * math.log(2.0)))


def window_centre(c):
    """Integer centre of the G-tap window for a continuous shift c (no gradient)."""
    return torch.floor(c.detach() + 0.5)


def shift_matrix(c, n, k, cv: Conventions = DEFAULT):
    """Banded matrix A (..., n, nu) with  (A @ row)[X] = D_k[ row (*) g(. - c) ][X].

    ``c`` (...,) is the shift in UPSAMPLED pixels.  out[u] = sum_t g(t - c) row[u - t] over the
    G taps t in [ic-G/2+1, ic+G/2], ic = floor(c+0.5); zero outside the array ('same').
    D_k is the block mean (or sum) along this axis.
    """
    G = cv.gauss_taps
    sig = gauss_sigma(cv)
    nu = n * k
    dt = c.dtype
    ic = window_centre(c)
    u = torch.arange(nu, dtype=dt)
    t = u[:, None] - u[None, :]                       # t = u_out - u_in
    tt = t - ic[..., None, None]
    inwin = (tt >= -(G // 2) + 1) & (tt <= G // 2)
    arg = t - c[..., None, None]
    g = torch.exp(-arg * arg / (2.0 * sig * sig)) / (math.sqrt(2.0 * math.pi) * sig)
    g = g * inwin.to(dt)
    g = g.reshape(*g.shape[:-2], n, k, nu)
    return g.mean(-2) if cv.downsample_mean else g.sum(-2)


def moffat_image(fwhm_x, fwhm_y, phi, beta, n, k, dtype=torch.float64):
    """Elliptical Moffat on the nu x nu grid, unit sum (A.1).  FWHMs in DATA pixels."""
    nu = n * k
    fwhm_x, fwhm_y, phi, beta = [torch.as_tensor(v, dtype=dtype) for v in (fwhm_x, fwhm_y, phi, beta)]
    ctr = (nu - 1) / 2.0
    ax = torch.arange(nu, dtype=dtype) - ctr
    y, x = torch.meshgrid(ax, ax, indexing='ij')
    sh = fwhm_x.shape
    x = x.reshape((1,) * len(sh) + x.shape)
    y = y.reshape((1,) * len(sh) + y.shape)
    e = lambda v: v[..., None, None]
    cp, sp = torch.cos(e(phi)), torch.sin(e(phi))
    xr = x * cp + y * sp
    yr = -x * sp + y * cp
    fac = 2.0 * torch.sqrt(torch.pow(torch.as_tensor(2.0, dtype=dtype), 1.0 / e(beta)) - 1.0)
    rx = e(fwhm_x) * k / fac
    ry = e(fwhm_y) * k / fac
    m = torch.pow(1.0 + (xr / rx) ** 2 + (yr / ry) ** 2, -e(beta))
    return m / m.sum((-1, -2), keepdim=True)


def distortion_terms(theta, xy):
    """kwargs_distortion at the stamps' frame positions [R]: theta (..., 6) = dilation_x (2), dilation_y (2), shear (2), each a
    first-order polynomial without constant term in the rescaled position xy (..., N, 2) of utilities/image_coordinates.py:4-25
    (psf_modelling.py:122-124).  Returns ex, ey, sh (..., N)."""
    X, Y = xy[..., 0], xy[..., 1]
    th = theta[..., None, :]
    return th[..., 0] * X + th[..., 1] * Y, th[..., 2] * X + th[..., 3] * Y, th[..., 4] * X + th[..., 5] * Y


def distort_psf(s, theta, xy, cv: Conventions = DEFAULT):
    """apply_distortion (star_photometry.py:303, roi_file_preparation.py:179) [R]: the narrow PSF s (..., nu, nu) seen at the
    positions xy (..., N, 2): s_i[v][u] = det_i * bilinear(s; c0 + A_i (u - c0, v - c0)), A_i = [[1+ex, sh], [sh, 1+ey]],
    c0 = (nu-1)/2, zeros outside the grid, det_i = |A_i| (``distortion_conserve_flux``) or 1.  Returns (..., N, nu, nu)."""
    nu = s.shape[-1]
    dt = s.dtype
    ex, ey, sh = distortion_terms(theta, xy)
    c0 = 0.5 * (nu - 1)
    r = torch.arange(nu, dtype=dt) - c0
    ry, rx = r[:, None], r[None, :]
    e = lambda t: t[..., None, None]
    qx = c0 + (1.0 + e(ex)) * rx + e(sh) * ry
    qy = c0 + e(sh) * rx + (1.0 + e(ey)) * ry
    i0, j0 = torch.floor(qx.detach()), torch.floor(qy.detach())
    fx, fy = qx - i0, qy - j0
    sp = torch.nn.functional.pad(s, (1, 1, 1, 1))                        # one ring of zeros; far-away cells are clamped onto it
    N = xy.shape[-2]
    spf = sp.reshape(*sp.shape[:-2], 1, -1).expand(*sp.shape[:-2], N, -1)

    def tap(jj, ii):
        jc = (jj + 1).clamp(0, nu + 1).long()
        ic = (ii + 1).clamp(0, nu + 1).long()
        inside = ((jj >= -1) & (jj <= nu) & (ii >= -1) & (ii <= nu)).to(dt)
        idx = (jc * (nu + 2) + ic).reshape(*jc.shape[:-2], -1)
        return torch.gather(spf, -1, idx).reshape(jc.shape) * inside
    val = (1 - fy) * ((1 - fx) * tap(j0, i0) + fx * tap(j0, i0 + 1)) + fy * ((1 - fx) * tap(j0 + 1, i0) + fx * tap(j0 + 1, i0 + 1))
    if cv.distortion_conserve_flux:
        val = val * e((1.0 + ex) * (1.0 + ey) - sh * sh)
    return val


def psf_star_models(s, a, x0, y0, n, k, cv: Conventions = DEFAULT, theta=None, xy=None):
    """m_i = a_i D_k[ s_i (*) g(. ; k x0_i, k y0_i) ]   (A.1); s_i = s, or ``distort_psf(s, theta, xy)[i]`` with field distortion.

    s (..., nu, nu); a, x0, y0 (..., N)  ->  (..., N, n, n)
    """
    Ay = shift_matrix(k * y0, n, k, cv)               # (..., N, n, nu)
    Ax = shift_matrix(k * x0, n, k, cv)
    si = s[..., None, :, :] if theta is None else distort_psf(s, theta, xy, cv)
    core = Ay @ si @ Ax.transpose(-1, -2)
    return a[..., None, None] * core


# ----------------------------------------------------------------------------------------------
# Starlet transform (A.3): undecimated B3-spline a-trous, edge replication
# ----------------------------------------------------------------------------------------------

_B3 = (1.0 / 16, 4.0 / 16, 6.0 / 16, 4.0 / 16, 1.0 / 16)


def _atrous_axis(x, j, axis):
    nax = x.shape[axis]
    idx = torch.arange(nax)
    out = 0.0
    for t, hv in enumerate(_B3):
        src = torch.clamp(idx + (t - 2) * (2 ** j), 0, nax - 1)
        out = out + hv * torch.index_select(x, axis, src)
    return out


def starlet_n_scales(nu: int) -> int:
    return int(math.log2(nu))


def starlet(b, n_scales=None):
    """Returns (alpha (..., J, nu, nu), coarse (..., nu, nu)) of the last two axes of b."""
    nu = b.shape[-1]
    J = starlet_n_scales(nu) if n_scales is None else n_scales
    c = b
    planes = []
    for j in range(J):
        cn = _atrous_axis(_atrous_axis(c, j, -1), j, -2)
        planes.append(c - cn)
        c = cn
    return torch.stack(planes, dim=-3), c


def starlet_l1(b, W, lam_scales, lam_hf):
    """lam_hf sum W_0|alpha_0| + lam_scales sum_{j>=1} sum W_j|alpha_j|  (A.2), coarse excluded.

    b (..., nu, nu); W (..., J, nu, nu) or None (== 1).  Returns (...,)
    """
    al, _ = starlet(b)
    if W is not None:
        al = al * W
    ab = al.abs().sum((-1, -2))                       # (..., J)
    return lam_hf * ab[..., 0] + lam_scales * ab[..., 1:].sum(-1)


def starlet_dirac_planes(nu, dtype=torch.float64):
    d = torch.zeros(nu, nu, dtype=dtype)
    d[nu // 2, nu // 2] = 1.0
    al, _ = starlet(d)
    return al


def starlet_noise_levels(var_grad):
    """W (J, nu, nu) = sqrt( var_grad (*) psi_j^2 ): std of the starlet coefficients of a white-ish
    field of per-pixel variance var_grad; psi_j = j-th starlet plane of a centred Dirac ('same'
    convolution, zero padded, anchored at the Dirac position nu//2)."""
    nu = var_grad.shape[-1]
    dtype = var_grad.dtype
    psi2 = starlet_dirac_planes(nu, dtype) ** 2       # (J, nu, nu)
    L = 2 * nu
    F = torch.fft.rfft2(var_grad, s=(L, L))
    K = torch.fft.rfft2(psi2, s=(L, L))
    full = torch.fft.irfft2(F[None] * K, s=(L, L))
    a0 = nu // 2
    out = full[:, a0:a0 + nu, a0:a0 + nu]
    return torch.sqrt(torch.clamp(out, min=0.0))


def psf_noise_weights(weight, a, x0, y0, n, k, cv: Conventions = DEFAULT, dtype=torch.float64):
    """Restates ``propagate_noise(model, noisemap, kwargs, ['starlet'], method='SLIT',
    likelihood_type='chi2')`` for the PSF grid (star_photometry.py:108, roi_modelling.py:299 show
    the call shape; build_psf makes the same call with the stage-1 kwargs) [R].

    For a chi2 likelihood the quantity thresholded by the weighted L1 term is the gradient of the
    chi2 in starlet space, whose noise is J^T C^-1 n.  SLIT's diagonal propagation gives per grid
    pixel  var_grad(p) = sum_i a_i^2 sum_q A_i[q,p]^2 w_i[q]  (A_i = shift+smooth+decimate operator of
    star i, w = mask/sigma^2), and W_j = sqrt(var_grad (*) psi_j^2).  One frame: weight (N,n,n).
    """
    weight = _const(weight, dtype)
    a, x0, y0 = _const(a, dtype), _const(x0, dtype), _const(y0, dtype)
    Ay = shift_matrix(k * y0, n, k, cv)               # (N, n, nu)
    Ax = shift_matrix(k * x0, n, k, cv)
    var = ((Ay * Ay).transpose(-1, -2) @ weight @ (Ax * Ax))      # (N, nu, nu)
    var = (a[:, None, None] ** 2 * var).sum(0)
    return starlet_noise_levels(var)


def psf_noise_weights_mc_limit(weight, a, x0, y0, n, k, cv: Conventions = DEFAULT, dtype=torch.float64):
    """The limit (num_samples -> infinity) of ``propagate_noise(..., method='MC')`` for the PSF grid [R]: draw noise
    maps n_i ~ N(0, sigma_i^2), push them to the grid as the chi2-gradient noise g = sum_i a_i A_i^T (n_i / sigma_i^2),
    take the starlet transform, W_j = std of the coefficients.  Exactly: W_j[p]^2 = sum_q,q' Phi_j[p,q] Cov_g[q,q']
    Phi_j[p,q'] with Cov_g = sum_i a_i^2 A_i^T diag(w_i) A_i -- the SLIT form (psf_noise_weights) keeps only the
    diagonal of Cov_g.  Dense linear algebra: small grids only."""
    weight = _const(weight, dtype)
    a, x0, y0 = _const(a, dtype), _const(x0, dtype), _const(y0, dtype)
    N, nu = weight.shape[0], n * k
    Ay = shift_matrix(k * y0, n, k, cv)               # (N, n, nu)
    Ax = shift_matrix(k * x0, n, k, cv)
    # B maps the N*n*n white unit-variance draws to the grid: g = sum_i a_i Ay_i^T (sqrt(w_i) z_i) Ax_i
    cols = []
    for i in range(N):
        Bi = torch.einsum('yv,xu->vuyx', Ay[i], Ax[i]) * (a[i] * torch.sqrt(weight[i]))[None, None]
        cols.append(Bi.reshape(nu * nu, n * n))
    B = torch.cat(cols, dim=1)                         # (nu^2, N n^2)
    J = starlet_n_scales(nu)
    # starlet of every column of B (linear), then the row norms
    planes = B.T.reshape(-1, nu, nu)
    al, _ = starlet(planes, J)                         # (N n^2, J, nu, nu)
    return torch.sqrt((al ** 2).sum(0))


# ----------------------------------------------------------------------------------------------
# PSF loss (A.2)
# ----------------------------------------------------------------------------------------------

def psf_loss(s_fixed, b, a, x0, y0, data, weight, W, n, k, lam_scales, lam_hf,
             cv: Conventions = DEFAULT, theta=None, xy=None):
    """Per-frame loss.  s_fixed, b (..., nu, nu); a,x0,y0 (..., N); data, weight (..., N, n, n); with field distortion
    theta (..., 6) and the stamps' rescaled frame positions xy (..., N, 2).

    weight = mask / sigma^2.  Returns (...,).
    """
    m = psf_star_models(s_fixed + b, a, x0, y0, n, k, cv, theta, xy)
    chi = (weight * (m - data) ** 2).sum((-1, -2, -3))
    if cv.chi2_half:
        chi = 0.5 * chi
    if lam_scales == 0.0 and lam_hf == 0.0:
        return chi
    return chi + starlet_l1(b, W, lam_scales, lam_hf)


# ----------------------------------------------------------------------------------------------
# Deconvolution model (A.6) and loss (A.7)
# ----------------------------------------------------------------------------------------------

def gauss_window(pc, nu, cv: Conventions = DEFAULT):
    """(..., nu): g(u - pc) for u inside the G-tap window around floor(pc+0.5), else 0."""
    G = cv.gauss_taps
    sig = gauss_sigma(cv)
    u = torch.arange(nu, dtype=pc.dtype)
    ic = window_centre(pc)
    tt = u - ic[..., None]
    inwin = (tt >= -(G // 2) + 1) & (tt <= G // 2)
    arg = u - pc[..., None]
    g = torch.exp(-arg * arg / (2.0 * sig * sig)) / (math.sqrt(2.0 * math.pi) * sig)
    return g * inwin.to(pc.dtype)


def bilinear_warp(h, dx, dy, alpha, k, ctr):
    """Scene h moved by rotation alpha (about ctr) then translation k*(dx,dy); zeros outside.

    h (nu,nu); dx,dy,alpha (E,) -> (E,nu,nu).  out(p) = h( R_alpha^-1 (p - ctr - k d) + ctr ).
    """
    nu = h.shape[-1]
    dt = h.dtype
    ax = torch.arange(nu, dtype=dt)
    v, u = torch.meshgrid(ax, ax, indexing='ij')
    e = lambda t: t[:, None, None]
    pu = u[None] - ctr - k * e(dx)
    pv = v[None] - ctr - k * e(dy)
    ca, sa = torch.cos(e(alpha)), torch.sin(e(alpha))
    qu = ca * pu + sa * pv + ctr
    qv = -sa * pu + ca * pv + ctr
    u0 = torch.floor(qu.detach())
    v0 = torch.floor(qv.detach())
    fu = qu - u0
    fv = qv - v0
    u0 = u0.long()
    v0 = v0.long()
    hf = h.reshape(-1)

    def tap(vv, uu):
        ok = (vv >= 0) & (vv < nu) & (uu >= 0) & (uu < nu)
        idx = torch.clamp(vv, 0, nu - 1) * nu + torch.clamp(uu, 0, nu - 1)
        return hf[idx] * ok.to(dt)

    return ((1 - fv) * ((1 - fu) * tap(v0, u0) + fu * tap(v0, u0 + 1))
            + fv * ((1 - fu) * tap(v0 + 1, u0) + fu * tap(v0 + 1, u0 + 1)))


def conv_same(f, s):
    """out[v,u] = sum_j s[jv,ju] f[v + j0 - jv, u + j0 - ju],  j0 = (P-1)//2  (scipy 'same').

    f (E,nu,nu), s (E,P,P).  FFT based (STARRED's default route [R]); exact up to rounding.
    """
    nu = f.shape[-1]
    P = s.shape[-1]
    L = nu + P - 1
    F = torch.fft.rfft2(f, s=(L, L))
    S = torch.fft.rfft2(s, s=(L, L))
    full = torch.fft.irfft2(F * S, s=(L, L))
    j0 = (P - 1) // 2
    return full[..., j0:j0 + nu, j0:j0 + nu]


def conv_same_direct(f, s):
    """Same as conv_same by direct summation (small sizes only; used to validate the FFT route)."""
    nu = f.shape[-1]
    P = s.shape[-1]
    j0 = (P - 1) // 2
    E = f.shape[0]
    fp = torch.nn.functional.pad(f, (P - 1 - j0, j0, P - 1 - j0, j0))
    # cross-correlation with flipped kernel == convolution
    w = torch.flip(s, (-1, -2))[:, None]
    return torch.nn.functional.conv2d(fp[None], w, groups=E)[0]


def downsample(x, k, cv: Conventions = DEFAULT):
    if k == 1:
        return x
    n = x.shape[-1] // k
    y = x.reshape(*x.shape[:-2], n, k, n, k)
    return y.mean((-1, -3)) if cv.downsample_mean else y.sum((-1, -3))


def deconv_positions(c_x, c_y, dx, dy, alpha, k, nu, P):
    """Point-source centres on the f grid, (E,M) each.  delta compensates the even-P 'same' anchor."""
    j0 = (P - 1) // 2
    delta = (P - 1) / 2.0 - j0
    ctr_f = (nu - 1) / 2.0 - delta
    ca, sa = torch.cos(alpha)[:, None], torch.sin(alpha)[:, None]
    px = ca * c_x[None] - sa * c_y[None] + dx[:, None]
    py = sa * c_x[None] + ca * c_y[None] + dy[:, None]
    return ctr_f + k * px, ctr_f + k * py, ctr_f


def deconv_highres(h, a, c_x, c_y, dx, dy, alpha, n, k, P, cv: Conventions = DEFAULT, with_h=True):
    """f_e = Warp_e[h] + sum_m a_em g(. ; centres)   (E,nu,nu)."""
    nu = n * k
    uc, vc, ctr_f = deconv_positions(c_x, c_y, dx, dy, alpha, k, nu, P)
    gx = gauss_window(uc, nu, cv)
    gy = gauss_window(vc, nu, cv)
    f = torch.einsum('em,emv,emu->evu', a, gy, gx)
    if with_h:
        f = f + bilinear_warp(h, dx, dy, alpha, k, ctr_f)
    return f


def deconv_model(h, mean, a, c_x, c_y, dx, dy, alpha, psf, n, k, cv: Conventions = DEFAULT,
                 with_h=True, direct=False):
    """m_e = D_k[ s_e (*) f_e ] + mean_e   (A.6).  a (E,M); returns (E,n,n)."""
    P = psf.shape[-1]
    f = deconv_highres(h, a, c_x, c_y, dx, dy, alpha, n, k, P, cv, with_h)
    conv = conv_same_direct(f, psf) if direct else conv_same(f, psf)
    return downsample(conv, k, cv) + mean[:, None, None]


def pts_source_l1(a, c_x, c_y, dx, dy, alpha, n, k, P, W, lam_pts, cv: Conventions = DEFAULT, first_epoch_is_local=True):
    """regularization_strength_pts_source (roi_modelling.py:311; Millon et al. 2024) [R, low confidence]: weighted L1 of
    the FIRST starlet scale of the point-source channel p_e = sum_m a_em g(. ; centres), weights W[0] (1 if W is None),
    summed over the epochs (cv.pts_source_all_epochs) or taken on the first epoch only."""
    p = deconv_highres(None, a, c_x, c_y, dx, dy, alpha, n, k, P, cv, with_h=False)
    c1 = _atrous_axis(_atrous_axis(p, 0, -1), 0, -2)
    al0 = (p - c1).abs()
    if W is not None:
        al0 = al0 * W[0]
    per = al0.sum((-1, -2))
    if cv.pts_source_all_epochs:
        return lam_pts * per.sum()
    return lam_pts * per[0] if first_epoch_is_local else 0.0 * per[0]


def flux_uniformity(a, lam_fu, cv: Conventions = DEFAULT):
    """regularization_strength_flux_uniformity (roi_modelling.py:275-276, 312) [R, low confidence]: scatter of the fluxes
    of every source over the epochs, a (E,M): lam * sum_m std_e(a_em) (population std) [/ |mean_e(a_em)|]."""
    mean = a.mean(0)
    var = ((a - mean) ** 2).mean(0)
    sd = torch.sqrt(torch.where(var > 0, var, torch.ones_like(var))) * (var > 0).to(a.dtype)
    if cv.flux_uniformity_relative:
        sd = sd / mean.abs()
    return lam_fu * sd.sum()


def deconv_loss(h, mean, a, c_x, c_y, dx, dy, alpha, psf, data, weight, W, n, k,
                lam_scales=0.0, lam_hf=0.0, lam_pos=0.0, prior=None,
                cv: Conventions = DEFAULT, with_h=True, direct=False, per_epoch=False,
                lam_pts=0.0, lam_fu=0.0):
    """A.7: chi2 + starlet-L1(h) + positivity(h) + Gaussian prior on (c_x, c_y) + pts-source L1 + flux uniformity.

    prior = (mu_x, sig_x, mu_y, sig_y) or None.  per_epoch=True returns the (E,) chi2 terms only.
    """
    m = deconv_model(h, mean, a, c_x, c_y, dx, dy, alpha, psf, n, k, cv, with_h, direct)
    chi = (weight * (m - data) ** 2).sum((-1, -2))
    if cv.chi2_half:
        chi = 0.5 * chi
    if per_epoch:
        return chi
    L = chi.sum()
    if with_h and (lam_scales != 0.0 or lam_hf != 0.0):
        L = L + starlet_l1(h, W, lam_scales, lam_hf)
    if with_h and lam_pos != 0.0:
        L = L - lam_pos * torch.clamp(h, max=0.0).sum()
    if prior is not None:
        mux, sgx, muy, sgy = prior
        L = L + 0.5 * (((c_x - mux) / sgx) ** 2).sum() + 0.5 * (((c_y - muy) / sgy) ** 2).sum()
    if lam_pts != 0.0:
        L = L + pts_source_l1(a, c_x, c_y, dx, dy, alpha, n, k, psf.shape[-1], W, lam_pts, cv)
    if lam_fu != 0.0:
        L = L + flux_uniformity(a, lam_fu, cv)
    return L


def deconv_noise_weights(psf, weight, dx, dy, alpha, n, k, cv: Conventions = DEFAULT, dtype=torch.float64):
    """W (J,nu,nu) for the shared background h: var(p) = sum_e sum_q (d m_e[q] / d h[p])^2 w_e[q] with the
    Jacobian taken explicitly by autograd (small sizes only), then starlet_noise_levels.  Same
    definition as psf_noise_weights (SLIT diagonal propagation of the chi2-gradient noise)."""
    psf, weight = _const(psf, dtype), _const(weight, dtype)
    dx, dy, alpha = _const(dx, dtype), _const(dy, dtype), _const(alpha, dtype)
    E, nu = psf.shape[0], n * k
    zM = torch.zeros(1, dtype=dtype)

    def model_of_h(hflat):
        return deconv_model(hflat.reshape(nu, nu), torch.zeros(E, dtype=dtype), torch.zeros(E, 1, dtype=dtype), zM, zM,
                            dx, dy, alpha, psf, n, k, cv, direct=True)

    Jac = torch.autograd.functional.jacobian(model_of_h, torch.zeros(nu * nu, dtype=dtype))   # (E,n,n,nu^2)
    var = (Jac ** 2 * weight[..., None]).sum((0, 1, 2)).reshape(nu, nu)
    return starlet_noise_levels(var)


def phot_models(psf, a, dx, dy, n, k, cv: Conventions = DEFAULT):
    """Fixed-PSF single point source at c=0, h=0, mean=0 (star_photometry.py:52-87), P == nu.

    Identical to the deconvolution model restricted to one source (tests check this); written in
    banded form so that B independent (frame,star) items batch.  psf (B,nu,nu); a,dx,dy (B,).
    """
    Ay = shift_matrix(k * dy, n, k, cv)
    Ax = shift_matrix(k * dx, n, k, cv)
    return a[:, None, None] * (Ay @ psf @ Ax.transpose(-1, -2))


def phot_loss(psf, a, dx, dy, data, weight, n, k, cv: Conventions = DEFAULT):
    m = phot_models(psf, a, dx, dy, n, k, cv)
    chi = (weight * (m - data) ** 2).sum((-1, -2))
    return 0.5 * chi if cv.chi2_half else chi


def flux_sigma(psf, dx, dy, weight, n, k, cv: Conventions = DEFAULT):
    """starred_utilities.py:10-39 in closed form: sigma_a = (sum w (dm/da)^2)^-1/2."""
    one = torch.ones_like(dx)
    dm = phot_models(psf, one, dx, dy, n, k, cv)
    H = (weight * dm * dm).sum((-1, -2))
    if not cv.chi2_half:
        H = 2.0 * H
    return H.rsqrt()


# ----------------------------------------------------------------------------------------------
# Optimiser (A.5): optax-equivalent AdaBelief, optional clip_by_global_norm + exponential schedule
# ----------------------------------------------------------------------------------------------

class AdaBelief:
    """``groups``: list of tensors updated together; ``problem_dims``: number of leading dims that
    index independent problems (per-problem global-norm clip)."""

    def __init__(self, params, lr, n_iter, schedule, cv: Conventions = DEFAULT, problem_dims=0):
        self.params = params
        self.mu = [torch.zeros_like(p) for p in params]
        self.nu = [torch.zeros_like(p) for p in params]
        self.t = 0
        self.lr0, self.T, self.schedule, self.cv, self.pd = lr, n_iter, schedule, cv, problem_dims

    def _norms(self, grads):
        tot = 0.0
        for g in grads:
            red = tuple(range(self.pd, g.dim()))
            tot = tot + ((g * g).sum(red) if red else g * g)
        return torch.sqrt(tot)

    def step(self, grads):
        cv = self.cv
        if self.schedule:
            gn = self._norms(grads)
            scale = torch.where(gn < cv.clip_global_norm, torch.ones_like(gn), cv.clip_global_norm / gn)
            grads = [g * scale.reshape(scale.shape + (1,) * (g.dim() - scale.dim())) for g in grads]
            lr = self.lr0 * cv.lr_decay_rate ** (self.t / self.T)
        else:
            lr = self.lr0
        self.t += 1
        bc1 = 1.0 - cv.belief_b1 ** self.t
        bc2 = 1.0 - cv.belief_b2 ** self.t
        with torch.no_grad():
            for p, g, mu, nu in zip(self.params, grads, self.mu, self.nu):
                mu.mul_(cv.belief_b1).add_(g, alpha=1.0 - cv.belief_b1)
                d = g - mu
                nu.mul_(cv.belief_b2).add_(d * d, alpha=1.0 - cv.belief_b2).add_(cv.belief_eps_root)
                p.sub_(lr * (mu / bc1) / (torch.sqrt(nu / bc2) + cv.belief_eps))


def _leaf(x, dtype):
    return torch.as_tensor(np.asarray(x), dtype=dtype).clone().requires_grad_(True)


def _const(x, dtype):
    return torch.as_tensor(np.asarray(x), dtype=dtype)


# ----------------------------------------------------------------------------------------------
# Fit drivers
# ----------------------------------------------------------------------------------------------

def fit_phot(psf, data, weight, a0, n, k, n_iter, lr=1e-3, schedule=True,
             cv: Conventions = DEFAULT, dtype=torch.float32, dx0=None, dy0=None):
    """B independent amplitude+shift fits (north-star formulation of star_photometry.py:113-122:
    c fixed at 0, per-item clip).  Returns dict of numpy arrays incl. loss_hist (B,T)."""
    psf, data, weight = _const(psf, dtype), _const(data, dtype), _const(weight, dtype)
    B = data.shape[0]
    a = _leaf(a0, dtype)
    dx = _leaf(np.zeros(B) if dx0 is None else dx0, dtype)
    dy = _leaf(np.zeros(B) if dy0 is None else dy0, dtype)
    opt = AdaBelief([a, dx, dy], lr, n_iter, schedule, cv, problem_dims=1)
    hist = np.zeros((B, n_iter), dtype=np.float64)
    for it in range(n_iter):
        L = phot_loss(psf, a, dx, dy, data, weight, n, k, cv)
        g = torch.autograd.grad(L.sum(), [a, dx, dy])
        hist[:, it] = L.detach().double().numpy()
        opt.step(list(g))
    with torch.no_grad():
        m = phot_models(psf, a, dx, dy, n, k, cv)
        res = data - m
        chi2 = (weight * res * res).sum((-1, -2)) / (n * n)
        sig = flux_sigma(psf, dx, dy, weight, n, k, cv)
    return dict(a=a.detach().numpy(), dx=dx.detach().numpy(), dy=dy.detach().numpy(),
                sigma_a=sig.numpy(), chi2=chi2.numpy(), residuals=res.numpy(), loss_hist=hist)


def phot_loss_grad(psf, data, weight, a, dx, dy, n, k, cv: Conventions = DEFAULT, dtype=torch.float64):
    psf, data, weight = _const(psf, dtype), _const(data, dtype), _const(weight, dtype)
    a, dx, dy = _leaf(a, dtype), _leaf(dx, dtype), _leaf(dy, dtype)
    L = phot_loss(psf, a, dx, dy, data, weight, n, k, cv)
    g = torch.autograd.grad(L.sum(), [a, dx, dy])
    return L.detach().numpy(), [t.numpy() for t in g]


def psf_loss_grad(s_fixed, b, a, x0, y0, data, weight, W, n, k, lam_scales, lam_hf,
                  cv: Conventions = DEFAULT, dtype=torch.float64, theta=None, xy=None):
    """Loss and gradient wrt (b, a, x0, y0 [, theta]) for ONE frame or a batch with uniform N."""
    s_fixed, data, weight = _const(s_fixed, dtype), _const(data, dtype), _const(weight, dtype)
    W = None if W is None else _const(W, dtype)
    b, a, x0, y0 = _leaf(b, dtype), _leaf(a, dtype), _leaf(x0, dtype), _leaf(y0, dtype)
    leaves = [b, a, x0, y0]
    if theta is not None:
        theta, xy = _leaf(theta, dtype), _const(xy, dtype)
        leaves.append(theta)
    L = psf_loss(s_fixed, b, a, x0, y0, data, weight, W, n, k, lam_scales, lam_hf, cv, theta, xy)
    g = torch.autograd.grad(L.sum(), leaves)
    return L.detach().numpy(), [t.numpy() for t in g]


def fit_psf_stage2(s_fixed, b0, a0, x00, y00, data, weight, W, n, k, n_iter, lr=None,
                   lam_scales=None, lam_hf=None, cv: Conventions = DEFAULT, dtype=torch.float32, theta0=None, xy=None):
    """AdaBelief on {background grid, a, x0, y0}, Moffat fixed (A.4 stage 2).  Leading batch dim
    (frames, uniform N) optional; clip is per frame."""
    lr = cv.psf_stage2_lr if lr is None else lr
    lam_scales = cv.psf_lambda_scales if lam_scales is None else lam_scales
    lam_hf = cv.psf_lambda_hf if lam_hf is None else lam_hf
    s_fixed, data, weight = _const(s_fixed, dtype), _const(data, dtype), _const(weight, dtype)
    W = None if W is None else _const(W, dtype)
    b, a, x0, y0 = _leaf(b0, dtype), _leaf(a0, dtype), _leaf(x00, dtype), _leaf(y00, dtype)
    pd = a.dim() - 1
    leaves = [b, a, x0, y0]
    theta = None
    if theta0 is not None:                                # field distortion: six more free parameters per frame
        theta, xy = _leaf(theta0, dtype), _const(xy, dtype)
        leaves.append(theta)
    opt = AdaBelief(leaves, lr, n_iter, True, cv, problem_dims=pd)
    hist = []
    for it in range(n_iter):
        L = psf_loss(s_fixed, b, a, x0, y0, data, weight, W, n, k, lam_scales, lam_hf, cv, theta, xy)
        g = torch.autograd.grad(L.sum(), leaves)
        hist.append(L.detach().double().numpy().copy())
        opt.step(list(g))
    out = dict(b=b.detach().numpy(), a=a.detach().numpy(), x0=x0.detach().numpy(),
               y0=y0.detach().numpy(), loss_hist=np.stack(hist, -1))
    if theta is not None:
        out['theta'] = theta.detach().numpy()
    return out


def psf_products(s, a, x0, y0, data, weight, n, k, cv: Conventions = DEFAULT, dtype=torch.float64):
    """narrow_psf, full_psf, residuals (data - model), reduced chi2 of one frame (A.1, A.4)."""
    s, data, weight = _const(s, dtype), _const(data, dtype), _const(weight, dtype)
    a, x0, y0 = _const(a, dtype), _const(x0, dtype), _const(y0, dtype)
    m = psf_star_models(s, a, x0, y0, n, k, cv)
    res = data - m
    npix = (weight > 0).sum().clamp(min=1)
    chi2 = (weight * res * res).sum() / npix
    nu = n * k
    cvs = Conventions(**{**cv.as_dict(), 'downsample_mean': False})
    z = torch.zeros(1, dtype=dtype)
    G0 = shift_matrix(z, nu, 1, cvs)[0]               # (nu,nu) unshifted Gaussian, no downsample
    full = G0 @ s @ G0.T
    return dict(narrow_psf=(s / s.sum()).numpy(), full_psf=(full / full.sum()).numpy(),
                residuals=res.numpy(), chi2=float(chi2))


def fit_psf_stage1(data, weight, n, k, fwhm_guess, a0, n_iter, cv: Conventions = DEFAULT, strict_tol=True):
    """Analytic stage (A.4 stage 1): scipy L-BFGS-B over {fwhm_x, fwhm_y, phi, beta, a, x0, y0},
    background 0, lambda 0, C fixed at 1.  float64.  One frame."""
    from scipy.optimize import minimize
    dtype = torch.float64
    data, weight = _const(data, dtype), _const(weight, dtype)
    N = data.shape[0]
    x_init = np.concatenate([[fwhm_guess, fwhm_guess, 0.0, cv.moffat_beta_init],
                             np.asarray(a0, dtype=np.float64), np.zeros(2 * N)])
    lo = [cv.moffat_fwhm_min, cv.moffat_fwhm_min, -np.inf, cv.moffat_beta_min] + [0.0] * N + [-n / 4.0] * (2 * N)
    hi = [n / 2.0, n / 2.0, np.inf, cv.moffat_beta_max] + [np.inf] * N + [n / 4.0] * (2 * N)

    def unpack(x):
        return x[0], x[1], x[2], x[3], x[4:4 + N], x[4 + N:4 + 2 * N], x[4 + 2 * N:]

    last = {}

    def fun(xv):
        x = torch.tensor(xv, dtype=dtype, requires_grad=True)
        fx, fy, ph, be, a, x0, y0 = unpack(x)
        s = moffat_image(fx, fy, ph, be, n, k, dtype)
        m = psf_star_models(s, a, x0, y0, n, k, cv)
        L = (weight * (m - data) ** 2).sum()
        if cv.chi2_half:
            L = 0.5 * L
        L.backward()
        last['L'] = float(L.detach())
        return last['L'], x.grad.numpy().copy()

    hist = []
    res = minimize(fun, x_init, jac=True, method='L-BFGS-B', bounds=list(zip(lo, hi)),
                   options=({'maxiter': n_iter, 'maxfun': 20 * n_iter, 'ftol': 1e-15, 'gtol': 1e-10} if strict_tol else
                            {'maxiter': n_iter, 'maxfun': 20 * n_iter}),      # strict: convergence tests; else scipy's defaults
                   callback=lambda xk: hist.append(last['L']))
    fx, fy, ph, be, a, x0, y0 = unpack(res.x)
    return dict(fwhm_x=fx, fwhm_y=fy, phi=ph, beta=be, C=1.0, a=a, x0=x0, y0=y0,
                loss=float(res.fun), loss_hist=np.array(hist))


def deconv_loss_grad(params, fixed, psf, data, weight, W, n, k, reg, cv: Conventions = DEFAULT,
                     dtype=torch.float64):
    """params/fixed: dicts over {h, mean, a, c_x, c_y, dx, dy, alpha}; returns loss and grads of params."""
    leaves = {kk: _leaf(v, dtype) for kk, v in params.items()}
    allp = {**{kk: _const(v, dtype) for kk, v in fixed.items()}, **leaves}
    prior = reg.get('prior')
    if prior is not None:
        prior = tuple(_const(p, dtype) for p in prior)
    L = deconv_loss(allp['h'].reshape(n * k, n * k), allp['mean'], allp['a'], allp['c_x'], allp['c_y'],
                    allp['dx'], allp['dy'], allp['alpha'], _const(psf, dtype), _const(data, dtype),
                    _const(weight, dtype), None if W is None else _const(W, dtype), n, k,
                    reg.get('lam_scales', 0.0), reg.get('lam_hf', 0.0), reg.get('lam_pos', 0.0),
                    prior, cv, lam_pts=reg.get('lam_pts', 0.0), lam_fu=reg.get('lam_fu', 0.0))
    names = list(leaves)
    g = torch.autograd.grad(L, [leaves[kk] for kk in names])
    return float(L.detach()), {kk: t.numpy() for kk, t in zip(names, g)}


def fit_deconv(params, fixed, psf, data, weight, W, n, k, reg, n_iter, lr=1e-4, schedule=False,
               cv: Conventions = DEFAULT, dtype=torch.float32):
    """Stage 2 of roi_modelling.py:326-335: AdaBelief over the free params, no clip/schedule."""
    leaves = {kk: _leaf(v, dtype) for kk, v in params.items()}
    consts = {kk: _const(v, dtype) for kk, v in fixed.items()}
    psf, data, weight = _const(psf, dtype), _const(data, dtype), _const(weight, dtype)
    W = None if W is None else _const(W, dtype)
    prior = reg.get('prior')
    if prior is not None:
        prior = tuple(_const(p, dtype) for p in prior)
    names = list(leaves)
    opt = AdaBelief([leaves[kk] for kk in names], lr, n_iter, schedule, cv, problem_dims=0)
    hist = np.zeros(n_iter)
    for it in range(n_iter):
        p = {**consts, **leaves}
        L = deconv_loss(p['h'].reshape(n * k, n * k), p['mean'], p['a'], p['c_x'], p['c_y'], p['dx'],
                        p['dy'], p['alpha'], psf, data, weight, W, n, k, reg.get('lam_scales', 0.0),
                        reg.get('lam_hf', 0.0), reg.get('lam_pos', 0.0), prior, cv,
                        lam_pts=reg.get('lam_pts', 0.0), lam_fu=reg.get('lam_fu', 0.0))
        g = torch.autograd.grad(L, [leaves[kk] for kk in names])
        hist[it] = float(L.detach())
        opt.step(list(g))
    out = {kk: v.detach().numpy() for kk, v in leaves.items()}
    out['loss_hist'] = hist
    return out
