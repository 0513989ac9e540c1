"""The normalisation-coefficient and zero-point reductions that follow ``star_photometry`` in lightcurver's workflow
(lightcurver/processes/normalization_calculation.py:133-213 ``calculate_coefficient``,
lightcurver/processes/absolute_zeropoint_calculation.py:95-100), on the device and fed straight from the K2 outputs
(SURVEY.md section 8, row f4).

The reference pulls the fluxes out of SQLite into pandas, pivots them to (star x frame) tables and minimises the "scatter in
each frame" of the star light curves with SLSQP under ``mean(coefficients) = 1`` (:160-187).  That cost is a quadratic form
c^T Q c in the star scaling factors (``lcb_norm_scatter_matrix`` accumulates Q in double precision over all frames), so its
constrained minimum is the solution of one (S+1) x (S+1) KKT system -- no optimiser iterations, no pivot tables.  The per-frame
coefficient, its weighted scatter and the zero points are one thread per frame.

Arrays are (F, S) frame-major -- exactly ``star_photometry_batch``'s ``fluxes`` / ``fluxes_uncertainties`` -- with NaN wherever a
star has no (accepted) measurement in a frame (the reference's chi2 window of :29-48 is applied by ``mask_by_chi2``).
"""
import ctypes as C

import numpy as np

from .. import _lib


def _dev(x):
    import torch
    if _lib.is_torch(x):
        return x.to(device='cuda', dtype=torch.float32).contiguous()
    return torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).to('cuda')


def _stream():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def mask_by_chi2(fluxes, chi2, chi2_min, chi2_max):
    """normalization_calculation.py:44-46 (``sff.chi2 BETWEEN ? AND ?``): measurements outside the window do not exist."""
    import torch
    f = _dev(fluxes).clone()
    c = _dev(chi2)
    f[~((c >= chi2_min) & (c <= chi2_max))] = float('nan')
    return f


def solve_star_scaling(Q):
    """argmin c^T Q c subject to mean(c) = 1 (normalization_calculation.py:182-187) from the KKT system
    [[2Q, 1], [1^T, 0]] [c; lambda] = [0; S].  Stars without any measurement (zero row and column) keep c = 1."""
    Q = np.asarray(Q, np.float64)
    S = Q.shape[0]
    live = np.abs(Q).sum(1) > 0
    c = np.ones(S)
    n = int(live.sum())
    if n >= 2:
        K = np.zeros((n + 1, n + 1))
        K[:n, :n] = 2.0 * Q[np.ix_(live, live)]
        K[:n, n] = 1.0
        K[n, :n] = 1.0
        rhs = np.zeros(n + 1)
        rhs[n] = float(n)
        c[live] = np.linalg.lstsq(K, rhs, rcond=None)[0][:n]
    return c


def calculate_coefficient_arrays(fluxes, fluxes_uncertainties):
    """calculate_coefficient (normalization_calculation.py:157-206) on arrays: fluxes, fluxes_uncertainties (F, S), NaN = missing
    (numpy, or torch CUDA tensors straight from ``star_photometry_batch``).
    Returns dict(coefficient (F,), coefficient_uncertainty (F,), star_scaling (S,), median_flux (S,)) as numpy arrays: the rows
    of the ``normalization_coefficients`` table (:206-211) are (frame_id, hash, coefficient[f], coefficient_uncertainty[f])."""
    import torch
    _lib.require_device()
    flux, dflux = _dev(fluxes), _dev(fluxes_uncertainties)
    F, S = int(flux.shape[0]), int(flux.shape[1])
    if tuple(dflux.shape) != (F, S):
        raise ValueError("fluxes and fluxes_uncertainties must both be (F, S)")
    # a measurement needs both numbers (the pivots of :166-167 share their NaN pattern after the chi2 filter)
    bad = torch.isnan(flux) | torch.isnan(dflux)
    flux = torch.where(bad, torch.full_like(flux, float('nan')), flux)
    dflux = torch.where(bad, torch.full_like(dflux, float('nan')), dflux)
    st = _stream()
    med = torch.empty(S, device='cuda')
    _lib.check(_lib.lib.lcb_norm_medians(flux.data_ptr(), F, S, med.data_ptr(), st), 'lcb_norm_medians')
    Q = torch.empty((S, S), dtype=torch.float64, device='cuda')
    work = torch.empty(int(_lib.lib.lcb_norm_scatter_work_doubles(F, S)), dtype=torch.float64, device='cuda')
    _lib.check(_lib.lib.lcb_norm_scatter_matrix(flux.data_ptr(), dflux.data_ptr(), med.data_ptr(), F, S, Q.data_ptr(),
                                                work.data_ptr(), st), 'lcb_norm_scatter_matrix')
    c = solve_star_scaling(Q.cpu().numpy())                       # S x S doubles: the only host step
    scale = torch.from_numpy(c.astype(np.float32)).to('cuda')
    coef, err = torch.empty(F, device='cuda'), torch.empty(F, device='cuda')
    _lib.check(_lib.lib.lcb_norm_coefficients(flux.data_ptr(), dflux.data_ptr(), med.data_ptr(), scale.data_ptr(), F, S,
                                              coef.data_ptr(), err.data_ptr(), st), 'lcb_norm_coefficients')
    return dict(coefficient=coef.cpu().numpy(), coefficient_uncertainty=err.cpu().numpy(), star_scaling=c,
                median_flux=med.cpu().numpy())


def calculate_zeropoints_arrays(fluxes, catalog_mag):
    """absolute_zeropoint_calculation.py:95-100: per frame, median and standard deviation (pandas ``std``, ddof = 1) over the
    stars of ``catalog_mag - (-2.5 log10 flux)``.  fluxes (F, S) NaN = missing; catalog_mag (S,).
    Returns (zeropoint (F,), zeropoint_uncertainty (F,)) as numpy arrays."""
    import torch
    _lib.require_device()
    flux, cmag = _dev(fluxes), _dev(catalog_mag)
    F, S = int(flux.shape[0]), int(flux.shape[1])
    if tuple(cmag.shape) != (S,):
        raise ValueError("catalog_mag must be (S,)")
    zp, zs = torch.empty(F, device='cuda'), torch.empty(F, device='cuda')
    _lib.check(_lib.lib.lcb_zeropoints(flux.data_ptr(), cmag.data_ptr(), F, S, zp.data_ptr(), zs.data_ptr(), _stream()), 'lcb_zeropoints')
    return zp.cpu().numpy(), zs.cpu().numpy()


NORM_DDL = """CREATE TABLE IF NOT EXISTS normalization_coefficients (
    frame_id INTEGER, combined_footprint_hash INTEGER, coefficient REAL, coefficient_uncertainty REAL,
    PRIMARY KEY (combined_footprint_hash, frame_id))"""


def update_normalization_coefficients(db, norm_data):
    """normalization_calculation.py:53-72: bulk upsert of (frame_id, hash, coefficient, uncertainty)."""
    db.execute(NORM_DDL)
    db.executemany("INSERT INTO normalization_coefficients (frame_id, combined_footprint_hash, coefficient, coefficient_uncertainty) "
                   "VALUES (?, ?, ?, ?) ON CONFLICT(combined_footprint_hash, frame_id) DO UPDATE SET "
                   "coefficient=excluded.coefficient, coefficient_uncertainty=excluded.coefficient_uncertainty", norm_data)
    db.commit()
