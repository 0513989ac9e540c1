"""The normalisation / zero-point reductions without a GPU: the oracle's restatement against the
golden vectors written by the UNMODIFIED reference modules (tools/make_golden_reductions.py ran
lightcurver/processes/normalization_calculation.py::calculate_coefficient and
absolute_zeropoint_calculation.py::calculate_zeropoints in the build container), and the closed-form KKT solution the product
uses instead of SLSQP against the reference's optimiser."""
from pathlib import Path

import numpy as np

GOLD = Path(__file__).resolve().parent / 'golden' / 'reference_reductions.npz'


def golden_inputs():
    g = np.load(GOLD)
    lo, hi = g['chi2_bounds']
    flux = g['flux'].copy()
    flux[g['missing']] = np.nan
    d_flux = g['d_flux'].copy()
    d_flux[g['missing']] = np.nan
    norm_flux, norm_dflux = flux.copy(), d_flux.copy()
    bad = ~((g['chi2'] >= lo) & (g['chi2'] <= hi))                        # the SQL filter of normalization_calculation.py:44-46
    norm_flux[bad] = np.nan
    norm_dflux[bad] = np.nan
    return g, flux, norm_flux, norm_dflux


def test_oracle_matches_reference_golden_vectors():
    from oracle import normalization as on
    g, flux, nf, nd = golden_inputs()
    r = on.calculate_coefficient(nf, nd, tol=None)                        # the reference calls SLSQP with its default tolerance
    np.testing.assert_allclose(r['coefficient'], g['ref_coefficient'], rtol=1e-9)
    np.testing.assert_allclose(r['coefficient_uncertainty'], g['ref_coefficient_uncertainty'], rtol=1e-9)
    zp, zs = on.zeropoints(flux, g['catalog_mag'])                        # the zero-point query has no chi2 filter (:64-84)
    np.testing.assert_allclose(zp, g['ref_zeropoint'], rtol=1e-12)
    np.testing.assert_allclose(zs, g['ref_zeropoint_uncertainty'], rtol=1e-9)
    # the recovered coefficients follow the transparency that generated the fluxes (up to one global factor)
    ratio = g['ref_coefficient'] / g['transparency']
    assert np.nanstd(ratio) / np.nanmean(ratio) < 0.01


def test_closed_form_star_scaling_is_the_slsqp_minimum():
    """cost_function_scatter_in_frame (normalization_calculation.py:75-98) is the quadratic form c^T Q c; its minimum under
    mean(c) = 1 from the KKT system equals what SLSQP converges to (tight tolerance), and is never worse than the reference's
    default-tolerance result."""
    from oracle import normalization as on
    from lightcurver_b200.processes.normalization_calculation import solve_star_scaling
    g, flux, nf, nd = golden_inputs()
    F, S = nf.shape
    med = np.nanmedian(nf, axis=0)
    x, d = nf / med, nd / med
    ok = ~(np.isnan(x) | np.isnan(d))
    w = np.where(ok, 1.0 / np.where(ok, d, 1.0), 0.0)
    xx = np.where(ok, x, 0.0)
    W = w.sum(1)
    u = w * xx / W[:, None]
    Q = np.diag((w * xx * xx / W[:, None]).sum(0)) - u.T @ u
    c = solve_star_scaling(Q)
    assert abs(c.mean() - 1) < 1e-12
    tight = on.calculate_coefficient(nf, nd, tol=1e-14)
    np.testing.assert_allclose(c, tight['star_scaling'], rtol=2e-5)
    import pandas as pd
    fp, dp = pd.DataFrame(x.T), pd.DataFrame(d.T)
    cost = lambda cc: on.cost_function_scatter_in_frame(cc, fp, dp)
    np.testing.assert_allclose(cost(c), c @ Q @ c, rtol=1e-9)
    loose = on.calculate_coefficient(nf, nd, tol=None)
    assert cost(c) <= cost(loose['star_scaling']) * (1 + 1e-12)
