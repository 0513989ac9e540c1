"""Device-side normalisation coefficients and zero points against the golden vectors of the
unmodified reference code and against the pandas oracle on a cfg3-shaped table (transparency truth)."""
import numpy as np
import pytest

from test_reductions_cpu import golden_inputs

pytestmark = pytest.mark.gpu


def test_device_reductions_match_reference_golden_vectors(cuda_device):
    from lightcurver_b200.processes.normalization_calculation import calculate_coefficient_arrays, calculate_zeropoints_arrays
    g, flux, nf, nd = golden_inputs()
    r = calculate_coefficient_arrays(nf, nd)
    # float32 kernels against the float64 pandas code; SLSQP's default tolerance leaves the reference ~1e-6 from the minimum
    np.testing.assert_allclose(r['coefficient'], g['ref_coefficient'], rtol=2e-5)
    np.testing.assert_allclose(r['coefficient_uncertainty'], g['ref_coefficient_uncertainty'], rtol=2e-3, atol=1e-6)
    np.testing.assert_allclose(r['median_flux'], np.nanmedian(nf, axis=0), rtol=1e-6)
    zp, zs = calculate_zeropoints_arrays(flux, g['catalog_mag'])
    np.testing.assert_allclose(zp, g['ref_zeropoint'], atol=2e-5)
    np.testing.assert_allclose(zs, g['ref_zeropoint_uncertainty'], rtol=2e-3, atol=1e-6)


def test_device_reductions_on_a_cfg3_shaped_table(cuda_device):
    """2,000 frames x 20 stars with per-frame transparency (cfg3's generator), NaN holes, torch CUDA inputs (what
    star_photometry_batch hands over): coefficients follow the transparency, everything equals the pandas oracle."""
    import torch
    from oracle import normalization as on
    from lightcurver_b200.processes.normalization_calculation import calculate_coefficient_arrays, calculate_zeropoints_arrays, mask_by_chi2
    rng = np.random.default_rng(3)
    F, S = 2000, 20
    star_flux = 10.0 ** rng.uniform(3.5, 5.0, S)
    c_f = rng.lognormal(0.0, 0.1, F)
    flux = (c_f[:, None] * star_flux[None] * (1 + 0.01 * rng.standard_normal((F, S)))).astype(np.float32)
    dflux = (0.01 * flux * rng.uniform(0.5, 2.0, (F, S))).astype(np.float32)
    chi2 = rng.uniform(0.5, 1.5, (F, S)).astype(np.float32)
    chi2[rng.random((F, S)) < 0.03] = 5.0
    fd = mask_by_chi2(torch.from_numpy(flux).cuda(), torch.from_numpy(chi2).cuda(), 0.0, 2.0)
    r = calculate_coefficient_arrays(fd, torch.from_numpy(dflux).cuda())
    nf = np.where(chi2 <= 2.0, flux, np.nan).astype(np.float64)
    nd = np.where(chi2 <= 2.0, dflux, np.nan).astype(np.float64)
    o = on.calculate_coefficient(nf, nd, tol=1e-14)
    np.testing.assert_allclose(r['star_scaling'], o['star_scaling'], rtol=1e-4)
    np.testing.assert_allclose(r['coefficient'], o['coefficient'], rtol=2e-5)
    np.testing.assert_allclose(r['coefficient_uncertainty'], o['coefficient_uncertainty'], rtol=5e-3, atol=1e-6)
    ratio = r['coefficient'] / c_f
    assert np.std(ratio) / np.mean(ratio) < 0.005
    cmag = (25.0 - 2.5 * np.log10(star_flux)).astype(np.float32)
    zp, zs = calculate_zeropoints_arrays(fd, cmag)
    ozp, ozs = on.zeropoints(nf, cmag.astype(np.float64))
    np.testing.assert_allclose(zp, ozp, atol=3e-5)
    np.testing.assert_allclose(zs, ozs, rtol=5e-3, atol=1e-6)
