"""Field distortion: ``build_psf(field_distortion=True, stamp_coordinates=...)`` and
``apply_distortion`` (lightcurver/processes/psf_modelling.py:169-170, star_photometry.py:293-304, roi_file_preparation.py:
169-180; switched on by the reference's second pipeline pass, tests/test_entire_pipeline/test_run_pipeline_example_config.py:
111-128).  CUDA (through the C ABI) against the CPU oracle: loss and the FULL gradient -- grid, a, x0, y0 and the six distortion
coefficients -- within 1e-5, short fits, the resampling itself, and the two batched drivers end to end."""
import sqlite3

import numpy as np
import pytest

from lightcurver_b200 import synthetic
from lightcurver_b200.conventions import Conventions, DEFAULT

pytestmark = pytest.mark.gpu


def _problem(F, N, n, k, seed):
    from oracle import starred_model as sm
    d = synthetic.make_psf_frames(F, N, n, k, seed=seed)
    data = d['data'].astype(np.float64)
    nm = d['noisemap'].astype(np.float64)
    sc = data.max() / 100.0
    data, nm = data / sc, nm / sc
    weight = d['masks'] / nm ** 2
    rng = np.random.default_rng(seed)
    nu = n * k
    moffat = np.stack([np.full(F, 3.2), np.full(F, 3.6), np.full(F, 0.4), np.full(F, 2.8), np.ones(F)], -1)
    s_fixed = sm.moffat_image(moffat[:, 0], moffat[:, 1], moffat[:, 2], moffat[:, 3], n, k).numpy()
    # The bilinear resampling makes the loss only piecewise smooth in the coefficients: d loss / d theta jumps whenever a sample
    # position crosses a grid line, and on a 192-wide grid a handful of the ~10^5 sample positions of a random theta sit within
    # float32 rounding of such a crossing (measured: the float32 ORACLE then takes the other cell, 1-2 % away from float64 in
    # d loss / d theta, exactly like the kernel).  For a comparison that is well posed in every precision the coefficients and
    # frame positions are dyadic (theta in 1/128, positions in 1/16): all sample positions are then multiples of 1/4096,
    # exactly representable and computed without rounding in float32 and float64 alike.
    xy = (rng.integers(-8, 9, (F, N, 2)) / 16.0).astype(np.float32)
    theta = (rng.integers(-7, 8, (F, 6)) / 128.0).astype(np.float32)
    p = dict(data=data.astype(np.float32), weight=weight.astype(np.float32), off=np.arange(F + 1, dtype=np.int32) * N, moffat=moffat,
             s_fixed=s_fixed, b0=(1e-4 * rng.standard_normal((F, nu, nu))).astype(np.float32),
             a0=((data * d['masks']).sum((-1, -2)) * rng.uniform(0.9, 1.1, (F, N))).astype(np.float32),
             x0=rng.uniform(-0.6, 0.6, (F, N)).astype(np.float32), y0=rng.uniform(-0.6, 0.6, (F, N)).astype(np.float32),
             xy=xy, theta=theta)
    return p


@pytest.mark.parametrize("n,k,N", [(16, 2, 4), (12, 3, 3), (32, 2, 3), (64, 3, 2)])
@pytest.mark.parametrize("conserve", [True, False])
def test_psf_distortion_loss_grad_parity(cuda_device, n, k, N, conserve):
    from lightcurver_b200 import engine
    from oracle import starred_model as sm
    from oracle.conventions import Conventions as OC
    F, nu = 2, n * k
    p = _problem(F, N, n, k, seed=7 + n)
    cv = Conventions(distortion_conserve_flux=conserve)
    J = engine.starlet_scales(nu)
    W = np.random.default_rng(3).uniform(0.5, 2.0, (F, J, nu, nu)).astype(np.float32)
    flat = lambda x: x.reshape(-1, *x.shape[2:])
    out = engine.psf_fit_batch(flat(p['data']), flat(p['weight']), p['off'], k, p['moffat'], p['a0'].ravel(), p['x0'].ravel(), p['y0'].ravel(),
                               background0=p['b0'], W=W, n_iter_analytic=0, n_iter_adabelief=1, lr=1e-3, lam_scales=0.7, lam_hf=1.3,
                               want=('loss0', 'grad_b0', 'grad_s0', 'grad_dist0', 'status'),
                               field_distortion=cv.distortion_mode(), stamp_xy=p['xy'].reshape(-1, 2), distortion0=p['theta'])
    L, (gb, ga, gx, gy, gt) = sm.psf_loss_grad(p['s_fixed'], p['b0'], p['a0'], p['x0'], p['y0'], p['data'], p['weight'], W, n, k, 0.7, 1.3,
                                               cv=OC(distortion_conserve_flux=conserve), theta=p['theta'], xy=p['xy'])
    np.testing.assert_allclose(out['loss0'], L, rtol=1e-5)
    np.testing.assert_allclose(out['grad_b0'], gb, rtol=1e-5, atol=1e-5 * np.abs(gb).max())
    gs = np.stack([ga, gx, gy], -1).reshape(-1, 3)
    np.testing.assert_allclose(out['grad_s0'], gs, rtol=2e-5, atol=1e-5 * np.abs(gs).max(0).max())
    np.testing.assert_allclose(out['grad_dist0'], gt, rtol=2e-5, atol=1e-5 * np.abs(gt).max())
    # the distortion matters in this problem: the same point without it has a visibly different loss
    L0, _ = sm.psf_loss_grad(p['s_fixed'], p['b0'], p['a0'], p['x0'], p['y0'], p['data'], p['weight'], W, n, k, 0.7, 1.3)
    assert np.any(np.abs(L0 - L) > 1e-4 * np.abs(L))


def test_psf_distortion_fit_parity_short(cuda_device):
    """30 AdaBelief iterations with the six coefficients free, against the oracle run in float32 (north star: "STARRED/JAX run in
    float32") and, more loosely, in float64.  Starting from theta = 0 every sample position sits exactly ON a grid point, and
    for the first steps the positions near the centre stay within float32 rounding of it: float32 takes the forward difference
    of the bilinear interpolant where float64 takes the backward one, so the float32 and float64 TRAJECTORIES of theta differ at
    the per-cent level (measured with the oracle alone) while fluxes and loss agree to 1e-4 / 2e-5."""
    import torch
    from lightcurver_b200 import engine
    from oracle import starred_model as sm
    n, k, N, F, T = 16, 2, 5, 2, 30
    p = _problem(F, N, n, k, seed=21)
    z = np.zeros((F, N))
    W = np.stack([sm.psf_noise_weights(p['weight'][f], p['a0'][f], z[f], z[f], n, k).numpy() for f in range(F)]).astype(np.float32)
    flat = lambda x: x.reshape(-1, *x.shape[2:])
    th0 = np.zeros((F, 6), np.float32)
    out = engine.psf_fit_batch(flat(p['data']), flat(p['weight']), p['off'], k, p['moffat'], p['a0'].ravel(), background0=p['b0'], W=W,
                               n_iter_analytic=0, n_iter_adabelief=T, lr=2e-5, lam_scales=1.0, lam_hf=1.0,
                               want=('loss_hist', 'status'), field_distortion=1, stamp_xy=p['xy'].reshape(-1, 2), distortion0=th0)
    kw = dict(lr=2e-5, lam_scales=1.0, lam_hf=1.0, theta0=th0, xy=p['xy'])
    r64 = sm.fit_psf_stage2(p['s_fixed'], p['b0'], p['a0'], z, z, p['data'], p['weight'], W, n, k, T, dtype=torch.float64, **kw)
    r32 = sm.fit_psf_stage2(p['s_fixed'], p['b0'], p['a0'], z, z, p['data'], p['weight'], W, n, k, T, dtype=torch.float32, **kw)
    np.testing.assert_allclose(out['loss_hist'], r64['loss_hist'], rtol=2e-5)
    np.testing.assert_allclose(out['a'].reshape(F, N), r64['a'], rtol=1e-4)
    assert np.abs(r64['theta']).max() > 1e-4                                 # the coefficients moved
    tmax = np.abs(r64['theta']).max()
    err_gpu_32 = np.abs(out['distortion'] - r32['theta']).max() / tmax
    err_gpu_64 = np.abs(out['distortion'] - r64['theta']).max() / tmax
    err_32_64 = np.abs(r32['theta'] - r64['theta']).max() / tmax
    print(f"[parity] distortion coefficients after {T} its: CUDA vs f32 oracle {err_gpu_32:.2e}, CUDA vs f64 {err_gpu_64:.2e}, "
          f"f32 oracle vs f64 {err_32_64:.2e} (of max |theta|)")
    assert err_gpu_32 <= 2e-3
    assert err_gpu_64 <= max(2e-3, 1.5 * err_32_64)
    np.testing.assert_allclose(out['background'], r64['b'], atol=1e-3 * np.abs(p['s_fixed']).max())
    assert (out['status'] == 0).all()


@pytest.mark.parametrize("conserve", [True, False])
def test_apply_distortion_matches_oracle(cuda_device, conserve):
    import torch
    from lightcurver_b200 import engine
    from lightcurver_b200.starred_api import apply_distortion
    from oracle import starred_model as sm
    from oracle.conventions import Conventions as OC
    rng = np.random.default_rng(5)
    Fp, B, nu = 3, 7, 48
    psfs = rng.random((Fp, nu, nu)).astype(np.float32)
    theta = rng.uniform(-0.1, 0.1, (Fp, 6)).astype(np.float32)
    idx = rng.integers(0, Fp, B).astype(np.int32)
    xy = rng.uniform(-0.5, 0.5, (B, 2)).astype(np.float32)
    cv = Conventions(distortion_conserve_flux=conserve)
    out = engine.apply_distortion_batch(psfs, theta, idx, xy, mode=cv.distortion_mode())
    for i in range(B):
        ref = sm.distort_psf(torch.tensor(psfs[idx[i]], dtype=torch.float64), torch.tensor(theta[idx[i]], dtype=torch.float64),
                             torch.tensor(xy[i:i + 1], dtype=torch.float64), OC(distortion_conserve_flux=conserve))[0].numpy()
        np.testing.assert_allclose(out[i], ref, rtol=1e-4, atol=2e-5)
    # the STARRED-shaped single call (star_photometry.py:303) and the empty-kwargs identity
    kd = {'dilation_x': theta[1, 0:2], 'dilation_y': theta[1, 2:4], 'shear': theta[1, 4:6]}
    one = apply_distortion(narrow_psf=psfs[1], kwargs_distortion=kd, star_xy_coordinates=xy[0], conventions=cv)
    np.testing.assert_array_equal(one, engine.apply_distortion_batch(psfs, theta, np.array([1], np.int32), xy[:1], mode=cv.distortion_mode())[0])
    p0 = psfs[0]
    assert apply_distortion(p0, {}, xy[0]) is p0


def test_drivers_with_field_distortion(cuda_device):
    """The reference's second pipeline pass (test_run_pipeline_example_config.py:111-128) sets field_distortion: true: the PSF
    step stores kwargs_distortion next to every PSF (psf_modelling.py:199-202) and the photometry step resamples the narrow PSF
    at every star's frame position before the fit (star_photometry.py:293-304)."""
    from lightcurver_b200.processes.psf_modelling import MemoryStore, model_all_psfs_batched
    from lightcurver_b200.processes.star_photometry import do_star_photometry_batched
    F, N, n, k = 2, 4, 24, 2
    d = synthetic.make_psf_frames(F, N, n, k, seed=91)
    store = MemoryStore()
    gaia = [str(1000 + i) for i in range(N)]
    frames = []
    rng = np.random.default_rng(0)
    for f in range(F):
        rel = f"frames/img{f}.fits"
        store[f"{rel}/frame_shape"] = np.array([2048, 4096])
        for i in range(N):
            store[f"{rel}/data/{gaia[i]}"] = d['data'][f, i]
            store[f"{rel}/noisemap/{gaia[i]}"] = d['noisemap'][f, i]
            store[f"{rel}/cosmicsmask/{gaia[i]}"] = ~d['masks'][f, i]
            store[f"{rel}/image_pixel_coordinates/{gaia[i]}"] = np.array([rng.uniform(0, 4095), rng.uniform(0, 2047)])
        frames.append(dict(id=f + 1, image_relpath=rel, seeing_pixels=float(d['fwhm'][f]), pixel_scale=0.2))
    stars = [dict(name=nm_, gaia_id=g) for nm_, g in zip('abcd', gaia)]
    db = sqlite3.connect(':memory:')
    cfg = dict(subsampling_factor=k, psf_n_iter_analytic=40, psf_n_iter_pixels=200, redo_psf=False, field_distortion=True,
               star_deconv_n_iter=300)
    all_ones = lambda data, nm: np.ones(data.shape, bool)
    written = model_all_psfs_batched(store, db, frames, lambda fid: stars, cfg, 5, automatic_mask_fn=all_ones)
    assert [w[0] for w in written] == [1, 2] and all(w[2] < 2 for w in written)
    for fr in frames:
        g = store[f"{fr['image_relpath']}/psf_abcd/distortion"]
        assert set(g.keys()) == {'dilation_x', 'dilation_y', 'shear'}
        assert all(g[key][...].shape == (2,) and np.isfinite(g[key][...]).all() for key in g.keys())
    res = do_star_photometry_batched(store, db, stars, lambda gid: frames, lambda fid: 'psf_abcd', cfg, 5)
    rows = db.execute("SELECT frame_id, star_gaia_id, flux, flux_uncertainty FROM star_flux_in_frame").fetchall()
    assert len(rows) == F * N
    truth = {(f + 1, gaia[i]): d['flux'][f, i] for f in range(F) for i in range(N)}
    for fid, gid, flux, sig in rows:
        assert sig > 0 and abs(flux - truth[(fid, gid)]) < 8 * sig + 0.05 * truth[(fid, gid)]
    assert all(np.isfinite(r['chi2']) for r in res.values())
