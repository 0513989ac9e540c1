"""The reference-facing Python surface, exercised the way lightcurver's own tests exercise STARRED
(tests/test_starred_calls/test_starred_calls.py in the reference): same inputs, same assertions on
keys / types / shapes, plus the reference's acceptance bound chi2 < 2
(tests/test_entire_pipeline/test_run_pipeline_example_config.py:18-21) on synthetic frames."""
import numpy as np
import pytest

from lightcurver_b200.conventions import DEFAULT

pytestmark = pytest.mark.gpu


def _ref_test_inputs():
    x, y = np.meshgrid(np.arange(-8, 8), np.arange(-8, 8))
    gauss = np.exp(-0.1 * (x ** 2 + y ** 2))
    rng = np.random.default_rng(0)
    data = 0.1 * rng.random((5, 16, 16)) + np.repeat(gauss[None, :, :], repeats=5, axis=0)
    noisemap = 0.1 * np.ones((5, 16, 16))
    psf = np.repeat(gauss[None, :, :], repeats=5, axis=0)
    return data, noisemap, psf


@pytest.mark.parametrize("flags", [dict(starlet_global_background=False)])
def test_do_one_star_forward_modelling_contract(cuda_device, flags):
    from lightcurver_b200.processes.star_photometry import do_one_star_forward_modelling
    data, noisemap, psf = _ref_test_inputs()
    d0 = data.copy()
    n_iter = 50
    result = do_one_star_forward_modelling(data, noisemap, psf, 1, n_iter, **flags)
    assert isinstance(result, dict)
    for key in ('scale', 'kwargs_final', 'fluxes', 'fluxes_uncertainties', 'chi2', 'chi2_per_frame', 'loss_curve', 'residuals'):
        assert key in result
    assert isinstance(result['scale'], float) and result['scale'] > 0
    assert isinstance(result['kwargs_final'], dict)
    assert isinstance(result['fluxes'], np.ndarray) and isinstance(result['fluxes_uncertainties'], np.ndarray)
    assert result['fluxes'].ndim == 1 and result['fluxes_uncertainties'].ndim == 1
    assert result['fluxes'].size == result['fluxes_uncertainties'].size == data.shape[0]
    assert isinstance(result['chi2'], float) and result['chi2'] >= 0
    assert isinstance(result['chi2_per_frame'], np.ndarray) and result['chi2_per_frame'].ndim == 1
    assert len(result['chi2_per_frame']) == data.shape[0]
    assert len(result['loss_curve']) == n_iter
    assert result['residuals'].shape == data.shape
    # star_photometry.py:47-49: the caller's arrays are rescaled in place
    np.testing.assert_allclose(data * result['scale'], d0, rtol=1e-12)
    assert result['deconvolved_image'].shape == (16, 16) and result['starlet_background'].shape == (16, 16)


def test_build_psf_contract(cuda_device):
    from lightcurver_b200.procedures.psf_routines import build_psf
    data, noisemap, _ = _ref_test_inputs()
    result = build_psf(data, noisemap, subsampling_factor=1, n_iter_analytic=5, n_iter_adabelief=10,
                       masks=np.ones_like(data), guess_method_star_position='center')
    assert isinstance(result, dict)
    for key in ('full_psf', 'adabelief_extra_fields', 'narrow_psf', 'chi2', 'residuals'):
        assert key in result
    assert 'loss_history' in result['adabelief_extra_fields'] and len(result['adabelief_extra_fields']['loss_history']) == 10
    assert result['narrow_psf'].shape == (16, 16) and result['full_psf'].shape == (16, 16)
    assert result['residuals'].shape == data.shape
    km = result['kwargs_psf']['kwargs_moffat']
    assert float(0.5 * (km['fwhm_x'] + km['fwhm_y']).item()) > 0          # psf_modelling.py:177-179
    assert set(result['kwargs_psf']) >= {'kwargs_moffat', 'kwargs_gaussian', 'kwargs_background', 'kwargs_distortion'}
    assert abs(result['narrow_psf'].sum() - 1) < 1e-4 and abs(result['full_psf'].sum() - 1) < 1e-4
    assert result['kwargs_psf']['kwargs_distortion'] == {}
    with pytest.raises(ValueError):                                       # the positions of the stamps in the frame are needed
        build_psf(data, noisemap, 1, field_distortion=True)
    # the call of psf_modelling.py:164-171 with field_distortion on: same keys + the distortion coefficients (:199-202)
    xy = np.array([[-0.3, 0.2], [0.1, -0.4], [0.4, 0.4], [-0.2, -0.1], [0.0, 0.3]])
    r2 = build_psf(data, noisemap, subsampling_factor=1, n_iter_analytic=5, n_iter_adabelief=10, masks=np.ones_like(data),
                   guess_method_star_position='center', guess_fwhm_pixels=3.0, field_distortion=True, stamp_coordinates=xy)
    assert set(r2['kwargs_psf']['kwargs_distortion']) == {'dilation_x', 'dilation_y', 'shear'}
    assert all(v.shape == (2,) and np.isfinite(v).all() for v in r2['kwargs_psf']['kwargs_distortion'].values())
    assert r2['narrow_psf'].shape == (16, 16) and len(r2['adabelief_extra_fields']['loss_history']) == 10


def test_pipeline_shaped_run_chi2_below_2(cuda_device):
    """Stand-in for BASELINE cfg1 (2 frames x 2 stars x 24x24, k=2, as in the reference's pipeline test):
    PSF fit then photometry through the public API; the reference's own acceptance bound is chi2 < 2."""
    from lightcurver_b200 import synthetic
    from lightcurver_b200.procedures.psf_routines import build_psf_batch
    from lightcurver_b200.processes.star_photometry import star_photometry_batch
    F, N, n, k = 2, 2, 24, 2
    d = synthetic.make_psf_frames(F, N, n, k, seed=77)
    res = build_psf_batch(d['data'], d['noisemap'], k, masks=d['masks'], n_iter_analytic=100, n_iter_adabelief=500,
                          guess_method_star_position='center', guess_fwhm_pixels=d['fwhm'])
    assert all(r['chi2'] < 2 for r in res)
    assert all(r['status'] == 0 for r in res)
    psfs = np.stack([r['narrow_psf'] for r in res])
    ph = star_photometry_batch(d['data'], d['noisemap'], psfs, k, n_iter=500, masks=None)
    assert (ph['chi2_per_frame'] < 2).all()
    rel = np.abs(ph['fluxes'] - d['flux']) / d['flux']        # pixel-sum units whatever the D_k convention
    assert rel.max() < 0.05
    # ragged: second frame loses a star (psf_modelling.py:144-153)
    res2 = build_psf_batch([d['data'][0], d['data'][1][:1]], [d['noisemap'][0], d['noisemap'][1][:1]], k,
                           masks=[d['masks'][0], d['masks'][1][:1]], n_iter_analytic=50, n_iter_adabelief=50,
                           guess_method_star_position='center', guess_fwhm_pixels=d['fwhm'])
    assert res2[1]['residuals'].shape == (1, n, n) and res2[0]['residuals'].shape == (2, n, n)


def test_device_prepare_matches_host_policies(cuda_device):
    """lcb_psf_prepare_batch / lcb_phot_prepare_batch against a numpy statement of the same policies
    (psf_modelling.py:136-140 + build_psf normalisation and smart guess; star_photometry.py:47-64, 309-316)."""
    from lightcurver_b200 import engine
    rng = np.random.default_rng(5)
    # ---- PSF side: ragged frames, NaNs, masks, non-positive noise
    counts = [3, 1, 4]
    n, k = 16, 2
    off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    sumN = int(off[-1])
    img = rng.normal(5.0, 2.0, (sumN, n, n)).astype(np.float32)
    img[:, 6:10, 6:10] += 200.0
    nm = rng.uniform(0.5, 2.0, (sumN, n, n)).astype(np.float32)
    mk = rng.random((sumN, n, n)) > 0.05
    img[0, 3, 4] = np.nan; nm[0, 3, 4] = np.nan; img[2, 1, 1] = np.inf; nm[5, 0, 0] = 0.0; nm[6, 2, 2] = np.nan
    for method, dmean in (('center', True), ('max', False), ('barycenter', False)):
        cv_apf = k * k if dmean else 1.0
        prep = engine.psf_prepare_batch(img, nm, mk, off, k, norm_scale=100.0, downsample_mean=dmean, guess_method=method)
        star_max = np.fmax.reduce(img.reshape(sumN, -1), axis=1)
        norms = np.fmax.reduceat(star_max, off[:-1]).astype(np.float64) / 100.0
        norms[~np.isfinite(norms) | (norms <= 0)] = 1.0
        inv = np.repeat(1.0 / norms, counts).astype(np.float32)[:, None, None]
        d, s = img * inv, nm * inv
        good = np.isfinite(d) & np.isfinite(s) & (s > 0) & mk
        d = np.where(np.isfinite(d), d, 0.0).astype(np.float32)
        w = np.where(good, 1.0 / np.where(good, s, 1.0) ** 2, 0.0)
        flux = np.where(good, d, 0.0).sum((-1, -2), dtype=np.float64)
        np.testing.assert_allclose(prep['norm'].cpu().numpy(), norms, rtol=1e-6)
        np.testing.assert_allclose(prep['data'].cpu().numpy(), d, rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(prep['weight'].cpu().numpy(), w, rtol=2e-6)
        np.testing.assert_allclose(prep['a0'].cpu().numpy(), np.maximum(flux, 1e-6) * cv_apf, rtol=1e-5)
        ctr = (n - 1) / 2.0
        if method == 'center':
            x0 = y0 = np.zeros(sumN)
        elif method == 'max':
            flat = np.where(good, d, -np.inf).reshape(sumN, -1).argmax(-1)
            x0, y0 = (flat % n) - ctr, (flat // n) - ctr
        else:
            ww = np.clip(np.where(good, d, 0.0), 0.0, None)
            tot = np.maximum(ww.sum((-1, -2)), 1e-30)
            ax = np.arange(n) - ctr
            x0, y0 = (ww.sum(-2) * ax).sum(-1) / tot, (ww.sum(-1) * ax).sum(-1) / tot
        np.testing.assert_allclose(prep['x0'].cpu().numpy(), x0, atol=1e-4)
        np.testing.assert_allclose(prep['y0'].cpu().numpy(), y0, atol=1e-4)
    # ---- photometry side
    F, S, n = 5, 3, 17
    data = rng.normal(1.0, 0.3, (F, S, n, n)).astype(np.float32)
    data[:, :, 7:10, 7:10] += rng.uniform(50, 100, (F, S, 1, 1)).astype(np.float32)
    noise = rng.uniform(0.5, 2.0, (F, S, n, n)).astype(np.float32)
    masks = np.ones((F, S, n, n), bool)
    masks[1, 2, 4, 4] = False; masks[3, 0, 0, 0] = False
    data[0, 1, 2, 3] = np.nan; noise[2, 2, 5, 5] = np.nan
    prep = engine.phot_prepare_batch(data, noise, masks, k)
    d, s = data.copy(), noise.copy()
    isn = np.isnan(d) | np.isnan(s)
    d[isn] = 0.0; s[isn] = 1e7
    s[~masks.all((-1, -2))] *= 1000.0
    scale = d.max(axis=(0, 2, 3))
    inv = (1.0 / scale).astype(np.float32)[None, :, None, None]
    d *= inv; s *= inv
    edges = np.stack([np.median(d[:, :, 0, :], -1), np.median(d[:, :, :, 0], -1), np.median(d[:, :, -1, :], -1), np.median(d[:, :, :, -1], -1)])
    bg = np.nan_to_num(edges.mean((0, 1)), nan=0.0)
    a_est = (d.sum((-1, -2), dtype=np.float64) - n * n * bg[None]) * DEFAULT.amplitude_per_flux(k)
    np.testing.assert_allclose(prep['scale'].cpu().numpy(), scale, rtol=1e-6)
    np.testing.assert_allclose(prep['data'].cpu().numpy().reshape(F, S, n, n), d, rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(prep['weight'].cpu().numpy().reshape(F, S, n, n), 1.0 / s.astype(np.float64) ** 2, rtol=3e-6)
    np.testing.assert_allclose(prep['a0'].cpu().numpy().reshape(F, S), a_est, rtol=2e-5, atol=1e-4)


def test_alternate_conventions_parity(cuda_device):
    """Every recalled STARRED convention is a switch: with D_k = block MEAN (the default is the block SUM, which is what
    lightcurver's use of pixel sums as initial_a and of `a` as the flux implies) and chi2 without the
    1/2, K2, K1 (fast and generic path) and K3 still match the oracle to 1e-5."""
    import dataclasses
    from lightcurver_b200 import engine, _lib, synthetic
    from lightcurver_b200.conventions import DEFAULT, apply_to_library
    from lightcurver_b200.processes.roi_modelling import JointDeconvolution
    from oracle import starred_model as sm
    from oracle.conventions import DEFAULT as ODEF
    cvp = dataclasses.replace(DEFAULT, downsample_mean=True, chi2_half=False)
    cvo = dataclasses.replace(ODEF, downsample_mean=True, chi2_half=False)
    rng = np.random.default_rng(8)
    try:
        apply_to_library(cvp)
        # K2
        n, k, F, S = 16, 2, 2, 3
        d = synthetic.make_phot_frames(F, S, n, k, seed=4)
        data = d['data'].reshape(F * S, n, n); w = (1.0 / d['noisemap'].reshape(F * S, n, n).astype(np.float64) ** 2).astype(np.float32)
        idx = np.repeat(np.arange(F), S).astype(np.int32)
        a0 = (data.sum((-1, -2)) * rng.uniform(0.8, 1.2, F * S)).astype(np.float32)
        dx0 = rng.uniform(-0.5, 0.5, F * S).astype(np.float32); dy0 = rng.uniform(-0.5, 0.5, F * S).astype(np.float32)
        out = engine.phot_fit_batch(data, w, d['psf'], idx, a0, k, n_iter=1, dx0=dx0, dy0=dy0, want_grad0=True)
        L, (ga, gx, gy) = sm.phot_loss_grad(d['psf'][idx], data, w, a0, dx0, dy0, n, k, cv=cvo)
        np.testing.assert_allclose(out['loss0'], L, rtol=1e-5)
        g = np.stack([ga, gx, gy], -1)
        np.testing.assert_allclose(out['grad0'], g, rtol=2e-5, atol=1e-5 * np.abs(g).max())
        # K1, fast (32 x 2) and generic (18 x 3) paths
        for (n, k, N) in [(32, 2, 3), (18, 3, 2)]:
            Fp = 2
            dd = synthetic.make_psf_frames(Fp, N, n, k, seed=20 + n)
            sc = dd['data'].max() / 100.0
            dat = (dd['data'] / sc).astype(np.float32); nm = (dd['noisemap'] / sc).astype(np.float32)
            wt = (dd['masks'] / nm.astype(np.float64) ** 2).astype(np.float32)
            nu = n * k
            J = engine.starlet_scales(nu)
            W = rng.uniform(0.5, 2.0, (Fp, J, nu, nu)).astype(np.float32)
            b0 = (1e-4 * rng.standard_normal((Fp, nu, nu))).astype(np.float32)
            a00 = (dat.sum((-1, -2)) * rng.uniform(0.9, 1.1, (Fp, N))).astype(np.float32)
            x00 = rng.uniform(-0.6, 0.6, (Fp, N)).astype(np.float32); y00 = rng.uniform(-0.6, 0.6, (Fp, N)).astype(np.float32)
            mof = np.stack([np.full(Fp, 3.2), np.full(Fp, 3.5), np.full(Fp, 0.3), np.full(Fp, 2.8), np.ones(Fp)], -1)
            off = np.arange(Fp + 1, dtype=np.int32) * N
            o = engine.psf_fit_batch(dat.reshape(-1, n, n), wt.reshape(-1, n, n), off, k, mof, a00.ravel(), x00.ravel(), y00.ravel(),
                                     background0=b0, W=W, n_iter_analytic=0, n_iter_adabelief=1, lr=1e-5, lam_scales=0.7, lam_hf=1.3,
                                     want=('loss0', 'grad_b0', 'grad_s0'))
            s_fixed = sm.moffat_image(mof[:, 0], mof[:, 1], mof[:, 2], mof[:, 3], n, k).numpy()
            L, (gb, ga, gx, gy) = sm.psf_loss_grad(s_fixed, b0, a00, x00, y00, dat, wt, W, n, k, 0.7, 1.3, cv=cvo)
            np.testing.assert_allclose(o['loss0'], L, rtol=1e-5)
            np.testing.assert_allclose(o['grad_b0'], gb, rtol=1e-5, atol=1e-5 * np.abs(gb).max())
            gs = np.stack([ga, gx, gy], -1).reshape(-1, 3)
            np.testing.assert_allclose(o['grad_s0'], gs, rtol=2e-5, atol=1e-5 * np.abs(gs).max(0).max())
        # K3
        import torch
        E, n, k, M = 2, 16, 2, 2
        nu = n * k
        psf = sm.moffat_image(torch.tensor([3.0, 3.3]), torch.tensor([3.2, 3.1]), torch.tensor([0.2, 1.0]), torch.tensor([3.0, 2.7]), 12, k).numpy().astype(np.float32)
        prm = dict(h=(0.05 * rng.standard_normal(nu * nu)).astype(np.float32), mean=rng.uniform(-0.01, 0.01, E).astype(np.float32),
                   a=rng.uniform(20, 60, (E, M)).astype(np.float32), c_x=rng.uniform(-3, 3, M).astype(np.float32),
                   c_y=rng.uniform(-3, 3, M).astype(np.float32), dx=rng.uniform(-1, 1, E).astype(np.float32), dy=rng.uniform(-1, 1, E).astype(np.float32))
        alpha = rng.uniform(-0.1, 0.1, E).astype(np.float32)
        dat3 = rng.standard_normal((E, n, n)).astype(np.float32); w3 = rng.uniform(0.5, 2.0, (E, n, n)).astype(np.float32)
        jd = JointDeconvolution(dat3, w3, psf, k, M, cvp)
        jd.set_params(alpha=alpha, **prm)
        jd.set_reg(0.8, 1.2, 50.0, lam_pts=0.2, lam_fu=3.0, conventions=cvp)
        g3 = jd.loss_grad()
        jd.close()
        L3, go = sm.deconv_loss_grad(prm, dict(alpha=alpha), psf, dat3, w3, None, n, k,
                                     dict(lam_scales=0.8, lam_hf=1.2, lam_pos=50.0, lam_pts=0.2, lam_fu=3.0), cv=cvo)
        assert abs(g3['loss'] - L3) <= 1e-5 * abs(L3)
        for nm_ in ('h', 'mean', 'a', 'c_x', 'c_y', 'dx', 'dy'):
            ref = go[nm_].reshape(-1)
            np.testing.assert_allclose(g3[nm_].reshape(-1), ref, rtol=2e-5, atol=2e-5 * np.abs(ref).max(), err_msg=nm_)
    finally:
        apply_to_library(DEFAULT)


def test_in_process_multi_gpu_fan_out(cuda_device):
    """build_psf_batch / star_photometry_batch with devices=2: frames (resp. stars) split over two GPUs from two host
    threads of ONE process, no collective; results identical to the single-GPU call (items are independent)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    from lightcurver_b200 import synthetic, engine
    from lightcurver_b200.procedures.psf_routines import build_psf_batch
    from lightcurver_b200.processes.star_photometry import star_photometry_batch
    assert engine.split_by_work([3, 1, 1, 1, 3, 3], 2) == [(0, 4), (4, 6)] and engine.split_by_work([1, 1], 4) == [(0, 1), (1, 2)]
    F, N, n, k = 7, 4, 16, 2
    d = synthetic.make_psf_frames(F, N, n, k, seed=5)
    images = [d['data'][f][:N - (f % 2)] for f in range(F)]            # ragged
    noise = [d['noisemap'][f][:N - (f % 2)] for f in range(F)]
    masks = [d['masks'][f][:N - (f % 2)] for f in range(F)]
    kw = dict(n_iter_analytic=30, n_iter_adabelief=60, guess_method_star_position='center', guess_fwhm_pixels=d['fwhm'])
    one = build_psf_batch(images, noise, k, masks=masks, return_dicts=False, **kw)
    two = build_psf_batch(images, noise, k, masks=masks, return_dicts=False, devices=2, **kw)
    for key in ('narrow_psf', 'a', 'x0', 'chi2', 'loss_hist', 'moffat', 'norms', 'star_off'):
        assert np.array_equal(one[key], two[key]), key
    dicts = build_psf_batch(images, noise, k, masks=masks, devices='all', **kw)
    assert len(dicts) == F and np.array_equal(dicts[3]['narrow_psf'], one['narrow_psf'][3])
    p1 = star_photometry_batch(d['data'], d['noisemap'], one['narrow_psf'], k, n_iter=80, masks=d['masks'], want_residuals=True)
    p2 = star_photometry_batch(d['data'], d['noisemap'], one['narrow_psf'], k, n_iter=80, masks=d['masks'], want_residuals=True, devices=[0, 1])
    for key in p1:
        assert np.array_equal(p1[key], p2[key]), key


@pytest.mark.parametrize("flags", [dict(starlet_global_background=True, uniform_background_per_epoch=False),
                                   dict(starlet_global_background=False, uniform_background_per_epoch=True)])
def test_coupled_photometry_of_many_stars_in_one_call(cuda_device, flags):
    """star_photometry.py:257 loops over the stars; with the background flags every star is one joint fit (shared h / c / clip norm,
    :74-87).  ``do_stars_forward_modelling_coupled`` runs all those fits in ONE library call (lcb_deconv_run_many: a handle per
    star, iterations interleaved on separate streams): the results must be those of the one-star function, star by star, also
    for stars with different numbers of epochs."""
    from lightcurver_b200 import synthetic
    from lightcurver_b200.processes.star_photometry import do_one_star_forward_modelling, do_stars_forward_modelling_coupled
    n, k, n_iter = 16, 2, 40
    stacks = []
    for s, E in enumerate((5, 3, 6)):
        d = synthetic.make_phot_frames(E, 1, n, k, seed=40 + s)
        stacks.append((d['data'][:, 0].astype(np.float64), d['noisemap'][:, 0].astype(np.float64), d['psf']))
    one = [do_one_star_forward_modelling(dd.copy(), nm.copy(), psf, k, n_iter, **flags) for dd, nm, psf in stacks]
    many = do_stars_forward_modelling_coupled([(dd.copy(), nm.copy(), psf) for dd, nm, psf in stacks], k, n_iter, **flags)
    assert len(many) == len(one)
    for a, b in zip(one, many):
        assert a['scale'] == b['scale']
        np.testing.assert_allclose(b['fluxes'], a['fluxes'], rtol=1e-6)
        np.testing.assert_allclose(b['fluxes_uncertainties'], a['fluxes_uncertainties'], rtol=1e-5)
        np.testing.assert_allclose(b['loss_curve'], a['loss_curve'], rtol=1e-6)
        np.testing.assert_allclose(b['residuals'], a['residuals'], atol=1e-5 * np.abs(a['residuals']).max() + 1e-9)
        assert len(b['loss_curve']) == n_iter and np.isfinite(b['chi2'])


def test_strided_batch_slices_go_up_in_one_pitched_copy(cuda_device):
    """The in-process fan-out hands every device a slice ``batch[:, lo:hi]`` of the (F, S, n, n) arrays: the host layer uploads
    it with ONE pitched copy (lcb_copy_2d) instead of gathering it on the host.  Same results as with a contiguous copy of the
    slice (bit for bit), outputs staged through page-locked memory."""
    from lightcurver_b200 import engine, synthetic
    from lightcurver_b200.processes.star_photometry import star_photometry_batch
    import torch
    F, S, n, k = 6, 5, 16, 2
    d = synthetic.make_phot_frames(F, S, n, k, seed=5)
    view_d, view_n = d['data'][:, 1:4], d['noisemap'][:, 1:4]
    assert not view_d.flags.c_contiguous
    up = engine._to_device(view_d, torch.float32)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(up.cpu().numpy(), np.ascontiguousarray(view_d))
    a = star_photometry_batch(view_d, view_n, d['psf'], k, n_iter=30)
    b = star_photometry_batch(np.ascontiguousarray(view_d), np.ascontiguousarray(view_n), d['psf'], k, n_iter=30)
    for key in ('fluxes', 'fluxes_uncertainties', 'chi2_per_frame', 'dx', 'dy', 'scale', 'loss_curve'):
        np.testing.assert_array_equal(a[key], b[key], err_msg=key)
    # a pinned source takes the same route
    pinned = torch.from_numpy(d['data']).pin_memory().numpy()
    up2 = engine._to_device(pinned[:, 2:5], torch.float32)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(up2.cpu().numpy(), d['data'][:, 2:5])


def test_two_host_threads_with_host_buffers_and_different_conventions(cuda_device):
    """The host layer is called from one thread per GPU by the in-process fan-out, and lightcurver may keep several handles
    alive: LCB_MEM_HOST calls lease a staging arena per call and device, conventions are per thread.  Two threads hammer the
    SAME device with host-buffer calls, one under the default conventions (block sum) and one under the block mean; every
    result must equal the one obtained serially under that thread's conventions."""
    import threading
    from lightcurver_b200 import engine, synthetic
    from lightcurver_b200.conventions import Conventions, apply_to_library
    n, k, F, S = 16, 2, 4, 3
    d = synthetic.make_phot_frames(F, S, n, k, seed=8)
    data = d['data'].reshape(-1, n, n)
    w = (1.0 / d['noisemap'].reshape(-1, n, n).astype(np.float64) ** 2).astype(np.float32)
    idx = np.repeat(np.arange(F), S).astype(np.int32)
    cvs = [Conventions(), Conventions(downsample_mean=True)]

    def run(cv, reps):
        apply_to_library(cv)                                   # this thread only
        a0 = (data.sum((-1, -2)) * cv.amplitude_per_flux(k)).astype(np.float32)
        outs = []
        for _ in range(reps):
            o = engine.phot_fit_batch(data, w, d['psf'], idx, a0, k, 25)
            outs.append((o['a'].copy(), o['dx'].copy(), o['loss_hist'].copy()))
        return outs
    serial = [run(cv, 1)[0] for cv in cvs]
    assert not np.allclose(serial[0][0], serial[1][0], rtol=1e-3)          # the two conventions give different amplitudes (x k^2)
    results, errors = [None, None], []

    def worker(i):
        try:
            results[i] = run(cvs[i], 12)
        except Exception as exc:                                  # pragma: no cover
            errors.append(exc)
    ts = [threading.Thread(target=worker, args=(i,)) for i in range(2)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    apply_to_library(DEFAULT)
    assert not errors, errors
    for i in range(2):
        for got in results[i]:
            for x, y in zip(got, serial[i]):
                np.testing.assert_array_equal(x, y)
