"""The STARRED-shaped front end (lightcurver_b200.starred_api): the call sequences lightcurver makes into `starred`
(star_photometry.py:66-137, starred_utilities.py:27-38, roi_modelling.py:213-335, 387) run on the sm_100a kernels with
the argument names, kwargs dict-of-dicts and return shapes those call sites rely on."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture
def starred_installed():
    import lightcurver_b200.starred_api as sa
    sa.install()
    yield sa
    sa.uninstall()


def _star_stack(E=5, n=16, k=2, seed=2):
    from lightcurver_b200 import synthetic
    d = synthetic.make_phot_frames(E, 1, n, k, seed=seed)
    data = d['data'][:, 0].astype(np.float64)
    noisemap = d['noisemap'][:, 0].astype(np.float64)
    psf = d['psf'].astype(np.float64)                  # (E, nu, nu), unit sum
    return d, data, noisemap, psf


@pytest.mark.parametrize("uniform_background_per_epoch,starlet_global_background", [(False, True), (False, False), (True, True)])
def test_star_photometry_call_sequence(cuda_device, starred_installed, uniform_background_per_epoch, starlet_global_background):
    """The STARRED calls of do_one_star_forward_modelling (star_photometry.py:66-137), written against the shim."""
    from copy import deepcopy
    from starred.deconvolution.deconvolution import setup_model
    from starred.deconvolution.loss import Loss
    from starred.deconvolution.parameters import ParametersDeconv
    from starred.optim.optimization import Optimizer
    from starred.optim.inference_base import FisherCovariance
    from starred.utils.noise_utils import propagate_noise
    E, n, k, n_iter = 5, 16, 2, 120
    d, data, noisemap, psf = _star_stack(E, n, k)
    scale = np.nanmax(data)
    data /= scale
    noisemap /= scale
    sigma_2 = noisemap ** 2
    a_est = list(np.nansum(data, axis=(1, 2)))
    model, kwargs_init, kwargs_up, kwargs_down, kwargs_fixed = setup_model(data, sigma_2, psf, np.array([0.]), np.array([0.]), k, a_est)
    assert set(kwargs_init) == {'kwargs_analytic', 'kwargs_background', 'kwargs_sersic'}
    assert set(kwargs_init['kwargs_analytic']) == {'c_x', 'c_y', 'dx', 'dy', 'a', 'alpha'}
    assert kwargs_init['kwargs_background']['h'].shape == ((n * k) ** 2,) and np.all(kwargs_init['kwargs_background']['h'] == 0)
    kwargs_fixed = {'kwargs_analytic': {'alpha': kwargs_init['kwargs_analytic']['alpha']},
                    'kwargs_background': {'h': kwargs_init['kwargs_background']['h'], 'mean': np.ravel([0. for _ in range(len(data))])},
                    'kwargs_sersic': {}}
    if uniform_background_per_epoch:
        del kwargs_fixed['kwargs_background']['mean']
    if starlet_global_background:
        del kwargs_fixed['kwargs_background']['h']
    parameters = ParametersDeconv(kwargs_init=kwargs_init, kwargs_fixed=kwargs_fixed, kwargs_up=kwargs_up, kwargs_down=kwargs_down)
    kwargs_loss = {'data': data, 'deconv_class': model, 'param_class': parameters, 'sigma_2': sigma_2,
                   'regularization_terms': 'l1_starlet', 'regularization_strength_scales': 3.0,
                   'regularization_strength_hf': 3.0, 'regularization_strength_flux_uniformity': 0.}
    if starlet_global_background:
        W = propagate_noise(model, noisemap, kwargs_init, wavelet_type_list=['starlet'], method='SLIT', num_samples=200, seed=1,
                            likelihood_type='chi2', verbose=False, upsampling_factor=k)[0]
        assert W.shape[1:] == (n * k, n * k) and np.isfinite(W).all()
        kwargs_loss['W'] = W
    loss = Loss(**kwargs_loss)
    optim = Optimizer(loss, parameters, method='adabelief')
    out = optim.minimize(max_iterations=n_iter, min_iterations=None, init_learning_rate=1e-3, schedule_learning_rate=True,
                         restart_from_init=True, stop_at_loss_increase=False, progress_bar=True, return_param_history=True)
    assert len(out) == 4 and len(out[2]['loss_history']) == n_iter
    kwargs_final = parameters.best_fit_values(as_kwargs=True)
    modelled_pixels = model.model(kwargs_final)
    residuals = data - np.array(modelled_pixels)
    chi2_per_frame = np.nansum((residuals ** 2 / sigma_2), axis=(1, 2)) / model.image_size ** 2
    fluxes = scale * np.array(kwargs_final['kwargs_analytic']['a'])
    assert len(optim.loss_history) == n_iter and optim.loss_history[-1] < optim.loss_history[0]
    assert fluxes.shape == (E,) and np.all(chi2_per_frame < 3.0)
    # `a` is the flux (pixel sum) in this front end: within a few per cent of the truth after 120 iterations
    np.testing.assert_allclose(fluxes, d['transparency'] * d['star_flux'][0], rtol=0.05)
    if not uniform_background_per_epoch:
        assert np.all(kwargs_final['kwargs_background']['mean'] == 0)
    if not starlet_global_background:
        assert np.all(kwargs_final['kwargs_background']['h'] == 0)
    # get_flux_uncertainties (starred_utilities.py:27-38)
    kf = deepcopy(kwargs_final)
    del kf['kwargs_analytic']['a']
    p2 = ParametersDeconv(kwargs_init=kwargs_final, kwargs_fixed=kf, kwargs_up=kwargs_up, kwargs_down=kwargs_down)
    l2 = Loss(data, model, p2, noisemap ** 2, regularization_terms='l1_starlet')
    o2 = Optimizer(l2, p2, method='l-bfgs-b')
    o2.minimize(maxiter=10)
    fish = FisherCovariance(p2, o2, diagonal_only=True)
    fish.compute_fisher_information()
    sig = np.array(fish.get_kwargs_sigma()['kwargs_analytic']['a'])
    assert sig.shape == (E,) and np.all(sig > 0) and np.all(scale * sig < 0.05 * fluxes)
    deconv, bkg = model.getDeconvolved(kwargs_final, 0)
    assert deconv.shape == (n * k, n * k) and bkg.shape == (n * k, n * k)


def test_roi_call_sequence(cuda_device, starred_installed):
    """The STARRED calls of do_modelling_of_roi (roi_modelling.py:213-335): stage 1 L-BFGS-B with a prior and the flux
    scatter penalty, SLIT weights, stage 2 AdaBelief with every regularisation strength of the reference."""
    from copy import deepcopy
    import torch
    from starred.deconvolution.deconvolution import setup_model
    from starred.deconvolution.loss import Loss, Prior
    from starred.deconvolution.parameters import ParametersDeconv
    from starred.optim.optimization import Optimizer
    from starred.utils.noise_utils import propagate_noise
    from oracle import starred_model as sm
    rng = np.random.default_rng(3)
    E, n, k, M = 6, 16, 2, 2
    nu = n * k
    fw = rng.uniform(2.5, 3.5, E)
    s = sm.moffat_image(torch.tensor(fw), torch.tensor(fw * 1.05), torch.tensor(rng.uniform(0, 3, E)), torch.full((E,), 3.0, dtype=torch.float64), 12, k).numpy()
    cx, cy = np.array([-2.0, 2.5]), np.array([1.0, -1.5])
    a_true = rng.uniform(2.0, 4.0, (E, M))
    dxt, dyt = rng.uniform(-0.5, 0.5, E), rng.uniform(-0.5, 0.5, E)
    t = lambda v: torch.tensor(v, dtype=torch.float64)
    import dataclasses
    from oracle.conventions import DEFAULT as ODEF
    cvo = dataclasses.replace(ODEF, downsample_mean=False)
    clean = sm.deconv_model(torch.zeros(nu, nu, dtype=torch.float64), torch.zeros(E, dtype=torch.float64), t(a_true), t(cx), t(cy), t(dxt), t(dyt),
                            torch.zeros(E, dtype=torch.float64), t(s), n, k, cvo).numpy()
    noisemap = np.sqrt(1e-4 + 1e-3 * np.abs(clean))
    data = clean + noisemap * rng.standard_normal(clean.shape)
    initial_a = list(a_true.mean(0) * 0.8) * E
    model, kwargs_init, kwargs_up, kwargs_down, kwargs_fixed = setup_model(data, noisemap ** 2, s, cx + 0.1, cy - 0.1, k, initial_a)
    astrometric_prior = Prior(prior_analytic=[['c_x', cx + 0.1, np.array(M * [0.5])], ['c_y', cy - 0.1, np.array(M * [0.5])]])
    kwargs_fixed = deepcopy(kwargs_init)
    del kwargs_fixed['kwargs_analytic']['dx']
    del kwargs_fixed['kwargs_analytic']['dy']
    del kwargs_fixed['kwargs_analytic']['a']
    parameters = ParametersDeconv(kwargs_init=kwargs_init, kwargs_fixed=kwargs_fixed, kwargs_up=kwargs_up, kwargs_down=kwargs_down)
    loss = Loss(data, model, parameters, noisemap ** 2, prior=astrometric_prior, regularization_strength_flux_uniformity=10.0)
    optim = Optimizer(loss, parameters, method='l-bfgs-b')
    best_fit, logL_best_fit, extra_fields, runtime = optim.minimize(maxiter=60)
    kwargs_partial1 = deepcopy(parameters.best_fit_values(as_kwargs=True))
    assert np.abs(kwargs_partial1['kwargs_analytic']['dx'] - dxt).max() < 0.25           # epochs registered
    kwargs_fixed = deepcopy(kwargs_partial1)
    for grp, nm in (('kwargs_background', 'h'), ('kwargs_background', 'mean'), ('kwargs_analytic', 'a'), ('kwargs_analytic', 'c_x'),
                    ('kwargs_analytic', 'c_y'), ('kwargs_analytic', 'dx'), ('kwargs_analytic', 'dy')):
        del kwargs_fixed[grp][nm]
    W = propagate_noise(model, noisemap, kwargs_init, wavelet_type_list=['starlet'], method='SLIT', num_samples=500, seed=1,
                        likelihood_type='chi2', verbose=False, upsampling_factor=k)[0]
    parameters = ParametersDeconv(kwargs_init=kwargs_partial1, kwargs_fixed=kwargs_fixed, kwargs_up=kwargs_up, kwargs_down=kwargs_down)
    loss = Loss(data, model, parameters, noisemap ** 2, regularization_terms='l1_starlet', regularization_strength_scales=1.0,
                regularization_strength_hf=1.0, regularization_strength_positivity=100.0, regularization_strength_pts_source=0.01,
                regularization_strength_flux_uniformity=10.0, W=W, prior=astrometric_prior)
    optim = Optimizer(loss, parameters, method='adabelief')
    best_fit, logL_best_fit, extra_fields, runtime = optim.minimize(max_iterations=300, init_learning_rate=1e-4, schedule_learning_rate=False,
                                                                    restart_from_init=False, stop_at_loss_increase=False,
                                                                    progress_bar=True, return_param_history=True)
    kwargs_final = deepcopy(parameters.best_fit_values(as_kwargs=True))
    hist = np.asarray(extra_fields['loss_history'])
    assert hist.shape == (300,) and hist[-1] < hist[0] and np.isfinite(hist).all()
    a_fit = np.asarray(kwargs_final['kwargs_analytic']['a']).reshape(E, M)
    assert np.median(np.abs(a_fit - a_true) / a_true) < 0.1
    x_pixels = np.array(kwargs_final['kwargs_analytic']['c_x'] + kwargs_final['kwargs_analytic']['dx'][0])     # roi_modelling.py:339
    assert x_pixels.shape == (M,)
    res = data - np.array(model.model(kwargs_final))
    assert (np.nansum(res ** 2 / noisemap ** 2, axis=(1, 2)) / model.image_size ** 2 < 3).all()


def test_three_paths_agree_on_fluxes_at_2000_iterations(cuda_device, starred_installed):
    """One convention everywhere (D_k = block sum, amplitude == pixel-sum flux): at n_iter = 2000 the batched driver
    (star_photometry_batch, per-(frame, star) K2 fits), do_one_star_forward_modelling and the STARRED-shaped call sequence
    (setup_model / Loss / Optimizer, which couples the epochs through c and the clip norm) return the same fluxes in the same
    units -- a k^2 slip in any of them would show as a factor 4 -- and the fitted amplitude of a unit-max stamp is its pixel
    sum, the scale relation of the reference's notebook (example_roi_modelling.ipynb cells 13 -> 21 -> 36)."""
    from starred.deconvolution.deconvolution import setup_model
    from starred.deconvolution.loss import Loss
    from starred.deconvolution.parameters import ParametersDeconv
    from starred.optim.optimization import Optimizer
    from lightcurver_b200.processes.star_photometry import star_photometry_batch, do_one_star_forward_modelling
    from lightcurver_b200.utilities.starred_utilities import get_flux_uncertainties
    E, n, k, n_iter = 6, 16, 2, 2000
    d, data, noisemap, psf = _star_stack(E, n, k, seed=21)
    truth = d['transparency'] * d['star_flux'][0]
    # (1) batched driver: (F, S = 1, n, n)
    b = star_photometry_batch(data[:, None].astype(np.float32), noisemap[:, None].astype(np.float32), psf.astype(np.float32), k,
                              n_iter=n_iter)
    f_batch, s_batch = b['fluxes'][:, 0], b['fluxes_uncertainties'][:, 0]
    # (2) the reference-shaped single-star function (in-place scaling of its arguments, like the reference)
    d2, n2 = data.copy(), noisemap.copy()
    r = do_one_star_forward_modelling(d2, n2, psf, k, n_iter=n_iter, uniform_background_per_epoch=False, starlet_global_background=False)
    # (3) the STARRED call sequence with h and mean fixed
    scale = np.nanmax(data)
    ds, ns = data / scale, noisemap / scale
    model, kwargs_init, kwargs_up, kwargs_down, _ = setup_model(ds, ns ** 2, psf, np.array([0.]), np.array([0.]), k, list(np.nansum(ds, axis=(1, 2))))
    kwargs_fixed = {'kwargs_analytic': {'alpha': kwargs_init['kwargs_analytic']['alpha']},
                    'kwargs_background': {'h': kwargs_init['kwargs_background']['h'], 'mean': np.zeros(E)}, 'kwargs_sersic': {}}
    parameters = ParametersDeconv(kwargs_init=kwargs_init, kwargs_fixed=kwargs_fixed, kwargs_up=kwargs_up, kwargs_down=kwargs_down)
    loss = Loss(ds, model, parameters, ns ** 2, regularization_terms='l1_starlet', regularization_strength_scales=3.0,
                regularization_strength_hf=3.0, regularization_strength_flux_uniformity=0.)
    Optimizer(loss, parameters, method='adabelief').minimize(max_iterations=n_iter, init_learning_rate=1e-3, schedule_learning_rate=True,
                                                             restart_from_init=True)
    kw = parameters.best_fit_values(as_kwargs=True)
    f_api = scale * np.asarray(kw['kwargs_analytic']['a'])
    s_api = scale * get_flux_uncertainties(kw, kwargs_up, kwargs_down, ds, ns, model=model)
    print("[parity] fluxes at 2000 iterations: batched/one-star max rel diff", float(np.max(np.abs(f_batch / r['fluxes'] - 1))),
          "batched/starred-api", float(np.max(np.abs(f_batch / f_api - 1))), "vs truth (sigma units)",
          float(np.max(np.abs(f_batch - truth) / s_batch)))
    np.testing.assert_allclose(r['fluxes'], f_batch, rtol=2e-4)           # same K2 kernel behind both, same scale rule
    np.testing.assert_allclose(f_api, f_batch, rtol=2e-3)                 # coupled trajectory, same optimum
    np.testing.assert_allclose(r['fluxes_uncertainties'], s_batch, rtol=1e-3)
    np.testing.assert_allclose(s_api, s_batch, rtol=5e-3)
    assert np.all(np.abs(f_batch - truth) < 6 * s_batch)
    # the notebook's scale relation: amplitude of a stamp scaled to unit maximum ~ its pixel sum
    a_unit = np.asarray(kw['kwargs_analytic']['a'])
    np.testing.assert_allclose(a_unit, np.nansum(ds, axis=(1, 2)), rtol=0.1)
