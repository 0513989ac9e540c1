#!/usr/bin/env python
"""Golden vectors from REAL STARRED (run this where `starred-astro` and jax are installed).

The build container has neither STARRED nor jax (no network), so parity in this repository is pinned on a
restatement (oracle/).  This script closes the gap: it calls the same STARRED entry points lightcurver
calls (psf_modelling.py:164-171, star_photometry.py:66-128) on the small seeded inputs of
tools/make_golden.py and writes `tests/golden/starred_*.npz` with the inputs, the loss, its gradient and the
fitted parameters.  Drop the files into tests/golden/; tests/test_starred_golden.py (skipped when the files
are absent) then compares the CUDA path against them and tells which field of
lightcurver_b200.conventions.Conventions must flip.
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    try:
        import jax
        import jax.numpy as jnp
        from starred.psf.psf import PSF
        from starred.psf.loss import Loss as PsfLoss
        from starred.psf.parameters import ParametersPSF
        from starred.deconvolution.deconvolution import setup_model
        from starred.deconvolution.loss import Loss as DeconvLoss
        from starred.deconvolution.parameters import ParametersDeconv
        from starred.procedures.psf_routines import build_psf
    except ImportError as e:      # pragma: no cover
        raise SystemExit(f"STARRED/JAX not importable here ({e}); run this script where starred-astro is installed")
    from lightcurver_b200 import synthetic
    out = ROOT / 'tests' / 'golden'
    out.mkdir(parents=True, exist_ok=True)

    # --- PSF: build_psf on 1 frame x 3 stars x 16x16, k = 2 (the call of psf_modelling.py:164-171)
    n, k, N = 16, 2, 3
    d = synthetic.make_psf_frames(1, N, n, k, seed=43)
    res = build_psf(image=d['data'][0], noisemap=d['noisemap'][0], subsampling_factor=k, n_iter_analytic=50,
                    n_iter_adabelief=100, masks=d['masks'][0], guess_method_star_position='center',
                    guess_fwhm_pixels=float(d['fwhm'][0]))
    kw = res['kwargs_psf']
    np.savez_compressed(out / 'starred_build_psf_n16_k2.npz', kind='starred_build_psf', n=n, k=k,
                        data=d['data'][0], noisemap=d['noisemap'][0], masks=d['masks'][0], fwhm_guess=d['fwhm'][0],
                        narrow_psf=np.asarray(res['narrow_psf']), full_psf=np.asarray(res['full_psf']),
                        chi2=float(res['chi2']), residuals=np.asarray(res['residuals']),
                        loss_history=np.asarray(res['adabelief_extra_fields']['loss_history']),
                        **{f'moffat_{kk}': np.asarray(v) for kk, v in kw['kwargs_moffat'].items()},
                        **{f'gaussian_{kk}': np.asarray(v) for kk, v in kw['kwargs_gaussian'].items()},
                        background=np.asarray(kw['kwargs_background']['background']))

    # --- field distortion: build_psf(field_distortion=True, stamp_coordinates=...) (psf_modelling.py:169-170) and apply_distortion
    #     (star_photometry.py:303) on the same frame: pins the parametrisation recalled in csrc/lcb_distort.cuh (which keys
    #     kwargs_distortion holds, the order of its polynomial, the resampling, the determinant factor)
    try:
        from starred.psf.psf import apply_distortion
        xy = np.array([[-0.3, 0.2], [0.1, -0.4], [0.4, 0.4]])
        resd = build_psf(image=d['data'][0], noisemap=d['noisemap'][0], subsampling_factor=k, n_iter_analytic=50,
                         n_iter_adabelief=100, masks=d['masks'][0], guess_method_star_position='center',
                         guess_fwhm_pixels=float(d['fwhm'][0]), field_distortion=True, stamp_coordinates=xy)
        kd = {kk: np.asarray(v) for kk, v in resd['kwargs_psf']['kwargs_distortion'].items()}
        probe = {kk: np.full_like(v, 0.03) for kk, v in kd.items()}          # a fixed, non-trivial distortion for the resampling itself
        np.savez_compressed(out / 'starred_distortion_n16_k2.npz', kind='starred_distortion', n=n, k=k, data=d['data'][0],
                            noisemap=d['noisemap'][0], masks=d['masks'][0], fwhm_guess=d['fwhm'][0], stamp_coordinates=xy,
                            narrow_psf=np.asarray(resd['narrow_psf']),
                            distorted_fitted=np.stack([np.asarray(apply_distortion(resd['narrow_psf'], kd, p_)) for p_ in xy]),
                            distorted_probe=np.stack([np.asarray(apply_distortion(resd['narrow_psf'], probe, p_)) for p_ in xy]),
                            **{f'distortion_{kk}': v for kk, v in kd.items()})
    except Exception as exc:      # pragma: no cover - depends on the installed STARRED version
        print('field distortion vectors skipped:', exc)

    # --- photometry: loss and gradient of the deconvolution Loss at fixed parameters (star_photometry.py:66-111)
    F, S = 2, 3
    p = synthetic.make_phot_frames(F, S, n, k, seed=42)
    data = p['data'][:, 0].astype(np.float64)
    nm = p['noisemap'][:, 0].astype(np.float64)
    scale = data.max()
    data, nm = data / scale, nm / scale
    model, kwargs_init, kwargs_up, kwargs_down, kwargs_fixed = setup_model(data, nm ** 2, p['psf'].astype(np.float64),
                                                                           np.array([0.]), np.array([0.]), k,
                                                                           list(data.sum((1, 2))))
    kwargs_fixed = {'kwargs_analytic': {'alpha': kwargs_init['kwargs_analytic']['alpha'],
                                        'c_x': kwargs_init['kwargs_analytic']['c_x'], 'c_y': kwargs_init['kwargs_analytic']['c_y']},
                    'kwargs_background': {'h': kwargs_init['kwargs_background']['h'], 'mean': np.zeros(F)}, 'kwargs_sersic': {}}
    params = ParametersDeconv(kwargs_init=kwargs_init, kwargs_fixed=kwargs_fixed, kwargs_up=kwargs_up, kwargs_down=kwargs_down)
    loss = DeconvLoss(data, model, params, nm ** 2, regularization_terms='l1_starlet', regularization_strength_scales=0.,
                      regularization_strength_hf=0.)
    x0 = params.kwargs2args(kwargs_init)
    val, grad = jax.value_and_grad(loss)(jnp.asarray(x0))
    np.savez_compressed(out / 'starred_phot_n16_k2.npz', kind='starred_phot', n=n, k=k, data=data, noisemap=nm, psf=p['psf'],
                        args=np.asarray(x0), loss=float(val), grad=np.asarray(grad), model=np.asarray(model.model(kwargs_init)),
                        a=np.asarray(kwargs_init['kwargs_analytic']['a']))
    # --- joint deconvolution: every term of the Loss separately at a non-trivial point (roi_modelling.py:275-321), and the
    #     SLIT weights (roi_modelling.py:299).  tests/test_starred_golden.py tells which Conventions field matches.
    from starred.utils.noise_utils import propagate_noise
    rng = np.random.default_rng(7)
    E, M = 3, 2
    d3 = rng.normal(0.0, 0.02, (E, n, n)) + 0.05
    nm3 = np.full((E, n, n), 0.02)
    xs, ys = np.array([-2.0, 2.5]), np.array([1.0, -1.5])
    a3 = rng.uniform(1.0, 2.0, E * M)
    model, kwargs_init, kwargs_up, kwargs_down, kwargs_fixed = setup_model(d3, nm3 ** 2, p['psf'][:1].repeat(E, 0).astype(np.float64),
                                                                           xs, ys, k, list(a3))
    kw = {'kwargs_analytic': dict(kwargs_init['kwargs_analytic']), 'kwargs_background': dict(kwargs_init['kwargs_background']),
          'kwargs_sersic': {}}
    kw['kwargs_analytic']['dx'] = rng.uniform(-0.5, 0.5, E); kw['kwargs_analytic']['dy'] = rng.uniform(-0.5, 0.5, E)
    kw['kwargs_background']['h'] = 0.01 * rng.standard_normal((n * k) ** 2)
    kw['kwargs_background']['mean'] = rng.uniform(-0.01, 0.01, E)
    fixed = {'kwargs_analytic': {'alpha': kwargs_init['kwargs_analytic']['alpha']}, 'kwargs_background': {}, 'kwargs_sersic': {}}
    params = ParametersDeconv(kwargs_init=kw, kwargs_fixed=fixed, kwargs_up=kwargs_up, kwargs_down=kwargs_down)
    W = np.asarray(propagate_noise(model, nm3, kwargs_init, wavelet_type_list=['starlet'], method='SLIT', num_samples=100, seed=1,
                                   likelihood_type='chi2', verbose=False, upsampling_factor=k)[0])
    x = jnp.asarray(params.kwargs2args(kw))
    terms = {}
    base = dict(regularization_terms='l1_starlet', regularization_strength_scales=0., regularization_strength_hf=0.,
                regularization_strength_positivity=0., regularization_strength_pts_source=0., regularization_strength_flux_uniformity=0., W=W)
    for name, over in (('chi2', {}), ('starlet_scales', dict(regularization_strength_scales=1.0)), ('starlet_hf', dict(regularization_strength_hf=1.0)),
                       ('positivity', dict(regularization_strength_positivity=100.0)), ('pts_source', dict(regularization_strength_pts_source=0.5)),
                       ('flux_uniformity', dict(regularization_strength_flux_uniformity=5.0))):
        lo = DeconvLoss(d3, model, params, nm3 ** 2, **{**base, **over})
        v, g = jax.value_and_grad(lo)(x)
        terms['loss_' + name] = float(v); terms['grad_' + name] = np.asarray(g)
    np.savez_compressed(out / 'starred_deconv_terms_n16_k2.npz', kind='starred_deconv_terms', n=n, k=k, E=E, M=M, data=d3, noisemap=nm3,
                        psf=p['psf'][:1].repeat(E, 0), xs=xs, ys=ys, W=W, args=np.asarray(x), model=np.asarray(model.model(kw)),
                        **{f'kw_{g_}_{kk}': np.asarray(v) for g_ in ('kwargs_analytic', 'kwargs_background') for kk, v in kw[g_].items()}, **terms)
    print('wrote', sorted(f.name for f in out.glob('starred_*.npz')))


if __name__ == '__main__':
    main()
