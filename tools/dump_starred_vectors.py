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
    print('wrote', sorted(f.name for f in out.glob('starred_*.npz')))


if __name__ == '__main__':
    main()
