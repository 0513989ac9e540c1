"""No-GPU checks of the STARRED-shaped front end: module registration under STARRED's names, the kwargs dict-of-dicts of
setup_model (star_photometry.py:66-87, roi_modelling.py:213-263), fixed / free bookkeeping of ParametersDeconv and Prior."""
import sys

import numpy as np
import pytest


def test_install_registers_starred_modules_and_refuses_to_shadow():
    import lightcurver_b200.starred_api as sa
    sa.install()
    try:
        from starred.deconvolution.deconvolution import setup_model
        from starred.deconvolution.loss import Loss, Prior                      # noqa: F401
        from starred.deconvolution.parameters import ParametersDeconv           # noqa: F401
        from starred.optim.optimization import Optimizer                        # noqa: F401
        from starred.optim.inference_base import FisherCovariance               # noqa: F401
        from starred.utils.noise_utils import propagate_noise                   # noqa: F401
        from starred.psf.psf import PSF, apply_distortion                       # noqa: F401
        from starred.procedures.psf_routines import build_psf                   # noqa: F401
        assert setup_model is sa.setup_model
        with pytest.raises(RuntimeError):
            sa.install()                                                        # never shadows an importable `starred`
        psf = np.ones((4, 4))
        assert apply_distortion(psf, {}, np.zeros(2)) is psf
        # with coefficients the resampling runs on the device: without one the library refuses (no CPU fallback)
        kd = {'dilation_x': np.zeros(2), 'dilation_y': np.zeros(2), 'shear': np.zeros(2)}
        from lightcurver_b200 import _lib
        if _lib.device_count() == 0:
            with pytest.raises(_lib.LcbError):
                apply_distortion(psf, kd, np.zeros(2))
    finally:
        sa.uninstall()
    assert not any(m == 'starred' or m.startswith('starred.') for m in sys.modules)


def test_setup_model_kwargs_and_parameter_bookkeeping():
    import lightcurver_b200.starred_api as sa
    E, n, k, M = 3, 8, 2, 2
    data, s = np.zeros((E, n, n)), np.ones((E, n * k, n * k))
    model, kw_init, kw_up, kw_down, kw_fixed = sa.setup_model(data, np.ones_like(data), s, np.array([1.0, -2.0]), np.array([0.5, 0.0]), k,
                                                               [3.0, 4.0] * E)
    assert model.image_size == n and model.epochs == E and model.M == M and model._cv.downsample_mean is False
    assert kw_init['kwargs_analytic']['a'].shape == (E * M,) and kw_init['kwargs_analytic']['dx'].shape == (E,)
    assert kw_init['kwargs_background']['h'].shape == ((n * k) ** 2,) and kw_init['kwargs_sersic'] == {}
    assert np.all(kw_down['kwargs_analytic']['a'] == 0) and np.all(np.isinf(kw_up['kwargs_analytic']['a']))
    assert set(kw_fixed['kwargs_analytic']) == {'alpha'}
    # star_photometry.py:74-87: alpha, h, mean fixed
    fixed = {'kwargs_analytic': {'alpha': kw_init['kwargs_analytic']['alpha']},
             'kwargs_background': {'h': kw_init['kwargs_background']['h'], 'mean': np.zeros(E)}, 'kwargs_sersic': {}}
    p = sa.ParametersDeconv(kwargs_init=kw_init, kwargs_fixed=fixed, kwargs_up=kw_up, kwargs_down=kw_down)
    assert p.free_flags() == dict(free_h=False, free_mean=False, free_a=True, free_c=True, free_d=True)
    del fixed['kwargs_background']['h']
    assert sa.ParametersDeconv(kw_init, fixed, kw_up, kw_down).free_flags()['free_h'] is True
    # roi_modelling.py:230-232: astrometry fixed at given values -> those values are the parameters
    fixed['kwargs_analytic']['c_x'] = np.array([9.0, 9.5]); fixed['kwargs_analytic']['c_y'] = np.array([1.0, 1.5])
    p = sa.ParametersDeconv(kw_init, fixed, kw_up, kw_down)
    assert p.free_flags()['free_c'] is False and np.all(p.best_fit_values(as_kwargs=True)['kwargs_analytic']['c_x'] == [9.0, 9.5])
    del fixed['kwargs_analytic']['c_y']
    with pytest.raises(NotImplementedError):
        sa.ParametersDeconv(kw_init, fixed, kw_up, kw_down).free_flags()
    # best_fit_values returns copies
    b = p.best_fit_values(as_kwargs=True)
    b['kwargs_analytic']['a'][:] = -1
    assert np.all(p.best_fit_values(as_kwargs=True)['kwargs_analytic']['a'] > 0)
    # Prior (roi_modelling.py:240-244)
    pr = sa.Prior(prior_analytic=[['c_x', np.array([1.0, -2.0]), np.array([0.3, 0.3])], ['c_y', np.array([0.5, 0.0]), np.array([0.4, 0.4])]])
    mux, sgx, muy, sgy = pr.as_tuple(M)
    assert np.all(mux == [1.0, -2.0]) and np.all(sgy == 0.4)
    assert sa.Prior().as_tuple(M) is None
    with pytest.raises(NotImplementedError):
        sa.Prior(prior_analytic=[['a', 1.0, 1.0]]).as_tuple(M)
