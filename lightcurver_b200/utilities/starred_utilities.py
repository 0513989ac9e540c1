"""``get_flux_uncertainties`` of lightcurver/utilities/starred_utilities.py:10-39.

The reference fixes everything but ``a``, runs 10 L-BFGS-B iterations and takes the diagonal Fisher
sigma.  The model is linear in ``a``, so the diagonal Fisher information is exactly
sum_p w_p (dm/da)^2 at the given (dx, dy) and the 10 iterations only touch the
value of ``a``, which is not returned; the kernel evaluates the closed form.
"""
import numpy as np

from .. import engine
from ..conventions import Conventions, DEFAULT, apply_to_library


def get_flux_uncertainties(kwargs, kwargs_up=None, kwargs_down=None, data=None, noisemap=None, model=None,
                           psf=None, subsampling_factor=None, conventions: Conventions = None):
    """Single point source form (star photometry).  ``model`` may be the dict returned by
    ``setup_model`` of this package (carrying psf and subsampling_factor) or None with explicit
    ``psf`` and ``subsampling_factor``.  Returns (E,) sigmas in the units of ``data``."""
    if model is not None and hasattr(model, '_s'):            # starred_api.Deconv
        psf = model._s if psf is None else psf
        subsampling_factor = model._upsampling_factor if subsampling_factor is None else subsampling_factor
    elif model is not None:
        psf = model['psf'] if psf is None else psf
        subsampling_factor = model['subsampling_factor'] if subsampling_factor is None else subsampling_factor
    if psf is None or subsampling_factor is None:
        raise ValueError("get_flux_uncertainties needs the PSFs and the subsampling factor")
    # conventions are explicit per call (never "whatever the last call left"): the model's own if it carries them, else
    # the argument, else the defaults
    cv = conventions if conventions is not None else (getattr(model, '_cv', None) or
                                                      (model.get('conventions') if isinstance(model, dict) else None) or DEFAULT)
    apply_to_library(cv)
    ka = kwargs['kwargs_analytic']
    E = data.shape[0]
    M = len(np.atleast_1d(ka['c_x']))
    if M != 1:
        from ..processes.roi_modelling import flux_sigma_multi
        return flux_sigma_multi(kwargs, data, noisemap, psf, subsampling_factor, cv)
    from ..processes.star_photometry import stamps_and_weights
    d32, weight = stamps_and_weights(data, noisemap)
    dx = np.asarray(ka['dx'], np.float32) + np.float32(np.atleast_1d(ka['c_x'])[0])
    dy = np.asarray(ka['dy'], np.float32) + np.float32(np.atleast_1d(ka['c_y'])[0])
    out = engine.phot_fit_batch(d32, weight,
                                np.asarray(psf, np.float32), np.arange(E, dtype=np.int32),
                                np.asarray(ka['a'], np.float32), int(subsampling_factor), 0, dx0=dx, dy0=dy,
                                want_residuals=False, want_loss_hist=False)
    return out['sigma_a']
