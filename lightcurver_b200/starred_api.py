"""A STARRED-shaped front end over liblcb: the ~12 symbols lightcurver imports from ``starred`` (SURVEY.md section 8,
row a8), with the call signatures, kwargs dict-of-dicts and return shapes its code relies on, so that the bodies of
``do_one_star_forward_modelling`` (star_photometry.py:23-151), ``get_flux_uncertainties`` (starred_utilities.py:10-39)
and ``do_modelling_of_roi`` (roi_modelling.py:213-335, 387-401) run UNCHANGED on the sm_100a kernels:

    import lightcurver_b200.starred_api as starred_api
    starred_api.install()          # registers starred, starred.deconvolution.deconvolution, ... in sys.modules
    import lightcurver             # its `from starred... import ...` lines now resolve to this module

Names and argument meaning: [V] where lightcurver's own call sites show them (cited per symbol), [R] otherwise.
Everything numerical happens in ``processes.roi_modelling.JointDeconvolution`` (the ``lcb_deconv_*`` handle API);
no array mathematics lives here apart from dictionary plumbing and scipy's L-BFGS-B driver.
"""
import sys
import time
import types
from copy import deepcopy

import numpy as np

from .conventions import Conventions, DEFAULT
from .processes.roi_modelling import JointDeconvolution, flux_sigma_multi
from .processes.star_photometry import point_source_image
from .procedures.psf_routines import build_psf          # starred.procedures.psf_routines.build_psf (psf_modelling.py:7,164)

import dataclasses

_GROUPS = {'kwargs_analytic': ('c_x', 'c_y', 'dx', 'dy', 'a', 'alpha'), 'kwargs_background': ('h', 'mean')}

# lightcurver hands pixel SUMS to setup_model as initial_a (star_photometry.py:55-69, roi_modelling.py:199-212) and reads
# `a` back as the flux (:128, :462): that is the D_k = block-sum normalisation, which is the library-wide default since
# round 2 (the block-mean alternative, amplitude k^2 times larger, stays switchable and is tested against the oracle).
STARRED_CONVENTIONS = DEFAULT


class Deconv:
    """starred.deconvolution.deconvolution.Deconv as lightcurver uses it: ``.model(kwargs)`` (star_photometry.py:124,
    roi_modelling.py:470), ``.getDeconvolved(kwargs, epoch)`` (:137, :387), ``.image_size`` (:127)."""

    def __init__(self, data, sigma_2, s, subsampling_factor, n_sources, conventions: Conventions = DEFAULT):
        self.epochs, self.image_size = int(data.shape[0]), int(data.shape[-1])
        self.M, self._upsampling_factor = int(n_sources), int(subsampling_factor)
        self.image_size_up = self.image_size * self._upsampling_factor
        self._s = np.ascontiguousarray(s, np.float32)
        self._cv = conventions
        self._engine_key, self._engine = None, None

    # one device handle per (data, weights): created lazily, reused by Loss / Optimizer / model evaluation
    def _jd(self, data, weight):
        key = (id(data), id(weight), data.shape)
        if self._engine is None or self._engine_key != key:
            if self._engine is not None:
                self._engine.close()
            self._engine = JointDeconvolution(np.nan_to_num(np.asarray(data, np.float32)), np.asarray(weight, np.float32),
                                              self._s, self._upsampling_factor, self.M, self._cv)
            self._engine_key = key
        return self._engine

    def _eval_engine(self):
        if self._engine is None:          # model(kwargs) before any Loss: data-free evaluation handle
            E, n = self.epochs, self.image_size
            self._engine = JointDeconvolution(np.zeros((E, n, n), np.float32), np.ones((E, n, n), np.float32), self._s,
                                              self._upsampling_factor, self.M, self._cv)
            self._engine_key = None
        return self._engine

    @staticmethod
    def _push(jd, kwargs, **free):
        ka, kb = kwargs['kwargs_analytic'], kwargs['kwargs_background']
        jd.set_params(h=np.asarray(kb['h']).reshape(-1), mean=kb['mean'], a=ka['a'], c_x=ka['c_x'], c_y=ka['c_y'],
                      dx=ka['dx'], dy=ka['dy'], alpha=ka['alpha'], **free)

    def model(self, kwargs):
        jd = self._eval_engine()
        self._push(jd, kwargs)
        return jd.get()['model']

    def getDeconvolved(self, kwargs, epoch=0):
        ka = kwargs['kwargs_analytic']
        n, k, nu = self.image_size, self._upsampling_factor, self.image_size_up
        h = np.asarray(kwargs['kwargs_background']['h'], np.float64).reshape(nu, nu)
        a = np.asarray(ka['a']).reshape(self.epochs, self.M)
        deconv = h.copy()
        for m in range(self.M):
            deconv += point_source_image(a[epoch, m], ka['c_x'][m] + ka['dx'][epoch], ka['c_y'][m] + ka['dy'][epoch], n, k, self._cv)
        return deconv, h


def setup_model(data, sigma_2, s, xs, ys, subsampling_factor, initial_a, conventions: Conventions = STARRED_CONVENTIONS):
    """starred.deconvolution.deconvolution.setup_model [V: star_photometry.py:66-69, roi_modelling.py:213-219]:
    returns (model, kwargs_init, kwargs_up, kwargs_down, kwargs_fixed); initial h = 0, mean = 0, dx = dy = 0, alpha = 0."""
    E, n = int(data.shape[0]), int(data.shape[-1])
    xs, ys = np.atleast_1d(np.asarray(xs, float)), np.atleast_1d(np.asarray(ys, float))
    M, nu = len(xs), n * int(subsampling_factor)
    model = Deconv(data, sigma_2, s, subsampling_factor, M, conventions)
    kwargs_init = {'kwargs_analytic': {'c_x': xs.copy(), 'c_y': ys.copy(), 'dx': np.zeros(E), 'dy': np.zeros(E),
                                       'a': np.asarray(initial_a, float).reshape(-1).copy(), 'alpha': np.zeros(E)},
                   'kwargs_background': {'h': np.zeros(nu * nu), 'mean': np.zeros(E)},
                   'kwargs_sersic': {}}
    inf = np.inf
    lim = {'c_x': n / 2.0, 'c_y': n / 2.0, 'dx': n / 2.0, 'dy': n / 2.0, 'alpha': np.pi}
    kwargs_up, kwargs_down = deepcopy(kwargs_init), deepcopy(kwargs_init)
    for grp, names in _GROUPS.items():
        for nm in names:
            shape = np.shape(kwargs_init[grp][nm])
            kwargs_up[grp][nm] = np.full(shape, lim.get(nm, inf))
            kwargs_down[grp][nm] = np.full(shape, 0.0 if nm == 'a' else -lim.get(nm, inf))
    kwargs_fixed = {'kwargs_analytic': {'alpha': kwargs_init['kwargs_analytic']['alpha'].copy()}, 'kwargs_background': {},
                    'kwargs_sersic': {}}
    return model, kwargs_init, kwargs_up, kwargs_down, kwargs_fixed


class ParametersDeconv:
    """starred.deconvolution.parameters.ParametersDeconv [V: star_photometry.py:89-92, roi_modelling.py:264-267]: a parameter
    is fixed when its key is present in kwargs_fixed (at the value given there)."""

    def __init__(self, kwargs_init, kwargs_fixed, kwargs_up=None, kwargs_down=None):
        self.kwargs_init, self.kwargs_fixed = deepcopy(kwargs_init), deepcopy(kwargs_fixed)
        self.kwargs_up, self.kwargs_down = kwargs_up, kwargs_down
        self._current = self.initial_values(as_kwargs=True)

    def _is_fixed(self, grp, nm):
        return nm in self.kwargs_fixed.get(grp, {})

    def initial_values(self, as_kwargs=True):
        kw = deepcopy(self.kwargs_init)
        for grp, names in _GROUPS.items():
            for nm in names:
                if self._is_fixed(grp, nm):
                    kw[grp][nm] = np.array(self.kwargs_fixed[grp][nm], dtype=float, copy=True)
        kw.setdefault('kwargs_sersic', {})
        return kw

    def free_flags(self):
        fx = self._is_fixed
        pairs = (('c_x', 'c_y'), ('dx', 'dy'))
        for a_, b_ in pairs:
            if fx('kwargs_analytic', a_) != fx('kwargs_analytic', b_):
                raise NotImplementedError(f"{a_} and {b_} must be fixed or free together")
        return dict(free_h=not fx('kwargs_background', 'h'), free_mean=not fx('kwargs_background', 'mean'),
                    free_a=not fx('kwargs_analytic', 'a'), free_c=not fx('kwargs_analytic', 'c_x'),
                    free_d=not fx('kwargs_analytic', 'dx'))

    def best_fit_values(self, as_kwargs=True):
        return deepcopy(self._current)


class Prior:
    """starred.deconvolution.loss.Prior [V: roi_modelling.py:240-244]: prior_analytic = [[name, mean, sigma], ...]."""

    def __init__(self, prior_analytic=None, prior_background=None, prior_sersic=None):
        self.prior_analytic = prior_analytic or []

    def as_tuple(self, M):
        mu = {'c_x': None, 'c_y': None}
        sg = {'c_x': None, 'c_y': None}
        for name, mean, sigma in self.prior_analytic:
            if name not in mu:
                raise NotImplementedError(f"Gaussian prior on '{name}' (supported: c_x, c_y)")
            mu[name], sg[name] = np.broadcast_to(np.asarray(mean, float), (M,)), np.broadcast_to(np.asarray(sigma, float), (M,))
        if mu['c_x'] is None and mu['c_y'] is None:
            return None
        big = np.full(M, 1e30)
        z = np.zeros(M)
        return (z if mu['c_x'] is None else mu['c_x'], big if sg['c_x'] is None else sg['c_x'],
                z if mu['c_y'] is None else mu['c_y'], big if sg['c_y'] is None else sg['c_y'])


class Loss:
    """starred.deconvolution.loss.Loss [V signature: star_photometry.py:95-111, roi_modelling.py:275-276, 313-321]."""

    def __init__(self, data, deconv_class, param_class, sigma_2, regularization_terms='l1_starlet',
                 regularization_strength_scales=1.0, regularization_strength_hf=1.0, regularization_strength_positivity=0.,
                 regularization_strength_positivity_ps=0., regularization_strength_pts_source=0.,
                 regularization_strength_flux_uniformity=0., W=None, prior=None):
        if regularization_terms != 'l1_starlet':
            raise NotImplementedError("regularization_terms must be 'l1_starlet'")
        if regularization_strength_positivity_ps:
            raise NotImplementedError("regularization_strength_positivity_ps")
        self.data, self.model, self.parameters = data, deconv_class, param_class
        with np.errstate(divide='ignore', invalid='ignore'):
            s2 = np.asarray(sigma_2, np.float64)
            self.weight = np.where(np.isfinite(s2) & (s2 > 0), 1.0 / s2, 0.0).astype(np.float32)
        self.reg = dict(lam_scales=regularization_strength_scales, lam_hf=regularization_strength_hf,
                        lam_pos=regularization_strength_positivity, lam_pts=regularization_strength_pts_source,
                        lam_fu=regularization_strength_flux_uniformity)
        self.W, self.prior = W, prior

    def engine(self):
        jd = self.model._jd(self.data, self.weight)
        W = None
        if self.W is not None:
            W = np.asarray(self.W, np.float32)[:jd.J]          # the coarsest plane, if present, is ignored
        prior = None if self.prior is None else self.prior.as_tuple(self.model.M)
        jd.set_reg(W=W, prior=prior, conventions=self.model._cv, **self.reg)
        return jd


class Optimizer:
    """starred.optim.optimization.Optimizer [V: star_photometry.py:113-122, roi_modelling.py:278-280, 326-334]:
    ``minimize(**opts)`` returns (best_fit, logL_best_fit, extra_fields, runtime); ``.loss_history``."""

    def __init__(self, loss, parameters, method='adabelief'):
        if method not in ('adabelief', 'l-bfgs-b'):
            raise NotImplementedError(f"optimiser '{method}' (supported: adabelief, l-bfgs-b)")
        self.loss, self.parameters, self.method = loss, parameters, method
        self.loss_history = []

    def _pull(self, jd):
        fin = jd.get(want_model=False)
        cur = self.parameters._current
        for nm in ('c_x', 'c_y', 'dx', 'dy', 'a', 'alpha'):
            cur['kwargs_analytic'][nm] = fin[nm].astype(np.float64)
        cur['kwargs_background']['h'] = fin['h'].astype(np.float64)
        cur['kwargs_background']['mean'] = fin['mean'].astype(np.float64)

    def minimize(self, max_iterations=None, min_iterations=None, init_learning_rate=1e-2, schedule_learning_rate=True,
                 restart_from_init=False, stop_at_loss_increase=False, progress_bar=False, return_param_history=False,
                 maxiter=None, **_ignored):
        t0 = time.time()
        jd = self.loss.engine()
        start = self.parameters.initial_values() if restart_from_init else self.parameters._current
        free = self.parameters.free_flags()
        Deconv._push(jd, start, **free)
        if self.method == 'adabelief':
            hist = jd.run(int(max_iterations), lr=float(init_learning_rate), schedule=bool(schedule_learning_rate))
        else:
            hist = _lbfgsb(jd, free, int(maxiter if maxiter is not None else (max_iterations or 100)), self.parameters)
        self._pull(jd)
        self.loss_history = [float(v) for v in np.asarray(hist)]
        extra = {'loss_history': np.asarray(hist)}
        final = float(hist[-1]) if len(hist) else float('nan')
        best = np.concatenate([np.ravel(self.parameters._current[g][nm]) for g, names in _GROUPS.items() for nm in names
                               if not self.parameters._is_fixed(g, nm)]) if any(free.values()) else np.zeros(0)
        return best, -final, extra, time.time() - t0


def _lbfgsb(jd, free, maxiter, parameters):
    """scipy L-BFGS-B over the free parameter groups, loss and gradient from the device (lcb_deconv_loss_grad)."""
    from scipy.optimize import minimize as sp_minimize
    E, M, nu2 = jd.E, jd.M, jd.nu * jd.nu
    layout = []
    if free['free_h']: layout.append(('h', nu2))
    if free['free_mean']: layout.append(('mean', E))
    if free['free_a']: layout.append(('a', E * M))
    if free['free_c']: layout += [('c_x', M), ('c_y', M)]
    if free['free_d']: layout += [('dx', E), ('dy', E)]
    cur = jd.get(want_model=False)
    x0 = np.concatenate([np.asarray(cur[nm], np.float64).reshape(-1) for nm, _ in layout]) if layout else np.zeros(0)
    if x0.size == 0:
        return np.asarray([jd.loss_grad()['loss']])
    bounds = []
    for nm, cnt in layout:
        grp = 'kwargs_background' if nm in ('h', 'mean') else 'kwargs_analytic'
        up = None if parameters.kwargs_up is None else np.ravel(parameters.kwargs_up[grp][nm])
        dn = None if parameters.kwargs_down is None else np.ravel(parameters.kwargs_down[grp][nm])
        for i in range(cnt):
            lo = None if dn is None or not np.isfinite(dn[i % len(dn)]) else float(dn[i % len(dn)])
            hi = None if up is None or not np.isfinite(up[i % len(up)]) else float(up[i % len(up)])
            bounds.append((lo, hi))
    hist, last = [], {}

    def fun(x):
        kw, o = {}, 0
        for nm, cnt in layout:
            kw[nm] = x[o:o + cnt]; o += cnt
        jd.set_params(**kw, **free)
        g = jd.loss_grad()
        last['L'] = g['loss']
        return g['loss'], np.concatenate([np.asarray(g[nm], np.float64).reshape(-1) for nm, _ in layout])

    res = sp_minimize(fun, x0, jac=True, method='L-BFGS-B', bounds=bounds,
                      options={'maxiter': maxiter, 'maxfun': 20 * maxiter + 20}, callback=lambda xk: hist.append(last['L']))
    fun(res.x)
    return np.asarray(hist if hist else [res.fun])


def propagate_noise(model, noisemap, kwargs, wavelet_type_list=('starlet',), method='SLIT', num_samples=200, seed=1,
                    likelihood_type='chi2', verbose=False, upsampling_factor=1, **_ignored):
    """starred.utils.noise_utils.propagate_noise [V: star_photometry.py:108-110, roi_modelling.py:299-301]: returns a list
    with one weight cube per wavelet type, (J + 1, nu, nu) with the coarsest plane last (ignored by Loss)."""
    if tuple(wavelet_type_list) != ('starlet',) or likelihood_type != 'chi2':
        raise NotImplementedError("propagate_noise: wavelet_type_list=['starlet'], likelihood_type='chi2'")
    if method != 'SLIT':
        raise NotImplementedError("propagate_noise for the deconvolution model: method='SLIT' (the 'MC' form exists for build_psf)")
    with np.errstate(divide='ignore', invalid='ignore'):
        nm = np.asarray(noisemap, np.float64)
        weight = np.where(np.isfinite(nm) & (nm > 0), 1.0 / nm ** 2, 0.0).astype(np.float32)
    E, n = weight.shape[0], weight.shape[-1]
    jd = JointDeconvolution(np.zeros((E, n, n), np.float32), weight, model._s, model._upsampling_factor, model.M, model._cv)
    try:
        Deconv._push(jd, kwargs)
        W = jd.noise_weights()
    finally:
        jd.close()
    return [np.concatenate([W, np.ones((1,) + W.shape[1:], W.dtype)])]


class FisherCovariance:
    """starred.optim.inference_base.FisherCovariance [V: starred_utilities.py:36-38], diagonal_only: with everything but
    ``a`` fixed the model is linear in a, sigma_a = (sum_p w (dm/da)^2)^-1/2."""

    def __init__(self, parameters, optim, diagonal_only=True):
        if not diagonal_only:
            raise NotImplementedError("FisherCovariance(diagonal_only=False)")
        self.parameters, self.optim, self._sigma = parameters, optim, None

    def compute_fisher_information(self, recompute=False):
        loss = self.optim.loss
        kw = self.parameters.best_fit_values(as_kwargs=True)
        with np.errstate(divide='ignore', invalid='ignore'):
            noisemap = np.where(loss.weight > 0, 1.0 / np.sqrt(loss.weight.astype(np.float64)), np.inf)
        self._sigma = flux_sigma_multi(kw, np.asarray(loss.data), noisemap, loss.model._s, loss.model._upsampling_factor, loss.model._cv)

    def get_kwargs_sigma(self):
        if self._sigma is None:
            self.compute_fisher_information()
        kw = self.parameters.best_fit_values(as_kwargs=True)
        out = {g: {nm: np.zeros_like(np.asarray(kw[g][nm], float)) for nm in names} for g, names in _GROUPS.items()}
        out['kwargs_analytic']['a'] = np.asarray(self._sigma)
        out['kwargs_sersic'] = {}
        return out


class PSF:
    """starred.psf.psf.PSF is only imported by lightcurver for type reasons (star_photometry.py:12); the fit itself goes
    through build_psf."""

    def __init__(self, *a, **k):
        raise NotImplementedError("use lightcurver_b200.procedures.psf_routines.build_psf")


def distortion_theta(kwargs_distortion):
    """kwargs_distortion {dilation_x, dilation_y, shear} (2 coefficients each) -> the (6,) vector of the C ABI; {} -> None."""
    if not kwargs_distortion:
        return None
    return np.concatenate([np.asarray(kwargs_distortion[key], np.float32).reshape(2) for key in ('dilation_x', 'dilation_y', 'shear')])


def apply_distortion(narrow_psf, kwargs_distortion, star_xy_coordinates, conventions: Conventions = STARRED_CONVENTIONS):
    """starred.psf.psf.apply_distortion (star_photometry.py:303, roi_file_preparation.py:179): the narrow PSF seen at the rescaled
    frame position ``star_xy_coordinates`` (utilities/image_coordinates.py:4-25) under ``kwargs_distortion`` (the dict build_psf
    returned and psf_modelling.py:200-202 stored).  Runs on the device (lcb_apply_distortion_batch); an empty kwargs_distortion
    (field_distortion was off) returns the PSF unchanged."""
    theta = distortion_theta(kwargs_distortion)
    if theta is None:
        return narrow_psf
    from . import engine
    psf = np.ascontiguousarray(narrow_psf, np.float32)
    out = engine.apply_distortion_batch(psf[None], theta[None], np.zeros(1, np.int32),
                                        np.asarray(star_xy_coordinates, np.float32).reshape(1, 2), mode=conventions.distortion_mode())
    return out[0]


def install():
    """Registers this front end under STARRED's module names (only if the real package is absent)."""
    import importlib.util
    if 'starred' in sys.modules or importlib.util.find_spec('starred') is not None:
        raise RuntimeError("a 'starred' package is importable: refusing to shadow it")
    me = sys.modules[__name__]
    tree = {
        'starred': {},
        'starred.deconvolution': {},
        'starred.deconvolution.deconvolution': dict(setup_model=setup_model, Deconv=Deconv),
        'starred.deconvolution.loss': dict(Loss=Loss, Prior=Prior),
        'starred.deconvolution.parameters': dict(ParametersDeconv=ParametersDeconv),
        'starred.optim': {},
        'starred.optim.optimization': dict(Optimizer=Optimizer),
        'starred.optim.inference_base': dict(FisherCovariance=FisherCovariance),
        'starred.utils': {},
        'starred.utils.noise_utils': dict(propagate_noise=propagate_noise),
        'starred.psf': {},
        'starred.psf.psf': dict(PSF=PSF, apply_distortion=apply_distortion),
        'starred.procedures': {},
        'starred.procedures.psf_routines': dict(build_psf=build_psf),
    }
    for name, symbols in tree.items():
        mod = types.ModuleType(name)
        mod.__dict__.update(symbols)
        mod.__lcb_shim__ = me
        sys.modules[name] = mod
    for name in tree:
        if '.' in name:
            parent, child = name.rsplit('.', 1)
            setattr(sys.modules[parent], child, sys.modules[name])
    return sys.modules['starred']


def uninstall():
    for name in [m for m in sys.modules if m == 'starred' or m.startswith('starred.')]:
        if getattr(sys.modules[name], '__lcb_shim__', None) is not None:
            del sys.modules[name]
