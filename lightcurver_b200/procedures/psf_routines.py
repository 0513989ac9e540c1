"""``build_psf``: the call lightcurver makes at lightcurver/processes/psf_modelling.py:164-171
(``starred.procedures.psf_routines.build_psf``), served by the sm_100a kernels of liblcb.

``build_psf`` keeps the reference's signature, argument meaning and result keys (those consumed at
psf_modelling.py:177-210 and pinned by tests/test_starred_calls/test_starred_calls.py:66-80) for ONE
frame; ``build_psf_batch`` is the batched form used by the patched ``model_all_psfs`` driver: gather
every pending frame, ONE library call, scatter the per-frame dicts.
"""
import numpy as np

from .. import engine
from ..conventions import Conventions, DEFAULT


def build_psf_batch(images, noisemaps, subsampling_factor, masks=None, n_iter_analytic=40,
                    n_iter_adabelief=2000, guess_method_star_position='barycenter', guess_fwhm_pixels=3.,
                    field_distortion=False, stamp_coordinates=None, regularization_strength_scales=None,
                    regularization_strength_hf=None, adabelief_learning_rate=None,
                    conventions: Conventions = DEFAULT, return_dicts=True, noise_propagation='SLIT', noise_samples=100,
                    noise_seed=1, devices=None, star_counts=None):
    """Fits F frames in one library call.

    images / noisemaps / masks: sequences (length F) of arrays (N_f, n, n) -- N_f may differ per
    frame (psf_modelling.py:144-153 drops stars per frame).  guess_fwhm_pixels: scalar or (F,).
    noise_propagation: 'SLIT' (deterministic diagonal propagation, default) or 'MC' (``noise_samples`` noise draws, the
    method STARRED's build_psf passes to propagate_noise [R]); both give the starlet-space weights W of stage 2.
    devices: None (current CUDA device), 'all', a count or a list of device indices: the frames are split into contiguous
    blocks balanced by their star counts, one host thread per GPU, no collective (frames are independent).
    star_counts: when given, ``images`` / ``noisemaps`` / ``masks`` are already CONCATENATED arrays (sumN, n, n) -- e.g. views
    of the page-locked staging buffer of ``stamp_store`` -- and star_counts[f] is the number of stars of frame f.
    field_distortion (psf_modelling.py:169-170): every star sees the affine resampling of the frame's narrow PSF given by
    ``kwargs_distortion`` = {dilation_x, dilation_y, shear} (each a first-order polynomial, 2 coefficients, in the rescaled frame
    position) at ``stamp_coordinates`` -- a sequence (length F) of (N_f, 2) arrays from ``rescale_image_coordinates``
    (utilities/image_coordinates.py:4-25), or one (sumN, 2) array with ``star_counts``; stage 2 fits the six coefficients.
    Returns a list of per-frame result dicts shaped like STARRED's (``return_dicts``), or the raw
    batched arrays.
    """
    devs = engine.resolve_devices(devices)
    n_frames = len(images) if star_counts is None else len(star_counts)
    if star_counts is not None and len(devs) > 1 and n_frames > 1:      # per-device blocks need per-frame slices
        offs = np.concatenate([[0], np.cumsum(np.asarray(star_counts, np.int64))])
        split = lambda a: None if a is None else [a[offs[f]:offs[f + 1]] for f in range(n_frames)]
        images, noisemaps, masks, star_counts = split(images), split(noisemaps), split(masks), None
        stamp_coordinates = split(None if stamp_coordinates is None else np.asarray(stamp_coordinates))
    if len(devs) > 1 and n_frames > 1:
        counts = [int(np.shape(im)[0]) for im in images]
        blocks = engine.split_by_work(counts, len(devs))
        fw = np.broadcast_to(np.asarray(3.0 if guess_fwhm_pixels is None else guess_fwhm_pixels, dtype=np.float64), (len(images),))
        kw = dict(n_iter_analytic=n_iter_analytic, n_iter_adabelief=n_iter_adabelief,
                  guess_method_star_position=guess_method_star_position, field_distortion=field_distortion,
                  regularization_strength_scales=regularization_strength_scales, regularization_strength_hf=regularization_strength_hf,
                  adabelief_learning_rate=adabelief_learning_rate, conventions=conventions, return_dicts=return_dicts,
                  noise_propagation=noise_propagation, noise_samples=noise_samples, noise_seed=noise_seed, devices=None)

        def one(lo, hi):
            return build_psf_batch(images[lo:hi], noisemaps[lo:hi], subsampling_factor, None if masks is None else masks[lo:hi],
                                   guess_fwhm_pixels=fw[lo:hi],
                                   stamp_coordinates=None if stamp_coordinates is None else stamp_coordinates[lo:hi], **kw)
        parts = engine.fan_out(blocks, devs, one)
        if return_dicts:
            return [r for part in parts for r in part]
        out = {kk: np.concatenate([part[kk] for part in parts]) for kk in parts[0] if kk != 'star_off'}
        off = np.zeros(len(images) + 1, np.int32)
        off[1:] = np.cumsum(counts)
        out['star_off'] = off
        return out
    if field_distortion and stamp_coordinates is None:
        raise ValueError("field_distortion=True needs stamp_coordinates (psf_modelling.py:122-124, 170)")
    cv = conventions
    from ..conventions import apply_to_library
    apply_to_library(cv)                       # the kernels read the library-wide conventions at call time
    F = n_frames
    k = int(subsampling_factor)
    counts = [int(c) for c in star_counts] if star_counts is not None else [int(np.shape(im)[0]) for im in images]
    if F == 0:
        return []
    if min(counts) < 1:
        raise ValueError("every frame needs at least one star (psf_modelling.py:154-160 skips empty frames)")
    n = int(np.shape(images)[-1]) if star_counts is not None else int(np.shape(images[0])[-1])
    off = np.zeros(F + 1, np.int32)
    off[1:] = np.cumsum(counts)
    sumN = int(off[-1])
    def cat(seq):
        if isinstance(seq, np.ndarray):            # already concatenated (star_counts form): no copy
            return np.asarray(seq).reshape(sumN, n, n)
        return np.concatenate([np.asarray(x) for x in seq])
    # normalisation (A.4), NaN / mask policy, weights and the smart guess run on the device (lcb_psf_prepare_batch):
    # the raw arrays are uploaded once (asynchronously when they are pinned) and stay there for the fit
    prep = engine.psf_prepare_batch(cat(images), cat(noisemaps), None if masks is None else cat(masks), off, k,
                                    norm_scale=cv.psf_norm_scale, downsample_mean=cv.downsample_mean,
                                    guess_method=guess_method_star_position)
    data, weight, a0, x00, y00 = prep['data'], prep['weight'], prep['a0'], prep['x0'], prep['y0']
    fwhm = np.broadcast_to(np.asarray(3.0 if guess_fwhm_pixels is None else guess_fwhm_pixels, dtype=np.float64), (F,))
    moffat0 = np.stack([fwhm, fwhm, np.zeros(F), np.full(F, cv.moffat_beta_init), np.ones(F)], -1)
    lam_s = cv.psf_lambda_scales if regularization_strength_scales is None else regularization_strength_scales
    lam_h = cv.psf_lambda_hf if regularization_strength_hf is None else regularization_strength_hf
    dist = {}
    if field_distortion:
        xy = np.asarray(stamp_coordinates, np.float32) if isinstance(stamp_coordinates, np.ndarray) else \
            np.concatenate([np.asarray(c, np.float32).reshape(-1, 2) for c in stamp_coordinates])
        if xy.shape != (sumN, 2):
            raise ValueError(f"stamp_coordinates must hold one (x, y) pair per star: expected ({sumN}, 2), got {xy.shape}")
        dist = dict(field_distortion=cv.distortion_mode(), stamp_xy=xy)
    out = engine.psf_fit_batch(
        data, weight, prep['star_off'], k, moffat0, a0, x00, y00, n_iter_analytic=n_iter_analytic,
        n_iter_adabelief=n_iter_adabelief, lr=cv.psf_stage2_lr if adabelief_learning_rate is None else adabelief_learning_rate,
        lam_scales=lam_s, lam_hf=lam_h, noise_weights=noise_propagation, mc_samples=noise_samples, mc_seed=noise_seed,
        bounds=dict(fwhm_min=cv.moffat_fwhm_min, fwhm_max=n / 2.0, beta_min=cv.moffat_beta_min, beta_max=cv.moffat_beta_max),
        want=('narrow_psf', 'full_psf', 'residuals', 'chi2', 'loss_hist', 'loss_hist_analytic', 'status'), **dist)
    out = engine.to_host(out)                  # one device -> host copy per product, through page-locked staging
    norms = prep['norm'].cpu().numpy().astype(np.float64)
    out['norms'] = norms
    out['star_off'] = off
    if not return_dicts:
        return out
    res = []
    nu = n * k
    for f in range(F):
        sl = slice(off[f], off[f + 1])
        mo = out['moffat'][f]
        res.append({
            'full_psf': out['full_psf'][f],
            'narrow_psf': out['narrow_psf'][f],
            'chi2': float(out['chi2'][f]),
            'residuals': out['residuals'][sl] * norms[f],
            'kwargs_psf': {
                'kwargs_moffat': {'fwhm_x': np.array([mo[0]]), 'fwhm_y': np.array([mo[1]]), 'phi': np.array([mo[2]]),
                                  'beta': np.array([mo[3]]), 'C': np.array([mo[4]])},
                'kwargs_gaussian': {'a': out['a'][sl], 'x0': out['x0'][sl], 'y0': out['y0'][sl]},
                'kwargs_background': {'background': out['background'][f].reshape(nu * nu)},
                # empty without field distortion (psf_modelling.py:200-202 iterates over the items)
                'kwargs_distortion': ({'dilation_x': out['distortion'][f, 0:2].copy(), 'dilation_y': out['distortion'][f, 2:4].copy(),
                                       'shear': out['distortion'][f, 4:6].copy()} if field_distortion else {}),
            },
            'adabelief_extra_fields': {'loss_history': out['loss_hist'][f]},
            'analytical_extra_fields': {'loss_history': out['loss_hist_analytic'][f]},
            'norm': float(norms[f]),
            'status': int(out['status'][f]),
        })
    return res


def build_psf(image, noisemap, subsampling_factor, masks=None, n_iter_analytic=40, n_iter_adabelief=2000,
              guess_method_star_position='barycenter', guess_fwhm_pixels=3., field_distortion=False,
              stamp_coordinates=None, **kwargs):
    """One frame; same call as psf_modelling.py:164-171.  image, noisemap, masks: (N, n, n)."""
    return build_psf_batch([image], [noisemap], subsampling_factor, None if masks is None else [masks],
                           n_iter_analytic=n_iter_analytic, n_iter_adabelief=n_iter_adabelief,
                           guess_method_star_position=guess_method_star_position,
                           guess_fwhm_pixels=guess_fwhm_pixels, field_distortion=field_distortion,
                           stamp_coordinates=stamp_coordinates, **kwargs)[0]
