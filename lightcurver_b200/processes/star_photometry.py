"""``do_one_star_forward_modelling`` of lightcurver/processes/star_photometry.py:23-151, served by
liblcb (K2 for the default pipeline configuration, the joint-deconvolution engine when a shared
starlet background or per-epoch constants are requested).

Same signature, same in-place rescaling of the caller's arrays (:47-49), same result keys (:139-150)
and types (tests/test_starred_calls/test_starred_calls.py:20-64).  In the north-star formulation the
epochs of one star are independent fits (c fixed at the stamp centre, per-epoch gradient clip; see
DESIGN.md "reference-coupled mode"), which is what makes frames shard across GPUs with no collective.
"""
import logging
import math

import numpy as np

from .. import engine
from ..conventions import Conventions, DEFAULT


def _initial_flux_guess(data):
    """star_photometry.py:55-64: sum of pixels minus n^2 * (mean over the 4 edges AND all epochs of the
    edge medians) -- one scalar background for the whole stack."""
    with np.errstate(all='ignore'):
        background_values = np.nanmean([
            np.nanmedian(data[:, :1, :], axis=(1, 2)),
            np.nanmedian(data[:, :, :1], axis=(1, 2)),
            np.nanmedian(data[:, -1:, :], axis=(1, 2)),
            np.nanmedian(data[:, :, -1:], axis=(1, 2))
        ])
    background_values = np.nan_to_num(background_values, nan=0)
    return np.nansum(data, axis=(1, 2)) - data[0].size * background_values


def stamps_and_weights(data, noisemap):
    """The ONE weight rule of every host path of this package (the device rule of k_phot_prep_item is the same): a pixel
    counts with weight 1/sigma^2 only where the data AND the noise are finite and the noise is positive; everywhere else
    the stamp value is 0 and the weight 0 (the reference's `data 0, noise 1e7` for doubly-NaN pixels, star_photometry.py:
    309-311, is the same thing to 1e-14 of a typical weight; a pixel where only the data is NaN would make STARRED's
    loss NaN, here it is ignored).  Returns (float32 stamps, float32 weights); the inputs are not modified."""
    d = np.asarray(data)
    nm = np.asarray(noisemap, dtype=np.float64)
    ok = np.isfinite(d) & np.isfinite(nm) & (nm > 0)
    with np.errstate(divide='ignore', invalid='ignore', over='ignore'):
        weight = np.where(ok, 1.0 / nm ** 2, 0.0)
    weight = np.where(np.isfinite(weight), weight, 0.0)
    return np.where(ok, d, 0.0).astype(np.float32), weight.astype(np.float32)


def point_source_image(a, x, y, n, k, cv: Conventions = DEFAULT):
    """a * G(.; k x, k y) on the nu x nu grid (what model.getDeconvolved shows for a point source)."""
    nu = n * k
    sig = cv.gauss_fwhm_up / (2.0 * math.sqrt(2.0 * math.log(2.0)))
    u = np.arange(nu) - (nu - 1) / 2.0
    gx = np.exp(-(u - k * x) ** 2 / (2 * sig * sig)) / (math.sqrt(2 * math.pi) * sig)
    gy = np.exp(-(u - k * y) ** 2 / (2 * sig * sig)) / (math.sqrt(2 * math.pi) * sig)
    return a * np.outer(gy, gx)


COUPLED_FIT = dict(lr=1e-3, schedule=True, free_c=True, regularization_strength_scales=3.0, regularization_strength_hf=3.0,
                   regularization_strength_positivity=0.0)     # star_photometry.py:76-122 [R]


def _scale_and_guess(data, noisemap, k, cv):
    """star_photometry.py:47-64 -- IN PLACE on the caller's arrays, like the reference: divide the stack by its nanmax, then the
    initial flux guess.  Returns (scale, initial amplitudes, float32 stamps, float32 weights)."""
    scale = float(np.nanmax(data))
    data /= scale
    noisemap /= scale
    a_est = _initial_flux_guess(data) * cv.amplitude_per_flux(k)
    d32, weight = stamps_and_weights(data, noisemap)
    return scale, a_est, d32, weight


def _coupled_result(res, scale, data, noisemap, n, k, cv):
    """Result dict of do_one_star_forward_modelling (:124-150) from a finished joint fit with M = 1."""
    kw = res['kwargs_final']
    a = np.asarray(kw['kwargs_analytic']['a'])
    deconv, bkg = res['deconvolved_epoch0']
    return _result_dict(scale, kw, a, res['flux_sigma'], res['loss_history'], data - res['model'], noisemap, deconv, bkg, n, k, cv)


def _result_dict(scale, kw, a, sigma, loss_curve, residuals, noisemap, deconv, bkg, n, k, cv):
    sigma_2 = noisemap ** 2
    with np.errstate(divide='ignore', invalid='ignore'):
        chi2_per_frame = np.nansum(residuals ** 2 / sigma_2, axis=(1, 2)) / n ** 2   # :127 (image_size^2, not dof)
    chi2 = np.nanmean(chi2_per_frame)
    return {
        'scale': scale,
        'kwargs_final': kw,
        # fluxes are reported as pixel sums (the unit of lightcurver's star_flux_in_frame.flux, star_photometry.py:128) whatever the
        # normalisation of D_k in the kernels: with the block mean the amplitude a is k^2 x the flux
        'fluxes': scale * a / cv.amplitude_per_flux(k),
        'fluxes_uncertainties': scale * np.asarray(sigma) / cv.amplitude_per_flux(k),
        'chi2': float(chi2),
        'chi2_per_frame': np.array(chi2_per_frame),
        'loss_curve': list(np.asarray(loss_curve)),
        'residuals': scale * residuals,
        'deconvolved_image': scale * np.asarray(deconv),
        'starlet_background': scale * np.asarray(bkg),
    }


def do_one_star_forward_modelling(data, noisemap, psf, subsampling_factor, n_iter=2000,
                                  uniform_background_per_epoch=False, starlet_global_background=True,
                                  conventions: Conventions = DEFAULT):
    """See lightcurver/processes/star_photometry.py:23-151.  data, noisemap (E,n,n); psf (E,n*k,n*k)."""
    cv = conventions
    from ..conventions import apply_to_library
    apply_to_library(cv)
    k = int(subsampling_factor)
    E, n = data.shape[0], data.shape[-1]
    scale, a_est, d32, weight = _scale_and_guess(data, noisemap, k, cv)
    psf = np.ascontiguousarray(psf, dtype=np.float32)
    if starlet_global_background or uniform_background_per_epoch:
        from .roi_modelling import joint_deconvolution
        res = joint_deconvolution(
            d32, weight, psf, k, xs=np.zeros(1), ys=np.zeros(1), initial_a=a_est,
            n_iter=n_iter, free_h=bool(starlet_global_background), free_mean=bool(uniform_background_per_epoch),
            conventions=cv, **COUPLED_FIT)
        return _coupled_result(res, scale, data, noisemap, n, k, cv)
    out = engine.phot_fit_batch(d32, weight, psf, np.arange(E, dtype=np.int32),
                                a_est.astype(np.float32), k, n_iter, lr=1e-3, schedule=True)
    a = out['a'].astype(np.float64)
    kw = {
        'kwargs_analytic': {'c_x': np.zeros(1), 'c_y': np.zeros(1), 'dx': out['dx'].astype(np.float64),
                            'dy': out['dy'].astype(np.float64), 'a': a, 'alpha': np.zeros(E)},
        'kwargs_background': {'h': np.zeros((n * k) ** 2), 'mean': np.zeros(E)},
        'kwargs_sersic': {},
    }
    model = d32 - out['residuals']
    loss_curve = out['loss_hist'].astype(np.float64).sum(0)   # the joint loss is the sum over epochs
    deconv = point_source_image(a[0], out['dx'][0], out['dy'][0], n, k, cv)
    return _result_dict(scale, kw, a, out['sigma_a'].astype(np.float64), loss_curve, data - model, noisemap, deconv,
                        np.zeros((n * k, n * k)), n, k, cv)


def do_stars_forward_modelling_coupled(stacks, subsampling_factor, n_iter=2000, uniform_background_per_epoch=False,
                                       starlet_global_background=True, conventions: Conventions = DEFAULT):
    """``do_one_star_forward_modelling`` with the background flags for MANY stars in one library call (the reference loops over the
    stars, star_photometry.py:257): stacks = list of (data, noisemap, psf) per star (scaled in place like the one-star function);
    every star is one joint fit (shared h / c / clip norm over its epochs), all fits are iterated together by
    ``lcb_deconv_run_many``.  Returns the list of result dicts."""
    cv = conventions
    from ..conventions import apply_to_library
    apply_to_library(cv)
    from .roi_modelling import joint_deconvolution_many
    k = int(subsampling_factor)
    problems, meta = [], []
    for data, noisemap, psf in stacks:
        scale, a_est, d32, weight = _scale_and_guess(data, noisemap, k, cv)
        problems.append(dict(data=d32, weight=weight, psf=np.ascontiguousarray(psf, dtype=np.float32), xs=np.zeros(1), ys=np.zeros(1),
                             initial_a=a_est))
        meta.append((scale, data, noisemap))
    res = joint_deconvolution_many(problems, k, n_iter=n_iter, free_h=bool(starlet_global_background),
                                   free_mean=bool(uniform_background_per_epoch), conventions=cv, **COUPLED_FIT)
    return [_coupled_result(r, scale, data, noisemap, data.shape[-1], k, cv) for r, (scale, data, noisemap) in zip(res, meta)]


def star_photometry_batch(data, noisemap, psfs, subsampling_factor, n_iter=2000, masks=None,
                          conventions: Conventions = DEFAULT, want_residuals=False, want_loss_hist=True, devices=None):
    """Batched driver replacing the serial loop of do_star_photometry (star_photometry.py:257-366):
    every (frame, star) item of a footprint in ONE library call.

    data, noisemap (F,S,n,n) [frame, star]; psfs (F,nu,nu) the narrow PSF of each frame; masks
    optional (F,S,n,n) bool, True = good (handled like star_photometry.py:309-316: NaN -> data 0 /
    noise 1e7, and the noise of an epoch with ANY masked pixel is multiplied by 1000).
    Per star the stack is scaled by its nanmax over all frames (:47-49) and a single background
    scalar enters the initial flux guess (:55-64).  Arrays are NOT modified in place.
    devices: None (current CUDA device), 'all', a count or a list of device indices: the STARS are split over the GPUs
    (every per-star quantity -- scale, background scalar -- needs all the epochs of that star), one host thread per GPU.
    Returns dict(fluxes, fluxes_uncertainties, chi2_per_frame (F,S), dx, dy, scale (S,), loss_curve (S,T)).
    """
    devs = engine.resolve_devices(devices)
    if len(devs) > 1 and data.shape[1] > 1:
        blocks = engine.split_by_work(np.ones(data.shape[1]), len(devs))

        def one(lo, hi):
            return star_photometry_batch(data[:, lo:hi], noisemap[:, lo:hi], psfs, subsampling_factor, n_iter=n_iter,
                                         masks=None if masks is None else masks[:, lo:hi], conventions=conventions,
                                         want_residuals=want_residuals, want_loss_hist=want_loss_hist, devices=None)
        parts = engine.fan_out(blocks, devs, one)
        star_axis = dict(scale=0, loss_curve=0)
        return {kk: np.concatenate([part[kk] for part in parts], axis=star_axis.get(kk, 1)) for kk in parts[0]}
    cv = conventions
    from ..conventions import apply_to_library
    apply_to_library(cv)
    k = int(subsampling_factor)
    F, S, n, _ = data.shape
    # NaN / mask policies, per-star scale, initial flux guess and weights run on the device (lcb_phot_prepare_batch)
    import torch
    prep = engine.phot_prepare_batch(data, noisemap, masks, k, downsample_mean=cv.downsample_mean)
    idx = torch.arange(F, dtype=torch.int32, device='cuda').repeat_interleave(S)
    psf_d = engine._to_device(psfs, torch.float32)
    out = engine.phot_fit_batch(prep['data'], prep['weight'], psf_d, idx, prep['a0'], k, n_iter, lr=1e-3, schedule=True,
                                want_residuals=want_residuals, want_loss_hist=want_loss_hist)
    out = engine.to_host(dict(out, scale=prep['scale']))
    scale = out.pop('scale')
    res = {
        'scale': scale,
        'fluxes': out['a'].reshape(F, S) * scale[None] / cv.amplitude_per_flux(k),          # pixel-sum units
        'fluxes_uncertainties': out['sigma_a'].reshape(F, S) * scale[None] / cv.amplitude_per_flux(k),
        'chi2_per_frame': out['chi2'].reshape(F, S),
        'dx': out['dx'].reshape(F, S), 'dy': out['dy'].reshape(F, S),
        'status': out['status'].reshape(F, S),
    }
    if want_loss_hist:
        res['loss_curve'] = out['loss_hist'].reshape(F, S, -1).sum(0)
    if want_residuals:
        res['residuals'] = out['residuals'].reshape(F, S, n, n) * scale[None, :, None, None]
    return res


STAR_FLUX_DDL = """CREATE TABLE IF NOT EXISTS star_flux_in_frame (
    frame_id INTEGER, star_gaia_id TEXT, combined_footprint_hash INTEGER, flux REAL, flux_uncertainty REAL, chi2 REAL,
    relative_loss_differential REAL,
    PRIMARY KEY (combined_footprint_hash, frame_id, star_gaia_id))"""   # columns of structure/database.py:396-408


def update_star_fluxes(db, flux_data):
    """star_photometry.py:201-229: upsert; on conflict ONLY flux and flux_uncertainty are refreshed."""
    db.execute(STAR_FLUX_DDL)
    db.executemany(
        "INSERT INTO star_flux_in_frame (combined_footprint_hash, frame_id, star_gaia_id, flux, flux_uncertainty, chi2, "
        "relative_loss_differential) VALUES (?, ?, ?, ?, ?, ?, ?) "
        "ON CONFLICT(combined_footprint_hash, frame_id, star_gaia_id) DO UPDATE SET "
        "flux=excluded.flux, flux_uncertainty=excluded.flux_uncertainty", flux_data)
    db.commit()


def gather_star_stack(store, frames, gaia_id, psf_ref_for_frame):
    """star_photometry.py:272-316 for one star: stamps, noise maps, PSFs of its frames from the stamp store, with the
    reference's NaN policy (both NaN -> data 0, noise 1e7) and its mask policy (the noise map of every epoch that has
    ANY masked pixel is multiplied by 1000 -- `noisemap[np.where(~mask)[0]] *= 1000.`, once per epoch)."""
    data, noisemap, mask, psf = [], [], [], []
    for frame in frames:
        rel = frame['image_relpath']
        data.append(store[f"{rel}/data/{gaia_id}"][...])
        noisemap.append(store[f"{rel}/noisemap/{gaia_id}"][...])
        mask.append(store[f"{rel}/cosmicsmask/{gaia_id}"][...])
        psf.append(store[f"{rel}/{psf_ref_for_frame(frame['id'])}/narrow_psf"][...])
    data, noisemap, psf = np.array(data, dtype=np.float64), np.array(noisemap, dtype=np.float64), np.array(psf)
    isnan = np.isnan(data) & np.isnan(noisemap)
    data[isnan] = 0.
    noisemap[isnan] = 1e7
    mask = ~(np.array(mask).astype(bool))
    noisemap[np.unique(np.where(~mask)[0])] *= 1000.
    return data, noisemap, psf


def do_star_photometry_batched(store, db, stars, frames_for_star, psf_ref_for_frame, user_config,
                               combined_footprint_hash, on_result=None, stager=None):
    """Batched form of do_star_photometry (star_photometry.py:232-373).

    stars: iterable of mappings with 'name', 'gaia_id'; frames_for_star(gaia_id) -> list of frame mappings (id,
    image_relpath) that need a measurement (get_frames_for_star's rows); psf_ref_for_frame(frame_id) -> psf_ref.
    The stamps of every (star, frame) item are gathered in bulk (``stamp_store.gather_photometry_batch``: groups resolved once
    per frame, the narrow PSF of a frame read once for all its stars, one page-locked staging buffer).  With the default
    pipeline flags (no shared background, no per-epoch constant) every item of every star goes into ONE K2 call; with the
    background flags the stars are joint fits (shared h / c / clip norm per star).  Rows are upserted with the reference's
    conflict rule, one executemany per star batch.
    Returns {gaia_id: result dict}.
    """
    logger = logging.getLogger('lightcurver.star_photometry')
    from .. import stamp_store
    cv = DEFAULT
    from ..conventions import apply_to_library
    apply_to_library(cv)
    k = int(user_config['subsampling_factor'])
    n_iter = int(user_config['star_deconv_n_iter'])
    coupled = bool(user_config.get('star_photometry_uniform_background_per_epoch', False)
                   or user_config.get('star_photometry_starlet_global_background', False))
    work = []
    for star in stars:
        frames = list(frames_for_star(star['gaia_id']))
        if len(frames) == 0:
            logger.info(f"Star {star['name']}: no new frames to process for this one. Skipping")
            continue
        work.append(dict(star=star, frames=frames))
    results = {}
    if not work:
        return results
    star_frames = [(wk['star']['gaia_id'], wk['frames']) for wk in work]
    if user_config.get('field_distortion', False):
        # star_photometry.py:291-304: every (star, frame) item sees the frame's narrow PSF resampled at the star's frame position
        d32, n32, cosmic, psfs, psf_index, off, thetas, xy = stamp_store.gather_photometry_batch(
            store, star_frames, psf_ref_for_frame, stager, with_distortion=True)
        psfs = engine.apply_distortion_batch(psfs, thetas, psf_index, xy, mode=cv.distortion_mode())   # one launch for the batch
        psf_index = np.arange(len(psf_index), dtype=np.int32)
    else:
        d32, n32, cosmic, psfs, psf_index, off = stamp_store.gather_photometry_batch(store, star_frames, psf_ref_for_frame, stager)
    # star_photometry.py:309-316 on the whole batch: doubly-NaN pixels -> (0, 1e7); the noise map of every epoch that has ANY
    # masked pixel is multiplied by 1000 (`noisemap[np.where(~mask)[0]] *= 1000.`, once per epoch)
    data, noisemap = d32.astype(np.float64), n32.astype(np.float64)
    isnan = np.isnan(data) & np.isnan(noisemap)
    data[isnan] = 0.
    noisemap[isnan] = 1e7
    noisemap[cosmic.any(axis=(1, 2))] *= 1000.
    for i, wk in enumerate(work):
        sl = slice(int(off[i]), int(off[i + 1]))
        wk['sl'], wk['data'], wk['noisemap'] = sl, data[sl], noisemap[sl]
    if coupled:
        # every star is a joint fit over its epochs (shared background / centre / clip norm); all stars in ONE library call
        res = do_stars_forward_modelling_coupled(
            [(wk['data'], wk['noisemap'], psfs[psf_index[wk['sl']]]) for wk in work], k, n_iter=n_iter,
            uniform_background_per_epoch=user_config.get('star_photometry_uniform_background_per_epoch', False),
            starlet_global_background=user_config.get('star_photometry_starlet_global_background', False))
        for wk, r in zip(work, res):
            wk['result'] = r
    else:
        # per star: scale (:47-49), initial flux guess (:55-64); then one K2 launch over all stars' epochs
        n = data.shape[-1]
        a0 = np.empty(data.shape[0], np.float32)
        for wk in work:
            d, nm = wk['data'], wk['noisemap']
            wk['scale'] = float(np.nanmax(d))
            d /= wk['scale']                       # in place on the batch arrays, like the reference on its own (:48-49)
            nm /= wk['scale']
            a0[wk['sl']] = _initial_flux_guess(d) * cv.amplitude_per_flux(k)
        ds, ws = stamps_and_weights(data, noisemap)
        out = engine.phot_fit_batch(ds, ws, psfs, psf_index, a0, k, n_iter, lr=1e-3, schedule=True)
        for wk in work:
            sl = wk['sl']
            residuals = out['residuals'][sl].astype(np.float64)             # data - model, scaled units
            with np.errstate(divide='ignore', invalid='ignore'):
                chi2_per_frame = np.nansum(residuals ** 2 / wk['noisemap'] ** 2, axis=(1, 2)) / n ** 2
            wk['result'] = {
                'scale': wk['scale'], 'fluxes': wk['scale'] * out['a'][sl].astype(np.float64) / cv.amplitude_per_flux(k),
                'fluxes_uncertainties': wk['scale'] * out['sigma_a'][sl].astype(np.float64) / cv.amplitude_per_flux(k),
                'status': out['status'][sl],
                'chi2': float(np.nanmean(chi2_per_frame)), 'chi2_per_frame': chi2_per_frame,
                'loss_curve': list(out['loss_hist'][sl].astype(np.float64).sum(0)), 'residuals': wk['scale'] * residuals,
                'kwargs_final': {'kwargs_analytic': {'c_x': np.zeros(1), 'c_y': np.zeros(1), 'dx': out['dx'][sl], 'dy': out['dy'][sl],
                                                     'a': out['a'][sl], 'alpha': np.zeros(sl.stop - sl.start)},
                                 'kwargs_background': {'h': np.zeros((n * k) ** 2), 'mean': np.zeros(sl.stop - sl.start)},
                                 'kwargs_sersic': {}},
            }
    flux_rows = []
    from .psf_modelling import relative_loss_differential
    for wk in work:
        result, star = wk['result'], wk['star']
        if on_result is not None:
            on_result(star, wk['data'], wk['noisemap'], result)
        rld = relative_loss_differential(result['loss_curve'])
        # non-finite measurements (per-item status of the library, or a NaN that slipped through) never reach the database:
        # the reference has no such guard because a NaN loss aborts its whole task (failure detection)
        status = np.asarray(result.get('status', np.zeros(len(wk['frames']), np.int32)))
        good = [j for j in range(len(wk['frames']))
                if status[j] == 0 and np.isfinite(result['fluxes'][j]) and np.isfinite(result['fluxes_uncertainties'][j])]
        if len(good) != len(wk['frames']):
            logger.warning(f"Star {star['name']}: {len(wk['frames']) - len(good)} non-finite measurement(s) skipped "
                           f"(frames {[wk['frames'][j]['id'] for j in range(len(wk['frames'])) if j not in good]})")
        flux_rows += [(combined_footprint_hash, wk['frames'][j]['id'], star['gaia_id'], float(result['fluxes'][j]),
                       float(result['fluxes_uncertainties'][j]), float(result['chi2_per_frame'][j]), rld) for j in good]
        results[star['gaia_id']] = result
        logger.info(f"Measured star {star['name']} in {len(wk['frames'])} frames. "
                    f"The global reduced chi2 is {result['chi2']:.02f}.")
    update_star_fluxes(db, flux_rows)              # one executemany + one commit for every star of the batch
    return results
