"""Batched driver for the PSF-modelling step: lightcurver/processes/psf_modelling.py:64-225 with the serial
per-frame STARRED call (:92, :164-171) replaced by ONE library call for every pending frame.

Everything around the fit keeps the reference's behaviour, in the reference's order:
  * star list per frame -> ``psf_ref = 'psf_' + ''.join(sorted(names))`` (:103), skip if already modelled and not
    ``redo_psf`` (:106-110);
  * stamps, noise maps, cosmic masks from the stamp store ``<image_relpath>/{data,noisemap,cosmicsmask}/<gaia_id>``
    (:113-127);
  * masks = ~cosmics * automatic (:135); where BOTH data and noise are NaN: data 0, noise 1, mask False (:136-140);
    stars with more than 40 % masked pixels are dropped (:144-153); frames left without stars are skipped (:154-160);
  * products: ``<frame>/<psf_ref>/{narrow_psf, full_psf, subsampling_factor, distortion/}`` (:190-202),
    ``fwhm_moffat_arcseconds`` (:177-179), ``relative_loss_differential`` (:205-208), ``REPLACE INTO PSFs`` (:209-217).

The store is any h5py.File-like object (``MemoryStore`` below offers the same subset in memory; h5py is not
installed in the build container).  The database is a ``sqlite3`` connection; star selection is injected as a
callable because it is SQL bookkeeping that stays lightcurver's.
"""
import logging
from pathlib import Path

import numpy as np

from ..procedures.psf_routines import build_psf_batch

MASK_THRESHOLD_FRACTION = 0.4            # psf_modelling.py:146

PSFS_DDL = """CREATE TABLE IF NOT EXISTS PSFs (
    combined_footprint_hash INTEGER, frame_id INTEGER, chi2 REAL, psf_ref TEXT, subsampling_factor INTEGER,
    relative_loss_differential REAL, fwhm_moffat_arcseconds REAL DEFAULT NULL,
    PRIMARY KEY (combined_footprint_hash, frame_id, psf_ref))"""   # columns of structure/database.py:381-392


class MemoryStore(dict):
    """In-memory stand-in for the h5py.File subset the pipeline uses: ``store['a/b/c'][...]``, ``name in group``,
    ``group.create_group(name)``, ``group[name] = array``, ``del group[name]``, ``group.keys()``."""

    class _Dataset:
        def __init__(self, value):
            self.value = np.asarray(value)

        def __getitem__(self, key):
            return self.value[key]

        @property
        def shape(self):
            return self.value.shape

    def _walk(self, path, create=False):
        node = self
        parts = [p for p in str(path).split('/') if p]
        for p in parts[:-1]:
            if p not in dict.keys(node):
                if not create:
                    raise KeyError(path)
                dict.__setitem__(node, p, MemoryStore())
            node = dict.__getitem__(node, p)
        return node, parts[-1]

    def __getitem__(self, path):
        node, leaf = self._walk(path)
        return dict.__getitem__(node, leaf)

    def __setitem__(self, path, value):
        node, leaf = self._walk(path, create=True)
        dict.__setitem__(node, leaf, value if isinstance(value, (MemoryStore, MemoryStore._Dataset)) else MemoryStore._Dataset(value))

    def __delitem__(self, path):
        node, leaf = self._walk(path)
        dict.__delitem__(node, leaf)

    def __contains__(self, path):
        try:
            self[path]
            return True
        except KeyError:
            return False

    def create_group(self, path):
        g = MemoryStore()
        self[path] = g
        return g


def mask_surrounding_stars(data, noisemap):
    """psf_modelling.py:35-61: mask every detected object but the central one.  Needs the optional ``sep``
    package (a C source extractor that stays host side); without it nothing is masked."""
    try:
        import sep
    except ImportError:
        return np.ones(data.shape, dtype=bool)
    objects, seg_map = sep.extract(np.ascontiguousarray(data, dtype=np.float32), thresh=3., err=noisemap, minarea=15,
                                   segmentation_map=True, deblend_cont=0.001)
    mask = np.ones_like(seg_map, dtype=bool)
    if len(objects) == 0:
        return mask
    cy, cx = (data.shape[0] - 1) / 2.0, (data.shape[1] - 1) / 2.0
    central = np.argmin(np.hypot(objects['x'] - cx, objects['y'] - cy))
    for i in range(len(objects)):
        if i != central:
            mask[seg_map == i + 1] = False
    return mask


def prepare_psf_inputs(datas, noisemaps, cosmics_masks, automatic_masks=None):
    """psf_modelling.py:128-153 for one frame.  cosmics_masks: True = cosmic (bad).  Returns
    (datas, noisemaps, masks, keep) with the >40 %-masked stars removed; ``keep`` indexes the original stars."""
    datas = np.array(datas, dtype=np.float32)
    noisemaps = np.array(noisemaps, dtype=np.float32)
    good = ~np.asarray(cosmics_masks, dtype=bool)
    if automatic_masks is None:
        automatic_masks = np.array([mask_surrounding_stars(d, n) for d, n in zip(datas, noisemaps)])
    masks = good & np.asarray(automatic_masks, dtype=bool)
    isnan = np.isnan(datas) & np.isnan(noisemaps)
    datas[isnan] = 0.
    noisemaps[isnan] = 1.0
    masks[isnan] = False
    masked_counts = np.sum(~masks, axis=(1, 2))
    keep = ~(masked_counts > MASK_THRESHOLD_FRACTION * datas.shape[1] * datas.shape[2])
    return datas[keep], noisemaps[keep], masks[keep], np.nonzero(keep)[0]


def relative_loss_differential(loss_history):
    """psf_modelling.py:205-208 / star_photometry.py:349-352."""
    lh = np.asarray(loss_history, dtype=np.float64)
    idx = int(0.9 * lh.size)
    with np.errstate(all='ignore'):
        initial = np.nanmax(lh[:idx]) - np.nanmin(lh[:idx])
        end = np.nanmax(lh[idx:]) - np.nanmin(lh[idx:])
        return float(end / initial)


def check_psf_exists(db, frame_id, psf_ref, combined_footprint_hash):
    cur = db.execute("SELECT 1 FROM PSFs WHERE frame_id = ? AND psf_ref = ? and combined_footprint_hash = ?",
                     (frame_id, psf_ref, combined_footprint_hash))
    return cur.fetchone() is not None


def model_all_psfs_batched(store, db, frames, stars_for_frame, user_config, combined_footprint_hash,
                           automatic_mask_fn=mask_surrounding_stars, on_result=None, devices=None, stager=None):
    """One pass over ``frames`` (iterable of mappings with id, image_relpath, seeing_pixels, pixel_scale):
    gather every frame that needs a PSF, fit them all in ONE ``build_psf_batch`` call, then write the per-frame
    products to the store and the PSFs table in the original order.

    stars_for_frame(frame_id) -> list of mappings with 'name' and 'gaia_id' (select_stars_for_a_frame's rows).
    on_result(frame, result, datas, noisemaps, masks, names) is called per frame (diagnostic plot hook, :182-187).
    automatic_mask_fn: per-stamp masking of neighbouring objects (:129-135; ``sep`` based by default), None = no masking.
    stager: a reusable ``stamp_store.StampStager`` (page-locked staging buffers; one is created per call otherwise).
    The stamps are gathered in bulk (``stamp_store.gather_psf_batch``), the products are written back group by group and the
    PSFs rows in ONE executemany.
    Returns the list of (frame_id, psf_ref, chi2) written.
    """
    logger = logging.getLogger('lightcurver.psf_modelling')
    from .. import stamp_store
    db.execute(PSFS_DDL)
    k = int(user_config['subsampling_factor'])
    cand = []
    for frame in frames:
        stars = list(stars_for_frame(frame['id']))
        if len(stars) == 0:
            logger.warning(f"The frame with id {frame['id']} does not have available reference stars. Skipping.")
            continue
        psf_ref = 'psf_' + ''.join(sorted(s['name'] for s in stars))
        if check_psf_exists(db, frame['id'], psf_ref, combined_footprint_hash) and not user_config.get('redo_psf', False):
            logger.info(f"The frame with id {frame['id']} already has a PSF (ref {psf_ref}), redo flag not set. Skipping.")
            continue
        cand.append((frame, stars, psf_ref))
    if not cand:
        return []
    # ---- gather: every stamp of every pending frame into ONE page-locked staging buffer (three group lookups per frame)
    stager = stager if stager is not None else stamp_store.StampStager()
    datas, noisemaps, cosmics, off = stamp_store.gather_psf_batch(store, [c[0] for c in cand], [[s['gaia_id'] for s in c[1]] for c in cand],
                                                                  stager)
    # ---- host policies of psf_modelling.py:128-153, vectorised over the whole batch (in place in the staging buffer)
    if automatic_mask_fn is None:
        masks = ~cosmics
    else:
        masks = ~cosmics & np.array([automatic_mask_fn(d, nm) for d, nm in zip(datas, noisemaps)], dtype=bool).reshape(cosmics.shape)
    isnan = np.isnan(datas) & np.isnan(noisemaps)
    datas[isnan] = 0.
    noisemaps[isnan] = 1.0
    masks[isnan] = False
    keep = ~((~masks).sum(axis=(1, 2)) > MASK_THRESHOLD_FRACTION * datas.shape[1] * datas.shape[2])
    pending, counts, sel = [], [], []
    for f, (frame, stars, psf_ref) in enumerate(cand):
        kf = np.nonzero(keep[off[f]:off[f + 1]])[0]
        if len(kf) == 0:
            logger.warning(f"The frame with id {frame['id']} had {len(stars)} reference stars, but none could be used "
                           "due to too many masked pixels. Skipping.")
            continue
        pending.append(dict(frame=frame, psf_ref=psf_ref, names=[stars[i]['name'] for i in kf], n_before=len(stars),
                            rows=off[f] + kf))
        counts.append(len(kf))
        sel.append(off[f] + kf)
    if not pending:
        return []
    sel = np.concatenate(sel)
    field_distortion = bool(user_config.get('field_distortion', False))
    positions = None
    if field_distortion:                      # psf_modelling.py:121-124: rescaled frame positions of the stamps
        positions = stamp_store.gather_positions(store, [c[0] for c in cand], [[s['gaia_id'] for s in c[1]] for c in cand])[sel]
    if len(sel) != len(keep):                 # some stars dropped: compact (one copy); otherwise the staging views go straight up
        datas, noisemaps, masks = datas[sel], noisemaps[sel], masks[sel]
    results = build_psf_batch(datas, noisemaps, k, masks=masks, star_counts=counts,
                              n_iter_analytic=user_config['psf_n_iter_analytic'],
                              n_iter_adabelief=user_config['psf_n_iter_pixels'],
                              guess_method_star_position='center',
                              guess_fwhm_pixels=np.array([float(p['frame']['seeing_pixels']) for p in pending]),
                              field_distortion=field_distortion, stamp_coordinates=positions,
                              devices=devices)          # None: current GPU; 'all': every visible GPU, one host thread each
    written, rows = [], []
    pos = 0
    for p, result in zip(pending, results):
        frame, psf_ref = p['frame'], p['psf_ref']
        km = result['kwargs_psf']['kwargs_moffat']
        fwhm_moffat_arcseconds = float((0.5 * (km['fwhm_x'] + km['fwhm_y']) * frame['pixel_scale']).item())
        loss_history = result['adabelief_extra_fields']['loss_history']
        if on_result is not None:
            sl = slice(pos, pos + len(p['names']))
            on_result(frame, result, datas[sl], noisemaps[sl], masks[sl], p['names'])
        pos += len(p['names'])
        stamp_store.write_psf_products(store, frame, psf_ref, result['narrow_psf'], result['full_psf'], k,
                                       result['kwargs_psf']['kwargs_distortion'])
        rld = relative_loss_differential(loss_history)
        rows.append((frame['id'], float(result['chi2']), rld, psf_ref, combined_footprint_hash, k, fwhm_moffat_arcseconds))
        written.append((frame['id'], psf_ref, float(result['chi2'])))
        logger.info(f"PSF built for frame with id {frame['id']}. The reference is {psf_ref}, that is {p['n_before']} stars "
                    f"available, and {len(p['names'])} actually used after filtering of masked pixels. "
                    f"The reduced chi2 is {result['chi2']:.02f}.")
    stamp_store.replace_psf_rows(db, rows)               # one executemany + one commit for the batch
    return written
