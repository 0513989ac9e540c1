"""The batched pipeline drivers on an in-memory stamp store and an
in-memory SQLite database: psf_modeling -> star_photometry, with the reference's bookkeeping rules."""
import sqlite3

import numpy as np
import pytest

from lightcurver_b200 import synthetic
from lightcurver_b200.processes.psf_modelling import MemoryStore, model_all_psfs_batched
from lightcurver_b200.processes.star_photometry import do_star_photometry_batched

pytestmark = pytest.mark.gpu


def _world(F=3, N=4, n=24, k=2, seed=91):
    d = synthetic.make_psf_frames(F, N, n, k, seed=seed)
    store = MemoryStore()
    names = ['a', 'b', 'c', 'd'][:N]
    gaia = [str(1000 + i) for i in range(N)]
    frames = []
    for f in range(F):
        rel = f"frames/img{f}.fits"
        for i in range(N):
            data = d['data'][f, i].copy()
            nm = d['noisemap'][f, i].copy()
            cosmic = ~d['masks'][f, i]
            if f == 1 and i == 2:                       # more than 40 % masked -> dropped from that frame's PSF
                cosmic[:12, :] = True
            if f == 0 and i == 0:                       # a dead pixel: NaN in both
                data[3, 3] = np.nan; nm[3, 3] = np.nan
            store[f"{rel}/data/{gaia[i]}"] = data
            store[f"{rel}/noisemap/{gaia[i]}"] = nm
            store[f"{rel}/cosmicsmask/{gaia[i]}"] = cosmic
        frames.append(dict(id=f + 1, image_relpath=rel, seeing_pixels=float(d['fwhm'][f]), pixel_scale=0.2))
    stars = [dict(name=nm_, gaia_id=g) for nm_, g in zip(names, gaia)]
    return d, store, frames, stars


def test_psf_then_photometry_drivers(cuda_device):
    d, store, frames, stars = _world()
    db = sqlite3.connect(':memory:')
    cfg = dict(subsampling_factor=2, psf_n_iter_analytic=60, psf_n_iter_pixels=300, redo_psf=False, field_distortion=False,
               star_deconv_n_iter=400)
    h = 42
    all_ones = lambda data, nm: np.ones(data.shape, bool)
    written = model_all_psfs_batched(store, db, frames, lambda fid: stars, cfg, h, automatic_mask_fn=all_ones)
    assert [w[0] for w in written] == [1, 2, 3]
    rows = db.execute("SELECT frame_id, chi2, psf_ref, subsampling_factor, relative_loss_differential, fwhm_moffat_arcseconds "
                      "FROM PSFs ORDER BY frame_id").fetchall()
    assert len(rows) == 3 and all(r[2] == 'psf_abcd' and r[3] == 2 for r in rows)
    assert all(r[1] < 2 for r in rows)                                  # the reference's acceptance bound
    assert all(0 < r[5] < 2.0 for r in rows) and all(np.isfinite(r[4]) for r in rows)
    for fr in frames:
        g = store[f"{fr['image_relpath']}/psf_abcd"]
        assert g['narrow_psf'][...].shape == (48, 48) and g['full_psf'][...].shape == (48, 48)
        assert int(g['subsampling_factor'][...][0]) == 2 and 'distortion' in g.keys()
    # second call: nothing pending (redo_psf false)
    assert model_all_psfs_batched(store, db, frames, lambda fid: stars, cfg, h, automatic_mask_fn=all_ones) == []
    # photometry of every star in every frame, one library call
    res = do_star_photometry_batched(store, db, stars, lambda gid: frames, lambda fid: 'psf_abcd', cfg, h)
    assert set(res) == {s['gaia_id'] for s in stars}
    rows = db.execute("SELECT frame_id, star_gaia_id, flux, flux_uncertainty, chi2 FROM star_flux_in_frame").fetchall()
    assert len(rows) == 12
    truth = {(f + 1, str(1000 + i)): d['flux'][f, i] for f in range(3) for i in range(4)}
    for fid, gid, flux, sig, chi2 in rows:
        if (fid, gid) == (2, '1002'):
            continue                                    # the half-masked epoch: noise x1000, flux unconstrained
        assert abs(flux - truth[(fid, gid)]) < 8 * sig + 0.05 * truth[(fid, gid)], (fid, gid, flux, truth[(fid, gid)])   # pixel-sum units
        assert sig > 0
    # upsert semantics (star_photometry.py:220-225): a redo refreshes flux and flux_uncertainty only
    db.execute("UPDATE star_flux_in_frame SET chi2 = -1, flux = 0")
    do_star_photometry_batched(store, db, stars[:1], lambda gid: frames, lambda fid: 'psf_abcd', cfg, h)
    again = db.execute("SELECT flux, chi2 FROM star_flux_in_frame WHERE star_gaia_id = '1000'").fetchall()
    assert all(fl > 0 and c == -1 for fl, c in again)


def test_photometry_driver_with_shared_background_flag(cuda_device):
    """star_photometry_starlet_global_background=True routes each star through the joint-deconvolution engine."""
    d, store, frames, stars = _world(F=3, N=2, n=16, k=2, seed=12)
    db = sqlite3.connect(':memory:')
    cfg = dict(subsampling_factor=2, psf_n_iter_analytic=40, psf_n_iter_pixels=100, star_deconv_n_iter=60,
               star_photometry_starlet_global_background=True)
    all_ones = lambda data, nm: np.ones(data.shape, bool)
    model_all_psfs_batched(store, db, frames, lambda fid: stars, cfg, 7, automatic_mask_fn=all_ones)
    res = do_star_photometry_batched(store, db, stars, lambda gid: frames, lambda fid: 'psf_ab', cfg, 7)
    for r in res.values():
        assert len(r['loss_curve']) == 60 and r['starlet_background'].shape == (32, 32)
        assert np.isfinite(r['fluxes']).all() and np.isfinite(r['fluxes_uncertainties']).all()
    assert db.execute("SELECT COUNT(*) FROM star_flux_in_frame").fetchone()[0] == 6
