"""Bulk gather / scatter between lightcurver's stamp store and the batched fits.

The reference reads the stamp store one dataset at a time inside its serial loops: four h5py lookups per (frame, star) in
``model_all_psfs`` (lightcurver/processes/psf_modelling.py:113-127), three stamps plus one PSF per (star, frame) in
``do_star_photometry`` (star_photometry.py:272-306, the PSF of a frame is re-read for every star), and one ``REPLACE`` /
upsert statement per frame or star (psf_modelling.py:209-217, star_photometry.py:354-366).  Once the fits run at ~10^3
frames per second that bookkeeping is the wall-clock term, so the batched drivers go through this module instead:

  * the groups of a frame (``<image_relpath>/{data,noisemap,cosmicsmask}``) are resolved ONCE per frame and
    every dataset is read straight into its slot of ONE page-locked staging buffer per batch (``Dataset.read_direct`` when the
    store is a real ``h5py.File``: no intermediate array), which then goes to the device in one asynchronous copy;
  * the narrow PSF of a (frame, psf_ref) is read once and shared by all the stars of that frame;
  * PSF products are written back group by group in the original order and the SQL rows go through ONE ``executemany`` per
    table inside one transaction.

The store is any object with the ``h5py.File`` subset the reference uses (``store[path]``, ``group[name]``,
``group.keys()``, ``create_group``, item assignment, ``del``); ``open_h5`` opens the real thing when ``h5py`` is installed
(it is not in the build container: tests fall back to ``MemoryStore`` and to a dataset double with ``read_direct``).
"""
import numpy as np


def open_h5(path, mode='r'):
    """``h5py.File(path, mode)`` -- the reference's ``regions.h5`` (user_config.py:46).  Import-guarded: h5py is an
    optional dependency of this package (the kernels never need it)."""
    try:
        import h5py
    except ImportError as exc:                        # pragma: no cover - depends on the environment
        raise ImportError("opening a regions.h5 stamp store needs the optional dependency h5py "
                          "(pass any h5py.File-like object, e.g. processes.psf_modelling.MemoryStore, otherwise)") from exc
    return h5py.File(str(path), mode)


def pinned_empty(shape, dtype):
    """Page-locked host array when a CUDA device is present (async H2D of the whole batch), plain numpy otherwise."""
    try:
        import torch
        if torch.cuda.is_available():
            tdt = {np.dtype(np.float32): torch.float32, np.dtype(np.uint8): torch.uint8, np.dtype(np.bool_): torch.bool,
                   np.dtype(np.int32): torch.int32}[np.dtype(dtype)]
            return torch.empty(tuple(shape), dtype=tdt, pin_memory=True).numpy()
    except Exception:
        pass
    return np.empty(shape, dtype=dtype)


def read_into(dataset, out):
    """One dataset -> its slot of the staging buffer, without an intermediate array when the store supports it."""
    rd = getattr(dataset, 'read_direct', None)
    if rd is not None and getattr(dataset, 'dtype', None) == out.dtype and tuple(dataset.shape) == out.shape:
        rd(out)
    else:
        out[...] = dataset[...]


class StampStager:
    """Reusable page-locked staging buffers for one batch of stamps: data, noisemap (float32) and cosmics (bool)."""

    def __init__(self):
        self.cap, self.n = 0, 0
        self.data = self.noise = self.cosmic = None

    def ensure(self, count, n):
        if count > self.cap or n != self.n:
            self.cap, self.n = max(count, int(self.cap * 1.5)), n
            self.data = pinned_empty((self.cap, n, n), np.float32)
            self.noise = pinned_empty((self.cap, n, n), np.float32)
            self.cosmic = pinned_empty((self.cap, n, n), np.bool_)
        return self

    def read_frame(self, store, image_relpath, ids, lo):
        """Stamps of the objects ``ids`` of one frame into rows lo .. lo + len(ids): three group lookups per FRAME."""
        frame_group = store[image_relpath]
        dg, ng, mg = frame_group['data'], frame_group['noisemap'], frame_group['cosmicsmask']
        for j, gid in enumerate(ids):
            read_into(dg[gid], self.data[lo + j])
            read_into(ng[gid], self.noise[lo + j])
            self.cosmic[lo + j] = np.asarray(mg[gid][...], dtype=bool)
        return frame_group


def stamp_side(store, image_relpath, gid):
    return int(store[image_relpath]['data'][gid].shape[-1])


def gather_psf_batch(store, frames, ids_per_frame, stager=None):
    """Every stamp of every pending frame into one staging buffer.  Returns (data, noisemap, cosmics, star_off): views of the
    first sumN rows of the staging arrays (valid until the stager is reused) and the CSR offsets per frame."""
    counts = [len(ids) for ids in ids_per_frame]
    off = np.zeros(len(frames) + 1, np.int64)
    off[1:] = np.cumsum(counts)
    total = int(off[-1])
    if total == 0:
        z = np.zeros((0, 0, 0), np.float32)
        return z, z, z.astype(bool), off
    first = next(i for i, c in enumerate(counts) if c)
    n = stamp_side(store, frames[first]['image_relpath'], ids_per_frame[first][0])
    st = (stager or StampStager()).ensure(total, n)
    for f, (frame, ids) in enumerate(zip(frames, ids_per_frame)):
        if ids:
            st.read_frame(store, frame['image_relpath'], ids, int(off[f]))
    return st.data[:total], st.noise[:total], st.cosmic[:total], off


def rescale_image_coordinates(xy_coordinates_array, image_shape):
    """utilities/image_coordinates.py:4-25: pixel coordinates (x, y) with the origin at the bottom left of the frame -> origin at
    the frame centre, divided by the frame dimensions (values in [-1/2, 1/2]); image_shape is (rows, columns)."""
    dims = np.asarray(image_shape, dtype=np.float64).reshape(-1)[::-1]        # (x extent, y extent)
    return (np.asarray(xy_coordinates_array, dtype=np.float64) - (dims - 1) / 2.) / dims


def gather_positions(store, frames, ids_per_frame):
    """Rescaled frame positions of the stamps (psf_modelling.py:121-124: ``image_pixel_coordinates/<id>`` and ``frame_shape`` of
    every frame), one (x, y) pair per star in the order of ``gather_psf_batch``.  Returns (sumN, 2) float32."""
    out = []
    for frame, ids in zip(frames, ids_per_frame):
        if not ids:
            continue
        frame_group = store[frame['image_relpath']]
        pg = frame_group['image_pixel_coordinates']
        shape = np.asarray(frame_group['frame_shape'][...])
        pos = np.array([np.asarray(pg[gid][...], dtype=np.float64).reshape(2) for gid in ids])
        out.append(rescale_image_coordinates(pos, shape))
    return np.concatenate(out).astype(np.float32) if out else np.zeros((0, 2), np.float32)


def read_distortion(psf_group):
    """star_photometry.py:293-297: the ``distortion`` group of a stored PSF -> (6,) coefficients (dilation_x, dilation_y, shear),
    zeros when the group is empty (the PSF was built without field distortion)."""
    dg = psf_group['distortion']
    keys = set(dg.keys())
    if not {'dilation_x', 'dilation_y', 'shear'} <= keys:
        return np.zeros(6, np.float32)
    return np.concatenate([np.asarray(dg[key][...], dtype=np.float32).reshape(2) for key in ('dilation_x', 'dilation_y', 'shear')])


def gather_photometry_batch(store, star_frames, psf_ref_for_frame, stager=None, with_distortion=False):
    """``star_frames``: list (one entry per star) of (gaia_id, [frame mappings]).  Frames are walked in the OUTER loop so that
    the groups of a frame are resolved once and its narrow PSF is read once, whatever the number of stars measured in it.
    Returns (data, noisemap, cosmics, psf_stack (n_psf, nu, nu), psf_index (B,), star_off (S + 1,)) with the items of a star
    contiguous and in the order of its frame list (star_photometry.py:272-306).  with_distortion (:293-304): two more values,
    the distortion coefficients of every PSF (n_psf, 6) and the rescaled frame position of every item (B, 2)."""
    counts = [len(fr) for _, fr in star_frames]
    off = np.zeros(len(star_frames) + 1, np.int64)
    off[1:] = np.cumsum(counts)
    total = int(off[-1])
    if total == 0:
        z = np.zeros((0, 0, 0), np.float32)
        empty = (z, z, z.astype(bool), z, np.zeros(0, np.int32), off)
        return empty + (np.zeros((0, 6), np.float32), np.zeros((0, 2), np.float32)) if with_distortion else empty
    by_frame = {}                                  # image_relpath -> (frame, [(item index, gaia_id)])
    for s, (gid, frs) in enumerate(star_frames):
        for j, fr in enumerate(frs):
            by_frame.setdefault(fr['image_relpath'], (fr, []))[1].append((int(off[s]) + j, gid))
    any_fr, any_items = next(iter(by_frame.values()))
    n = stamp_side(store, any_fr['image_relpath'], any_items[0][1])
    st = (stager or StampStager()).ensure(total, n)
    psf_index = np.empty(total, np.int32)
    psfs, psf_slot = [], {}
    thetas, xy = [], np.zeros((total, 2), np.float32)
    for rel, (fr, items) in by_frame.items():
        frame_group = store[rel]
        dg, ng, mg = frame_group['data'], frame_group['noisemap'], frame_group['cosmicsmask']
        ref = psf_ref_for_frame(fr['id'])
        key = (rel, ref)
        if key not in psf_slot:
            psf_slot[key] = len(psfs)
            psfs.append(np.asarray(frame_group[ref]['narrow_psf'][...], dtype=np.float32))
            if with_distortion:
                thetas.append(read_distortion(frame_group[ref]))
        if with_distortion:
            pg, shape = frame_group['image_pixel_coordinates'], np.asarray(frame_group['frame_shape'][...])
            for idx, gid in items:
                xy[idx] = rescale_image_coordinates(np.asarray(pg[gid][...], dtype=np.float64).reshape(2), shape)
        for idx, gid in items:
            read_into(dg[gid], st.data[idx])
            read_into(ng[gid], st.noise[idx])
            st.cosmic[idx] = np.asarray(mg[gid][...], dtype=bool)
            psf_index[idx] = psf_slot[key]
    res = (st.data[:total], st.noise[:total], st.cosmic[:total], np.stack(psfs), psf_index, off)
    return res + (np.stack(thetas), xy) if with_distortion else res


def write_psf_products(store, frame, psf_ref, narrow_psf, full_psf, subsampling_factor, kwargs_distortion):
    """psf_modelling.py:190-202: <frame>/<psf_ref>/{narrow_psf, full_psf, subsampling_factor, distortion/*}, replacing an
    existing group of the same reference."""
    frame_group = store[frame['image_relpath']]
    if psf_ref in frame_group.keys():
        del frame_group[psf_ref]
    psf_group = frame_group.create_group(psf_ref)
    psf_group['narrow_psf'] = np.asarray(narrow_psf)
    psf_group['full_psf'] = np.asarray(full_psf)
    psf_group['subsampling_factor'] = np.array([subsampling_factor])
    distortion_group = psf_group.create_group('distortion')
    for key, value in kwargs_distortion.items():
        distortion_group[key] = value
    return psf_group


def replace_psf_rows(db, rows):
    """psf_modelling.py:209-217 for a whole batch: one executemany, one commit.  rows: (frame_id, chi2,
    relative_loss_differential, psf_ref, combined_footprint_hash, subsampling_factor, fwhm_moffat_arcseconds)."""
    db.executemany("REPLACE INTO PSFs (frame_id, chi2, relative_loss_differential, psf_ref, combined_footprint_hash, "
                   "subsampling_factor, fwhm_moffat_arcseconds) VALUES (?,?,?,?,?,?,?)", rows)
    db.commit()
