"""Bulk gather / scatter between the stamp store and the batched fits: host logic, no GPU.

h5py is not installed in the build container, so the h5py.File protocol is exercised through a small double that offers
what the adapter uses of it (path lookups, groups, ``Dataset.read_direct``, item assignment) and COUNTS the calls; the
same tests run against a real regions.h5 written by the test itself when h5py is importable."""
import sqlite3

import numpy as np
import pytest

from lightcurver_b200 import stamp_store
from lightcurver_b200.processes.psf_modelling import MemoryStore, PSFS_DDL


class CountingDataset:
    """h5py.Dataset double: shape, dtype, [...], read_direct."""
    def __init__(self, value, counter):
        self.value, self.counter = np.ascontiguousarray(value), counter
        self.shape, self.dtype = self.value.shape, self.value.dtype

    def __getitem__(self, key):
        self.counter['getitem'] += 1
        return self.value[key]

    def read_direct(self, dest):
        self.counter['read_direct'] += 1
        dest[...] = self.value


class CountingGroup(dict):
    """h5py.Group double that resolves 'a/b/c' paths and counts group lookups."""
    def __init__(self, counter):
        super().__init__()
        self.counter = counter

    def __getitem__(self, path):
        node = self
        for p in [q for q in str(path).split('/') if q]:
            node = dict.__getitem__(node, p)
            if isinstance(node, CountingGroup):
                self.counter['group'] += 1
        return node

    def __setitem__(self, path, value):
        parts = [q for q in str(path).split('/') if q]
        node = self
        for p in parts[:-1]:
            if p not in dict.keys(node):
                dict.__setitem__(node, p, CountingGroup(self.counter))
            node = dict.__getitem__(node, p)
        dict.__setitem__(node, parts[-1], value if isinstance(value, (CountingGroup, CountingDataset)) else CountingDataset(value, self.counter))

    def create_group(self, name):
        g = CountingGroup(self.counter)
        self[name] = g
        return g


def _fill(store, F, S, n, rng):
    frames, ids = [], [str(5000 + i) for i in range(S)]
    truth = {}
    for f in range(F):
        rel = f"frames/img{f}.fits"
        for g in ids:
            d = rng.normal(size=(n, n)).astype(np.float32)
            nm = rng.uniform(0.5, 2, (n, n)).astype(np.float32)
            cm = rng.random((n, n)) < 0.02
            store[f"{rel}/data/{g}"] = d
            store[f"{rel}/noisemap/{g}"] = nm
            store[f"{rel}/cosmicsmask/{g}"] = cm
            truth[(rel, g)] = (d, nm, cm)
        store[f"{rel}/psf_ab/narrow_psf"] = rng.random((2 * n, 2 * n)).astype(np.float32)
        frames.append(dict(id=f + 1, image_relpath=rel))
    return frames, ids, truth


@pytest.mark.parametrize("kind", ['memory', 'counting'])
def test_gather_psf_batch_matches_per_dataset_reads(kind):
    rng = np.random.default_rng(0)
    counter = dict(group=0, getitem=0, read_direct=0)
    store = MemoryStore() if kind == 'memory' else CountingGroup(counter)
    F, S, n = 4, 3, 8
    frames, ids, truth = _fill(store, F, S, n, rng)
    ragged = [ids, ids[:2], [], ids[1:]]                          # ragged star lists, one empty frame
    counter.update(group=0, getitem=0, read_direct=0)
    stager = stamp_store.StampStager()
    data, noise, cosmic, off = stamp_store.gather_psf_batch(store, frames, ragged, stager)
    assert list(off) == [0, 3, 5, 5, 7] and data.shape == (7, n, n) and data.dtype == np.float32 and cosmic.dtype == bool
    pos = 0
    for fr, lst in zip(frames, ragged):
        for g in lst:
            d, nm, cm = truth[(fr['image_relpath'], g)]
            assert np.array_equal(data[pos], d) and np.array_equal(noise[pos], nm) and np.array_equal(cosmic[pos], cm)
            pos += 1
    if kind == 'counting':
        # float32 stamps go through read_direct (no intermediate array); groups are resolved per FRAME: the frame group (2 path
        # components) + its three sub-groups for each of the 3 non-empty frames (+ the stamp-side probe), not per (frame, star)
        assert counter['read_direct'] == 2 * 7 and counter['getitem'] == 7
        assert counter['group'] <= 3 * (2 + 3) + 3
    # the staging buffers are reused by the next batch of the same size
    buf = stager.data
    stamp_store.gather_psf_batch(store, frames[:1], [ids], stager)
    assert stager.data is buf


def test_gather_photometry_batch_reads_each_psf_once():
    rng = np.random.default_rng(1)
    counter = dict(group=0, getitem=0, read_direct=0)
    store = CountingGroup(counter)
    F, S, n = 5, 4, 8
    frames, ids, truth = _fill(store, F, S, n, rng)
    star_frames = [(ids[0], frames), (ids[1], frames[1:4]), (ids[2], []), (ids[3], frames[::2])]
    counter.update(group=0, getitem=0, read_direct=0)
    data, noise, cosmic, psfs, psf_index, off = stamp_store.gather_photometry_batch(store, star_frames, lambda fid: 'psf_ab')
    assert list(off) == [0, 5, 8, 8, 11] and psfs.shape == (5, 2 * n, 2 * n)
    B = 11
    # every frame's PSF read once (5 reads for 11 items), stamps through read_direct
    assert counter['read_direct'] == 2 * B and counter['getitem'] == B + 5
    pos = 0
    for gid, frs in star_frames:
        for fr in frs:
            d, nm, cm = truth[(fr['image_relpath'], gid)]
            assert np.array_equal(data[pos], d) and np.array_equal(noise[pos], nm) and np.array_equal(cosmic[pos], cm)
            assert np.array_equal(psfs[psf_index[pos]], store[f"{fr['image_relpath']}/psf_ab/narrow_psf"][...])
            pos += 1
    empty = stamp_store.gather_photometry_batch(store, [(ids[0], [])], lambda fid: 'psf_ab')
    assert empty[0].shape[0] == 0 and list(empty[-1]) == [0, 0]


def test_batched_writes():
    store = MemoryStore()
    store['frames/a.fits/data/1'] = np.zeros((4, 4), np.float32)
    fr = dict(id=3, image_relpath='frames/a.fits')
    stamp_store.write_psf_products(store, fr, 'psf_ab', np.ones((8, 8)), 2 * np.ones((8, 8)), 2, {'dilation_x': np.zeros(3)})
    stamp_store.write_psf_products(store, fr, 'psf_ab', 3 * np.ones((8, 8)), 4 * np.ones((8, 8)), 2, {})      # replaces the group
    g = store['frames/a.fits/psf_ab']
    assert g['narrow_psf'][...].mean() == 3 and int(g['subsampling_factor'][...][0]) == 2 and list(g['distortion'].keys()) == []
    db = sqlite3.connect(':memory:')
    db.execute(PSFS_DDL)
    rows = [(i, 1.0 + i, 0.1, 'psf_ab', 7, 2, 0.8) for i in range(50)]
    stamp_store.replace_psf_rows(db, rows)
    stamp_store.replace_psf_rows(db, [(0, 9.0, 0.1, 'psf_ab', 7, 2, 0.8)])        # REPLACE semantics
    assert db.execute("SELECT COUNT(*), MAX(chi2) FROM PSFs").fetchone() == (50, 50.0)
    assert db.execute("SELECT chi2 FROM PSFs WHERE frame_id = 0").fetchone()[0] == 9.0


def test_real_h5py_store(tmp_path):
    h5py = pytest.importorskip('h5py')
    rng = np.random.default_rng(2)
    path = tmp_path / 'regions.h5'
    with h5py.File(path, 'w') as f:
        frames, ids, truth = _fill(f, 3, 3, 8, rng)
    with stamp_store.open_h5(path, 'r') as f:
        data, noise, cosmic, off = stamp_store.gather_psf_batch(f, frames, [ids] * 3)
        assert np.array_equal(data[4], truth[(frames[1]['image_relpath'], ids[1])][0])
    with stamp_store.open_h5(path, 'r+') as f:
        stamp_store.write_psf_products(f, frames[0], 'psf_xy', np.ones((16, 16)), np.ones((16, 16)), 2, {'shear': np.zeros(3)})
        assert f[f"{frames[0]['image_relpath']}/psf_xy/distortion/shear"][...].shape == (3,)


def test_open_h5_is_import_guarded():
    try:
        import h5py  # noqa: F401
        pytest.skip("h5py present")
    except ImportError:
        with pytest.raises(ImportError, match="h5py"):
            stamp_store.open_h5('/nonexistent/regions.h5')
