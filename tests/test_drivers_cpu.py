"""Host logic of the batched pipeline drivers (no GPU): stamp-store stand-in, mask / NaN policies, products."""
import numpy as np

from lightcurver_b200.processes.psf_modelling import (MemoryStore, prepare_psf_inputs, relative_loss_differential)


def test_memory_store_h5py_subset():
    st = MemoryStore()
    st['frames/a.fits/data/123'] = np.arange(4.0).reshape(2, 2)
    assert st['frames/a.fits/data/123'][...].shape == (2, 2)
    g = st['frames/a.fits']
    assert 'data' in g.keys() and 'frames/a.fits/data/123' in st and 'frames/a.fits/nope' not in st
    pg = g.create_group('psf_ab')
    pg['narrow_psf'] = np.ones((4, 4))
    assert st['frames/a.fits/psf_ab/narrow_psf'][...].sum() == 16
    del g['psf_ab']
    assert 'psf_ab' not in g.keys()


def test_prepare_psf_inputs_policies():
    """psf_modelling.py:135-153: NaN policy only where BOTH are NaN; >40 % masked stars are dropped."""
    rng = np.random.default_rng(0)
    d = rng.random((3, 10, 10)); nm = np.full((3, 10, 10), 0.1)
    cos = np.zeros((3, 10, 10), bool)
    d[0, 1, 1] = np.nan; nm[0, 1, 1] = np.nan          # both NaN -> data 0, noise 1, masked
    d[0, 2, 2] = np.nan                               # only data NaN -> untouched by this policy
    cos[1, :5, :] = True                              # 50 % masked -> dropped
    cos[2, :4, :] = True                              # exactly 40 % -> kept (strict >)
    dd, nn, mk, keep = prepare_psf_inputs(d, nm, cos, automatic_masks=np.ones((3, 10, 10), bool))
    assert list(keep) == [0, 2] and dd.shape[0] == 2
    assert dd[0, 1, 1] == 0 and nn[0, 1, 1] == 1 and not mk[0, 1, 1]
    assert np.isnan(dd[0, 2, 2]) and mk[0, 2, 2]
    assert (~mk[1]).sum() == 40


def test_relative_loss_differential():
    lh = np.concatenate([np.linspace(100, 10, 90), np.linspace(10, 9, 10)])
    assert abs(relative_loss_differential(lh) - 1.0 / 90.0) < 1e-12
