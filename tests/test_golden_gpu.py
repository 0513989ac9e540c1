"""CUDA path (through the C ABI) against the committed golden vectors of tests/golden (1e-5 relative)."""
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).parent / 'golden'


def test_golden_phot(cuda_device):
    from lightcurver_b200 import engine
    g = np.load(GOLD / 'phot_n16_k2.npz')
    B = g['data'].shape[0]
    out = engine.phot_fit_batch(g['data'], g['weight'], g['psf'], np.arange(B, dtype=np.int32), g['a'], int(g['k']), 1,
                                dx0=g['dx'], dy0=g['dy'], want_grad0=True)
    np.testing.assert_allclose(out['loss0'], g['loss'], rtol=1e-5)
    np.testing.assert_allclose(out['grad0'], g['grad'], rtol=2e-5, atol=1e-5 * np.abs(g['grad']).max())


def test_golden_psf(cuda_device):
    from lightcurver_b200 import engine
    from oracle import starred_model as sm
    g = np.load(GOLD / 'psf_n16_k2.npz')
    n, k = int(g['n']), int(g['k'])
    N = g['data'].shape[0]
    # the golden s_fixed is Moffat(3.1, 3.4, 0.5, 2.7); the library rebuilds it from the parameters
    moffat = np.array([[3.1, 3.4, 0.5, 2.7, 1.0]])
    out = engine.psf_fit_batch(g['data'], g['weight'], np.array([0, N], np.int32), k, moffat, g['a'], g['x0'], g['y0'],
                               background0=g['b'][None], W=g['W'][None], n_iter_analytic=0, n_iter_adabelief=1,
                               lam_scales=float(g['lam_scales']), lam_hf=float(g['lam_hf']),
                               want=('loss0', 'grad_b0', 'grad_s0'))
    np.testing.assert_allclose(out['loss0'][0], g['loss'], rtol=1e-5)
    np.testing.assert_allclose(out['grad_b0'][0], g['grad_b'], rtol=1e-5, atol=1e-5 * np.abs(g['grad_b']).max())
    np.testing.assert_allclose(out['grad_s0'], g['grad_s'], rtol=2e-5, atol=1e-5 * np.abs(g['grad_s']).max())


def test_golden_psf_with_field_distortion(cuda_device):
    from lightcurver_b200 import engine
    g = np.load(GOLD / 'psfdist_n16_k2.npz')
    n, k = int(g['n']), int(g['k'])
    N = g['data'].shape[0]
    moffat = np.array([[3.0, 3.3, 0.3, 2.9, 1.0]])
    out = engine.psf_fit_batch(g['data'], g['weight'], np.array([0, N], np.int32), k, moffat, g['a'], g['x0'], g['y0'],
                               background0=g['b'][None], W=g['W'][None], n_iter_analytic=0, n_iter_adabelief=1,
                               lam_scales=float(g['lam_scales']), lam_hf=float(g['lam_hf']),
                               want=('loss0', 'grad_b0', 'grad_s0', 'grad_dist0'),
                               field_distortion=1, stamp_xy=g['xy'], distortion0=g['theta'][None])
    np.testing.assert_allclose(out['loss0'][0], g['loss'], rtol=1e-5)
    np.testing.assert_allclose(out['grad_b0'][0], g['grad_b'], rtol=1e-5, atol=1e-5 * np.abs(g['grad_b']).max())
    np.testing.assert_allclose(out['grad_s0'], g['grad_s'], rtol=2e-5, atol=1e-5 * np.abs(g['grad_s']).max())
    np.testing.assert_allclose(out['grad_dist0'][0], g['grad_theta'], rtol=2e-5, atol=1e-5 * np.abs(g['grad_theta']).max())
