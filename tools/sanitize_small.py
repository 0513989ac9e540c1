"""Small invocations of K1 (fast and generic paths), K1a, K5, K2, K3 (cluster sizes 1 and 2, with the starlet on the second
stream, reduce and update) for compute-sanitizer (racecheck / memcheck):

    compute-sanitizer --tool racecheck python tools/sanitize_small.py
    compute-sanitizer --tool memcheck  python tools/sanitize_small.py

Sizes are tiny because the tools slow kernels down by two orders of magnitude.  Prints 'sanitize_small ok' at the end."""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from lightcurver_b200 import engine, synthetic                     # noqa: E402
from lightcurver_b200.processes.roi_modelling import JointDeconvolution   # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else 'all'
rng = np.random.default_rng(0)
if which in ('all', 'psf'):
    for (n, k, N) in [(16, 2, 2), (12, 3, 2)]:                     # fast path (32-wide grid) and generic path
        F = 2
        d = synthetic.make_psf_frames(F, N, n, k, seed=1)
        sc = d['data'].max() / 100.0
        data = (d['data'] / sc).astype(np.float32).reshape(-1, n, n)
        nm = (d['noisemap'] / sc).astype(np.float32).reshape(-1, n, n)
        w = (d['masks'].reshape(-1, n, n) / nm.astype(np.float64) ** 2).astype(np.float32)
        off = (np.arange(F + 1) * N).astype(np.int32)
        mof = np.stack([d['fwhm'], d['fwhm'], np.zeros(F), np.full(F, 2.5), np.ones(F)], -1)
        out = engine.psf_fit_batch(data, w, off, k, mof, data.sum((-1, -2)), n_iter_analytic=2, n_iter_adabelief=3, lr=1e-5,
                                   noise_weights=True)
        assert np.isfinite(out['narrow_psf']).all()
        print('psf', n, k, 'ok', flush=True)
if which in ('all', 'phot'):
    n, k, F, S = 16, 2, 2, 2
    d = synthetic.make_phot_frames(F, S, n, k, seed=2)
    data = d['data'].reshape(-1, n, n)
    w = (1.0 / d['noisemap'].reshape(-1, n, n).astype(np.float64) ** 2).astype(np.float32)
    out = engine.phot_fit_batch(data, w, d['psf'], np.repeat(np.arange(F), S).astype(np.int32), data.sum((-1, -2)), k, 3)
    assert np.isfinite(out['a']).all()
    print('phot ok', flush=True)
if which in ('all', 'deconv'):
    for cs in (1, 2):
        E, n, k, M, npsf = 2, 16, 2, 2, 8
        t = synthetic.make_deconv_epochs(E, n, k, M=M, n_psf=npsf, seed=3)
        nu = n * k
        data = rng.standard_normal((E, n, n)).astype(np.float32)
        jd = JointDeconvolution(data, np.ones((E, n, n), np.float32), t['psf'], k, M)
        jd.set_cluster(cs)
        jd.set_params(h=np.zeros(nu * nu), mean=np.zeros(E), a=np.ones((E, M)), c_x=t['c_x'] / 4, c_y=t['c_y'] / 4, dx=np.zeros(E), dy=np.zeros(E),
                      alpha=np.array([0.0, 0.05]))
        jd.set_reg(1.0, 1.0, 100.0, lam_pts=0.01, lam_fu=10.0)
        jd.noise_weights()
        hist = jd.run(2, lr=1e-4, schedule=True)
        g = jd.loss_grad()
        jd.close()
        assert np.isfinite(hist).all() and np.isfinite(g['h']).all()
        print('deconv cs', cs, 'ok', flush=True)
print('sanitize_small ok')
