"""Where does the end-to-end (host API) time go?  cfg2, one step."""
import cProfile, pstats, sys, time
import numpy as np
import torch
sys.path.insert(0, '.')
from lightcurver_b200 import synthetic
from lightcurver_b200.procedures.psf_routines import build_psf_batch
from lightcurver_b200.processes.star_photometry import star_photometry_batch
F, N, n, k = 1000, 10, 32, 2
d = synthetic.make_psf_frames(F, N, n, k)
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
data, nm, mk = pin(d['data']), pin(d['noisemap']), pin(d['masks'])


def step():
    res = build_psf_batch(data, nm, k, masks=mk, n_iter_analytic=100, n_iter_adabelief=3000,
                          guess_method_star_position='center', guess_fwhm_pixels=d['fwhm'], return_dicts=False)
    ph = star_photometry_batch(data, nm, res['narrow_psf'], k, n_iter=2000, masks=mk, want_loss_hist=False)
    return res, ph


step()
t = time.perf_counter(); step(); print('step', time.perf_counter() - t)
pr = cProfile.Profile(); pr.enable(); step(); pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(18)
