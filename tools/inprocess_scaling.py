"""The pipeline's OWN multi-GPU entry, measured: lightcurver runs its tasks in ONE Python process
(pipeline/workflow_manager.py:201-207), so the drop-in fans out below the API -- ``build_psf_batch(..., devices=...)`` and
``star_photometry_batch(..., devices=...)``, one host thread per GPU, no collective.  Strong scaling (total work fixed):

    cfg3  10,000 frames x 20 stars x 32x32 zero-point photometry (T = 2000)     star_photometry_batch
    cfg2   1,000 frames x 10 stars x 32x32 PSF fits (T1 = 100, T2 = 3000)        build_psf_batch

    python tools/inprocess_scaling.py [--devices 1,2,4,8] [--frames3 10000] [--frames2 1000]

Prints one JSON line per device count: wall-clock of the whole call from pinned host arrays to numpy results (H2D, kernels, D2H
inside), frames/s and the efficiency against one device."""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def main():
    import torch
    from lightcurver_b200 import synthetic
    from lightcurver_b200.procedures.psf_routines import build_psf_batch
    from lightcurver_b200.processes.star_photometry import star_photometry_batch
    ap = argparse.ArgumentParser()
    ap.add_argument('--devices', default='1,2,4,8')
    ap.add_argument('--frames3', type=int, default=10000)
    ap.add_argument('--frames2', type=int, default=1000)
    ap.add_argument('--repeats', type=int, default=2)
    args = ap.parse_args()
    have = torch.cuda.device_count()
    counts = [int(c) for c in args.devices.split(',') if int(c) <= have]
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()

    def sync_all():
        for dv in range(have):
            torch.cuda.synchronize(dv)

    d3 = synthetic.make_phot_frames(args.frames3, 20, 32, 2, seed=synthetic.SEEDS['cfg3'])
    data3, nm3, psf3 = pin(d3['data']), pin(d3['noisemap']), pin(d3['psf'])
    d2 = synthetic.make_psf_frames(args.frames2, 10, 32, 2, seed=synthetic.SEEDS['cfg2'])
    F2 = args.frames2
    data2, nm2, mk2 = pin(d2['data'].reshape(-1, 32, 32)), pin(d2['noisemap'].reshape(-1, 32, 32)), pin(d2['masks'].reshape(-1, 32, 32))
    base = {}
    for nd in counts:
        devs = list(range(nd))
        res = {}
        for name, call, frames in (
                ('cfg3_star_photometry_batch', lambda: star_photometry_batch(data3, nm3, psf3, 2, n_iter=2000, want_loss_hist=False, devices=devs),
                 args.frames3),
                ('cfg2_build_psf_batch', lambda: build_psf_batch(data2, nm2, 2, masks=mk2, star_counts=[10] * F2, n_iter_analytic=100,
                                                                 n_iter_adabelief=3000, guess_method_star_position='center',
                                                                 guess_fwhm_pixels=d2['fwhm'], return_dicts=False, devices=devs), F2)):
            call()                                   # warm-up (allocations, function attributes) on every device
            sync_all()
            best = None
            for _ in range(args.repeats):
                t0 = time.perf_counter()
                out = call()
                sync_all()
                dt = time.perf_counter() - t0
                best = dt if best is None else min(best, dt)
            base.setdefault(name, best if nd == counts[0] else None)
            res[name] = dict(seconds=best, frames_per_s=frames / best,
                             efficiency_vs_first=(base[name] * counts[0] / (best * nd)) if base[name] else None)
        print(json.dumps(dict(devices=nd, api='in-process fan-out, one host thread per GPU, pinned host arrays in, numpy out', **res)), flush=True)


if __name__ == '__main__':
    main()
