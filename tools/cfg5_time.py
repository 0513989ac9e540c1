"""Device timing of the PSF fit and the photometry at cfg5 shapes (64x64 stamps, subsampling 3, 30 stars) on a
small batch and few iterations; prints us per iteration per frame and the implied frames/s at T2=3000 / T=2000."""
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
from lightcurver_b200 import engine, synthetic


def ev_time(fn, reps=2):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)


def main():
    F = int(sys.argv[1]) if len(sys.argv) > 1 else 148
    T2 = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    N, n, k = 30, 64, 3
    d = synthetic.make_psf_frames(F, N, n, k, seed=synthetic.SEEDS['cfg5'])
    sc = d['data'].max() / 100
    data = torch.as_tensor(d['data'] / sc).reshape(F * N, n, n).cuda()
    nm = torch.as_tensor(d['noisemap'] / sc).reshape(F * N, n, n).cuda()
    w = (torch.as_tensor(d['masks']).reshape(F * N, n, n).cuda() / nm ** 2).contiguous()
    off = torch.arange(F + 1, dtype=torch.int32).cuda() * N
    a0 = data.sum((-1, -2))   # block-sum convention: amplitude = pixel-sum flux
    mof = torch.tensor([[3.5, 3.5, 0.0, 2.5, 1.0]]).repeat(F, 1).cuda()
    res = {}
    for T in (T2, 2 * T2):
        res[T] = ev_time(lambda: engine.psf_fit_batch(data, w, off, k, mof, a0, n_iter_analytic=0, n_iter_adabelief=T,
                                                       noise_weights=False, lam_scales=1.0, lam_hf=1.0, want=('narrow_psf', 'chi2')))
    per_it = (res[2 * T2] - res[T2]) / T2 / F * 1e3
    from lightcurver_b200 import _lib
    _lib.profile_enable(True)
    engine.psf_fit_batch(data, w, off, k, mof, a0, n_iter_analytic=0, n_iter_adabelief=T2, noise_weights=False, lam_scales=1.0, lam_hf=1.0, want=('narrow_psf', 'chi2'))
    print('kernels:', _lib.profile_summary()); _lib.profile_enable(False)
    print(f"cfg5 psf: F={F} T2={T2}: {res[T2]:.1f} ms, {res[2*T2]:.1f} ms -> {per_it:.1f} us / iteration / frame (batch of {F}); "
          f"3000 its -> {1e6 / (per_it * 3000):.2f} frames/s", flush=True)
    ms = ev_time(lambda: engine.psf_fit_batch(data, w, off, k, mof, a0, n_iter_analytic=30, n_iter_adabelief=0, noise_weights=True,
                                               want=('narrow_psf', 'chi2')))
    print(f"cfg5 stage 1 (30 LM its) + W: {ms:.1f} ms for {F} frames", flush=True)
    out = engine.psf_fit_batch(data, w, off, k, mof, a0, n_iter_analytic=30, n_iter_adabelief=T2, noise_weights=True, want=('narrow_psf', 'chi2'))
    B = F * N
    idx = torch.arange(F, dtype=torch.int32).repeat_interleave(N).cuda()
    wph = (1.0 / nm ** 2).contiguous()
    rp = {}
    for T in (50, 100):
        rp[T] = ev_time(lambda: engine.phot_fit_batch(data, wph, out['narrow_psf'], idx, out['a'], k, T, want_residuals=False, want_loss_hist=False))
    pit = (rp[100] - rp[50]) / 50 / B * 1e3
    print(f"cfg5 phot: B={B}: {pit:.2f} us / iteration / item -> {1e6 / (pit * 2000 * N):.2f} frames/s at T=2000, {N} stars", flush=True)


if __name__ == '__main__':
    main()
