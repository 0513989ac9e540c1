"""Quick device-side timing of K1/K2 at reduced iteration counts (development helper)."""
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
from lightcurver_b200 import engine, synthetic


def ev_time(fn, reps=3):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)


def main():
    F, N, n, k = int(sys.argv[1]) if len(sys.argv) > 1 else 296, 10, 32, 2
    T2 = int(sys.argv[2]) if len(sys.argv) > 2 else 300
    d = synthetic.make_psf_frames(F, N, n, k)
    sc = d['data'].max() / 100
    data = torch.as_tensor(d['data'] / sc).reshape(F * N, n, n).cuda()
    nm = torch.as_tensor(d['noisemap'] / sc).reshape(F * N, n, n).cuda()
    w = (torch.as_tensor(d['masks']).reshape(F * N, n, n).cuda() / nm ** 2).contiguous()
    off = torch.arange(F + 1, dtype=torch.int32).cuda() * N
    a0 = data.sum((-1, -2))   # block-sum convention: amplitude = pixel-sum flux
    mof = torch.tensor([[3.0, 3.0, 0.0, 2.5, 1.0]]).repeat(F, 1).cuda()
    for (t1, t2, nw, lam) in [(0, T2, False, 1.0), (0, T2, False, 0.0), (100, 0, False, 1.0), (0, 1, True, 1.0), (100, T2, True, 1.0)]:
        ms = ev_time(lambda: engine.psf_fit_batch(data, w, off, k, mof, a0, n_iter_analytic=t1,
                                                   n_iter_adabelief=t2, noise_weights=nw, lam_scales=lam, lam_hf=lam,
                                                   want=('narrow_psf', 'chi2')))
        print(f"psf F={F} N={N} T1={t1} T2={t2} W={nw} lam={lam}: {ms:.2f} ms  -> {ms / max(t2, 1) / F * 148 * 1e3:.2f} us/iter/frame-slot", flush=True)
    out = engine.psf_fit_batch(data, w, off, k, mof, a0, n_iter_analytic=100, n_iter_adabelief=T2, noise_weights=True,
                               want=('narrow_psf', 'chi2', 'loss_hist_analytic'))
    print('chi2 median', float(out['chi2'].median()), 'fwhm', out['moffat'][:3].cpu().numpy())
    # photometry
    B = F * 20
    psf = out['narrow_psf']
    d3 = synthetic.make_phot_frames(F, 20, n, k)
    dat3 = torch.as_tensor(d3['data']).reshape(B, n, n).cuda()
    sc3 = dat3.max()
    w3 = (sc3 ** 2 / torch.as_tensor(d3['noisemap']).reshape(B, n, n).cuda() ** 2).contiguous()
    dat3 = dat3 / sc3
    idx = torch.arange(F, dtype=torch.int32).repeat_interleave(20).cuda()
    a3 = dat3.sum((-1, -2))
    psf3 = torch.as_tensor(d3['psf']).cuda()
    T = 200
    ms = ev_time(lambda: engine.phot_fit_batch(dat3, w3, psf3, idx, a3, k, T, want_residuals=False, want_loss_hist=False))
    print(f"phot B={B} T={T}: {ms:.2f} ms -> {B / (ms * 1e-3) * T / 2000 / 20:.1f} frames/s at T=2000, 20 stars", flush=True)


if __name__ == '__main__':
    main()
