import sys
import numpy as np
import torch
sys.path.insert(0, '.')
from lightcurver_b200 import engine, synthetic
F, N, n, k = 64, 10, 32, 2
d = synthetic.make_psf_frames(F, N, n, k)
sc = d['data'].max() / 100
data = torch.as_tensor(d['data'] / sc).reshape(F * N, n, n).cuda()
nm = torch.as_tensor(d['noisemap'] / sc).reshape(F * N, n, n).cuda()
w = (torch.as_tensor(d['masks']).reshape(F * N, n, n).cuda() / nm ** 2).contiguous()
off = torch.arange(F + 1, dtype=torch.int32).cuda() * N
a0 = data.sum((-1, -2))   # block-sum convention: amplitude = pixel-sum flux
mof = torch.tensor([[3.0, 3.0, 0.0, 2.5, 1.0]]).repeat(F, 1).cuda()
truth = torch.as_tensor(d['psf']).cuda()
for T in (1000, 3000):
    for lr in (0.0, 1e-6, 3e-6, 1e-5, 3e-5, 1e-4, 1e-3):
        for lam in (1.0,):
            out = engine.psf_fit_batch(data, w, off, k, mof, a0, n_iter_analytic=100, n_iter_adabelief=T if lr > 0 else 1,
                                       lr=max(lr, 1e-12), noise_weights=True, lam_scales=lam, lam_hf=lam, want=('narrow_psf', 'chi2', 'loss_hist'))
            err = ((out['narrow_psf'] - truth).abs().amax((-1, -2)) / truth.amax((-1, -2)))
            lh = out['loss_hist']
            print(f"T={T} lr={lr:g} lam={lam}: chi2 med {float(out['chi2'].median()):.3f} max {float(out['chi2'].max()):.3f}; psf err/peak med {float(err.median()):.4f}; loss first {float(lh[:,0].median()):.1f} last {float(lh[:,-1].median()):.1f}", flush=True)
