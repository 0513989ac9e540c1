"""Per-phase cycle breakdown of k_psf_fit (needs the -DLCB_PHASE_TIMERS build: LCB_LIBRARY=...liblcb_timers.so)."""
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
from lightcurver_b200 import engine, synthetic
F, N, n, k, T2 = 148, 30, 64, 3, 10
d = synthetic.make_psf_frames(F, N, n, k)
sc = d['data'].max() / 100
data = torch.as_tensor(d['data'] / sc).reshape(F * N, n, n).cuda()
nm = torch.as_tensor(d['noisemap'] / sc).reshape(F * N, n, n).cuda()
w = (torch.as_tensor(d['masks']).reshape(F * N, n, n).cuda() / nm ** 2).contiguous()
off = torch.arange(F + 1, dtype=torch.int32).cuda() * N
a0 = data.sum((-1, -2))   # block-sum convention: amplitude = pixel-sum flux
mof = torch.tensor([[3.0, 3.0, 0.0, 2.5, 1.0]]).repeat(F, 1).cuda()
out = engine.psf_fit_batch(data, w, off, k, mof, a0, n_iter_analytic=0, n_iter_adabelief=T2, lr=1e-5,
                           noise_weights=True, lam_scales=1.0, lam_hf=1.0, want=('loss_hist',))
ph = out['loss_hist'][:, :8].cpu().numpy().mean(0) / (T2 + 1)
names = ['taps+zero', 'pass1 (x10)', 'pass2 (x10)', 'pass2T (x10)', 'pass1T (x10)', 'starlet', 'update', '-']
tot = ph.sum()
for nm_, c in zip(names, ph):
    print(f"{nm_:14s} {c:9.0f} cycles/iter  {100 * c / tot:5.1f}%")
print(f"total {tot:.0f} cycles/iter = {tot / 1.965e3:.1f} us at 1965 MHz")
