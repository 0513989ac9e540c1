"""Per-phase cycle breakdown of the PSF-fit kernels (needs the -DLCB_PHASE_TIMERS build:
python -m lightcurver_b200.build --timers; LCB_LIBRARY=lightcurver_b200/liblcb_timers.so python tools/phase_time.py [cfg2|cfg5|cluster]).

cfg2: k_psf_fit<2,12,32> (148 frames x 10 stars x 32^2, k = 2); cfg5: k_psf_fit<3,12,0> (30 stars x 64^2, k = 3, planes in L2);
cluster: k_psf_fit_cl (LCB_PSF_CLUSTER=1, 18 frames of the cfg5 shape)."""
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
from lightcurver_b200 import engine, synthetic
mode = sys.argv[1] if len(sys.argv) > 1 else 'cfg2'
F, N, n, k, T2 = {'cfg2': (148, 10, 32, 2, 200), 'cfg5': (148, 30, 64, 3, 10), 'cluster': (18, 30, 64, 3, 20)}[mode]
names = (['halo+taps+zero', 'star loop', 'syncA+fold', 'starlet fwd', 'starlet bwd', 'grad+syncB', 'update+syncC', '-'] if mode == 'cluster'
         else ['taps+zero', f'pass1 (x{N})', f'pass2 (x{N})', f'pass2T (x{N})', f'pass1T (x{N})', 'starlet', 'update', '-'])
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
ph = out['loss_hist'][:, :8].cpu().numpy().mean(0) / (T2 if mode == 'cluster' else T2 + 1)
tot = ph.sum()
for nm_, c in zip(names, ph):
    print(f"{nm_:16s} {c:9.0f} cycles/iter  {100 * c / tot:5.1f}%")
print(f"total {tot:.0f} cycles/iter = {tot / 1.965e3:.1f} us at 1965 MHz")
