"""A/B timing of liblcb variants: LCB_LIBRARY=<path> python tools/ab_time.py"""
import sys
import torch
sys.path.insert(0, '.')
from lightcurver_b200 import engine, synthetic
import os
F, N, n, k, T2 = int(os.environ.get("AB_F", 296)), 10, 32, 2, int(os.environ.get("AB_T", 200))
d = synthetic.make_psf_frames(F, N, n, k)
sc = d['data'].max() / 100
data = torch.as_tensor(d['data'] / sc).reshape(F * N, n, n).cuda()
nm = torch.as_tensor(d['noisemap'] / sc).reshape(F * N, n, n).cuda()
w = (torch.as_tensor(d['masks']).reshape(F * N, n, n).cuda() / nm ** 2).contiguous()
off = torch.arange(F + 1, dtype=torch.int32).cuda() * N
a0 = data.sum((-1, -2))   # block-sum convention: amplitude = pixel-sum flux
mof = torch.tensor([[3.0, 3.0, 0.0, 2.5, 1.0]]).repeat(F, 1).cuda()
res = []
for lam in (1.0, 0.0):
    fn = lambda: engine.psf_fit_batch(data, w, off, k, mof, a0, n_iter_analytic=0, n_iter_adabelief=T2, lr=1e-5,
                                      noise_weights=True, lam_scales=lam, lam_hf=lam, want=('narrow_psf', 'chi2'))
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    res.append(min(ts) / T2 / 2 * 1e3)
print(f"us/iter/frame: lam=1 {res[0]:.1f}  lam=0 {res[1]:.1f}", flush=True)
