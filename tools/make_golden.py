"""Generates tests/golden/*.npz: small seeded inputs + loss/gradient from the float64 oracle.

STARRED itself cannot be run in the build container (see oracle/__init__.py), so these vectors pin
the RESTATEMENT, not STARRED; tools/dump_starred_vectors.py produces the same files from real STARRED
wherever it is installed."""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import starred_model as sm            # noqa: E402
from lightcurver_b200 import synthetic           # noqa: E402

out = ROOT / 'tests' / 'golden'
out.mkdir(exist_ok=True, parents=True)
rng = np.random.default_rng(20260101)

n, k, F, S = 16, 2, 2, 3
d = synthetic.make_phot_frames(F, S, n, k, seed=42)
data = d['data'].reshape(-1, n, n).astype(np.float64)
sc = data.max()
w = sc ** 2 / d['noisemap'].reshape(-1, n, n).astype(np.float64) ** 2
psf = np.repeat(d['psf'], S, 0).astype(np.float64)
a = data.sum((-1, -2)) * sm.DEFAULT.amplitude_per_flux(k) / sc * rng.uniform(0.9, 1.1, F * S)
dx, dy = rng.uniform(-0.8, 0.8, F * S), rng.uniform(-0.8, 0.8, F * S)
L, g = sm.phot_loss_grad(psf, data / sc, w, a, dx, dy, n, k)
np.savez_compressed(out / 'phot_n16_k2.npz', kind='phot', n=n, k=k, psf=psf.astype(np.float32), data=(data / sc).astype(np.float32),
                    weight=w.astype(np.float32), a=a.astype(np.float32), dx=dx.astype(np.float32), dy=dy.astype(np.float32),
                    loss=None, grad=None)
# recompute with the float32-rounded inputs so that the file is self-consistent
z = np.load(out / 'phot_n16_k2.npz')
L, g = sm.phot_loss_grad(z['psf'], z['data'], z['weight'], z['a'], z['dx'], z['dy'], n, k)
np.savez_compressed(out / 'phot_n16_k2.npz', **{kk: z[kk] for kk in z.files if kk not in ('loss', 'grad')}, loss=L, grad=np.stack(g, -1))

n, k, N = 16, 2, 3
nu = n * k
d = synthetic.make_psf_frames(1, N, n, k, seed=43)
sc = d['data'].max() / 100
data = (d['data'][0] / sc).astype(np.float32)
nm = d['noisemap'][0] / sc
weight = (d['masks'][0] / nm ** 2).astype(np.float32)
s_fixed = sm.moffat_image(3.1, 3.4, 0.5, 2.7, n, k).numpy().astype(np.float32)
b = (1e-4 * rng.standard_normal((nu, nu))).astype(np.float32)
a = ((data * d['masks'][0]).sum((-1, -2)) * sm.DEFAULT.amplitude_per_flux(k)).astype(np.float32)
x0, y0 = rng.uniform(-0.6, 0.6, N).astype(np.float32), rng.uniform(-0.6, 0.6, N).astype(np.float32)
W = rng.uniform(0.5, 2.0, (4 + 1, nu, nu)).astype(np.float32)
L, g = sm.psf_loss_grad(s_fixed, b, a, x0, y0, data, weight, W, n, k, 0.7, 1.3)
np.savez_compressed(out / 'psf_n16_k2.npz', kind='psf', n=n, k=k, s_fixed=s_fixed, b=b, a=a, x0=x0, y0=y0, data=data,
                    weight=weight, W=W, lam_scales=0.7, lam_hf=1.3, loss=L, grad_b=g[0], grad_s=np.stack(g[1:], -1))

# ---- PSF loss / gradient with field distortion (dyadic coefficients and positions: every sample position of the bilinear
#      resampling is exactly representable, so the vector is well posed in float32 and float64 alike)
rng2 = np.random.default_rng(20260218)
n, k, N = 16, 2, 4
nu = n * k
d = synthetic.make_psf_frames(1, N, n, k, seed=44)
sc = d['data'].max() / 100
data = (d['data'][0] / sc).astype(np.float32)
weight = (d['masks'][0] / (d['noisemap'][0] / sc) ** 2).astype(np.float32)
s_fixed = sm.moffat_image(3.0, 3.3, 0.3, 2.9, n, k).numpy().astype(np.float32)
b = (1e-4 * rng2.standard_normal((nu, nu))).astype(np.float32)
a = ((data * d['masks'][0]).sum((-1, -2))).astype(np.float32)
x0, y0 = rng2.uniform(-0.6, 0.6, N).astype(np.float32), rng2.uniform(-0.6, 0.6, N).astype(np.float32)
W = rng2.uniform(0.5, 2.0, (5, nu, nu)).astype(np.float32)
theta = (rng2.integers(-7, 8, 6) / 128.0).astype(np.float32)
xy = (rng2.integers(-8, 9, (N, 2)) / 16.0).astype(np.float32)
L, g = sm.psf_loss_grad(s_fixed, b, a, x0, y0, data, weight, W, n, k, 0.7, 1.3, theta=theta, xy=xy)
np.savez_compressed(out / 'psfdist_n16_k2.npz', kind='psfdist', n=n, k=k, s_fixed=s_fixed, b=b, a=a, x0=x0, y0=y0, data=data,
                    weight=weight, W=W, lam_scales=0.7, lam_hf=1.3, theta=theta, xy=xy, loss=L, grad_b=g[0],
                    grad_s=np.stack(g[1:4], -1), grad_theta=g[4])
print('wrote', sorted(p.name for p in out.glob('*.npz')))
