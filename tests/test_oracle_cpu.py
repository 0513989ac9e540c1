"""CPU tests of the oracle itself (no GPU): internal consistency of the restated STARRED model.

The reference holds no golden vector for this path (parity unpinned, see oracle/__init__.py); what
can be pinned on CPU is that the restatement is self-consistent: banded form == general
deconvolution form, FFT == direct convolution, starlet reconstruction, optimiser semantics of
optax.scale_by_belief on a known-answer case, and the committed golden vectors under tests/golden.
"""
import json
from pathlib import Path

import numpy as np
import torch

from oracle import starred_model as sm
from oracle.conventions import Conventions, DEFAULT

GOLD = Path(__file__).parent / 'golden'


def test_conventions_twins_identical():
    from lightcurver_b200.conventions import Conventions as P
    assert P().as_dict() == Conventions().as_dict()


def test_banded_equals_general_deconvolution_model():
    n, k, E = 16, 2, 3
    nu = n * k
    psf = sm.moffat_image(torch.tensor([3.0] * E), torch.tensor([3.3] * E), torch.tensor([0.3] * E), torch.tensor([2.5] * E), n, k)
    a = torch.tensor([1.0, 2.0, 3.0], dtype=torch.float64)
    dx = torch.tensor([0.2, -0.7, 1.3], dtype=torch.float64)
    dy = torch.tensor([-0.4, 0.1, 0.6], dtype=torch.float64)
    z1, zE = torch.zeros(1, dtype=torch.float64), torch.zeros(E, dtype=torch.float64)
    m1 = sm.phot_models(psf, a, dx, dy, n, k)
    m2 = sm.deconv_model(torch.zeros(nu, nu, dtype=torch.float64), zE, a[:, None], z1, z1, dx, dy, zE, psf, n, k, with_h=False)
    m3 = sm.deconv_model(torch.zeros(nu, nu, dtype=torch.float64), zE, a[:, None], z1, z1, dx, dy, zE, psf, n, k, with_h=False, direct=True)
    assert (m1 - m2).abs().max() < 1e-14 and (m2 - m3).abs().max() < 1e-14
    # block-sum convention (default): the amplitude IS the pixel-sum flux when nothing falls off the stamp -- the scale
    # relation of the reference's notebook (example_roi_modelling.ipynb cells 13 -> 21 -> 36: a ~ sum of pixels / scale)
    np.testing.assert_allclose(m1.sum((-1, -2)).numpy(), a.numpy() / sm.DEFAULT.amplitude_per_flux(k), rtol=2e-2)
    assert sm.DEFAULT.amplitude_per_flux(k) == 1.0


def test_fft_equals_direct_with_background_and_rotation():
    n, k, E = 12, 2, 2
    nu = n * k
    torch.manual_seed(0)
    h = torch.rand(nu, nu, dtype=torch.float64)
    psf = sm.moffat_image(torch.tensor([3.0] * E), torch.tensor([2.7] * E), torch.tensor([0.1] * E), torch.tensor([3.0] * E), n, k)[:, 2:-2, 2:-2]
    args = (h, torch.tensor([0.1, -0.2], dtype=torch.float64), torch.tensor([[1.0, 2.0], [1.5, 0.5]], dtype=torch.float64),
            torch.tensor([-2.0, 2.5], dtype=torch.float64), torch.tensor([1.0, -1.5], dtype=torch.float64),
            torch.tensor([0.3, -0.6], dtype=torch.float64), torch.tensor([-0.2, 0.9], dtype=torch.float64),
            torch.tensor([0.0, 0.15], dtype=torch.float64), psf, n, k)
    assert (sm.deconv_model(*args) - sm.deconv_model(*args, direct=True)).abs().max() < 1e-13


def test_starlet_reconstruction_and_adjoint_identity():
    torch.manual_seed(1)
    b = torch.rand(24, 24, dtype=torch.float64, requires_grad=True)
    al, c = sm.starlet(b)
    assert al.shape[0] == 4
    assert (al.sum(0) + c - b).abs().max() < 1e-14
    # <Phi b, y> == <b, Phi^T y> with Phi^T from autograd
    y = torch.rand_like(al)
    (g,) = torch.autograd.grad((al * y).sum(), b)
    b2 = torch.rand(24, 24, dtype=torch.float64)
    al2, _ = sm.starlet(b2)
    assert abs(float((al2 * y).sum()) - float((b2 * g).sum())) < 1e-10


def test_adabelief_matches_optax_formulas_known_answer():
    """Two steps of scale_by_belief (b1=.9, b2=.999, eps=eps_root=1e-16) worked by hand."""
    p = torch.tensor([1.0], dtype=torch.float64)
    opt = sm.AdaBelief([p], lr=0.1, n_iter=10, schedule=False)
    g1 = torch.tensor([2.0], dtype=torch.float64)
    opt.step([g1])
    mu = 0.2; s = 0.001 * (2.0 - 0.2) ** 2 + 1e-16
    exp1 = 1.0 - 0.1 * (mu / 0.1) / (np.sqrt(s / 0.001) + 1e-16)
    assert abs(float(p) - exp1) < 1e-12
    g2 = torch.tensor([-1.0], dtype=torch.float64)
    opt.step([g2])
    mu2 = 0.9 * mu + 0.1 * -1.0
    s2 = 0.999 * s + 0.001 * (-1.0 - mu2) ** 2 + 1e-16
    exp2 = exp1 - 0.1 * (mu2 / (1 - 0.81)) / (np.sqrt(s2 / (1 - 0.999 ** 2)) + 1e-16)
    assert abs(float(p) - exp2) < 1e-12


def test_clip_and_schedule():
    p = torch.zeros(3, dtype=torch.float64)
    opt = sm.AdaBelief([p], lr=1e-3, n_iter=100, schedule=True)
    g = torch.tensor([3.0, 4.0, 0.0], dtype=torch.float64)       # norm 5 -> clipped to norm 1
    opt.step([g])
    # first step of AdaBelief is -lr * g/(0.9|g|): independent of the clip scale, schedule gives lr0 at t=0
    np.testing.assert_allclose(p.numpy()[:2], [-1e-3 / 0.9, -1e-3 / 0.9], rtol=1e-9)


def test_phot_fit_recovers_flux_cpu():
    from lightcurver_b200 import synthetic
    n, k = 16, 2
    d = synthetic.make_phot_frames(2, 2, n, k, seed=5)
    data = d['data'].reshape(-1, n, n)
    sc = data.max()
    w = sc ** 2 / d['noisemap'].reshape(-1, n, n).astype(np.float64) ** 2
    a0 = data.sum((-1, -2)) * sm.DEFAULT.amplitude_per_flux(k) / sc
    r = sm.fit_phot(np.repeat(d['psf'], 2, 0), data / sc, w, a0, n, k, 300, dtype=torch.float64)
    truth = (d['transparency'][:, None] * d['star_flux'][None]).reshape(-1)
    flux = r['a'] * sc / sm.DEFAULT.amplitude_per_flux(k)
    assert np.all(np.abs(flux - truth) < 6 * r['sigma_a'] * sc / sm.DEFAULT.amplitude_per_flux(k))
    assert r['loss_hist'][:, -1].sum() < r['loss_hist'][:, 0].sum()


def test_golden_vectors():
    """Committed golden vectors (tests/golden/*.npz, generated by tools/make_golden.py from the oracle
    in float64) pin the oracle against silent edits."""
    files = sorted(f for f in GOLD.glob('*.npz') if not f.name.startswith('reference_'))    # reference_*: tests/test_reductions_*
    assert files, "tests/golden is empty"
    for f in files:
        g = np.load(f)
        kind = str(g['kind'])
        n, k = int(g['n']), int(g['k'])
        if kind == 'phot':
            L, gr = sm.phot_loss_grad(g['psf'], g['data'], g['weight'], g['a'], g['dx'], g['dy'], n, k)
            np.testing.assert_allclose(L, g['loss'], rtol=1e-12)
            np.testing.assert_allclose(np.stack(gr, -1), g['grad'], rtol=1e-10, atol=1e-12)
        elif kind == 'psf':
            L, gr = sm.psf_loss_grad(g['s_fixed'], g['b'], g['a'], g['x0'], g['y0'], g['data'], g['weight'], g['W'], n, k,
                                     float(g['lam_scales']), float(g['lam_hf']))
            np.testing.assert_allclose(L, g['loss'], rtol=1e-12)
            np.testing.assert_allclose(gr[0], g['grad_b'], rtol=1e-9, atol=1e-12)
            np.testing.assert_allclose(np.stack(gr[1:], -1), g['grad_s'], rtol=1e-9, atol=1e-12)
        elif kind == 'psfdist':
            L, gr = sm.psf_loss_grad(g['s_fixed'], g['b'], g['a'], g['x0'], g['y0'], g['data'], g['weight'], g['W'], n, k,
                                     float(g['lam_scales']), float(g['lam_hf']), theta=g['theta'], xy=g['xy'])
            np.testing.assert_allclose(L, g['loss'], rtol=1e-12)
            np.testing.assert_allclose(gr[0], g['grad_b'], rtol=1e-9, atol=1e-12)
            np.testing.assert_allclose(np.stack(gr[1:4], -1), g['grad_s'], rtol=1e-9, atol=1e-12)
            np.testing.assert_allclose(gr[4], g['grad_theta'], rtol=1e-9, atol=1e-12)
        else:
            raise AssertionError(f"unknown golden vector kind {kind!r} in {f.name}")


def test_pts_source_and_flux_uniformity_closed_forms():
    """Hand-derived forms used by k_deconv_epoch / k_deconv_update against autograd, float64:
    the first starlet scale of a separable Gaussian is g_y g_x - (B g_y)(B g_x) with B the edge-replicated B3 filter, so
    dR/dtheta = sum_x T[x] alpha_0(dp/dtheta)[x] with T = lam W_0 sign(alpha_0(p)) (overlapping windows counted once);
    flux uniformity: d/da_e [lam std/|mean|] = A (a_e - mean) - B."""
    import numpy as np
    import torch
    from oracle import starred_model as sm
    from oracle.conventions import DEFAULT as cv
    E,n,k,M,P=1,12,2,3,24
    nu=n*k
    rng=np.random.default_rng(0)
    t=lambda v: torch.tensor(v,dtype=torch.float64,requires_grad=True)
    a=t(rng.uniform(1,2,(E,M))); cx=t(np.array([-5.2,-4.0,4.9])); cy=t(np.array([-5.4,-3.1,0.3])); dx=t(rng.uniform(-1,1,E)); dy=t(rng.uniform(-1,1,E))
    alpha=torch.tensor([0.1],dtype=torch.float64)
    W=torch.tensor(rng.uniform(0.5,2,(4,nu,nu)))
    lam=0.7
    L=sm.pts_source_l1(a,cx,cy,dx,dy,alpha,n,k,P,W,lam,cv)
    gr=torch.autograd.grad(L,[a,cx,cy,dx,dy])
    # closed form as in the kernel
    G=cv.gauss_taps; sig=sm.gauss_sigma(cv)
    uc,vc,ctr=sm.deconv_positions(cx.detach(),cy.detach(),dx.detach(),dy.detach(),alpha,k,nu,P)
    uc=uc[0].numpy(); vc=vc[0].numpy()
    B=np.array([1,4,6,4,1])/16
    def taps(pc):
        ic=int(np.floor(pc+0.5)); w0=ic-G//2+1
        u=np.arange(w0,w0+G); x=u-pc
        g=np.exp(-x*x/(2*sig*sig))/(np.sqrt(2*np.pi)*sig)
        return w0,g,x/sig**2*g
    def ext(w0,g):
        gE=np.zeros(G+4); BE=np.zeros(G+4)
        for t_ in range(G+4):
            u=w0-2+t_
            if 2<=t_<G+2: gE[t_]=g[t_-2]
            s=0
            for tt in range(-2,3):
                uu=min(max(u+tt,0),nu-1)-w0
                if 0<=uu<G: s+=B[tt+2]*g[uu]
            BE[t_]=s
        return gE,BE
    ex=[];ey=[];wx=[];wy=[]
    for m in range(M):
        w0,g,d=taps(uc[m]); wx.append(w0); ex.append(ext(w0,g)+ext(w0,d))
        w0,g,d=taps(vc[m]); wy.append(w0); ey.append(ext(w0,g)+ext(w0,d))
    av=a.detach().numpy()[0]
    loss=0; pa=np.zeros(M); pu=np.zeros(M); pv=np.zeros(M)
    GEX=G+4
    for m in range(M):
        for i in range(GEX*GEX):
            v=wy[m]-2+i//GEX; u=wx[m]-2+i%GEX
            if v<0 or v>=nu or u<0 or u>=nu: continue
            dup=any(0<=v-(wy[q]-2)<GEX and 0<=u-(wx[q]-2)<GEX for q in range(m))
            if dup: continue
            al0=0; ca=np.zeros(M);cu=np.zeros(M);cvv=np.zeros(M)
            for q in range(M):
                tv=v-(wy[q]-2); tu=u-(wx[q]-2)
                if 0<=tv<GEX and 0<=tu<GEX:
                    gxu,bxu,dxu,bdxu=[ex[q][j][tu] for j in range(4)]
                    gyv,byv,dyv,bdyv=[ey[q][j][tv] for j in range(4)]
                    ca[q]=gyv*gxu-byv*bxu; cu[q]=gyv*dxu-byv*bdxu; cvv[q]=dyv*gxu-bdyv*bxu
                    al0+=av[q]*ca[q]
            lw=lam*W[0,v,u].item()
            loss+=lw*abs(al0); T=lw*np.sign(al0)
            pa+=T*ca; pu+=T*cu; pv+=T*cvv
    assert abs(L.item() - loss) <= 1e-12 * abs(loss)
    np.testing.assert_allclose(gr[0].numpy()[0], pa, rtol=1e-10, atol=1e-13)
    ca_,sa_=np.cos(0.1),np.sin(0.1)
    GU=av*pu; GV=av*pv
    np.testing.assert_allclose(gr[1].numpy(), k * (ca_ * GU + sa_ * GV), rtol=1e-10, atol=1e-13)
    np.testing.assert_allclose(gr[2].numpy(), k * (-sa_ * GU + ca_ * GV), rtol=1e-10, atol=1e-13)
    np.testing.assert_allclose(gr[3].numpy(), [k * GU.sum()], rtol=1e-10); np.testing.assert_allclose(gr[4].numpy(), [k * GV.sum()], rtol=1e-10)
    # flux uniformity check
    a2=t(rng.uniform(1,2,(7,3)))
    for rel in (True,False):
        import dataclasses
        c2=dataclasses.replace(cv,flux_uniformity_relative=rel)
        L2=sm.flux_uniformity(a2,10.0,c2); g2=torch.autograd.grad(L2,[a2])[0].numpy()
        A_=a2.detach().numpy(); Et=7; mean=A_.mean(0); sd=A_.std(0)
        if rel: Ac=10/(Et*sd*abs(mean)); Bc=10*sd*np.sign(mean)/(Et*mean**2)
        else: Ac=10/(Et*sd); Bc=0
        assert np.abs(g2 - (Ac * (A_ - mean) - Bc)).max() < 1e-12


def test_distortion_oracle_identity_and_gradient():
    """Field distortion of the oracle (distort_psf): theta = 0 is the identity, a smooth PSF keeps its integral with the
    determinant factor, and autograd agrees with central differences in the six coefficients."""
    import torch
    rng = np.random.default_rng(0)
    n, k, N = 12, 2, 3
    nu = n * k
    yy, xx = np.mgrid[:nu, :nu] - (nu - 1) / 2
    s = torch.tensor(np.exp(-(xx ** 2 + yy ** 2) / 18.0))
    xy = torch.tensor(rng.uniform(-0.5, 0.5, (N, 2)))
    assert float((sm.distort_psf(s, torch.zeros(6, dtype=torch.float64), xy) - s[None]).abs().max()) == 0.0
    th = rng.uniform(-0.1, 0.1, 6)
    ratio = sm.distort_psf(s, torch.tensor(th), xy).sum((-1, -2)) / s.sum()
    assert float((ratio - 1).abs().max()) < 1e-2
    data, w = rng.random((N, n, n)), np.ones((N, n, n))
    b = 0.01 * rng.standard_normal((nu, nu))
    args = (s.numpy(), b, np.ones(N), np.zeros(N), np.zeros(N), data, w, None, n, k, 0.0, 0.0)
    L, g = sm.psf_loss_grad(*args, theta=th, xy=xy.numpy())
    eps = 1e-6
    for q in range(6):
        tp, tm = th.copy(), th.copy()
        tp[q] += eps
        tm[q] -= eps
        fd = (sm.psf_loss_grad(*args, theta=tp, xy=xy.numpy())[0] - sm.psf_loss_grad(*args, theta=tm, xy=xy.numpy())[0]) / (2 * eps)
        assert abs(fd - g[4][q]) < 1e-5 * max(1.0, abs(g[4][q])), (q, fd, g[4][q])


def test_rescale_image_coordinates_and_position_gather():
    """utilities/image_coordinates.py:4-25 restated in stamp_store (the product cannot import lightcurver): centre -> (0, 0),
    corners -> about +-1/2; gather_positions follows the order of gather_psf_batch."""
    from lightcurver_b200 import stamp_store
    from lightcurver_b200.processes.psf_modelling import MemoryStore
    shape = (100, 200)                                         # rows (y), columns (x)
    out = stamp_store.rescale_image_coordinates(np.array([[99.5, 49.5], [0.0, 0.0], [199.0, 99.0]]), shape)
    np.testing.assert_allclose(out, [[0, 0], [-99.5 / 200, -49.5 / 100], [99.5 / 200, 49.5 / 100]])
    store = MemoryStore()
    store['f0/frame_shape'] = np.array(shape)
    store['f0/image_pixel_coordinates/s1'] = np.array([10.0, 20.0])
    store['f0/image_pixel_coordinates/s2'] = np.array([150.0, 80.0])
    xy = stamp_store.gather_positions(store, [dict(image_relpath='f0')], [['s2', 's1']])
    np.testing.assert_allclose(xy, stamp_store.rescale_image_coordinates(np.array([[150.0, 80.0], [10.0, 20.0]]), shape), rtol=1e-6)
