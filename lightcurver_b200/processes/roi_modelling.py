"""Joint multi-epoch deconvolution: the STARRED calls of lightcurver/processes/roi_modelling.py:213-335
(``setup_model``, ``Prior``, ``Loss``, ``propagate_noise``, ``Optimizer('adabelief').minimize``,
``model.model``, ``model.getDeconvolved``) served by liblcb's ``lcb_deconv_*`` handle API.

``joint_deconvolution`` is the single entry point a patched ``do_modelling_of_roi`` calls for its
stage 2 (and ``do_one_star_forward_modelling`` for the shared-background variants).  Epochs can be
sharded over ranks: every rank passes ITS epochs plus a ``torch.distributed`` process group, and one
all-reduce of nu^2 + 6M + 2 floats per iteration keeps the shared parameters (h, c_x, c_y)
bit-identical on all ranks.  The exchange runs inside the kernels over NVLink peer
memory (``comm='p2p'``: CUDA IPC mapped receive buffers, push + flag, summed in rank order; the process
group is only used once to exchange the 64-byte IPC handles) or as one NCCL all-reduce between the two
halves of an iteration (``comm='nccl'``).
"""
import ctypes as C

import numpy as np

from .. import _lib
from .._lib import as_f32, ptr
from ..conventions import Conventions, DEFAULT


class JointDeconvolution:
    """Thin owner of an ``lcb_deconv`` handle (host-pointer mode for inputs/outputs)."""

    def __init__(self, data, weight, psf, subsampling_factor, n_sources, conventions: Conventions = DEFAULT):
        _lib.require_device()
        from ..conventions import apply_to_library
        apply_to_library(conventions)
        self.data, self.weight, self.psf = as_f32(data), as_f32(weight), as_f32(psf)
        self.E, self.n = int(self.data.shape[0]), int(self.data.shape[-1])
        self.k, self.M, self.P = int(subsampling_factor), int(n_sources), int(self.psf.shape[-1])
        self.nu = self.n * self.k
        if tuple(self.psf.shape) != (self.E, self.P, self.P) or tuple(self.weight.shape) != tuple(self.data.shape):
            raise ValueError("data/weight must be (E,n,n) and psf (E,P,P)")
        prob = _lib.DeconvProblem(self.E, self.n, self.k, self.P, self.M, ptr(self.data), ptr(self.weight), ptr(self.psf))
        self.handle = C.c_void_p()
        _lib.check(_lib.lib.lcb_deconv_create(C.byref(prob), _lib.MEM_HOST, None, C.byref(self.handle)), 'lcb_deconv_create')
        self.J = int(_lib.lib.lcb_starlet_scales(self.nu))
        self.group, self.world, self.rank, self.comm = None, 1, 0, None
        self.E_total, self.e0 = self.E, 0

    def close(self):
        """Frees the handle.  With comm='p2p' this is a COLLECTIVE call (device synchronise + barrier: nobody may unmap a
        receive buffer a peer can still write to), so sharded users must call close() explicitly, or use the object as
        a context manager, on every rank."""
        if getattr(self, 'handle', None):
            if self.comm == 'p2p':
                import torch
                import torch.distributed as dist
                torch.cuda.synchronize()
                dist.barrier(group=self.group)
            _lib.lib.lcb_deconv_destroy(self.handle)
            self.handle = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False

    def __del__(self):
        # finalisers run at garbage-collection / interpreter-shutdown time, which differs per rank: NO collective here.
        # A connected p2p handle that was never closed is leaked on purpose (its receive buffer may still be mapped by peers).
        try:
            if getattr(self, 'handle', None) and self.comm != 'p2p':
                _lib.lib.lcb_deconv_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    # ---- multi-GPU ---------------------------------------------------------------------------
    def connect(self, group, comm='p2p'):
        """Epochs are sharded over the ranks of ``group`` (contiguous blocks in rank order, this rank holds its own).
        comm='p2p': in-kernel all-reduce over NVLink peer memory; 'nccl': one NCCL all-reduce per iteration."""
        import torch
        import torch.distributed as dist
        if comm not in ('p2p', 'nccl'):
            raise ValueError("comm must be 'p2p' or 'nccl'")
        self.group, self.world, self.rank = group, dist.get_world_size(group), dist.get_rank(group)
        counts = [None] * self.world
        dist.all_gather_object(counts, self.E, group=group)
        self.E_total, self.e0 = int(sum(counts)), int(sum(counts[:self.rank]))
        _lib.check(_lib.lib.lcb_deconv_set_global(self.handle, self.E_total, self.e0, None, _lib.MEM_HOST), 'lcb_deconv_set_global')
        self.comm = comm if self.world > 1 else None
        if self.comm == 'p2p':
            mine = C.create_string_buffer(64)
            _lib.check(_lib.lib.lcb_deconv_comm_init(self.handle, self.rank, self.world, mine), 'lcb_deconv_comm_init')
            hs = [None] * self.world
            dist.all_gather_object(hs, mine.raw, group=group)
            allh = C.create_string_buffer(b''.join(hs), 64 * self.world)
            _lib.check(_lib.lib.lcb_deconv_comm_connect(self.handle, allh), 'lcb_deconv_comm_connect')
            torch.cuda.synchronize()
            dist.barrier(group=group)
        return self

    def set_cluster(self, ctas_per_epoch=0):
        """CTAs per epoch of the per-epoch kernel (0 = automatic)."""
        _lib.check(_lib.lib.lcb_deconv_set_cluster(self.handle, int(ctas_per_epoch)), 'lcb_deconv_set_cluster')

    def _global_mean_flux(self, a):
        """Per-source mean flux over ALL epochs (the shift of the flux-uniformity sums; identical on all ranks)."""
        sm_ = np.asarray(a, np.float64).reshape(self.E, self.M).sum(0) if self.M else np.zeros(0)
        if self.world > 1:
            import torch.distributed as dist
            parts = [None] * self.world
            dist.all_gather_object(parts, sm_, group=self.group)
            sm_ = np.sum(parts, axis=0)
        return (sm_ / max(self.E_total, 1)).astype(np.float32)

    # ---- parameters -------------------------------------------------------------------------
    def set_params(self, h=None, mean=None, a=None, c_x=None, c_y=None, dx=None, dy=None, alpha=None,
                   free_h=True, free_mean=True, free_a=True, free_c=True, free_d=True):
        E, M = self.E, self.M
        arrs = dict(h=(h, self.nu * self.nu), mean=(mean, E), a=(a, E * M), c_x=(c_x, M), c_y=(c_y, M),
                    dx=(dx, E), dy=(dy, E), alpha=(alpha, E))
        keep = {}
        for nm, (v, cnt) in arrs.items():
            if v is None:
                keep[nm] = None
                continue
            v = np.ascontiguousarray(np.asarray(v, dtype=np.float32).reshape(-1))
            if v.size != cnt:
                raise ValueError(f"{nm} must have {cnt} elements (got {v.size})")
            keep[nm] = v
        q = _lib.DeconvParams(*[ptr(keep[nm]) for nm in ('h', 'mean', 'a', 'c_x', 'c_y', 'dx', 'dy', 'alpha')],
                              int(free_h), int(free_mean), int(free_a), int(free_c), int(free_d))
        _lib.check(_lib.lib.lcb_deconv_set_params(self.handle, C.byref(q), _lib.MEM_HOST), 'lcb_deconv_set_params')
        if keep['a'] is not None and self.M:
            sh = np.ascontiguousarray(self._global_mean_flux(keep['a']))
            _lib.check(_lib.lib.lcb_deconv_set_global(self.handle, self.E_total, self.e0, ptr(sh), _lib.MEM_HOST), 'lcb_deconv_set_global')

    def set_reg(self, lam_scales=0.0, lam_hf=0.0, lam_pos=0.0, W=None, prior=None, lam_pts=0.0, lam_fu=0.0,
                conventions: Conventions = DEFAULT):
        W = None if W is None else np.ascontiguousarray(W, dtype=np.float32)
        if W is not None and W.size != self.J * self.nu * self.nu:
            raise ValueError(f"W must be ({self.J},{self.nu},{self.nu})")
        pr = [None] * 4
        if prior is not None:
            pr = [np.ascontiguousarray(np.broadcast_to(np.asarray(p, dtype=np.float32), (self.M,))) for p in prior]
        r = _lib.DeconvReg(float(lam_scales), float(lam_hf), float(lam_pos), ptr(W), *[ptr(p) for p in pr],
                           float(lam_pts), float(lam_fu), int(conventions.pts_source_all_epochs),
                           int(conventions.flux_uniformity_relative))
        _lib.check(_lib.lib.lcb_deconv_set_reg(self.handle, C.byref(r), _lib.MEM_HOST), 'lcb_deconv_set_reg')

    # ---- evaluation -------------------------------------------------------------------------
    def get(self, want_model=True):
        E, M, nu, n = self.E, self.M, self.nu, self.n
        out = dict(h=np.empty(nu * nu, np.float32), mean=np.empty(E, np.float32), a=np.empty(E * M, np.float32),
                   c_x=np.empty(M, np.float32), c_y=np.empty(M, np.float32), dx=np.empty(E, np.float32),
                   dy=np.empty(E, np.float32), alpha=np.empty(E, np.float32))
        q = _lib.DeconvParams(*[ptr(out[nm]) for nm in ('h', 'mean', 'a', 'c_x', 'c_y', 'dx', 'dy', 'alpha')], 0, 0, 0, 0, 0)
        model = np.empty((E, n, n), np.float32) if want_model else None
        loss = np.empty(1, np.float32) if want_model else None
        _lib.check(_lib.lib.lcb_deconv_get(self.handle, C.byref(q), ptr(model), ptr(loss), _lib.MEM_HOST), 'lcb_deconv_get')
        if want_model:
            out['model'] = model
            out['loss_local'] = float(loss[0])
        return out

    def loss_grad(self):
        """Loss and gradient at the current parameters.  With comm='p2p' this is a collective call: the loss is the
        global one, h / c gradients are all-reduced, per-epoch gradients are those of the local epochs."""
        if self.comm == 'nccl':
            raise NotImplementedError("loss_grad over sharded epochs needs comm='p2p'")
        E, M, nu = self.E, self.M, self.nu
        g = dict(loss=np.empty(1, np.float32), h=np.empty(nu * nu, np.float32), mean=np.empty(E, np.float32),
                 a=np.empty(E * M, np.float32), c_x=np.empty(M, np.float32), c_y=np.empty(M, np.float32),
                 dx=np.empty(E, np.float32), dy=np.empty(E, np.float32))
        s = _lib.DeconvGrad(*[ptr(g[nm]) for nm in ('loss', 'h', 'mean', 'a', 'c_x', 'c_y', 'dx', 'dy')])
        _lib.check(_lib.lib.lcb_deconv_loss_grad(self.handle, C.byref(s), _lib.MEM_HOST), 'lcb_deconv_loss_grad')
        g['loss'] = float(g['loss'][0])
        return g

    # ---- optimisation -----------------------------------------------------------------------
    def reduce_tensor(self):
        """torch view (no copy) of the device buffer that must be sum-all-reduced between the halves."""
        import torch
        p, cnt = C.c_void_p(), C.c_int()
        _lib.check(_lib.lib.lcb_deconv_reduce_buffer(self.handle, C.byref(p), C.byref(cnt)), 'lcb_deconv_reduce_buffer')

        class _Ext:      # __cuda_array_interface__ holder
            pass
        ext = _Ext()
        ext.__cuda_array_interface__ = dict(shape=(cnt.value,), typestr='<f4', data=(p.value, False), version=2)
        return torch.as_tensor(ext, device='cuda')

    def noise_weights(self, group=None):
        group = self.group if group is None else group
        _lib.check(_lib.lib.lcb_deconv_noise_weights(self.handle, 0, None, _lib.MEM_HOST), 'lcb_deconv_noise_weights')
        if group is not None:
            import torch.distributed as dist
            dist.all_reduce(self.reduce_tensor(), group=group)
        W = np.empty((self.J, self.nu, self.nu), np.float32)
        _lib.check(_lib.lib.lcb_deconv_noise_weights(self.handle, 1, ptr(W), _lib.MEM_HOST), 'lcb_deconv_noise_weights')
        return W

    def run(self, n_iter, lr=1e-4, schedule=False, group=None):
        """n_iter AdaBelief iterations; returns the loss history (global loss when sharded)."""
        if group is not None and self.group is None:
            self.connect(group, 'nccl')          # historical signature: a bare group means the NCCL route
        if self.comm != 'nccl':                   # single rank, or in-kernel exchange over peer memory
            group = None
        else:
            group = self.group
        if group is None:
            hist = np.empty(n_iter, np.float32)
            opts = _lib.FitOpts(int(n_iter), float(lr), int(bool(schedule)))
            _lib.check(_lib.lib.lcb_deconv_run(self.handle, C.byref(opts), ptr(hist), _lib.MEM_HOST), 'lcb_deconv_run')
            return hist
        import torch
        import torch.distributed as dist
        red = self.reduce_tensor()
        lossidx = self.nu * self.nu + 2 * self.M
        hist = torch.empty(n_iter, device='cuda')
        for it in range(n_iter):
            _lib.check(_lib.lib.lcb_deconv_step_local(self.handle, 0), 'lcb_deconv_step_local')
            dist.all_reduce(red, group=group)                     # ONE collective per iteration
            _lib.check(_lib.lib.lcb_deconv_step_update(self.handle, it, n_iter, float(lr), int(bool(schedule))), 'lcb_deconv_step_update')
            hist[it] = red[lossidx]
        if n_iter:
            _lib.check(_lib.lib.lcb_deconv_flush(self.handle), 'lcb_deconv_flush')   # pending per-epoch update, flag cleared
        torch.cuda.synchronize()
        return hist.cpu().numpy()


def epoch_shard(E, rank, world):
    """Contiguous block partition of E epochs: returns slice(lo, hi)."""
    base, rem = divmod(E, world)
    lo = rank * base + min(rank, rem)
    return slice(lo, lo + base + (1 if rank < rem else 0))


def flux_sigma_multi(kwargs, data, noisemap, psf, subsampling_factor, conventions: Conventions = DEFAULT):
    """starred_utilities.py:10-39 for M point sources: sigma_a[e,m] = (sum_p w (dm/da_em)^2)^-1/2, from M
    model evaluations with unit amplitudes.  Returns (E*M,) epoch-major."""
    ka = kwargs['kwargs_analytic']
    E = data.shape[0]
    M = len(np.atleast_1d(ka['c_x']))
    with np.errstate(divide='ignore', invalid='ignore'):
        w = np.where(np.isfinite(noisemap) & (noisemap > 0), 1.0 / np.asarray(noisemap, np.float64) ** 2, 0.0)
    jd = JointDeconvolution(np.nan_to_num(np.asarray(data, np.float32)), w.astype(np.float32), psf, subsampling_factor, M, conventions)
    H = np.empty((E, M))
    for m in range(M):
        a = np.zeros((E, M), np.float32)
        a[:, m] = 1.0
        jd.set_params(h=np.zeros(jd.nu ** 2), mean=np.zeros(E), a=a, c_x=ka['c_x'], c_y=ka['c_y'], dx=ka['dx'], dy=ka['dy'],
                      alpha=ka.get('alpha', np.zeros(E)))
        mod = jd.get()['model'].astype(np.float64)
        H[:, m] = (w * mod * mod).sum((-1, -2)) * (1.0 if conventions.chi2_half else 2.0)
    jd.close()
    with np.errstate(divide='ignore'):
        return (1.0 / np.sqrt(H)).reshape(-1)


def _kwargs_of(fin):
    return {
        'kwargs_analytic': {'c_x': fin['c_x'].astype(np.float64), 'c_y': fin['c_y'].astype(np.float64),
                            'dx': fin['dx'].astype(np.float64), 'dy': fin['dy'].astype(np.float64),
                            'a': fin['a'].astype(np.float64), 'alpha': fin['alpha'].astype(np.float64)},
        'kwargs_background': {'h': fin['h'].astype(np.float64), 'mean': fin['mean'].astype(np.float64)},
        'kwargs_sersic': {},
    }


def _flux_sigma_from_handle(jd, fin, weight, cv):
    """flux_sigma_multi on a handle that already holds the stamps and folded PSFs (no second upload): M model evaluations with
    unit amplitudes, h = 0, mean = 0 at the fitted positions.  Overwrites the handle's parameters: call after get()."""
    E, M = jd.E, jd.M
    w = np.asarray(weight, np.float64)
    H = np.empty((E, M))
    for m in range(M):
        a = np.zeros((E, M), np.float32)
        a[:, m] = 1.0
        jd.set_params(h=np.zeros(jd.nu ** 2), mean=np.zeros(E), a=a, c_x=fin['c_x'], c_y=fin['c_y'], dx=fin['dx'], dy=fin['dy'],
                      alpha=fin['alpha'])
        mod = jd.get()['model'].astype(np.float64)
        H[:, m] = (w * mod * mod).sum((-1, -2)) * (1.0 if cv.chi2_half else 2.0)
    with np.errstate(divide='ignore'):
        return (1.0 / np.sqrt(H)).reshape(-1)


def _finish(jd, hist, Wused, data, weight, psf, cv):
    """Products of a finished fit (roi_modelling.py:335-401): kwargs_final, model, sigma of the fluxes, deconvolved epoch 0."""
    fin = jd.get()
    kw = _kwargs_of(fin)
    n, k, nu, M = jd.n, jd.k, jd.nu, jd.M
    if jd.world > 1:          # the sharded handle keeps its exchange state: evaluate sigma on a local single-rank handle
        with np.errstate(divide='ignore', invalid='ignore'):
            noisemap = np.where(weight > 0, 1.0 / np.sqrt(weight), np.inf)
        sig = flux_sigma_multi(kw, data, noisemap, psf, k, cv)
    else:
        sig = _flux_sigma_from_handle(jd, fin, weight, cv)
    # model.getDeconvolved(kwargs, 0): high-resolution scene of epoch 0 and its background only
    from .star_photometry import point_source_image
    h2 = fin['h'].reshape(nu, nu).astype(np.float64)
    deconv = h2.copy()
    for m in range(M):
        deconv += point_source_image(fin['a'][m], fin['c_x'][m] + fin['dx'][0], fin['c_y'][m] + fin['dy'][0], n, k, cv)
    return dict(kwargs_final=kw, model=fin['model'], loss_history=hist, W=Wused, flux_sigma=sig,
                deconvolved_epoch0=(deconv, h2),
                amplitude_per_flux=cv.amplitude_per_flux(k))     # kernel amplitude a = amplitude_per_flux * pixel-sum flux (1 by default)


def joint_deconvolution(data, weight, psf, subsampling_factor, xs, ys, initial_a, n_iter=2000, lr=1e-4,
                        schedule=False, alpha=None, h0=None, dx0=None, dy0=None, mean0=None,
                        free_h=True, free_mean=True, free_a=True, free_c=True, free_d=True,
                        regularization_strength_scales=1.0, regularization_strength_hf=1.0,
                        regularization_strength_positivity=100.0, W='propagate', prior=None,
                        regularization_strength_pts_source=0.0, regularization_strength_flux_uniformity=0.0,
                        conventions: Conventions = DEFAULT, group=None, comm='p2p'):
    """Stage 2 of roi_modelling.py:285-335 (defaults: lr 1e-4, no schedule/clip, strengths 1/1/100).

    data, weight (E,n,n) LOCAL epochs; psf (E,P,P); xs, ys (M,) point-source positions in data pixels
    from the stamp centre (roi_modelling.py:207-210); initial_a (E*M,) epoch-major or (E,M).
    ``prior`` = (mu_x, sigma_x, mu_y, sigma_y) Gaussian astrometric prior (roi_modelling.py:240-244).
    ``W`` = 'propagate' (SLIT weights from the model), an array (J,nu,nu), or None (== 1).
    ``group``: torch.distributed process group over which the epochs are sharded (``comm`` = 'p2p' | 'nccl').
    Returns dict(kwargs_final, model, loss_history, W, flux_sigma, deconvolved_epoch0).
    """
    cv = conventions
    E = data.shape[0]
    M = len(np.atleast_1d(xs))
    jd = JointDeconvolution(data, weight, psf, subsampling_factor, M, cv)
    if group is not None:
        jd.connect(group, comm)
    nu = jd.nu
    a0 = np.asarray(initial_a, dtype=np.float32).reshape(E, M)
    jd.set_params(h=np.zeros(nu * nu) if h0 is None else h0, mean=np.zeros(E) if mean0 is None else mean0, a=a0,
                  c_x=np.atleast_1d(xs), c_y=np.atleast_1d(ys), dx=np.zeros(E) if dx0 is None else dx0,
                  dy=np.zeros(E) if dy0 is None else dy0, alpha=np.zeros(E) if alpha is None else alpha,
                  free_h=free_h, free_mean=free_mean, free_a=free_a, free_c=free_c, free_d=free_d)
    propagate = isinstance(W, str)
    if propagate and W != 'propagate':
        raise ValueError("W must be 'propagate', an array (J,nu,nu) or None")
    Wused = None if propagate else W
    reg = dict(lam_scales=regularization_strength_scales, lam_hf=regularization_strength_hf,
               lam_pos=regularization_strength_positivity, prior=prior, lam_pts=regularization_strength_pts_source,
               lam_fu=regularization_strength_flux_uniformity, conventions=cv)
    jd.set_reg(W=Wused, **reg)
    need_W = (free_h and (regularization_strength_scales or regularization_strength_hf)) or regularization_strength_pts_source
    if propagate and need_W:
        Wused = jd.noise_weights()                    # installs the weights in the handle
    hist = jd.run(n_iter, lr=lr, schedule=schedule)
    out = _finish(jd, hist, Wused, data, weight, psf, cv)
    jd.close()
    return out


def joint_deconvolution_many(problems, subsampling_factor, n_iter=2000, lr=1e-4, schedule=False,
                             free_h=True, free_mean=True, free_a=True, free_c=True, free_d=True,
                             regularization_strength_scales=1.0, regularization_strength_hf=1.0,
                             regularization_strength_positivity=100.0, prior=None,
                             regularization_strength_pts_source=0.0, regularization_strength_flux_uniformity=0.0,
                             conventions: Conventions = DEFAULT):
    """Several INDEPENDENT joint fits in one library call (``lcb_deconv_run_many``): the reference-coupled star photometry
    (star_photometry.py:74-87 with ``starlet_global_background`` / ``uniform_background_per_epoch``) is one joint fit per star --
    shared h, c and clip norm over the star's epochs -- and the reference runs the stars one after the other (:257); here every
    star gets a handle and the iterations of all handles are interleaved on their own streams.

    problems: list of dict(data, weight, psf, xs, ys, initial_a) with the meaning of ``joint_deconvolution``; epoch counts may
    differ.  The remaining arguments are shared.  Returns the list of ``joint_deconvolution`` result dicts, in order."""
    cv = conventions
    jds, Ws = [], []
    need_W = (free_h and (regularization_strength_scales or regularization_strength_hf)) or regularization_strength_pts_source
    try:
        for pr in problems:
            E = pr['data'].shape[0]
            M = len(np.atleast_1d(pr['xs']))
            jd = JointDeconvolution(pr['data'], pr['weight'], pr['psf'], subsampling_factor, M, cv)
            jds.append(jd)
            jd.set_params(h=np.zeros(jd.nu ** 2), mean=np.zeros(E), a=np.asarray(pr['initial_a'], np.float32).reshape(E, M),
                          c_x=np.atleast_1d(pr['xs']), c_y=np.atleast_1d(pr['ys']), dx=np.zeros(E), dy=np.zeros(E), alpha=np.zeros(E),
                          free_h=free_h, free_mean=free_mean, free_a=free_a, free_c=free_c, free_d=free_d)
            jd.set_reg(lam_scales=regularization_strength_scales, lam_hf=regularization_strength_hf,
                       lam_pos=regularization_strength_positivity, W=None, prior=prior, lam_pts=regularization_strength_pts_source,
                       lam_fu=regularization_strength_flux_uniformity, conventions=cv)
            Ws.append(jd.noise_weights() if need_W else None)
        hists = [np.empty(n_iter, np.float32) for _ in jds]
        if jds:
            hs = (C.c_void_p * len(jds))(*[jd.handle for jd in jds])
            lh = (C.c_void_p * len(jds))(*[ptr(h) for h in hists])
            opts = _lib.FitOpts(int(n_iter), float(lr), int(bool(schedule)))
            _lib.check(_lib.lib.lcb_deconv_run_many(hs, len(jds), C.byref(opts), lh, _lib.MEM_HOST), 'lcb_deconv_run_many')
        return [_finish(jd, hist, W, pr['data'], pr['weight'], pr['psf'], cv) for jd, hist, W, pr in zip(jds, hists, Ws, problems)]
    finally:
        for jd in jds:
            jd.close()


def lbfgsb_translations_and_fluxes(jd, maxiter, a_lower=0.0):
    """Stage 1 of do_modelling_of_roi (roi_modelling.py:260-281): scipy L-BFGS-B over {dx, dy, a} with everything else
    fixed, exactly the reference's structure (host optimiser, device loss and gradient: ``lcb_deconv_loss_grad``).
    With sharded epochs every rank runs the same optimiser on the concatenated vector (local blocks are
    all-gathered), so the trajectories are identical on all ranks.  Returns (loss history, scipy result)."""
    from scipy.optimize import minimize
    E, M = jd.E, jd.M
    cur = jd.get(want_model=False)
    x_loc = np.concatenate([cur['dx'], cur['dy'], cur['a']]).astype(np.float64)
    nloc = x_loc.size
    if jd.world > 1:
        import torch.distributed as dist
        sizes = [None] * jd.world
        dist.all_gather_object(sizes, nloc, group=jd.group)
        offs = np.concatenate([[0], np.cumsum(sizes)])

        def gather(v):
            parts = [None] * jd.world
            dist.all_gather_object(parts, np.asarray(v, np.float64), group=jd.group)
            return np.concatenate(parts)
        x0 = gather(x_loc)
        lo = int(offs[jd.rank])
    else:
        gather = lambda v: np.asarray(v, np.float64)
        x0, lo = x_loc, 0
    hist, last = [], {}

    def fun(x):
        xl = x[lo:lo + nloc]
        jd.set_params(dx=xl[:E], dy=xl[E:2 * E], a=xl[2 * E:], free_h=False, free_mean=False, free_a=True, free_c=False, free_d=True)
        g = jd.loss_grad()
        last['L'] = g['loss']
        return g['loss'], gather(np.concatenate([g['dx'], g['dy'], g['a']]))

    # bounds (kwargs_down / kwargs_up of setup_model [R]): fluxes are non-negative, shifts stay inside the stamp
    bl = [(-jd.n / 2.0, jd.n / 2.0)] * (2 * E) + [(a_lower, None)] * (E * M)
    if jd.world > 1:
        bparts = [None] * jd.world
        import torch.distributed as dist
        dist.all_gather_object(bparts, bl, group=jd.group)
        bl = [b for part in bparts for b in part]
    res = minimize(fun, x0, jac=True, method='L-BFGS-B', bounds=bl,
                   options={'maxiter': int(maxiter), 'maxfun': 20 * int(maxiter) + 20},
                   callback=lambda xk: hist.append(last['L']))
    xl = res.x[lo:lo + nloc]
    jd.set_params(dx=xl[:E], dy=xl[E:2 * E], a=xl[2 * E:], free_h=False, free_mean=False, free_a=True, free_c=False, free_d=True)
    return np.asarray(hist), res


def lbfgs_translations_and_fluxes_device(jd, maxiter, a_lower=0.0, ftol=2.220446049250313e-09, pgtol=1e-5):
    """Stage 1 of do_modelling_of_roi (roi_modelling.py:260-281) with the optimiser on the device (``lcb_deconv_lbfgs``):
    projected L-BFGS over {dx, dy, a}, scipy's default stopping rules (ftol = 1e7 eps, pgtol = 1e-5), no host round trip per
    evaluation.  Single-rank handles.  Returns (loss history per iteration, info dict)."""
    jd.set_params(free_h=False, free_mean=False, free_a=True, free_c=False, free_d=True)
    hist = np.zeros(int(maxiter), np.float32)
    info = np.zeros(5, np.float32)
    _lib.check(_lib.lib.lcb_deconv_lbfgs(jd.handle, int(maxiter), float(a_lower), float(ftol), float(pgtol), ptr(hist), ptr(info),
                                         _lib.MEM_HOST), 'lcb_deconv_lbfgs')
    reasons = {0: 'evaluation budget exhausted', 1: 'relative reduction of the loss <= ftol', 2: 'projected gradient <= pgtol',
               3: 'iteration limit reached', 4: 'line search could not improve the loss'}
    nit = int(info[0])
    return hist[:nit].astype(np.float64), dict(nit=nit, nfev=int(info[1]), message=reasons.get(int(info[2]), '?'), fun=float(info[3]),
                                               projected_gradient=float(info[4]))


def model_roi_arrays(data, noisemap, psf, subsampling_factor, xs, ys, initial_a, angles_to_north=None,
                     fix_point_source_astrometry=False, starting_background=None, further_optimize_background=True,
                     roi_model_regularization=None, roi_deconv_translations_iters=300, roi_deconv_all_iters=2000,
                     conventions: Conventions = DEFAULT, group=None, comm='p2p', stage1='device'):
    """The two optimisation stages of do_modelling_of_roi (roi_modelling.py:213-335) on arrays that are already
    loaded and scaled (:154-170): data, noisemap (E,n,n); psf (E,P,P); xs, ys the point-source guesses in data
    pixels from the stamp centre (:207-210); initial_a (E*M,) (:211-212).

    Stage 1 (:260-281): scipy L-BFGS-B, ``roi_deconv_translations_iters`` iterations over {dx, dy, a}, chi2 +
    flux-uniformity penalty ``regularization_scatter_fluxes_pre_optim`` (default 10.0, :273) + the astrometric prior
    (constant in this stage).  Stage 2 (:285-335): SLIT noise weights, free {h (if further_optimize_background),
    mean, a, c_x, c_y, dx, dy}, AdaBelief lr 1e-4, no schedule; strengths from ``roi_model_regularization`` with the
    reference's defaults (scales 1, hf 1, positivity 100, pts_source 0.01, flux scatter 10.0; :305-312).
    ``fix_point_source_astrometry``: True fixes c, a float is the sigma (data pixels) of a Gaussian prior around the
    initial positions (:225-244).  One device handle serves both stages (the stamps are uploaded once).
    ``stage1``: 'device' (default: L-BFGS with its state on the GPU, ``lcb_deconv_lbfgs``) or 'scipy' (the reference's own
    structure: host L-BFGS-B over device loss / gradient; always used when the epochs are sharded over ranks).
    """
    reg = roi_model_regularization or {}
    cv = conventions
    E = data.shape[0]
    k = int(subsampling_factor)
    xs, ys = np.atleast_1d(np.asarray(xs, float)), np.atleast_1d(np.asarray(ys, float))
    M = len(xs)
    from .star_photometry import stamps_and_weights
    d32, weight = stamps_and_weights(data, noisemap)
    # initial_a are pixel-sum (aperture) fluxes (roi_modelling.py:199-212): amplitude = amplitude_per_flux x flux
    initial_a = np.asarray(initial_a, np.float64) * cv.amplitude_per_flux(k)
    alpha = np.zeros(E) if angles_to_north is None else np.asarray(angles_to_north, float)
    h0 = None if starting_background is None else np.asarray(starting_background, np.float32).reshape(-1)
    fix_c = isinstance(fix_point_source_astrometry, bool) and fix_point_source_astrometry
    prior = None
    if isinstance(fix_point_source_astrometry, float):
        sg = np.full(M, fix_point_source_astrometry)
        prior = (xs, sg, ys, sg)
    jd = JointDeconvolution(d32, weight, psf, k, M, cv)
    if group is not None:
        jd.connect(group, comm)
    nu = jd.nu
    jd.set_params(h=np.zeros(nu * nu) if h0 is None else h0, mean=np.zeros(E), a=np.asarray(initial_a, np.float32).reshape(E, M),
                  c_x=xs, c_y=ys, dx=np.zeros(E), dy=np.zeros(E), alpha=alpha,
                  free_h=False, free_mean=False, free_a=True, free_c=False, free_d=True)
    # stage 1: translations and fluxes
    jd.set_reg(0.0, 0.0, 0.0, W=None, prior=prior, lam_fu=reg.get('regularization_scatter_fluxes_pre_optim', 10.0), conventions=cv)
    if stage1 not in ('device', 'scipy'):
        raise ValueError("stage1 must be 'device' or 'scipy'")
    if stage1 == 'device' and jd.world == 1:
        hist1, info1 = lbfgs_translations_and_fluxes_device(jd, roi_deconv_translations_iters)
    else:
        hist1, res1 = lbfgsb_translations_and_fluxes(jd, roi_deconv_translations_iters)
        info1 = dict(nit=int(res1.nit), nfev=int(res1.nfev), message=str(res1.message), fun=float(res1.fun))
    kwargs_partial1 = _kwargs_of(jd.get(want_model=False))
    # stage 2: everything
    free_h = bool(further_optimize_background)
    lam_s, lam_hf = reg.get('regularization_strength_scales', 1.0), reg.get('regularization_strength_hf', 1.0)
    lam_pts = reg.get('regularization_strength_pts_source', 0.01)
    jd.set_params(free_h=free_h, free_mean=True, free_a=True, free_c=not fix_c, free_d=True)
    jd.set_reg(lam_s, lam_hf, reg.get('regularization_strength_positivity', 100.0), W=None, prior=prior, lam_pts=lam_pts,
               lam_fu=reg.get('regularization_scatter_fluxes_main_optim', 10.0), conventions=cv)
    Wused = None
    if (free_h and (lam_s or lam_hf)) or lam_pts:
        Wused = jd.noise_weights()
    hist = jd.run(int(roi_deconv_all_iters), lr=1e-4, schedule=False)
    out = _finish(jd, hist, Wused, d32, weight, psf, cv)
    jd.close()
    out['stage1'] = dict(kwargs_final=kwargs_partial1, loss_history=hist1, **info1)
    return out


def get_fluxes_dataframe_from_model(result, data, noisemap, point_sources_names, model_scale, normalization_errors,
                                    frame_ids, mjds, seeings, zeropoint, sky_level_electron_per_second):
    """roi_modelling.py:420-497 up to the per-epoch table: fluxes per source ``a[i::M] * scale`` (:462), uncertainties
    sqrt(sigma_Fisher^2 + (normalisation error * flux)^2) (:465-467), reduced chi2 per frame = sum r^2/sigma^2 / n^2
    (:470-472).  Returns (per-epoch DataFrame indexed by frame_id, residuals); the grouping per night and the
    magnitude conversion stay lightcurver's (utilities/lightcurves_postprocessing.py)."""
    import pandas as pd
    kw = result['kwargs_final']
    M = len(point_sources_names)
    fluxes = np.asarray(kw['kwargs_analytic']['a'])
    sig = np.asarray(result['flux_sigma'])
    curves, d_curves = {}, {}
    for i, ps in enumerate(point_sources_names):
        # fluxes in pixel-sum units (what `a * scale` is in the reference, roi_modelling.py:462)
        curve = fluxes[i::M] * model_scale / result.get('amplitude_per_flux', 1.0)
        photon = sig[i::M] * model_scale / result.get('amplitude_per_flux', 1.0)
        curves[ps] = curve
        d_curves[ps] = (photon ** 2 + (np.asarray(normalization_errors) * curve) ** 2) ** 0.5
    residuals = np.asarray(data) - np.asarray(result['model'])
    with np.errstate(divide='ignore', invalid='ignore'):
        chi2_per_frame = np.nansum(residuals ** 2 / np.asarray(noisemap) ** 2, axis=(1, 2)) / data.shape[-1] ** 2
    rows = []
    for e in range(len(frame_ids)):
        row = {'frame_id': frame_ids[e], 'mjd': mjds[e], 'zeropoint': zeropoint, 'reduced_chi2': chi2_per_frame[e],
               'seeing': seeings[e], 'sky_level_electron_per_second': sky_level_electron_per_second[e]}
        for ps in point_sources_names:
            row[f'{ps}_flux'] = curves[ps][e]
            row[f'{ps}_d_flux'] = d_curves[ps][e]
        rows.append(row)
    return pd.DataFrame(rows).set_index('frame_id'), residuals
