"""N > 1 host logic on CPU (gloo, world_size 2): epoch sharding of the joint deconvolution.

The data-path collective of the joint fit is ONE sum all-reduce of [dL/dh, dL/dc_x, dL/dc_y, loss,
|g_epoch|^2] per iteration.  Here the local contribution of each rank is produced by the CPU oracle
on its epoch shard; the all-reduced buffer must equal the unsharded gradient, which is what makes the
replicated update of the shared parameters identical on every rank."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    s = socket.socket(); s.bind(('127.0.0.1', 0)); p = s.getsockname()[1]; s.close(); return p


def _problem():
    sys.path.insert(0, str(ROOT))
    from oracle import starred_model as sm
    rng = np.random.default_rng(3)
    E, n, k, M = 5, 12, 2, 2
    nu = n * k
    fw = rng.uniform(2.5, 3.5, E)
    psf = sm.moffat_image(torch.tensor(fw), torch.tensor(fw), torch.zeros(E, dtype=torch.float64), torch.full((E,), 3.0, dtype=torch.float64), 8, k).numpy()
    prm = dict(h=0.1 * rng.standard_normal(nu * nu), mean=0.01 * rng.standard_normal(E),
               a=rng.uniform(1, 2, (E, M)) * sm.DEFAULT.amplitude_per_flux(k) / (k * k),      # pixel-sum fluxes of 0.25 .. 0.5
              
               c_x=rng.uniform(-2, 2, M), c_y=rng.uniform(-2, 2, M), dx=rng.uniform(-1, 1, E), dy=rng.uniform(-1, 1, E))
    data = rng.standard_normal((E, n, n)); weight = rng.uniform(0.5, 2, (E, n, n))
    return E, n, k, M, psf, prm, data, weight


def _worker(rank, world, port, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'; os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    sys.path.insert(0, str(ROOT))
    from oracle import starred_model as sm
    from lightcurver_b200.processes.roi_modelling import epoch_shard
    E, n, k, M, psf, prm, data, weight = _problem()
    sl = epoch_shard(E, rank, world)
    local = {kk: (v[sl] if kk in ('mean', 'a', 'dx', 'dy') else v) for kk, v in prm.items()}
    L, g = sm.deconv_loss_grad(local, dict(alpha=np.zeros(sl.stop - sl.start)), psf[sl], data[sl], weight[sl], None, n, k, {})
    buf = torch.tensor(np.concatenate([g['h'], g['c_x'], g['c_y'], [L]]))
    dist.all_reduce(buf)
    q.put((rank, buf.numpy(), sl.start, sl.stop, g['a']))
    dist.destroy_process_group()


def test_epoch_shard_partition():
    sys.path.insert(0, str(ROOT))
    from lightcurver_b200.processes.roi_modelling import epoch_shard
    for E in (1, 5, 200, 201):
        for world in (1, 2, 3, 8):
            sl = [epoch_shard(E, r, world) for r in range(world)]
            assert sl[0].start == 0 and sl[-1].stop == E
            assert all(a.stop == b.start for a, b in zip(sl, sl[1:]))
            sizes = [s.stop - s.start for s in sl]
            assert max(sizes) - min(sizes) <= 1


def test_sharded_gradient_allreduce_equals_unsharded():
    world, port = 2, _free_port()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    sys.path.insert(0, str(ROOT))
    from oracle import starred_model as sm
    E, n, k, M, psf, prm, data, weight = _problem()
    L, g = sm.deconv_loss_grad(prm, dict(alpha=np.zeros(E)), psf, data, weight, None, n, k, {})
    full = np.concatenate([g['h'], g['c_x'], g['c_y'], [L]])
    for rank, buf, lo, hi, ga in res:
        np.testing.assert_allclose(buf, full, rtol=1e-10, atol=1e-12)          # every rank holds the global sums
        np.testing.assert_allclose(ga, g['a'][lo:hi], rtol=1e-10, atol=1e-12)  # per-epoch gradients stay local


class _OracleShard:
    """Stands in for JointDeconvolution on CPU: same attributes and methods as the host-side stage-1 driver uses
    (E, M, n, world, rank, group, get, set_params, loss_grad), with the oracle evaluating the LOCAL epochs and the loss
    (and the flux-uniformity sums) all-reduced, like lcb_deconv_loss_grad with a connected communicator."""

    def __init__(self, sl, rank, world, group, lam_fu):
        from oracle import starred_model as sm
        self.sm = sm
        E, n, k, M, psf, prm, data, weight = _problem()
        self.n, self.k, self.M, self.E = n, k, M, sl.stop - sl.start
        self.E_total, self.sl = E, sl
        self.rank, self.world, self.group = rank, world, group
        self.psf, self.data, self.weight = psf[sl], data[sl], weight[sl]
        self.fixed = dict(h=np.zeros((n * k) ** 2), mean=np.zeros(self.E), c_x=prm['c_x'], c_y=prm['c_y'], alpha=np.zeros(self.E))
        self.p = dict(dx=np.zeros(self.E), dy=np.zeros(self.E), a=prm['a'][sl].copy())
        self.lam_fu = lam_fu

    def get(self, want_model=False):
        return dict(dx=self.p['dx'].copy(), dy=self.p['dy'].copy(), a=self.p['a'].reshape(-1).copy())

    def set_params(self, dx=None, dy=None, a=None, **_):
        self.p = dict(dx=np.asarray(dx, float), dy=np.asarray(dy, float), a=np.asarray(a, float).reshape(self.E, self.M))

    def loss_grad(self):
        L, g = self.sm.deconv_loss_grad(self.p, self.fixed, self.psf, self.data, self.weight, None, self.n, self.k, {})
        a = torch.tensor(self.p['a'])
        stats = torch.cat([a.sum(0), (a * a).sum(0), torch.tensor([L])])
        if self.world > 1:
            dist.all_reduce(stats, group=self.group)
        M = self.M
        mean = stats[:M] / self.E_total
        var = stats[M:2 * M] / self.E_total - mean ** 2
        sd = var.clamp_min(0).sqrt()
        Lfu = float(self.lam_fu * (sd / mean.abs()).sum())
        A = self.lam_fu / (self.E_total * sd * mean.abs())
        B = self.lam_fu * sd * torch.sign(mean) / (self.E_total * mean ** 2)
        ga = g['a'] + (A[None] * (a - mean[None]) - B[None]).numpy()
        return dict(loss=float(stats[-1]) + Lfu, dx=g['dx'], dy=g['dy'], a=ga.reshape(-1))


def _lbfgs_worker(rank, world, port, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'; os.environ['MASTER_PORT'] = str(port)
    if world > 1:
        dist.init_process_group('gloo', rank=rank, world_size=world)
    sys.path.insert(0, str(ROOT))
    from lightcurver_b200.processes.roi_modelling import epoch_shard, lbfgsb_translations_and_fluxes
    E = _problem()[0]
    jd = _OracleShard(epoch_shard(E, rank, world), rank, world, dist.group.WORLD if world > 1 else None, lam_fu=5.0)
    hist, res = lbfgsb_translations_and_fluxes(jd, 25)
    q.put((rank, jd.sl.start, jd.sl.stop, jd.get(), float(res.fun), int(res.nit)))
    if world > 1:
        dist.destroy_process_group()


def test_sharded_lbfgsb_stage1_equals_unsharded():
    """Stage 1 of the ROI modelling (roi_modelling.py:260-281) with epochs sharded over 2 ranks: every rank runs the same
    scipy L-BFGS-B on the all-gathered vector; the result equals the single-rank run (host logic, oracle as the evaluator)."""
    ctx = mp.get_context('spawn')
    out = {}
    for world in (1, 2):
        port = _free_port()
        q = ctx.Queue()
        procs = [ctx.Process(target=_lbfgs_worker, args=(r, world, port, q)) for r in range(world)]
        for p in procs:
            p.start()
        out[world] = sorted([q.get(timeout=300) for _ in range(world)], key=lambda t: t[0])
        for p in procs:
            p.join(timeout=60)
            assert p.exitcode == 0
    ref = out[1][0]
    assert ref[5] >= 3
    for rank, lo, hi, prm, fun, nit in out[2]:
        # (the sharded loss is a sum of partial sums: rounding-level differences that 25 quasi-Newton steps amplify)
        assert abs(fun - ref[4]) <= 1e-5 * abs(ref[4])
        np.testing.assert_allclose(prm['dx'], ref[3]['dx'][lo:hi], atol=2e-3)
        M = 2
        np.testing.assert_allclose(prm['a'], ref[3]['a'][lo * M:hi * M], rtol=2e-3)


def test_split_by_work_properties():
    """In-process multi-GPU fan-out: contiguous blocks in order, every item exactly once, no empty
    block, and no block heavier than the ideal share plus one item."""
    sys.path.insert(0, str(ROOT))
    from lightcurver_b200.engine import split_by_work
    rng = np.random.default_rng(0)
    for trial in range(200):
        nitem = int(rng.integers(1, 40))
        parts = int(rng.integers(1, 9))
        work = rng.integers(1, 31, nitem)
        blocks = split_by_work(work, parts)
        assert 1 <= len(blocks) <= min(parts, nitem)
        assert blocks[0][0] == 0 and blocks[-1][1] == nitem
        assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:])) and all(hi > lo for lo, hi in blocks)
        heaviest = max(work[lo:hi].sum() for lo, hi in blocks)
        assert heaviest <= work.sum() / parts + work.max()
