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
    prm = dict(h=0.1 * rng.standard_normal(nu * nu), mean=0.01 * rng.standard_normal(E), a=rng.uniform(1, 2, (E, M)),
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
