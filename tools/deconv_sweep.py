"""Timing sweep of the joint-deconvolution iteration (cfg4 shapes): local epochs x CTAs per epoch.

    python tools/deconv_sweep.py [--iters 100]

Prints, per (E_local, cluster size), ms per iteration twice -- with CUDA-graph replays (what lcb_deconv_run does) and with eager
launches under per-kernel CUDA events (lcb_profile_*), whose per-kernel split follows.  E_local = 200/100/50/25 are the shards of
cfg4 on 1/2/4/8 GPUs.  With a -DLCB_DC_TIMERS build (python -m lightcurver_b200.build --variant dctim -DLCB_DC_TIMERS;
LCB_LIBRARY=lightcurver_b200/liblcb_dctim.so) the clock64 stamps of the phases of k_deconv_epoch are printed as median / max cycles
over the CTAs of the last launch.  LCB_DECONV_REDUCE=fused|separate and LCB_DECONV_CS=n select the reduction route / cluster size."""
import argparse
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--iters', type=int, default=100)
    ap.add_argument('--epochs', default='200,100,50,25')
    ap.add_argument('--cs', default='0,1,2,4,8')
    ap.add_argument('--alpha', type=float, default=0.0)
    args = ap.parse_args()
    import torch
    from lightcurver_b200 import _lib, synthetic
    from lightcurver_b200.processes.roi_modelling import JointDeconvolution
    n, k, M, npsf = 64, 2, 4, 32
    nu = n * k
    t = synthetic.make_deconv_epochs(200, n, k, M=M, n_psf=npsf)
    rng = np.random.default_rng(0)
    for E in [int(x) for x in args.epochs.split(',')]:
        data = rng.standard_normal((E, n, n)).astype(np.float32)
        weight = np.ones((E, n, n), np.float32)
        for cs in [int(x) for x in args.cs.split(',')]:
            jd = JointDeconvolution(data, weight, t['psf'][:E], k, M)
            jd.set_cluster(cs)
            jd.set_params(h=np.zeros(nu * nu), mean=np.zeros(E), a=t['a'][:E] / 3000.0, c_x=t['c_x'], c_y=t['c_y'],
                          dx=t['dx'][:E], dy=t['dy'][:E], alpha=np.full(E, args.alpha))
            jd.set_reg(1.0, 1.0, 100.0, lam_pts=0.01, lam_fu=10.0)
            jd.run(10, lr=1e-4)
            torch.cuda.synchronize()
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record()
            jd.run(args.iters, lr=1e-4)          # CUDA-graph replays (no per-kernel events)
            g1.record()
            torch.cuda.synchronize()
            ms_graph = g0.elapsed_time(g1) / args.iters
            _lib.profile_enable(True)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            jd.run(args.iters, lr=1e-4)
            e1.record()
            torch.cuda.synchronize()
            prof = _lib.profile_summary()
            _lib.profile_enable(False)
            ms = e0.elapsed_time(e1) / args.iters
            used = int(_lib.lib.lcb_deconv_get_cluster(jd.handle))
            parts = ' '.join(f"{kn.replace('k_deconv_', '')}={v['ms'] / max(v['launches'], 1):.3f}" for kn, v in sorted(prof.items()))
            print(f"E={E:4d} cs={cs} (used {used}) {ms:.3f} ms/it  {1e3 / ms:7.1f} it/s eager | graph {ms_graph:.3f} ms/it {1e3 / ms_graph:7.1f} it/s   [{parts}]", flush=True)
            if hasattr(_lib.lib, 'lcb_debug_dc_timers'):     # -DLCB_DC_TIMERS build: phase stamps of the LAST launch of k_deconv_epoch
                import ctypes
                nc = E * used
                buf = np.zeros((nc, 16), np.int64)
                _lib.lib.lcb_debug_dc_timers.argtypes = [ctypes.c_void_p, ctypes.c_int]
                _lib.lib.lcb_debug_dc_timers(buf.ctypes.data, nc)
                names = ['zero', 'fbuild', 'cl0', 'halo', 'tma', 'fwd', 'cl1', 'adj', 'grads', 'ptsreg', 'cl2', 'warpT', 'cl3', 'tail']
                d = np.diff(np.concatenate([np.zeros((nc, 1), np.int64), buf[:, :13]], axis=1), axis=1)
                med, mx = np.median(d, axis=0), d.max(axis=0)
                print('      cycles median/max: ' + ' '.join(f"{nm}={int(a)}/{int(b)}" for nm, a, b in zip(names, med, mx)))
                tail = buf[:, 13][buf[:, 13] > 0]
                span = (buf[:, 15].max() - buf[:, 14].min()) / 1e3
                start_skew = (buf[:, 14].max() - buf[:, 14].min()) / 1e3
                print(f"      end-of-body total median {int(np.median(buf[:, 12]))} max {int(buf[:, 12].max())} cycles; tail CTAs {tail.size}: "
                      f"{[int(x) for x in np.sort(tail)[-4:]]}; grid span {span:.1f} us, start skew {start_skew:.1f} us", flush=True)
            jd.close()


if __name__ == '__main__':
    main()
