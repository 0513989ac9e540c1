#!/usr/bin/env python
"""bench.py -- frames/s of the PSF + photometry fits on synthetic stamp stacks (BASELINE.json cfg2).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One step = one pass of the hot path over the cfg2 batch: 1,000 frames x 10 stars x 32x32 stamps,
subsampling 2: per-frame PSF fit (analytic Moffat stage <=100 its, noise weights, 3000 AdaBelief
iterations on the 64x64 pixel grid) followed by the fixed-PSF photometry of the same 10 stars
(2000 AdaBelief iterations) with the PSFs just fitted.  N > 1 (torchrun): every rank owns its own
1,000 frames (weak scaling, no data-path collective); value = all frames / max-over-ranks time.

JSON keys: see the task contract.  `value` is timed with inputs resident in HBM (device pointers
through the C ABI); `e2e` is the same step through the public Python API with pinned HOST buffers
(host preparation, H2D, kernels, D2H inside the timed region).  `roofline` refers to k_psf_fit (the
dominant kernel), timed live with CUDA events on its launch stream (lcb_profile_*); the bound is the
FP32 SIMT FMA pipe, peak measured live by lcb_fp32_peak.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

CFG = dict(F=1000, N=10, n=32, k=2, T1=100, T2=3000, Tphot=2000, G=12)
METRIC = "frames/sec PSF+photometry fit (cfg2: 1000 frames x 10 stars x 32x32, ss2, Moffat+grid)"


# ------------------------------------------------------------------ algorithmic work
def algorithmic_flops():
    F, N, n, k, G = CFG['F'], CFG['N'], CFG['n'], CFG['k'], CFG['G']
    nu = n * k
    J = int(np.log2(nu))
    per_it = 14 * G * nu * nu * N + N * (2 * nu * nu + 12 * n * n) + 2 * J * 21 * nu * nu + 16 * (nu * nu + 3 * N)
    psf = per_it * CFG['T2']
    phot_it = 10 * G * nu * nu + 3 * nu * nu + 18 * n * n
    phot = phot_it * CFG['Tphot'] * N
    # flops the kernels EXECUTE per PSF iteration and frame: the k-box is folded into GE = G + k - 1 decimating
    # taps, so one star costs 2 nu n GE (vertical, 2 kernels) + 3 n^2 GE (horizontal, 3 kernels) FMAs forward and n nu GE/k + nu^2 GE/k
    # for the two transposed passes; the starlet is 2 x 2 five-tap passes per scale (+ the point-wise steps)
    GE = G + k - 1
    fma_star = 2 * nu * n * GE + 3 * n * n * GE + (n * nu * GE) // k + (nu * nu * GE) // k
    fma_it = N * fma_star + J * (4 * 5 + 4) * nu * nu + 8 * (nu * nu + 3 * N)
    return dict(psf_per_frame=psf, phot_per_frame=phot, psf_per_it=per_it, phot_per_it_item=phot_it, psf_executed_per_it=2 * fma_it)


def algorithmic_bytes_per_frame():
    N, n, k = CFG['N'], CFG['n'], CFG['k']
    nu = n * k
    return 2 * N * n * n * 4 + (2 * nu * nu + N * n * n + 3 * N + 5 + CFG['T2']) * 4


# ------------------------------------------------------------------ clocks sampler
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,power.draw"

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            p = [x.strip() for x in r.split(',')]
            if len(p) < 6:
                continue
            try:
                sm.append(float(p[0])); mx.append(float(p[1]))
            except ValueError:
                continue
            for nm, v in zip(names, p[2:6]):
                if v.lower().startswith('active'):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------ CPU baseline (oracle port)
def cpu_baseline_run(sample_frames=8, t1=20, t2=200, tphot=200, threads=None, first_frame=0):
    """Times the oracle (restated STARRED model, PyTorch CPU float32) on a bounded sample of cfg2 -- frames
    first_frame .. first_frame + sample_frames - 1 of the workload -- and scales linearly in the iteration counts to
    (T1, T2, Tphot).  threads: torch intra-op threads (None = all host cores).  L-BFGS-B runs with scipy's default tolerances
    (what STARRED's Optimizer('l-bfgs-b') passes [R]); callers run one untimed warm call first."""
    import torch
    from oracle import starred_model as sm
    from lightcurver_b200 import synthetic
    cores = (os.cpu_count() or 1) if threads is None else int(threads)
    torch.set_num_threads(cores)
    n, k, N = CFG['n'], CFG['k'], CFG['N']
    d = synthetic.make_psf_frames(first_frame + sample_frames, N, n, k, seed=synthetic.SEEDS['cfg2'])
    d = {kk: (v[first_frame:] if isinstance(v, np.ndarray) and v.shape[:1] == (first_frame + sample_frames,) else v) for kk, v in d.items()}
    sc = d['data'].reshape(sample_frames, -1).max(1)[:, None, None, None] / 100.0
    data = d['data'] / sc
    nm = d['noisemap'] / sc
    weight = d['masks'] / nm ** 2
    a0 = (data * d['masks']).sum((-1, -2)) * sm.DEFAULT.amplitude_per_flux(k)
    nu = n * k
    t0 = time.perf_counter()
    st1 = [sm.fit_psf_stage1(data[f], weight[f], n, k, float(d['fwhm'][f]), a0[f], t1, strict_tol=False) for f in range(sample_frames)]
    t_stage1 = time.perf_counter() - t0
    s_fixed = np.stack([sm.moffat_image(r['fwhm_x'], r['fwhm_y'], r['phi'], r['beta'], n, k).numpy() for r in st1])
    a1 = np.stack([r['a'] for r in st1]); x1 = np.stack([r['x0'] for r in st1]); y1 = np.stack([r['y0'] for r in st1])
    t0 = time.perf_counter()
    W = np.stack([sm.psf_noise_weights(weight[f], a1[f], x1[f], y1[f], n, k).numpy() for f in range(sample_frames)])
    t_w = time.perf_counter() - t0
    t0 = time.perf_counter()
    r2 = sm.fit_psf_stage2(s_fixed, np.zeros((sample_frames, nu, nu)), a1, x1, y1, data, weight, W, n, k, t2,
                           lr=1e-5, dtype=torch.float32)
    t_stage2 = time.perf_counter() - t0
    s = s_fixed + r2['b']
    psf = (s / s.sum((-1, -2), keepdims=True)).astype(np.float32)
    t0 = time.perf_counter()
    sm.fit_phot(np.repeat(psf, N, 0), data.reshape(-1, n, n), (1.0 / nm ** 2).reshape(-1, n, n),
                a1.reshape(-1), n, k, tphot, dtype=torch.float32)
    t_phot = time.perf_counter() - t0
    n_lbfgs = max(1, int(np.mean([len(r['loss_hist']) for r in st1])))
    per_frame = (t_stage1 * CFG['T1'] / n_lbfgs + t_w + t_stage2 * CFG['T2'] / t2 + t_phot * CFG['Tphot'] / tphot) / sample_frames
    return dict(value=1.0 / per_frame, unit="frames/s", cores=cores, kind="port",
                sample=f"{sample_frames} frames x {N} stars x {n}x{n} (subsampling {k}) of the workload; oracle (restated STARRED model, PyTorch CPU f32 + autograd, "
                       f"scipy L-BFGS-B f64): stage1 {n_lbfgs} its {t_stage1:.2f}s, W {t_w:.2f}s, stage2 {t2} its {t_stage2:.2f}s, "
                       f"phot {tphot} its {t_phot:.2f}s; scaled linearly to {CFG['T1']}/{CFG['T2']}/{CFG['Tphot']} iterations",
                seconds=t_stage1 + t_w + t_stage2 + t_phot)


CPU_BATCH = 8       # frames per worker: the oracle batches its tensors over frames, which amortises PyTorch's per-operation overhead


def _cpu_worker(args):
    """CPU_BATCH frames on one core (spawned process, one torch thread): warm call, then the timed sample."""
    idx, t1, t2, tphot = args
    cpu_baseline_run(sample_frames=2, t1=2, t2=5, tphot=5, threads=1, first_frame=idx * CPU_BATCH)
    return cpu_baseline_run(sample_frames=CPU_BATCH, t1=t1, t2=t2, tphot=tphot, threads=1, first_frame=idx * CPU_BATCH)


def cpu_baseline_parallel(t1=20, t2=200, tphot=200, workers=None):
    """The CPU arm with ALL host cores, the way independent frames are processed on a CPU: one worker process per core, a batch of
    CPU_BATCH frames each, one torch thread per worker (measured on the GPU box: intra-op threading makes the small tensors of
    this path SLOWER -- the scipy L-BFGS-B stage of an 8-frame sample takes 14.5 s with 16 torch threads and 0.35 s with one --
    so frame-level parallelism is what uses the cores, and batching over frames inside a worker amortises PyTorch's per-operation
    overhead).  All workers run at the same time; the rate is the sum of their rates."""
    import multiprocessing as mp
    workers = workers or (os.cpu_count() or 1)
    ctx = mp.get_context('spawn')                          # the parent may hold a CUDA context: never fork it
    t0 = time.perf_counter()
    with ctx.Pool(workers) as pool:
        res = pool.map(_cpu_worker, [(i, t1, t2, tphot) for i in range(workers)])
    wall = time.perf_counter() - t0
    rate = float(sum(r['value'] for r in res))
    secs = [r['seconds'] for r in res]
    return dict(value=rate, unit="frames/s", cores=workers, kind="port",
                sample=f"{workers * CPU_BATCH} frames x {CFG['N']} stars x {CFG['n']}x{CFG['n']} (subsampling {CFG['k']}) of the workload, {CPU_BATCH} frames per worker "
                       f"process (one torch thread each, all {workers} running at once); oracle (restated STARRED model, PyTorch CPU f32 + autograd, scipy "
                       f"L-BFGS-B f64): stage1 <= {t1} its, stage2 {t2} its, phot {tphot} its, {np.mean(secs):.2f} s of CPU work per worker "
                       f"(min {np.min(secs):.2f}, max {np.max(secs):.2f}); scaled linearly to {CFG['T1']}/{CFG['T2']}/{CFG['Tphot']} iterations; "
                       f"value = sum of the workers' frames/s",
                seconds=wall, per_worker_frames_per_s=[float(r['value']) for r in res])


def workload_string(F, N, n, k, cfg_name='cfg2'):
    nu = n * k
    return (f"{cfg_name}: {F} frames x {N} stars x {n}x{n} per GPU, subsampling {k}: PSF fit (Moffat LM<= {CFG['T1']} its, "
            f"SLIT noise weights, {CFG['T2']} AdaBelief its on the {nu}x{nu} grid) + photometry of the same stars "
            f"({CFG['Tphot']} AdaBelief its)")


def run_reference(args):
    """Reference arm: the CPU restatement of the reference's path (oracle port; STARRED itself is not installable here) on the
    box's host cores, on OUR arm's config / metric / unit.  Every step is the SAME bounded sample as the `cpu_baseline` of our arm
    (`cpu_baseline_parallel`: one single-threaded worker process per host core, 8 frames of the workload batched per worker,
    <= 20 / 200 / 200 iterations of the three stages scaled linearly to 100 / 3000 / 2000); at most three timed steps so that the
    run ends within a few minutes."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    sample = dict(t1=20, t2=200, tphot=200)                # a few seconds of CPU work per core and step; the workers warm themselves
    vals, secs = [], []
    r = None
    for i in range(max(1, min(args.steps, 3))):
        r = cpu_baseline_parallel(**sample)
        vals.append(r['value']); secs.append(r['seconds'])
    v = float(np.median(vals))
    r['value'] = v
    r['spread'] = [float(np.min(vals)), float(np.max(vals))]
    line = {"metric": METRIC, "value": v, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(secs)), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference",
            "config": {"workload": workload_string(CFG['F'], CFG['N'], CFG['n'], CFG['k']), **CFG,
                       "l2": "n/a (CPU arm)", "seed": 20260102},
            "cpu_baseline": r,
            "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------ cfg4: joint deconvolution
DECONV = dict(E=200, n=64, k=2, M=4, npsf=32)
DECONV_METRIC = "joint deconvolution iterations/s (cfg4: 200 epochs x 64x64, ss2, 4 point sources, starlet reg)"


def deconv_flops(E, n, k, M, P):
    """(executed, survey) flops per iteration.  SURVEY 8d counts a full-resolution 'same' P x P convolution; the kernel folds the
    k x k decimation into the PSF: each of the k^2 polyphase planes (n x n) is correlated with an NA x NA
    kernel.  In-bounds MACs of that form, forward + adjoint, 2 flop each, + the point-source windows and the bilinear warp
    and its transpose -- the restated formula SURVEY asks for when the build changes the pass structure."""
    nu = n * k
    C = sum(min(nu, v + (P - 1) // 2 + 1) - max(0, v - P // 2) for v in range(nu)) ** 2
    survey = E * (4 * C + M * 14 * CFG['G'] * nu * nu + 16 * nu * nu)
    j0 = (P - 1) // 2
    A0 = -((P - 1 - j0 + k - 1) // k)
    NA = (j0 + k - 1) // k - A0 + 1
    rows = sum(min(n, Y + A0 + NA) - max(0, Y + A0) for Y in range(n))
    executed = E * (4 * k * k * rows * rows + M * 10 * CFG['G'] ** 2 + 24 * nu * nu)
    return executed, survey


def deconv_cpu_baseline(t, scale, n_iter=20):
    """Oracle (restated STARRED deconvolution, PyTorch CPU f32 + autograd, all host threads): ALL 200 epochs of cfg4, a bounded
    number of AdaBelief iterations (20 of the 2000) after one untimed iteration."""
    import torch
    from oracle import starred_model as sm
    E, n, k, M = DECONV['E'], DECONV['n'], DECONV['k'], DECONV['M']
    nu = n * k
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    rng = np.random.default_rng(7)
    data = rng.standard_normal((E, n, n)).astype(np.float32)
    w = np.ones((E, n, n), np.float32)
    params = dict(h=np.zeros(nu * nu), mean=np.zeros(E), a=t['a'] * scale, c_x=t['c_x'], c_y=t['c_y'], dx=np.zeros(E), dy=np.zeros(E))
    reg = dict(lam_scales=1.0, lam_hf=1.0, lam_pos=100.0, lam_pts=0.01, lam_fu=10.0)
    sm.fit_deconv(params, dict(alpha=np.zeros(E)), t['psf'], data, w, None, n, k, reg, 1, lr=1e-4, dtype=torch.float32)
    t0 = time.perf_counter()
    sm.fit_deconv(params, dict(alpha=np.zeros(E)), t['psf'], data, w, None, n, k, reg, n_iter, lr=1e-4, dtype=torch.float32)
    dt = time.perf_counter() - t0
    return dict(value=n_iter / dt, unit="it/s", cores=cores, kind="port", seconds=dt,
                sample=f"all {E} epochs x {n}x{n} (subsampling {k}, {M} sources, P={DECONV['npsf'] * k}), {n_iter} AdaBelief iterations after one "
                       f"untimed; oracle (restated STARRED deconvolution, PyTorch CPU f32 + autograd, FFT convolutions)")


def deconv_parity_vs_single_rank(world, rank, group, comm):
    """N >= 2: a small joint fit (9 epochs x 32x32, k = 2, M = 2, alpha on, every regulariser, 20 scheduled iterations) with its
    epochs sharded over the N ranks, against the same fit on rank 0 alone -- what tests/test_deconv_gpu.py::
    test_deconv_two_ranks_match_single_rank checks, run inside the bench because the GPU-test box has one GPU.
    Returns (on rank 0) dict(shared_bit_identical, max_abs_dh, median_abs_dh, max_rel_da, loss_rel)."""
    import torch
    import torch.distributed as dist
    from lightcurver_b200.processes.roi_modelling import JointDeconvolution, epoch_shard
    from lightcurver_b200 import synthetic
    E, n, k, M, npsf = max(9, 2 * world + 1), 32, 2, 2, 16
    nu = n * k
    t = synthetic.make_deconv_epochs(E, n, k, M=M, n_psf=npsf, seed=777)
    rng = np.random.default_rng(778)
    alpha = rng.uniform(-0.05, 0.05, E)
    scale = 1.0 / 3000.0

    # the stamps are rendered ONCE per rank over all E epochs (same handle shape on every rank -> identical bits), then sliced
    gen = JointDeconvolution(np.zeros((E, n, n), np.float32), np.ones((E, n, n), np.float32), t['psf'], k, M)
    gen.set_params(h=t['h'].reshape(-1), mean=np.zeros(E), a=t['a'], c_x=t['c_x'], c_y=t['c_y'], dx=t['dx'], dy=t['dy'], alpha=alpha)
    clean = gen.get()['model'].astype(np.float64)
    gen.close()
    sky = t['sky'][:, None, None]
    data_all = clean + np.sqrt(sky ** 2 + np.abs(clean)) * np.random.default_rng(900).standard_normal((E, n, n))
    sig_all = np.sqrt(sky ** 2 + np.abs(data_all))

    def fit(sl, grp):
        El = sl.stop - sl.start
        data, sig = data_all[sl], sig_all[sl]
        jd = JointDeconvolution((data * scale).astype(np.float32), (1.0 / (sig * scale) ** 2).astype(np.float32), t['psf'][sl], k, M)
        if grp is not None:
            jd.connect(grp, comm)
        jd.set_params(h=np.zeros(nu * nu), mean=np.zeros(El), a=(t['a'] * scale * 0.9)[sl], c_x=t['c_x'], c_y=t['c_y'],
                      dx=np.zeros(El), dy=np.zeros(El), alpha=alpha[sl])
        jd.set_reg(1.0, 1.0, 100.0, lam_pts=0.01, lam_fu=10.0)
        jd.noise_weights()
        hist = jd.run(20, lr=1e-4, schedule=True)
        fin = jd.get(want_model=False)
        jd.close()
        return hist, fin

    hist, fin = fit(epoch_shard(E, rank, world), group)
    shared = torch.from_numpy(np.concatenate([fin['h'], fin['c_x'], fin['c_y']])).cuda()
    gathered = [torch.empty_like(shared) for _ in range(world)]
    dist.all_gather(gathered, shared, group=group)
    a_parts = [None] * world
    dist.all_gather_object(a_parts, fin['a'], group=group)
    out = None
    if rank == 0:
        identical = all(bool(torch.equal(gathered[0], g)) for g in gathered[1:])
        hist1, fin1 = fit(slice(0, E), None)
        dh = np.abs(fin['h'] - fin1['h'])
        a_all = np.concatenate(a_parts)
        out = dict(shared_bit_identical=identical, max_abs_dh=float(dh.max()), median_abs_dh=float(np.median(dh)),
                   max_rel_da=float(np.max(np.abs(a_all - fin1['a']) / np.abs(fin1['a']))),
                   loss_rel=float(np.max(np.abs(hist - hist1) / np.abs(hist1))), epochs=E, iterations=20, ranks=world,
                   h_peak=float(np.abs(fin1['h']).max()), lr=1e-4,
                   # shared parameters bit-identical on every rank; against the single-rank fit: fluxes 1e-5, loss 1e-4, and the background
                   # within a few AdaBelief steps (a pixel whose gradient is at the rounding level moves by ~lr per iteration in a
                   # direction that depends on the summation order of the reduction, which differs between 1 and N ranks)
                   ok=bool(identical and dh.max() <= 10 * 1e-4 and np.median(dh) <= 0.5 * 1e-4 and
                           np.max(np.abs(a_all - fin1['a']) / np.abs(fin1['a'])) <= 1e-5 and
                           np.max(np.abs(hist - hist1) / np.abs(hist1)) <= 1e-4),
                   note="h, c_x, c_y compared bit for bit across ranks; against the single-rank fit: max |dh| <= 10 lr, median |dh| <= lr / 2, "
                        "fluxes 1e-5, loss history 1e-4 (the summation order of the gradient reduction differs between 1 and N ranks, and "
                        "AdaBelief moves a pixel whose gradient is at the rounding level by ~lr per iteration)")
    dist.barrier(group=group)
    return out


def deconv_section(world, rank, local, group, T, steps, warmup, comm='p2p', cpu_baseline=False):
    """BASELINE cfg4: 200 epochs x 64x64, subsampling 2 (128x128 background), 4 point sources, P = 64; stage 2
    of roi_modelling.py:326-335 (AdaBelief lr 1e-4, no schedule), epochs block-sharded over the ranks (strong
    scaling), nu^2 + 6M + 2 floats exchanged per iteration.  One step = T iterations.  Returns the result dict on rank 0."""
    import torch
    import torch.distributed as dist
    from lightcurver_b200 import _lib, synthetic
    from lightcurver_b200.processes.roi_modelling import JointDeconvolution, epoch_shard
    E, n, k, M, npsf = DECONV['E'], DECONV['n'], DECONV['k'], DECONV['M'], DECONV['npsf']
    nu, P = n * k, npsf * k
    sl = epoch_shard(E, rank, world)
    Eloc = sl.stop - sl.start
    t = synthetic.make_deconv_epochs(E, n, k, M=M, n_psf=npsf)
    rng = np.random.default_rng(synthetic.SEEDS['cfg4'] + 1)
    # render the truth with the library itself (data generator), add noise with the cutout_making.py:45 law
    gen = JointDeconvolution(np.zeros((Eloc, n, n), np.float32), np.ones((Eloc, n, n), np.float32), t['psf'][sl], k, M)
    gen.set_params(h=t['h'].reshape(-1), mean=np.zeros(Eloc), a=t['a'][sl], c_x=t['c_x'], c_y=t['c_y'], dx=t['dx'][sl],
                   dy=t['dy'][sl], alpha=np.zeros(Eloc))
    clean = gen.get()['model'].astype(np.float64)
    gen.close()
    sky = t['sky'][sl][:, None, None]
    noise = np.random.default_rng(1000 + rank).standard_normal(clean.shape)
    data = clean + np.sqrt(sky ** 2 + np.abs(clean)) * noise
    sig = np.sqrt(sky ** 2 + np.abs(data))
    scale = 1.0 / 3000.0
    jd = JointDeconvolution((data * scale).astype(np.float32), (1.0 / (sig * scale) ** 2).astype(np.float32), t['psf'][sl], k, M)
    if group is not None and world > 1:
        jd.connect(group, comm)
    jd.set_params(h=np.zeros(nu * nu), mean=np.zeros(Eloc), a=t['a'][sl] * scale * rng.uniform(0.9, 1.1, (E, M))[sl],
                  c_x=t['c_x'], c_y=t['c_y'], dx=np.zeros(Eloc), dy=np.zeros(Eloc), alpha=np.zeros(Eloc))
    jd.set_reg(1.0, 1.0, 100.0, lam_pts=0.01, lam_fu=10.0)        # roi_modelling.py:305-312 defaults
    jd.noise_weights()
    fp32_peak, _ = _lib.fp32_peak(8192)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        jd.run(T, lr=1e-4, schedule=False)
    barrier()
    # timed region: per-kernel profiling OFF, so that the iterations replay the captured CUDA graph (the product path)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    barrier()
    hist = None
    t0 = time.perf_counter()
    for i in range(steps):
        ev[i][0].record()
        hist = jd.run(T, lr=1e-4, schedule=False)
        ev[i][1].record()
    barrier()
    wall = time.perf_counter() - t0
    # per-kernel breakdown (CUDA events around every launch, eager launches): one extra untimed step
    _lib.profile_enable(True)
    jd.run(T, lr=1e-4, schedule=False)
    barrier()
    prof = _lib.profile_summary()
    _lib.profile_enable(False)
    ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
    tt = torch.tensor([ms], device='cuda')
    prof_ranks = None
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        prof_ranks = [None] * world                   # per-rank kernel times: shows which rank the others wait for in the update kernel
        dist.all_gather_object(prof_ranks, {kn: round(v['ms'] / max(v['launches'], 1), 4) for kn, v in prof.items()})
    ms = float(tt.item())
    ctas = int(_lib.lib.lcb_deconv_get_cluster(jd.handle))
    jd.close()
    if rank != 0:
        return None
    flop_it, flop_survey = deconv_flops(E, n, k, M, P)
    kep = prof.get('k_deconv_epoch', {'ms': 0.0, 'launches': 1})
    ach = (flop_it / world) / (kep['ms'] / max(kep['launches'], 1) * 1e-3) / 1e12 if kep['ms'] else 0.0
    res = {"metric": DECONV_METRIC, "value": T / (ms * 1e-3), "unit": "it/s", "it_s": T / (ms * 1e-3), "ms_per_iter": ms / T,
           "n_gpus": world, "steps": steps, "warmup": warmup, "iterations_per_step": T, "ms_per_step": ms, "scaling": "strong",
           "comm": ("none" if world == 1 else comm),
           "config": {"workload": f"cfg4 joint deconvolution, {T} AdaBelief iterations per step, {E} epochs sharded {world}-way, P={P}, M={M}",
                      "E": E, "n": n, "k": k, "M": M, "P": P,
                      "collective": ("none" if world == 1 else "in-kernel all-reduce of nu^2+6M+2 floats per iteration over NVLink peer memory "
                                     "(push + flag, summed in rank order)" if comm == 'p2p' else "1 NCCL all-reduce of nu^2+6M+2 floats per iteration"),
                      "ctas_per_epoch": ctas, "loss_first_last": [float(hist[0]), float(hist[-1])]},
           # kernels per iteration: starlet, epoch, update, + the reduce kernel unless the reduction runs fused in the epoch kernel's tail
           "gpu_launches": steps * ((4 if (comm == 'nccl' or 'k_deconv_reduce' in prof) else 3) * T + 1), "kernels": prof, "wall_s_timed_region": wall,
           "launch_mode": "iteration 0 eager, iterations 1.. replay one captured CUDA graph (starlet | epoch | reduce + NVLink push | update; "
                          "LCB_DECONV_REDUCE=fused folds the reduction and the push into the tail of the epoch kernel); "
                          "`kernels` comes from one extra step with per-launch CUDA events (eager launches)",
           "roofline": {"bound": "fp32", "kernel": "k_deconv_epoch", "achieved": ach, "peak": fp32_peak, "unit": "TFLOP/s",
                        "frac": ach / fp32_peak if fp32_peak else None, "traffic": None,
                        "algorithmic_flop_per_iteration": flop_it, "flop_per_iteration_survey_formula": flop_survey,
                        "note": "flops of the decimation-folded polyphase convolution the kernel executes (the restated SURVEY 8d formula: 4x fewer "
                                "MACs than the full-resolution count, same result); per-rank share over the mean k_deconv_epoch launch time"}}
    if prof_ranks is not None:
        res["kernel_ms_per_launch_by_rank"] = prof_ranks
    if cpu_baseline:
        res["cpu_baseline"] = deconv_cpu_baseline(t, scale)
    return res


def run_deconv(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    group = None
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
        group = dist.group.WORLD
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    res = deconv_section(world, rank, local, group, args.iters_per_step, args.steps, args.warmup, comm=args.comm,
                         cpu_baseline=(world == 1 and not args.no_cpu_baseline))
    clocks = sampler.stop() if rank == 0 else None
    par = deconv_parity_vs_single_rank(world, rank, group, args.comm) if world > 1 else None
    if rank == 0:
        line = {**res, "higher_is_better": True, "vs_baseline": None, "dtype": "f32", "data": "synthetic", "clocks": clocks}
        if par is not None:
            line["parity_vs_single_rank"] = par
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------ cfg3: zero-point photometry
def run_phot(args):
    """BASELINE cfg3: 10,000 frames x 20 stars x 32x32, subsampling 2, fixed (true) narrow PSF per frame, free a, dx, dy per
    (frame, star), 2000 scheduled AdaBelief iterations; weak scaling (every rank owns its own 10,000 frames, no collective).
    value: prepared stamps resident in HBM; e2e: star_photometry_batch from pinned host arrays (raw stamps, noise maps, PSFs)."""
    import torch
    import torch.distributed as dist
    from lightcurver_b200 import _lib, engine, synthetic
    from lightcurver_b200.processes.star_photometry import star_photometry_batch
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    F, S, n, k, T = (args.frames if args.frames != 1000 else 10000), 20, 32, 2, 2000
    d = synthetic.make_phot_frames(F, S, n, k, seed=synthetic.SEEDS['cfg3'] + rank)
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    h_data, h_nm, h_psf = pin(d['data']), pin(d['noisemap']), pin(d['psf'])
    prep = engine.phot_prepare_batch(h_data.numpy(), h_nm.numpy(), None, k)
    g_psf = h_psf.cuda()
    g_idx = torch.arange(F, dtype=torch.int32, device='cuda').repeat_interleave(S)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device='cuda')
    fp32_peak, _ = _lib.fp32_peak(8192)

    def step_device():
        return engine.phot_fit_batch(prep['data'], prep['weight'], g_psf, g_idx, prep['a0'], k, T, lr=1e-3, schedule=True,
                                     want_residuals=False, want_loss_hist=False)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_device(); flush.fill_(1)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    _lib.profile_enable(True)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for i in range(args.steps):
        ev[i][0].record(); out = step_device(); ev[i][1].record(); flush.fill_(i)
    barrier()
    prof = _lib.profile_summary()
    _lib.profile_enable(False)
    clocks = sampler.stop() if rank == 0 else None
    ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
    tt = torch.tensor([ms], device='cuda')
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms = float(tt.item())
    star_photometry_batch(h_data.numpy(), h_nm.numpy(), h_psf.numpy(), k, n_iter=T, want_loss_hist=False)
    barrier()
    t0 = time.perf_counter()
    ph = star_photometry_batch(h_data.numpy(), h_nm.numpy(), h_psf.numpy(), k, n_iter=T, want_loss_hist=False)
    barrier()
    te = torch.tensor([time.perf_counter() - t0], device='cuda')
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    if rank == 0:
        truth = d['transparency'][:, None] * d['star_flux'][None]
        rel = float(np.median(np.abs(ph['fluxes'] - truth) / truth))
        nu = n * k
        # SURVEY.md section 8d counts full-resolution separable passes (10 G nu^2 + 3 nu^2 + 18 n^2 = 522 kFLOP per item-iteration);
        # the kernel folds the k-box into G + k - 1 decimating taps and executes 2 (2 nu n + 3 n^2)(G + k - 1) = 186 kFLOP for the
        # same result, so SURVEY's count is not a bound (it gave frac > 1): the roofline uses the restated, executed count
        flop_survey = 10 * CFG['G'] * nu * nu + 3 * nu * nu + 18 * n * n
        flop_item_it = 2 * (2 * nu * n + 3 * n * n) * (CFG['G'] + k - 1)
        kp = prof.get('k_phot_fit', {'ms': 0.0, 'launches': 1})
        ach = flop_item_it * F * S * T / (kp['ms'] / max(kp['launches'], 1) * 1e-3) / 1e12 if kp['ms'] else 0.0
        line = {"metric": "frames/sec zero-point photometry (cfg3: 10,000 frames x 20 stars x 32x32, fixed PSF, amplitude+shift fit)",
                "value": world * F / (ms * 1e-3), "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"cfg3: {F} frames x {S} stars x {n}x{n} per GPU, subsampling {k}, {T} scheduled AdaBelief iterations "
                                       f"per (frame, star) item", "F": F, "S": S, "n": n, "k": k, "T": T,
                           "l2": "256 MiB buffer written between timed steps (L2 flush)", "seed": synthetic.SEEDS['cfg3'],
                           "quality": {"median_relative_flux_error": rel, "chi2_median": float(np.median(ph['chi2_per_frame']))}},
                "clocks": clocks,
                "e2e": {"value": world * F / float(te.item()), "unit": "frames/s",
                        "h2d_bytes_per_step": 2 * F * S * n * n * 4 + F * nu * nu * 4, "d2h_bytes_per_step": 6 * F * S * 4 + S * 4,
                        "api": "star_photometry_batch: pinned host numpy in (raw stamps, noise maps, PSFs), numpy out", "steps": 1},
                "gpu_launches": sum(v['launches'] for v in prof.values()), "kernels": prof,
                "roofline": {"bound": "fp32", "kernel": "k_phot_fit", "achieved": ach, "peak": fp32_peak, "unit": "TFLOP/s",
                             "frac": ach / fp32_peak if fp32_peak else None, "traffic": None,
                             "algorithmic_flop_per_launch": flop_item_it * F * S * T,
                             "flop_per_item_iteration": flop_item_it, "flop_per_item_iteration_survey_formula": flop_survey,
                             "note": "flops of the box-folded decimating passes the kernel executes (the restated SURVEY 8d formula: "
                                     "2 (2 nu n + 3 n^2)(G + k - 1) = 186 kFLOP per item-iteration; the full-resolution count, 522 kFLOP, is "
                                     "2.8x larger for the same result and is not a bound)"}}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------ our arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--frames', type=int, default=CFG['F'], help=argparse.SUPPRESS)
    ap.add_argument('--no-cpu-baseline', action='store_true', help=argparse.SUPPRESS)
    ap.add_argument('--workload', default='psfphot', choices=['psfphot', 'deconv', 'cfg5', 'cfg3'],
                    help='psfphot (default, BASELINE cfg2), deconv (cfg4: joint deconvolution iterations/s, epochs sharded over ranks) '
                         'cfg3 (zero-point photometry: 10,000 frames x 20 stars, fixed PSF) '
                         'or cfg5 (PSF + photometry at the large-survey shapes: 64x64 stamps, subsampling 3, 30 stars; one 148-frame wave per GPU)')
    ap.add_argument('--iters-per-step', type=int, default=200, help=argparse.SUPPRESS)
    ap.add_argument('--no-deconv', action='store_true', help=argparse.SUPPRESS)
    ap.add_argument('--comm', default='p2p', choices=['p2p', 'nccl'], help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference(args)
    if args.workload == 'deconv':
        return run_deconv(args)
    if args.workload == 'cfg3':
        return run_phot(args)
    global METRIC
    if args.workload == 'cfg5':
        # BASELINE cfg5 is 20,000 frames; throughput is linear in the number of 148-frame waves, so one wave per GPU is timed
        CFG.update(F=148, N=30, n=64, k=3)
        METRIC = "frames/sec PSF+photometry fit (cfg5 shapes: 30 stars x 64x64, ss3, Moffat+grid; one 148-frame wave per GPU)"
        if args.frames == 1000:
            args.frames = 148

    import torch
    import torch.distributed as dist
    from lightcurver_b200 import _lib, engine, synthetic
    from lightcurver_b200.procedures.psf_routines import build_psf_batch
    from lightcurver_b200.processes.star_photometry import star_photometry_batch

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    dev = torch.device('cuda', local)
    F, N, n, k = args.frames, CFG['N'], CFG['n'], CFG['k']
    nu = n * k

    # ---- synthetic workload (seed differs per rank: every rank owns its own frames)
    cfg_name = 'cfg5' if args.workload == 'cfg5' else 'cfg2'
    d = synthetic.make_psf_frames(F, N, n, k, seed=synthetic.SEEDS[cfg_name] + rank)
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    h_data, h_nm, h_mask = pin(d['data']), pin(d['noisemap']), pin(d['masks'])
    fwhm_guess = d['fwhm'] * np.random.default_rng(1).uniform(0.9, 1.1, F)

    # device-resident, already-prepared inputs for the `value` arm
    sc = torch.from_numpy(d['data']).reshape(F, -1).max(1).values[:, None, None, None] / 100.0
    g_data = (torch.from_numpy(d['data']) / sc).reshape(F * N, n, n).to(dev)
    g_nm = (torch.from_numpy(d['noisemap']) / sc).reshape(F * N, n, n).to(dev)
    g_w = (torch.from_numpy(d['masks']).reshape(F * N, n, n).to(dev) / g_nm ** 2).contiguous()
    g_wphot = (1.0 / g_nm ** 2).contiguous()
    g_off = (torch.arange(F + 1, dtype=torch.int32) * N).to(dev)
    from lightcurver_b200.conventions import DEFAULT as CV
    g_a0 = ((g_data * (g_w > 0)).sum((-1, -2)) * CV.amplitude_per_flux(k)).contiguous()
    g_mof = torch.tensor(np.stack([fwhm_guess, fwhm_guess, np.zeros(F), np.full(F, 2.5), np.ones(F)], -1), dtype=torch.float32).to(dev)
    g_idx = torch.arange(F, dtype=torch.int32).repeat_interleave(N).to(dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    def step_device():
        out = engine.psf_fit_batch(g_data, g_w, g_off, k, g_mof, g_a0, n_iter_analytic=CFG['T1'],
                                   n_iter_adabelief=CFG['T2'], lr=1e-5, lam_scales=1.0, lam_hf=1.0, noise_weights=True,
                                   want=('narrow_psf', 'full_psf', 'residuals', 'chi2', 'loss_hist', 'status'))
        ph = engine.phot_fit_batch(g_data, g_wphot, out['narrow_psf'], g_idx, out['a'], k, CFG['Tphot'], lr=1e-3,
                                   schedule=True, want_residuals=False, want_loss_hist=True)
        return out, ph

    def step_e2e():
        res = build_psf_batch(h_data.numpy(), h_nm.numpy(), k, masks=h_mask.numpy(), n_iter_analytic=CFG['T1'],
                              n_iter_adabelief=CFG['T2'], guess_method_star_position='center',
                              guess_fwhm_pixels=fwhm_guess, return_dicts=False)
        ph = star_photometry_batch(h_data.numpy(), h_nm.numpy(), res['narrow_psf'], k, n_iter=CFG['Tphot'],
                                   masks=h_mask.numpy(), want_loss_hist=False)
        return res, ph

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- FP32 peak (roofline denominator), before the run heats the chip
    fp32_peak, _ = _lib.fp32_peak(8192)

    for _ in range(args.warmup):
        step_device()
        flush.fill_(1)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    _lib.profile_enable(True)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    t_wall0 = time.perf_counter()
    for i in range(args.steps):
        ev[i][0].record()
        out, ph = step_device()
        ev[i][1].record()
        flush.fill_(i)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    prof = _lib.profile_summary()
    _lib.profile_enable(False)
    clocks = sampler.stop() if rank == 0 else None
    ms_steps = [a.elapsed_time(b) for a, b in ev]
    ms_step = float(np.mean(ms_steps))
    t = torch.tensor([ms_step], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step_max = float(t.item())
    value = world * F / (ms_step_max * 1e-3)
    chi2_med = float(out['chi2'].median())
    phot_chi2_med = float(ph['chi2'].median())

    # ---- e2e through the public API with pinned host buffers
    e2e_steps = 5 if args.steps >= 3 else max(1, args.steps)      # at least five end-to-end steps in the default run
    step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        res, ph2 = step_e2e()
    barrier()
    t_e2e = (time.perf_counter() - t0) / e2e_steps
    te = torch.tensor([t_e2e], device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * F / float(te.item())
    # bytes moved by the public API per step (counted from the tensors it copies): raw stamps, noise maps (f32) and masks
    # (u8) for the PSF step and again for the photometry step, the fitted PSFs back up for the photometry; down: every
    # product of build_psf_batch (parameters, grids, PSFs, residuals, loss histories) and of star_photometry_batch
    h2d = 2 * (2 * F * N * n * n * 4 + F * N * n * n) + (F + 1) * 4 + F * 5 * 4 + F * nu * nu * 4
    d2h = (3 * F * nu * nu + F * N * n * n + F * (CFG['T1'] + CFG['T2']) + 3 * F * N + 5 * F + 3 * F) * 4 \
        + 6 * F * N * 4 + N * 4

    # ---- cfg4 joint deconvolution in the same run (strong-scaled over the ranks; the one path with an exchange step)
    deconv, deconv_par = None, None
    if args.workload == 'psfphot' and not args.no_deconv:
        grp = dist.group.WORLD if world > 1 else None
        deconv = deconv_section(world, rank, local, grp, args.iters_per_step, 3, 1, comm=args.comm,
                                cpu_baseline=(world == 1 and not args.no_cpu_baseline))
        if world > 1:
            deconv_par = deconv_parity_vs_single_rank(world, rank, grp, args.comm)
        if rank == 0 and deconv is not None and deconv_par is not None:
            deconv["parity_vs_single_rank"] = deconv_par

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    fl = algorithmic_flops()
    kfit = prof.get('k_psf_fit', {'ms': 0.0, 'launches': 0})
    launches = sum(v['launches'] for v in prof.values())
    fit_launches = max(kfit['launches'], 1)                 # more than one when the batch exceeds the 1184-frame workspace (chunks)
    fit_ms_per_launch = kfit['ms'] / fit_launches
    F_launch = F / fit_launches                             # frames one launch processes on average
    achieved = fl['psf_per_frame'] * F_launch / (fit_ms_per_launch * 1e-3) / 1e12 if fit_ms_per_launch > 0 else 0.0
    hbm_peak = None
    try:
        hbm_peak = json.load(open(ROOT / 'MEASURED_PEAKS.json'))['hbm_gbs']
    except Exception:
        pass
    traffic = None
    try:
        if args.workload == 'psfphot':
            tj = json.load(open(ROOT / 'profiles' / 'k_psf_fit_traffic.json'))      # from the committed ncu --set full capture
            traffic = tj['dram_bytes_per_launch'] * F_launch / tj['frames']       # per launch
    except Exception:
        pass
    hbm_ach = algorithmic_bytes_per_frame() * F_launch / (fit_ms_per_launch * 1e-3) / 1e9 if fit_ms_per_launch > 0 else 0.0
    line = {
        "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step_max, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_string(F, N, n, k, cfg_name), **{kk: (F if kk == 'F' else v) for kk, v in CFG.items()},
                   "l2": "256 MiB buffer written between timed steps (L2 flush)", "seed": synthetic.SEEDS[cfg_name],
                   "quality": {"psf_chi2_median": chi2_med, "phot_chi2_median": phot_chi2_med}},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "api": "build_psf_batch + star_photometry_batch: pinned host numpy in (raw stamps, noise maps, masks), numpy out; data policies, fits and products on the device", "steps": e2e_steps},
        "gpu_launches": launches,
        "kernels": prof,
        "roofline": {"bound": "fp32", "kernel": "k_psf_fit", "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s",
                     "frac": achieved / fp32_peak if fp32_peak else None, "traffic": traffic,
                     "peak_source": "lcb_fp32_peak measured live (FFMA chains on all SMs); MEASURED_PEAKS.json has no FP32 SIMT figure",
                     "algorithmic_flop_per_launch": fl['psf_per_frame'] * F_launch, "ms_per_launch": fit_ms_per_launch,
                     "launches_per_step": fit_launches,
                     "frac_executed": (achieved / fp32_peak if fp32_peak else 0.0) * fl['psf_executed_per_it'] / fl['psf_per_it'],
                     "note": "frac uses SURVEY 8d's algorithmic count (full-resolution separable passes); frac_executed is the FP32 pipe "
                             "utilisation on the flops the kernel executes after folding the k-box into the taps",
                     "hbm": {"achieved": hbm_ach, "peak": hbm_peak, "unit": "GB/s",
                             "frac": (hbm_ach / hbm_peak) if hbm_peak else None, "peak_source": "MEASURED_PEAKS.json (of measured)"}},
        "wall_s_timed_region": t_wall,
    }
    if deconv is not None:
        line["deconv"] = deconv
    if not args.no_cpu_baseline and world == 1:
        if args.workload == 'cfg5':
            cpu_baseline_run(sample_frames=1, t1=2, t2=2, tphot=2)          # untimed warm call
            line["cpu_baseline"] = cpu_baseline_run(sample_frames=1, t1=4, t2=20, tphot=20)
        else:
            line["cpu_baseline"] = cpu_baseline_parallel(t1=20, t2=200, tphot=200)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
