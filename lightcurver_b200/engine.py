"""Batched entry points over the C ABI (numpy host arrays or torch CUDA tensors).

Host arrays go through the library's LCB_MEM_HOST path (H2D + kernels + D2H inside the call);
torch CUDA tensors are passed by device pointer and the work is enqueued on torch's current
stream.  torch is used for device memory and streams only.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import as_f32, as_i32, ptr, empty_like_kind, mem_kind, current_stream


def phot_fit_batch(data, weight, psf, psf_index, a0, k, n_iter, lr=1e-3, schedule=True,
                   dx0=None, dy0=None, want_residuals=True, want_loss_hist=True, want_grad0=False):
    """K2: B independent fixed-PSF amplitude+shift fits (lcb_phot_fit_batch).

    data, weight (B,n,n); psf (Fp,n*k,n*k); psf_index (B,) int; a0 (B,).
    Returns dict(a, dx, dy, sigma_a, chi2, status [, residuals, loss_hist, loss0, grad0]).
    """
    _lib.require_device()
    data, weight, psf, a0 = as_f32(data), as_f32(weight), as_f32(psf), as_f32(a0)
    dx0, dy0 = as_f32(dx0), as_f32(dy0)
    psf_index = as_i32(psf_index)
    mem = mem_kind(data, weight, psf, psf_index, a0, dx0, dy0)
    B, n = int(data.shape[0]), int(data.shape[-1])
    if psf.ndim == 2:
        psf = psf[None]
    if tuple(psf.shape[-2:]) != (n * k, n * k):
        raise ValueError(f"psf must be (Fp,{n * k},{n * k}) for stamps of side {n} and subsampling {k}; got {tuple(psf.shape)}")
    if tuple(data.shape) != (B, n, n) or tuple(weight.shape) != (B, n, n):
        raise ValueError("data and weight must both be (B,n,n)")
    out = {nm: empty_like_kind(data, (B,)) for nm in ('a', 'dx', 'dy', 'sigma_a', 'chi2')}
    out['status'] = empty_like_kind(data, (B,), 'i')
    if want_residuals:
        out['residuals'] = empty_like_kind(data, (B, n, n))
    if want_loss_hist:
        out['loss_hist'] = empty_like_kind(data, (B, max(n_iter, 0)))
    if want_grad0:
        out['loss0'] = empty_like_kind(data, (B,))
        out['grad0'] = empty_like_kind(data, (B, 3))
    bi = _lib.PhotBatch(B, n, k, ptr(data), ptr(weight), ptr(psf), ptr(psf_index), int(psf.shape[0]),
                        ptr(a0), ptr(dx0), ptr(dy0))
    bo = _lib.PhotOut(*[ptr(out.get(nm)) for nm in ('a', 'dx', 'dy', 'sigma_a', 'chi2', 'residuals',
                                                    'loss_hist', 'loss0', 'grad0', 'status')])
    opts = _lib.FitOpts(int(n_iter), float(lr), int(bool(schedule)))
    rc = _lib.lib.lcb_phot_fit_batch(C.byref(bi), C.byref(opts), C.byref(bo), mem, current_stream(data))
    _lib.check(rc, 'lcb_phot_fit_batch')
    return out


def starlet_scales(nu):
    return int(_lib.lib.lcb_starlet_scales(int(nu)))


def _clone_f32(x, ref):
    """float32 C-contiguous private copy of the same kind as ``ref`` (in/out arrays of the ABI)."""
    if _lib.is_torch(ref):
        import torch
        return torch.as_tensor(x, dtype=torch.float32, device=ref.device).contiguous().clone()
    return np.array(x, dtype=np.float32, order='C', copy=True)


def psf_fit_batch(data, weight, star_off, k, moffat0, a0, x00=None, y00=None, background0=None,
                  W=None, n_iter_analytic=100, n_iter_adabelief=3000, lr=1e-3,
                  lam_scales=1.0, lam_hf=1.0, noise_weights=False, bounds=None, mc_samples=100, mc_seed=1,
                  want=('narrow_psf', 'full_psf', 'residuals', 'chi2', 'loss_hist', 'status'),
                  field_distortion=0, stamp_xy=None, distortion0=None):
    """K1: ragged batch of per-frame PSF fits (lcb_psf_fit_batch).

    data, weight (sumN,n,n); star_off (F+1,) CSR offsets; moffat0 (F,5) = fwhm_x, fwhm_y, phi, beta, C
    guesses; a0 (sumN,).  ``want`` lists optional outputs (see lcb_psf_out in include/lcb.h).
    noise_weights: False (W from the caller, None == 1), True / 'SLIT' (diagonal propagation), 'MC' (mc_samples draws).
    field_distortion: 0 off, 1 flux-conserving / 2 plain affine resampling of the narrow PSF per star (include/lcb.h);
    stamp_xy (sumN,2) rescaled frame coordinates of the stamps, distortion0 (F,6) initial coefficients (default 0).
    Returns a dict with moffat, a, x0, y0, background (+ distortion) and the requested outputs.
    """
    _lib.require_device()
    data, weight = as_f32(data), as_f32(weight)
    W = as_f32(W)
    star_off = as_i32(star_off)
    mem = mem_kind(data, weight, star_off, W)
    sumN, n = int(data.shape[0]), int(data.shape[-1])
    F = int(star_off.shape[0]) - 1
    nu = n * k
    J = starlet_scales(nu)
    if W is not None and tuple(W.shape) != (F, J, nu, nu):
        raise ValueError(f"W must be ({F},{J},{nu},{nu}); got {tuple(W.shape)}")
    out = dict(moffat=_clone_f32(moffat0, data), a=_clone_f32(a0, data),
               x0=_clone_f32(np.zeros(sumN) if x00 is None else x00, data),
               y0=_clone_f32(np.zeros(sumN) if y00 is None else y00, data),
               background=_clone_f32(np.zeros((F, nu, nu)) if background0 is None else background0, data))
    if tuple(out['moffat'].shape) != (F, 5) or tuple(out['a'].shape) != (sumN,):
        raise ValueError("moffat0 must be (F,5) and a0 (sumN,)")
    xy = None
    if field_distortion:
        if stamp_xy is None:
            raise ValueError("field_distortion needs stamp_xy (sumN,2)")
        xy = _clone_f32(stamp_xy, data)
        if tuple(xy.shape) != (sumN, 2):
            raise ValueError(f"stamp_xy must be ({sumN},2); got {tuple(xy.shape)}")
        out['distortion'] = _clone_f32(np.zeros((F, 6)) if distortion0 is None else distortion0, data)
    shapes = dict(narrow_psf=(F, nu, nu), full_psf=(F, nu, nu), residuals=(sumN, n, n), chi2=(F,),
                  loss_hist=(F, n_iter_adabelief), loss_hist_analytic=(F, n_iter_analytic),
                  W_out=(F, J, nu, nu), loss0=(F,), grad_b0=(F, nu, nu), grad_s0=(sumN, 3), status=(F,), grad_dist0=(F, 6))
    for nm in want:
        out[nm] = empty_like_kind(data, shapes[nm], 'i' if nm == 'status' else 'f')
    b = bounds or {}
    opts = _lib.PsfOpts(int(n_iter_analytic), int(n_iter_adabelief), float(lr), float(lam_scales), float(lam_hf),
                        (2 if noise_weights == 'MC' else int(bool(noise_weights))), float(b.get('fwhm_min', 1.0)),
                        float(b.get('fwhm_max', n / 2.0)), float(b.get('beta_min', 1.1)), float(b.get('beta_max', 12.0)),
                        int(mc_samples), int(mc_seed) & 0xffffffff, int(field_distortion))
    bi = _lib.PsfBatch(F, ptr(star_off), n, k, ptr(data), ptr(weight), ptr(W), ptr(xy))
    bo = _lib.PsfOut(*[ptr(out.get(nm)) for nm in _lib.PSF_OUT_FIELDS])
    rc = _lib.lib.lcb_psf_fit_batch(C.byref(bi), C.byref(opts), C.byref(bo), mem, current_stream(data))
    _lib.check(rc, 'lcb_psf_fit_batch')
    return out


def apply_distortion_batch(psfs, theta, psf_index, xy, mode=1):
    """lcb_apply_distortion_batch: psfs (Fp,nu,nu), theta (Fp,6), psf_index (B,), xy (B,2) -> (B,nu,nu) distorted PSFs."""
    _lib.require_device()
    psfs, theta, xy = as_f32(psfs), as_f32(theta), as_f32(xy)
    psf_index = as_i32(psf_index)
    mem = mem_kind(psfs, theta, psf_index, xy)
    B, Fp, nu = int(psf_index.shape[0]), int(psfs.shape[0]), int(psfs.shape[-1])
    if tuple(theta.shape) != (Fp, 6) or tuple(xy.shape) != (B, 2):
        raise ValueError("theta must be (Fp,6) and xy (B,2)")
    out = empty_like_kind(psfs, (B, nu, nu), 'f')
    _lib.check(_lib.lib.lcb_apply_distortion_batch(ptr(psfs), ptr(theta), ptr(psf_index), ptr(xy), B, Fp, nu, int(mode), ptr(out), mem,
                                                   current_stream(psfs)), 'lcb_apply_distortion_batch')
    return out


def _to_device(x, dtype):
    """numpy (pinned or pageable) / torch host or device array -> contiguous CUDA tensor of ``dtype`` on the current device.
    A numpy view whose rows are contiguous blocks at a constant pitch (``batch[:, lo:hi]`` of a C-contiguous array: what the
    in-process fan-out hands every device) goes up in ONE pitched asynchronous copy (``lcb_copy_2d``), without a host gather."""
    import torch
    if _lib.is_torch(x):
        return x.to(device='cuda', dtype=dtype, non_blocking=True).contiguous()
    want = {torch.float32: np.float32, torch.uint8: None, torch.int32: np.int32}.get(dtype)
    a = np.asarray(x)
    if (want is not None and a.dtype == want and a.ndim >= 2 and not a.flags.c_contiguous and a[0].flags.c_contiguous
            and a.strides[0] >= a[0].nbytes and a.shape[0] > 0 and a[0].nbytes > 0):
        out = torch.empty(a.shape, dtype=dtype, device='cuda')
        _lib.check(_lib.lib.lcb_copy_2d(out.data_ptr(), a[0].nbytes, a.ctypes.data, a.strides[0], a[0].nbytes, a.shape[0], 1,
                                        C.c_void_p(torch.cuda.current_stream().cuda_stream)), 'lcb_copy_2d')
        out._lcb_keepalive = a                     # the source must outlive the asynchronous copy
        return out
    a = np.ascontiguousarray(a)
    if dtype == torch.uint8:
        a = a.view(np.uint8) if a.dtype == np.bool_ else (a > 0).view(np.uint8)
    elif dtype == torch.int32:
        a = a.astype(np.int32, copy=False)
    elif a.dtype != np.float32:
        a = a.astype(np.float32)
    return torch.from_numpy(a).to('cuda', non_blocking=True)


def to_host(tensors):
    """dict of CUDA tensors -> dict of numpy arrays through PAGE-LOCKED staging (torch's caching host allocator keeps the
    buffers for the next call): the copies run at the PCIe rate and overlap each other; one synchronisation at the end."""
    import torch
    staged = {}
    for kk, v in tensors.items():
        if not v.is_cuda:
            staged[kk] = v
            continue
        h = torch.empty(v.shape, dtype=v.dtype, pin_memory=True)
        h.copy_(v, non_blocking=True)
        staged[kk] = h
    torch.cuda.current_stream().synchronize()
    return {kk: h.numpy() for kk, h in staged.items()}


GUESS_METHODS = {'center': 0, 'max': 1, 'barycenter': 2}


def psf_prepare_batch(images, noisemaps, masks, star_off, k, norm_scale=100.0, downsample_mean=False, guess_method='center'):
    """lcb_psf_prepare_batch: raw stamps (sumN,n,n) -> device tensors (data, weight, a0, x0, y0, norm)."""
    import torch
    _lib.require_device()
    if guess_method not in GUESS_METHODS:
        raise ValueError(f"guess_method_star_position must be 'center', 'max' or 'barycenter' (got {guess_method!r})")
    img, nm = _to_device(images, torch.float32), _to_device(noisemaps, torch.float32)
    mk = None if masks is None else _to_device(masks, torch.uint8)
    off = _to_device(np.asarray(star_off, np.int32), torch.int32) if not _lib.is_torch(star_off) else star_off.to('cuda', torch.int32)
    sumN, n = int(img.shape[0]), int(img.shape[-1])
    F = int(off.shape[0]) - 1
    out = dict(data=torch.empty_like(img), weight=torch.empty_like(img), a0=torch.empty(sumN, device='cuda'),
               x0=torch.empty(sumN, device='cuda'), y0=torch.empty(sumN, device='cuda'), norm=torch.empty(F, device='cuda'))
    pi = _lib.PsfPrepareIn(F, ptr(off), n, int(k), ptr(img), ptr(nm), ptr(mk), float(norm_scale), int(bool(downsample_mean)),
                           GUESS_METHODS[guess_method])
    po = _lib.PsfPrepareOut(*[ptr(out[nm_]) for nm_ in ('data', 'weight', 'a0', 'x0', 'y0', 'norm')])
    _lib.check(_lib.lib.lcb_psf_prepare_batch(C.byref(pi), C.byref(po), current_stream(img)), 'lcb_psf_prepare_batch')
    out['star_off'] = off
    return out


def phot_prepare_batch(data, noisemap, masks, k, downsample_mean=False):
    """lcb_phot_prepare_batch: raw stamps (F,S,n,n) -> device tensors (data, weight (F*S,n,n), a0 (F*S,), scale (S,))."""
    import torch
    _lib.require_device()
    d, nm = _to_device(data, torch.float32), _to_device(noisemap, torch.float32)
    mk = None if masks is None else _to_device(masks, torch.uint8)
    F, S, n, _ = d.shape
    out = dict(data=torch.empty((F * S, n, n), device='cuda'), weight=torch.empty((F * S, n, n), device='cuda'),
               a0=torch.empty(F * S, device='cuda'), scale=torch.empty(S, device='cuda'))
    work = torch.empty(int(_lib.lib.lcb_phot_prepare_work_floats(F, S)), device='cuda')
    pi = _lib.PhotPrepareIn(int(F), int(S), int(n), int(k), ptr(d), ptr(nm), ptr(mk), int(bool(downsample_mean)))
    po = _lib.PhotPrepareOut(*[ptr(out[nm_]) for nm_ in ('data', 'weight', 'a0', 'scale')])
    _lib.check(_lib.lib.lcb_phot_prepare_batch(C.byref(pi), C.byref(po), ptr(work), current_stream(d)), 'lcb_phot_prepare_batch')
    return out


def resolve_devices(devices):
    """devices: None / 1 (current device), 'all', an int count or a list of CUDA device indices -> list of indices."""
    _lib.require_device()                        # no CPU fallback: fail loudly before touching torch.cuda
    import torch
    if devices is None or devices == 1:
        return [torch.cuda.current_device()]
    if devices == 'all':
        return list(range(torch.cuda.device_count()))
    if isinstance(devices, int):
        return list(range(min(devices, torch.cuda.device_count())))
    return [int(d) for d in devices]


def split_by_work(work, parts):
    """Contiguous partition of len(work) items into ``parts`` blocks of roughly equal total work (SURVEY.md section 8e:
    'contiguous blocks of frames per GPU, balanced by sum N'); returns [(lo, hi), ...], empty blocks dropped."""
    work = np.asarray(work, np.float64)
    csum = np.concatenate([[0.0], np.cumsum(work)])
    bounds = [0]
    for p in range(1, parts):
        bounds.append(int(np.searchsorted(csum, csum[-1] * p / parts)))
    bounds.append(len(work))
    bounds = np.maximum.accumulate(bounds)
    return [(int(a), int(b)) for a, b in zip(bounds[:-1], bounds[1:]) if b > a]


def fan_out(blocks, devices, fn):
    """Runs fn(lo, hi) for every block on its own device from its own host thread (the library calls release the GIL;
    no data-path collective: the items are independent) and returns the results in block order."""
    import concurrent.futures
    import torch
    if len(blocks) == 1:
        with torch.cuda.device(devices[0]):
            return [fn(*blocks[0])]

    def run(i):
        with torch.cuda.device(devices[i % len(devices)]):
            return fn(*blocks[i])
    with concurrent.futures.ThreadPoolExecutor(max_workers=len(blocks)) as ex:
        return list(ex.map(run, range(len(blocks))))
