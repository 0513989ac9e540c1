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
