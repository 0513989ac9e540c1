"""ctypes binding of liblcb.so (the C ABI declared in include/lcb.h).

The product has NO CPU fallback: importing this module without the built library raises, and every
compute entry returns LCB_ERR_CUDA (-> RuntimeError) when no CUDA device is visible.
"""
import ctypes as C
import os
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ.get('LCB_LIBRARY', _HERE / 'liblcb.so'))

if not LIB_PATH.exists():
    raise ImportError(
        f"{LIB_PATH} is missing: build the sm_100a kernels first "
        f"(python -m lightcurver_b200.build, or __graft_entry__.build()). "
        f"lightcurver_b200 has no CPU fallback.")

lib = C.CDLL(str(LIB_PATH))

MEM_DEVICE, MEM_HOST = 0, 1
c_fp = C.POINTER(C.c_float)
c_ip = C.POINTER(C.c_int)


class Conventions(C.Structure):
    _fields_ = [('gauss_fwhm_up', C.c_float), ('gauss_taps', C.c_int), ('downsample_mean', C.c_int),
                ('chi2_half', C.c_int), ('clip_global_norm', C.c_float), ('lr_decay_rate', C.c_float),
                ('belief_b1', C.c_float), ('belief_b2', C.c_float), ('belief_eps', C.c_float),
                ('belief_eps_root', C.c_float)]


class FitOpts(C.Structure):
    _fields_ = [('n_iter', C.c_int), ('lr', C.c_float), ('schedule', C.c_int)]


class PhotBatch(C.Structure):
    _fields_ = [('B', C.c_int), ('n', C.c_int), ('k', C.c_int),
                ('data', C.c_void_p), ('weight', C.c_void_p), ('psf', C.c_void_p), ('psf_index', C.c_void_p),
                ('Fp', C.c_int), ('a0', C.c_void_p), ('dx0', C.c_void_p), ('dy0', C.c_void_p)]


class PhotOut(C.Structure):
    _fields_ = [(nm, C.c_void_p) for nm in ('a', 'dx', 'dy', 'sigma_a', 'chi2', 'residuals', 'loss_hist',
                                            'loss0', 'grad0', 'status')]


lib.lcb_last_error.restype = C.c_char_p
lib.lcb_version.restype = C.c_int
lib.lcb_device_count.restype = C.c_int
lib.lcb_conventions_get.argtypes = [C.POINTER(Conventions)]
lib.lcb_conventions_set.argtypes = [C.POINTER(Conventions)]
lib.lcb_phot_fit_batch.argtypes = [C.POINTER(PhotBatch), C.POINTER(FitOpts), C.POINTER(PhotOut), C.c_int, C.c_void_p]
lib.lcb_phot_fit_batch.restype = C.c_int
lib.lcb_fp32_peak.argtypes = [C.c_int, c_fp, c_fp]
lib.lcb_fp32_peak.restype = C.c_int


class LcbError(RuntimeError):
    pass


def check(rc, what):
    if rc != 0:
        raise LcbError(f"{what} failed (status {rc}): {lib.lcb_last_error().decode(errors='replace')}")


def device_count():
    return int(lib.lcb_device_count())


def require_device():
    if device_count() == 0:
        raise LcbError("lightcurver_b200 needs a CUDA device (sm_100a); there is no CPU fallback")


# ------------------------------------------------------------------ array plumbing
def is_torch(x):
    return type(x).__module__.startswith('torch')


def ptr(x):
    """Raw address of a C-contiguous numpy array or torch tensor (None -> NULL)."""
    if x is None:
        return None
    if is_torch(x):
        assert x.is_contiguous()
        return C.c_void_p(x.data_ptr())
    assert x.flags['C_CONTIGUOUS']
    return C.c_void_p(x.ctypes.data)


def as_f32(x, like_torch=None):
    """C-contiguous float32 view/copy; numpy stays numpy, torch stays torch."""
    if x is None:
        return None
    if is_torch(x):
        import torch
        return x.to(torch.float32).contiguous()
    return np.ascontiguousarray(x, dtype=np.float32)


def as_i32(x):
    if x is None:
        return None
    if is_torch(x):
        import torch
        return x.to(torch.int32).contiguous()
    return np.ascontiguousarray(x, dtype=np.int32)


def empty_like_kind(ref, shape, dtype='f'):
    """Uninitialised output of the same kind (numpy / torch device) as ``ref``."""
    if is_torch(ref):
        import torch
        return torch.empty(shape, dtype=torch.float32 if dtype == 'f' else torch.int32, device=ref.device)
    return np.empty(shape, dtype=np.float32 if dtype == 'f' else np.int32)


def mem_kind(*arrays):
    kinds = set()
    for a in arrays:
        if a is None:
            continue
        if is_torch(a):
            kinds.add(MEM_DEVICE if a.is_cuda else MEM_HOST)
        else:
            kinds.add(MEM_HOST)
    if len(kinds) != 1:
        raise ValueError("all arrays of one call must live in the same memory (all host or all device)")
    return kinds.pop()


def current_stream(ref):
    if is_torch(ref) and ref.is_cuda:
        import torch
        return C.c_void_p(torch.cuda.current_stream(ref.device).cuda_stream)
    return None


def get_conventions():
    c = Conventions()
    check(lib.lcb_conventions_get(C.byref(c)), 'lcb_conventions_get')
    return c


def set_conventions(**kw):
    c = get_conventions()
    for k, v in kw.items():
        if not hasattr(c, k):
            raise KeyError(k)
        setattr(c, k, v)
    check(lib.lcb_conventions_set(C.byref(c)), 'lcb_conventions_set')


lib.lcb_fp32_peak_rrr.argtypes = [C.c_int, c_fp, c_fp]
lib.lcb_fp32x2_peak.argtypes = [C.c_int, C.c_int, c_fp, c_fp]


def fp32_peak_rrr(iters=4096):
    require_device()
    t, ms = C.c_float(0), C.c_float(0)
    check(lib.lcb_fp32_peak_rrr(iters, C.byref(t), C.byref(ms)), 'lcb_fp32_peak_rrr')
    return float(t.value), float(ms.value)


def fp32x2_peak(which=0, iters=4096):
    require_device()
    t, ms = C.c_float(0), C.c_float(0)
    check(lib.lcb_fp32x2_peak(iters, which, C.byref(t), C.byref(ms)), 'lcb_fp32x2_peak')
    return float(t.value), float(ms.value)


def fp32_peak(iters=4096):
    require_device()
    t, ms = C.c_float(0), C.c_float(0)
    check(lib.lcb_fp32_peak(iters, C.byref(t), C.byref(ms)), 'lcb_fp32_peak')
    return float(t.value), float(ms.value)


class PsfBatch(C.Structure):
    _fields_ = [('F', C.c_int), ('star_off', C.c_void_p), ('n', C.c_int), ('k', C.c_int),
                ('data', C.c_void_p), ('weight', C.c_void_p), ('W', C.c_void_p), ('stamp_xy', C.c_void_p)]


class PsfOpts(C.Structure):
    _fields_ = [('n_iter_analytic', C.c_int), ('n_iter_adabelief', C.c_int), ('lr', C.c_float),
                ('lam_scales', C.c_float), ('lam_hf', C.c_float), ('noise_weights', C.c_int),
                ('fwhm_min', C.c_float), ('fwhm_max', C.c_float), ('beta_min', C.c_float), ('beta_max', C.c_float),
                ('mc_samples', C.c_int), ('mc_seed', C.c_uint), ('field_distortion', C.c_int)]


PSF_OUT_FIELDS = ('moffat', 'a', 'x0', 'y0', 'background', 'narrow_psf', 'full_psf', 'residuals', 'chi2',
                  'loss_hist', 'loss_hist_analytic', 'W_out', 'loss0', 'grad_b0', 'grad_s0', 'status', 'distortion', 'grad_dist0')


class PsfOut(C.Structure):
    _fields_ = [(nm, C.c_void_p) for nm in PSF_OUT_FIELDS]


lib.lcb_starlet_scales.argtypes = [C.c_int]
lib.lcb_starlet_scales.restype = C.c_int
lib.lcb_psf_fit_batch.argtypes = [C.POINTER(PsfBatch), C.POINTER(PsfOpts), C.POINTER(PsfOut), C.c_int, C.c_void_p]
lib.lcb_psf_fit_batch.restype = C.c_int
lib.lcb_apply_distortion_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                           C.c_void_p, C.c_int, C.c_void_p]
lib.lcb_apply_distortion_batch.restype = C.c_int
lib.lcb_copy_2d.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, C.c_void_p]
lib.lcb_norm_medians.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
lib.lcb_norm_scatter_work_doubles.argtypes = [C.c_int, C.c_int]
lib.lcb_norm_scatter_matrix.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
lib.lcb_norm_coefficients.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
lib.lcb_zeropoints.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]


lib.lcb_profile_enable.argtypes = [C.c_int]
lib.lcb_profile_summary.argtypes = [C.c_char_p, C.c_int]


def profile_enable(on=True):
    check(lib.lcb_profile_enable(int(bool(on))), 'lcb_profile_enable')


def profile_summary():
    import json
    buf = C.create_string_buffer(1 << 16)
    check(lib.lcb_profile_summary(buf, len(buf)), 'lcb_profile_summary')
    return json.loads(buf.value.decode())


class DeconvProblem(C.Structure):
    _fields_ = [('E', C.c_int), ('n', C.c_int), ('k', C.c_int), ('P', C.c_int), ('M', C.c_int),
                ('data', C.c_void_p), ('weight', C.c_void_p), ('psf', C.c_void_p)]


class DeconvParams(C.Structure):
    _fields_ = [(nm, C.c_void_p) for nm in ('h', 'mean', 'a', 'c_x', 'c_y', 'dx', 'dy', 'alpha')] + \
               [(nm, C.c_int) for nm in ('free_h', 'free_mean', 'free_a', 'free_c', 'free_d')]


class DeconvReg(C.Structure):
    _fields_ = [('lam_scales', C.c_float), ('lam_hf', C.c_float), ('lam_pos', C.c_float), ('W', C.c_void_p),
                ('prior_mu_x', C.c_void_p), ('prior_sig_x', C.c_void_p), ('prior_mu_y', C.c_void_p), ('prior_sig_y', C.c_void_p),
                ('lam_pts', C.c_float), ('lam_fu', C.c_float), ('pts_all_epochs', C.c_int), ('fu_relative', C.c_int)]


class DeconvGrad(C.Structure):
    _fields_ = [(nm, C.c_void_p) for nm in ('loss', 'h', 'mean', 'a', 'c_x', 'c_y', 'dx', 'dy')]


lib.lcb_deconv_create.argtypes = [C.POINTER(DeconvProblem), C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]
lib.lcb_deconv_set_params.argtypes = [C.c_void_p, C.POINTER(DeconvParams), C.c_int]
lib.lcb_deconv_set_reg.argtypes = [C.c_void_p, C.POINTER(DeconvReg), C.c_int]
lib.lcb_deconv_run.argtypes = [C.c_void_p, C.POINTER(FitOpts), C.c_void_p, C.c_int]
lib.lcb_deconv_run_many.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.POINTER(FitOpts), C.POINTER(C.c_void_p), C.c_int]
lib.lcb_deconv_step_local.argtypes = [C.c_void_p, C.c_int]
lib.lcb_deconv_reduce_buffer.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int)]
lib.lcb_deconv_step_update.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_int]
lib.lcb_deconv_flush.argtypes = [C.c_void_p]
lib.lcb_deconv_loss_grad.argtypes = [C.c_void_p, C.POINTER(DeconvGrad), C.c_int]
lib.lcb_deconv_lbfgs.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_int]
lib.lcb_deconv_get.argtypes = [C.c_void_p, C.POINTER(DeconvParams), C.c_void_p, C.c_void_p, C.c_int]
lib.lcb_deconv_destroy.argtypes = [C.c_void_p]
lib.lcb_deconv_noise_weights.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
lib.lcb_deconv_set_cluster.argtypes = [C.c_void_p, C.c_int]
lib.lcb_deconv_get_cluster.argtypes = [C.c_void_p]
lib.lcb_deconv_set_global.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int]
lib.lcb_deconv_comm_init.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
lib.lcb_deconv_comm_connect.argtypes = [C.c_void_p, C.c_void_p]


class PsfPrepareIn(C.Structure):
    _fields_ = [('F', C.c_int), ('star_off', C.c_void_p), ('n', C.c_int), ('k', C.c_int), ('image', C.c_void_p),
                ('noisemap', C.c_void_p), ('mask', C.c_void_p), ('norm_scale', C.c_float), ('downsample_mean', C.c_int),
                ('guess_method', C.c_int)]


class PsfPrepareOut(C.Structure):
    _fields_ = [(nm, C.c_void_p) for nm in ('data', 'weight', 'a0', 'x0', 'y0', 'norm')]


class PhotPrepareIn(C.Structure):
    _fields_ = [('F', C.c_int), ('S', C.c_int), ('n', C.c_int), ('k', C.c_int), ('data', C.c_void_p),
                ('noisemap', C.c_void_p), ('mask', C.c_void_p), ('downsample_mean', C.c_int)]


class PhotPrepareOut(C.Structure):
    _fields_ = [(nm, C.c_void_p) for nm in ('data', 'weight', 'a0', 'scale')]


lib.lcb_psf_prepare_batch.argtypes = [C.POINTER(PsfPrepareIn), C.POINTER(PsfPrepareOut), C.c_void_p]
lib.lcb_phot_prepare_work_floats.argtypes = [C.c_int, C.c_int]
lib.lcb_phot_prepare_work_floats.restype = C.c_size_t
lib.lcb_phot_prepare_batch.argtypes = [C.POINTER(PhotPrepareIn), C.POINTER(PhotPrepareOut), C.c_void_p, C.c_void_p]
