"""Conventions of the STARRED model as implemented by liblcb.

Everything tagged [R] in SURVEY.md (recalled from STARRED's public code, not verifiable in this
container) lives here, so that a golden vector produced with real STARRED can flip a convention
without touching the model code.  This is the product-side twin of ``oracle.conventions`` (the product never imports the oracle); a
CPU test asserts that both have identical fields and defaults.
"""
from dataclasses import dataclass, asdict


@dataclass(frozen=True)
class Conventions:
    # point-source / PSF "target resolution" Gaussian: FWHM in upsampled pixels, number of taps (even)
    gauss_fwhm_up: float = 2.0
    gauss_taps: int = 12
    # D_k: k x k block mean (True) or block sum (False).  Default SUM: lightcurver hands pixel sums to STARRED as initial
    # amplitudes and reads the fitted amplitudes back as fluxes (star_photometry.py:55-69,128; roi_modelling.py:199-212,462;
    # docs/example_starred_notebooks/example_roi_modelling.ipynb cells 13 -> 21 -> 36: a ~ sum of pixels / scale)
    downsample_mean: bool = False
    # chi2 term carries a factor 1/2
    chi2_half: bool = True
    # optax chain used when schedule_learning_rate=True
    clip_global_norm: float = 1.0
    lr_decay_rate: float = 0.99
    # optax.scale_by_belief defaults
    belief_b1: float = 0.9
    belief_b2: float = 0.999
    belief_eps: float = 1e-16
    belief_eps_root: float = 1e-16
    # build_psf: stamps are divided by max(image)/psf_norm_scale before fitting
    psf_norm_scale: float = 100.0
    # build_psf stage 2 (AdaBelief on the pixel grid) initial learning rate
    psf_stage2_lr: float = 1e-5
    # Moffat initial beta and bounds used by the analytic stage
    moffat_beta_init: float = 2.5
    moffat_beta_min: float = 1.1
    moffat_beta_max: float = 12.0
    moffat_fwhm_min: float = 1.0
    # regularisation strengths of build_psf stage 2
    psf_lambda_scales: float = 1.0
    psf_lambda_hf: float = 1.0
    # deconvolution Loss: regularization_strength_pts_source = L1 of the first starlet scale of the point-source
    # channel, weighted by W[0]; summed over all epochs (True) or evaluated on the first epoch only (False)
    pts_source_all_epochs: bool = True
    # regularization_strength_flux_uniformity: sum_m std_e(a_em) / |mean_e(a_em)| (True) or sum_m std_e(a_em) (False)
    flux_uniformity_relative: bool = True
    # field distortion (PSF(field_distortion=True), apply_distortion): dilation_x, dilation_y, shear are first-order polynomials
    # (no constant term) in the rescaled frame position; the resampled PSF is multiplied by |det A| (True) so that its integral is
    # kept, or left as sampled (False)
    distortion_conserve_flux: bool = True

    def as_dict(self):
        return asdict(self)

    def distortion_mode(self):
        """lcb_psf_opts.field_distortion / lcb_apply_distortion_batch mode."""
        return 1 if self.distortion_conserve_flux else 2

    def amplitude_per_flux(self, k):
        """Amplitude of a point source per unit of pixel-sum flux: 1 with the block sum, k^2 with the block mean."""
        return float(k * k) if self.downsample_mean else 1.0


DEFAULT = Conventions()


def apply_to_library(cv: Conventions = DEFAULT):
    """Pushes the kernel-visible subset of ``cv`` into liblcb (lcb_conventions_set)."""
    from . import _lib
    _lib.set_conventions(gauss_fwhm_up=cv.gauss_fwhm_up, gauss_taps=cv.gauss_taps,
                         downsample_mean=int(cv.downsample_mean), chi2_half=int(cv.chi2_half),
                         clip_global_norm=cv.clip_global_norm, lr_decay_rate=cv.lr_decay_rate,
                         belief_b1=cv.belief_b1, belief_b2=cv.belief_b2, belief_eps=cv.belief_eps,
                         belief_eps_root=cv.belief_eps_root)
