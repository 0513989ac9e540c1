"""lightcurver_b200 -- B200 (sm_100a) drop-in for the STARRED fits of lightcurver.

Host code is Python and mirrors the reference's call surface for this path only:

  * ``lightcurver_b200.procedures.psf_routines.build_psf``            (starred build_psf as called at
    lightcurver/processes/psf_modelling.py:164-171)
  * ``lightcurver_b200.processes.star_photometry.do_one_star_forward_modelling``
    (lightcurver/processes/star_photometry.py:23-151)
  * ``lightcurver_b200.processes.roi_modelling.joint_deconvolution``  (roi_modelling.py:213-335)
  * ``lightcurver_b200.utilities.starred_utilities.get_flux_uncertainties``

Everything numerical happens in ``liblcb.so`` (CUDA, C ABI in ``include/lcb.h``); there is no CPU
fallback and importing the compute modules without the built library raises ImportError.
"""
__version__ = '0.1.0'
