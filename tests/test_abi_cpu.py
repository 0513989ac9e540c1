"""No-GPU checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/lcb.h declares, round-trips the conventions, and refuses to compute without a device
(there is no CPU fallback)."""
import ctypes
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / 'include' / 'lcb.h').read_text()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(lcb_[a-z0-9_]+)\s*\(', text)))


def test_header_symbols_exported(lcb):
    syms = declared_symbols()
    assert 'lcb_psf_fit_batch' in syms and 'lcb_phot_fit_batch' in syms
    for s in syms:
        assert hasattr(lcb.lib, s), f"liblcb.so does not export {s}"


def test_conventions_roundtrip(lcb):
    c = lcb.get_conventions()
    assert c.gauss_taps == 12 and abs(c.gauss_fwhm_up - 2.0) < 1e-7 and c.downsample_mean == 0
    lcb.set_conventions(gauss_taps=16)
    assert lcb.get_conventions().gauss_taps == 16
    with pytest.raises(lcb.LcbError):
        lcb.set_conventions(gauss_taps=11)
    lcb.set_conventions(gauss_taps=12)
    from lightcurver_b200.conventions import apply_to_library, DEFAULT
    apply_to_library(DEFAULT)
    assert lcb.get_conventions().gauss_taps == 12


def test_starlet_scales(lcb):
    assert [lcb.lib.lcb_starlet_scales(v) for v in (16, 32, 48, 64, 128, 192)] == [4, 5, 5, 6, 7, 7]


def test_no_cpu_fallback(lcb):
    """Without a device the product fails loudly (skipped on the GPU box)."""
    if lcb.device_count() > 0:
        pytest.skip("a CUDA device is visible")
    from lightcurver_b200 import engine
    z = np.zeros((1, 16, 16), np.float32)
    with pytest.raises(lcb.LcbError):
        engine.phot_fit_batch(z, z, z, np.zeros(1, np.int32), np.ones(1, np.float32), 1, 5)
    from lightcurver_b200.procedures.psf_routines import build_psf
    with pytest.raises(lcb.LcbError):
        build_psf(np.ones((2, 16, 16)), np.ones((2, 16, 16)), 1, n_iter_analytic=1, n_iter_adabelief=1)


def test_product_does_not_import_oracle():
    import subprocess, sys
    code = ("import sys; import lightcurver_b200, lightcurver_b200.engine, lightcurver_b200.procedures.psf_routines, "
            "lightcurver_b200.processes.star_photometry, lightcurver_b200.processes.roi_modelling, lightcurver_b200.starred_api, "
            "lightcurver_b200.utilities.starred_utilities; "
            "assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules), 'product imported the oracle'")
    subprocess.run([sys.executable, '-c', code], check=True, cwd=str(ROOT))
