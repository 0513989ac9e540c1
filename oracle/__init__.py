"""CPU oracle for the lightcurver -> STARRED hot path.  TEST INFRASTRUCTURE ONLY.

This package is a CPU restatement (PyTorch float64 / float32 with autograd standing in for
``jax.grad``; ``scipy.optimize`` L-BFGS-B for the analytic stages) of the algorithm that
lightcurver's ``psf_modeling``, ``star_photometry`` and ROI-modelling steps hand to the third-party
package ``starred-astro >= 1.4.7`` (reference ``pyproject.toml:26``; NOT vendored under
``/root/reference`` and not installable here: no network, no jax).

PARITY UNPINNED: the reference's own tests hold no golden vector on this path
(``tests/test_starred_calls/test_starred_calls.py`` pins dict keys/shapes/types only;
``tests/test_entire_pipeline/test_run_pipeline_example_config.py:18-21`` pins ``chi2 < 2``), and
STARRED cannot be run here.  Every constant that is recalled rather than verified is a field of
:class:`oracle.conventions.Conventions`; ``tools/dump_starred_vectors.py`` produces golden vectors
wherever STARRED is installed.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this
package, and only as the checker / reported CPU baseline.  The product (``lightcurver_b200``) never
imports it and has no CPU fallback.
"""
from .conventions import Conventions  # noqa: F401
