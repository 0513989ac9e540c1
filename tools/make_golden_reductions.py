"""Golden vectors for the normalisation-coefficient and zero-point reductions, produced by the UNMODIFIED reference code.

Runs only in the build container (needs /root/reference): lightcurver's own modules
``lightcurver/processes/normalization_calculation.py`` and ``lightcurver/processes/absolute_zeropoint_calculation.py`` are
loaded from where they lie (never copied) with their package-level collaborators (config, SQL helpers, catalog look-ups,
plotting -- all outside the hot path and not importable here: astropy, h5py ... are missing) replaced by small stand-ins that
serve a synthetic flux table; ``calculate_coefficient()`` and ``calculate_zeropoints()`` then run as they are and write their
rows into a temporary SQLite database, which is read back.

    python tools/make_golden_reductions.py      ->  tests/golden/reference_reductions.npz
"""
import importlib.util
import sqlite3
import sys
import tempfile
import types
from pathlib import Path

import numpy as np
import pandas as pd

ROOT = Path(__file__).resolve().parent.parent
REF = Path('/root/reference/lightcurver')


def synthetic_table(F=80, S=7, seed=20260118):
    rng = np.random.default_rng(seed)
    star_flux = 10.0 ** rng.uniform(3.5, 5.0, S)
    transparency = rng.lognormal(0.0, 0.1, F)
    flux = transparency[:, None] * star_flux[None] * (1 + 0.01 * rng.standard_normal((F, S)))
    d_flux = 0.01 * flux * rng.uniform(0.5, 2.0, (F, S))
    chi2 = rng.uniform(0.5, 1.5, (F, S))
    chi2[rng.random((F, S)) < 0.05] = 9.0                      # outside the chi2 window -> the SQL filter drops the row
    missing = rng.random((F, S)) < 0.04                         # star not measured in that frame
    catalog_mag = 25.0 - 2.5 * np.log10(star_flux) + 0.02 * rng.standard_normal(S)
    return dict(flux=flux, d_flux=d_flux, chi2=chi2, missing=missing, catalog_mag=catalog_mag, transparency=transparency,
                chi2_bounds=np.array([0.0, 2.0]))


def load_reference(t, db_path):
    """The two reference modules with stand-in collaborators."""
    F, S = t['flux'].shape
    lo, hi = t['chi2_bounds']
    ok = ~t['missing']
    rows = [dict(name=f's{s}', frame_id=f + 1, mjd=60000.0 + f, star_gaia_id=1000 + s, combined_footprint_hash=7,
                 flux=t['flux'][f, s], d_flux=t['d_flux'][f, s], chi2=t['chi2'][f, s]) for s in range(S) for f in range(F) if ok[f, s]]
    table = pd.DataFrame(rows)

    def execute_sqlite_query(query, params=(), is_select=True, use_pandas=False):
        if 'normalization' in query or 'star_flux_in_frame sff ON f.id' in query:        # get_fluxes (:29-48), BETWEEN filter
            df = table[(table['chi2'] >= params[1]) & (table['chi2'] <= params[2])]
            return df.sort_values(['name', 'frame_id'])[['name', 'frame_id', 'mjd', 'star_gaia_id', 'combined_footprint_hash', 'flux', 'd_flux']]
        if 'SELECT DISTINCT star_gaia_id' in query:
            return [(g,) for g in sorted(table['star_gaia_id'].unique())]
        if 'catalog_star_photometry csp' in query:                                         # absolute_zeropoint_calculation.py:64-84
            df = table[['frame_id', 'flux', 'star_gaia_id']].rename(columns={'star_gaia_id': 'gaia_id'}).copy()
            df['catalog_mag'] = t['catalog_mag'][df['gaia_id'].values - 1000]
            return df
        raise AssertionError(query)

    cfg = dict(database_path=db_path, stars_to_use_norm='all', reference_absolute_photometric_survey='gaia',
               plots_dir=Path(tempfile.mkdtemp()))
    stubs = {
        'lightcurver': {}, 'lightcurver.processes': {}, 'lightcurver.structure': {}, 'lightcurver.utilities': {}, 'lightcurver.plotting': {},
        'lightcurver.structure.database': dict(execute_sqlite_query=execute_sqlite_query,
                                               get_pandas=lambda **kw: pd.DataFrame(dict(id=np.arange(1, F + 1)))),
        'lightcurver.structure.user_config': dict(get_user_config=lambda: cfg),
        'lightcurver.utilities.footprint': dict(get_combined_footprint_hash=lambda *a, **k: 7),
        'lightcurver.utilities.chi2_selector': dict(get_chi2_bounds=lambda psf_or_fluxes: (float(lo), float(hi))),
        'lightcurver.plotting.normalization_plotting': dict(plot_normalized_star_curves=lambda **kw: None),
        'lightcurver.utilities.absolute_magnitudes_from_panstarrs': dict(save_panstarrs_catalog_photometry_to_database=lambda g: None),
        'lightcurver.utilities.absolute_magnitudes_from_gaia': dict(save_gaia_catalog_photometry_to_database=lambda g: None),
    }
    for name, syms in stubs.items():
        mod = types.ModuleType(name)
        mod.__dict__.update(syms)
        if not syms:
            mod.__path__ = []
        sys.modules[name] = mod
    out = {}
    for nm in ('normalization_calculation', 'absolute_zeropoint_calculation'):
        spec = importlib.util.spec_from_file_location(f'lightcurver.processes.{nm}', REF / 'processes' / f'{nm}.py')
        mod = importlib.util.module_from_spec(spec)
        sys.modules[spec.name] = mod
        spec.loader.exec_module(mod)
        out[nm] = mod

    # pandas >= 3 no longer accepts a plain list in pd.unique (absolute_zeropoint_calculation.py:58 passes one): give that one
    # module the pandas < 3 behaviour (an environment shim, the reference file is untouched)
    class _Pandas:
        def __getattr__(self, name):
            return getattr(pd, name)

        @staticmethod
        def unique(values):
            return pd.unique(np.asarray(values))
    out['absolute_zeropoint_calculation'].pd = _Pandas()
    return out


def main():
    t = synthetic_table()
    F, S = t['flux'].shape
    db_path = Path(tempfile.mkdtemp()) / 'db.sqlite3'
    with sqlite3.connect(db_path) as conn:
        conn.execute("CREATE TABLE normalization_coefficients (frame_id INTEGER, combined_footprint_hash INTEGER, coefficient REAL, "
                     "coefficient_uncertainty REAL, PRIMARY KEY (combined_footprint_hash, frame_id))")
        conn.execute("CREATE TABLE absolute_zeropoints (frame_id INTEGER, combined_footprint_hash INTEGER, zeropoint REAL, "
                     "zeropoint_uncertainty REAL, source_catalog TEXT, PRIMARY KEY (combined_footprint_hash, frame_id))")
    mods = load_reference(t, str(db_path))
    mods['normalization_calculation'].calculate_coefficient()
    mods['absolute_zeropoint_calculation'].calculate_zeropoints()
    with sqlite3.connect(db_path) as conn:
        norm = conn.execute("SELECT frame_id, coefficient, coefficient_uncertainty FROM normalization_coefficients ORDER BY frame_id").fetchall()
        zps = conn.execute("SELECT frame_id, zeropoint, zeropoint_uncertainty FROM absolute_zeropoints ORDER BY frame_id").fetchall()
    coef, err, zp, zs = (np.full(F, np.nan) for _ in range(4))
    for fid, c, e in norm:
        coef[fid - 1], err[fid - 1] = c, e
    for fid, z, s in zps:
        zp[fid - 1] = z
        zs[fid - 1] = np.nan if s is None else s
    dst = ROOT / 'tests' / 'golden' / 'reference_reductions.npz'
    np.savez(dst, flux=t['flux'], d_flux=t['d_flux'], chi2=t['chi2'], missing=t['missing'], catalog_mag=t['catalog_mag'],
             chi2_bounds=t['chi2_bounds'], transparency=t['transparency'],
             ref_coefficient=coef, ref_coefficient_uncertainty=err, ref_zeropoint=zp, ref_zeropoint_uncertainty=zs)
    print('wrote', dst, 'coef[:4]', coef[:4], 'zp[:4]', zp[:4])


if __name__ == '__main__':
    main()
