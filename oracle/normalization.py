"""CPU restatement of the reductions that follow star_photometry (TEST INFRASTRUCTURE, see ``oracle/__init__.py``):

  * ``calculate_coefficient``  lightcurver/processes/normalization_calculation.py:157-206 -- kept in the reference's own terms
    (pandas pivots, ``cost_function_scatter_in_frame`` :75-98, scipy SLSQP under mean(c) = 1 :182-187, ``weighted_std`` :118-131)
    because pandas and scipy are available here even though lightcurver itself cannot be imported (its module imports need
    astropy / h5py); each step cites the line it follows.
  * ``zeropoints``             lightcurver/processes/absolute_zeropoint_calculation.py:95-100.
"""
import numpy as np
import pandas as pd
from scipy.optimize import minimize


def cost_function_scatter_in_frame(scaling_factors, normalized_flux_pivot, normalized_d_flux_pivot):
    """normalization_calculation.py:75-98."""
    scaled_fluxes = normalized_flux_pivot.mul(scaling_factors, axis=0)
    weights = 1 / normalized_d_flux_pivot
    weighted_means = (scaled_fluxes * weights).sum(axis=0) / weights.sum(axis=0)
    return ((weights.mul((scaled_fluxes.sub(weighted_means, axis='columns')) ** 2)).sum(axis=0) / weights.sum(axis=0)).sum()


def weighted_std(values, weights):
    """normalization_calculation.py:118-131."""
    values, weights = np.asarray(values, float), np.asarray(weights, float)
    isnan = np.isnan(values) | np.isnan(weights)
    values, weights = values[~isnan], weights[~isnan]
    average = np.average(values, weights=weights)
    return np.sqrt(np.average((values - average) ** 2, weights=weights))


def calculate_coefficient(flux, d_flux, tol=1e-12):
    """flux, d_flux (F, S) with NaN = missing -> dict(coefficient (F,), coefficient_uncertainty (F,), star_scaling (S,),
    median_flux (S,)).  Follows normalization_calculation.py:157-204 on the long-format table the SQL query of :29-48 returns."""
    F, S = flux.shape
    fid, sid = np.meshgrid(np.arange(F), np.arange(S), indexing='ij')
    df = pd.DataFrame(dict(frame_id=fid.ravel(), star_gaia_id=sid.ravel(), flux=flux.ravel(), d_flux=d_flux.ravel()))
    df = df[~(df['flux'].isna() | df['d_flux'].isna())]
    median_flux_per_star = df.groupby('star_gaia_id')['flux'].median().rename('median_flux')          # :158
    df2 = df.merge(median_flux_per_star, on='star_gaia_id')
    df2['normalized_flux'] = df2['flux'] / df2['median_flux']                                             # :160-161
    df2['normalized_d_flux'] = df2['d_flux'] / df2['median_flux']
    fp = df2.pivot(index='star_gaia_id', columns='frame_id', values='normalized_flux')                    # :165-167
    dp = df2.pivot(index='star_gaia_id', columns='frame_id', values='normalized_d_flux')
    constraint = ({'type': 'eq', 'fun': lambda coeffs: 1 - np.nanmean(coeffs)})                          # :182
    result = minimize(cost_function_scatter_in_frame, np.ones(fp.shape[0]), args=(fp, dp), constraints=constraint,
                      method='SLSQP', tol=tol)                                                            # :184-186
    c = result.x
    adj, dadj = fp.mul(c, axis=0), dp.mul(c, axis=0)                                                      # :189-190
    w = 1. / dadj ** 2                                                                                    # :199
    norm_err = pd.Series([weighted_std(adj[f], w[f]) for f in adj.columns], index=adj.columns)            # :200-205
    norm_coeff = (adj.multiply(w)).sum(axis=0) / w.sum(axis=0)                                            # :202
    norm_err.loc[norm_err == 0.] = 0.1 * norm_coeff.loc[norm_err == 0.]                                   # :207
    coef, err = np.full(F, np.nan), np.full(F, np.nan)
    coef[norm_coeff.index.values] = norm_coeff.values
    err[norm_err.index.values] = norm_err.values
    scaling, med = np.ones(S), np.full(S, np.nan)
    scaling[fp.index.values] = c
    med[median_flux_per_star.index.values] = median_flux_per_star.values
    return dict(coefficient=coef, coefficient_uncertainty=err, star_scaling=scaling, median_flux=med, slsqp=result)


def zeropoints(flux, catalog_mag):
    """absolute_zeropoint_calculation.py:95-100: flux (F, S) NaN = missing, catalog_mag (S,) -> (median, std) per frame."""
    F, S = flux.shape
    fid, sid = np.meshgrid(np.arange(F), np.arange(S), indexing='ij')
    fd = pd.DataFrame(dict(frame_id=fid.ravel(), flux=flux.ravel(), catalog_mag=np.broadcast_to(catalog_mag, (F, S)).ravel()))
    fd = fd[~fd['flux'].isna()]
    with np.errstate(invalid='ignore', divide='ignore'):
        fd['instrumental_mag'] = -2.5 * np.log10(fd['flux'])
    fd['mag_difference'] = fd['catalog_mag'] - fd['instrumental_mag']
    res = fd.groupby('frame_id')['mag_difference'].agg(['median', 'std'])
    zp, zs = np.full(F, np.nan), np.full(F, np.nan)
    zp[res.index.values] = res['median'].values
    zs[res.index.values] = res['std'].values
    return zp, zs
