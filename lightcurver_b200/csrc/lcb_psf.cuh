// lcb_psf.cuh -- argument block shared by the PSF-fit kernels (K1a Moffat LM, K5 noise weights,
// K1 pixel-grid AdaBelief).
#pragma once
#include "lcb_passes.cuh"

#define PSF_THREADS 512
#define PSF_WARPS (PSF_THREADS / 32)
#define LCB_JMAX 8

struct PsfArgs {
    int F, n, k, nu, J, n_iter, n_iter_lm, Nmax, planes_in_smem, jim_in_smem;
    float lr, lam_scales, lam_hf;
    const int* star_off;            // [F+1] CSR offsets into the star arrays
    const float *data, *weight;     // [sumN][n][n]
    const float* W;                 // [F][J][nu*nu] or NULL (== 1)
    float* s_fixed;                 // [F][nu*nu]  C * Moffat (input of stage 2, output of stage 1)
    float* b;                       // [F][nu*nu]  background grid, in: initial, out: fitted
    float *a, *x0, *y0;             // [sumN]      in: initial, out: fitted
    float* moffat;                  // [F][5]      fwhm_x, fwhm_y, phi, beta, C (stage 1 in/out)
    float* loss_hist;               // [F][n_iter] or NULL
    float* loss_hist_lm;            // [F][n_iter_lm] or NULL
    float* residuals;               // [sumN][n][n] or NULL
    float* chi2;                    // [F] or NULL
    float *narrow_psf, *full_psf;   // [F][nu*nu] or NULL
    float *loss0, *grad_b0, *grad_s0;  // [F], [F][nu*nu], [sumN][3]: loss and gradient at the initial point
    float* work;                    // workspace, see lcb_psf_workspace_floats()
    size_t work_per_frame;          // floats
    int* status;                    // [F]
    float fwhm_min, fwhm_max, beta_min, beta_max;
    // field distortion (lcb_distort.cuh): 0 = off, 1 = flux-conserving affine resampling, 2 = plain resampling
    int distort;
    const float* stamp_xy;          // [sumN][2] rescaled frame coordinates of the stamps (image_coordinates.py:4-25)
    float* distortion;              // [F][6] theta, in: initial, out: fitted
    float* grad_dist0;              // [F][6] d loss / d theta at the initial point, or NULL
    DevConv cv;
};
