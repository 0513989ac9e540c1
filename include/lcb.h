/* lcb.h -- C ABI of liblcb.so: B200 (sm_100a) kernels for lightcurver's STARRED hot path.
 *
 * Every entry point replaces one group of calls that lightcurver makes into the third-party
 * `starred` package (paths relative to the lightcurver checkout, v1.2.3):
 *
 *   lcb_phot_fit_batch     star_photometry.py:66-128  setup_model / Loss / Optimizer('adabelief').minimize /
 *                                                      model.model  (fixed-PSF amplitude+shift fit)
 *                          + starred_utilities.py:10-39 get_flux_uncertainties (closed form, sigma_a)
 *   lcb_psf_fit_batch      psf_modelling.py:164-171    starred.procedures.psf_routines.build_psf  (loss0 / grad_b0 / grad_s0 of
 *                                                      lcb_psf_out expose one evaluation of its Loss object to the parity tests;
 *                                                      noise_weights = 1 / 2 is propagate_noise(method='SLIT' / 'MC') [R])
 *   lcb_deconv_noise_weights  star_photometry.py:108, roi_modelling.py:299  propagate_noise(model, ..., method='SLIT')
 *   lcb_deconv_*           roi_modelling.py:213-335    setup_model / Loss (chi2, starlet-L1, positivity, prior, pts-source,
 *                                                      flux uniformity) / Optimizer('adabelief').minimize / the loss-gradient
 *                                                      pair of Optimizer('l-bfgs-b'); epochs shard over GPUs (lcb_deconv_comm_*)
 *   lcb_psf_prepare_batch  psf_modelling.py:136-140 + build_psf's normalisation and smart guess (device-side data policies)
 *   lcb_phot_prepare_batch star_photometry.py:47-64, 309-316 (scale, flux guess, NaN and mask rules)
 *
 * Conventions: all arrays are C-contiguous float32, row-major [item][y][x]; positions are in data
 * pixels with the origin at the stamp centre (n-1)/2 (roi_modelling.py:207-210).  `mem` selects
 * where EVERY pointer of a call lives: LCB_MEM_DEVICE (device pointers, work is enqueued on
 * `stream`; the only host synchronisations are the ones documented per entry point: lcb_psf_fit_batch reads star_off back
 * to size its launches) or LCB_MEM_HOST (host pointers; the library stages through a device arena leased for the call,
 * H2D and D2H copies included, and returns after synchronising).  Calls may come from several host threads (one per
 * GPU): conventions and the last-error string are per thread, staging arenas are leased per call and device.
 * The caller owns every buffer; nothing is retained after return except explicit handles.
 * Return value: 0 on success, negative lcb_status otherwise; lcb_last_error() gives the text.
 * There is no CPU fallback: without a CUDA device every compute entry returns LCB_ERR_CUDA.
 */
#ifndef LCB_H
#define LCB_H

#ifdef __cplusplus
extern "C" {
#endif

enum lcb_status {
    LCB_OK = 0,
    LCB_ERR_ARG = -1,     /* bad argument / unsupported shape */
    LCB_ERR_CUDA = -2,    /* CUDA runtime error (text in lcb_last_error) */
    LCB_ERR_NOMEM = -3
};

enum lcb_mem { LCB_MEM_DEVICE = 0, LCB_MEM_HOST = 1 };

/* per-item status flags written to status[] */
enum lcb_item_status { LCB_ITEM_OK = 0, LCB_ITEM_NONFINITE = 1 };

/* every recalled STARRED constant, switchable at run time, PER HOST THREAD (lcb_conventions_set
 * touches only the calling thread; a deconvolution handle keeps the conventions of the thread that created it). */
typedef struct {
    float gauss_fwhm_up;     /* FWHM of the target-resolution Gaussian, upsampled px (2.0) */
    int   gauss_taps;        /* G, even, one of 8/12/16 (12) */
    int   downsample_mean;   /* 0 (default): D_k is the block sum, amplitudes are pixel-sum fluxes; 1: block mean */
    int   chi2_half;         /* 1: chi2 term is 1/2 sum w r^2 */
    float clip_global_norm;  /* optax.clip_by_global_norm when scheduled (1.0) */
    float lr_decay_rate;     /* exponential_decay rate over max_iterations (0.99) */
    float belief_b1, belief_b2, belief_eps, belief_eps_root;
} lcb_conventions;

int lcb_conventions_get(lcb_conventions* out);
int lcb_conventions_set(const lcb_conventions* in);
const char* lcb_last_error(void);
int lcb_version(void);
/* number of visible CUDA devices (0 when none); never fails */
int lcb_device_count(void);

/* Optimiser options (starred Optimizer.minimize(**opts), star_photometry.py:115-120) */
typedef struct {
    int   n_iter;     /* max_iterations; loss history has exactly n_iter entries */
    float lr;         /* init_learning_rate */
    int   schedule;   /* schedule_learning_rate: 1 = clip_by_global_norm + exponential decay */
} lcb_fit_opts;

/* ---------------- K2: fixed-PSF amplitude + shift photometry ---------------------------------- */
typedef struct {
    int B;                  /* number of (frame,star) items */
    int n, k;               /* stamp side (data px), subsampling factor; PSF side is n*k */
    const float* data;      /* [B][n][n] */
    const float* weight;    /* [B][n][n]  1/sigma^2 (0 = ignored pixel) */
    const float* psf;       /* [Fp][n*k][n*k] narrow PSFs */
    const int*   psf_index; /* [B] index into psf */
    int Fp;                 /* number of PSFs */
    const float* a0;        /* [B] initial amplitude */
    const float* dx0;       /* [B] initial shifts, may be NULL (= 0) */
    const float* dy0;
} lcb_phot_batch;

typedef struct {
    float* a; float* dx; float* dy;   /* [B] fitted parameters */
    float* sigma_a;                   /* [B] Fisher sigma of a (may be NULL) */
    float* chi2;                      /* [B] sum w r^2 / n^2 at the final parameters (may be NULL) */
    float* residuals;                 /* [B][n][n] data - model (may be NULL) */
    float* loss_hist;                 /* [B][n_iter] (may be NULL) */
    float* loss0;                     /* [B] loss at the initial parameters (may be NULL) */
    float* grad0;                     /* [B][3] d loss / d (a,dx,dy) at the initial parameters (may be NULL) */
    int*   status;                    /* [B] lcb_item_status (may be NULL) */
} lcb_phot_out;

int lcb_phot_fit_batch(const lcb_phot_batch* in, const lcb_fit_opts* opt, lcb_phot_out* out,
                       int mem, void* stream);

/* starred.psf.psf.apply_distortion (star_photometry.py:303, roi_file_preparation.py:179): the narrow PSF of frame psf_index[i]
 * seen at the rescaled frame position xy[i] under that frame's distortion coefficients.  psf [Fp][nu][nu]; theta [Fp][6];
 * psf_index [B]; xy [B][2]; out [B][nu][nu]; mode as lcb_psf_opts.field_distortion (1 or 2). */
int lcb_apply_distortion_batch(const float* psf, const float* theta, const int* psf_index, const float* xy, int B, int Fp,
                               int nu, int mode, float* out, int mem, void* stream);

/* ---------------- reductions on the fitted fluxes (DEVICE pointers only, enqueued on `stream`) ---------------------------
 * flux, dflux: [F][S] frame-major (item f*S + s, the order of lcb_phot_fit_batch's outputs for F frames x S stars); NaN = no
 * measurement.  S <= 64 for the scatter matrix and the zero points.
 *
 * lightcurver/processes/normalization_calculation.py:157-206 (calculate_coefficient):
 *   lcb_norm_medians         per-star median flux over the frames (:158)                                        -> median [S]
 *   lcb_norm_scatter_matrix  Q [S][S] (double) with cost_function_scatter_in_frame(c) = c^T Q c (:75-98) for the fluxes divided
 *                            by the star medians; work: lcb_norm_scatter_work_doubles(F, S) doubles
 *   lcb_norm_coefficients    per-frame coefficient = weighted mean over the stars of star_scale * flux / median, weights
 *                            1 / (scaled uncertainty)^2, and its weighted standard deviation (0 -> 10 % of the coefficient) (:185-204)
 * lightcurver/processes/absolute_zeropoint_calculation.py:95-100:
 *   lcb_zeropoints           per-frame median and standard deviation (ddof 1) of catalog_mag[s] + 2.5 log10(flux[f][s]) */
int lcb_norm_medians(const float* flux, int F, int S, float* median, void* stream);
int lcb_norm_scatter_work_doubles(int F, int S);
int lcb_norm_scatter_matrix(const float* flux, const float* dflux, const float* median, int F, int S, double* Q, double* work,
                            void* stream);
int lcb_norm_coefficients(const float* flux, const float* dflux, const float* median, const float* star_scale, int F, int S,
                          float* coefficient, float* coefficient_uncertainty, void* stream);
int lcb_zeropoints(const float* flux, const float* catalog_mag, int F, int S, float* zeropoint, float* zeropoint_uncertainty,
                   void* stream);

/* ---------------- K1: per-frame PSF fit (starred build_psf) ------------------------------------ */
/* Ragged batch: frame f owns stars star_off[f] .. star_off[f+1]-1 of every per-star array. */
typedef struct {
    int F;                   /* frames */
    const int* star_off;     /* [F+1] */
    int n, k;                /* stamp side, subsampling factor (nu = n*k) */
    const float* data;       /* [sumN][n][n] stamps, already normalised by the caller */
    const float* weight;     /* [sumN][n][n] mask / sigma^2 */
    const float* W;          /* [F][J][nu*nu] starlet-space weights supplied by the caller, or NULL */
    const float* stamp_xy;   /* [sumN][2] rescaled frame coordinates (x, y) of the stamps (utilities/image_coordinates.py:4-25);
                                needed when lcb_psf_opts.field_distortion != 0, else may be NULL */
} lcb_psf_batch;

typedef struct {
    int   n_iter_analytic;   /* stage 1 iterations (reference: L-BFGS-B maxiter; here LM, early stop) ; 0 = skip */
    int   n_iter_adabelief;  /* stage 2 iterations */
    float lr;                /* stage 2 init_learning_rate (scheduled, clipped) */
    float lam_scales, lam_hf;/* regularization_strength_scales / _hf */
    int   noise_weights;     /* 0: W from batch (NULL -> 1);  1: SLIT (diagonal, deterministic) propagation of the weights through the
                                stage-1 model;  2: Monte-Carlo propagation (propagate_noise(method='MC')): mc_samples noise draws */
    float fwhm_min, fwhm_max, beta_min, beta_max;   /* bounds of the analytic stage */
    int   mc_samples;        /* noise_weights == 2: number of noise realisations (<= 0: 100) */
    unsigned mc_seed;        /* noise_weights == 2: seed of the counter-based generator */
    int   field_distortion;  /* build_psf(field_distortion=...), psf_modelling.py:169: 0 off; 1: every star sees the affine
                                resampling of the frame's narrow PSF given by kwargs_distortion at its frame position (flux
                                conserving); 2: the same without the determinant factor.  Stage 2 fits the 6 coefficients. */
} lcb_psf_opts;

typedef struct {
    float* moffat;           /* [F][5] fwhm_x, fwhm_y, phi, beta, C   in: guess, out: fitted */
    float* a; float* x0; float* y0;   /* [sumN]                       in: guess, out: fitted */
    float* background;       /* [F][nu*nu]                            in: initial grid, out: fitted */
    float* narrow_psf;       /* [F][nu*nu] (may be NULL) */
    float* full_psf;         /* [F][nu*nu] (may be NULL) */
    float* residuals;        /* [sumN][n][n] data - model (may be NULL) */
    float* chi2;             /* [F] sum w r^2 / #(w>0) (may be NULL) */
    float* loss_hist;        /* [F][n_iter_adabelief] (may be NULL) */
    float* loss_hist_analytic; /* [F][n_iter_analytic] (may be NULL) */
    float* W_out;            /* [F][J][nu*nu] the weights used (may be NULL) */
    float* loss0;            /* [F] stage-2 loss at its initial point (may be NULL) */
    float* grad_b0;          /* [F][nu*nu] d loss / d background at the initial point (may be NULL) */
    float* grad_s0;          /* [sumN][3] d loss / d (a, x0, y0) at the initial point (may be NULL) */
    int*   status;           /* [F] (may be NULL) */
    float* distortion;       /* [F][6] kwargs_distortion: dilation_x (2), dilation_y (2), shear (2), each linear in the rescaled
                                frame position (X, Y); in: initial, out: fitted.  Mandatory when field_distortion != 0 */
    float* grad_dist0;       /* [F][6] d loss / d distortion at the initial point (may be NULL) */
} lcb_psf_out;

/* number of starlet scales used for a grid of side nu: int(log2(nu)) */
int lcb_starlet_scales(int nu);
int lcb_psf_fit_batch(const lcb_psf_batch* in, const lcb_psf_opts* opt, lcb_psf_out* out,
                      int mem, void* stream);

/* ---------------- K3: joint multi-epoch deconvolution (roi_modelling.py:213-335) ----------------- */
typedef struct {
    int E, n, k;            /* epochs (local to this rank), stamp side, subsampling factor */
    int P;                  /* narrow PSF side (upsampled px), any parity */
    int M;                  /* point sources (<= 8) */
    const float* data;      /* [E][n][n]  (already scaled, roi_modelling.py:162-164) */
    const float* weight;    /* [E][n][n]  1/sigma^2 */
    const float* psf;       /* [E][P][P] */
} lcb_deconv_problem;

/* kwargs of the STARRED model (roi_modelling.py:221-263): a is epoch-major a[e*M+m] (:462) */
typedef struct {
    const float* h;         /* [nu*nu] shared background (NULL keeps the current one; initial 0) */
    const float* mean;      /* [E] */
    const float* a;         /* [E*M] */
    const float* c_x; const float* c_y;   /* [M] */
    const float* dx; const float* dy;     /* [E] */
    const float* alpha;     /* [E] fixed rotation of each epoch (radians) */
    int free_h, free_mean, free_a, free_c, free_d;   /* which groups the optimiser moves */
} lcb_deconv_params;

typedef struct {
    float lam_scales, lam_hf, lam_pos;    /* regularization_strength_{scales,hf,positivity} */
    const float* W;                       /* [J][nu*nu] or NULL (== 1) */
    const float* prior_mu_x; const float* prior_sig_x;   /* [M] Gaussian prior on c_x (Prior(prior_analytic=...)), or NULL */
    const float* prior_mu_y; const float* prior_sig_y;
    float lam_pts;          /* regularization_strength_pts_source (roi_modelling.py:311): L1 of the first starlet scale of the
                               point-source channel, weighted by W[0] */
    float lam_fu;           /* regularization_strength_flux_uniformity (roi_modelling.py:275-276, 312): scatter of a over the epochs */
    int pts_all_epochs;     /* 1: the pts-source term is summed over all epochs, 0: first epoch only */
    int fu_relative;        /* 1: sum_m std_e(a_em)/|mean_e(a_em)|, 0: sum_m std_e(a_em) */
} lcb_deconv_reg;

typedef struct {           /* gradient of the loss at the current parameters */
    float* loss;            /* [1] */
    float* h;               /* [nu*nu] */
    float* mean; float* a; float* c_x; float* c_y; float* dx; float* dy;
} lcb_deconv_grad;

int lcb_deconv_create(const lcb_deconv_problem* p, int mem, void* stream, void** handle);
int lcb_deconv_set_params(void* handle, const lcb_deconv_params* q, int mem);   /* also restarts the optimiser state */
int lcb_deconv_set_reg(void* handle, const lcb_deconv_reg* r, int mem);
/* CTAs per epoch of the per-epoch kernel (thread-block cluster size): 0 = automatic, or 1 .. 8 (any size: bands of ceil(n / size) rows) */
int lcb_deconv_set_cluster(void* handle, int ctas_per_epoch);
int lcb_deconv_get_cluster(void* handle);
/* Epoch sharding: total number of epochs over all ranks, global index of this rank's first epoch, and
 * (may be NULL) one value per source near its mean flux, identical on all ranks (shift of the flux sums). */
int lcb_deconv_set_global(void* handle, int E_total, int e0, const float* flux_shift, int mem);
/* In-kernel all-reduce over NVLink peer memory (no NCCL launch per iteration): comm_init allocates this
 * rank's receive buffer and writes its 64-byte CUDA IPC handle to ipc_handle_out; the caller exchanges the
 * handles (any transport) and passes all of them, in rank order, to comm_connect.  world <= 8, one node.
 * Afterwards lcb_deconv_run and lcb_deconv_loss_grad are COLLECTIVE calls (same arguments on every rank). */
int lcb_deconv_comm_init(void* handle, int rank, int world, void* ipc_handle_out);
int lcb_deconv_comm_connect(void* handle, const void* all_handles);
/* n_iter AdaBelief iterations, enqueued without host synchronisation; loss_hist [n_iter] may be NULL.
 * Single rank, or every rank of a connected communicator. */
int lcb_deconv_run(void* handle, const lcb_fit_opts* opt, float* loss_hist, int mem);
/* The same run on `count` independent single-rank handles at once: the iterations of the handles are interleaved, each handle on
 * its own stream, so that many small joint fits (the reference-coupled star photometry of star_photometry.py:74-87, one joint fit
 * per star) fill the GPU together instead of one after the other.  loss_hist: `count` pointers ([n_iter] each) or NULL. */
int lcb_deconv_run_many(void* const* handles, int count, const lcb_fit_opts* opt, float* const* loss_hist, int mem);
/* Alternative multi-rank driver with an external collective (e.g. NCCL): one iteration = step_local ;
 * all-reduce(sum) of reduce_buffer ; step_update.
 * After the last iteration call lcb_deconv_flush to apply the pending per-epoch update. */
int lcb_deconv_step_local(void* handle, int want_model);
int lcb_deconv_reduce_buffer(void* handle, float** device_ptr, int* count);
int lcb_deconv_step_update(void* handle, int it, int n_iter, float lr, int schedule);
/* applies the pending per-epoch update of the last step_update and clears the pending flag (end of an external loop) */
int lcb_deconv_flush(void* handle);
int lcb_deconv_loss_grad(void* handle, lcb_deconv_grad* g, int mem);
/* Stage 1 of do_modelling_of_roi (roi_modelling.py:260-281: Optimizer('l-bfgs-b').minimize over {dx, dy, a}) with the optimiser
 * state resident on the device: projected L-BFGS (10 pairs) + Armijo backtracking, bounds a >= a_lower and |dx|, |dy| <= n/2,
 * scipy's stopping rules (relative decrease <= ftol, projected gradient <= pgtol, maxiter).  No host round trip per evaluation:
 * the host enqueues 16 evaluate-and-step rounds at a time and reads one flag.  Single-rank handles.
 * loss_hist [maxiter] (may be NULL): loss after every accepted iteration; info [5] (may be NULL): iterations, evaluations, stop
 * reason (1 ftol, 2 pgtol, 3 maxiter, 4 line search stalled, 0 budget), final loss, |projected gradient|_inf. */
int lcb_deconv_lbfgs(void* handle, int maxiter, float a_lower, float ftol, float pgtol, float* loss_hist, float* info, int mem);
/* current parameters; model [E][n][n] and loss [1] are evaluated when non-NULL */
int lcb_deconv_get(void* handle, lcb_deconv_params* q, float* model, float* loss, int mem);
/* starlet-space noise weights of h: stage 0 fills the reduce buffer with the local variance plane
 * (all-reduce it when epochs are sharded), stage 1 builds W [J][nu*nu], installs it for the
 * regulariser and copies it to W_out when non-NULL */
int lcb_deconv_noise_weights(void* handle, int stage, float* W_out, int mem);
int lcb_deconv_destroy(void* handle);

/* ---------------- host-side data policies of the batched drivers, on the device ------------------
 * All pointers are DEVICE pointers; work is enqueued on `stream` (no synchronisation). */
typedef struct {           /* psf_modelling.py:136-140 + build_psf's normalisation / smart guess */
    int F; const int* star_off;      /* ragged batch, as lcb_psf_batch */
    int n, k;
    const float* image;              /* [sumN][n][n] raw stamps (NaN allowed) */
    const float* noisemap;           /* [sumN][n][n] */
    const unsigned char* mask;       /* [sumN][n][n] nonzero = good pixel, or NULL */
    float norm_scale;                /* stamps are divided by max(frame) / norm_scale (100) */
    int downsample_mean;             /* conventions: a0 = flux * k^2 when D_k is the block mean */
    int guess_method;                /* guess_method_star_position: 0 'center', 1 'max', 2 'barycenter' */
} lcb_psf_prepare_in;

typedef struct {
    float* data; float* weight;      /* [sumN][n][n] normalised stamps, mask / sigma^2 */
    float* a0; float* x0; float* y0; /* [sumN] initial amplitudes and positions */
    float* norm;                     /* [F] normalisation of each frame (may be NULL) */
} lcb_psf_prepare_out;

int lcb_psf_prepare_batch(const lcb_psf_prepare_in* in, lcb_psf_prepare_out* out, void* stream);

typedef struct {           /* star_photometry.py:47-64, 309-316 for every star of a footprint at once */
    int F, S, n, k;                  /* frames, stars, stamp side, subsampling factor */
    const float* data;               /* [F][S][n][n] raw stamps */
    const float* noisemap;           /* [F][S][n][n] */
    const unsigned char* mask;       /* [F][S][n][n] nonzero = good pixel, or NULL */
    int downsample_mean;
} lcb_phot_prepare_in;

typedef struct {
    float* data; float* weight;      /* [F][S][n][n] stamps / scale[s], 1 / sigma^2 */
    float* a0;                       /* [F][S] initial flux guess */
    float* scale;                    /* [S] nanmax of each star over all its epochs */
} lcb_phot_prepare_out;

size_t lcb_phot_prepare_work_floats(int F, int S);
int lcb_phot_prepare_batch(const lcb_phot_prepare_in* in, lcb_phot_prepare_out* out, float* work, void* stream);

/* Pitched copy host <-> device (cudaMemcpy2DAsync on `stream`): `height` rows of `width_bytes`, row pitches in bytes.  Used by the
 * host layer to upload a strided slice of a pinned batch array (one device's share of the stars) without a host-side gather. */
int lcb_copy_2d(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width_bytes, size_t height, int to_device, void* stream);

/* ---------------- measurement helper ---------------------------------------------------------- */
/* FP32 FMA micro-benchmark: runs `iters` dependent-chain FFMA loops on every SM and returns the
 * achieved TFLOP/s in *tflops (used as the measured roofline denominator by bench.py). */
int lcb_fp32_peak(int iters, float* tflops, float* ms);
/* same with three-register FFMAs in an 8x8 outer-product pattern (what a stencil inner loop issues) */
int lcb_fp32_peak_rrr(int iters, float* tflops, float* ms);
/* packed FP32 (FFMA2, sm_100): which = 0 FFMA2 chains; 1 / 2 = scalar / packed FMAs interleaved with one shared-memory load per
 * two FMAs (how many issue slots the packed form leaves for the loads of a stencil loop).  2 flops per scalar FMA. */
int lcb_fp32x2_peak(int iters, int which, float* tflops, float* ms);

/* Per-kernel device timing: after lcb_profile_enable(1) every kernel launched by the library is
 * bracketed by CUDA events on its launch stream; lcb_profile_summary() synchronises on them and
 * writes a JSON object {"kernel": {"ms": total, "launches": count}, ...} into buf. */
int lcb_profile_enable(int on);
int lcb_profile_summary(char* buf, int buflen);

#ifdef __cplusplus
}
#endif
#endif /* LCB_H */
