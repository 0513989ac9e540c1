// lcb_psf_setup.cu -- the setup stages of starred build_psf (psf_modelling.py:164-171):
//   K1a  k_psf_moffat_lm : stage 1, analytic fit of {fwhm_x, fwhm_y, phi, beta, a_i, x0_i, y0_i}
//        with background 0 and lambda 0.  The reference runs scipy L-BFGS-B
//        (n_iter_analytic iterations); this stage is a pure weighted least-squares problem, so the
//        kernel runs Levenberg-Marquardt on the arrow-head normal equations (4 shared Moffat
//        parameters + 3 per star), one CTA per frame.  Parity is on the converged loss/parameters,
//        not the trajectory (hard part 3).  C is held at 1.
//   K5   k_noise_weights : W_j = sqrt( var_up (*) psi_j^2 ), the SLIT-style propagation of the noise
//        into starlet space (star_photometry.py:108, roi_modelling.py:299), separable.
#include "lcb_psf.cuh"

template <int NV>
__device__ __forceinline__ void block_reduce_s(float (&v)[NV], float* red, int tid) {
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = warp_sum(v[i]);
    __syncthreads();                                  // previous readers of red are done
    if ((tid & 31) == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) red[(tid >> 5) * NV + i] = v[i];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        float s = 0.f;
        for (int w = 0; w < PSF_WARPS; ++w) s += red[w * NV + i];
        v[i] = s;
    }
}

// Raw elliptical Moffat (1+q)^-beta and its derivatives wrt (fwhm_x, fwhm_y, phi, beta)  (A.1)
__device__ __forceinline__ void moffat_raw(const float* th, int k, float x, float y, float& m, float (&dm)[4]) {
    const float fx = th[0], fy = th[1], ph = th[2], be = th[3];
    float sp_, cp_;
    sincosf(ph, &sp_, &cp_);
    const float xr = x * cp_ + y * sp_, yr = -x * sp_ + y * cp_;
    const float p2 = exp2f(1.f / be);
    const float fac = 2.f * sqrtf(p2 - 1.f);
    const float rx = fx * (float)k / fac, ry = fy * (float)k / fac;
    const float qx = (xr / rx) * (xr / rx), qy = (yr / ry) * (yr / ry);
    const float q = qx + qy;
    const float l1q = log1pf(q);
    m = expf(-be * l1q);
    const float dmdq = -be * m / (1.f + q);
    dm[0] = dmdq * (-2.f * qx / fx);
    dm[1] = dmdq * (-2.f * qy / fy);
    dm[2] = dmdq * (2.f * xr * yr * (1.f / (rx * rx) - 1.f / (ry * ry)));
    // d fac / d beta = (p2-1)^-1/2 * p2 * ln2 * (-1/beta^2);  dq/dbeta = 2 q (dfac/dbeta)/fac
    const float dfac = -p2 * 0.6931471805599453f / (be * be * sqrtf(p2 - 1.f));
    dm[3] = -l1q * m + dmdq * (2.f * q * dfac / fac);
}

__global__ void __launch_bounds__(PSF_THREADS) k_moffat_image(PsfArgs A) {
    __shared__ float red[PSF_WARPS];
    const int f = blockIdx.x, nu = A.nu, pp = nu * nu, tid = threadIdx.x;
    const float* th = A.moffat + f * 5;
    const float ctr = 0.5f * (float)(nu - 1);
    float* out = A.s_fixed + (size_t)f * pp;
    float tot[1] = {0.f};
    for (int i = tid; i < pp; i += PSF_THREADS) {
        float m, dm[4];
        moffat_raw(th, A.k, (float)(i % nu) - ctr, (float)(i / nu) - ctr, m, dm);
        out[i] = m;
        tot[0] += m;
    }
    block_reduce_s<1>(tot, red, tid);
    const float sc = th[4] / tot[0];
    for (int i = tid; i < pp; i += PSF_THREADS) out[i] *= sc;
}

// ---------------------------------------------------------------- K1a Levenberg-Marquardt
#define LM_GRAM 36   // 28 upper-triangle entries of the 7x7 block + 7 gradient entries + chi2

__device__ __forceinline__ int tri_idx(int r, int c) { return r * 7 - r * (r - 1) / 2 + (c - r); }  // r <= c

__device__ __forceinline__ void inv3(const double (&A)[3][3], double (&inv)[3][3]) {
    const double c00 = A[1][1] * A[2][2] - A[1][2] * A[2][1];
    const double c01 = A[1][2] * A[2][0] - A[1][0] * A[2][2];
    const double c02 = A[1][0] * A[2][1] - A[1][1] * A[2][0];
    double det = A[0][0] * c00 + A[0][1] * c01 + A[0][2] * c02;
    if (fabs(det) < 1e-300) det = (det < 0.0) ? -1e-300 : 1e-300;
    const double id = 1.0 / det;
    inv[0][0] = c00 * id;
    inv[0][1] = (A[0][2] * A[2][1] - A[0][1] * A[2][2]) * id;
    inv[0][2] = (A[0][1] * A[1][2] - A[0][2] * A[1][1]) * id;
    inv[1][0] = c01 * id;
    inv[1][1] = (A[0][0] * A[2][2] - A[0][2] * A[2][0]) * id;
    inv[1][2] = (A[0][2] * A[1][0] - A[0][0] * A[1][2]) * id;
    inv[2][0] = c02 * id;
    inv[2][1] = (A[0][1] * A[2][0] - A[0][0] * A[2][1]) * id;
    inv[2][2] = (A[0][0] * A[1][1] - A[0][1] * A[1][0]) * id;
}

// Solves the damped arrow-head system in double; writes the step into dth[4], dsp[N][3].
__device__ void lm_solve(const float* gram, int N, double lam, double* dth, float* dsp) {
    double H[4][4], rhs[4];
    for (int r = 0; r < 4; ++r) { rhs[r] = 0.0; for (int c = 0; c < 4; ++c) H[r][c] = 0.0; }
    for (int st = 0; st < N; ++st) {
        const float* g = gram + st * LM_GRAM;
        double Ass[3][3], Ats[4][3], gs[3];
        for (int r = 0; r < 4; ++r) {
            for (int c = r; c < 4; ++c) { H[r][c] += g[tri_idx(r, c)]; if (c != r) H[c][r] += g[tri_idx(r, c)]; }
            rhs[r] -= g[28 + r];
            for (int c = 0; c < 3; ++c) Ats[r][c] = g[tri_idx(r, 4 + c)];
        }
        for (int r = 0; r < 3; ++r) {
            for (int c = r; c < 3; ++c) { Ass[r][c] = g[tri_idx(4 + r, 4 + c)]; Ass[c][r] = Ass[r][c]; }
            gs[r] = g[28 + 4 + r];
        }
        for (int r = 0; r < 3; ++r) Ass[r][r] = Ass[r][r] * (1.0 + lam) + 1e-30;
        double inv[3][3];
        inv3(Ass, inv);
        // Schur complement
        double AtsInv[4][3];
        for (int r = 0; r < 4; ++r)
            for (int c = 0; c < 3; ++c) AtsInv[r][c] = Ats[r][0] * inv[0][c] + Ats[r][1] * inv[1][c] + Ats[r][2] * inv[2][c];
        for (int r = 0; r < 4; ++r) {
            for (int c = 0; c < 4; ++c) H[r][c] -= AtsInv[r][0] * Ats[c][0] + AtsInv[r][1] * Ats[c][1] + AtsInv[r][2] * Ats[c][2];
            rhs[r] += AtsInv[r][0] * gs[0] + AtsInv[r][1] * gs[1] + AtsInv[r][2] * gs[2];
        }
    }
    // damp the shared block with the un-Schur'd diagonal
    for (int r = 0; r < 4; ++r) {
        double d = 0.0;
        for (int st = 0; st < N; ++st) d += gram[st * LM_GRAM + tri_idx(r, r)];
        H[r][r] += lam * d + 1e-30;
    }
    // 4x4 Gaussian elimination with partial pivoting
    int perm[4] = {0, 1, 2, 3};
    for (int c = 0; c < 4; ++c) {
        int piv = c;
        for (int r = c + 1; r < 4; ++r) if (fabs(H[perm[r]][c]) > fabs(H[perm[piv]][c])) piv = r;
        const int tmp = perm[c]; perm[c] = perm[piv]; perm[piv] = tmp;
        const double d = H[perm[c]][c];
        if (fabs(d) < 1e-300) continue;
        for (int r = c + 1; r < 4; ++r) {
            const double m = H[perm[r]][c] / d;
            for (int cc = c; cc < 4; ++cc) H[perm[r]][cc] -= m * H[perm[c]][cc];
            rhs[perm[r]] -= m * rhs[perm[c]];
        }
    }
    for (int c = 3; c >= 0; --c) {
        double s = rhs[perm[c]];
        for (int cc = c + 1; cc < 4; ++cc) s -= H[perm[c]][cc] * dth[cc];
        const double d = H[perm[c]][c];
        dth[c] = (fabs(d) < 1e-300) ? 0.0 : s / d;
    }
    // back-substitution per star
    for (int st = 0; st < N; ++st) {
        const float* g = gram + st * LM_GRAM;
        double Ass[3][3], r3[3];
        for (int r = 0; r < 3; ++r) {
            for (int c = r; c < 3; ++c) { Ass[r][c] = g[tri_idx(4 + r, 4 + c)]; Ass[c][r] = Ass[r][c]; }
            r3[r] = -g[28 + 4 + r];
            for (int c = 0; c < 4; ++c) r3[r] -= g[tri_idx(c, 4 + r)] * dth[c];
        }
        for (int r = 0; r < 3; ++r) Ass[r][r] = Ass[r][r] * (1.0 + lam) + 1e-30;
        double inv[3][3];
        inv3(Ass, inv);
        for (int r = 0; r < 3; ++r)
            dsp[st * 3 + r] = (float)(inv[r][0] * r3[0] + inv[r][1] * r3[1] + inv[r][2] * r3[2]);
    }
}

template <int K, int G>
__global__ void __launch_bounds__(PSF_THREADS) k_psf_moffat_lm(PsfArgs A) {
    using P = LcbPass<K, G>;
    extern __shared__ __align__(16) float sm[];
    const int n = A.n, nu = A.nu, nn = n * n, pp = nu * nu, tid = threadIdx.x;
    const int ldv = n + 1, ldt = n + 1;
    const int f = blockIdx.x;
    const int i0 = A.star_off[f], N = A.star_off[f + 1] - i0;
    const DevConv cv = A.cv;
    const float fk = (float)K;
    const float ctr = 0.5f * (float)(nu - 1);

    float* taps = sm;                                   // [Nmax][4][GE_MAX]
    float* cur = taps + A.Nmax * 4 * LCB_GE_MAX;        // [4 + 3 Nmax] accepted parameters
    float* tri = cur + 4 + 3 * A.Nmax;                  // [4 + 3 Nmax] trial parameters
    float* gram = tri + 4 + 3 * A.Nmax;                 // [2][Nmax][LM_GRAM]
    float* red = gram + 2 * A.Nmax * LM_GRAM;           // [PSF_WARPS][LM_GRAM]
    float* ctl = red + PSF_WARPS * LM_GRAM;             // [8] control scalars broadcast by thread 0
    float* Vg = ctl + 8;                                // [nu][ldv]
    float* Vd = Vg + nu * ldv;
    float* Jim = Vd + nu * ldv;                         // [7][n][ldt] Jacobian images, [X][Y]
    float* planes = Jim + (A.jim_in_smem ? 8 * n * ldt : 0);   // [5][pp]: s, ds/dtheta_c
    if (!A.planes_in_smem) planes = A.work + (size_t)f * A.work_per_frame;
    if (!A.jim_in_smem) Jim = A.work + (size_t)f * A.work_per_frame + (size_t)5 * pp;   // large stamps: L2
    float* dfT = Jim + 7 * n * ldt;                     // [n][ldt] a*m0 - d

    const float* dat = A.data + (size_t)i0 * nn;
    const float* wgt = A.weight + (size_t)i0 * nn;

    if (tid < 4) cur[tid] = A.moffat[f * 5 + tid];
    for (int i = tid; i < N; i += PSF_THREADS) {
        cur[4 + 3 * i] = A.a[i0 + i]; cur[5 + 3 * i] = A.x0[i0 + i]; cur[6 + 3 * i] = A.y0[i0 + i];
    }
    __syncthreads();

    // evaluate Jacobian Gram + loss at parameters `par` into gram buffer `gb`; returns loss
    auto evaluate = [&](const float* par, float* gb) -> float {
        // planes
        float sums[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
        for (int i = tid; i < pp; i += PSF_THREADS) {
            float m, dm[4];
            moffat_raw(par, K, (float)(i % nu) - ctr, (float)(i / nu) - ctr, m, dm);
            planes[i] = m; sums[0] += m;
#pragma unroll
            for (int c = 0; c < 4; ++c) { planes[(c + 1) * pp + i] = dm[c]; sums[c + 1] += dm[c]; }
        }
        block_reduce_s<5>(sums, red, tid);
        const float inv = 1.f / sums[0];
        for (int i = tid; i < pp; i += PSF_THREADS) {
            const float s = planes[i] * inv;
            planes[i] = s;
#pragma unroll
            for (int c = 0; c < 4; ++c) planes[(c + 1) * pp + i] = (planes[(c + 1) * pp + i] - s * sums[c + 1]) * inv;
        }
        for (int idx = tid; idx < N * 2 * P::GE; idx += PSF_THREADS) {
            const int st = idx / (2 * P::GE), rem = idx % (2 * P::GE), which = rem / P::GE, p = rem % P::GE;
            const float c = fk * par[4 + 3 * st + (which ? 1 : 2)];
            const float ic = floorf(c + 0.5f);
            float e, de;
            lcb_tap(cv, K, c - ic, p, e, de);
            taps[(st * 4 + (which ? 2 : 0)) * LCB_GE_MAX + p] = e;
            taps[(st * 4 + (which ? 3 : 1)) * LCB_GE_MAX + p] = de;
        }
        __syncthreads();
        float loss = 0.f;
        for (int st = 0; st < N; ++st) {
            const float a = par[4 + 3 * st], cx = fk * par[5 + 3 * st], cy = fk * par[6 + 3 * st];
            const int icx = (int)floorf(cx + 0.5f), icy = (int)floorf(cy + 0.5f);
            const float* tp = taps + st * 4 * LCB_GE_MAX;
            const float* ds = dat + (size_t)st * nn;
            for (int c = 0; c < 5; ++c) {
                lcb_pass1<K, G>(planes + (size_t)c * pp, nu, nu, n, icy, tp, tp + LCB_GE_MAX, Vg, Vd, ldv, tid, PSF_THREADS);
                __syncthreads();
                lcb_pass2<K, G>(Vg, Vd, ldv, nu, n, icx, tp + 2 * LCB_GE_MAX, tp + 3 * LCB_GE_MAX, nullptr, nullptr, 0, tid, PSF_THREADS,
                                [&](int Y, int X, float m0, float mx, float my, float, float) {
                                    const int o = X * ldt + Y;
                                    if (c == 0) {
                                        Jim[4 * n * ldt + o] = m0;
                                        Jim[5 * n * ldt + o] = a * fk * mx;
                                        Jim[6 * n * ldt + o] = a * fk * my;
                                        dfT[o] = fmaf(a, m0, -__ldg(ds + Y * n + X));
                                    } else {
                                        Jim[(c - 1) * n * ldt + o] = a * m0;
                                    }
                                });
                __syncthreads();
            }
            float acc[LM_GRAM];
#pragma unroll
            for (int q = 0; q < LM_GRAM; ++q) acc[q] = 0.f;
            const float* ws = wgt + (size_t)st * nn;
            for (int i = tid; i < nn; i += PSF_THREADS) {
                const int Y = i % n, X = i / n, o = X * ldt + Y;
                const float w = __ldg(ws + Y * n + X), df = dfT[o];
                float jv[7];
#pragma unroll
                for (int c = 0; c < 7; ++c) jv[c] = Jim[c * n * ldt + o];
#pragma unroll
                for (int r = 0; r < 7; ++r) {
                    const float wj = w * jv[r];
#pragma unroll
                    for (int c = r; c < 7; ++c) {
                        const int q = r * 7 - r * (r - 1) / 2 + (c - r);
                        acc[q] = fmaf(wj, jv[c], acc[q]);
                    }
                }
#pragma unroll
                for (int r = 0; r < 7; ++r) acc[28 + r] = fmaf(w * df, jv[r], acc[28 + r]);
                acc[35] = fmaf(w * df, df, acc[35]);
            }
            block_reduce_s<LM_GRAM>(acc, red, tid);
            if (tid < LM_GRAM) gb[st * LM_GRAM + tid] = acc[tid];
            loss += acc[35];
            __syncthreads();
        }
        return cv.half * loss;
    };

    int which = 0;
    float loss = evaluate(cur, gram);
    float lam = 1e-3f;
    int stall = 0, bad = 0;
    const int np = 4 + 3 * N;
    for (int it = 0; it < A.n_iter_lm; ++it) {
        if (tid == 0) {
            double dth[4];
            float* dsp = red;                                 // scratch (N*3 floats) owned by thread 0 here
            lm_solve(gram + which * A.Nmax * LM_GRAM, N, (double)lam, dth, dsp);
            const float lim = 0.25f * (float)n;
            tri[0] = fminf(fmaxf(cur[0] + (float)dth[0], A.fwhm_min), A.fwhm_max);
            tri[1] = fminf(fmaxf(cur[1] + (float)dth[1], A.fwhm_min), A.fwhm_max);
            tri[2] = cur[2] + (float)dth[2];
            tri[3] = fminf(fmaxf(cur[3] + (float)dth[3], A.beta_min), A.beta_max);
            for (int st = 0; st < N; ++st) {
                tri[4 + 3 * st] = fmaxf(cur[4 + 3 * st] + dsp[st * 3], 0.f);
                tri[5 + 3 * st] = fminf(fmaxf(cur[5 + 3 * st] + dsp[st * 3 + 1], -lim), lim);
                tri[6 + 3 * st] = fminf(fmaxf(cur[6 + 3 * st] + dsp[st * 3 + 2], -lim), lim);
            }
        }
        __syncthreads();
        const float lt = evaluate(tri, gram + (1 - which) * A.Nmax * LM_GRAM);
        bool stop = false;
        if (isfinite(lt) && lt < loss) {
            const float rel = (loss - lt) / fmaxf(loss, 1e-30f);
            for (int i = tid; i < np; i += PSF_THREADS) cur[i] = tri[i];
            which = 1 - which;
            loss = lt;
            lam = fmaxf(lam * 0.3f, 1e-9f);
            stall = (rel < 1e-7f) ? stall + 1 : 0;
            if (stall >= 2) stop = true;
        } else {
            if (!isfinite(lt)) bad = 1;
            lam *= 5.f;
            if (lam > 1e10f) stop = true;
        }
        if (tid == 0 && A.loss_hist_lm) A.loss_hist_lm[(size_t)f * A.n_iter_lm + it] = loss;
        __syncthreads();
        if (stop) {
            if (A.loss_hist_lm)
                for (int j = it + 1 + tid; j < A.n_iter_lm; j += PSF_THREADS) A.loss_hist_lm[(size_t)f * A.n_iter_lm + j] = loss;
            break;
        }
    }
    // ---- write back: parameters and the fixed analytic component s_fixed = C * Moffat(theta), C = 1
    __syncthreads();
    if (tid < 4) A.moffat[f * 5 + tid] = cur[tid];
    if (tid == 4) A.moffat[f * 5 + 4] = 1.f;
    for (int i = tid; i < N; i += PSF_THREADS) {
        A.a[i0 + i] = cur[4 + 3 * i]; A.x0[i0 + i] = cur[5 + 3 * i]; A.y0[i0 + i] = cur[6 + 3 * i];
    }
    if (tid == 0 && A.status && bad) A.status[f] = LCB_ITEM_NONFINITE;
    {
        float sums[1] = {0.f};
        for (int i = tid; i < pp; i += PSF_THREADS) {
            float m, dm[4];
            moffat_raw(cur, K, (float)(i % nu) - ctr, (float)(i / nu) - ctr, m, dm);
            planes[i] = m; sums[0] += m;
        }
        block_reduce_s<1>(sums, red, tid);
        const float inv = 1.f / sums[0];
        for (int i = tid; i < pp; i += PSF_THREADS) A.s_fixed[(size_t)f * pp + i] = planes[i] * inv;
    }
}

// without the Jacobian images (8 n (n+1) floats) and the 5 planes, which move to L2 when they do not fit
size_t lcb_psf_lm_smem_small(int n, int nu, int Nmax) {
    return (size_t)(Nmax * 4 * LCB_GE_MAX + 2 * (4 + 3 * Nmax) + 2 * Nmax * LM_GRAM + PSF_WARPS * LM_GRAM + 8 +
                    2 * nu * (n + 1)) * 4;
}

template <int K, int G>
static int launch_lm(const PsfArgs& A, size_t smem, cudaStream_t st) {
    LCB_CUDA(cudaFuncSetAttribute(k_psf_moffat_lm<K, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    { LcbProfScope ps("k_psf_moffat_lm", st); k_psf_moffat_lm<K, G><<<A.F, PSF_THREADS, smem, st>>>(A); }
    LCB_CUDA(cudaGetLastError());
    return LCB_OK;
}

int lcb_psf_lm_dispatch(const PsfArgs& A, size_t smem, cudaStream_t st) {
    const int G = A.cv.G;
#define CASE(KK, GG) if (A.k == KK && G == GG) return launch_lm<KK, GG>(A, smem, st);
    CASE(1, 12) CASE(2, 12) CASE(3, 12) CASE(4, 12)
    CASE(2, 8) CASE(2, 16)
#undef CASE
    lcb_set_error("psf moffat stage: unsupported (subsampling_factor=%d, gauss_taps=%d)", A.k, G);
    return LCB_ERR_ARG;
}

int lcb_moffat_image_launch(const PsfArgs& A, cudaStream_t st) {
    { LcbProfScope ps("k_moffat_image", st); k_moffat_image<<<A.F, PSF_THREADS, 0, st>>>(A); }
    LCB_CUDA(cudaGetLastError());
    return LCB_OK;
}

// ---------------------------------------------------------------- K5 noise weights
// Restates propagate_noise(model, noisemap, kwargs, ['starlet'], method='SLIT', likelihood_type='chi2')
// for the PSF grid: var_grad(p) = sum_i a_i^2 sum_q A_i[q,p]^2 w_i[q]  (diagonal of J^T C^-1 J, i.e. the
// transposed passes with SQUARED taps applied to the weights), then W_j = sqrt(var_grad (*) psi_j^2).
template <int K, int G>
__global__ void __launch_bounds__(PSF_THREADS) k_noise_var(PsfArgs A) {
    using P = LcbPass<K, G>;
    extern __shared__ __align__(16) float sm[];
    const int n = A.n, nu = A.nu, nn = n * n, pp = nu * nu, tid = threadIdx.x;
    const int ldt = n + 1, ldb = nu + 1;
    const int f = blockIdx.x;
    const int i0 = A.star_off[f], N = A.star_off[f + 1] - i0;
    const float fk = (float)K;
    float* taps2 = sm;                               // [Nmax][2][GE_MAX]  ey^2, ex^2
    float* wT = taps2 + A.Nmax * 2 * LCB_GE_MAX;     // [n][ldt]
    float* Vbar = wT + n * ldt;                      // [n][ldb]
    float* var = A.work + (size_t)f * A.work_per_frame;
    for (int idx = tid; idx < N * 2 * P::GE; idx += PSF_THREADS) {
        const int st = idx / (2 * P::GE), rem = idx % (2 * P::GE), which = rem / P::GE, p = rem % P::GE;
        const float c = fk * (which ? A.x0[i0 + st] : A.y0[i0 + st]);
        const float ic = floorf(c + 0.5f);
        float e, de;
        lcb_tap(A.cv, K, c - ic, p, e, de);
        taps2[(st * 2 + which) * LCB_GE_MAX + p] = e * e;
    }
    for (int i = tid; i < pp; i += PSF_THREADS) var[i] = 0.f;
    __syncthreads();
    for (int st = 0; st < N; ++st) {
        const float a = A.a[i0 + st];
        const float cx = fk * A.x0[i0 + st], cy = fk * A.y0[i0 + st];
        const int icx = (int)floorf(cx + 0.5f), icy = (int)floorf(cy + 0.5f);
        const float* ws = A.weight + (size_t)(i0 + st) * nn;
        for (int i = tid; i < nn; i += PSF_THREADS) wT[(i % n) * ldt + i / n] = ws[i];
        __syncthreads();
        lcb_pass2T<K, G>(wT, ldt, nu, n, icx, taps2 + (st * 2 + 1) * LCB_GE_MAX, Vbar, ldb, tid, PSF_THREADS);
        __syncthreads();
        lcb_pass1T<K, G>(Vbar, ldb, nu, n, icy, taps2 + (st * 2) * LCB_GE_MAX, tid, PSF_THREADS,
                         [&](int v, int u, float val) { var[v * nu + u] = fmaf(a * a, val, var[v * nu + u]); });
        __syncthreads();
    }
}

template <int K, int G>
static int launch_nvar(const PsfArgs& A, cudaStream_t st) {
    const size_t smem = (size_t)(A.Nmax * 2 * LCB_GE_MAX + A.n * (A.n + 1) + A.n * (A.nu + 1)) * 4;
    LCB_CUDA(cudaFuncSetAttribute(k_noise_var<K, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    { LcbProfScope ps("k_noise_var", st); k_noise_var<K, G><<<A.F, PSF_THREADS, smem, st>>>(A); }
    LCB_CUDA(cudaGetLastError());
    return LCB_OK;
}

int lcb_noise_var_dispatch(const PsfArgs& A, cudaStream_t st) {
    const int G = A.cv.G;
#define CASE(KK, GG) if (A.k == KK && G == GG) return launch_nvar<KK, GG>(A, st);
    CASE(1, 12) CASE(2, 12) CASE(3, 12) CASE(4, 12)
    CASE(2, 8) CASE(2, 16)
#undef CASE
    lcb_set_error("noise weights: unsupported (subsampling_factor=%d, gauss_taps=%d)", A.k, G);
    return LCB_ERR_ARG;
}

// ---------------------------------------------------------------- K5b: Monte-Carlo noise propagation
// propagate_noise(model, noisemap, kwargs, ['starlet'], method='MC', num_samples, seed, likelihood_type='chi2') [R]:
// for every sample draw z ~ N(0,1) per stamp pixel, push the chi2-gradient noise g = sum_i a_i A_i^T (sqrt(w_i) z_i) to
// the grid (the same transposed passes as the gradient of the fit), take its starlet transform and accumulate the
// squares; W_j = sqrt(mean).  Unlike the SLIT form this keeps the correlations of g between grid pixels.
// Counter-based generator: every (seed, star, pixel, sample) hashes to its own normal deviate (reproducible, order free).
__device__ __forceinline__ unsigned mc_mix(unsigned x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}
__device__ __forceinline__ float mc_normal(unsigned seed, unsigned item, unsigned sample) {
    const unsigned h1 = mc_mix(mc_mix(seed ^ 0x9e3779b9U) + item * 0x85ebca6bU + mc_mix(sample * 0xc2b2ae35U + 0x27d4eb2fU));
    const unsigned h2 = mc_mix(h1 ^ 0x165667b1U);
    const float u1 = ((float)(h1 >> 8) + 0.5f) * (1.f / 16777216.f);       // (0, 1)
    const float u2 = ((float)(h2 >> 8) + 0.5f) * (1.f / 16777216.f);
    return sqrtf(-2.f * logf(u1)) * cospif(2.f * u2);
}

__device__ __forceinline__ float mc_atrous(const float* __restrict__ c, int nu, int v, int u, int D, int axis) {
    const float h0 = 1.f / 16.f, h1 = 4.f / 16.f, h2 = 6.f / 16.f;
    if (axis == 0) {
        const float* row = c + v * nu;
        return h0 * (row[max(u - 2 * D, 0)] + row[min(u + 2 * D, nu - 1)]) + h1 * (row[max(u - D, 0)] + row[min(u + D, nu - 1)]) + h2 * row[u];
    }
    return h0 * (c[max(v - 2 * D, 0) * nu + u] + c[min(v + 2 * D, nu - 1) * nu + u]) +
           h1 * (c[max(v - D, 0) * nu + u] + c[min(v + D, nu - 1) * nu + u]) + h2 * c[v * nu + u];
}

template <int K, int G>
__global__ void __launch_bounds__(PSF_THREADS) k_noise_mc(PsfArgs A, float* W, int n_samples, unsigned seed, int frame0) {
    using P = LcbPass<K, G>;
    extern __shared__ __align__(16) float sm[];
    const int n = A.n, nu = A.nu, nn = n * n, pp = nu * nu, tid = threadIdx.x, J = A.J;
    const int ldt = n + 1, ldb = nu + 1;
    const int f = blockIdx.x;
    const int i0 = A.star_off[f], N = A.star_off[f + 1] - i0;
    const float fk = (float)K;
    float* taps = sm;                                // [Nmax][2][GE_MAX]  ey, ex
    float* wT = taps + A.Nmax * 2 * LCB_GE_MAX;      // [n][ldt]
    float* Vbar = wT + n * ldt;                      // [n][ldb]
    float* Gp = A.work + (size_t)f * A.work_per_frame;   // gradient-noise plane, then c_j
    float* C1 = Gp + pp;
    float* Wf = W + (size_t)f * J * pp;              // accumulators, then the weights
    (void)frame0;
    for (int idx = tid; idx < N * 2 * P::GE; idx += PSF_THREADS) {
        const int st = idx / (2 * P::GE), rem = idx % (2 * P::GE), which = rem / P::GE, p = rem % P::GE;
        const float c = fk * (which ? A.x0[i0 + st] : A.y0[i0 + st]);
        const float ic = floorf(c + 0.5f);
        float e, de;
        lcb_tap(A.cv, K, c - ic, p, e, de);
        taps[(st * 2 + which) * LCB_GE_MAX + p] = e;
    }
    for (int i = tid; i < J * pp; i += PSF_THREADS) Wf[i] = 0.f;
    __syncthreads();
    for (int smp = 0; smp < n_samples; ++smp) {
        for (int i = tid; i < pp; i += PSF_THREADS) Gp[i] = 0.f;
        __syncthreads();
        for (int st = 0; st < N; ++st) {
            const float a = A.a[i0 + st];
            const float cx = fk * A.x0[i0 + st], cy = fk * A.y0[i0 + st];
            const int icx = (int)floorf(cx + 0.5f), icy = (int)floorf(cy + 0.5f);
            const float* ws = A.weight + (size_t)(i0 + st) * nn;
            for (int i = tid; i < nn; i += PSF_THREADS)
                wT[(i % n) * ldt + i / n] = sqrtf(ws[i]) * mc_normal(seed, (unsigned)((i0 + st) * nn + i), (unsigned)smp);
            __syncthreads();
            lcb_pass2T<K, G>(wT, ldt, nu, n, icx, taps + (st * 2 + 1) * LCB_GE_MAX, Vbar, ldb, tid, PSF_THREADS);
            __syncthreads();
            lcb_pass1T<K, G>(Vbar, ldb, nu, n, icy, taps + (st * 2) * LCB_GE_MAX, tid, PSF_THREADS,
                             [&](int v, int u, float val) { Gp[v * nu + u] = fmaf(a, val, Gp[v * nu + u]); });
            __syncthreads();
        }
        for (int j = 0; j < J; ++j) {
            const int D = 1 << j;
            for (int i = tid; i < pp; i += PSF_THREADS) C1[i] = mc_atrous(Gp, nu, i / nu, i % nu, D, 0);
            __syncthreads();
            for (int i = tid; i < pp; i += PSF_THREADS) {
                const float nxt = mc_atrous(C1, nu, i / nu, i % nu, D, 1);
                const float al = Gp[i] - nxt;
                Wf[(size_t)j * pp + i] = fmaf(al, al, Wf[(size_t)j * pp + i]);
                Gp[i] = nxt;                         // in place: Gp[i] is only read by the owner of pixel i in this pass
            }
            __syncthreads();
        }
    }
    const float inv = 1.f / (float)n_samples;
    for (int i = tid; i < J * pp; i += PSF_THREADS) Wf[i] = sqrtf(Wf[i] * inv);
}

template <int K, int G>
static int launch_nmc(const PsfArgs& A, float* W, int n_samples, unsigned seed, cudaStream_t st) {
    const size_t smem = (size_t)(A.Nmax * 2 * LCB_GE_MAX + A.n * (A.n + 1) + A.n * (A.nu + 1)) * 4;
    LCB_CUDA(cudaFuncSetAttribute(k_noise_mc<K, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    { LcbProfScope ps("k_noise_mc", st); k_noise_mc<K, G><<<A.F, PSF_THREADS, smem, st>>>(A, W, n_samples, seed, 0); }
    LCB_CUDA(cudaGetLastError());
    return LCB_OK;
}

// W: [A.F][J][nu^2] of the frames of this chunk
int lcb_noise_mc_dispatch(const PsfArgs& A, float* W, int n_samples, unsigned seed, cudaStream_t st) {
    const int G = A.cv.G;
#define CASE(KK, GG) if (A.k == KK && G == GG) return launch_nmc<KK, GG>(A, W, n_samples, seed, st);
    CASE(1, 12) CASE(2, 12) CASE(3, 12) CASE(4, 12)
    CASE(2, 8) CASE(2, 16)
#undef CASE
    lcb_set_error("noise weights (MC): unsupported (subsampling_factor=%d, gauss_taps=%d)", A.k, G);
    return LCB_ERR_ARG;
}

// tab: [J][3][nu] 1-D kernels f_j^2, f_j f_{j+1}, f_{j+1}^2 (host-computed, clamped cascade of a
// Dirac at nu/2).  W[f][j] = sqrt(max(0, sep(f_j^2) - 2 sep(f_j f_j+1) + sep(f_j+1^2))) of the
// variance plane at work[f] (written by k_noise_var or by the caller).
__global__ void __launch_bounds__(PSF_THREADS) k_noise_weights(int nu, int J, const float* tab,
                                                               float* W, float* work, size_t work_per_frame) {
    const int f = blockIdx.x, tid = threadIdx.x, pp = nu * nu;
    float* var = work + (size_t)f * work_per_frame;   // [pp]
    float* tmp = var + pp;                            // [pp]
    const int a0 = nu / 2;
    float* Wf = W + (size_t)f * J * pp;
    for (int j = 0; j < J; ++j) {
        for (int term = 0; term < 3; ++term) {
            const float* ker = tab + ((size_t)j * 3 + term) * nu;
            const float sgn = (term == 1) ? -2.f : 1.f;
            for (int i = tid; i < pp; i += PSF_THREADS) {       // along x
                const int v = i / nu, u = i % nu;
                float acc = 0.f;
                for (int m = 0; m < nu; ++m) {
                    const int uu = u + a0 - m;
                    if (uu >= 0 && uu < nu) acc = fmaf(__ldg(ker + m), var[v * nu + uu], acc);
                }
                tmp[i] = acc;
            }
            __syncthreads();
            for (int i = tid; i < pp; i += PSF_THREADS) {       // along y, accumulate into W
                const int v = i / nu, u = i % nu;
                float acc = 0.f;
                for (int m = 0; m < nu; ++m) {
                    const int vv = v + a0 - m;
                    if (vv >= 0 && vv < nu) acc = fmaf(__ldg(ker + m), tmp[vv * nu + u], acc);
                }
                const float prev = (term == 0) ? 0.f : Wf[(size_t)j * pp + i];
                const float tot = fmaf(sgn, acc, prev);
                Wf[(size_t)j * pp + i] = (term == 2) ? sqrtf(fmaxf(tot, 0.f)) : tot;
            }
            __syncthreads();
        }
    }
}

int lcb_noise_weights_launch(int F, int nu, int J, const float* tab, float* W, float* work,
                             size_t work_per_frame, cudaStream_t st) {
    { LcbProfScope ps("k_noise_weights", st); k_noise_weights<<<F, PSF_THREADS, 0, st>>>(nu, J, tab, W, work, work_per_frame); }
    LCB_CUDA(cudaGetLastError());
    return LCB_OK;
}
