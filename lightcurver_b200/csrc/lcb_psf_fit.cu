// lcb_psf_fit.cu -- K1: per-frame PSF pixel-grid fit (stage 2 of starred build_psf, called at
// lightcurver/processes/psf_modelling.py:164-171): AdaBelief over {background grid b (nu^2), a_i,
// x0_i, y0_i} with the Moffat fixed, loss = 1/2 sum_i sum_p w (m_i - d_i)^2 + starlet-L1(b; W)
// (SURVEY.md A.1-A.5).  One CTA per frame runs ALL iterations: forward model, hand-derived adjoint
// (SURVEY.md B.1-B.3), global-norm clip and the AdaBelief update are fused in one kernel, with the
// grid planes (s, b, grad, mu, nu, 2 starlet scratch) resident in shared memory.
#include "lcb_psf.cuh"

// ---------------------------------------------------------------- block reduction of NV scalars
template <int NV>
__device__ __forceinline__ void block_reduce(float (&v)[NV], float* red /* [PSF_WARPS][NV] */, int tid) {
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = warp_sum(v[i]);
    if ((tid & 31) == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) red[(tid >> 5) * NV + i] = v[i];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < PSF_WARPS; ++w) s += red[w * NV + i];
        v[i] = s;
    }
}

// ---------------------------------------------------------------- starlet building blocks (A.3)
// B3-spline a-trous, edge replication.  Forward 1-D pass along x (axis=0) or y (axis=1).
__device__ __forceinline__ float atrous_fwd(const float* __restrict__ c, int nu, int v, int u, int D, int axis) {
    const float h0 = 1.f / 16.f, h1 = 4.f / 16.f, h2 = 6.f / 16.f;
    if (axis == 0) {
        const float* row = c + v * nu;
        const int um2 = max(u - 2 * D, 0), um1 = max(u - D, 0), up1 = min(u + D, nu - 1), up2 = min(u + 2 * D, nu - 1);
        return h0 * (row[um2] + row[up2]) + h1 * (row[um1] + row[up1]) + h2 * row[u];
    } else {
        const int vm2 = max(v - 2 * D, 0), vm1 = max(v - D, 0), vp1 = min(v + D, nu - 1), vp2 = min(v + 2 * D, nu - 1);
        return h0 * (c[vm2 * nu + u] + c[vp2 * nu + u]) + h1 * (c[vm1 * nu + u] + c[vp1 * nu + u]) + h2 * c[v * nu + u];
    }
}

// Transposed 1-D pass: (H^T y)[i] = sum_t h_t sum_{i'} [clamp(i' + (t-2)D) == i] y[i'].
// `stride` walks along the transformed axis, `base` points at element 0 of the line.
__device__ __forceinline__ float atrous_adj_line(const float* __restrict__ base, int stride, int nu, int i, int D) {
    const float h[5] = {1.f / 16.f, 4.f / 16.f, 6.f / 16.f, 4.f / 16.f, 1.f / 16.f};
    float acc = 0.f;
    if (i > 0 && i < nu - 1) {
#pragma unroll
        for (int t = 0; t < 5; ++t) {
            const int ip = i - (t - 2) * D;
            if (ip >= 0 && ip < nu) acc = fmaf(h[t], base[ip * stride], acc);
        }
    } else if (i == 0) {
        // clamp(i' + off) == 0  <=>  i' + off <= 0
#pragma unroll
        for (int t = 0; t < 5; ++t) {
            const int off = (t - 2) * D;
            const int hi = min(-off, nu - 1);
            float s = 0.f;
            for (int ip = 0; ip <= hi; ++ip) s += base[ip * stride];
            if (hi >= 0) acc = fmaf(h[t], s, acc);
        }
    } else {
        // clamp(i' + off) == nu-1  <=>  i' + off >= nu-1
#pragma unroll
        for (int t = 0; t < 5; ++t) {
            const int off = (t - 2) * D;
            const int lo = max(nu - 1 - off, 0);
            float s = 0.f;
            for (int ip = lo; ip < nu; ++ip) s += base[ip * stride];
            if (lo < nu) acc = fmaf(h[t], s, acc);
        }
    }
    return acc;
}

// ---------------------------------------------------------------- the kernel
template <int K, int G>
__global__ void __launch_bounds__(PSF_THREADS) k_psf_fit(PsfArgs A) {
    using P = LcbPass<K, G>;
    extern __shared__ __align__(16) float sm[];
    const int n = A.n, nu = A.nu, nn = n * n, pp = nu * nu, tid = threadIdx.x;
    const int ldv = n + 1, ldt = n + 1, ldb = nu + 1;
    const int f = blockIdx.x;
    const int i0 = A.star_off[f], N = A.star_off[f + 1] - i0;
    const DevConv cv = A.cv;
    const float fk = (float)K;
    const int J = A.J;

    // ---- shared layout (small arrays first; the 7 planes last so that they can move to global)
    float* taps = sm;                                   // [Nmax][4][LCB_GE_MAX]
    float* sp = taps + A.Nmax * 4 * LCB_GE_MAX;         // [Nmax][12] a,x0,y0, mu3, nu3, g3
    float* redS = sp + A.Nmax * 12;                     // [Nmax][PSF_WARPS][4]
    float* red = redS + A.Nmax * PSF_WARPS * 4;         // [2][PSF_WARPS][4]
    float* Vg = red + 2 * PSF_WARPS * 4;                // [nu][ldv]
    float* Vd = Vg + nu * ldv;                          // [nu][ldv]
    float* rT = Vd + nu * ldv;                          // [n][ldt]
    float* Vbar = rT + n * ldt;                         // [n][ldb]
    float* planes = Vbar + n * ldb;
    if (!A.planes_in_smem) planes = A.work + (size_t)f * A.work_per_frame + (size_t)J * pp;
    float* S = planes;            // s = s_fixed + b
    float* Bp = S + pp;           // b
    float* GR = Bp + pp;          // d chi2 / d s
    float* MU = GR + pp;
    float* NU = MU + pp;
    float* C0 = NU + pp;
    float* C1 = C0 + pp;
    float* Tj = A.work + (size_t)f * A.work_per_frame;  // [J][pp] lambda_j W_j sign(alpha_j)

    const float* sfix = A.s_fixed + (size_t)f * pp;
    const float* Wf = A.W ? A.W + (size_t)f * J * pp : nullptr;
    const float* dat = A.data + (size_t)i0 * nn;
    const float* wgt = A.weight + (size_t)i0 * nn;

    for (int i = tid; i < pp; i += PSF_THREADS) {
        const float b = A.b[(size_t)f * pp + i];
        Bp[i] = b; MU[i] = 0.f; NU[i] = 0.f; S[i] = sfix[i] + b;
    }
    for (int i = tid; i < N * 12; i += PSF_THREADS) {
        const int st = i / 12, c = i % 12;
        sp[i] = (c == 0) ? A.a[i0 + st] : (c == 1) ? A.x0[i0 + st] : (c == 2) ? A.y0[i0 + st] : 0.f;
    }
    __syncthreads();

    float b1t = 1.f, b2t = 1.f;
    int bad = 0;
    const float sc = (cv.half == 0.5f) ? 1.f : 2.f;

    for (int it = 0; it <= A.n_iter; ++it) {
        const bool last = (it == A.n_iter);
        // ---- taps of every star, zero the gradient plane
        for (int idx = tid; idx < N * 2 * P::GE; idx += PSF_THREADS) {
            const int st = idx / (2 * P::GE), rem = idx % (2 * P::GE), which = rem / P::GE, p = rem % P::GE;
            const float c = fk * sp[st * 12 + (which ? 1 : 2)];      // which=0: y axis, 1: x axis
            const float ic = floorf(c + 0.5f);
            float e, de;
            lcb_tap(cv, K, c - ic, p, e, de);
            taps[(st * 4 + (which ? 2 : 0)) * LCB_GE_MAX + p] = e;
            taps[(st * 4 + (which ? 3 : 1)) * LCB_GE_MAX + p] = de;
        }
        for (int i = tid; i < pp; i += PSF_THREADS) GR[i] = 0.f;
        __syncthreads();

        float chi = 0.f, cnt = 0.f;
        for (int st = 0; st < N; ++st) {
            const float a = sp[st * 12], cx = fk * sp[st * 12 + 1], cy = fk * sp[st * 12 + 2];
            const int icx = (int)floorf(cx + 0.5f), icy = (int)floorf(cy + 0.5f);
            const float* tp = taps + st * 4 * LCB_GE_MAX;
            lcb_pass1<K, G>(S, nu, nu, n, icy, tp, tp + LCB_GE_MAX, Vg, Vd, ldv, tid, PSF_THREADS);
            __syncthreads();
            float ga = 0.f, gx = 0.f, gy = 0.f;
            const float* ds = dat + (size_t)st * nn;
            const float* ws = wgt + (size_t)st * nn;
            float* resid = (last && A.residuals) ? A.residuals + (size_t)(i0 + st) * nn : nullptr;
            lcb_pass2<K, G>(Vg, Vd, ldv, nu, n, icx, tp + 2 * LCB_GE_MAX, tp + 3 * LCB_GE_MAX, tid, PSF_THREADS,
                            [&](int Y, int X, float m0, float mx, float my) {
                                const float d = __ldg(ds + Y * n + X), w = __ldg(ws + Y * n + X);
                                const float diff = fmaf(a, m0, -d);
                                const float r = w * diff;
                                rT[X * ldt + Y] = r;
                                chi = fmaf(r, diff, chi);
                                ga = fmaf(r, m0, ga);
                                gx = fmaf(r, mx, gx);
                                gy = fmaf(r, my, gy);
                                if (last) {
                                    cnt += (w > 0.f) ? 1.f : 0.f;
                                    if (resid) resid[Y * n + X] = -diff;
                                }
                            });
            ga = warp_sum(ga); gx = warp_sum(gx); gy = warp_sum(gy);
            if ((tid & 31) == 0) {
                float* q = redS + (st * PSF_WARPS + (tid >> 5)) * 4;
                q[0] = ga; q[1] = gx; q[2] = gy;
            }
            __syncthreads();
            if (last) continue;
            lcb_pass2T<K, G>(rT, ldt, nu, n, icx, tp + 2 * LCB_GE_MAX, Vbar, ldb, tid, PSF_THREADS);
            __syncthreads();
            lcb_pass1T<K, G>(Vbar, ldb, nu, n, icy, tp, tid, PSF_THREADS,
                             [&](int v, int u, float val) { GR[v * nu + u] = fmaf(a, val, GR[v * nu + u]); });
        }
        __syncthreads();
        // ---- per-star gradients (threads st < N)
        float gn2 = 0.f;
        if (tid < N) {
            float ga = 0.f, gx = 0.f, gy = 0.f;
            for (int w = 0; w < PSF_WARPS; ++w) {
                const float* q = redS + (tid * PSF_WARPS + w) * 4;
                ga += q[0]; gx += q[1]; gy += q[2];
            }
            const float a = sp[tid * 12];
            ga *= sc; gx *= sc * a * fk; gy *= sc * a * fk;
            sp[tid * 12 + 9] = ga; sp[tid * 12 + 10] = gx; sp[tid * 12 + 11] = gy;
            gn2 = ga * ga + gx * gx + gy * gy;
        }
        if (last) {
            float v2[2] = {chi, cnt};
            block_reduce<2>(v2, red, tid);
            if (tid == 0 && A.chi2) A.chi2[f] = v2[0] / fmaxf(v2[1], 1.f);
            break;
        }

        // ---- starlet regulariser: forward transform, loss, t_j = lambda_j W_j sign(alpha_j)
        float reg = 0.f;
        const bool do_reg = (A.lam_scales != 0.f || A.lam_hf != 0.f);
        if (do_reg) {
            for (int j = 0; j < J; ++j) {
                const int D = 1 << j;
                const float* cur = (j == 0) ? Bp : C0;
                for (int i = tid; i < pp; i += PSF_THREADS) C1[i] = atrous_fwd(cur, nu, i / nu, i % nu, D, 0);
                __syncthreads();
                const float lam = (j == 0) ? A.lam_hf : A.lam_scales;
                for (int i = tid; i < pp; i += PSF_THREADS) {
                    const float nxt = atrous_fwd(C1, nu, i / nu, i % nu, D, 1);
                    const float al = cur[i] - nxt;
                    const float lw = lam * (Wf ? __ldg(Wf + (size_t)j * pp + i) : 1.f);
                    reg = fmaf(lw, fabsf(al), reg);
                    Tj[(size_t)j * pp + i] = (al > 0.f) ? lw : (al < 0.f) ? -lw : 0.f;
                    C0[i] = nxt;
                }
                __syncthreads();
            }
            // adjoint recursion: g_J = 0; g_j = t_j + H_j^T (g_{j+1} - t_j), H_j = Hcol Hrow
            for (int j = J - 1; j >= 0; --j) {
                const int D = 1 << j;
                for (int i = tid; i < pp; i += PSF_THREADS) {
                    const float t = Tj[(size_t)j * pp + i];
                    C0[i] = ((j == J - 1) ? 0.f : C0[i]) - t;
                }
                __syncthreads();
                for (int i = tid; i < pp; i += PSF_THREADS)      // Hcol^T : along v for fixed u
                    C1[i] = atrous_adj_line(C0 + (i % nu), nu, nu, i / nu, D);
                __syncthreads();
                for (int i = tid; i < pp; i += PSF_THREADS)      // Hrow^T : along u for fixed v
                    C0[i] = Tj[(size_t)j * pp + i] + atrous_adj_line(C1 + (i / nu) * nu, 1, nu, i % nu, D);
                __syncthreads();
            }
        }
        // ---- total gradient, norm, loss
        for (int i = tid; i < pp; i += PSF_THREADS) {
            const float g = sc * GR[i] + (do_reg ? C0[i] : 0.f);
            GR[i] = g;
            gn2 = fmaf(g, g, gn2);
        }
        float v3[3] = {chi, reg, gn2};
        block_reduce<3>(v3, red + (it & 1) * PSF_WARPS * 4, tid);
        const float L = cv.half * v3[0] + v3[1];
        if (tid == 0 && A.loss_hist) A.loss_hist[(size_t)f * A.n_iter + it] = L;
        if (it == 0) {
            if (tid == 0 && A.loss0) A.loss0[f] = L;
            if (A.grad_b0) for (int i = tid; i < pp; i += PSF_THREADS) A.grad_b0[(size_t)f * pp + i] = GR[i];
            if (A.grad_s0 && tid < N) {
                A.grad_s0[(i0 + tid) * 3] = sp[tid * 12 + 9];
                A.grad_s0[(i0 + tid) * 3 + 1] = sp[tid * 12 + 10];
                A.grad_s0[(i0 + tid) * 3 + 2] = sp[tid * 12 + 11];
            }
        }
        if (!isfinite(L)) bad = 1;
        // ---- clip + schedule + AdaBelief (optax chain)
        const float gn = sqrtf(v3[2]);
        const float cs = (gn < cv.clip) ? 1.f : cv.clip / gn;
        const float lr = A.lr * powf(cv.decay, (float)it / (float)A.n_iter);
        b1t *= cv.b1; b2t *= cv.b2;
        const BeliefCoef bc = {lr, cv.b1, cv.b2, 1.f - cv.b1, 1.f - cv.b2, 1.f / (1.f - b1t), 1.f / (1.f - b2t),
                               cv.eps, cv.eps_root};
        for (int i = tid; i < pp; i += PSF_THREADS) {
            float b = Bp[i], mu = MU[i], nv = NU[i];
            belief_update(bc, cs * GR[i], b, mu, nv);
            Bp[i] = b; MU[i] = mu; NU[i] = nv;
            S[i] = __ldg(sfix + i) + b;
        }
        if (tid < N) {
            float* q = sp + tid * 12;
            belief_update(bc, cs * q[9], q[0], q[3], q[6]);
            belief_update(bc, cs * q[10], q[1], q[4], q[7]);
            belief_update(bc, cs * q[11], q[2], q[5], q[8]);
            const float lim = 0.25f * (float)n;
            q[1] = fminf(fmaxf(q[1], -lim), lim);
            q[2] = fminf(fmaxf(q[2], -lim), lim);
        }
        __syncthreads();
    }

    // ---- products: fitted parameters, narrow_psf = s / sum s, full_psf = (s (*) g0) / sum
    __syncthreads();
    for (int i = tid; i < pp; i += PSF_THREADS) A.b[(size_t)f * pp + i] = Bp[i];
    if (tid < N) { A.a[i0 + tid] = sp[tid * 12]; A.x0[i0 + tid] = sp[tid * 12 + 1]; A.y0[i0 + tid] = sp[tid * 12 + 2]; }
    if (tid == 0 && A.status) A.status[f] = bad ? LCB_ITEM_NONFINITE : LCB_ITEM_OK;
    if (A.narrow_psf || A.full_psf) {
        float tot[1] = {0.f};
        for (int i = tid; i < pp; i += PSF_THREADS) tot[0] += S[i];
        block_reduce<1>(tot, red, tid);
        const float inv = 1.f / tot[0];
        if (A.narrow_psf) for (int i = tid; i < pp; i += PSF_THREADS) A.narrow_psf[(size_t)f * pp + i] = S[i] * inv;
        if (A.full_psf) {
            float g0[G];
#pragma unroll
            for (int t = 0; t < G; ++t) {           // tau = t - G/2 + 1
                const float x = (float)(t - G / 2 + 1);
                g0[t] = cv.gnorm * expf(-x * x * cv.inv2s2);
            }
            __syncthreads();
            for (int i = tid; i < pp; i += PSF_THREADS) {
                const int v = i / nu, u = i % nu;
                float acc = 0.f;
#pragma unroll
                for (int t = 0; t < G; ++t) {
                    const int uu = u - (t - G / 2 + 1);
                    if (uu >= 0 && uu < nu) acc = fmaf(g0[t], S[v * nu + uu], acc);
                }
                C1[i] = acc;
            }
            __syncthreads();
            float tf[1] = {0.f};
            for (int i = tid; i < pp; i += PSF_THREADS) {
                const int v = i / nu, u = i % nu;
                float acc = 0.f;
#pragma unroll
                for (int t = 0; t < G; ++t) {
                    const int vv = v - (t - G / 2 + 1);
                    if (vv >= 0 && vv < nu) acc = fmaf(g0[t], C1[vv * nu + u], acc);
                }
                C0[i] = acc;
                tf[0] += acc;
            }
            block_reduce<1>(tf, red + PSF_WARPS * 4, tid);
            const float invf = 1.f / tf[0];
            for (int i = tid; i < pp; i += PSF_THREADS) A.full_psf[(size_t)f * pp + i] = C0[i] * invf;
        }
    }
}

size_t lcb_psf_fit_smem_small(int n, int nu, int Nmax) {
    return (size_t)(Nmax * 4 * LCB_GE_MAX + Nmax * 12 + Nmax * PSF_WARPS * 4 + 2 * PSF_WARPS * 4 +
                    2 * nu * (n + 1) + n * (n + 1) + n * (nu + 1)) * 4;
}

template <int K, int G>
static int launch_psf_fit(const PsfArgs& A, size_t smem, cudaStream_t st) {
    LCB_CUDA(cudaFuncSetAttribute(k_psf_fit<K, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    { LcbProfScope ps("k_psf_fit", st); k_psf_fit<K, G><<<A.F, PSF_THREADS, smem, st>>>(A); }
    LCB_CUDA(cudaGetLastError());
    return LCB_OK;
}

int lcb_psf_fit_dispatch(const PsfArgs& A, size_t smem, cudaStream_t st) {
    const int G = A.cv.G;
#define CASE(KK, GG) if (A.k == KK && G == GG) return launch_psf_fit<KK, GG>(A, smem, st);
    CASE(1, 12) CASE(2, 12) CASE(3, 12) CASE(4, 12)
    CASE(1, 8) CASE(2, 8) CASE(3, 8)
    CASE(1, 16) CASE(2, 16) CASE(3, 16)
#undef CASE
    lcb_set_error("psf fit: unsupported (subsampling_factor=%d, gauss_taps=%d)", A.k, G);
    return LCB_ERR_ARG;
}
