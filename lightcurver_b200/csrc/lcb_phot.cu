// lcb_phot.cu -- K2: fixed-PSF amplitude + shift photometry, one CTA per (frame, star) item.
//
// Replaces, per item, the STARRED calls of lightcurver/processes/star_photometry.py:66-128
// (setup_model / Loss / Optimizer('adabelief').minimize / model.model) in the per-frame form of
// SURVEY.md section 7 "hard part 2" (c fixed at 0, h = 0, mean = 0, clip per item), plus the closed
// form of lightcurver/utilities/starred_utilities.py:10-39 (sigma_a).
//
// The whole fit (all n_iter AdaBelief iterations) runs inside one kernel: the narrow PSF, the
// stamp, the weights and the intermediate planes stay in shared memory; per iteration
//   taps -> pass 1 (rows, g and dg/dy) -> pass 2 (columns, g and dg/dx) + weighted residual
//        -> block reduction (warp shuffles) of loss, dL/da, dL/ddx, dL/ddy -> AdaBelief in registers.
#include "lcb_passes.cuh"

struct PhotArgs {
    int B, n, k, nu, n_iter, schedule, s_in_smem;
    float lr;
    const float *data, *weight, *psf, *a0, *dx0, *dy0;
    const int* psf_index;
    float *a, *dx, *dy, *sigma_a, *chi2, *residuals, *loss_hist, *loss0, *grad0;
    int* status;
    DevConv cv;
};

#define PHOT_THREADS 256

// NS > 0: compile-time stamp side (no integer divisions in the passes); NS == 0: runtime sizes.
template <int K, int G, int NS>
__global__ void __launch_bounds__(PHOT_THREADS, (NS > 0 ? 4 : 1)) k_phot_fit(PhotArgs A) {
    using P = LcbPass<K, G>;
    extern __shared__ __align__(16) float sm[];
    const int n = (NS > 0) ? NS : A.n, nu = (NS > 0) ? NS * K : A.nu, tid = threadIdx.x;
    const int ldv = n + 1, ldt = n + 1;
    const int item = blockIdx.x;
    // shared layout.  NS > 0: HB zero rows before and after s and Vg/Vd (no bounds predicates in the passes)
    constexpr int HB = (NS > 0) ? 8 : 0;
    float* taps = sm;                              // ey, dey, ex, dex  [4][LCB_GE_MAX]
    float* red = taps + 4 * LCB_GE_MAX;            // [2][8 warps][4] double-buffered partials
    float* dT = red + 2 * 8 * 4;                   // [n][ldt]  data, stored [X][Y]
    float* wT = dT + n * ldt;                      // [n][ldt]
    float* Vg = wT + n * ldt + HB * ldv;           // [HB + nu + HB][ldv]
    float* Vd = Vg + (nu + 2 * HB) * ldv;          // [HB + nu + HB][ldv]
    float* s_sm = Vd + (nu + HB) * ldv + HB * nu;  // [HB + nu + HB][nu] when it fits
    const float* psf_g = A.psf + (size_t)A.psf_index[item] * nu * nu;
    const float* s = A.s_in_smem ? s_sm : psf_g;

    // The narrow PSF of the item's frame (nu x nu floats, contiguous in HBM: 16 KB at n = 32, k = 2) is fetched by the TMA unit:
    // one thread arms an mbarrier and issues ONE bulk asynchronous copy into the interior of the haloed tile while the other
    // warps zero the halos and transpose the stamp and its weights; everybody waits on the barrier before the first pass.
    __shared__ __align__(8) unsigned long long s_bar;
    const bool tma_ok = A.s_in_smem && ((nu * nu) & 3) == 0 && (reinterpret_cast<size_t>(psf_g) & 15) == 0;
    if (tid == 0) lcb_mbar_init(&s_bar, 1);
    __syncthreads();
    if (tma_ok && tid == 0) {
        lcb_mbar_expect_tx(&s_bar, (unsigned)(nu * nu) * 4u);
        lcb_bulk_g2s(s_sm, psf_g, (unsigned)(nu * nu) * 4u, &s_bar);
    }
    if (HB > 0) {
        // zero rows of the halos: everything from the end of wT up to the interior of the PSF tile, and the rows after it
        // (the interior itself belongs to the bulk copy in flight)
        float* z0 = wT + n * ldt;
        const int zc = A.s_in_smem ? (int)(s_sm - z0) : 2 * (nu + 2 * HB) * ldv;
        for (int i = tid; i < zc; i += PHOT_THREADS) z0[i] = 0.f;
        float* z1 = s_sm + nu * nu;
        if (A.s_in_smem) for (int i = tid; i < HB * nu; i += PHOT_THREADS) z1[i] = 0.f;
    }
    // ---- load: stamp + weight transposed (the transposition is why these two are not bulk copies)
    const float* dg = A.data + (size_t)item * n * n;
    const float* wg = A.weight + (size_t)item * n * n;
    for (int i = tid; i < n * n; i += PHOT_THREADS) {
        const int Y = i / n, X = i % n;
        dT[X * ldt + Y] = dg[i];
        wT[X * ldt + Y] = wg[i];
    }
    if (A.s_in_smem && !tma_ok) {
        for (int i = tid; i < nu * nu; i += PHOT_THREADS) s_sm[i] = __ldg(psf_g + i);
    }
    if (tma_ok) lcb_mbar_wait(&s_bar, 0);

    __syncthreads();
    // Only warp 0 runs the scalar part of an iteration (sum of the per-warp partials, loss, clip, schedule,
    // AdaBelief on the 3 parameters, the 2*GE taps) and publishes (a, icx, icy) + taps through shared memory:
    // with three CTAs per SM the issue slots the other seven warps would spend on the same redundant arithmetic
    // go to the other CTAs' passes.
    float* par = red + 2 * 8 * 4 - 4;               // [a, icx, icy, -] (tail of the partials area: 8 warps use 2*32 floats)
    float a = A.a0[item];
    float dx = A.dx0 ? A.dx0[item] : 0.f;
    float dy = A.dy0 ? A.dy0[item] : 0.f;
    float mu[3] = {0.f, 0.f, 0.f}, nv[3] = {0.f, 0.f, 0.f};
    float b1t = 1.f, b2t = 1.f;
    int bad = 0;
    const DevConv cv = A.cv;
    const float fk = (float)K;
    const float sched_c = log2f(cv.decay) / (float)max(A.n_iter, 1);   // lr_t = lr0 * 2^(t * log2(decay)/T)
    const bool w0 = (tid < 32);

    // iteration n_iter is the final evaluation (outputs only, no update)
    for (int it = 0; it <= A.n_iter; ++it) {
        const bool last = (it == A.n_iter);
        if (w0) {
            const float cx = fk * dx, cy = fk * dy;
            const float fx = floorf(cx + 0.5f), fy = floorf(cy + 0.5f);
            if (tid < 2 * P::GE) {                  // (readers of the previous taps passed the reduce barrier)
                const int which = tid / P::GE, p = tid % P::GE;
                float e, de;
                lcb_tap(cv, K, which ? (cx - fx) : (cy - fy), p, e, de);
                taps[(which ? 2 : 0) * LCB_GE_MAX + p] = e;
                taps[(which ? 3 : 1) * LCB_GE_MAX + p] = de;
            }
            if (tid == 0) { par[0] = a; par[1] = fx; par[2] = fy; }
        }
        __syncthreads();
        const float ac = par[0];
        const int icx = (int)par[1], icy = (int)par[2];
        const bool hal = (HB > 0) && A.s_in_smem && abs(icx) <= HB - G / 2 && abs(icy) <= HB - G / 2;
        if (hal) lcb_pass1<K, G, (NS > 0), (NS > 0 ? 8 : 4)>(s, nu, nu, n, icy, taps, taps + LCB_GE_MAX, Vg, Vd, ldv, tid, PHOT_THREADS);
        else lcb_pass1<K, G, false>(s, nu, nu, n, icy, taps, taps + LCB_GE_MAX, Vg, Vd, ldv, tid, PHOT_THREADS);
        __syncthreads();
        float loss = 0.f, ga = 0.f, gx = 0.f, gy = 0.f;   // on the last pass: chi2, H, -, -
        float* resid = (last && A.residuals) ? A.residuals + (size_t)item * n * n : nullptr;
        auto consume = [&](int Y, int X, float m0, float mx, float my, float d, float w) {
            const float diff = fmaf(ac, m0, -d);
            const float r = w * diff;
            if (!last) {
                loss = fmaf(r, diff, loss);
                ga = fmaf(r, m0, ga);
                gx = fmaf(r, mx, gx);
                gy = fmaf(r, my, gy);
            } else {
                loss = fmaf(r, diff, loss);
                ga = fmaf(w * m0, m0, ga);
                if (resid) resid[Y * n + X] = -diff;
            }
        };
        if (hal) lcb_pass2<K, G, 4, (NS > 0)>(Vg, Vd, ldv, nu, n, icx, taps + 2 * LCB_GE_MAX, taps + 3 * LCB_GE_MAX, dT, wT, ldt, tid, PHOT_THREADS, consume);
        else lcb_pass2<K, G, 4, false>(Vg, Vd, ldv, nu, n, icx, taps + 2 * LCB_GE_MAX, taps + 3 * LCB_GE_MAX, dT, wT, ldt, tid, PHOT_THREADS, consume);
        loss = warp_sum(loss); ga = warp_sum(ga); gx = warp_sum(gx); gy = warp_sum(gy);
        if ((tid & 31) == 0) reinterpret_cast<float4*>(red)[tid >> 5] = make_float4(loss, ga, gx, gy);
        __syncthreads();
        if (!w0) continue;                          // (uniform per warp; warp 0 finishes the iteration)
        float L = 0.f, Ga = 0.f, Gx = 0.f, Gy = 0.f;
#pragma unroll
        for (int w = 0; w < PHOT_THREADS / 32; ++w) {
            const float4 v = reinterpret_cast<const float4*>(red)[w];
            L += v.x; Ga += v.y; Gx += v.z; Gy += v.w;
        }
        if (last) {
            if (tid == 0) {
                if (A.chi2) A.chi2[item] = L / (float)(n * n);
                if (A.sigma_a) A.sigma_a[item] = rsqrtf((cv.half == 0.5f ? 1.f : 2.f) * Ga);
            }
            break;
        }
        L *= cv.half;
        const float sc = (cv.half == 0.5f) ? 1.f : 2.f;   // d/dp of (half * sum w diff^2) = 2*half * sum r dm/dp
        Ga *= sc;
        float Gdx = sc * a * fk * Gx, Gdy = sc * a * fk * Gy;
        if (tid == 0) {
            if (A.loss_hist) A.loss_hist[(size_t)item * A.n_iter + it] = L;
            if (it == 0) {
                if (A.loss0) A.loss0[item] = L;
                if (A.grad0) { A.grad0[item * 3] = Ga; A.grad0[item * 3 + 1] = Gdx; A.grad0[item * 3 + 2] = Gdy; }
            }
        }
        if (!isfinite(L)) bad = 1;
        // ---- optimiser (warp 0, every lane redundantly; 3 parameters)
        float lr = A.lr;
        if (A.schedule) {
            const float gn = sqrtf(Ga * Ga + Gdx * Gdx + Gdy * Gdy);
            const float cs = (gn < cv.clip) ? 1.f : cv.clip / gn;
            Ga *= cs; Gdx *= cs; Gdy *= cs;
            lr = A.lr * exp2f((float)it * sched_c);
        }
        b1t *= cv.b1; b2t *= cv.b2;
        BeliefCoef bc = {lr, cv.b1, cv.b2, 1.f - cv.b1, 1.f - cv.b2, 1.f / (1.f - b1t), 1.f / (1.f - b2t),
                         cv.eps, cv.eps_root};
        belief_update(bc, Ga, a, mu[0], nv[0]);
        belief_update(bc, Gdx, dx, mu[1], nv[1]);
        belief_update(bc, Gdy, dy, mu[2], nv[2]);
        // keep the G-tap windows inside what the halo logic supports
        const float lim = 0.25f * (float)n;
        dx = fminf(fmaxf(dx, -lim), lim);
        dy = fminf(fmaxf(dy, -lim), lim);
    }
    if (tid == 0) {
        A.a[item] = a; A.dx[item] = dx; A.dy[item] = dy;
        if (A.status) A.status[item] = bad ? LCB_ITEM_NONFINITE : LCB_ITEM_OK;
    }
}

template <int K, int G, int NS>
static int launch_phot(const PhotArgs& A, size_t smem, cudaStream_t st) {
    LCB_CUDA(cudaFuncSetAttribute(k_phot_fit<K, G, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    { LcbProfScope ps("k_phot_fit", st); k_phot_fit<K, G, NS><<<A.B, PHOT_THREADS, smem, st>>>(A); }
    LCB_CUDA(cudaGetLastError());
    return LCB_OK;
}

static int dispatch_phot(const PhotArgs& A, size_t smem, cudaStream_t st) {
    const int G = A.cv.G;
    if (G == 12 && A.k == 2 && A.n == 32) return launch_phot<2, 12, 32>(A, smem, st);
    if (G == 12 && A.k == 2 && A.n == 16) return launch_phot<2, 12, 16>(A, smem, st);
    if (G == 12 && A.k == 1 && A.n == 32) return launch_phot<1, 12, 32>(A, smem, st);
#define CASE(KK, GG) if (A.k == KK && G == GG) return launch_phot<KK, GG, 0>(A, smem, st);
    CASE(1, 12) CASE(2, 12) CASE(3, 12) CASE(4, 12)
    CASE(2, 8) CASE(2, 16)
#undef CASE
    lcb_set_error("photometry: unsupported (subsampling_factor=%d, gauss_taps=%d)", A.k, G);
    return LCB_ERR_ARG;
}

extern "C" int lcb_phot_fit_batch(const lcb_phot_batch* in, const lcb_fit_opts* opt, lcb_phot_out* out,
                                  int mem, void* stream) {
    LcbRange nvtx_range("lcb_phot_fit_batch");
    LCB_REQUIRE(in && opt && out, "lcb_phot_fit_batch: NULL argument");
    LCB_REQUIRE(in->B >= 0 && in->n >= 4 && in->k >= 1 && in->Fp >= 1, "lcb_phot_fit_batch: bad sizes B=%d n=%d k=%d Fp=%d",
                in->B, in->n, in->k, in->Fp);
    LCB_REQUIRE(opt->n_iter >= 0, "n_iter must be >= 0");
    LCB_REQUIRE(in->data && in->weight && in->psf && in->psf_index && in->a0, "lcb_phot_fit_batch: NULL input array");
    LCB_REQUIRE(out->a && out->dx && out->dy, "lcb_phot_fit_batch: a/dx/dy outputs are mandatory");
    LCB_REQUIRE(((size_t)in->n * in->k * in->n * in->k) % 4 == 0, "PSF plane must be a multiple of 4 floats");
    if (in->B == 0) return LCB_OK;
    if (lcb_device_count() == 0) { lcb_set_error("no CUDA device: liblcb has no CPU fallback"); return LCB_ERR_CUDA; }
    cudaStream_t st = (cudaStream_t)stream;
    const int B = in->B, n = in->n, k = in->k, nu = n * k, T = opt->n_iter;
    PhotArgs A;
    memset(&A, 0, sizeof(A));
    A.B = B; A.n = n; A.k = k; A.nu = nu; A.n_iter = T; A.schedule = opt->schedule; A.lr = opt->lr;
    A.cv = lcb_devconv();
    const size_t fixed = (size_t)(4 * LCB_GE_MAX + 64 + 2 * n * (n + 1) + 2 * (nu + 16) * (n + 1)) * 4;
    const size_t with_s = fixed + (size_t)(nu + 16) * nu * 4;
    int dev = 0, maxsm = 0;
    LCB_CUDA(cudaGetDevice(&dev));
    LCB_CUDA(cudaDeviceGetAttribute(&maxsm, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    LCB_REQUIRE(fixed <= (size_t)maxsm, "photometry: stamp %dx%d (k=%d) needs %zu B of shared memory (> %d)", n, n, k, fixed, maxsm);
    A.s_in_smem = (with_s <= (size_t)maxsm) ? 1 : 0;
    const size_t smem = A.s_in_smem ? with_s : fixed;

    if (mem == LCB_MEM_DEVICE) {
        A.data = in->data; A.weight = in->weight; A.psf = in->psf; A.psf_index = in->psf_index;
        A.a0 = in->a0; A.dx0 = in->dx0; A.dy0 = in->dy0;
        A.a = out->a; A.dx = out->dx; A.dy = out->dy; A.sigma_a = out->sigma_a; A.chi2 = out->chi2;
        A.residuals = out->residuals; A.loss_hist = out->loss_hist; A.loss0 = out->loss0; A.grad0 = out->grad0;
        A.status = out->status;
        return dispatch_phot(A, smem, st);
    }
    LCB_REQUIRE(mem == LCB_MEM_HOST, "mem must be LCB_MEM_DEVICE or LCB_MEM_HOST");
    // ---- host pointers: stage through the arena
    LcbArenaLease lease;
    LcbArena& ar = *lease.a;
    const size_t nn = (size_t)n * n, pp = (size_t)nu * nu;
    size_t need = 2 * B * nn * 4 + (size_t)in->Fp * pp * 4 + (size_t)B * 4 * 16 + (size_t)B * nn * 4 +
                  (size_t)B * T * 4 + 64 * 256;
    int rc = ar.reserve(need);
    if (rc) return rc;
    ar.rewind();
    auto up = [&](const void* h, size_t bytes, const void** d) -> int {
        if (!h) { *d = nullptr; return LCB_OK; }
        void* p = ar.take(bytes);
        if (!p) { lcb_set_error("arena overflow"); return LCB_ERR_NOMEM; }
        LCB_CUDA(cudaMemcpyAsync(p, h, bytes, cudaMemcpyHostToDevice, st));
        *d = p;
        return LCB_OK;
    };
    auto dn = [&](void* h, size_t bytes, void** d) -> int {
        if (!h) { *d = nullptr; return LCB_OK; }
        *d = ar.take(bytes);
        if (!*d) { lcb_set_error("arena overflow"); return LCB_ERR_NOMEM; }
        return LCB_OK;
    };
#define UP(f, bytes) if ((rc = up(in->f, bytes, (const void**)&A.f))) return rc;
    UP(data, B * nn * 4) UP(weight, B * nn * 4) UP(psf, in->Fp * pp * 4) UP(psf_index, (size_t)B * 4)
    UP(a0, (size_t)B * 4) UP(dx0, (size_t)B * 4) UP(dy0, (size_t)B * 4)
#undef UP
#define DN(f, bytes) if ((rc = dn(out->f, bytes, (void**)&A.f))) return rc;
    DN(a, (size_t)B * 4) DN(dx, (size_t)B * 4) DN(dy, (size_t)B * 4) DN(sigma_a, (size_t)B * 4) DN(chi2, (size_t)B * 4)
    DN(residuals, B * nn * 4) DN(loss_hist, (size_t)B * T * 4) DN(loss0, (size_t)B * 4) DN(grad0, (size_t)B * 12)
    DN(status, (size_t)B * 4)
#undef DN
    rc = dispatch_phot(A, smem, st);
    if (rc) return rc;
#define BACK(f, bytes) if (out->f) LCB_CUDA(cudaMemcpyAsync(out->f, A.f, bytes, cudaMemcpyDeviceToHost, st));
    BACK(a, (size_t)B * 4) BACK(dx, (size_t)B * 4) BACK(dy, (size_t)B * 4) BACK(sigma_a, (size_t)B * 4) BACK(chi2, (size_t)B * 4)
    BACK(residuals, B * nn * 4) BACK(loss_hist, (size_t)B * T * 4) BACK(loss0, (size_t)B * 4) BACK(grad0, (size_t)B * 12)
    BACK(status, (size_t)B * 4)
#undef BACK
    LCB_CUDA(cudaStreamSynchronize(st));
    return LCB_OK;
}
