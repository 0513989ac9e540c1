// lcb_psf_api.cu -- host side of lcb_psf_fit_batch: stage 1 (Moffat LM) -> noise weights -> stage 2
// (pixel grid AdaBelief), chunked over frames so that the per-frame workspace stays bounded.
#include "lcb_psf.cuh"
#include <vector>
#include <cstdlib>
#include <mutex>
#include <atomic>

size_t lcb_psf_fit_smem_small(int n, int nu, int Nmax);
size_t lcb_psf_lm_smem_small(int n, int nu, int Nmax);
int lcb_psf_fit_dispatch(const PsfArgs& A, size_t smem, bool fast, cudaStream_t st);
size_t lcb_psf_fit_smem_fast_extra(int n, int nu, int J);
bool lcb_psf_fit_has_fast(int n, int k, int G);
int lcb_psf_lm_dispatch(const PsfArgs& A, size_t smem, cudaStream_t st);
bool lcb_psf_fit_cluster_ok(int n, int k, int G, int Nmax, int J, int max_smem);
int lcb_psf_fit_cluster_dispatch(const PsfArgs& A, cudaStream_t st);
int lcb_moffat_image_launch(const PsfArgs& A, cudaStream_t st);
int lcb_noise_var_dispatch(const PsfArgs& A, cudaStream_t st);
int lcb_noise_mc_dispatch(const PsfArgs& A, float* W, int n_samples, unsigned seed, cudaStream_t st);
int lcb_noise_weights_launch(int F, int nu, int J, const float* tab, float* W, float* work,
                             size_t work_per_frame, cudaStream_t st);

extern "C" int lcb_starlet_scales(int nu) {
    int J = 0;
    while ((1 << (J + 1)) <= nu) ++J;
    return J;
}

// 1-D kernels of the starlet-space noise propagation: f_j = H_{j-1}..H_0 delta_{nu/2} (clamped a-trous
// cascade, double precision), table [J][3][nu] = f_j^2, f_j f_{j+1}, f_{j+1}^2.
void lcb_build_noise_table(int nu, int J, std::vector<float>& tab) {
    std::vector<double> f(nu, 0.0), g(nu, 0.0);
    f[nu / 2] = 1.0;
    tab.assign((size_t)J * 3 * nu, 0.f);
    const double h[5] = {1.0 / 16, 4.0 / 16, 6.0 / 16, 4.0 / 16, 1.0 / 16};
    for (int j = 0; j < J; ++j) {
        const int D = 1 << j;
        for (int i = 0; i < nu; ++i) {
            double s = 0.0;
            for (int t = 0; t < 5; ++t) {
                int ip = i + (t - 2) * D;
                ip = ip < 0 ? 0 : (ip >= nu ? nu - 1 : ip);
                s += h[t] * f[ip];
            }
            g[i] = s;
        }
        for (int i = 0; i < nu; ++i) {
            tab[((size_t)j * 3 + 0) * nu + i] = (float)(f[i] * f[i]);
            tab[((size_t)j * 3 + 1) * nu + i] = (float)(f[i] * g[i]);
            tab[((size_t)j * 3 + 2) * nu + i] = (float)(g[i] * g[i]);
        }
        f = g;
    }
}

// Keep freed stream-ordered allocations cached in the device's default pool: with the default release
// threshold (0) every call would hand ~400 MB of workspace back to the driver at the next synchronisation
// and pay for a fresh allocation on the following call.
static void keep_pool_cached() {
    static std::atomic<unsigned long long> done_mask{0};     // one bit per device; host threads of several GPUs call this
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return;
    const unsigned long long bit = 1ull << dev;
    if (done_mask.load(std::memory_order_relaxed) & bit) return;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        unsigned long long thr = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    done_mask.fetch_or(bit, std::memory_order_relaxed);
}

struct DevTemp {       // stream-ordered temporary
    void* p = nullptr; cudaStream_t st;
    explicit DevTemp(cudaStream_t s) : st(s) { keep_pool_cached(); }
    int alloc(size_t bytes) {
        cudaError_t e = cudaMallocAsync(&p, bytes ? bytes : 4, st);
        if (e != cudaSuccess) { lcb_set_error("cudaMallocAsync(%zu): %s", bytes, cudaGetErrorString(e)); p = nullptr; return LCB_ERR_NOMEM; }
        return LCB_OK;
    }
    ~DevTemp() { if (p) cudaFreeAsync(p, st); }
};

// All pointers are device pointers here.
static int psf_run_device(const lcb_psf_batch* in, const lcb_psf_opts* opt, lcb_psf_out* out, int sumN,
                          int Nmax, cudaStream_t st) {
    const int F = in->F, n = in->n, k = in->k, nu = n * k, pp = nu * nu;
    const int J = lcb_starlet_scales(nu);
    LCB_REQUIRE(J <= LCB_JMAX, "grid side %d needs %d starlet scales (> %d)", nu, J, LCB_JMAX);
    int dev = 0, maxsm = 0;
    LCB_CUDA(cudaGetDevice(&dev));
    LCB_CUDA(cudaDeviceGetAttribute(&maxsm, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));

    const size_t fit_small = lcb_psf_fit_smem_small(n, nu, Nmax);
    const size_t lm_small = lcb_psf_lm_smem_small(n, nu, Nmax);
    LCB_REQUIRE(fit_small <= (size_t)maxsm && lm_small <= (size_t)maxsm,
                "psf fit: n=%d k=%d Nmax=%d needs %zu B of shared memory (> %d)", n, k, Nmax,
                fit_small > lm_small ? fit_small : lm_small, maxsm);
    const bool fit_planes_sm = fit_small + (size_t)7 * pp * 4 <= (size_t)maxsm;
    const size_t lm_jim = (size_t)8 * n * (n + 1) * 4;
    const bool lm_jim_sm = lm_small + lm_jim <= (size_t)maxsm;
    const bool lm_planes_sm = lm_jim_sm && lm_small + lm_jim + (size_t)5 * pp * 4 <= (size_t)maxsm;
    const bool distort = opt->field_distortion != 0;
    LCB_REQUIRE(!distort || (in->stamp_xy && out->distortion), "field_distortion needs stamp_xy and the distortion in/out array");
    LCB_REQUIRE(opt->field_distortion >= 0 && opt->field_distortion <= 2, "field_distortion must be 0, 1 or 2");
    // floats per frame of workspace (+ the resampled PSF of the current star and its gradient plane with field distortion)
    const size_t wpf = (size_t)(J + 7) * pp + (size_t)2 * Nmax * n * n + (distort ? (size_t)2 * pp : 0);
    const bool fit_fast = !distort && lcb_psf_fit_has_fast(n, k, lcb_conv().gauss_taps) &&
                          fit_small + lcb_psf_fit_smem_fast_extra(n, nu, J) <= (size_t)maxsm;
    int chunk_max = 1184;                                    // 8 waves of 148 CTAs
    if (const char* ce = getenv("LCB_PSF_CHUNK")) { const int c = atoi(ce); if (c >= 1) chunk_max = c; }   // tests exercise the chunk loop
    const int chunk = F < chunk_max ? F : chunk_max;
    // grids too large for one SM: one 8-CTA cluster per frame with the planes distributed over its shared memories
    // (LCB_PSF_CLUSTER=0 / 1 forces the single-CTA / the cluster kernel: parity tests compare the two)
    const char* cl_env = getenv("LCB_PSF_CLUSTER");
    // Measured on cfg5 shapes (profiles/README.md): the cluster kernel is bound by the 17-21 B/clk DSMEM port of an SM
    // (column passes of the starlet) and by the latency of ~150 barrier-separated phases per iteration; at 20 us per
    // iteration-frame it does not yet beat the single-CTA kernel with L2-resident planes (13.6 us), so it is opt-in.
    const bool cl_want = cl_env ? (cl_env[0] == '1') : false;
    const bool fit_cluster = cl_want && !distort && lcb_psf_fit_cluster_ok(n, k, lcb_conv().gauss_taps, Nmax, J, maxsm);

    DevTemp work(st), sfix(st), Wtmp(st), tabd(st);
    int rc;
    if ((rc = work.alloc((size_t)chunk * wpf * 4))) return rc;
    if ((rc = sfix.alloc((size_t)F * pp * 4))) return rc;
    float* Wuse = nullptr;
    const bool do_reg = (opt->lam_scales != 0.f || opt->lam_hf != 0.f);
    if (opt->noise_weights && do_reg) {
        if (out->W_out) Wuse = out->W_out;
        else { if ((rc = Wtmp.alloc((size_t)F * J * pp * 4))) return rc; Wuse = (float*)Wtmp.p; }
        std::vector<float> tab;
        lcb_build_noise_table(nu, J, tab);
        if ((rc = tabd.alloc(tab.size() * 4))) return rc;
        LCB_CUDA(cudaMemcpyAsync(tabd.p, tab.data(), tab.size() * 4, cudaMemcpyHostToDevice, st));
        LCB_CUDA(cudaStreamSynchronize(st));                 // tab is a host temporary
    } else if (in->W && do_reg) {
        Wuse = const_cast<float*>(in->W);
        if (out->W_out) LCB_CUDA(cudaMemcpyAsync(out->W_out, in->W, (size_t)F * J * pp * 4, cudaMemcpyDeviceToDevice, st));
    }
    if (out->status) LCB_CUDA(cudaMemsetAsync(out->status, 0, (size_t)F * 4, st));

    for (int f0 = 0; f0 < F; f0 += chunk) {
        const int Fc = (F - f0 < chunk) ? F - f0 : chunk;
        PsfArgs A;
        memset(&A, 0, sizeof(A));
        A.F = Fc; A.n = n; A.k = k; A.nu = nu; A.J = J; A.Nmax = Nmax;
        A.n_iter = opt->n_iter_adabelief; A.n_iter_lm = opt->n_iter_analytic;
        A.lr = opt->lr; A.lam_scales = opt->lam_scales; A.lam_hf = opt->lam_hf;
        A.star_off = in->star_off + f0;
        A.data = in->data; A.weight = in->weight;
        A.W = Wuse ? Wuse + (size_t)f0 * J * pp : nullptr;
        A.s_fixed = (float*)sfix.p + (size_t)f0 * pp;
        A.b = out->background + (size_t)f0 * pp;
        A.a = out->a; A.x0 = out->x0; A.y0 = out->y0;
        A.moffat = out->moffat + (size_t)f0 * 5;
        A.loss_hist = out->loss_hist ? out->loss_hist + (size_t)f0 * opt->n_iter_adabelief : nullptr;
        A.loss_hist_lm = out->loss_hist_analytic ? out->loss_hist_analytic + (size_t)f0 * opt->n_iter_analytic : nullptr;
        A.residuals = out->residuals;
        A.chi2 = out->chi2 ? out->chi2 + f0 : nullptr;
        A.narrow_psf = out->narrow_psf ? out->narrow_psf + (size_t)f0 * pp : nullptr;
        A.full_psf = out->full_psf ? out->full_psf + (size_t)f0 * pp : nullptr;
        A.loss0 = out->loss0 ? out->loss0 + f0 : nullptr;
        A.grad_b0 = out->grad_b0 ? out->grad_b0 + (size_t)f0 * pp : nullptr;
        A.grad_s0 = out->grad_s0;
        A.work = (float*)work.p; A.work_per_frame = wpf;
        A.status = out->status ? out->status + f0 : nullptr;
        A.fwhm_min = opt->fwhm_min; A.fwhm_max = opt->fwhm_max; A.beta_min = opt->beta_min; A.beta_max = opt->beta_max;
        A.cv = lcb_devconv();
        A.distort = opt->field_distortion;
        A.stamp_xy = in->stamp_xy;
        A.distortion = distort ? out->distortion + (size_t)f0 * 6 : nullptr;
        A.grad_dist0 = (distort && out->grad_dist0) ? out->grad_dist0 + (size_t)f0 * 6 : nullptr;
        if (opt->n_iter_analytic > 0) {
            A.planes_in_smem = lm_planes_sm;
            A.jim_in_smem = lm_jim_sm;
            if ((rc = lcb_psf_lm_dispatch(A, lm_small + (lm_jim_sm ? lm_jim : 0) + (lm_planes_sm ? (size_t)5 * pp * 4 : 0), st))) return rc;
        } else {
            if ((rc = lcb_moffat_image_launch(A, st))) return rc;
        }
        if (opt->noise_weights == 2 && do_reg) {
            // the per-(star, pixel) counters are local to the chunk: mix the first frame into the seed
            if ((rc = lcb_noise_mc_dispatch(A, Wuse + (size_t)f0 * J * pp, opt->mc_samples > 0 ? opt->mc_samples : 100,
                                            opt->mc_seed + 0x9e3779b9u * (unsigned)f0, st))) return rc;
        } else if (opt->noise_weights && do_reg) {
            if ((rc = lcb_noise_var_dispatch(A, st))) return rc;
            if ((rc = lcb_noise_weights_launch(Fc, nu, J, (const float*)tabd.p, Wuse + (size_t)f0 * J * pp,
                                               (float*)work.p, wpf, st))) return rc;
        }
        A.planes_in_smem = fit_planes_sm;
        const size_t fit_smem = fit_fast ? fit_small + lcb_psf_fit_smem_fast_extra(n, nu, J)
                                         : fit_small + (fit_planes_sm ? (size_t)7 * pp * 4 : 0);
        if (fit_cluster && A.n_iter > 0) {
            if ((rc = lcb_psf_fit_cluster_dispatch(A, st))) return rc;
            // products (residuals, chi2, narrow / full PSF) at the fitted parameters: the single-CTA kernel with 0 iterations
            A.n_iter = 0; A.loss_hist = nullptr; A.loss0 = nullptr; A.grad_b0 = nullptr; A.grad_s0 = nullptr; A.status = nullptr;
        }
        if ((rc = lcb_psf_fit_dispatch(A, fit_smem, fit_fast && !(fit_cluster && opt->n_iter_adabelief > 0), st))) return rc;
    }
    (void)sumN;
    return LCB_OK;
}

extern "C" int lcb_psf_fit_batch(const lcb_psf_batch* in, const lcb_psf_opts* opt, lcb_psf_out* out,
                                 int mem, void* stream) {
    LcbRange nvtx_range("lcb_psf_fit_batch");
    LCB_REQUIRE(in && opt && out, "lcb_psf_fit_batch: NULL argument");
    LCB_REQUIRE(in->F >= 0 && in->n >= 4 && in->k >= 1, "lcb_psf_fit_batch: bad sizes F=%d n=%d k=%d", in->F, in->n, in->k);
    LCB_REQUIRE(opt->n_iter_analytic >= 0 && opt->n_iter_adabelief >= 0, "iteration counts must be >= 0");
    LCB_REQUIRE(in->star_off && in->data && in->weight, "lcb_psf_fit_batch: NULL input array");
    LCB_REQUIRE(out->moffat && out->a && out->x0 && out->y0 && out->background,
                "lcb_psf_fit_batch: moffat/a/x0/y0/background are mandatory in/out arrays");
    if (in->F == 0) return LCB_OK;
    if (lcb_device_count() == 0) { lcb_set_error("no CUDA device: liblcb has no CPU fallback"); return LCB_ERR_CUDA; }
    cudaStream_t st = (cudaStream_t)stream;
    const int F = in->F, n = in->n, k = in->k, nu = n * k;
    const size_t nn = (size_t)n * n, pp = (size_t)nu * nu;
    const int J = lcb_starlet_scales(nu);

    // star_off is needed on the host to size things
    std::vector<int> off(F + 1);
    if (mem == LCB_MEM_HOST) memcpy(off.data(), in->star_off, (size_t)(F + 1) * 4);
    else {
        LCB_CUDA(cudaMemcpyAsync(off.data(), in->star_off, (size_t)(F + 1) * 4, cudaMemcpyDeviceToHost, st));
        LCB_CUDA(cudaStreamSynchronize(st));
    }
    int Nmax = 0;
    for (int f = 0; f < F; ++f) {
        const int N = off[f + 1] - off[f];
        LCB_REQUIRE(N >= 1, "frame %d has %d stars (need >= 1; drop empty frames in the caller)", f, N);
        if (N > Nmax) Nmax = N;
    }
    LCB_REQUIRE(off[0] == 0, "star_off[0] must be 0");
    LCB_REQUIRE(Nmax <= 128, "at most 128 stars per frame (got %d)", Nmax);
    const int sumN = off[F];

    if (mem == LCB_MEM_DEVICE) return psf_run_device(in, opt, out, sumN, Nmax, st);
    LCB_REQUIRE(mem == LCB_MEM_HOST, "mem must be LCB_MEM_DEVICE or LCB_MEM_HOST");

    // ---- host pointers: stage everything through the arena
    const int T1 = opt->n_iter_analytic, T2 = opt->n_iter_adabelief;
    size_t need = (size_t)(F + 1) * 4 + 2 * sumN * nn * 4 + (in->W || out->W_out ? (size_t)F * J * pp * 4 : 0) +
                  (size_t)F * 5 * 4 + 3 * (size_t)sumN * 4 + 4 * F * pp * 4 + sumN * nn * 4 + (size_t)F * 4 * 3 +
                  (size_t)F * (T1 + T2) * 4 + (size_t)sumN * 12 + (size_t)sumN * 8 + (size_t)F * 48 + 64 * 256;
    LcbArenaLease lease;
    LcbArena& ar = *lease.a;
    int rc = ar.reserve(need);
    if (rc) return rc;
    ar.rewind();
    lcb_psf_batch din = *in;
    lcb_psf_out dout;
    memset(&dout, 0, sizeof(dout));
    struct Back { void* h; void* d; size_t bytes; };
    std::vector<Back> back;
    auto up = [&](const void* h, size_t bytes, const void** d) -> int {
        *d = nullptr;
        if (!h) return LCB_OK;
        void* p = ar.take(bytes);
        if (!p) { lcb_set_error("arena overflow"); return LCB_ERR_NOMEM; }
        LCB_CUDA(cudaMemcpyAsync(p, h, bytes, cudaMemcpyHostToDevice, st));
        *d = p;
        return LCB_OK;
    };
    auto io = [&](void* h, size_t bytes, void** d, bool upload) -> int {
        *d = nullptr;
        if (!h) return LCB_OK;
        void* p = ar.take(bytes);
        if (!p) { lcb_set_error("arena overflow"); return LCB_ERR_NOMEM; }
        if (upload) LCB_CUDA(cudaMemcpyAsync(p, h, bytes, cudaMemcpyHostToDevice, st));
        *d = p;
        back.push_back({h, p, bytes});
        return LCB_OK;
    };
#define UP(f, bytes) if ((rc = up(in->f, bytes, (const void**)&din.f))) return rc;
    UP(star_off, (size_t)(F + 1) * 4) UP(data, sumN * nn * 4) UP(weight, sumN * nn * 4)
    UP(W, (size_t)F * J * pp * 4) UP(stamp_xy, (size_t)sumN * 8)
#undef UP
#define IO(f, bytes, upl) if ((rc = io(out->f, bytes, (void**)&dout.f, upl))) return rc;
    IO(moffat, (size_t)F * 20, true) IO(a, (size_t)sumN * 4, true) IO(x0, (size_t)sumN * 4, true) IO(y0, (size_t)sumN * 4, true)
    IO(background, F * pp * 4, true)
    IO(narrow_psf, F * pp * 4, false) IO(full_psf, F * pp * 4, false) IO(residuals, sumN * nn * 4, false)
    IO(chi2, (size_t)F * 4, false) IO(loss_hist, (size_t)F * T2 * 4, false) IO(loss_hist_analytic, (size_t)F * T1 * 4, false)
    IO(W_out, (size_t)F * J * pp * 4, false) IO(loss0, (size_t)F * 4, false) IO(grad_b0, F * pp * 4, false)
    IO(grad_s0, (size_t)sumN * 12, false) IO(status, (size_t)F * 4, false)
    IO(distortion, (size_t)F * 24, true) IO(grad_dist0, (size_t)F * 24, false)
#undef IO
    rc = psf_run_device(&din, opt, &dout, sumN, Nmax, st);
    if (rc) return rc;
    for (const Back& b : back) LCB_CUDA(cudaMemcpyAsync(b.h, b.d, b.bytes, cudaMemcpyDeviceToHost, st));
    LCB_CUDA(cudaStreamSynchronize(st));
    return LCB_OK;
}

// ---------------------------------------------------------------- apply_distortion (one CTA per item)
#include "lcb_distort.cuh"
__global__ void __launch_bounds__(256) k_apply_distortion(const float* __restrict__ psf, const float* __restrict__ theta,
                                                          const int* __restrict__ psf_index, const float* __restrict__ xy,
                                                          int nu, int mode, float* __restrict__ out) {
    const int i = blockIdx.x, f = psf_index[i];
    const float* s = psf + (size_t)f * nu * nu;
    const LcbAffine A = lcb_affine(theta + (size_t)f * 6, xy[2 * i], xy[2 * i + 1], mode);
    float* o = out + (size_t)i * nu * nu;
    for (int p = threadIdx.x; p < nu * nu; p += blockDim.x)
        o[p] = A.det * lcb_bilin_value(lcb_bilin(s, nu, nu, A, p % nu, p / nu));
}

extern "C" int lcb_apply_distortion_batch(const float* psf, const float* theta, const int* psf_index, const float* xy, int B, int Fp,
                                          int nu, int mode, float* out, int mem, void* stream) {
    LCB_REQUIRE(psf && theta && psf_index && xy && out, "lcb_apply_distortion_batch: NULL argument");
    LCB_REQUIRE(B >= 0 && Fp >= 1 && nu >= 2 && (mode == 1 || mode == 2), "lcb_apply_distortion_batch: bad sizes / mode");
    if (B == 0) return LCB_OK;
    if (lcb_device_count() == 0) { lcb_set_error("no CUDA device: liblcb has no CPU fallback"); return LCB_ERR_CUDA; }
    cudaStream_t st = (cudaStream_t)stream;
    const size_t pp = (size_t)nu * nu;
    if (mem == LCB_MEM_DEVICE) {
        { LcbProfScope ps("k_apply_distortion", st); k_apply_distortion<<<B, 256, 0, st>>>(psf, theta, psf_index, xy, nu, mode, out); }
        LCB_CUDA(cudaGetLastError());
        return LCB_OK;
    }
    LCB_REQUIRE(mem == LCB_MEM_HOST, "mem must be LCB_MEM_DEVICE or LCB_MEM_HOST");
    LcbArenaLease lease;
    LcbArena& ar = *lease.a;
    int rc = ar.reserve((size_t)Fp * pp * 4 + (size_t)Fp * 24 + (size_t)B * 12 + (size_t)B * pp * 4 + 8 * 256);
    if (rc) return rc;
    ar.rewind();
    float* dpsf = (float*)ar.take((size_t)Fp * pp * 4); float* dth = (float*)ar.take((size_t)Fp * 24);
    int* didx = (int*)ar.take((size_t)B * 4); float* dxy = (float*)ar.take((size_t)B * 8); float* dout = (float*)ar.take((size_t)B * pp * 4);
    if (!dpsf || !dth || !didx || !dxy || !dout) { lcb_set_error("arena overflow"); return LCB_ERR_NOMEM; }
    LCB_CUDA(cudaMemcpyAsync(dpsf, psf, (size_t)Fp * pp * 4, cudaMemcpyHostToDevice, st));
    LCB_CUDA(cudaMemcpyAsync(dth, theta, (size_t)Fp * 24, cudaMemcpyHostToDevice, st));
    LCB_CUDA(cudaMemcpyAsync(didx, psf_index, (size_t)B * 4, cudaMemcpyHostToDevice, st));
    LCB_CUDA(cudaMemcpyAsync(dxy, xy, (size_t)B * 8, cudaMemcpyHostToDevice, st));
    { LcbProfScope ps("k_apply_distortion", st); k_apply_distortion<<<B, 256, 0, st>>>(dpsf, dth, didx, dxy, nu, mode, dout); }
    LCB_CUDA(cudaGetLastError());
    LCB_CUDA(cudaMemcpyAsync(out, dout, (size_t)B * pp * 4, cudaMemcpyDeviceToHost, st));
    LCB_CUDA(cudaStreamSynchronize(st));
    return LCB_OK;
}
