// lcb_reduce.cu -- the reductions lightcurver runs on the fitted fluxes right after star_photometry, on the device and fed
// straight from K2's output arrays:
//
//   lcb_norm_medians / _scatter_matrix / _coefficients   lightcurver/processes/normalization_calculation.py:157-206
//       (calculate_coefficient): per-star median flux, the "scatter in each frame" cost of the star scaling factors
//       (cost_function_scatter_in_frame, :75-98) and the per-frame normalisation coefficient with its weighted scatter
//   lcb_zeropoints                                        lightcurver/processes/absolute_zeropoint_calculation.py:95-100
//       per-frame median and standard deviation of catalog_mag + 2.5 log10(flux)
//
// Layout: flux, dflux [F][S] frame-major (item index f*S + s, the order of K2's outputs); NaN = no measurement of that star
// in that frame (pandas skips NaN in every sum / median of the reference).  All pointers are DEVICE pointers.
//
// The cost of normalization_calculation.py:92-97 is a quadratic form in the scaling factors c:
//   cost(c) = sum_f [ sum_s w c_s^2 x^2 / W_f - (sum_s w c_s x)^2 / W_f^2 ],  w = 1/dx_sf, W_f = sum_s w   (x, dx normalised by the
//   star medians), i.e. c^T Q c with Q = sum_f [ diag(w x^2 / W) - u u^T ], u_s = w x / W.  The kernel accumulates Q in double
//   precision; the host solves the equality-constrained minimum (mean(c) = 1, :182) from its KKT system instead of iterating
//   SLSQP over pandas pivots.
#include "lcb_common.cuh"

#define RD_THREADS 256
#define RD_SMAX 64

__device__ __forceinline__ unsigned rd_key(float v) {          // order-preserving float -> uint
    const unsigned b = __float_as_uint(v);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float rd_unkey(unsigned k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// k-th smallest (0-based) of the non-NaN values x[i * stride], i < count: radix select, 4 passes of 8 bits
__device__ float rd_select(const float* __restrict__ x, int count, int stride, int kth, unsigned* hist, int tid) {
    unsigned prefix = 0, mask = 0;
    int k = kth;
    for (int shift = 24; shift >= 0; shift -= 8) {
        for (int i = tid; i < 256; i += RD_THREADS) hist[i] = 0;
        __syncthreads();
        for (int i = tid; i < count; i += RD_THREADS) {
            const float v = x[(size_t)i * stride];
            if (v == v) {
                const unsigned key = rd_key(v);
                if ((key & mask) == prefix) atomicAdd(hist + ((key >> shift) & 255u), 1u);
            }
        }
        __syncthreads();
        // every thread walks the 256 bins (cheap, and keeps the control flow uniform)
        int acc = 0, bin = 0;
        for (int b = 0; b < 256; ++b) {
            const int h = (int)hist[b];
            if (acc + h > k) { bin = b; break; }
            acc += h;
        }
        k -= acc;
        prefix |= (unsigned)bin << shift;
        mask |= 255u << shift;
        __syncthreads();
    }
    return rd_unkey(prefix);
}

// one CTA per star: pandas groupby('star_gaia_id')['flux'].median() (normalization_calculation.py:158)
__global__ void __launch_bounds__(RD_THREADS) k_star_median(const float* __restrict__ flux, int F, int S, float* __restrict__ median) {
    __shared__ unsigned hist[256];
    __shared__ int s_cnt;
    const int s = blockIdx.x, tid = threadIdx.x;
    const float* x = flux + s;
    if (tid == 0) s_cnt = 0;
    __syncthreads();
    int c = 0;
    for (int i = tid; i < F; i += RD_THREADS) { const float v = x[(size_t)i * S]; c += (v == v) ? 1 : 0; }
    c = (int)warp_sum((float)c);                  // counts stay far below 2^24
    if ((tid & 31) == 0) atomicAdd(&s_cnt, c);
    __syncthreads();
    const int cnt = s_cnt;
    if (cnt == 0) { if (tid == 0) median[s] = nanf(""); return; }
    const float lo = rd_select(x, F, S, (cnt - 1) / 2, hist, tid);
    const float hi = (cnt & 1) ? lo : rd_select(x, F, S, cnt / 2, hist, tid);
    if (tid == 0) median[s] = 0.5f * (lo + hi);
}

// Q partials: one CTA per tile of RD_TILE frames; thread (s, s') pairs accumulate in double
#define RD_TILE 64
__global__ void __launch_bounds__(RD_THREADS) k_norm_scatter(const float* __restrict__ flux, const float* __restrict__ dflux,
                                                             const float* __restrict__ median, int F, int S, double* __restrict__ Qpart) {
    __shared__ float u[RD_TILE][RD_SMAX + 1];     // w x / W
    __shared__ float dg[RD_TILE][RD_SMAX + 1];    // w x^2 / W
    const int tid = threadIdx.x, f0 = blockIdx.x * RD_TILE;
    const int nf = min(RD_TILE, F - f0);
    for (int t = tid; t < nf; t += RD_THREADS) {
        const int f = f0 + t;
        float W = 0.f;
        for (int s = 0; s < S; ++s) {
            const float m = median[s], x = flux[(size_t)f * S + s] / m, d = dflux[(size_t)f * S + s] / m;
            const float w = 1.f / d;
            const bool ok = (x == x) && (w == w);             // pandas: NaN entries are skipped by every sum
            u[t][s] = ok ? w * x : 0.f;
            dg[t][s] = ok ? w * x * x : 0.f;
            W += ok ? w : 0.f;
        }
        const float iw = (W != 0.f) ? 1.f / W : 0.f;
        for (int s = 0; s < S; ++s) { u[t][s] *= iw; dg[t][s] *= iw; }
    }
    __syncthreads();
    for (int p = tid; p < S * S; p += RD_THREADS) {
        const int a = p / S, b = p % S;
        double acc = 0.0;
        for (int t = 0; t < nf; ++t) {
            acc -= (double)u[t][a] * (double)u[t][b];
            if (a == b) acc += (double)dg[t][a];
        }
        Qpart[(size_t)blockIdx.x * S * S + p] = acc;
    }
}

__global__ void k_norm_scatter_sum(const double* __restrict__ Qpart, int nblk, int SS, double* __restrict__ Q) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= SS) return;
    double acc = 0.0;
    for (int b = 0; b < nblk; ++b) acc += Qpart[(size_t)b * SS + p];      // fixed order: deterministic
    Q[p] = acc;
}

// one thread per frame: normalization_calculation.py:185-206
__global__ void k_norm_coeff(const float* __restrict__ flux, const float* __restrict__ dflux, const float* __restrict__ median,
                             const float* __restrict__ scale, int F, int S, float* __restrict__ coef, float* __restrict__ err) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    double sw = 0.0, swv = 0.0;
    for (int s = 0; s < S; ++s) {
        const float g = scale[s] / median[s];
        const float v = flux[(size_t)f * S + s] * g, d = dflux[(size_t)f * S + s] * g;
        const float w = 1.f / (d * d);
        if (v == v && w == w) { sw += (double)w; swv += (double)w * (double)v; }
    }
    const double avg = swv / sw;                   // NaN when the frame has no usable star (0/0), as in pandas
    double var = 0.0;
    for (int s = 0; s < S; ++s) {
        const float g = scale[s] / median[s];
        const float v = flux[(size_t)f * S + s] * g, d = dflux[(size_t)f * S + s] * g;
        const float w = 1.f / (d * d);
        if (v == v && w == w) var += (double)w * ((double)v - avg) * ((double)v - avg);
    }
    float e = (float)sqrt(var / sw);
    if (e == 0.f) e = 0.1f * (float)avg;           // one star only (:203-204)
    coef[f] = (float)avg;
    err[f] = e;
}

// one thread per frame: median and std (ddof = 1, pandas) of catalog_mag - (-2.5 log10 flux) over the stars of the frame
__global__ void k_zeropoint(const float* __restrict__ flux, const float* __restrict__ cmag, int F, int S,
                            float* __restrict__ zp, float* __restrict__ zp_std) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    float v[RD_SMAX];
    int c = 0;
    double sum = 0.0;
    for (int s = 0; s < S; ++s) {
        const float d = cmag[s] + 2.5f * log10f(flux[(size_t)f * S + s]);
        if (d == d) {                              // NaN (missing or negative flux) is skipped, like pandas' median / std
            int j = c++;
            while (j > 0 && v[j - 1] > d) { v[j] = v[j - 1]; --j; }    // insertion sort: S <= 64
            v[j] = d;
            sum += (double)d;
        }
    }
    if (c == 0) { zp[f] = nanf(""); zp_std[f] = nanf(""); return; }
    zp[f] = 0.5f * (v[(c - 1) / 2] + v[c / 2]);
    const double mean = sum / c;
    double var = 0.0;
    for (int i = 0; i < c; ++i) var += ((double)v[i] - mean) * ((double)v[i] - mean);
    zp_std[f] = (c > 1) ? (float)sqrt(var / (c - 1)) : nanf("");
}

extern "C" {

int lcb_norm_medians(const float* flux, int F, int S, float* median, void* stream) {
    LCB_REQUIRE(flux && median && F >= 1 && S >= 1, "lcb_norm_medians: bad arguments");
    if (lcb_device_count() == 0) { lcb_set_error("no CUDA device: liblcb has no CPU fallback"); return LCB_ERR_CUDA; }
    cudaStream_t st = (cudaStream_t)stream;
    { LcbProfScope ps("k_star_median", st); k_star_median<<<S, RD_THREADS, 0, st>>>(flux, F, S, median); }
    LCB_CUDA(cudaGetLastError());
    return LCB_OK;
}

int lcb_norm_scatter_matrix(const float* flux, const float* dflux, const float* median, int F, int S, double* Q, double* work,
                            void* stream) {
    LCB_REQUIRE(flux && dflux && median && Q && work && F >= 1, "lcb_norm_scatter_matrix: bad arguments");
    LCB_REQUIRE(S >= 1 && S <= RD_SMAX, "lcb_norm_scatter_matrix: at most %d stars (got %d)", RD_SMAX, S);
    if (lcb_device_count() == 0) { lcb_set_error("no CUDA device: liblcb has no CPU fallback"); return LCB_ERR_CUDA; }
    cudaStream_t st = (cudaStream_t)stream;
    const int nblk = (F + RD_TILE - 1) / RD_TILE;
    LcbProfScope ps("k_norm_scatter", st);
    k_norm_scatter<<<nblk, RD_THREADS, 0, st>>>(flux, dflux, median, F, S, work);
    k_norm_scatter_sum<<<(S * S + 127) / 128, 128, 0, st>>>(work, nblk, S * S, Q);
    LCB_CUDA(cudaGetLastError());
    return LCB_OK;
}

int lcb_norm_coefficients(const float* flux, const float* dflux, const float* median, const float* star_scale, int F, int S,
                          float* coefficient, float* coefficient_uncertainty, void* stream) {
    LCB_REQUIRE(flux && dflux && median && star_scale && coefficient && coefficient_uncertainty && F >= 1 && S >= 1,
                "lcb_norm_coefficients: bad arguments");
    if (lcb_device_count() == 0) { lcb_set_error("no CUDA device: liblcb has no CPU fallback"); return LCB_ERR_CUDA; }
    cudaStream_t st = (cudaStream_t)stream;
    { LcbProfScope ps("k_norm_coeff", st);
      k_norm_coeff<<<(F + 127) / 128, 128, 0, st>>>(flux, dflux, median, star_scale, F, S, coefficient, coefficient_uncertainty); }
    LCB_CUDA(cudaGetLastError());
    return LCB_OK;
}

int lcb_zeropoints(const float* flux, const float* catalog_mag, int F, int S, float* zeropoint, float* zeropoint_uncertainty,
                   void* stream) {
    LCB_REQUIRE(flux && catalog_mag && zeropoint && zeropoint_uncertainty && F >= 1, "lcb_zeropoints: bad arguments");
    LCB_REQUIRE(S >= 1 && S <= RD_SMAX, "lcb_zeropoints: at most %d stars per frame (got %d)", RD_SMAX, S);
    if (lcb_device_count() == 0) { lcb_set_error("no CUDA device: liblcb has no CPU fallback"); return LCB_ERR_CUDA; }
    cudaStream_t st = (cudaStream_t)stream;
    { LcbProfScope ps("k_zeropoint", st); k_zeropoint<<<(F + 127) / 128, 128, 0, st>>>(flux, catalog_mag, F, S, zeropoint, zeropoint_uncertainty); }
    LCB_CUDA(cudaGetLastError());
    return LCB_OK;
}

int lcb_norm_scatter_work_doubles(int F, int S) { return ((F + RD_TILE - 1) / RD_TILE) * S * S; }

}  // extern "C"
