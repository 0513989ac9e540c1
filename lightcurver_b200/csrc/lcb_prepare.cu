// lcb_prepare.cu -- the host-side data policies of the two batched drivers, on the device.
//
//   lcb_psf_prepare_batch   psf_modelling.py:136-140 (NaN policy) and starred build_psf's normalisation / smart guess
//                           (SURVEY.md A.4): raw stamps, noise maps and masks -> normalised stamps, weights = mask/sigma^2,
//                           initial amplitudes and positions, per-frame normalisation
//   lcb_phot_prepare_batch  star_photometry.py:47-64, 309-316: NaN -> (0, 1e7), whole-epoch x1000 mask rule, per-star
//                           scale = max over all epochs, initial flux guess = sum - n^2 * (mean of the edge medians over
//                           the four edges and all epochs), weights = 1/sigma^2
// With these the public API uploads the raw (pinned) arrays once and no numpy pass over the stamps remains on the host.
// All pointers are DEVICE pointers; work is enqueued on `stream`.
#include "lcb_common.cuh"

#define PR_THREADS 256

__device__ __forceinline__ float pr_block_max(float v, float* red, int tid) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();
    if ((tid & 31) == 0) red[tid >> 5] = v;
    __syncthreads();
    float m = red[0];
    for (int w = 1; w < PR_THREADS / 32; ++w) m = fmaxf(m, red[w]);
    return m;
}

__device__ __forceinline__ float pr_block_sum(float v, float* red, int tid) {
    v = warp_sum(v);
    __syncthreads();
    if ((tid & 31) == 0) red[tid >> 5] = v;
    __syncthreads();
    float s = 0.f;
    for (int w = 0; w < PR_THREADS / 32; ++w) s += red[w];
    return s;
}

// one CTA per frame
__global__ void __launch_bounds__(PR_THREADS) k_psf_prepare(lcb_psf_prepare_in in, lcb_psf_prepare_out out) {
    __shared__ float red[PR_THREADS / 32];
    __shared__ int redi[PR_THREADS / 32];
    const int f = blockIdx.x, tid = threadIdx.x;
    const int n = in.n, nn = n * n;
    const int i0 = in.star_off[f], N = in.star_off[f + 1] - i0;
    const float* img = in.image + (size_t)i0 * nn;
    const float* nmp = in.noisemap + (size_t)i0 * nn;
    const unsigned char* mk = in.mask ? in.mask + (size_t)i0 * nn : nullptr;
    // global normalisation of the frame (A.4): stamps / (max(image) / norm_scale); fmax ignores NaN like np.fmax
    float mx = -INFINITY;
    for (int i = tid; i < N * nn; i += PR_THREADS) mx = fmaxf(mx, img[i]);
    mx = pr_block_max(mx, red, tid);
    float norm = mx / in.norm_scale;
    if (!isfinite(norm) || norm <= 0.f) norm = 1.f;
    const float inv = 1.f / norm;
    if (tid == 0 && out.norm) out.norm[f] = norm;
    const float ctr = 0.5f * (float)(n - 1);
    const float kk = in.downsample_mean ? (float)(in.k * in.k) : 1.f;
    for (int st = 0; st < N; ++st) {
        float flux = 0.f, sw = 0.f, swx = 0.f, swy = 0.f, best = -INFINITY;
        int besti = 0x7fffffff;
        for (int p = tid; p < nn; p += PR_THREADS) {
            const size_t i = (size_t)st * nn + p;
            const float d = img[i] * inv, s = nmp[i] * inv;
            const bool good = isfinite(d) && isfinite(s) && s > 0.f && (!mk || mk[i] != 0);
            const float dd = isfinite(d) ? d : 0.f;
            out.data[(size_t)i0 * nn + i] = dd;
            out.weight[(size_t)i0 * nn + i] = good ? 1.f / (s * s) : 0.f;
            if (good) {
                flux += dd;
                const float w = fmaxf(dd, 0.f);
                sw += w; swx += w * ((float)(p % n) - ctr); swy += w * ((float)(p / n) - ctr);
                if (dd > best) { best = dd; besti = p; }      // p increases per thread: first maximum kept
            }
        }
        flux = pr_block_sum(flux, red, tid);
        float x0 = 0.f, y0 = 0.f;
        if (in.guess_method == 2) {
            sw = pr_block_sum(sw, red, tid); swx = pr_block_sum(swx, red, tid); swy = pr_block_sum(swy, red, tid);
            const float tot = fmaxf(sw, 1e-30f);
            x0 = swx / tot; y0 = swy / tot;
        } else if (in.guess_method == 1) {
            const float m = pr_block_max(best, red, tid);
            int cand = (best == m) ? besti : 0x7fffffff;     // ties: lowest pixel index (numpy argmax)
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) cand = min(cand, __shfl_xor_sync(0xffffffffu, cand, o));
            __syncthreads();
            if ((tid & 31) == 0) redi[tid >> 5] = cand;
            __syncthreads();
            int bi = redi[0];
            for (int w = 1; w < PR_THREADS / 32; ++w) bi = min(bi, redi[w]);
            if (bi == 0x7fffffff) bi = 0;
            x0 = (float)(bi % n) - ctr; y0 = (float)(bi / n) - ctr;
        }
        if (tid == 0) {
            out.a0[i0 + st] = fmaxf(flux, 1e-6f) * kk;
            out.x0[i0 + st] = x0; out.y0[i0 + st] = y0;
        }
    }
}

// ---------------------------------------------------------------- photometry: [F][S][n][n]
// stage 1: one CTA per (frame, star) item: max of the NaN-cleaned stamp, "any masked pixel" flag
__global__ void __launch_bounds__(PR_THREADS) k_phot_prep_max(lcb_phot_prepare_in in, float* item_max, int* item_bad) {
    __shared__ float red[PR_THREADS / 32];
    const int it = blockIdx.x, tid = threadIdx.x, nn = in.n * in.n;
    const float* d = in.data + (size_t)it * nn;
    const float* s = in.noisemap + (size_t)it * nn;
    const unsigned char* mk = in.mask ? in.mask + (size_t)it * nn : nullptr;
    float mx = -INFINITY, bad = 0.f;
    for (int p = tid; p < nn; p += PR_THREADS) {
        const float dv = d[p], sv = s[p];
        const float dd = (isnan(dv) || isnan(sv)) ? 0.f : dv;
        mx = fmaxf(mx, dd);
        if (mk && mk[p] == 0) bad = 1.f;
    }
    mx = pr_block_max(mx, red, tid);
    bad = pr_block_max(bad, red, tid);
    if (tid == 0) { item_max[it] = mx; item_bad[it] = bad > 0.f ? 1 : 0; }
}

// stage 2: per star, scale = max over the frames
__global__ void k_phot_prep_scale(int F, int S, const float* item_max, float* scale) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    float m = -INFINITY;
    for (int f = 0; f < F; ++f) m = fmaxf(m, item_max[(size_t)f * S + s]);
    scale[s] = m;
}

// median of n values held by one warp (lane l holds v[l], v[l+32], ...; n <= 128), by rank counting
__device__ __forceinline__ float warp_median(const float* __restrict__ vals, int n, int lane) {
    // vals in shared memory, all lanes see all values
    float lo = 0.f, hi = 0.f;
    const int k0 = (n - 1) / 2, k1 = n / 2;
    for (int i = lane; i < n; i += 32) {
        const float vi = vals[i];
        int rank = 0;
        for (int j = 0; j < n; ++j) { const float vj = vals[j]; rank += (vj < vi || (vj == vi && j < i)) ? 1 : 0; }
        if (rank == k0) lo = vi;
        if (rank == k1) hi = vi;
    }
    lo = warp_sum(lo); hi = warp_sum(hi);          // exactly one lane contributes to each
    return 0.5f * (lo + hi);
}

// stage 3: one CTA (128 threads = 4 warps, one per edge) per item: scaled stamp, weight, sum, edge medians
__global__ void __launch_bounds__(128) k_phot_prep_item(lcb_phot_prepare_in in, lcb_phot_prepare_out out, const float* scale,
                                                        const int* item_bad, float* item_sum, float* item_edge) {
    extern __shared__ float edge[];                 // [4][n]
    __shared__ float red[4];
    const int it = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = in.n, nn = n * n, S = in.S, s = it % S;
    const float inv = 1.f / scale[s];
    const float nfac = item_bad[it] ? 1000.f : 1.f;
    const float* d = in.data + (size_t)it * nn;
    const float* sg = in.noisemap + (size_t)it * nn;
    float sum = 0.f;
    for (int p = tid; p < nn; p += 128) {
        const float dv = d[p], sv = sg[p];
        const bool isn = isnan(dv) || isnan(sv);
        const float dd = (isn ? 0.f : dv) * inv;
        const float ss = (isn ? 1e7f : sv) * nfac * inv;
        out.data[(size_t)it * nn + p] = dd;
        out.weight[(size_t)it * nn + p] = 1.f / (ss * ss);
        sum += dd;
        const int y = p / n, x = p % n;
        if (y == 0) edge[x] = dd;                   // d[0, :]
        if (x == 0) edge[n + y] = dd;               // d[:, 0]
        if (y == n - 1) edge[2 * n + x] = dd;       // d[-1, :]
        if (x == n - 1) edge[3 * n + y] = dd;       // d[:, -1]
    }
    sum = warp_sum(sum);
    if (lane == 0) red[warp] = sum;
    __syncthreads();
    const float med = warp_median(edge + warp * n, n, lane);
    if (lane == 0) edge[warp * n] = med;            // (each warp only overwrites its own edge, after using it)
    __syncthreads();
    if (tid == 0) {
        item_sum[it] = (red[0] + red[1]) + (red[2] + red[3]);
        item_edge[it] = (edge[0] + edge[n]) + (edge[2 * n] + edge[3 * n]);
    }
}

// stage 4: per star background = mean of the edge medians over the four edges and all frames; a_est per item
__global__ void k_phot_prep_guess(int F, int S, int n, float kk, const float* item_sum, const float* item_edge, float* a0) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    double acc = 0.0;
    for (int f = 0; f < F; ++f) acc += (double)item_edge[(size_t)f * S + s];
    double bg = acc / (4.0 * (double)F);
    if (!(bg == bg)) bg = 0.0;                      // nan_to_num
    if (isinf(bg)) bg = bg > 0 ? 3.4028234663852886e38 : -3.4028234663852886e38;
    for (int f = 0; f < F; ++f)
        a0[(size_t)f * S + s] = (float)(((double)item_sum[(size_t)f * S + s] - (double)(n * n) * bg) * (double)kk);
}

extern "C" {

int lcb_psf_prepare_batch(const lcb_psf_prepare_in* in, lcb_psf_prepare_out* out, void* stream) {
    LcbRange nvtx_range("lcb_psf_prepare_batch");
    LCB_REQUIRE(in && out, "lcb_psf_prepare_batch: NULL argument");
    LCB_REQUIRE(in->F >= 0 && in->n >= 1 && in->k >= 1 && in->star_off && in->image && in->noisemap,
                "lcb_psf_prepare_batch: bad input");
    LCB_REQUIRE(out->data && out->weight && out->a0 && out->x0 && out->y0, "lcb_psf_prepare_batch: NULL output array");
    LCB_REQUIRE(in->guess_method >= 0 && in->guess_method <= 2, "guess_method: 0 center, 1 max, 2 barycenter");
    if (in->F == 0) return LCB_OK;
    if (lcb_device_count() == 0) { lcb_set_error("no CUDA device: liblcb has no CPU fallback"); return LCB_ERR_CUDA; }
    cudaStream_t st = (cudaStream_t)stream;
    { LcbProfScope ps("k_psf_prepare", st); k_psf_prepare<<<in->F, PR_THREADS, 0, st>>>(*in, *out); }
    LCB_CUDA(cudaGetLastError());
    return LCB_OK;
}

// work: device scratch of lcb_phot_prepare_work_floats(F, S) floats
size_t lcb_phot_prepare_work_floats(int F, int S) { return (size_t)4 * F * S + 16; }

int lcb_phot_prepare_batch(const lcb_phot_prepare_in* in, lcb_phot_prepare_out* out, float* work, void* stream) {
    LcbRange nvtx_range("lcb_phot_prepare_batch");
    LCB_REQUIRE(in && out && work, "lcb_phot_prepare_batch: NULL argument");
    LCB_REQUIRE(in->F >= 1 && in->S >= 1 && in->n >= 2 && in->n <= 128 && in->k >= 1 && in->data && in->noisemap,
                "lcb_phot_prepare_batch: bad input (n <= 128)");
    LCB_REQUIRE(out->data && out->weight && out->a0 && out->scale, "lcb_phot_prepare_batch: NULL output array");
    if (lcb_device_count() == 0) { lcb_set_error("no CUDA device: liblcb has no CPU fallback"); return LCB_ERR_CUDA; }
    cudaStream_t st = (cudaStream_t)stream;
    const int B = in->F * in->S;
    float* item_max = work;
    int* item_bad = (int*)(work + B);
    float* item_sum = work + 2 * (size_t)B;
    float* item_edge = work + 3 * (size_t)B;
    const float kk = in->downsample_mean ? (float)(in->k * in->k) : 1.f;
    LcbProfScope ps("k_phot_prepare", st);
    k_phot_prep_max<<<B, PR_THREADS, 0, st>>>(*in, item_max, item_bad);
    k_phot_prep_scale<<<(in->S + 127) / 128, 128, 0, st>>>(in->F, in->S, item_max, out->scale);
    k_phot_prep_item<<<B, 128, (size_t)4 * in->n * sizeof(float), st>>>(*in, *out, out->scale, item_bad, item_sum, item_edge);
    k_phot_prep_guess<<<(in->S + 127) / 128, 128, 0, st>>>(in->F, in->S, in->n, kk, item_sum, item_edge, out->a0);
    LCB_CUDA(cudaGetLastError());
    return LCB_OK;
}

}  // extern "C"
