// lcb_api.cu -- library-wide state of liblcb: last error, conventions, staging arena, FP32 peak probe.
#include "lcb_common.cuh"
#include <nvtx3/nvToolsExt.h>
#include <cstdlib>

#include <mutex>
#include <vector>
#include <string>
#include <map>

static thread_local char g_err[512] = "";
// Conventions are PER HOST THREAD (every thread starts from the defaults below and lcb_conventions_set only touches the
// calling thread): engine.fan_out runs one host thread per GPU and each of them applies the conventions of its own call.
// Defaults: D_k is the block SUM -- lightcurver passes pixel sums as initial amplitudes and reads the fitted amplitudes back
// as fluxes (star_photometry.py:55-69,128; roi_modelling.py:199-212,462; notebook cells 13/21/36), see DESIGN.md section 2.
static thread_local lcb_conventions g_conv = {2.0f, 12, 0, 1, 1.0f, 0.99f, 0.9f, 0.999f, 1e-16f, 1e-16f};

void lcb_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

const lcb_conventions& lcb_conv() { return g_conv; }

DevConv lcb_devconv() {
    const lcb_conventions& c = g_conv;
    DevConv d;
    const double sig = (double)c.gauss_fwhm_up / (2.0 * sqrt(2.0 * log(2.0)));
    d.G = c.gauss_taps;
    d.inv2s2 = (float)(1.0 / (2.0 * sig * sig));
    d.invs2 = (float)(1.0 / (sig * sig));
    d.gnorm = (float)(1.0 / (sqrt(2.0 * M_PI) * sig));
    d.mean = c.downsample_mean;
    d.half = c.chi2_half ? 0.5f : 1.0f;
    d.clip = c.clip_global_norm;
    d.decay = c.lr_decay_rate;
    d.b1 = c.belief_b1; d.b2 = c.belief_b2; d.eps = c.belief_eps; d.eps_root = c.belief_eps_root;
    return d;
}

// Staging arenas of the LCB_MEM_HOST calls: a pool of arenas per device behind one mutex.  A call leases an arena of the
// current device for its duration (LcbArenaLease), so that concurrent host-pointer calls from several host threads -- on
// the same or on different devices -- never share or free each other's staging memory.
static std::mutex g_arena_mu;
static std::vector<LcbArena*> g_arena_free;

LcbArenaLease::LcbArenaLease() : a(nullptr) {
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess) { cudaGetLastError(); d = 0; }
    {
        std::lock_guard<std::mutex> lk(g_arena_mu);
        for (size_t i = 0; i < g_arena_free.size(); ++i)
            if (g_arena_free[i]->dev == d) { a = g_arena_free[i]; g_arena_free.erase(g_arena_free.begin() + i); break; }
    }
    if (!a) { a = new LcbArena(); a->dev = d; }
}
LcbArenaLease::~LcbArenaLease() {
    if (!a) return;
    a->rewind();
    std::lock_guard<std::mutex> lk(g_arena_mu);
    g_arena_free.push_back(a);
}

int LcbArena::reserve(size_t bytes) {
    int d = 0;
    LCB_CUDA(cudaGetDevice(&d));
    LCB_REQUIRE(dev < 0 || d == dev, "staging arena of device %d used while device %d is current", dev, d);
    if (base && cap >= bytes) return LCB_OK;
    if (base) { cudaFree(base); base = nullptr; cap = 0; }
    size_t want = bytes + (bytes >> 2) + (1u << 20);
    cudaError_t e = cudaMalloc((void**)&base, want);
    if (e != cudaSuccess) {
        lcb_set_error("arena cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
        base = nullptr;
        return LCB_ERR_NOMEM;
    }
    cap = want; dev = d; off = 0;
    return LCB_OK;
}

void* LcbArena::take(size_t bytes) {
    size_t a = (off + 255) & ~(size_t)255;
    if (a + bytes > cap) return nullptr;
    off = a + bytes;
    return base + a;
}

// ---------------------------------------------------------------- profiling
struct ProfRec { const char* name; cudaEvent_t e0, e1; };
static bool g_prof_on = false;
static std::vector<ProfRec> g_prof;
static std::mutex g_prof_mu;          // launches may come from several host threads (one per GPU)
bool lcb_profiling() { return g_prof_on; }

// NVTX ranges (the reference only logs wall-clock per step): opt-in with LCB_NVTX=1, one range per library
// call and one per kernel launch, visible to any NVTX-aware tool (nsys, ncu --nvtx)
static const bool g_nvtx = [] { const char* e = getenv("LCB_NVTX"); return e && e[0] == '1'; }();
LcbRange::LcbRange(const char* name) { if (g_nvtx) nvtxRangePushA(name); }
LcbRange::~LcbRange() { if (g_nvtx) nvtxRangePop(); }

LcbProfScope::LcbProfScope(const char* name, cudaStream_t s) : idx(-1), st(s) {
    if (g_nvtx) nvtxRangePushA(name);
    if (!g_prof_on) return;
    ProfRec r; r.name = name;
    if (cudaEventCreate(&r.e0) != cudaSuccess || cudaEventCreate(&r.e1) != cudaSuccess) return;
    cudaEventRecord(r.e0, st);
    e1 = r.e1;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof.push_back(r);
    idx = (int)g_prof.size() - 1;
}
LcbProfScope::~LcbProfScope() { if (idx >= 0) cudaEventRecord(e1, st); if (g_nvtx) nvtxRangePop(); }

extern "C" int lcb_profile_enable(int on) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    for (ProfRec& r : g_prof) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
    g_prof.clear();
    g_prof_on = (on != 0);
    return LCB_OK;
}

extern "C" int lcb_profile_summary(char* buf, int buflen) {
    LCB_REQUIRE(buf && buflen > 2, "lcb_profile_summary: bad buffer");
    std::map<std::string, std::pair<double, int>> agg;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    for (ProfRec& r : g_prof) {
        float ms = 0.f;
        if (cudaEventSynchronize(r.e1) == cudaSuccess && cudaEventElapsedTime(&ms, r.e0, r.e1) == cudaSuccess) {
            auto& a = agg[r.name]; a.first += ms; a.second += 1;
        }
    }
    std::string out = "{";
    bool first = true;
    for (auto& kv : agg) {
        char tmp[256];
        snprintf(tmp, sizeof(tmp), "%s\"%s\": {\"ms\": %.6f, \"launches\": %d}", first ? "" : ", ", kv.first.c_str(),
                 kv.second.first, kv.second.second);
        out += tmp; first = false;
    }
    out += "}";
    LCB_REQUIRE((int)out.size() + 1 <= buflen, "lcb_profile_summary: buffer too small (%zu needed)", out.size() + 1);
    memcpy(buf, out.c_str(), out.size() + 1);
    return LCB_OK;
}

extern "C" {

const char* lcb_last_error(void) { return g_err; }
int lcb_version(void) { return 100; }

int lcb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int lcb_conventions_get(lcb_conventions* out) {
    LCB_REQUIRE(out != nullptr, "lcb_conventions_get: NULL");
    *out = g_conv;
    return LCB_OK;
}

int lcb_conventions_set(const lcb_conventions* in) {
    LCB_REQUIRE(in != nullptr, "lcb_conventions_set: NULL");
    LCB_REQUIRE(in->gauss_taps == 8 || in->gauss_taps == 12 || in->gauss_taps == 16,
                "gauss_taps must be 8, 12 or 16 (got %d)", in->gauss_taps);
    LCB_REQUIRE(in->gauss_fwhm_up > 0.f, "gauss_fwhm_up must be > 0");
    g_conv = *in;
    return LCB_OK;
}

}  // extern "C"

// ---------------------------------------------------------------- FP32 FMA peak probe
// 8 independent FFMA chains per thread, 1024 threads/CTA, 2 CTAs/SM: measures the SIMT FP32 issue
// ceiling that the stencil kernels are bounded by.
__global__ void __launch_bounds__(1024) k_fp32_peak(int iters, float* sink) {
    float a0 = threadIdx.x * 1e-3f, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f;
    float a4 = a0 + 4.f, a5 = a0 + 5.f, a6 = a0 + 6.f, a7 = a0 + 7.f;
    const float m = 0.999f, c = 1e-3f;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            a0 = fmaf(a0, m, c); a1 = fmaf(a1, m, c); a2 = fmaf(a2, m, c); a3 = fmaf(a3, m, c);
            a4 = fmaf(a4, m, c); a5 = fmaf(a5, m, c); a6 = fmaf(a6, m, c); a7 = fmaf(a7, m, c);
        }
    }
    float s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (s == 123.456f) sink[0] = s;
}

// Same measurement with register-operand FFMAs in an 8x8 outer-product pattern (the shape of a stencil /
// SGEMM inner loop: acc[i][j] += a[i] * b[j], three register sources per FFMA, no immediates).
__global__ void __launch_bounds__(256) k_fp32_peak_rrr(int iters, const float* in, float* sink) {
    float a[8], b[8], acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = in[i + (threadIdx.x & 7)]; b[i] = in[8 + i + (threadIdx.x & 3)]; }
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) s += acc[i][j];
    if (s == 123.456f) sink[0] = s;
}

// The same probe with the packed FP32 instructions of sm_100 (FFMA2: one instruction, two FMAs per lane on a 64-bit register
// pair; __ffma2_rn).  8 independent FFMA2 chains per thread.  Tells whether a kernel that is bound by instruction ISSUE (the
// stencil kernels: DESIGN.md section 4) can halve the issue slots its arithmetic takes.
__global__ void __launch_bounds__(1024) k_fp32x2_peak(int iters, float* sink) {
    float2 a0 = make_float2(threadIdx.x * 1e-3f, 0.5f), a1 = a0, a2 = a0, a3 = a0, a4 = a0, a5 = a0, a6 = a0, a7 = a0;
    a1.x += 1.f; a2.x += 2.f; a3.x += 3.f; a4.x += 4.f; a5.x += 5.f; a6.x += 6.f; a7.x += 7.f;
    const float2 m = make_float2(0.999f, 0.998f), c = make_float2(1e-3f, 2e-3f);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            a0 = __ffma2_rn(a0, m, c); a1 = __ffma2_rn(a1, m, c); a2 = __ffma2_rn(a2, m, c); a3 = __ffma2_rn(a3, m, c);
            a4 = __ffma2_rn(a4, m, c); a5 = __ffma2_rn(a5, m, c); a6 = __ffma2_rn(a6, m, c); a7 = __ffma2_rn(a7, m, c);
        }
    }
    const float s = (a0.x + a1.x + a2.x + a3.x + a4.x + a5.x + a6.x + a7.x) + (a0.y + a1.y + a2.y + a3.y + a4.y + a5.y + a6.y + a7.y);
    if (s == 123.456f) sink[0] = s;
}

// mixed issue probe: per FFMA2 (or per pair of FFMAs when packed == 0) one conflict-free LDS.32 whose value feeds the chain --
// measures how many issue slots the packed form leaves for the loads of a stencil loop
template <int PACKED>
__global__ void __launch_bounds__(1024) k_fp32_mix(int iters, float* sink) {
    __shared__ float sm[2048];
    for (int i = threadIdx.x; i < 2048; i += 1024) sm[i] = 1e-3f * (float)i;
    __syncthreads();
    float2 a0 = make_float2(threadIdx.x * 1e-3f, 0.5f), a1 = a0, a2 = a0, a3 = a0;
    const float2 m = make_float2(0.999f, 0.998f);
    int idx = threadIdx.x;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const float v0 = sm[(idx + 32 * j) & 2047], v1 = sm[(idx + 32 * j + 512) & 2047];
            const float v2 = sm[(idx + 32 * j + 1024) & 2047], v3 = sm[(idx + 32 * j + 1536) & 2047];
            if (PACKED) {
                a0 = __ffma2_rn(a0, m, make_float2(v0, v0)); a1 = __ffma2_rn(a1, m, make_float2(v1, v1));
                a2 = __ffma2_rn(a2, m, make_float2(v2, v2)); a3 = __ffma2_rn(a3, m, make_float2(v3, v3));
            } else {
                a0.x = fmaf(a0.x, m.x, v0); a0.y = fmaf(a0.y, m.y, v0); a1.x = fmaf(a1.x, m.x, v1); a1.y = fmaf(a1.y, m.y, v1);
                a2.x = fmaf(a2.x, m.x, v2); a2.y = fmaf(a2.y, m.y, v2); a3.x = fmaf(a3.x, m.x, v3); a3.y = fmaf(a3.y, m.y, v3);
            }
        }
        idx += 7;
    }
    const float s = (a0.x + a1.x + a2.x + a3.x) + (a0.y + a1.y + a2.y + a3.y);
    if (s == 123.456f) sink[0] = s;
}

// which: 0 = FFMA2 chains, 1 = FFMA + LDS mix, 2 = FFMA2 + LDS mix.  tflops counts 2 flops per scalar FMA.
extern "C" int lcb_fp32x2_peak(int iters, int which, float* tflops, float* ms_out) {
    LCB_REQUIRE(iters > 0 && tflops != nullptr && which >= 0 && which <= 2, "lcb_fp32x2_peak: bad arguments");
    int dev = 0, sms = 0;
    LCB_CUDA(cudaGetDevice(&dev));
    LCB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    float* sink = nullptr;
    LCB_CUDA(cudaMalloc(&sink, 4));
    cudaEvent_t e0, e1;
    LCB_CUDA(cudaEventCreate(&e0));
    LCB_CUDA(cudaEventCreate(&e1));
    const int grid = sms * 2;
    auto launch = [&](int it) {
        if (which == 0) k_fp32x2_peak<<<grid, 1024>>>(it, sink);
        else if (which == 1) k_fp32_mix<0><<<grid, 1024>>>(it, sink);
        else k_fp32_mix<1><<<grid, 1024>>>(it, sink);
    };
    launch(iters / 8 + 1);
    LCB_CUDA(cudaEventRecord(e0));
    launch(iters);
    LCB_CUDA(cudaEventRecord(e1));
    LCB_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    LCB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    const double fma_per_thread = (which == 0 ? 2.0 * 8.0 : 2.0 * 4.0) * 16.0 * (double)iters;
    *tflops = (float)(2.0 * fma_per_thread * 1024.0 * (double)grid / (ms * 1e-3) / 1e12);
    if (ms_out) *ms_out = ms;
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(sink);
    return LCB_OK;
}

extern "C" int lcb_fp32_peak_rrr(int iters, float* tflops, float* ms_out) {
    LCB_REQUIRE(iters > 0 && tflops != nullptr, "lcb_fp32_peak_rrr: bad arguments");
    int dev = 0, sms = 0;
    LCB_CUDA(cudaGetDevice(&dev));
    LCB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    float *sink = nullptr, *in = nullptr;
    LCB_CUDA(cudaMalloc(&sink, 4));
    LCB_CUDA(cudaMalloc(&in, 64 * 4));
    float h[64];
    for (int i = 0; i < 64; ++i) h[i] = 1e-3f * (float)(i + 1);
    LCB_CUDA(cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice));
    cudaEvent_t e0, e1;
    LCB_CUDA(cudaEventCreate(&e0));
    LCB_CUDA(cudaEventCreate(&e1));
    const int grid = sms * 8;
    k_fp32_peak_rrr<<<grid, 256>>>(iters / 8 + 1, in, sink);
    LCB_CUDA(cudaEventRecord(e0));
    k_fp32_peak_rrr<<<grid, 256>>>(iters, in, sink);
    LCB_CUDA(cudaEventRecord(e1));
    LCB_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    LCB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    const double flops = 2.0 * 64.0 * (double)iters * 256.0 * (double)grid;
    *tflops = (float)(flops / (ms * 1e-3) / 1e12);
    if (ms_out) *ms_out = ms;
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(sink); cudaFree(in);
    return LCB_OK;
}

extern "C" int lcb_fp32_peak(int iters, float* tflops, float* ms_out) {
    LCB_REQUIRE(iters > 0 && tflops != nullptr, "lcb_fp32_peak: bad arguments");
    int dev = 0, sms = 0;
    LCB_CUDA(cudaGetDevice(&dev));
    LCB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    float* sink = nullptr;
    LCB_CUDA(cudaMalloc(&sink, 4));
    cudaEvent_t e0, e1;
    LCB_CUDA(cudaEventCreate(&e0));
    LCB_CUDA(cudaEventCreate(&e1));
    const int grid = sms * 2;
    k_fp32_peak<<<grid, 1024>>>(iters / 8 + 1, sink);   // warm-up
    LCB_CUDA(cudaEventRecord(e0));
    k_fp32_peak<<<grid, 1024>>>(iters, sink);
    LCB_CUDA(cudaEventRecord(e1));
    LCB_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    LCB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    const double flops = 2.0 * 8.0 * 16.0 * (double)iters * 1024.0 * (double)grid;
    *tflops = (float)(flops / (ms * 1e-3) / 1e12);
    if (ms_out) *ms_out = ms;
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(sink);
    return LCB_OK;
}

// Strided (pitched) copy between host and device: the in-process multi-GPU fan-out hands every device a slice data[:, lo:hi]
// of a C-contiguous pinned (F, S, n, n) array -- F rows of (hi - lo) n^2 floats with a pitch of S n^2 floats -- and this moves
// it in ONE asynchronous 2-D copy instead of a host-side gather into pageable memory followed by a staged upload.
extern "C" int lcb_copy_2d(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width_bytes, size_t height,
                           int to_device, void* stream) {
    LCB_REQUIRE(dst && src && dpitch >= width_bytes && spitch >= width_bytes, "lcb_copy_2d: bad arguments");
    if (width_bytes == 0 || height == 0) return LCB_OK;
    LCB_CUDA(cudaMemcpy2DAsync(dst, dpitch, src, spitch, width_bytes, height,
                               to_device ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    return LCB_OK;
}
