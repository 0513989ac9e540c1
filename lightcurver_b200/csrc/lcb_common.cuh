// lcb_common.cuh -- shared host/device helpers of liblcb (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdarg>
#include <cstring>
#include <cmath>
#include "../../include/lcb.h"

// ---------------------------------------------------------------- host side: errors, conventions
void lcb_set_error(const char* fmt, ...);
const lcb_conventions& lcb_conv();

#define LCB_CUDA(call)                                                                   \
    do {                                                                                 \
        cudaError_t e__ = (call);                                                        \
        if (e__ != cudaSuccess) {                                                        \
            lcb_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call,                  \
                          cudaGetErrorString(e__));                                      \
            return LCB_ERR_CUDA;                                                         \
        }                                                                                \
    } while (0)

#define LCB_REQUIRE(cond, ...)                                                           \
    do {                                                                                 \
        if (!(cond)) { lcb_set_error(__VA_ARGS__); return LCB_ERR_ARG; }                 \
    } while (0)

// Device staging arena for LCB_MEM_HOST calls: a growing allocation on ONE device; take() hands out 256-byte aligned
// slices, rewind() releases them (no free between calls).  Calls lease an arena of the current device from a process-wide
// pool for their duration (thread safe: two host threads never share an arena).
struct LcbArena {
    char* base = nullptr; size_t cap = 0; size_t off = 0; int dev = -1;
    int reserve(size_t bytes);
    void* take(size_t bytes);
    void rewind() { off = 0; }
};
struct LcbArenaLease {
    LcbArena* a;
    LcbArenaLease();
    ~LcbArenaLease();
    LcbArenaLease(const LcbArenaLease&) = delete;
    LcbArenaLease& operator=(const LcbArenaLease&) = delete;
};

// Optional per-kernel timing (lcb_profile_enable): CUDA events recorded on the launch stream around
// every kernel launch of the library; lcb_profile_summary() synchronises and aggregates by name.
struct LcbProfScope {
    int idx; cudaStream_t st; cudaEvent_t e1;
    LcbProfScope(const char* name, cudaStream_t s);
    ~LcbProfScope();
};
// NVTX range around a library call (LCB_NVTX=1; every LcbProfScope, i.e. every kernel launch, is a nested range as well)
struct LcbRange {
    LcbRange(const char* name);
    ~LcbRange();
};

// ---------------------------------------------------------------- device side
struct DevConv {              // conventions in the form kernels consume
    int G;                    // taps
    float inv2s2;             // 1/(2 sigma^2)
    float invs2;              // 1/sigma^2
    float gnorm;              // 1/(sqrt(2pi) sigma)
    int mean;                 // downsample mean
    float half;               // 0.5 or 1.0 factor on chi2
    float clip, decay, b1, b2, eps, eps_root;
};
DevConv lcb_devconv();

#define LCB_GE_MAX 20        // G + k - 1 <= 16 + 4 - 1

// Effective decimating taps for one axis (DESIGN.md "taps"): for shift c (upsampled px),
//   ic = floor(c + 0.5), fr = c - ic,
//   out[X] = sum_{p=0}^{GE-1} e[p] * in[k*X - ic - G/2 + p],   GE = G + k - 1
//   e[p]  = (1/k) sum_{kap<k} g(kap - (p-G/2) - fr) [tau = kap-(p-G/2) in (-G/2, G/2]]
//   de[p] = d e[p] / d c  (window held fixed)
__device__ __forceinline__ void lcb_tap(const DevConv& cv, int k, float fr, int pidx, float& e, float& de) {
    const int G2 = cv.G / 2;
    const int p = pidx - G2;
    float s = 0.f, d = 0.f;
    for (int kap = 0; kap < k; ++kap) {
        const int tau = kap - p;
        if (tau >= -G2 + 1 && tau <= G2) {
            const float x = (float)tau - fr;
            const float g = cv.gnorm * expf(-x * x * cv.inv2s2);
            s += g;
            d += x * cv.invs2 * g;
        }
    }
    const float sc = cv.mean ? 1.f / (float)k : 1.f;
    e = s * sc;
    de = d * sc;
}

// ---------------------------------------------------------------- bulk asynchronous copies (TMA, 1-D form) + mbarrier
// cp.async.bulk moves a contiguous, 16-byte aligned block global -> shared through the TMA unit without occupying the
// issuing warps; completion is signalled on an mbarrier in shared memory (transaction-byte count).  One thread arms the
// barrier (expect_tx) and issues the copies; every consumer waits on the barrier's phase parity.
__device__ __forceinline__ unsigned lcb_smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void lcb_mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(lcb_smem_addr(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");      // visible to the async proxy
}

__device__ __forceinline__ void lcb_mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(lcb_smem_addr(bar)), "r"(bytes) : "memory");
}

// dst (shared, 16 B aligned), src (global, 16 B aligned), bytes (multiple of 16)
__device__ __forceinline__ void lcb_bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(lcb_smem_addr(dst)), "l"(src), "r"(bytes), "r"(lcb_smem_addr(bar)) : "memory");
}

__device__ __forceinline__ void lcb_mbar_wait(unsigned long long* bar, unsigned parity) {
    unsigned done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(lcb_smem_addr(bar)), "r"(parity) : "memory");
    }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// AdaBelief (optax.scale_by_belief) on one scalar parameter; g already clipped.
struct BeliefCoef { float lr, b1, b2, omb1, omb2, inv_bc1, inv_bc2, eps, eps_root; };

__device__ __forceinline__ void belief_update(const BeliefCoef& c, float g, float& p, float& mu, float& nu) {
    mu = c.b1 * mu + c.omb1 * g;
    const float d = g - mu;
    nu = c.b2 * nu + c.omb2 * d * d + c.eps_root;
    p -= c.lr * (mu * c.inv_bc1) / (sqrtf(nu * c.inv_bc2) + c.eps);
}
