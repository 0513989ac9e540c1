// lcb_passes.cuh -- the separable "shift, smooth, decimate" passes shared by K1 (PSF fit) and K2
// (photometry), and their transposes.  FP32 SIMT, shared-memory resident, register-tiled.
//
// Model of one star:  M[Y][X] = sum_{v,u} ey[v - (kY-icy-G/2)] ex[u - (kX-icx-G/2)] s[v][u]
// evaluated as two 1-D decimating FIR passes that TRANSPOSE between them so that both are
// bank-conflict free with one 4-byte shared load per (GE*2..3)/k FMAs:
//   pass 1  lanes <-> input column u : Vg/Vd[u][Y] = sum_p {ey,dey}[p] s[kY-icy-G/2+p][u]
//   pass 2  lanes <-> output row   Y : M0,Mx,My[Y][X] = sum_p {ex,dex}[p] V{g,d}[kX-icx-G/2+p][Y]
// Taps live in registers (compile-time GE), each loaded value feeds YB (XB) outputs x 2 (3) kernels.
#pragma once
#include "lcb_common.cuh"

// Taps are broadcast reads (every lane the same address): fetch them as 16-byte vectors, one shared-memory
// wavefront per four taps instead of one per tap.  Tap arrays are 16-byte aligned and LCB_GE_MAX long.
template <int GE>
__device__ __forceinline__ void lcb_load_taps(const float* __restrict__ src, float (&dst)[GE]) {
    constexpr int NV = (GE + 3) / 4;
    static_assert(NV * 4 <= LCB_GE_MAX, "tap arrays are LCB_GE_MAX floats");
    const float4* s4 = reinterpret_cast<const float4*>(src);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const float4 t = s4[i];
        if (4 * i + 0 < GE) dst[4 * i + 0] = t.x;
        if (4 * i + 1 < GE) dst[4 * i + 1] = t.y;
        if (4 * i + 2 < GE) dst[4 * i + 2] = t.z;
        if (4 * i + 3 < GE) dst[4 * i + 3] = t.w;
    }
}

template <int K, int G, int OBV = 4>
struct LcbPass {
    static constexpr int GE = G + K - 1;   // effective taps after folding the k-box
    static constexpr int OB = OBV;         // outputs per thread-task along the contracted axis
    static constexpr int NR = GE + K * (OB - 1);
};

// pass 1: s (nu x nu, leading dim lds, row-major, shared or global) -> Vg, Vd stored [u][Y] (ld ldv)
// HALO = true: the caller guarantees that every input index touched lies inside zero-filled halos
// (no bounds predicates, loads use immediate offsets from one base address).
template <int K, int G, bool HALO = false, int OBV = 4>
__device__ __forceinline__ void lcb_pass1(const float* __restrict__ s, int lds, int nu, int n, int icy,
                                          const float* __restrict__ ey_s, const float* __restrict__ dey_s,
                                          float* __restrict__ Vg, float* __restrict__ Vd, int ldv,
                                          int tid, int nthreads) {
    using P = LcbPass<K, G, OBV>;
    float ey[P::GE], dey[P::GE];
    lcb_load_taps<P::GE>(ey_s, ey);
    lcb_load_taps<P::GE>(dey_s, dey);
    const int nyb = (n + P::OB - 1) / P::OB;
    for (int task = tid; task < nu * nyb; task += nthreads) {
        const int u = task % nu;
        const int Y0 = (task / nu) * P::OB;
        const int vbase = K * Y0 - icy - G / 2;
        float ag[P::OB], ad[P::OB];
#pragma unroll
        for (int y = 0; y < P::OB; ++y) { ag[y] = 0.f; ad[y] = 0.f; }
#pragma unroll
        for (int r = 0; r < P::NR; ++r) {
            const int v = vbase + r;
            const float sv = (HALO || (v >= 0 && v < nu)) ? s[v * lds + u] : 0.f;
#pragma unroll
            for (int y = 0; y < P::OB; ++y) {
                const int p = r - K * y;
                if (p >= 0 && p < P::GE) { ag[y] = fmaf(ey[p], sv, ag[y]); ad[y] = fmaf(dey[p], sv, ad[y]); }
            }
        }
#pragma unroll
        for (int y = 0; y < P::OB; ++y)
            if (Y0 + y < n) { Vg[u * ldv + Y0 + y] = ag[y]; Vd[u * ldv + Y0 + y] = ad[y]; }
    }
}

// pass 2: V{g,d} [u][Y] -> (M0, Mx, My)[Y][X] handed to consume(Y, X, m0, mx, my).
//   m0 = sum ex*Vg, mx = sum dex*Vg (d/dcx), my = sum ex*Vd (d/dcy); c in upsampled px.
// `aux0`, `aux1` (may be NULL): two per-pixel arrays stored [X][Y] with leading dimension ldaux (stamp and
// weight); their values for the task's outputs are fetched BEFORE the FMA loop so that the global/L2 latency
// hides behind it, and handed to consume(Y, X, m0, mx, my, aux0[X][Y], aux1[X][Y]).
template <int K, int G, int OBV = 4, bool HALO = false, typename F>
__device__ __forceinline__ void lcb_pass2(const float* __restrict__ Vg, const float* __restrict__ Vd, int ldv,
                                          int nu, int n, int icx,
                                          const float* __restrict__ ex_s, const float* __restrict__ dex_s,
                                          const float* __restrict__ aux0, const float* __restrict__ aux1, int ldaux,
                                          int tid, int nthreads, F&& consume) {
    using P = LcbPass<K, G, OBV>;
    float ex[P::GE], dex[P::GE];
    lcb_load_taps<P::GE>(ex_s, ex);
    lcb_load_taps<P::GE>(dex_s, dex);
    const int nxb = (n + P::OB - 1) / P::OB;
    for (int task = tid; task < n * nxb; task += nthreads) {
        const int Y = task % n;
        const int X0 = (task / n) * P::OB;
        const int ubase = K * X0 - icx - G / 2;
        float pa[P::OB], pb[P::OB];
#pragma unroll
        for (int x = 0; x < P::OB; ++x) {
            const bool ok = (X0 + x < n);
            pa[x] = (ok && aux0) ? aux0[(X0 + x) * ldaux + Y] : 0.f;
            pb[x] = (ok && aux1) ? aux1[(X0 + x) * ldaux + Y] : 0.f;
        }
        float m0[P::OB], mx[P::OB], my[P::OB];
#pragma unroll
        for (int x = 0; x < P::OB; ++x) { m0[x] = 0.f; mx[x] = 0.f; my[x] = 0.f; }
#pragma unroll
        for (int r = 0; r < P::NR; ++r) {
            const int u = ubase + r;
            const bool ok = HALO || (u >= 0 && u < nu);
            const float vg = ok ? Vg[u * ldv + Y] : 0.f;
            const float vd = ok ? Vd[u * ldv + Y] : 0.f;
#pragma unroll
            for (int x = 0; x < P::OB; ++x) {
                const int p = r - K * x;
                if (p >= 0 && p < P::GE) {
                    m0[x] = fmaf(ex[p], vg, m0[x]);
                    mx[x] = fmaf(dex[p], vg, mx[x]);
                    my[x] = fmaf(ex[p], vd, my[x]);
                }
            }
        }
#pragma unroll
        for (int x = 0; x < P::OB; ++x)
            if (X0 + x < n) consume(Y, X0 + x, m0[x], mx[x], my[x], pa[x], pb[x]);
    }
}

__device__ __forceinline__ int lcb_floordiv(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }

// Transposed (interpolating) FIR, the adjoint of the decimating passes.  Output index u is
// handled in the shifted frame u' = u + ic + G/2 and tasks are aligned on multiples of K*OB in u',
// so that the tap index p = u' - K*X = j - K*i is a compile-time constant for every (i, j).
//
// pass 2^T: r stored [X][Y] (ld ldr) -> Vbar[Y][u] (ld ldb):  Vbar[Y][u] = sum_X ex[u'-K X] r[Y][X]
// lanes <-> Y.
template <int K, int G, int OBV = 4, bool HALO = false>
__device__ __forceinline__ void lcb_pass2T(const float* __restrict__ rT, int ldr, int nu, int n, int icx,
                                           const float* __restrict__ ex_s, float* __restrict__ Vbar, int ldb,
                                           int tid, int nthreads) {
    using P = LcbPass<K, G, OBV>;
    constexpr int UB = K * P::OB;
    constexpr int ILO = -((P::GE - 1 + K - 1) / K);
    float ex[P::GE];
    lcb_load_taps<P::GE>(ex_s, ex);
    const int off = icx + G / 2;
    const int S0 = K * lcb_floordiv(off, K);            // block grid origin: <= off, multiple of K
    const int nb = (nu + UB - 1) / UB;
    for (int task = tid; task < n * nb; task += nthreads) {
        const int Y = task % n;
        const int U0 = S0 + UB * (task / n);
        const int XB0 = lcb_floordiv(U0, K);
        float acc[UB];
#pragma unroll
        for (int j = 0; j < UB; ++j) acc[j] = 0.f;
#pragma unroll
        for (int i = ILO; i < P::OB; ++i) {
            const int X = XB0 + i;
            const float rv = (HALO || (X >= 0 && X < n)) ? rT[X * ldr + Y] : 0.f;
#pragma unroll
            for (int j = 0; j < UB; ++j) {
                const int p = j - K * i;
                if (p >= 0 && p < P::GE) acc[j] = fmaf(ex[p], rv, acc[j]);
            }
        }
#pragma unroll
        for (int j = 0; j < UB; ++j) {
            const int u = U0 + j - off;
            if (u >= 0 && u < nu) Vbar[Y * ldb + u] = acc[j];
        }
    }
    if (nu % UB == 0 && K > 1) {                        // trailing columns (see lcb_pass1T)
        const int nt = off - S0;
        const int Xb = (S0 + nb * UB) / K;
        for (int task = tid; task < n * nt; task += nthreads) {
            const int Y = task % n, t = task / n;
            float acc = 0.f;
#pragma unroll
            for (int tt = 0; tt < K - 1; ++tt) {
                if (tt == t) {
#pragma unroll
                    for (int q = 0; q < (P::GE + K - 1) / K; ++q) {
                        const int p = tt + K * q, X = Xb - q;
                        if (p < P::GE && (HALO || (X >= 0 && X < n))) acc = fmaf(ex[p], rT[X * ldr + Y], acc);
                    }
                }
            }
            Vbar[Y * ldb + nu - nt + t] = acc;
        }
    }
}

// pass 1^T: Vbar[Y][u] (ld ldb) -> emit(v, u, sum_Y ey[v'-K Y] Vbar[Y][u]);  lanes <-> u.
// The block grid starts at V0 = K*floor(off/K) (any multiple of K keeps p = j - K*i static), so
// nu/UB blocks cover all but at most K-1 trailing rows; those are finished by a short second loop
// instead of a whole extra block per column (keeps the task count a multiple of the CTA size).
template <int K, int G, bool HALO = false, int OBV = 4, typename F>
__device__ __forceinline__ void lcb_pass1T(const float* __restrict__ Vbar, int ldb, int nu, int n, int icy,
                                           const float* __restrict__ ey_s, int tid, int nthreads, F&& emit) {
    using P = LcbPass<K, G, OBV>;
    constexpr int UB = K * P::OB;
    constexpr int ILO = -((P::GE - 1 + K - 1) / K);
    float ey[P::GE];
    lcb_load_taps<P::GE>(ey_s, ey);
    const int off = icy + G / 2;
    const int S0 = K * lcb_floordiv(off, K);            // <= off, multiple of K
    const int nb = (nu + UB - 1) / UB;                  // blocks covering v' in [S0, S0 + nb*UB)
    for (int task = tid; task < nu * nb; task += nthreads) {
        const int u = task % nu;
        const int V0 = S0 + UB * (task / nu);
        const int YB0 = lcb_floordiv(V0, K);
        float acc[UB];
#pragma unroll
        for (int j = 0; j < UB; ++j) acc[j] = 0.f;
#pragma unroll
        for (int i = ILO; i < P::OB; ++i) {
            const int Y = YB0 + i;
            const float bv = (HALO || (Y >= 0 && Y < n)) ? Vbar[Y * ldb + u] : 0.f;
#pragma unroll
            for (int j = 0; j < UB; ++j) {
                const int p = j - K * i;
                if (p >= 0 && p < P::GE) acc[j] = fmaf(ey[p], bv, acc[j]);
            }
        }
#pragma unroll
        for (int j = 0; j < UB; ++j) {
            const int v = V0 + j - off;
            if (v >= 0 && v < nu) emit(v, u, acc[j]);
        }
    }
    // trailing rows v' = S0 + nb*UB + t, t < K-1 (only when nu % UB == 0): v' - t is a multiple of K, so
    // the tap index p = t + K*q stays a compile-time constant.
    if (nu % UB == 0 && K > 1) {
        const int nt = off - S0;                        // 0 .. K-1 uncovered rows
        const int Yb = (S0 + nb * UB) / K;              // exact
        for (int task = tid; task < nu * nt; task += nthreads) {
            const int u = task % nu, t = task / nu;
            float acc = 0.f;
#pragma unroll
            for (int tt = 0; tt < K - 1; ++tt) {
                if (tt == t) {
#pragma unroll
                    for (int q = 0; q < (P::GE + K - 1) / K; ++q) {
                        const int p = tt + K * q, Y = Yb - q;
                        if (p < P::GE && (HALO || (Y >= 0 && Y < n))) acc = fmaf(ey[p], Vbar[Y * ldb + u], acc);
                    }
                }
            }
            emit(nu - nt + t, u, acc);
        }
    }
}


// pass 1^T accumulating into a plane that lives in GLOBAL memory (grids too large for shared memory):
// Gp[v][u] += a * sum_Y ey[v'-K Y] Vbar[Y][u].  Same blocking as lcb_pass1T, but the UB values of Gp a task updates
// are loaded together BEFORE the first store -- one L2 round trip per task instead of UB dependent ones.
template <int K, int G, int OBV = 4>
__device__ __forceinline__ void lcb_pass1T_rmw(const float* __restrict__ Vbar, int ldb, int nu, int n, int icy,
                                               const float* __restrict__ ey_s, float a, float* __restrict__ Gp,
                                               int tid, int nthreads) {
    using P = LcbPass<K, G, OBV>;
    constexpr int UB = K * P::OB;
    constexpr int ILO = -((P::GE - 1 + K - 1) / K);
    float ey[P::GE];
    lcb_load_taps<P::GE>(ey_s, ey);
    const int off = icy + G / 2;
    const int S0 = K * lcb_floordiv(off, K);
    const int nb = (nu + K - 1 + UB - 1) / UB;          // covers every row (no trailing loop)
    for (int task = tid; task < nu * nb; task += nthreads) {
        const int u = task % nu;
        const int V0 = S0 + UB * (task / nu);
        const int YB0 = lcb_floordiv(V0, K);
        float old[UB];
#pragma unroll
        for (int j = 0; j < UB; ++j) {
            const int v = V0 + j - off;
            old[j] = (v >= 0 && v < nu) ? Gp[v * nu + u] : 0.f;
        }
        float acc[UB];
#pragma unroll
        for (int j = 0; j < UB; ++j) acc[j] = 0.f;
#pragma unroll
        for (int i = ILO; i < P::OB; ++i) {
            const int Y = YB0 + i;
            const float bv = (Y >= 0 && Y < n) ? Vbar[Y * ldb + u] : 0.f;
#pragma unroll
            for (int j = 0; j < UB; ++j) {
                const int p = j - K * i;
                if (p >= 0 && p < P::GE) acc[j] = fmaf(ey[p], bv, acc[j]);
            }
        }
#pragma unroll
        for (int j = 0; j < UB; ++j) {
            const int v = V0 + j - off;
            if (v >= 0 && v < nu) Gp[v * nu + u] = fmaf(a, acc[j], old[j]);
        }
    }
}
