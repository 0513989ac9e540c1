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

template <int K, int G>
struct LcbPass {
    static constexpr int GE = G + K - 1;   // effective taps after folding the k-box
    static constexpr int OB = 4;           // outputs per thread-task along the contracted axis
    static constexpr int NR = GE + K * (OB - 1);
};

// pass 1: s (nu x nu, leading dim lds, row-major, shared or global) -> Vg, Vd stored [u][Y] (ld ldv)
template <int K, int G>
__device__ __forceinline__ void lcb_pass1(const float* __restrict__ s, int lds, int nu, int n, int icy,
                                          const float* __restrict__ ey_s, const float* __restrict__ dey_s,
                                          float* __restrict__ Vg, float* __restrict__ Vd, int ldv,
                                          int tid, int nthreads) {
    using P = LcbPass<K, G>;
    float ey[P::GE], dey[P::GE];
#pragma unroll
    for (int p = 0; p < P::GE; ++p) { ey[p] = ey_s[p]; dey[p] = dey_s[p]; }
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
            const float sv = (v >= 0 && v < nu) ? s[v * lds + u] : 0.f;
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
template <int K, int G, typename F>
__device__ __forceinline__ void lcb_pass2(const float* __restrict__ Vg, const float* __restrict__ Vd, int ldv,
                                          int nu, int n, int icx,
                                          const float* __restrict__ ex_s, const float* __restrict__ dex_s,
                                          int tid, int nthreads, F&& consume) {
    using P = LcbPass<K, G>;
    float ex[P::GE], dex[P::GE];
#pragma unroll
    for (int p = 0; p < P::GE; ++p) { ex[p] = ex_s[p]; dex[p] = dex_s[p]; }
    const int nxb = (n + P::OB - 1) / P::OB;
    for (int task = tid; task < n * nxb; task += nthreads) {
        const int Y = task % n;
        const int X0 = (task / n) * P::OB;
        const int ubase = K * X0 - icx - G / 2;
        float m0[P::OB], mx[P::OB], my[P::OB];
#pragma unroll
        for (int x = 0; x < P::OB; ++x) { m0[x] = 0.f; mx[x] = 0.f; my[x] = 0.f; }
#pragma unroll
        for (int r = 0; r < P::NR; ++r) {
            const int u = ubase + r;
            const bool ok = (u >= 0 && u < nu);
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
            if (X0 + x < n) consume(Y, X0 + x, m0[x], mx[x], my[x]);
    }
}

__device__ __forceinline__ int lcb_floordiv(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }

// Transposed (interpolating) FIR, the adjoint of the decimating passes.  Output index u is
// handled in the shifted frame u' = u + ic + G/2 and tasks are aligned on multiples of K*OB in u',
// so that the tap index p = u' - K*X = j - K*i is a compile-time constant for every (i, j).
//
// pass 2^T: r stored [X][Y] (ld ldr) -> Vbar[Y][u] (ld ldb):  Vbar[Y][u] = sum_X ex[u'-K X] r[Y][X]
// lanes <-> Y.
template <int K, int G>
__device__ __forceinline__ void lcb_pass2T(const float* __restrict__ rT, int ldr, int nu, int n, int icx,
                                           const float* __restrict__ ex_s, float* __restrict__ Vbar, int ldb,
                                           int tid, int nthreads) {
    using P = LcbPass<K, G>;
    constexpr int UB = K * P::OB;
    constexpr int ILO = -((P::GE - 1 + K - 1) / K);
    float ex[P::GE];
#pragma unroll
    for (int p = 0; p < P::GE; ++p) ex[p] = ex_s[p];
    const int off = icx + G / 2;
    const int bmin = lcb_floordiv(off, UB);
    const int nb = lcb_floordiv(nu - 1 + off, UB) - bmin + 1;
    for (int task = tid; task < n * nb; task += nthreads) {
        const int Y = task % n;
        const int U0 = UB * (bmin + task / n);
        const int XB0 = U0 / K;                      // exact: U0 is a multiple of K (may be negative)
        float acc[UB];
#pragma unroll
        for (int j = 0; j < UB; ++j) acc[j] = 0.f;
#pragma unroll
        for (int i = ILO; i < P::OB; ++i) {
            const int X = XB0 + i;
            const float rv = (X >= 0 && X < n) ? rT[X * ldr + Y] : 0.f;
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
}

// pass 1^T: Vbar[Y][u] (ld ldb) -> emit(v, u, sum_Y ey[v'-K Y] Vbar[Y][u]);  lanes <-> u.
template <int K, int G, typename F>
__device__ __forceinline__ void lcb_pass1T(const float* __restrict__ Vbar, int ldb, int nu, int n, int icy,
                                           const float* __restrict__ ey_s, int tid, int nthreads, F&& emit) {
    using P = LcbPass<K, G>;
    constexpr int UB = K * P::OB;
    constexpr int ILO = -((P::GE - 1 + K - 1) / K);
    float ey[P::GE];
#pragma unroll
    for (int p = 0; p < P::GE; ++p) ey[p] = ey_s[p];
    const int off = icy + G / 2;
    const int bmin = lcb_floordiv(off, UB);
    const int nb = lcb_floordiv(nu - 1 + off, UB) - bmin + 1;
    for (int task = tid; task < nu * nb; task += nthreads) {
        const int u = task % nu;
        const int V0 = UB * (bmin + task / nu);
        const int YB0 = V0 / K;
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
            if (v >= 0 && v < nu) emit(v, u, acc[j]);
        }
    }
}
