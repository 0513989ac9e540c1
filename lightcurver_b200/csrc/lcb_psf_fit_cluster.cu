// lcb_psf_fit_cluster.cu -- K1c: the per-frame PSF pixel-grid fit (stage 2 of starred build_psf,
// lightcurver/processes/psf_modelling.py:164-171) for grids that do not fit the shared memory of one SM
// (BASELINE cfg5: 64x64 stamps, subsampling 3 -> 192x192 grid, 30 stars).
//
// One thread-block CLUSTER of 8 CTAs per frame.  CTA c owns a band of n/8 stamp rows = K n/8 grid rows of every
// plane (s, b, gradient, starlet scratch, sign planes) in its own shared memory:
//   * star passes: the vertical decimating pass reads the band of s plus HB halo rows that are refreshed from the two
//     neighbouring CTAs through distributed shared memory once per iteration; the horizontal passes are local; the
//     transposed vertical pass accumulates into a gradient band WITH halo rows, and the halos are folded into the
//     neighbours' bands once per iteration (fixed order: deterministic);
//   * per-star sums (chi2, d/da, d/dx0, d/dy0) are pushed to every CTA and summed in rank order, so that every CTA
//     updates an identical copy of the star parameters;
//   * starlet regulariser: rows local, columns across the cluster through DSMEM (same scheme as k_deconv_starlet_dsm),
//     sign(alpha_j) kept as int8 bands, t_j = lambda_j W_j sign(alpha_j) rebuilt from W (L2);
//   * AdaBelief moments live in the L2-resident workspace (touched once per iteration, coalesced).
// Same mathematics and conventions as k_psf_fit (lcb_psf_fit.cu); parity is tested against it and against the oracle.
// Star shifts are limited to +-(HB - G/2 - 1)/K stamp pixels (2.3 px at K = 3), the reach of the halo rows.
#include "lcb_psf.cuh"
#include <cooperative_groups.h>

namespace cg = cooperative_groups;

#define CL_CTAS 8
#define CL_THREADS 512
#define CL_WARPS (CL_THREADS / 32)
#define CL_HB 16              // halo rows (grid) above and below a band; also halo rows of Vg/Vd along u

// ---------------------------------------------------------------- banded pass variants (see lcb_passes.cuh)
// pass 2 on a band: V{g,d}[u][Yl] (zero halo rows in u) -> consume(Yl, X, m0, mx, my, aux0[X][Yl], aux1[X][Yl])
template <int K, int G, int OBV, typename F>
__device__ __forceinline__ void cl_pass2(const float* __restrict__ Vg, const float* __restrict__ Vd, int ldv,
                                         int nrows, int ncols, int icx,
                                         const float* __restrict__ ex_s, const float* __restrict__ dex_s,
                                         const float* __restrict__ aux0, const float* __restrict__ aux1, int ldaux,
                                         int tid, int nthreads, F&& consume) {
    using P = LcbPass<K, G, OBV>;
    float ex[P::GE], dex[P::GE];
    lcb_load_taps<P::GE>(ex_s, ex);
    lcb_load_taps<P::GE>(dex_s, dex);
    const int nxb = (ncols + P::OB - 1) / P::OB;
    for (int task = tid; task < nrows * nxb; task += nthreads) {
        const int Y = task % nrows;
        const int X0 = (task / nrows) * P::OB;
        const int ubase = K * X0 - icx - G / 2;
        float pa[P::OB], pb[P::OB];
#pragma unroll
        for (int x = 0; x < P::OB; ++x) {
            const bool ok = (X0 + x < ncols);
            pa[x] = ok ? aux0[(X0 + x) * ldaux + Y] : 0.f;
            pb[x] = ok ? aux1[(X0 + x) * ldaux + Y] : 0.f;
        }
        float m0[P::OB], mx[P::OB], my[P::OB];
#pragma unroll
        for (int x = 0; x < P::OB; ++x) { m0[x] = 0.f; mx[x] = 0.f; my[x] = 0.f; }
#pragma unroll
        for (int r = 0; r < P::NR; ++r) {
            const int u = ubase + r;
            const float vg = Vg[u * ldv + Y];
            const float vd = Vd[u * ldv + Y];
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
            if (X0 + x < ncols) consume(Y, X0 + x, m0[x], mx[x], my[x], pa[x], pb[x]);
    }
}

// pass 2^T on a band: r [X][Yl] -> Vbar[Yl][u], u in [0, nu)
template <int K, int G, int OBV>
__device__ __forceinline__ void cl_pass2T(const float* __restrict__ rT, int ldr, int nu, int nrows, int ncols, int icx,
                                          const float* __restrict__ ex_s, float* __restrict__ Vbar, int ldb,
                                          int tid, int nthreads) {
    using P = LcbPass<K, G, OBV>;
    constexpr int UB = K * P::OB;
    constexpr int ILO = -((P::GE - 1 + K - 1) / K);
    float ex[P::GE];
    lcb_load_taps<P::GE>(ex_s, ex);
    const int off = icx + G / 2;
    const int S0 = K * lcb_floordiv(off, K);
    const int nb = (nu + K - 1 + UB - 1) / UB;
    for (int task = tid; task < nrows * nb; task += nthreads) {
        const int Y = task % nrows;
        const int U0 = S0 + UB * (task / nrows);
        const int XB0 = lcb_floordiv(U0, K);
        float acc[UB];
#pragma unroll
        for (int j = 0; j < UB; ++j) acc[j] = 0.f;
#pragma unroll
        for (int i = ILO; i < P::OB; ++i) {
            const int X = XB0 + i;
            const float rv = (X >= 0 && X < ncols) ? rT[X * ldr + Y] : 0.f;
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

// pass 1^T on a band: Vbar[Yl][u] (Yl in [0, nrows_in)) -> emit(vv, u, .) for vv = v_local + vshift in [0, nrows_out)
// (v_local = grid row relative to the first row of the band)
template <int K, int G, int OBV, typename F>
__device__ __forceinline__ void cl_pass1T(const float* __restrict__ Vbar, int ldb, int nu, int nrows_in, int nrows_out,
                                          int vshift, int icy, const float* __restrict__ ey_s, int tid, int nthreads, F&& emit) {
    using P = LcbPass<K, G, OBV>;
    constexpr int UB = K * P::OB;
    constexpr int ILO = -((P::GE - 1 + K - 1) / K);
    float ey[P::GE];
    lcb_load_taps<P::GE>(ey_s, ey);
    const int off = icy + G / 2 - vshift;               // vv = v' - off,  v' = K Y + p
    const int S0 = K * lcb_floordiv(off, K);
    const int nb = (nrows_out + K - 1 + UB - 1) / UB;
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
            const float bv = (Y >= 0 && Y < nrows_in) ? Vbar[Y * ldb + u] : 0.f;
#pragma unroll
            for (int j = 0; j < UB; ++j) {
                const int p = j - K * i;
                if (p >= 0 && p < P::GE) acc[j] = fmaf(ey[p], bv, acc[j]);
            }
        }
#pragma unroll
        for (int j = 0; j < UB; ++j) {
            const int vv = V0 + j - off;
            if (vv >= 0 && vv < nrows_out) emit(vv, u, acc[j]);
        }
    }
}

template <int NV>
__device__ __forceinline__ void cl_block_reduce(float (&v)[NV], float* red /* [CL_WARPS][NV] */, int tid) {
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = warp_sum(v[i]);
    __syncthreads();
    if ((tid & 31) == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) red[(tid >> 5) * NV + i] = v[i];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < CL_WARPS; ++w) s += red[w * NV + i];
        v[i] = s;
    }
}

// shared-memory layout (floats), shared by host and device
struct ClLayout { int NB, RB, ldv, oTaps, oSp, oRed, oXch, oBp4, oS, oGR, oB, oC0, oSg, oScr, scr, total; };
__host__ __device__ inline ClLayout cl_layout(int n, int k, int Nmax, int J) {
    ClLayout L;
    L.NB = n / CL_CTAS; L.RB = L.NB * k;
    const int nu = n * k;
    L.ldv = L.NB + 1;
    int o = 0;
    L.oTaps = o; o += Nmax * 4 * LCB_GE_MAX;
    L.oSp = o; o += Nmax * 12;
    L.oRed = o; o += CL_WARPS * 4;
    L.oXch = o; o += CL_CTAS * (Nmax * 4 + 4);          // per-rank partial sums, pushed by every CTA
    L.oBp4 = o; o += 4 * nu;                            // per-band partial column sums for the folded border taps
    o = (o + 3) & ~3;
    L.oS = o; o += (L.RB + 2 * CL_HB) * nu;             // s with halo rows
    L.oGR = o; o += (L.RB + 2 * CL_HB) * nu;            // d chi2 / d s with halo rows
    L.oB = o; o += L.RB * nu;
    L.oC0 = o; o += L.RB * nu;
    L.oSg = o; o += (J * L.RB * nu + 3) / 4;            // int8 sign bands
    // star-pass scratch: Vg, Vd [(nu + 2 HB)][ldv], rT [n][ldv], Vbar [NB][nu + 1]; aliased by the starlet planes X, Q
    const int pass_scr = 2 * (nu + 2 * CL_HB) * L.ldv + n * L.ldv + L.NB * (nu + 1);
    const int star_scr = 2 * L.RB * nu;
    L.scr = pass_scr > star_scr ? pass_scr : star_scr;
    L.oScr = o; o += L.scr;
    L.total = o;
    return L;
}

template <int K, int G>
__global__ void __cluster_dims__(CL_CTAS, 1, 1) __launch_bounds__(CL_THREADS, 1) k_psf_fit_cl(PsfArgs A) {
    using P = LcbPass<K, G>;
    constexpr int NT = CL_THREADS;
    constexpr int HB = CL_HB;
    extern __shared__ __align__(16) float sm[];
    cg::cluster_group cl = cg::this_cluster();
    const int crank = (int)cl.block_rank();
    const int f = blockIdx.x / CL_CTAS;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = A.n, nu = A.nu, nn = n * n, pp = nu * nu, J = A.J;
    const ClLayout L = cl_layout(n, K, A.Nmax, J);
    const int NB = L.NB, RB = L.RB, ldv = L.ldv, ldb = nu + 1;
    const int Y0 = crank * NB, v0 = crank * RB;           // first stamp row / grid row of the band
    const int bsz = RB * nu;
    const int i0 = A.star_off[f], N = A.star_off[f + 1] - i0;
    const DevConv cv = A.cv;
    const float fk = (float)K;

    float* taps = sm + L.oTaps;                           // [Nmax][4][LCB_GE_MAX]
    float* sp = sm + L.oSp;                               // [Nmax][12] a,x0,y0, mu3, nu3, g3 (replicated in every CTA)
    float* red = sm + L.oRed;
    float* xch = sm + L.oXch;                             // [CL_CTAS][Nmax*4 + 4]
    float* bp4 = sm + L.oBp4;                             // [top P1, top P2, bottom P1, bottom P2][nu]
    const int xsz = A.Nmax * 4 + 4;
    float* Sh = sm + L.oS;                                // [HB + RB + HB][nu]
    float* S = Sh + HB * nu;                              // own rows
    float* GRh = sm + L.oGR;                              // [HB + RB + HB][nu]
    float* GR = GRh + HB * nu;
    float* Bp = sm + L.oB;
    float* C0 = sm + L.oC0;
    signed char* sg = reinterpret_cast<signed char*>(sm + L.oSg);
    float* scr = sm + L.oScr;
    float* Vg = scr + HB * ldv;                           // [HB + nu + HB][ldv]
    float* Vd = Vg + (nu + 2 * HB) * ldv;
    float* rT = Vd + (nu + HB) * ldv;                     // [n][ldv]
    float* Vbar = rT + n * ldv;                           // [NB][ldb]
    float* Xp = scr;                                      // starlet planes alias the star-pass scratch
    float* Qp = scr + bsz;
    __shared__ float* peer[CL_CTAS];
    if (tid < CL_CTAS) peer[tid] = (float*)cl.map_shared_rank(sm, tid);

    float* MU = A.work + (size_t)f * A.work_per_frame + v0 * nu;      // moments of the band, L2 resident
    float* NU = A.work + (size_t)f * A.work_per_frame + pp + v0 * nu;
    float* dT = A.work + (size_t)f * A.work_per_frame + (size_t)(J + 7) * pp;   // stamps transposed [st][X][Y]
    float* wT = dT + (size_t)A.Nmax * nn;
    const float* sfix = A.s_fixed + (size_t)f * pp + v0 * nu;
    const float* Wf = A.W ? A.W + (size_t)f * J * pp + v0 * nu : nullptr;
    const float* dat = A.data + (size_t)i0 * nn;
    const float* wgt = A.weight + (size_t)i0 * nn;

    // ---- initial state
    for (int i = tid; i < (RB + 2 * HB) * nu; i += NT) { Sh[i] = 0.f; GRh[i] = 0.f; }
    for (int i = tid; i < L.scr; i += NT) scr[i] = 0.f;
    for (int i = tid; i < N * NB * n; i += NT) {          // own rows of every stamp, transposed
        const int st = i / (NB * n), r = i % (NB * n), Yl = r / n, X = r % n;
        dT[(size_t)st * nn + X * n + Y0 + Yl] = dat[(size_t)st * nn + (Y0 + Yl) * n + X];
        wT[(size_t)st * nn + X * n + Y0 + Yl] = wgt[(size_t)st * nn + (Y0 + Yl) * n + X];
    }
    __syncthreads();
    for (int i = tid; i < bsz; i += NT) {
        const float b = A.b[(size_t)f * pp + v0 * nu + i];
        Bp[i] = b; MU[i] = 0.f; NU[i] = 0.f; S[i] = sfix[i] + b;
    }
    for (int i = tid; i < N * 12; i += NT) {
        const int st = i / 12, c = i % 12;
        sp[i] = (c == 0) ? A.a[i0 + st] : (c == 1) ? A.x0[i0 + st] : (c == 2) ? A.y0[i0 + st] : 0.f;
    }
    cl.sync();

    auto pl = [&](int c, int off) -> float* { return peer[c] + off; };
    float b1t = 1.f, b2t = 1.f;
    int bad = 0;
    const float sc = (cv.half == 0.5f) ? 1.f : 2.f;
    const float lim = fminf(0.25f * (float)n, (float)(HB - G / 2 - 1) / fk);
    const bool do_reg = (A.lam_scales != 0.f || A.lam_hf != 0.f);
    const float h0 = 1.f / 16.f, h1 = 4.f / 16.f, h2 = 6.f / 16.f;
    auto rd = [&](int plane_off, int v, int u) -> float {  // element (v, u) of a band-distributed plane
        const int c = v / RB;
        return peer[c][plane_off + (v - c * RB) * nu + u];
    };
    const int xo = L.oScr, qo = L.oScr + bsz;

#ifdef LCB_PHASE_TIMERS
    long long ph[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long tlast = clock64();
#define CPHASE(k) { const long long tnow = clock64(); ph[k] += tnow - tlast; tlast = tnow; }
#else
#define CPHASE(k)
#endif
    for (int it = 0; it < A.n_iter; ++it) {
        // ---- halo rows of s from the neighbours (their own rows are final: cluster barrier at the end of the update)
        for (int i = tid; i < HB * nu; i += NT) {
            Sh[i] = (crank > 0) ? pl(crank - 1, L.oS + HB * nu)[(RB - HB) * nu + i] : 0.f;
            Sh[(HB + RB) * nu + i] = (crank < CL_CTAS - 1) ? pl(crank + 1, L.oS + HB * nu)[i] : 0.f;
        }
        for (int idx = tid; idx < N * 2 * P::GE; idx += NT) {
            const int st = idx / (2 * P::GE), rem = idx % (2 * P::GE), which = rem / P::GE, p = rem % P::GE;
            const float c = fk * sp[st * 12 + (which ? 1 : 2)];      // which=0: y axis, 1: x axis
            const float ic = floorf(c + 0.5f);
            float e, de;
            lcb_tap(cv, K, c - ic, p, e, de);
            taps[(st * 4 + (which ? 2 : 0)) * LCB_GE_MAX + p] = e;
            taps[(st * 4 + (which ? 3 : 1)) * LCB_GE_MAX + p] = de;
        }
        for (int i = tid; i < (RB + 2 * HB) * nu; i += NT) GRh[i] = 0.f;
        for (int i = tid; i < L.scr; i += NT) scr[i] = 0.f;          // halo rows of Vg / Vd must read as zero
        __syncthreads();
        CPHASE(0)

        float chi = 0.f, reg = 0.f;
        for (int st = 0; st < N; ++st) {
            const float a = sp[st * 12], cx = fk * sp[st * 12 + 1], cy = fk * sp[st * 12 + 2];
            const int icx = (int)floorf(cx + 0.5f), icy = (int)floorf(cy + 0.5f);
            const float* tp = taps + st * 4 * LCB_GE_MAX;
            // vertical decimating pass over the band: rows K Yl - icy - G/2 + p of S (own rows + halo)
            lcb_pass1<K, G, true, 4>(S, nu, nu, NB, icy, tp, tp + LCB_GE_MAX, Vg, Vd, ldv, tid, NT);
            __syncthreads();
            float ga = 0.f, gx = 0.f, gy = 0.f;
            const float* ds = dT + (size_t)st * nn + Y0;
            const float* ws = wT + (size_t)st * nn + Y0;
            auto consume = [&](int Yl, int X, float m0, float mx, float my, float d, float w) {
                const float diff = fmaf(a, m0, -d);
                const float r = w * diff;
                rT[X * ldv + Yl] = r;
                chi = fmaf(r, diff, chi);
                ga = fmaf(r, m0, ga);
                gx = fmaf(r, mx, gx);
                gy = fmaf(r, my, gy);
            };
            cl_pass2<K, G, 4>(Vg, Vd, ldv, NB, n, icx, tp + 2 * LCB_GE_MAX, tp + 3 * LCB_GE_MAX, ds, ws, n, tid, NT, consume);
            float v3[3] = {ga, gx, gy};
            cl_block_reduce<3>(v3, red, tid);             // (barriers inside: rT complete)
            if (tid == 0) {
                for (int c = 0; c < CL_CTAS; ++c) {
                    float* q = pl(c, L.oXch) + crank * xsz + st * 4;
                    q[0] = v3[0]; q[1] = v3[1]; q[2] = v3[2];
                }
            }
            cl_pass2T<K, G, 4>(rT, ldv, nu, NB, n, icx, tp + 2 * LCB_GE_MAX, Vbar, ldb, tid, NT);
            __syncthreads();
            auto emit = [&](int vv, int u, float val) { GRh[vv * nu + u] = fmaf(a, val, GRh[vv * nu + u]); };
            cl_pass1T<K, G, 4>(Vbar, ldb, nu, NB, RB + 2 * HB, HB, icy, tp, tid, NT, emit);
            __syncthreads();
        }
        {
            float v1[1] = {chi};
            cl_block_reduce<1>(v1, red, tid);
            if (tid == 0) for (int c = 0; c < CL_CTAS; ++c) pl(c, L.oXch)[crank * xsz + A.Nmax * 4] = v1[0];
        }
        CPHASE(1)
        cl.sync();                                        // A: gradient bands with halos and per-star sums complete everywhere
        // ---- fold the neighbours' halo rows into the own band (fixed order), per-star gradients (identical in every CTA)
        if (crank > 0) for (int i = tid; i < HB * nu; i += NT) GR[i] += pl(crank - 1, L.oGR)[(HB + RB) * nu + i];
        __syncthreads();                                  // (the two row ranges overlap when RB < 2 HB)
        if (crank < CL_CTAS - 1) for (int i = tid; i < HB * nu; i += NT) GR[(RB - HB) * nu + i] += pl(crank + 1, L.oGR)[i];
        float gn2 = 0.f;
        if (tid < N) {
            float ga = 0.f, gx = 0.f, gy = 0.f;
            for (int c = 0; c < CL_CTAS; ++c) {
                const float* q = xch + c * xsz + tid * 4;
                ga += q[0]; gx += q[1]; gy += q[2];
            }
            const float a = sp[tid * 12];
            ga *= sc; gx *= sc * a * fk; gy *= sc * a * fk;
            sp[tid * 12 + 9] = ga; sp[tid * 12 + 10] = gx; sp[tid * 12 + 11] = gy;
            if (crank == 0) gn2 = ga * ga + gx * gx + gy * gy;
        }
        float chi_tot = 0.f;
        for (int c = 0; c < CL_CTAS; ++c) chi_tot += xch[c * xsz + A.Nmax * 4];
        __syncthreads();

        CPHASE(2)
        // ---- starlet regulariser on the distributed plane b: value + gradient (left in C0)
        if (do_reg) {
            for (int i = tid; i < bsz; i += NT) C0[i] = Bp[i];
            __syncthreads();
            for (int j = 0; j < J; ++j) {
                const int Dd = 1 << j;
                for (int i = tid; i < bsz; i += NT) {
                    const int u = i % nu;
                    const float* row = C0 + i - u;
                    Xp[i] = h0 * (row[max(u - 2 * Dd, 0)] + row[min(u + 2 * Dd, nu - 1)]) + h1 * (row[max(u - Dd, 0)] + row[min(u + Dd, nu - 1)]) + h2 * row[u];
                }
                cl.sync();
                const float lam = (j == 0) ? A.lam_hf : A.lam_scales;
                for (int i = tid; i < bsz; i += NT) {
                    const int v = v0 + i / nu, u = i % nu;
                    const float nxt = h0 * (rd(xo, max(v - 2 * Dd, 0), u) + rd(xo, min(v + 2 * Dd, nu - 1), u)) +
                                      h1 * (rd(xo, max(v - Dd, 0), u) + rd(xo, min(v + Dd, nu - 1), u)) + h2 * Xp[i];
                    const float al = C0[i] - nxt;
                    const float lw = lam * (Wf ? __ldg(Wf + (size_t)j * pp + i) : 1.f);
                    reg = fmaf(lw, fabsf(al), reg);
                    sg[j * bsz + i] = (al > 0.f) ? 1 : (al < 0.f) ? -1 : 0;
                    C0[i] = nxt;
                }
                cl.sync();                                // X is rewritten by the next scale
            }
            CPHASE(3)
            for (int j = J - 1; j >= 0; --j) {
                const int Dd = 1 << j;
                const float lam = (j == 0) ? A.lam_hf : A.lam_scales;
                for (int i = tid; i < bsz; i += NT) {
                    const float t = lam * (Wf ? __ldg(Wf + (size_t)j * pp + i) : 1.f) * (float)sg[j * bsz + i];
                    Qp[i] = ((j == J - 1) ? 0.f : C0[i]) - t;
                }
                __syncthreads();
                // The folded border taps of the column direction need, for every column, the sums of Q over the rows at
                // distance 1..Dd and Dd+1..2Dd from the top / bottom edge: every band sums ITS rows (local), the two edge
                // bands add the eight partial sums after the barrier.
                for (int task = tid; task < 4 * nu; task += NT) {
                    const int which = task / nu, u = task % nu;
                    const int side = which >> 1, far = which & 1;
                    const int rlo = far ? Dd + 1 : 1, rhi = min(far ? 2 * Dd : Dd, nu - 1);     // distances from the edge
                    // rows of this band: v = v0 + l, distance r = side ? nu - 1 - v : v
                    float sacc = 0.f;
                    for (int l = 0; l < RB; ++l) {
                        const int v = v0 + l, r = side ? nu - 1 - v : v;
                        if (r >= rlo && r <= rhi) sacc += Qp[l * nu + u];
                    }
                    bp4[which * nu + u] = sacc;
                }
                cl.sync();
                for (int i = tid; i < bsz; i += NT) {     // columns: H^T along v, other bands through DSMEM
                    const int v = v0 + i / nu, u = i % nu;
                    if (v > 0 && v < nu - 1) {
                        float acc = h2 * Qp[i];
                        if (v - Dd >= 0) acc = fmaf(h1, rd(qo, v - Dd, u), acc);
                        if (v + Dd < nu) acc = fmaf(h1, rd(qo, v + Dd, u), acc);
                        if (v - 2 * Dd >= 0) acc = fmaf(h0, rd(qo, v - 2 * Dd, u), acc);
                        if (v + 2 * Dd < nu) acc = fmaf(h0, rd(qo, v + 2 * Dd, u), acc);
                        Xp[i] = acc;
                    }
                }
                if (crank == 0 || crank == CL_CTAS - 1) {  // rows 0 and nu-1 collect the folded taps
                    const int side = (crank == 0) ? 0 : 1;
                    for (int u = tid; u < nu; u += NT) {
                        float P1 = 0.f, P2 = 0.f;
                        for (int c = 0; c < CL_CTAS; ++c) {     // fixed order
                            const float* q = pl(c, L.oBp4);
                            P1 += q[(2 * side) * nu + u]; P2 += q[(2 * side + 1) * nu + u];
                        }
                        const int vb = side ? nu - 1 : 0;
                        const float P0 = Qp[(vb - v0) * nu + u];
                        Xp[(vb - v0) * nu + u] = h0 * ((P0 + P1) + P2) + h1 * (P0 + P1) + h2 * P0;
                    }
                }
                cl.sync();                                // Q is rewritten by the next scale; X complete (local use only)
                for (int i = tid; i < bsz; i += NT) {     // rows: H^T along u, local
                    const int u = i % nu;
                    if (u > 0 && u < nu - 1) {
                        const float* row = Xp + i - u;
                        float acc = h2 * row[u];
                        if (u - Dd >= 0) acc = fmaf(h1, row[u - Dd], acc);
                        if (u + Dd < nu) acc = fmaf(h1, row[u + Dd], acc);
                        if (u - 2 * Dd >= 0) acc = fmaf(h0, row[u - 2 * Dd], acc);
                        if (u + 2 * Dd < nu) acc = fmaf(h0, row[u + 2 * Dd], acc);
                        const float t = lam * (Wf ? __ldg(Wf + (size_t)j * pp + i) : 1.f) * (float)sg[j * bsz + i];
                        C0[i] = t + acc;
                    }
                }
                for (int b = warp; b < 2 * RB; b += CL_WARPS) {
                    const int rl = b >> 1, side = b & 1;
                    const float* row = Xp + rl * nu;
                    float P1 = 0.f, P2 = 0.f;
                    for (int r = lane + 1; r <= 2 * Dd && r <= nu - 1; r += 32) {
                        const float x = row[side ? nu - 1 - r : r];
                        if (r <= Dd) P1 += x; else P2 += x;
                    }
                    P1 = warp_sum(P1); P2 = warp_sum(P2);
                    if (lane == 0) {
                        const int ub = side ? nu - 1 : 0, i = rl * nu + ub;
                        const float P0 = row[ub];
                        const float t = lam * (Wf ? __ldg(Wf + (size_t)j * pp + i) : 1.f) * (float)sg[j * bsz + i];
                        C0[i] = t + h0 * ((P0 + P1) + P2) + h1 * (P0 + P1) + h2 * P0;
                    }
                }
                __syncthreads();
            }
        }
        CPHASE(4)
        // ---- total gradient, norm, loss (cluster-wide sums: pushed to every CTA, summed in rank order)
        for (int i = tid; i < bsz; i += NT) {
            const float g = sc * GR[i] + (do_reg ? C0[i] : 0.f);
            GR[i] = g;
            gn2 = fmaf(g, g, gn2);
        }
        float v2[2] = {reg, gn2};
        cl_block_reduce<2>(v2, red, tid);
        if (tid == 0) for (int c = 0; c < CL_CTAS; ++c) { float* q = pl(c, L.oXch) + crank * xsz + A.Nmax * 4 + 1; q[0] = v2[0]; q[1] = v2[1]; }
        cl.sync();                                        // B
        CPHASE(5)
        float reg_tot = 0.f, gn2_tot = 0.f;
        for (int c = 0; c < CL_CTAS; ++c) { reg_tot += xch[c * xsz + A.Nmax * 4 + 1]; gn2_tot += xch[c * xsz + A.Nmax * 4 + 2]; }
        const float Lval = cv.half * chi_tot + reg_tot;
        if (crank == 0 && tid == 0 && A.loss_hist) A.loss_hist[(size_t)f * A.n_iter + it] = Lval;
        if (it == 0) {
            if (crank == 0 && tid == 0 && A.loss0) A.loss0[f] = Lval;
            if (A.grad_b0) for (int i = tid; i < bsz; i += NT) A.grad_b0[(size_t)f * pp + v0 * nu + i] = GR[i];
            if (A.grad_s0 && crank == 0 && tid < N) {
                A.grad_s0[(i0 + tid) * 3] = sp[tid * 12 + 9];
                A.grad_s0[(i0 + tid) * 3 + 1] = sp[tid * 12 + 10];
                A.grad_s0[(i0 + tid) * 3 + 2] = sp[tid * 12 + 11];
            }
        }
        if (!isfinite(Lval)) bad = 1;
        // ---- clip + schedule + AdaBelief (optax chain)
        const float gn = sqrtf(gn2_tot);
        const float cs = (gn < cv.clip) ? 1.f : cv.clip / gn;
        const float lr = A.lr * exp2f((float)it * (log2f(cv.decay) / (float)A.n_iter));
        b1t *= cv.b1; b2t *= cv.b2;
        const BeliefCoef bc = {lr, cv.b1, cv.b2, 1.f - cv.b1, 1.f - cv.b2, 1.f / (1.f - b1t), 1.f / (1.f - b2t),
                               cv.eps, cv.eps_root};
        for (int i = tid; i < bsz; i += NT) {
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
            q[1] = fminf(fmaxf(q[1], -lim), lim);
            q[2] = fminf(fmaxf(q[2], -lim), lim);
        }
        cl.sync();                                        // C: own rows of s final; xch / GR halos free for the next iteration
        CPHASE(6)
    }
#ifdef LCB_PHASE_TIMERS
    if (crank == 0 && tid == 0 && A.loss_hist && A.n_iter >= 8) for (int q = 0; q < 8; ++q) A.loss_hist[(size_t)f * A.n_iter + q] = (float)ph[q];
#endif
    for (int i = tid; i < bsz; i += NT) A.b[(size_t)f * pp + v0 * nu + i] = Bp[i];
    if (crank == 0 && tid < N) { A.a[i0 + tid] = sp[tid * 12]; A.x0[i0 + tid] = sp[tid * 12 + 1]; A.y0[i0 + tid] = sp[tid * 12 + 2]; }
    if (crank == 0 && tid == 0 && A.status) A.status[f] = bad ? LCB_ITEM_NONFINITE : LCB_ITEM_OK;
    cl.sync();
}

// ---------------------------------------------------------------- host side
bool lcb_psf_fit_cluster_ok(int n, int k, int G, int Nmax, int J, int max_smem) {
    if (G != 12 || k < 2 || k > 4 || n % CL_CTAS != 0) return false;
    const ClLayout L = cl_layout(n, k, Nmax, J);
    if (L.RB < CL_HB) return false;                       // halos must not reach beyond the adjacent band
    return (size_t)L.total * 4 + 64 <= (size_t)max_smem;
}

template <int K, int G>
static int launch_cl(const PsfArgs& A, size_t smem, cudaStream_t st) {
    LCB_CUDA(cudaFuncSetAttribute(k_psf_fit_cl<K, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    { LcbProfScope ps("k_psf_fit_cluster", st); k_psf_fit_cl<K, G><<<A.F * CL_CTAS, CL_THREADS, smem, st>>>(A); }
    LCB_CUDA(cudaGetLastError());
    return LCB_OK;
}

// runs the A.n_iter AdaBelief iterations (parameters in place, loss history, loss0 / grad outputs, status); the
// products (residuals, chi2, narrow / full PSF) come from k_psf_fit with n_iter = 0 afterwards
int lcb_psf_fit_cluster_dispatch(const PsfArgs& A, cudaStream_t st) {
    const size_t smem = (size_t)cl_layout(A.n, A.k, A.Nmax, A.J).total * 4;
    if (A.k == 2) return launch_cl<2, 12>(A, smem, st);
    if (A.k == 3) return launch_cl<3, 12>(A, smem, st);
    if (A.k == 4) return launch_cl<4, 12>(A, smem, st);
    lcb_set_error("psf fit (cluster): unsupported subsampling_factor=%d", A.k);
    return LCB_ERR_ARG;
}
