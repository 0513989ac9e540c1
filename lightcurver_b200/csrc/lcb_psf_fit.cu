// lcb_psf_fit.cu -- K1: per-frame PSF pixel-grid fit (stage 2 of starred build_psf, called at
// lightcurver/processes/psf_modelling.py:164-171): AdaBelief over {background grid b (nu^2), a_i,
// x0_i, y0_i} with the Moffat fixed, loss = 1/2 sum_i sum_p w (m_i - d_i)^2 + starlet-L1(b; W)
// (SURVEY.md A.1-A.5).  One CTA per frame runs ALL iterations: forward model, hand-derived adjoint
// (SURVEY.md B.1-B.3), global-norm clip and the AdaBelief update are fused in one kernel, with the
// grid planes (s, b, grad, mu, nu, 2 starlet scratch) resident in shared memory.
#include "lcb_psf.cuh"
#include "lcb_distort.cuh"
#include "lcb_starlet.cuh"

// ---------------------------------------------------------------- block reduction of NV scalars
template <int NV, int NW = PSF_WARPS>
__device__ __forceinline__ void block_reduce(float (&v)[NV], float* red /* [NW][NV] */, int tid) {
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
        for (int w = 0; w < NW; ++w) s += red[w * NV + i];
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


// ---------------------------------------------------------------- in-place starlet regulariser for large grids
// Grids whose seven planes do not fit in shared memory (BASELINE cfg5: 192 x 192) keep them in the L2 / HBM workspace,
// where every one of the 35 stencil passes per iteration is a latency-bound gather.  Here ONE plane lives in shared
// memory (in the star-pass scratch, which is dead during this phase; leading dimension nu + 1: conflict free along rows
// and columns) and every dilated 5-tap pass runs IN PLACE: a thread owns one residue class mod D of one line and sweeps
// it with a rolling register window of original values (the classes of a line are independent except through the
// edge-replicated / folded border pixels, whose original values or folded sums are saved before the sweep).  Global
// traffic per scale drops to coalesced streams: c_j read once, t_j written once and read twice.
// Returns the per-thread partial of the regulariser; leaves d reg / d b in the global plane C0g (after a barrier).
template <bool ADJ>
__device__ __forceinline__ void starlet_sweep(float* __restrict__ base, int stride, int nu, int D, int r, float edgeL, float edgeR) {
    const float h0 = 1.f / 16.f, h1 = 4.f / 16.f, h2 = 6.f / 16.f;
    const int m = (nu - r + D - 1) / D;
    float p2 = ADJ ? 0.f : edgeL, p1 = p2;
    auto ldo = [&](int i) -> float { const int u = r + i * D; return (u < nu) ? base[u * stride] : (ADJ ? 0.f : edgeR); };
    float c = ldo(0), n1 = ldo(1);
    for (int i = 0; i < m; ++i) {
        const float n2 = ldo(i + 2);
        float out = h2 * c + h1 * (p1 + n1) + h0 * (p2 + n2);
        const int u = r + i * D;
        if (ADJ) { if (u == 0) out = edgeL; else if (u == nu - 1) out = edgeR; }
        base[u * stride] = out;
        p2 = p1; p1 = c; c = n1; n1 = n2;
    }
}

// folded border value of H^T for one end of a line: h0 (P0 + S1 + S2) + h1 (P0 + S1) + h2 P0, S1 / S2 = sums of the
// elements at distance 1..D / D+1..2D from that end (one warp per line end)
__device__ __forceinline__ float starlet_fold(const float* __restrict__ base, int stride, int nu, int D, int side, int lane) {
    float S1 = 0.f, S2 = 0.f;
    for (int r = lane + 1; r <= 2 * D && r <= nu - 1; r += 32) {
        const float x = base[(side ? nu - 1 - r : r) * stride];
        if (r <= D) S1 += x; else S2 += x;
    }
    S1 = warp_sum(S1); S2 = warp_sum(S2);
    const float P0 = base[(side ? nu - 1 : 0) * stride];
    return (1.f / 16.f) * ((P0 + S1) + S2) + (4.f / 16.f) * (P0 + S1) + (6.f / 16.f) * P0;
}

template <int NTH>
__device__ __forceinline__ float starlet_reg_inplace(const float* __restrict__ Bp, float* __restrict__ C0g, float* __restrict__ Tj,
                                                     const float* __restrict__ Wf, float* __restrict__ P, float* __restrict__ side,
                                                     int nu, int J, float lam_hf, float lam_scales, int tid) {
    const int pp = nu * nu, ld = nu + 1;
    const int lane = tid & 31, warp = tid >> 5;
    constexpr int NWP = NTH / 32;
    constexpr int UQ = 16;       // pixels per lane and trip of the point-wise passes: their global loads are issued together
    float* eA = side;            // [nu] left / top edge values (forward: originals, adjoint: folded sums)
    float* eB = side + nu;       // [nu] right / bottom
    float reg = 0.f;
    // point-wise passes walk (row = warp, column = lane + 32 q): coalesced in global memory, conflict free in P, no division
    // (two rows per trip: q = 0..UQ/2-1 walks the columns of row v, q = UQ/2.. those of row v + NWP)
#define LCB_PW_LOOP(BODY_LOAD, BODY_USE)                                              \
    for (int v0 = warp; v0 < nu; v0 += 2 * NWP)                                       \
        for (int u0 = lane; u0 < nu; u0 += 32 * (UQ / 2)) {                           \
            _Pragma("unroll") for (int q = 0; q < UQ; ++q) { const int v = v0 + (q / (UQ / 2)) * NWP; const int u = u0 + 32 * (q % (UQ / 2)); const int i = v * nu + u; const bool ok = (u < nu) && (v < nu); BODY_LOAD }   \
            _Pragma("unroll") for (int q = 0; q < UQ; ++q) { const int v = v0 + (q / (UQ / 2)) * NWP; const int u = u0 + 32 * (q % (UQ / 2)); const int i = v * nu + u; const int o = v * ld + u; if ((u < nu) && (v < nu)) { BODY_USE } }  \
        }
    {
        float bv[UQ];
        LCB_PW_LOOP(bv[q] = ok ? Bp[i] : 0.f;, P[o] = bv[q]; (void)i;)
    }
    __syncthreads();
    for (int j = 0; j < J; ++j) {
        const int D = 1 << j;
        const float* cur = (j == 0) ? Bp : C0g;
        for (int axis = 0; axis < 2; ++axis) {            // rows (axis 0), then columns
            const int ls = axis ? 1 : ld, es = axis ? ld : 1;           // stride between lines / along a line
            for (int l = tid; l < nu; l += NTH) { eA[l] = P[l * ls]; eB[l] = P[l * ls + (nu - 1) * es]; }
            __syncthreads();
            for (int t = tid; t < nu * D; t += NTH) {
                const int l = t % nu, r = t / nu;
                starlet_sweep<false>(P + l * ls, es, nu, D, r, eA[l], eB[l]);
            }
            __syncthreads();
        }
        const float lam = (j == 0) ? lam_hf : lam_scales;
        const float* Wj = Wf ? Wf + (size_t)j * pp : nullptr;
        float* Tw = Tj + (size_t)j * pp;
        const bool keep = j < J - 1;
        {
            float cu[UQ], wv[UQ];
            LCB_PW_LOOP(cu[q] = ok ? cur[i] : 0.f; wv[q] = (ok && Wj) ? __ldg(Wj + i) : 1.f;,
                        const float nxt = P[o]; const float al = cu[q] - nxt; const float lw = lam * wv[q];
                        reg = fmaf(lw, fabsf(al), reg);
                        Tw[i] = (al > 0.f) ? lw : (al < 0.f) ? -lw : 0.f;
                        if (keep) C0g[i] = nxt;)
        }
        __syncthreads();
    }
    // adjoint recursion: g_J = 0; g_j = t_j + H_j^T (g_{j+1} - t_j)
    for (int j = J - 1; j >= 0; --j) {
        const int D = 1 << j;
        const float* T = Tj + (size_t)j * pp;
        const bool first = (j == J - 1);
        {
            float tv[UQ];
            LCB_PW_LOOP(tv[q] = ok ? T[i] : 0.f;, P[o] = (first ? 0.f : P[o]) - tv[q]; (void)i;)
        }
        __syncthreads();
        for (int axis = 1; axis >= 0; --axis) {           // columns, then rows
            const int ls = axis ? 1 : ld, es = axis ? ld : 1;
            for (int b = warp; b < 2 * nu; b += NWP) {
                const int l = b >> 1, sd = b & 1;
                const float fv = starlet_fold(P + l * ls, es, nu, D, sd, lane);
                if (lane == 0) (sd ? eB : eA)[l] = fv;
            }
            __syncthreads();
            for (int t = tid; t < nu * D; t += NTH) {
                const int l = t % nu, r = t / nu;
                starlet_sweep<true>(P + l * ls, es, nu, D, r, eA[l], eB[l]);
            }
            __syncthreads();
        }
        {
            float tv[UQ];
            LCB_PW_LOOP(tv[q] = ok ? T[i] : 0.f;, P[o] += tv[q]; (void)i;)
        }
        __syncthreads();
    }
    for (int v = warp; v < nu; v += NWP)
        for (int u = lane; u < nu; u += 32) C0g[v * nu + u] = P[v * ld + u];
#undef LCB_PW_LOOP
    __syncthreads();
    return reg;
}

// ---------------------------------------------------------------- fast starlet regulariser
// Compile-time grid side NU (32 or 64, PSF_THREADS % NU == 0): thread <-> (column u, ROWS consecutive rows),
// no integer division, clamped indices hoisted, sign(alpha_j) kept as int8 planes in shared memory
// (t_j = lambda_j W_j sign(alpha_j) is rebuilt from W, fetched one phase ahead), and a branch-free adjoint:
// H^T y = zero-extended symmetric stencil + the out-of-range taps folded onto the two border pixels
// (prefix / suffix sums of D and 2D elements).  The folded sums come from 8-element chunk
// sums that the PRODUCING phase leaves in shared memory (registers for the column direction, three
// shuffles per row for the row direction), so no thread ever walks a line serially.
// Returns the per-thread partial of the regulariser; leaves g_0 = d reg / d b in C0 (after a barrier).
// aux: [2*GROUPS*NU + NU*NU/8 + 2*NU] floats of shared scratch.
template <int NU, int NTH>
__device__ __forceinline__ float starlet_reg_fast(const float* __restrict__ Bp, float* __restrict__ C0,
                                                  float* __restrict__ C1, signed char* __restrict__ sg,
                                                  float* __restrict__ aux,
                                                  const float* __restrict__ Wf, float lam_hf, float lam_scales,
                                                  int J, int tid) {
    constexpr int GROUPS = NTH / NU;
    constexpr int ROWS = NU / GROUPS;
    static_assert(ROWS >= 2 && (ROWS & (ROWS - 1)) == 0, "rows per group must be a power of two");
    constexpr int CH = 8;                                // chunk length of the row-direction sums
    constexpr int PP = NU * NU;
    const float h0 = 1.f / 16.f, h1 = 4.f / 16.f, h2 = 6.f / 16.f;
    const int u = tid % NU, rg = tid / NU, v0 = rg * ROWS;
    float* chkC = aux;                                   // [GROUPS][NU] sums of the ROWS rows of a group
    float* chkR = chkC + GROUPS * NU;                    // [NU][NU/CH]  sums of CH consecutive columns
    float* ext = chkR + NU * (NU / CH);                  // [NU][2]      folded-tap extras of the row pass
    float reg = 0.f;
    float wreg[ROWS];
    for (int j = 0; j < J; ++j) {
        const int D = 1 << j;
        const float* cur = (j == 0) ? Bp : C0;
        const float lam = (j == 0) ? lam_hf : lam_scales;
#pragma unroll
        for (int r = 0; r < ROWS; ++r) wreg[r] = lam * (Wf ? __ldg(Wf + (size_t)j * PP + (v0 + r) * NU + u) : 1.f);
        const int um2 = max(u - 2 * D, 0), um1 = max(u - D, 0), up1 = min(u + D, NU - 1), up2 = min(u + 2 * D, NU - 1);
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
            const float* row = cur + (v0 + r) * NU;
            C1[(v0 + r) * NU + u] = h0 * (row[um2] + row[up2]) + h1 * (row[um1] + row[up1]) + h2 * row[u];
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
            const int v = v0 + r, idx = v * NU + u;
            const int vm2 = max(v - 2 * D, 0), vm1 = max(v - D, 0), vp1 = min(v + D, NU - 1), vp2 = min(v + 2 * D, NU - 1);
            const float nxt = h0 * (C1[vm2 * NU + u] + C1[vp2 * NU + u]) + h1 * (C1[vm1 * NU + u] + C1[vp1 * NU + u]) + h2 * C1[idx];
            const float al = cur[idx] - nxt;
            reg = fmaf(wreg[r], fabsf(al), reg);
            sg[j * PP + idx] = (al > 0.f) ? 1 : (al < 0.f) ? -1 : 0;
            C0[idx] = nxt;
        }
        __syncthreads();
    }
    // ---- backward sweep.  wreg holds lambda W of scale J-1.
    // prefix(m) / suffix(m) of a line from chunk sums (m is a power of two; chunks of c elements)
    for (int j = J - 1; j >= 0; --j) {
        const int D = 1 << j;
        const int m1 = min(D, NU), m2 = min(2 * D, NU);
        float tj[ROWS];
        // (1) q = g_{j+1} - t_j  (+ the row-pass border extras of the previous scale), column chunk sums
        float csum = 0.f;
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
            const int v = v0 + r, idx = v * NU + u;
            tj[r] = wreg[r] * (float)sg[j * PP + idx];
            float g = 0.f;
            if (j != J - 1) {
                g = C0[idx];
                if (u == 0) g += ext[v * 2];
                if (u == NU - 1) g += ext[v * 2 + 1];
            }
            const float q = g - tj[r];
            C0[idx] = q;
            csum += q;
        }
        chkC[rg * NU + u] = csum;
        if (j > 0) {
            const float lamn = (j - 1 == 0) ? lam_hf : lam_scales;
#pragma unroll
            for (int r = 0; r < ROWS; ++r) wreg[r] = lamn * (Wf ? __ldg(Wf + (size_t)(j - 1) * PP + (v0 + r) * NU + u) : 1.f);
        }
        __syncthreads();
        // (2) Hcol^T along v (fixed u) -> C1, with the folded taps on rows 0 and NU-1, + row chunk sums
        float c1v[ROWS];
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
            const int v = v0 + r;
            float acc = h2 * C0[v * NU + u];
            if (v - D >= 0) acc = fmaf(h1, C0[(v - D) * NU + u], acc);
            if (v + D < NU) acc = fmaf(h1, C0[(v + D) * NU + u], acc);
            if (v - 2 * D >= 0) acc = fmaf(h0, C0[(v - 2 * D) * NU + u], acc);
            if (v + 2 * D < NU) acc = fmaf(h0, C0[(v + 2 * D) * NU + u], acc);
            c1v[r] = acc;
        }
        if (rg == 0 || rg == GROUPS - 1) {
            // sums of the first / last m rows of column u
            float s1 = 0.f, s2 = 0.f;
            const bool top = (rg == 0);
            if (m2 >= ROWS) {
                for (int g = 0; g < m2 / ROWS; ++g) {
                    const float c = chkC[(top ? g : GROUPS - 1 - g) * NU + u];
                    s2 += c;
                    if (g < m1 / ROWS) s1 += c;
                }
                if (m1 < ROWS) for (int i = 0; i < m1; ++i) s1 += C0[(top ? i : NU - 1 - i) * NU + u];
            } else {
                for (int i = 0; i < m2; ++i) { const float y = C0[(top ? i : NU - 1 - i) * NU + u]; s2 += y; if (i < m1) s1 += y; }
            }
            const float e = h0 * s2 + h1 * s1;
            if (top) c1v[0] += e;
            if (rg == GROUPS - 1) {
                if (GROUPS == 1 && top) {                 // (never: GROUPS >= 8) keeps both borders correct
                }
                if (!top) c1v[ROWS - 1] += e;
            }
        }
        if (GROUPS == 1) { /* unreachable for NU in {32, 64} */ }
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
            C1[(v0 + r) * NU + u] = c1v[r];
            float cs = c1v[r];
            cs += __shfl_xor_sync(0xffffffffu, cs, 1);
            cs += __shfl_xor_sync(0xffffffffu, cs, 2);
            cs += __shfl_xor_sync(0xffffffffu, cs, 4);
            if ((u & (CH - 1)) == 0) chkR[(v0 + r) * (NU / CH) + (u >> 3)] = cs;
        }
        __syncthreads();
        // (3) Hrow^T along u (fixed v) + t_j -> C0 ; threads 0..2NU-1 also prepare the folded extras
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
            const float* row = C1 + (v0 + r) * NU;
            float acc = h2 * row[u];
            if (u - D >= 0) acc = fmaf(h1, row[u - D], acc);
            if (u + D < NU) acc = fmaf(h1, row[u + D], acc);
            if (u - 2 * D >= 0) acc = fmaf(h0, row[u - 2 * D], acc);
            if (u + 2 * D < NU) acc = fmaf(h0, row[u + 2 * D], acc);
            C0[(v0 + r) * NU + u] = tj[r] + acc;
        }
        if (tid < 2 * NU) {
            const int v = tid % NU;
            const bool left = tid < NU;
            const float* row = C1 + v * NU;
            float s1 = 0.f, s2 = 0.f;
            if (m2 >= CH) {
                for (int q = 0; q < m2 / CH; ++q) {
                    const float c = chkR[v * (NU / CH) + (left ? q : NU / CH - 1 - q)];
                    s2 += c;
                    if (q < m1 / CH) s1 += c;
                }
                if (m1 < CH) for (int i = 0; i < m1; ++i) s1 += row[left ? i : NU - 1 - i];
            } else {
                for (int i = 0; i < m2; ++i) { const float y = row[left ? i : NU - 1 - i]; s2 += y; if (i < m1) s1 += y; }
            }
            ext[v * 2 + (left ? 0 : 1)] = h0 * s2 + h1 * s1;
        }
        __syncthreads();
    }
    // apply the extras of the last scale
    if (u == 0 || u == NU - 1) {
#pragma unroll
        for (int r = 0; r < ROWS; ++r) C0[(v0 + r) * NU + u] += ext[(v0 + r) * 2 + (u == 0 ? 0 : 1)];
    }
    __syncthreads();
    return reg;
}

// shared scratch of the two fast starlet routines: starlet_reg_fast keeps column / row chunk sums and the folded extras
// ([2 GROUPS][NU] + [NU][NU/8] + [NU][2]); starlet_reg_fast4 quad sums with a padded leading dimension and the extras
// ([NU][NU/4 + 1] + [NU][2] -- 64 floats more than the first form on the 64-wide grid: until round 2 its last 32 rows of
// extras spilled into the first zero-halo row of the s plane, read only by the outermost Gaussian tap of a star shifted
// by two pixels)
__host__ __device__ constexpr int lcb_starlet_aux_floats(int nu, int nth) {
    const int a = 2 * (nth / nu) * nu + nu * nu / 8 + 2 * nu, b = nu * (nu / 4 + 1) + 2 * nu;
    return a > b ? a : b;
}

// ---------------------------------------------------------------- the kernel
// NS > 0: compile-time stamp side (fast path, requires NS*K in {32, 64}); NS == 0: runtime sizes.
// Fast path: 256-thread CTAs sized (128 registers, <= 113 KB of shared memory) so that TWO frames are resident
// per SM: the barrier-separated phases of one frame overlap with the other's instead of idling the SM.
// (Measured alternatives, see profiles/README.md: 512-thread CTAs one per SM; two star groups with named
// barriers; warp specialisation with the starlet on 4 / 8 dedicated warps -- all slower.)
#define FIT_THREADS(K, NS) ((NS) > 0 ? 256 : PSF_THREADS)
template <int K, int G, int NS>
__global__ void __launch_bounds__(FIT_THREADS(K, NS), ((NS) > 0 ? 2 : 1)) k_psf_fit(PsfArgs A) {
    using P = LcbPass<K, G>;
    constexpr bool FAST = (NS > 0);
    constexpr int NT = FIT_THREADS(K, NS);          // threads of the CTA
    constexpr int NW = NT / 32;
    static_assert(!FAST || (NS * K == 32 || NS * K == 64), "fast path needs a 32 or 64 wide grid");
    extern __shared__ __align__(16) float sm[];
    const int n = FAST ? NS : A.n, nu = FAST ? NS * K : A.nu, nn = n * n, pp = nu * nu, tid = threadIdx.x;
    const int ldv = n + 1, ldt = n + 1, ldb = nu + 1;
    const int f = blockIdx.x;
    const int i0 = A.star_off[f], N = A.star_off[f + 1] - i0;
    const DevConv cv = A.cv;
    const float fk = (float)K;
    const int J = A.J;

    // ---- shared layout.  FAST (<= 113 KB so that two CTAs share an SM): s (with halo), b, grad in shared memory;
    // the starlet scratch planes C0, C1 alias the star-pass scratch (Vg, Vd, r^T, Vbar: disjoint phases, the region
    // is re-zeroed every iteration for the halos); AdaBelief moments live in the L2-resident workspace (touched once
    // per iteration, coalesced); signs are packed 4 per byte (64-wide grid).  Generic: 7 planes, shared if they fit.
    float* taps = sm;                                   // [Nmax][4][LCB_GE_MAX]
    float* sp = taps + A.Nmax * 4 * LCB_GE_MAX;         // [Nmax][12] a,x0,y0, mu3, nu3, g3
    float* redS = sp + A.Nmax * 12;                     // [Nmax][PSF_WARPS][4]
    float* red = redS + A.Nmax * PSF_WARPS * 4;         // [2][32 warps][4]
    // FAST: every plane read by the passes carries HB zero rows before and after (no bounds predicates)
    constexpr int HB = FAST ? 8 : 0;
    float* dth = red + 2 * 32 * 4;                      // [4][6] field distortion: theta, mu, nu, gradient (+ 8 spare)
    float* scr = dth + 32;                              // start of the star-pass scratch
    float* Vg = scr + HB * ldv;                         // [HB + nu + HB][ldv]
    float* Vd = Vg + (nu + 2 * HB) * ldv;               // [HB + nu + HB][ldv]
    float* rT = Vd + (nu + HB) * ldv + HB * ldt;        // [HB + n + HB][ldt]
    float* Vbar = rT + (n + HB) * ldt + HB * ldb;       // [HB + n + HB][ldb]
    float* aux = Vbar + (n + HB) * ldb;                 // FAST: starlet chunk sums (see starlet_reg_fast)
    const int scr_count = (int)(aux - scr);
    float* planes = aux + (FAST ? lcb_starlet_aux_floats(nu, NT) : 0);
    if constexpr (!FAST) {
        if (!A.planes_in_smem) planes = A.work + (size_t)f * A.work_per_frame + (size_t)J * pp;
    }
    float* S = planes + HB * nu;  // s = s_fixed + b, [HB + nu + HB][nu]
    float* Bp = S + pp + HB * nu; // b
    float* GR = Bp + pp;          // d chi2 / d s
    float* MU = GR + pp;
    float* NU = MU + pp;
    float* C0 = NU + pp;
    float* C1 = C0 + pp;
    signed char* sg = reinterpret_cast<signed char*>(C1 + pp);   // [J][pp] sign(alpha_j)
    float* Tj = A.work + (size_t)f * A.work_per_frame;  // generic path: [J][pp] lambda_j W_j sign(alpha_j)
    // stamps transposed ([X][Y], lanes <-> Y read coalesced) in the global workspace
    float* dT = A.work + (size_t)f * A.work_per_frame + (size_t)(J + 7) * pp;
    float* wT = dT + (size_t)A.Nmax * nn;
    // field distortion (generic path only): the resampled PSF of the current star and d loss / d (that plane), in the workspace
    float* Sd = wT + (size_t)A.Nmax * nn;
    float* Gd = Sd + pp;
    const bool distort = !FAST && A.distort != 0;
    if constexpr (FAST) {
        sg = reinterpret_cast<signed char*>(GR + pp);
        static_assert(!FAST || (2 * (NS * K + 16) * (NS + 1) + (NS + 16) * (NS + 1) + (NS + 16) * (NS * K + 1) >= 2 * NS * K * NS * K),
                      "the star-pass scratch must be able to hold the two starlet planes");
        C0 = scr; C1 = scr + pp;                        // alias: disjoint phases
        MU = A.work + (size_t)f * A.work_per_frame;     // the t_j area of the generic path is free here
        NU = MU + pp;
    }


    const float* sfix = A.s_fixed + (size_t)f * pp;
    const float* Wf = A.W ? A.W + (size_t)f * J * pp : nullptr;
    const float* dat = A.data + (size_t)i0 * nn;
    const float* wgt = A.weight + (size_t)i0 * nn;

    if constexpr (FAST) {                               // halos must read as zero
        const int zc = (int)(Bp - scr);
        for (int i = tid; i < zc; i += NT) scr[i] = 0.f;
        __syncthreads();
    }
    for (int i = tid; i < N * nn; i += NT) {
        const int st = i / nn, r = i % nn, Y = r / n, X = r % n;
        dT[(size_t)st * nn + X * n + Y] = dat[i];
        wT[(size_t)st * nn + X * n + Y] = wgt[i];
    }
    for (int i = tid; i < pp; i += NT) {
        const float b = A.b[(size_t)f * pp + i];
        Bp[i] = b; MU[i] = 0.f; NU[i] = 0.f; S[i] = sfix[i] + b;
    }
    for (int i = tid; i < N * 12; i += NT) {
        const int st = i / 12, c = i % 12;
        sp[i] = (c == 0) ? A.a[i0 + st] : (c == 1) ? A.x0[i0 + st] : (c == 2) ? A.y0[i0 + st] : 0.f;
    }
    if (tid < 24) dth[tid] = (distort && tid < 6) ? A.distortion[(size_t)f * 6 + tid] : 0.f;
    __syncthreads();

    float b1t = 1.f, b2t = 1.f;
    int bad = 0;
    const float sc = (cv.half == 0.5f) ? 1.f : 2.f;
#ifdef LCB_PHASE_TIMERS
    long long ph[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long tlast = clock64();
#define PHASE(k) { const long long tnow = clock64(); ph[k] += tnow - tlast; tlast = tnow; }
#else
#define PHASE(k)
#endif

    for (int it = 0; it <= A.n_iter; ++it) {
        const bool last = (it == A.n_iter);
        // ---- taps of every star, zero the gradient plane
        for (int idx = tid; idx < N * 2 * P::GE; idx += NT) {
            const int st = idx / (2 * P::GE), rem = idx % (2 * P::GE), which = rem / P::GE, p = rem % P::GE;
            const float c = fk * sp[st * 12 + (which ? 1 : 2)];      // which=0: y axis, 1: x axis
            const float ic = floorf(c + 0.5f);
            float e, de;
            lcb_tap(cv, K, c - ic, p, e, de);
            taps[(st * 4 + (which ? 2 : 0)) * LCB_GE_MAX + p] = e;
            taps[(st * 4 + (which ? 3 : 1)) * LCB_GE_MAX + p] = de;
        }
        for (int i = tid; i < pp; i += NT) GR[i] = 0.f;
        if constexpr (FAST) { for (int i = tid; i < scr_count; i += NT) scr[i] = 0.f; }
        __syncthreads();
        PHASE(0)

        float chi = 0.f, cnt = 0.f;
        float reg = 0.f;
        const bool do_reg = (A.lam_scales != 0.f || A.lam_hf != 0.f);
        float gth[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};  // per-thread partial of d chi2 / d theta (field distortion)
        for (int st = 0; st < N; ++st) {
            const float a = sp[st * 12], cx = fk * sp[st * 12 + 1], cy = fk * sp[st * 12 + 2];
            const int icx = (int)floorf(cx + 0.5f), icy = (int)floorf(cy + 0.5f);
            const float* tp = taps + st * 4 * LCB_GE_MAX;
            // halo variant (no bounds checks) whenever the tap windows stay within HB rows of the planes
            const bool hal = FAST && abs(icx) <= HB - G / 2 && abs(icy) <= HB - G / 2;
            LcbAffine aff = {0.f, 0.f, 0.f, 1.f};
            float sX = 0.f, sY = 0.f;
            const float* Ssrc = S;
            if constexpr (!FAST) {
                if (distort) {                          // s_i = det * bilinear(s; c0 + A_i (p - c0))
                    sX = A.stamp_xy[(size_t)(i0 + st) * 2]; sY = A.stamp_xy[(size_t)(i0 + st) * 2 + 1];
                    aff = lcb_affine(dth, sX, sY, A.distort);
                    for (int i = tid; i < pp; i += NT) Sd[i] = aff.det * lcb_bilin_value(lcb_bilin(S, nu, nu, aff, i % nu, i / nu));
                    __syncthreads();
                    Ssrc = Sd;
                }
            }
            // FAST (256 threads): task shapes chosen so that every pass is exactly one task per thread at n = 32, k = 2
            // (pass 1: 64 columns x 4 blocks of 8 rows; pass 2 / 2^T: 32 rows x 8 blocks of 4; pass 1^T: 64 x 4 blocks of 16)
            if (hal) lcb_pass1<K, G, FAST, (FAST ? 8 : 4)>(S, nu, nu, n, icy, tp, tp + LCB_GE_MAX, Vg, Vd, ldv, tid, NT);
            else lcb_pass1<K, G, false>(Ssrc, nu, nu, n, icy, tp, tp + LCB_GE_MAX, Vg, Vd, ldv, tid, NT);
            __syncthreads();
            PHASE(1)
            float ga = 0.f, gx = 0.f, gy = 0.f;
            const float* ds = dT + (size_t)st * nn;
            const float* ws = wT + (size_t)st * nn;
            float* resid = (last && A.residuals) ? A.residuals + (size_t)(i0 + st) * nn : nullptr;
            auto consume = [&](int Y, int X, float m0, float mx, float my, float d, float w) {
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
            };
            if (hal) lcb_pass2<K, G, 4, FAST>(Vg, Vd, ldv, nu, n, icx, tp + 2 * LCB_GE_MAX, tp + 3 * LCB_GE_MAX, ds, ws, n, tid, NT, consume);
            else lcb_pass2<K, G, 4, false>(Vg, Vd, ldv, nu, n, icx, tp + 2 * LCB_GE_MAX, tp + 3 * LCB_GE_MAX, ds, ws, n, tid, NT, consume);
            ga = warp_sum(ga); gx = warp_sum(gx); gy = warp_sum(gy);
            if ((tid & 31) == 0) {
                float* q = redS + (st * PSF_WARPS + (tid >> 5)) * 4;
                q[0] = ga; q[1] = gx; q[2] = gy;
            }
            __syncthreads();
            PHASE(2)
            if (last) continue;
            if (hal) lcb_pass2T<K, G, 4, FAST>(rT, ldt, nu, n, icx, tp + 2 * LCB_GE_MAX, Vbar, ldb, tid, NT);
            else lcb_pass2T<K, G, 4, false>(rT, ldt, nu, n, icx, tp + 2 * LCB_GE_MAX, Vbar, ldb, tid, NT);
            __syncthreads();
            PHASE(3)
            auto emit = [&](int v, int u, float val) { GR[v * nu + u] = fmaf(a, val, GR[v * nu + u]); };
            if constexpr (!FAST) {
                if (distort) {
                    // d loss / d s_i into its own plane, then the transposed resampling: scatter onto the four source pixels of
                    // every output pixel (float atomics: the order of the additions is not fixed) and the chain rule to theta
                    auto emit_d = [&](int v, int u, float val) { Gd[v * nu + u] = a * val; };
                    lcb_pass1T<K, G, false>(Vbar, ldb, nu, n, icy, tp, tid, NT, emit_d);
                    __syncthreads();
                    float t_ex = 0.f, t_ey = 0.f, t_sh = 0.f;
                    for (int i = tid; i < pp; i += NT) {
                        const float g = Gd[i];
                        const LcbBilin bl = lcb_bilin(S, nu, nu, aff, i % nu, i / nu);
                        const float gd = g * aff.det;
                        const bool x0 = bl.i0 >= 0 && bl.i0 < nu, x1 = bl.i0 + 1 >= 0 && bl.i0 + 1 < nu;
                        const bool y0 = bl.j0 >= 0 && bl.j0 < nu, y1 = bl.j0 + 1 >= 0 && bl.j0 + 1 < nu;
                        if (x0 && y0) atomicAdd(GR + bl.j0 * nu + bl.i0, gd * (1.f - bl.fy) * (1.f - bl.fx));
                        if (x1 && y0) atomicAdd(GR + bl.j0 * nu + bl.i0 + 1, gd * (1.f - bl.fy) * bl.fx);
                        if (x0 && y1) atomicAdd(GR + (bl.j0 + 1) * nu + bl.i0, gd * bl.fy * (1.f - bl.fx));
                        if (x1 && y1) atomicAdd(GR + (bl.j0 + 1) * nu + bl.i0 + 1, gd * bl.fy * bl.fx);
                        const float dIx = (1.f - bl.fy) * (bl.s01 - bl.s00) + bl.fy * (bl.s11 - bl.s10);
                        const float dIy = (1.f - bl.fx) * (bl.s10 - bl.s00) + bl.fx * (bl.s11 - bl.s01);
                        const float gI = (A.distort == 1) ? g * lcb_bilin_value(bl) : 0.f;      // through det
                        t_ex += gd * dIx * bl.rx + gI * (1.f + aff.ey);
                        t_ey += gd * dIy * bl.ry + gI * (1.f + aff.ex);
                        t_sh += gd * (dIx * bl.ry + dIy * bl.rx) - 2.f * aff.sh * gI;
                    }
                    gth[0] = fmaf(sX, t_ex, gth[0]); gth[1] = fmaf(sY, t_ex, gth[1]);
                    gth[2] = fmaf(sX, t_ey, gth[2]); gth[3] = fmaf(sY, t_ey, gth[3]);
                    gth[4] = fmaf(sX, t_sh, gth[4]); gth[5] = fmaf(sY, t_sh, gth[5]);
                    continue;                           // the next star's resampling pass ends with a barrier before Gd / Vbar are reused
                }
            }
            if (hal) lcb_pass1T<K, G, FAST, (FAST ? 8 : 4)>(Vbar, ldb, nu, n, icy, tp, tid, NT, emit);
            else if (!FAST && !A.planes_in_smem) lcb_pass1T_rmw<K, G>(Vbar, ldb, nu, n, icy, tp, a, GR, tid, NT);   // gradient plane in L2
            else lcb_pass1T<K, G, false>(Vbar, ldb, nu, n, icy, tp, tid, NT, emit);
        }
        __syncthreads();
        PHASE(4)
        // ---- per-star gradients (threads st < N)
        float gn2 = 0.f;
        if (tid < N) {
            float ga = 0.f, gx = 0.f, gy = 0.f;
            for (int w = 0; w < NW; ++w) {
                const float* q = redS + (tid * PSF_WARPS + w) * 4;
                ga += q[0]; gx += q[1]; gy += q[2];
            }
            const float a = sp[tid * 12];
            ga *= sc; gx *= sc * a * fk; gy *= sc * a * fk;
            sp[tid * 12 + 9] = ga; sp[tid * 12 + 10] = gx; sp[tid * 12 + 11] = gy;
            gn2 = ga * ga + gx * gx + gy * gy;
        }
        if constexpr (!FAST) {
            if (distort && !last) {                     // d loss / d theta: block sum (the half of `red` the loss reduction does not use now)
                block_reduce<6, NW>(gth, red + ((it + 1) & 1) * 32 * 4, tid);
                if (tid < 6) dth[18 + tid] = sc * gth[tid];
                if (tid == 0) {
#pragma unroll
                    for (int q = 0; q < 6; ++q) gn2 = fmaf(sc * gth[q], sc * gth[q], gn2);
                }
            }
        }
        if (last) {
            float v2[2] = {chi, cnt};
            block_reduce<2, NW>(v2, red, tid);
            if (tid == 0 && A.chi2) A.chi2[f] = v2[0] / fmaxf(v2[1], 1.f);
            break;
        }

        // ---- starlet regulariser: forward transform, loss, t_j = lambda_j W_j sign(alpha_j)
        if (do_reg && FAST) {
            if constexpr (FAST) {
                if constexpr (NS * K == 64) reg = starlet_reg_fast4<64, NT>(Bp, C0, C1, sg, aux, Wf, A.lam_hf, A.lam_scales, J, tid);
                else reg = starlet_reg_fast<NS * K, NT>(Bp, C0, C1, sg, aux, Wf, A.lam_hf, A.lam_scales, J, tid);
            }
        } else if (do_reg && !A.planes_in_smem && scr_count >= nu * (nu + 1) + 2 * nu) {
            reg = starlet_reg_inplace<NT>(Bp, C0, Tj, Wf, scr, scr + nu * (nu + 1), nu, J, A.lam_hf, A.lam_scales, tid);
        } else if (do_reg) {
            for (int j = 0; j < J; ++j) {
                const int D = 1 << j;
                const float* cur = (j == 0) ? Bp : C0;
                for (int i = tid; i < pp; i += NT) C1[i] = atrous_fwd(cur, nu, i / nu, i % nu, D, 0);
                __syncthreads();
                const float lam = (j == 0) ? A.lam_hf : A.lam_scales;
                for (int i = tid; i < pp; i += NT) {
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
                for (int i = tid; i < pp; i += NT) {
                    const float t = Tj[(size_t)j * pp + i];
                    C0[i] = ((j == J - 1) ? 0.f : C0[i]) - t;
                }
                __syncthreads();
                for (int i = tid; i < pp; i += NT)      // Hcol^T : along v for fixed u
                    C1[i] = atrous_adj_line(C0 + (i % nu), nu, nu, i / nu, D);
                __syncthreads();
                for (int i = tid; i < pp; i += NT)      // Hrow^T : along u for fixed v
                    C0[i] = Tj[(size_t)j * pp + i] + atrous_adj_line(C1 + (i / nu) * nu, 1, nu, i % nu, D);
                __syncthreads();
            }
        }
        PHASE(5)
        // ---- total gradient, norm, loss
        if (!FAST && !A.planes_in_smem) {                // planes in L2: loads of 8 pixels in flight
            constexpr int UG = 8;
            for (int i0 = tid; i0 < pp; i0 += UG * NT) {
                float gv[UG], cv0[UG];
#pragma unroll
                for (int q = 0; q < UG; ++q) { const int i = i0 + q * NT; gv[q] = (i < pp) ? GR[i] : 0.f; cv0[q] = (i < pp && do_reg) ? C0[i] : 0.f; }
#pragma unroll
                for (int q = 0; q < UG; ++q) {
                    const int i = i0 + q * NT;
                    if (i < pp) { const float g = sc * gv[q] + cv0[q]; GR[i] = g; gn2 = fmaf(g, g, gn2); }
                }
            }
        } else
        for (int i = tid; i < pp; i += NT) {
            const float g = sc * GR[i] + (do_reg ? C0[i] : 0.f);
            GR[i] = g;
            gn2 = fmaf(g, g, gn2);
        }
        float v3[3] = {chi, reg, gn2};
        block_reduce<3, NW>(v3, red + (it & 1) * 32 * 4, tid);
        const float L = cv.half * v3[0] + v3[1];
        if (tid == 0 && A.loss_hist) A.loss_hist[(size_t)f * A.n_iter + it] = L;
        if (it == 0) {
            if (tid == 0 && A.loss0) A.loss0[f] = L;
            if (A.grad_b0) for (int i = tid; i < pp; i += NT) A.grad_b0[(size_t)f * pp + i] = GR[i];
            if (distort && A.grad_dist0 && tid < 6) A.grad_dist0[(size_t)f * 6 + tid] = dth[18 + tid];
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
        const float lr = A.lr * exp2f((float)it * (log2f(cv.decay) / (float)A.n_iter));   // lr0 * decay^(it/T)
        b1t *= cv.b1; b2t *= cv.b2;
        const BeliefCoef bc = {lr, cv.b1, cv.b2, 1.f - cv.b1, 1.f - cv.b2, 1.f / (1.f - b1t), 1.f / (1.f - b2t),
                               cv.eps, cv.eps_root};
        if constexpr (FAST) {
            // the moments and the fixed Moffat image come from L2: issue the loads of four pixels before the first
            // dependent store, otherwise every pixel pays a full L2 round trip (16 per thread and iteration)
            constexpr int UQ = 4;
            static_assert(!FAST || ((NS * K) * (NS * K)) % (UQ * NT) == 0, "update loop: grid must be a multiple of 4 * NT");
            for (int i0 = tid; i0 < pp; i0 += UQ * NT) {
                float mu[UQ], nv[UQ], sf[UQ];
#pragma unroll
                for (int q = 0; q < UQ; ++q) { mu[q] = MU[i0 + q * NT]; nv[q] = NU[i0 + q * NT]; sf[q] = __ldg(sfix + i0 + q * NT); }
#pragma unroll
                for (int q = 0; q < UQ; ++q) {
                    const int i = i0 + q * NT;
                    float b = Bp[i];
                    belief_update(bc, cs * GR[i], b, mu[q], nv[q]);
                    Bp[i] = b; MU[i] = mu[q]; NU[i] = nv[q];
                    S[i] = sf[q] + b;
                }
            }
        } else if (!A.planes_in_smem) {                   // planes in L2: loads of 4 pixels in flight
            constexpr int UQ = 4;
            for (int i0 = tid; i0 < pp; i0 += UQ * NT) {
                float bq[UQ], mu[UQ], nv[UQ], gq[UQ], sf[UQ];
#pragma unroll
                for (int q = 0; q < UQ; ++q) {
                    const int i = min(i0 + q * NT, pp - 1);
                    bq[q] = Bp[i]; mu[q] = MU[i]; nv[q] = NU[i]; gq[q] = GR[i]; sf[q] = __ldg(sfix + i);
                }
#pragma unroll
                for (int q = 0; q < UQ; ++q) {
                    const int i = i0 + q * NT;
                    if (i < pp) {
                        belief_update(bc, cs * gq[q], bq[q], mu[q], nv[q]);
                        Bp[i] = bq[q]; MU[i] = mu[q]; NU[i] = nv[q];
                        S[i] = sf[q] + bq[q];
                    }
                }
            }
        } else {
            for (int i = tid; i < pp; i += NT) {
                float b = Bp[i], mu = MU[i], nv = NU[i];
                belief_update(bc, cs * GR[i], b, mu, nv);
                Bp[i] = b; MU[i] = mu; NU[i] = nv;
                S[i] = __ldg(sfix + i) + b;
            }
        }
        if (distort && tid >= 32 && tid < 38) {         // warp 1: the six distortion coefficients
            const int q = tid - 32;
            belief_update(bc, cs * dth[18 + q], dth[q], dth[6 + q], dth[12 + q]);
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
        PHASE(6)
    }
#ifdef LCB_PHASE_TIMERS
    if (tid == 0 && A.loss_hist && A.n_iter >= 8) for (int k = 0; k < 8; ++k) A.loss_hist[(size_t)f * A.n_iter + k] = (float)ph[k];
#endif

    // ---- products: fitted parameters, narrow_psf = s / sum s, full_psf = (s (*) g0) / sum
    __syncthreads();
    for (int i = tid; i < pp; i += NT) A.b[(size_t)f * pp + i] = Bp[i];
    if (tid < N) { A.a[i0 + tid] = sp[tid * 12]; A.x0[i0 + tid] = sp[tid * 12 + 1]; A.y0[i0 + tid] = sp[tid * 12 + 2]; }
    if (tid == 0 && A.status) A.status[f] = bad ? LCB_ITEM_NONFINITE : LCB_ITEM_OK;
    if (distort && tid < 6) A.distortion[(size_t)f * 6 + tid] = dth[tid];
    if (A.narrow_psf || A.full_psf) {
        float tot[1] = {0.f};
        for (int i = tid; i < pp; i += NT) tot[0] += S[i];
        block_reduce<1, NW>(tot, red, tid);
        const float inv = 1.f / tot[0];
        if (A.narrow_psf) for (int i = tid; i < pp; i += NT) A.narrow_psf[(size_t)f * pp + i] = S[i] * inv;
        if (A.full_psf) {
            float g0[G];
#pragma unroll
            for (int t = 0; t < G; ++t) {           // tau = t - G/2 + 1
                const float x = (float)(t - G / 2 + 1);
                g0[t] = cv.gnorm * expf(-x * x * cv.inv2s2);
            }
            __syncthreads();
            for (int i = tid; i < pp; i += NT) {
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
            for (int i = tid; i < pp; i += NT) {
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
            block_reduce<1, NW>(tf, red + 32 * 4, tid);
            const float invf = 1.f / tf[0];
            for (int i = tid; i < pp; i += NT) A.full_psf[(size_t)f * pp + i] = C0[i] * invf;
        }
    }
}

size_t lcb_psf_fit_smem_small(int n, int nu, int Nmax) {
    return (size_t)(Nmax * 4 * LCB_GE_MAX + Nmax * 12 + Nmax * PSF_WARPS * 4 + 2 * 32 * 4 + 32 +
                    2 * nu * (n + 1) + n * (n + 1) + n * (nu + 1)) * 4;
}

// shared bytes of the fast path on top of smem_small: 7 planes + J int8 sign planes
// fast path (on top of smem_small): s, b, grad planes + sign planes (packed 4/byte on the 64-wide grid) + starlet
// chunk sums + the zero halos (8 rows each side of s, Vg, Vd, r^T, Vbar); C0/C1 alias the scratch, moments in L2
size_t lcb_psf_fit_smem_fast_extra(int n, int nu, int J) {
    const size_t halos = (size_t)16 * (nu + 2 * (n + 1) + (n + 1) + (nu + 1)) * 4;
    const size_t signs = (nu == 64) ? (size_t)J * nu * nu / 4 : (size_t)J * nu * nu;
    return (size_t)3 * nu * nu * 4 + signs + (size_t)lcb_starlet_aux_floats(nu, 256) * 4 + halos;
}

bool lcb_psf_fit_has_fast(int n, int k, int G) {
    return G == 12 && ((k == 2 && (n == 32 || n == 16)) || (k == 1 && (n == 32 || n == 64)));
}

template <int K, int G, int NS>
static int launch_psf_fit(const PsfArgs& A, size_t smem, cudaStream_t st) {
    LCB_CUDA(cudaFuncSetAttribute(k_psf_fit<K, G, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    LCB_CUDA(cudaFuncSetAttribute(k_psf_fit<K, G, NS>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    { LcbProfScope ps("k_psf_fit", st); k_psf_fit<K, G, NS><<<A.F, FIT_THREADS(K, NS), smem, st>>>(A); }
    LCB_CUDA(cudaGetLastError());
    return LCB_OK;
}

// `fast`: the caller verified lcb_psf_fit_has_fast() and that planes + sign planes fit in shared memory
int lcb_psf_fit_dispatch(const PsfArgs& A, size_t smem, bool fast, cudaStream_t st) {
    const int G = A.cv.G;
    if (fast) {
        if (A.k == 2 && A.n == 32) return launch_psf_fit<2, 12, 32>(A, smem, st);
        if (A.k == 2 && A.n == 16) return launch_psf_fit<2, 12, 16>(A, smem, st);
        if (A.k == 1 && A.n == 32) return launch_psf_fit<1, 12, 32>(A, smem, st);
        if (A.k == 1 && A.n == 64) return launch_psf_fit<1, 12, 64>(A, smem, st);
    }
#define CASE(KK, GG) if (A.k == KK && G == GG) return launch_psf_fit<KK, GG, 0>(A, smem, st);
    CASE(1, 12) CASE(2, 12) CASE(3, 12) CASE(4, 12)
    CASE(2, 8) CASE(2, 16)
#undef CASE
    lcb_set_error("psf fit: unsupported (subsampling_factor=%d, gauss_taps=%d)", A.k, G);
    return LCB_ERR_ARG;
}
