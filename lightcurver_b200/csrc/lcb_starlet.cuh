// lcb_starlet.cuh -- float4-vectorised starlet regulariser on a square grid held in shared memory (A.3, B.3), shared by K1
// (the 64 x 64 PSF grid of a frame, 256 threads) and K3 (the 128 x 128 background of the joint deconvolution, one CTA of 1024
// threads): value lambda_j sum W_j |alpha_j| and gradient g_0 = sum_j Psi_j^T (lambda_j W_j sign alpha_j).
#pragma once
#include "lcb_common.cuh"

// ---------------------------------------------------------------- float4-vectorised starlet (NU = 64 or 128)
// Same mathematics as starlet_reg_fast, with each thread owning a 4-pixel quad of 2 rows: every stencil tap
// is one LDS.128 (quads at +-D, +-2D are 16-byte aligned for D >= 4; D = 1, 2 use three neighbouring quads
// and register swizzles), edge replication / zero extension become whole-quad selects, signs are packed
// four to a word.  About half the instructions per pixel of the scalar version.
__device__ __forceinline__ float4 f4_splat(float v) { return make_float4(v, v, v, v); }
__device__ __forceinline__ float4 f4_add(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float4 f4_b3(float4 c, float4 l1, float4 r1, float4 l2, float4 r2) {
    const float h0 = 1.f / 16.f, h1 = 4.f / 16.f, h2 = 6.f / 16.f;
    return make_float4(h0 * (l2.x + r2.x) + h1 * (l1.x + r1.x) + h2 * c.x, h0 * (l2.y + r2.y) + h1 * (l1.y + r1.y) + h2 * c.y,
                       h0 * (l2.z + r2.z) + h1 * (l1.z + r1.z) + h2 * c.z, h0 * (l2.w + r2.w) + h1 * (l1.w + r1.w) + h2 * c.w);
}
// the four dilated neighbours of quad B along a row, from the quads A (4 left), C (4 right) for D = 1, 2
__device__ __forceinline__ void f4_near(int D, float4 A, float4 B, float4 C, float4& l1, float4& r1, float4& l2, float4& r2) {
    if (D == 1) {
        l1 = make_float4(A.w, B.x, B.y, B.z); r1 = make_float4(B.y, B.z, B.w, C.x);
        l2 = make_float4(A.z, A.w, B.x, B.y); r2 = make_float4(B.z, B.w, C.x, C.y);
    } else {
        l1 = make_float4(A.z, A.w, B.x, B.y); r1 = make_float4(B.z, B.w, C.x, C.y);
        l2 = A; r2 = C;
    }
}

template <int NU, int NTH>
__device__ __forceinline__ float starlet_reg_fast4(const float* __restrict__ Bp, float* __restrict__ C0,
                                                   float* __restrict__ C1, signed char* __restrict__ sg,
                                                   float* __restrict__ aux,
                                                   const float* __restrict__ Wf, float lam_hf, float lam_scales,
                                                   int J, int tid) {
    constexpr int PP = NU * NU, QR = NU / 4, LDC = QR + 1, ROWS = NU * QR / NTH, PQ = PP / 4;
    static_assert((NU == 64 || NU == 128) && NTH % QR == 0 && ROWS >= 1 && NTH >= 2 * NU, "starlet_reg_fast4: unsupported grid / CTA size");
    const float h0 = 1.f / 16.f, h1 = 4.f / 16.f;
    const int q = tid & (QR - 1), rg = tid / QR, v0 = ROWS * rg, u0 = 4 * q;
    float* chkR = aux;                                   // [NU][LDC] quad sums of the rows of C1
    float* ext = chkR + NU * LDC;                        // [NU][2]
    auto L4 = [](const float* p) { return *reinterpret_cast<const float4*>(p); };
    auto S4 = [](float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; };
    float reg = 0.f;
    float4 wreg[ROWS];
    for (int j = 0; j < J; ++j) {
        const int D = 1 << j;
        const float* cur = (j == 0) ? Bp : C0;
        const float lam = (j == 0) ? lam_hf : lam_scales;
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
            float4 w = Wf ? __ldg(reinterpret_cast<const float4*>(Wf + (size_t)j * PP + (v0 + r) * NU + u0)) : f4_splat(1.f);
            wreg[r] = make_float4(lam * w.x, lam * w.y, lam * w.z, lam * w.w);
        }
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
            const float* row = cur + (v0 + r) * NU;
            const float4 B = L4(row + u0);
            float4 l1, r1, l2, r2;
            if (D < 4) {
                const float4 A = (q == 0) ? f4_splat(row[0]) : L4(row + u0 - 4);
                const float4 C = (q == QR - 1) ? f4_splat(row[NU - 1]) : L4(row + u0 + 4);
                f4_near(D, A, B, C, l1, r1, l2, r2);
            } else {
                l1 = (u0 - D < 0) ? f4_splat(row[0]) : L4(row + u0 - D);
                r1 = (u0 + D >= NU) ? f4_splat(row[NU - 1]) : L4(row + u0 + D);
                l2 = (u0 - 2 * D < 0) ? f4_splat(row[0]) : L4(row + u0 - 2 * D);
                r2 = (u0 + 2 * D >= NU) ? f4_splat(row[NU - 1]) : L4(row + u0 + 2 * D);
            }
            S4(C1 + (v0 + r) * NU + u0, f4_b3(B, l1, r1, l2, r2));
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
            const int v = v0 + r, idx = v * NU + u0;
            const int vm2 = max(v - 2 * D, 0), vm1 = max(v - D, 0), vp1 = min(v + D, NU - 1), vp2 = min(v + 2 * D, NU - 1);
            const float4 nxt = f4_b3(L4(C1 + idx), L4(C1 + vm1 * NU + u0), L4(C1 + vp1 * NU + u0), L4(C1 + vm2 * NU + u0), L4(C1 + vp2 * NU + u0));
            const float4 c = L4(cur + idx);
            const float4 al = make_float4(c.x - nxt.x, c.y - nxt.y, c.z - nxt.z, c.w - nxt.w);
            reg = fmaf(wreg[r].x, fabsf(al.x), reg); reg = fmaf(wreg[r].y, fabsf(al.y), reg);
            reg = fmaf(wreg[r].z, fabsf(al.z), reg); reg = fmaf(wreg[r].w, fabsf(al.w), reg);
            // four signs packed in one byte, two bits each: 0 -> -1, 1 -> 0, 2 -> +1
            const int sx = (al.x > 0.f) ? 2 : (al.x < 0.f) ? 0 : 1, sy = (al.y > 0.f) ? 2 : (al.y < 0.f) ? 0 : 1;
            const int sz = (al.z > 0.f) ? 2 : (al.z < 0.f) ? 0 : 1, sw = (al.w > 0.f) ? 2 : (al.w < 0.f) ? 0 : 1;
            sg[j * PQ + (idx >> 2)] = (signed char)(sx | (sy << 2) | (sz << 4) | (sw << 6));
            S4(C0 + idx, nxt);
        }
        __syncthreads();
    }
    for (int j = J - 1; j >= 0; --j) {
        const int D = 1 << j;
        const int m1 = min(D, NU), m2 = min(2 * D, NU);
        float4 tj[ROWS];
        // (1) q = g_{j+1} (+ row-pass border extras of the previous scale) - t_j
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
            const int v = v0 + r, idx = v * NU + u0;
            const int pk = (int)(unsigned char)sg[j * PQ + (idx >> 2)];
            tj[r] = make_float4(wreg[r].x * (float)((pk & 3) - 1), wreg[r].y * (float)(((pk >> 2) & 3) - 1),
                                wreg[r].z * (float)(((pk >> 4) & 3) - 1), wreg[r].w * (float)(((pk >> 6) & 3) - 1));
            float4 g = f4_splat(0.f);
            if (j != J - 1) {
                g = L4(C0 + idx);
                if (q == 0) g.x += ext[v * 2];
                if (q == QR - 1) g.w += ext[v * 2 + 1];
            }
            S4(C0 + idx, make_float4(g.x - tj[r].x, g.y - tj[r].y, g.z - tj[r].z, g.w - tj[r].w));
        }
        if (j > 0) {
            const float lamn = (j - 1 == 0) ? lam_hf : lam_scales;
#pragma unroll
            for (int r = 0; r < ROWS; ++r) {
                float4 w = Wf ? __ldg(reinterpret_cast<const float4*>(Wf + (size_t)(j - 1) * PP + (v0 + r) * NU + u0)) : f4_splat(1.f);
                wreg[r] = make_float4(lamn * w.x, lamn * w.y, lamn * w.z, lamn * w.w);
            }
        }
        __syncthreads();
        // (2) Hcol^T (zero extension) + folded taps on rows 0 / NU-1, quad sums of the result
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
            const int v = v0 + r, idx = v * NU + u0;
            const float4 z = f4_splat(0.f);
            float4 acc = f4_b3(L4(C0 + idx), (v - D >= 0) ? L4(C0 + idx - D * NU) : z, (v + D < NU) ? L4(C0 + idx + D * NU) : z,
                               (v - 2 * D >= 0) ? L4(C0 + idx - 2 * D * NU) : z, (v + 2 * D < NU) ? L4(C0 + idx + 2 * D * NU) : z);
            if (v == 0 || v == NU - 1) {
                float4 s1 = z, s2 = z;
                for (int i = 0; i < m2; ++i) {
                    const float4 y = L4(C0 + ((v == 0) ? i : NU - 1 - i) * NU + u0);
                    s2 = f4_add(s2, y);
                    if (i < m1) s1 = f4_add(s1, y);
                }
                acc.x += h0 * s2.x + h1 * s1.x; acc.y += h0 * s2.y + h1 * s1.y;
                acc.z += h0 * s2.z + h1 * s1.z; acc.w += h0 * s2.w + h1 * s1.w;
            }
            S4(C1 + idx, acc);
            chkR[v * LDC + q] = (acc.x + acc.y) + (acc.z + acc.w);
        }
        __syncthreads();
        // (3) Hrow^T (zero extension) + t_j ; threads 0..127 prepare the folded extras of the rows
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
            const float* row = C1 + (v0 + r) * NU;
            const float4 z = f4_splat(0.f);
            const float4 B = L4(row + u0);
            float4 l1, r1, l2, r2;
            if (D < 4) {
                const float4 A = (q == 0) ? z : L4(row + u0 - 4);
                const float4 C = (q == QR - 1) ? z : L4(row + u0 + 4);
                f4_near(D, A, B, C, l1, r1, l2, r2);
            } else {
                l1 = (u0 - D < 0) ? z : L4(row + u0 - D);
                r1 = (u0 + D >= NU) ? z : L4(row + u0 + D);
                l2 = (u0 - 2 * D < 0) ? z : L4(row + u0 - 2 * D);
                r2 = (u0 + 2 * D >= NU) ? z : L4(row + u0 + 2 * D);
            }
            const float4 a4 = f4_b3(B, l1, r1, l2, r2);
            S4(C0 + (v0 + r) * NU + u0, make_float4(tj[r].x + a4.x, tj[r].y + a4.y, tj[r].z + a4.z, tj[r].w + a4.w));
        }
        if (tid < 2 * NU) {
            const int v = tid & (NU - 1);
            const bool left = tid < NU;
            const float* row = C1 + v * NU;
            float s1 = 0.f, s2 = 0.f;
            if (m2 >= 4) {
                for (int c = 0; c < m2 / 4; ++c) {
                    const float y = chkR[v * LDC + (left ? c : QR - 1 - c)];
                    s2 += y;
                    if (c < m1 / 4) s1 += y;
                }
                if (m1 < 4) for (int i = 0; i < m1; ++i) s1 += row[left ? i : NU - 1 - i];
            } else {
                for (int i = 0; i < m2; ++i) { const float y = row[left ? i : NU - 1 - i]; s2 += y; if (i < m1) s1 += y; }
            }
            ext[v * 2 + (left ? 0 : 1)] = h0 * s2 + h1 * s1;
        }
        __syncthreads();
    }
    if (q == 0 || q == QR - 1) {
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
            if (q == 0) C0[(v0 + r) * NU] += ext[(v0 + r) * 2];
            else C0[(v0 + r) * NU + NU - 1] += ext[(v0 + r) * 2 + 1];
        }
    }
    __syncthreads();
    return reg;
}

