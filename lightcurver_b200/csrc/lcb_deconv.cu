// lcb_deconv.cu -- K3: joint multi-epoch deconvolution with a shared high-resolution background.
//
// Replaces the STARRED calls of lightcurver/processes/roi_modelling.py:213-335 (setup_model, Loss,
// Optimizer('adabelief').minimize, model.model, model.getDeconvolved) and, with M = 1, the
// reference-coupled form of star_photometry.py:66-137.  Model:
//     f_e = Warp_{dx_e,dy_e,alpha_e}[h] + sum_m a_em G(. ; k (R_alpha c_m + d_e))        (nu x nu)
//     m_e = D_k[ s_e (*) f_e ] + mean_e                                                  (n x n)
// Loss (A.7): 1/2 sum w (m-d)^2 + starlet-L1(h; W) + positivity(h) + Gaussian prior on (c_x, c_y)
//             + pts-source L1 (first starlet scale of the point-source channel) + flux uniformity.
//
// One iteration =
//   k_deconv_epoch   one thread-block CLUSTER of CS CTAs per epoch (CS = 1, 2, 4, 8 chosen so that the local
//                    epochs fill the 148 SMs): every CTA owns a band of data rows; it applies the pending
//                    AdaBelief update of the per-epoch parameters, builds the rows of f_e it needs in shared
//                    memory (polyphase layout), runs the direct FP32 convolution with the k-box-folded PSF
//                    fused with the decimation for its band, broadcasts its rows of the weighted residual to
//                    the other CTAs through distributed shared memory, runs the adjoint convolution for its
//                    band, and the transposed warp reads dL/df of neighbouring bands through DSMEM
//   k_deconv_reduce  deterministic sum over the local epochs -> red[] ; multi-GPU: every rank PUSHES its
//                    partial sums into a slot of every peer's receive buffer over NVLink (peer memory mapped
//                    with CUDA IPC), then raises a flag there
//   k_deconv_update  multi-GPU: waits for the flags and sums the slots in rank order (bit-identical on all
//                    ranks); starlet / positivity / prior / flux-uniformity terms, global norm, AdaBelief on
//                    h, c_x, c_y.  No host round trip and no separate collective launch per iteration.
//   (lcb_deconv_step_local / _step_update keep the two halves separately callable so that the caller can
//    put ONE NCCL all-reduce of red[] between them instead.)
//
// Decimation is folded into the PSF: m[Y][X] = mean + 1/k^2 sum_phase sum_{av,au} S_ph[av][au] f_ph[Y+av][X+au]
// with f_ph[Y'][X'] = f[kY'+pv][kX'+pu] and S_ph the polyphase components of s (*) box_k; each phase is a
// plain 2-D correlation of an n x n plane with an NA x NA kernel (NA ~ P/k + 1): 4x fewer MACs than
// convolving at full resolution, and lanes <-> rows with an odd leading dimension is conflict free.
#include "lcb_common.cuh"
#include "lcb_starlet.cuh"
#include <cooperative_groups.h>
#include <vector>
#include <type_traits>

namespace cg = cooperative_groups;

void lcb_build_noise_table(int nu, int J, std::vector<float>& tab);
bool lcb_profiling();
int lcb_noise_weights_launch(int F, int nu, int J, const float* tab, float* W, float* work,
                             size_t work_per_frame, cudaStream_t st);

#define DC_THREADS 256
#define DC_XB 8
#define DC_MMAX 8
#define DC_MAXW 8            // ranks of one NVSwitch domain
#define DC_CSMAX 8           // portable cluster size
#define DC_EXT 20            // G + 4 <= 20

struct DeconvComm {          // peer memory of the in-kernel all-reduce
    int world, rank;
    float* slots[DC_MAXW];   // receive buffer of rank r: [2 parities][world][tot]
    int* flags[DC_MAXW];     // flags of rank r: [2][world], value = sequence number of the data in the slot
    unsigned* ctr;           // [2] local last-block counters of k_deconv_reduce
};

struct DeconvDev {
    int E, n, k, nu, P, M, NA, A0, J, G;
    int E_total, e0;                     // epochs over all ranks, global index of the first local epoch
    int tot;                             // length of red[]
    int tot_pad;                         // stride of a receive slot: tot rounded up to 4 floats (16-byte stores into peer memory)
    int free_h, free_mean, free_a, free_c, free_d;
    // inputs
    float *data, *weight, *S;            // [E][n][n], [E][n][n], [E][k*k][NA][NA]
    float *W;                            // [J][nu^2] or NULL
    float lam_scales, lam_hf, lam_pos, lam_pts, lam_fu;
    int pts_all_epochs, fu_relative;
    int has_prior; float *prior;         // [4][M] mu_x, sig_x, mu_y, sig_y
    // parameters + AdaBelief state
    float *h, *h_mu, *h_nu;              // [nu^2]
    float *c, *c_mu, *c_nu;              // [2M] c_x then c_y
    float *ep, *ep_mu, *ep_nu, *ep_g;    // [E][M+3]: a[M], dx, dy, mean  (value, moments, pending gradient)
    float *alpha;                        // [E]
    // per-iteration scratch
    float *Gh;                           // [E][nu^2] per-epoch partial dL/dh
    float *gc;                           // [E][2M]
    float *eloss;                        // [E]
    float *red;                          // [tot] reduced: dL/dh | dL/dc (2M) | loss | |g_epoch|^2 | flux sums (4M)
    float *ctl;                          // [8] clip scale, lr, 1/bc1, 1/bc2, pending flag, loss, -, peer timeout
    float *fu;                           // [3][DC_MMAX] flux uniformity: mean (= shift of the sums), A, B
    float *gpart;                        // [8][DU_CTAS] cross-CTA partial sums of k_deconv_update
    float *planes;                       // [3][nu^2] starlet scratch + Tj [J][nu^2]
    float *model;                        // [E][n][n] (written when requested)
    float *loss_hist;                    // [cap]
    unsigned *band_ctr;                  // [DC_CSMAX + 1] arrival counters of the fused reduction (self-resetting)
    unsigned *grp_ctr;                   // [NG][DC_CSMAX] arrivals per group of GS consecutive epochs and band (self-resetting)
    float *Gp;                           // [NG][nu^2] dL/dh summed over the epochs of a group (level 1 of the fused reduction)
    int GS, NG;                          // epochs per group (~ sqrt(E)), number of groups
    int *ictl;                           // [2] device-resident iteration index and exchange sequence number (graph replay)
    DevConv cv;
    DeconvComm cm;
};

// red[] layout
__host__ __device__ inline int red_loss(const DeconvDev& D) { return D.nu * D.nu + 2 * D.M; }
__host__ __device__ inline int red_flux(const DeconvDev& D) { return D.nu * D.nu + 2 * D.M + 2; }

// ---------------------------------------------------------------- shared-memory layout of k_deconv_epoch (floats)
// Planes read by the convolutions are stored with ZEROS between consecutive rows: row r keeps its n values at
// HL + r*ld + sh .. + n, the ld - n floats up to the next row are zero and serve as the right halo of row r and the
// left halo of row r+1 (ld - n >= the reach of the kernel on either side).  No bounds predicate is needed along a
// row, every window / tap load is an aligned LDS.128 (sh makes the first tap column a multiple of 4; ld = 4 mod 8
// spreads the rows of a warp over all banks), and out-of-range ROWS are simply skipped.
struct DcLayout {
    int NAp, NA8, ld, HL, shF, shR, FR, RR, pst;
    int oS, oF, oR, oG, oPar, oRed, oCp, oPts, oEx, zero_end, total;
};
__host__ __device__ inline DcLayout dc_layout(int n, int k, int NA, int A0, int CS) {
    const int kk = k * k;
    DcLayout L;
    L.NA8 = NA & ~7;
    L.NAp = (NA + 7) & ~7;
    const int A1 = A0 + NA - 1;
    int reach = (-A0 > A1 ? -A0 : A1);
    if (reach < 0) reach = 0;
    int ld = n + reach;
    while ((ld & 7) != 4) ++ld;
    L.ld = ld;
    L.HL = (reach + 3) & ~3;
    L.shF = ((-A0) % 4 + 4) % 4;                       // forward: first tap column X0 + A0 (+ sh) is a multiple of 4
    L.shR = ((A0 + 7) % 4 + 4) % 4;                    // adjoint: first window column X0 - A0 - 7 (+ sh)
    const int rpc = (n + CS - 1) / CS;
    int rows = rpc + NA - 1;
    if (rows > n) rows = n;
    L.FR = rows; L.RR = rows;
    L.pst = L.FR * ld;
    int o = 0;
    L.oS = o; o += kk * NA * L.NAp;
    L.oF = o; o += L.HL + kk * L.pst + 32;
    L.oR = o; o += L.HL + L.RR * ld + 32;
    L.zero_end = o;                                    // [oF, zero_end) is cleared at kernel start
    L.oG = o; o += 4 * DC_MMAX * 16;                   // gx | d gx | gy | d gy, [M][16] each
    L.oPar = o; o += 16;
    L.oRed = o; o += 16;
    L.oCp = o; o += DC_CSMAX * 32;                     // per-CTA partial sums, gathered in rank 0
    L.oPts = o; o += DC_MMAX * 32;                     // pts-source partial sums per warp
    L.oEx = o; o += 2 * DC_MMAX * 4 * DC_EXT;          // [axis][m][g, Bg, g', Bg'][G+4]
    L.total = o;
    return L;
}

// bands of CTA c of a cluster: own data rows [Y0,Y1), rows of f it stores [flo,fhi), rows of r it stores [rlo,rhi)
struct DcBand { int Y0, Y1, flo, fhi, rlo, rhi; };
__device__ __forceinline__ DcBand dc_band(int c, int n, int rpc, int A0, int NA) {
    DcBand b;
    b.Y0 = min(n, c * rpc); b.Y1 = min(n, b.Y0 + rpc);
    b.flo = min(b.Y0, max(0, b.Y0 + A0)); b.fhi = max(b.Y1, min(n, b.Y1 + A0 + NA - 1));
    b.rlo = min(b.Y0, max(0, b.Y0 - A0 - NA + 1)); b.rhi = max(b.Y1, min(n, b.Y1 - A0));
    return b;
}

__device__ __forceinline__ void ld8(float (&d)[8], const float* __restrict__ p) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    d[0] = a.x; d[1] = a.y; d[2] = a.z; d[3] = a.w; d[4] = b.x; d[5] = b.y; d[6] = b.z; d[7] = b.w;
}

// acc[x] += sum_{j<8} tap[j] * w[j + x],  w = lo | hi (16 consecutive floats)
template <bool REV>
__device__ __forceinline__ void blk8(float (&acc)[DC_XB], const float (&tap)[8], const float (&lo)[8], const float (&hi)[8]) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float tv = REV ? tap[7 - j] : tap[j];
#pragma unroll
        for (int x = 0; x < DC_XB; ++x) {
            const int i = j + x;
            acc[x] = fmaf(tv, (i < 8) ? lo[i] : hi[i - 8], acc[x]);
        }
    }
}

// forward, one kernel row: acc[x] += sum_t sr[t] * fr[t + x]   (fr = window start: column X0 + A0 of the f row; 16 B aligned)
__device__ __forceinline__ void corr_line(const float* __restrict__ fr, const float* __restrict__ sr, int NA, int NA8, float (&acc)[DC_XB]) {
    float lo[8], hi[8], tap[8];
    if (NA8 > 0) ld8(lo, fr);
    int t0 = 0;
    for (; t0 + 16 <= NA8; t0 += 16) {                 // two blocks per trip: the window registers ping-pong (no moves)
        ld8(hi, fr + t0 + 8);
        ld8(tap, sr + t0);
        blk8<false>(acc, tap, lo, hi);
        ld8(lo, fr + t0 + 16);
        ld8(tap, sr + t0 + 8);
        blk8<false>(acc, tap, hi, lo);
    }
    if (t0 < NA8) {
        ld8(hi, fr + t0 + 8);
        ld8(tap, sr + t0);
        blk8<false>(acc, tap, lo, hi);
    }
    const int R = NA - NA8;                            // tail taps: aligned loads, only the real taps are multiplied
    if (R > 0) {
        ld8(lo, fr + NA8); ld8(tap, sr + NA8);
        if (R > 1) ld8(hi, fr + NA8 + 8);
#pragma unroll
        for (int j = 0; j < 7; ++j) {
            if (j < R) {
#pragma unroll
                for (int x = 0; x < DC_XB; ++x) {
                    const int i = j + x;
                    acc[x] = fmaf(tap[j], (i < 8) ? lo[i] : hi[i - 8], acc[x]);
                }
            }
        }
    }
}

// adjoint, one kernel row: acc[x] += sum_t sr[t] * rr[(NA8 - 1 - t) + x]   (rr = column X0 - A0 - (NA8 - 1) of the r row).
// Blocks run over t0 = NA8-8 ... 0 with the taps used in reverse, so the window slides forwards; rr + 8*b is 16 B aligned.
template <bool SQ>
__device__ __forceinline__ void conv_line_T(const float* __restrict__ rr, const float* __restrict__ sr, int NA, int NA8, float (&acc)[DC_XB]) {
    float lo[8], hi[8], tap[8];
    // tail taps t = NA8 + j first: column offset (NA8 - 1 - t) + x = x - 1 - j, i.e. w[7 - j + x] of the window rr[-8 .. 7]
    const int R = NA - NA8;
    if (R > 0) {
        ld8(lo, rr - 8); ld8(hi, rr); ld8(tap, sr + NA8);
#pragma unroll
        for (int j = 0; j < 7; ++j) {
            if (j < R) {
                const float sv = SQ ? tap[j] * tap[j] : tap[j];
#pragma unroll
                for (int x = 0; x < DC_XB; ++x) {
                    const int i = 7 - j + x;
                    acc[x] = fmaf(sv, (i < 8) ? lo[i] : hi[i - 8], acc[x]);
                }
            }
        }
    }
    // block b (t0 = NA8 - 8 - 8 b): j = t0 + 7 - t, column offset = (NA8 - 1 - t0 - 7) + j + x = 8 b + j + x
    if (NA8 > 0) ld8(lo, rr);
    auto sq = [](float (&t)[8]) {
        if (SQ) {
#pragma unroll
            for (int i = 0; i < 8; ++i) t[i] *= t[i];
        }
    };
    int b = 0;
    for (; 8 * b + 16 <= NA8; b += 2) {                // two blocks per trip: the window registers ping-pong (no moves)
        ld8(hi, rr + 8 * b + 8);
        ld8(tap, sr + NA8 - 8 - 8 * b);
        sq(tap);
        blk8<true>(acc, tap, lo, hi);
        ld8(lo, rr + 8 * b + 16);
        ld8(tap, sr + NA8 - 16 - 8 * b);
        sq(tap);
        blk8<true>(acc, tap, hi, lo);
    }
    if (8 * b < NA8) {
        ld8(hi, rr + 8 * b + 8);
        ld8(tap, sr + NA8 - 8 - 8 * b);
        sq(tap);
        blk8<true>(acc, tap, lo, hi);
    }
}

// ---------------------------------------------------------------- the same two lines with XB outputs per thread
// With 8 outputs per thread a block of 8 taps costs 2 + 2 LDS.128 for 64 FFMA: the passes run at the shared-memory bandwidth
// (window loads are 4 wavefronts each).  XB = 16 halves the loads per FFMA (2 + 2 LDS.128 for 128 FFMA): FP32-issue bound.
// The window is a register array of XB + 8 floats that slides by 8 per block.
template <int XB>
__device__ __forceinline__ void ldx(float* d, const float* __restrict__ p) {
#pragma unroll
    for (int i = 0; i < XB / 4; ++i) {
        const float4 a = reinterpret_cast<const float4*>(p)[i];
        d[4 * i] = a.x; d[4 * i + 1] = a.y; d[4 * i + 2] = a.z; d[4 * i + 3] = a.w;
    }
}

template <int XB>
__device__ __forceinline__ void corr_line_x(const float* __restrict__ fr, const float* __restrict__ sr, int NA, int NA8, float (&acc)[XB]) {
    float w[XB + 8], tap[8];
    ldx<XB>(w, fr);
#pragma unroll 1
    for (int t0 = 0; t0 < NA8; t0 += 8) {
        ldx<8>(w + XB, fr + t0 + XB);
        ldx<8>(tap, sr + t0);
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
            for (int x = 0; x < XB; ++x) acc[x] = fmaf(tap[j], w[j + x], acc[x]);
#pragma unroll
        for (int i = 0; i < XB; ++i) w[i] = w[i + 8];
    }
    const int R = NA - NA8;                            // tail taps: aligned loads, only the real taps are multiplied
    if (R > 0) {
        ldx<8>(tap, sr + NA8);
        if (R > 1) ldx<8>(w + XB, fr + NA8 + XB);
#pragma unroll
        for (int j = 0; j < 7; ++j) {
            if (j < R) {
#pragma unroll
                for (int x = 0; x < XB; ++x) acc[x] = fmaf(tap[j], w[j + x], acc[x]);
            }
        }
    }
}

template <int XB>
__device__ __forceinline__ void conv_line_T_x(const float* __restrict__ rr, const float* __restrict__ sr, int NA, int NA8, float (&acc)[XB]) {
    float w[XB + 8], tap[8];
    const int R = NA - NA8;
    if (R > 0) {                                       // tail taps t = NA8 + j: w[7 - j + x] of the window rr[-8 .. XB)
        ldx<XB + 8>(w, rr - 8);
        ldx<8>(tap, sr + NA8);
#pragma unroll
        for (int j = 0; j < 7; ++j) {
            if (j < R) {
#pragma unroll
                for (int x = 0; x < XB; ++x) acc[x] = fmaf(tap[j], w[7 - j + x], acc[x]);
            }
        }
#pragma unroll
        for (int i = 0; i < XB; ++i) w[i] = w[i + 8];
    } else {
        ldx<XB>(w, rr);
    }
#pragma unroll 1
    for (int b = 0; 8 * b < NA8; ++b) {                // block b: taps NA8-8-8b .. NA8-1-8b in reverse, window rr + 8 b
        ldx<8>(w + XB, rr + 8 * b + XB);
        ldx<8>(tap, sr + NA8 - 8 - 8 * b);
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
            for (int x = 0; x < XB; ++x) acc[x] = fmaf(tap[7 - j], w[j + x], acc[x]);
#pragma unroll
        for (int i = 0; i < XB; ++i) w[i] = w[i + 8];
    }
}

// NB = NA8 / 8 known at compile time: straight-line code, the window blocks are distinct registers (no rotation moves) and the
// loads of the next block can be scheduled above the FFMAs of the current one
template <int XB, int NB>
__device__ __forceinline__ void corr_line_u(const float* __restrict__ fr, const float* __restrict__ sr, int R, float (&acc)[XB]) {
    float w[XB + 8 * NB + 8], tap[8];
    ldx<XB>(w, fr);
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        ldx<8>(w + XB + 8 * b, fr + XB + 8 * b);
        ldx<8>(tap, sr + 8 * b);
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
            for (int x = 0; x < XB; ++x) acc[x] = fmaf(tap[j], w[8 * b + j + x], acc[x]);
    }
    if (R > 0) {
        ldx<8>(tap, sr + 8 * NB);
        if (R > 1) ldx<8>(w + XB + 8 * NB, fr + XB + 8 * NB);
#pragma unroll
        for (int j = 0; j < 7; ++j) {
            if (j < R) {
#pragma unroll
                for (int x = 0; x < XB; ++x) acc[x] = fmaf(tap[j], w[8 * NB + j + x], acc[x]);
            }
        }
    }
}

template <int XB, int NB>
__device__ __forceinline__ void conv_line_T_u(const float* __restrict__ rr, const float* __restrict__ sr, int R, float (&acc)[XB]) {
    float w[XB + 8 * NB + 8], tap[8];                   // w[i] = rr[i - 8]
    if (R > 0) {
        ldx<XB + 8>(w, rr - 8);
        ldx<8>(tap, sr + 8 * NB);
#pragma unroll
        for (int j = 0; j < 7; ++j) {
            if (j < R) {
#pragma unroll
                for (int x = 0; x < XB; ++x) acc[x] = fmaf(tap[j], w[7 - j + x], acc[x]);
            }
        }
    } else {
        ldx<XB>(w + 8, rr);
    }
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        ldx<8>(w + 8 + XB + 8 * b, rr + XB + 8 * b);
        ldx<8>(tap, sr + 8 * (NB - 1 - b));
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
            for (int x = 0; x < XB; ++x) acc[x] = fmaf(tap[7 - j], w[8 + 8 * b + j + x], acc[x]);
    }
}

// ---------------------------------------------------------------- setup: polyphase box-folded PSF
// S[e][pv*k+pu][av-A0][au-A0] = stilde(tv = j0 - k av - pv, tu = j0 - k au - pu),
// stilde(tv,tu) = sum_{kv,ku<k} s[tv+kv][tu+ku]   (zero outside the P x P array).
// Rows are zero padded to NAp = roundup(NA, 8) taps, so that one epoch's kernel is ONE contiguous, 16-byte aligned block that the
// per-epoch kernel fetches with a single bulk asynchronous copy (TMA) straight into its shared-memory layout.
__global__ void k_deconv_fold_psf(const float* psf, float* S, int E, int P, int k, int NA, int A0) {
    const int e = blockIdx.x;
    const int j0 = (P - 1) / 2;
    const int NAp = (NA + 7) & ~7;
    const float* s = psf + (size_t)e * P * P;
    float* out = S + (size_t)e * k * k * NA * NAp;
    for (int i = threadIdx.x; i < k * k * NA * NAp; i += blockDim.x) {
        const int ph = i / (NA * NAp), r = i % (NA * NAp), av = r / NAp + A0, au = r % NAp + A0;
        if (r % NAp >= NA) { out[i] = 0.f; continue; }
        const int pv = ph / k, pu = ph % k;
        const int tv = j0 - k * av - pv, tu = j0 - k * au - pu;
        float acc = 0.f;
        for (int kv = 0; kv < k; ++kv)
            for (int ku = 0; ku < k; ++ku) {
                const int jv = tv + kv, ju = tu + ku;
                if (jv >= 0 && jv < P && ju >= 0 && ju < P) acc += s[jv * P + ju];
            }
        out[i] = acc;
    }
}

// ---------------------------------------------------------------- geometry helpers
struct Geo { float ca, sa, tx, ty, ctr; };   // q = R^-1 (p - ctr - k d) + ctr

__device__ __forceinline__ void geo_src(const Geo& g, float u, float v, float& qu, float& qv) {
    const float pu = u - g.ctr - g.tx, pv = v - g.ctr - g.ty;
    qu = g.ca * pu + g.sa * pv + g.ctr;
    qv = -g.sa * pu + g.ca * pv + g.ctr;
}

__device__ __forceinline__ float h_at(const float* __restrict__ h, int nu, int v, int u) {
    return (v >= 0 && v < nu && u >= 0 && u < nu) ? __ldg(h + v * nu + u) : 0.f;
}

#ifdef LCB_DC_TIMERS
// development build only (python -m lightcurver_b200.build --dc-timers): clock64 stamps of the phases of k_deconv_epoch per CTA
#define DC_TIM_SLOTS 16
#define DC_TIM_CTAS 2048
__device__ long long g_dc_tim[DC_TIM_CTAS][DC_TIM_SLOTS];
#define DC_STAMP(k) { if (tid == 0 && blockIdx.x < DC_TIM_CTAS) g_dc_tim[blockIdx.x][k] = clock64() - dc_t0; }
extern "C" int lcb_debug_dc_timers(long long* out, int ctas) {
    if (ctas > DC_TIM_CTAS) ctas = DC_TIM_CTAS;
    return cudaMemcpyFromSymbol(out, g_dc_tim, (size_t)ctas * DC_TIM_SLOTS * sizeof(long long)) == cudaSuccess ? 0 : -1;
}
#else
#define DC_STAMP(k)
#endif

// ---------------------------------------------------------------- per-epoch kernel (cluster of CS CTAs)
// flags: 1 = write the model image; 2 = propagate the weights instead of the residuals (noise weights); 4 = fused reduction:
// the LAST cluster to finish a band of rows sums that band of dL/dh over the local epochs (fixed order) into red[] -- or, with
// epochs sharded over GPUs (seq > 0), pushes the sums straight into the receive slots of every peer over NVLink -- and the last
// band raises the peers' flags: the exchange starts from the tail of the epoch kernel, there is no separate reduce launch.
// seq_arg < 0: the sequence number is read from D.ictl[1] (CUDA-graph replay: identical arguments every iteration).
template <int K>
__global__ void __launch_bounds__(DC_THREADS, 2) k_deconv_epoch(DeconvDev D, int flags, int CS, int seq_arg) {
    cg::cluster_group cl = cg::this_cluster();
    const int want_model = flags & 1;
    const bool noise = (flags & 2) != 0;
    extern __shared__ __align__(16) float sm[];
    const int tid = threadIdx.x;
    const int crank = (int)(blockIdx.x % CS);     // == cl.block_rank() for cluster dims (CS,1,1)
#ifdef LCB_DC_TIMERS
    const long long dc_t0 = clock64();
    if (tid == 0 && blockIdx.x < DC_TIM_CTAS) { unsigned long long gt; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt)); g_dc_tim[blockIdx.x][14] = (long long)gt; }
#endif
    const int e = blockIdx.x / CS;
    constexpr int k = K, kk = K * K;
    const int n = D.n, nu = D.nu, M = D.M, NA = D.NA, A0 = D.A0, G = D.G;
    const int np = M + 3;
    const DcLayout L = dc_layout(n, k, NA, A0, CS);
    const int pst = L.pst, ld = L.ld, NAp = L.NAp, NA8 = L.NA8;
    float* Ssm = sm + L.oS;                       // [kk][NA][NAp] zero padded
    float* fpl = sm + L.oF + L.HL + L.shF;        // [kk][FR rows][ld]: f, later dL/df (own band); row r of the band at (r - flo) * ld
    float* rsm = sm + L.oR + L.HL + L.shR;        // [RR rows][ld]: row Y at (Y - rlo) * ld
    float* gx = sm + L.oG;                        // [M][16] and derivative [M][16]
    float* gy = gx + 2 * DC_MMAX * 16;
    float* par = sm + L.oPar;                     // [np]
    float* cpart = sm + L.oCp;                    // [CS][32] (rank 0's copy is the one that is read)
    float* ptsacc = sm + L.oPts;                  // [M][32]
    float* ext = sm + L.oEx;                      // [axis][m][4][DC_EXT]
    __shared__ int iwin[2 * DC_MMAX];
    __shared__ const float* remf[DC_CSMAX];       // fpl of every CTA of the cluster (DSMEM)
    __shared__ float* remr[DC_CSMAX];             // rsm of every CTA of the cluster
    __shared__ int remflo[DC_CSMAX], remrlo[DC_CSMAX], remrhi[DC_CSMAX];

    // band of data rows (and of rows of every polyphase plane) owned by this CTA
    const int rpc = (n + CS - 1) / CS;
    const DcBand bd = dc_band(crank, n, rpc, A0, NA);
    const int Y0 = bd.Y0, Y1 = bd.Y1, own = Y1 - Y0, flo = bd.flo, fhi = bd.fhi, rlo = bd.rlo;

    // ---- pending AdaBelief update of the per-epoch parameters (gradients of the previous iteration):
    //      every CTA of the cluster computes it (identical arithmetic), rank 0 stores it after the first cluster barrier
    float* ep = D.ep + (size_t)e * np;
    float upd_p = 0.f, upd_mu = 0.f, upd_nv = 0.f;
    bool upd = false;
    // every global load of the prologue is issued up front (unconditionally: one L2 round trip instead of a chain of three),
    // the planes are cleared while the values are in flight
    float p_in = 0.f, mu_in = 0.f, nv_in = 0.f, g_in = 0.f, pend = 0.f, c_cs = 0.f, c_lr = 0.f, c_b1 = 0.f, c_b2 = 0.f;
    float fu_k = 0.f, fu_a = 0.f, fu_b = 0.f;
    if (tid < np) {
        p_in = ep[tid];
        pend = D.ctl[4]; c_cs = D.ctl[0]; c_lr = D.ctl[1]; c_b1 = D.ctl[2]; c_b2 = D.ctl[3];
        mu_in = D.ep_mu[(size_t)e * np + tid]; nv_in = D.ep_nu[(size_t)e * np + tid]; g_in = D.ep_g[(size_t)e * np + tid];
        if (tid < M && D.lam_fu != 0.f) { fu_k = D.fu[tid]; fu_a = D.fu[DC_MMAX + tid]; fu_b = D.fu[2 * DC_MMAX + tid]; }
    }
    // the folded PSF of the epoch (kk x NA x NAp floats, ~21 KB at cfg4) comes in through the TMA unit: one thread arms an
    // mbarrier and issues ONE bulk copy; the other warps go on building f, and wait just before the forward pass
    __shared__ __align__(8) unsigned long long s_bar;
    if (tid == 0) lcb_mbar_init(&s_bar, 1);
    {   // zero the planes (the gaps between rows are the halos)
        float4* z = reinterpret_cast<float4*>(sm + L.oF);
        for (int i = tid; i < (L.zero_end - L.oF) / 4; i += DC_THREADS) z[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (tid < np) {
        float p = p_in;
        if (pend != 0.f && !noise) {
            const bool is_free = (tid < M) ? D.free_a : (tid < M + 2) ? D.free_d : D.free_mean;
            if (is_free) {
                const BeliefCoef bc = {c_lr, D.cv.b1, D.cv.b2, 1.f - D.cv.b1, 1.f - D.cv.b2, c_b1, c_b2, D.cv.eps, D.cv.eps_root};
                float mu = mu_in, nv = nv_in, g = g_in;
                if (tid < M && D.lam_fu != 0.f)      // flux-uniformity gradient from the global statistics of the last evaluation
                    g += fu_a * (p - fu_k) - fu_b;
                belief_update(bc, c_cs * g, p, mu, nv);
                upd = true; upd_p = p; upd_mu = mu; upd_nv = nv;
            }
        }
        par[tid] = p;
    }
    if (tid < CS) {
        const DcBand bc = dc_band(tid, n, rpc, A0, NA);
        remf[tid] = (CS > 1) ? (const float*)cl.map_shared_rank(fpl, tid) : fpl;
        remr[tid] = (CS > 1) ? (float*)cl.map_shared_rank(rsm, tid) : rsm;
        remflo[tid] = bc.flo; remrlo[tid] = bc.rlo; remrhi[tid] = bc.rhi;
    }
    __syncthreads();                              // (remote writes of r only start after cluster barrier #0 below)
    if (tid == 0) {
        const unsigned bytes = (unsigned)(kk * NA * NAp) * 4u;
        lcb_mbar_expect_tx(&s_bar, bytes);
        lcb_bulk_g2s(Ssm, D.S + (size_t)e * kk * NA * NAp, bytes, &s_bar);
    }
    DC_STAMP(0)

    const int j0 = (D.P - 1) / 2;
    const float delta = 0.5f * (float)(D.P - 1) - (float)j0;
    const float ctr = 0.5f * (float)(nu - 1) - delta;
    const float al = D.alpha[e];
    float sa, ca;
    sincosf(al, &sa, &ca);
    const float dx = par[M], dy = par[M + 1], mean = par[M + 2];
    const Geo geo = {ca, sa, (float)k * dx, (float)k * dy, ctr};

    // ---- point-source windows and taps
    if (tid < 2 * M * G) {
        const int m = (tid / G) % M, t = tid % G, axis = tid / (M * G);
        const float cxm = D.c[m], cym = D.c[M + m];
        const float pc = axis == 0 ? ctr + (float)k * (ca * cxm - sa * cym + dx) : ctr + (float)k * (sa * cxm + ca * cym + dy);
        const int ic = (int)floorf(pc + 0.5f);
        const int u = ic - G / 2 + 1 + t;
        const float x = (float)u - pc;
        const float g = D.cv.gnorm * expf(-x * x * D.cv.inv2s2);
        float* dst = axis == 0 ? gx : gy;
        dst[m * 16 + t] = g;
        dst[DC_MMAX * 16 + m * 16 + t] = x * D.cv.invs2 * g;     // d g / d pc
        if (t == 0) iwin[axis * DC_MMAX + m] = ic - G / 2 + 1;
    }
    // ---- f = warp(h) + point sources in polyphase layout: every CTA builds the rows of its OWN band, then copies the
    //      other rows its forward pass reads from the CTAs that own them (distributed shared memory)
    // pure translation (alpha = 0, the common case): the bilinear taps sit at a constant integer offset with constant weights --
    // f(p) = (1-fy) ((1-fx) h[v0][u0] + fx h[v0][u0+1]) + fy (...), (v0, u0) = p - (bv, bu).  (bu, fx) reproduce floor() and the
    // fraction of the generic form at every shift, integer shifts included (fx = 0 there, so the one-sided derivative of the
    // shift gradient stays the right-sided one).  One warp per row: no per-pixel division, coalesced loads of h.
    const int wrp = tid >> 5, ln = tid & 31;
    const bool transl = (al == 0.f);
    const float t_itx = floorf(geo.tx), t_ity = floorf(geo.ty);
    const float t_wx1 = geo.tx - t_itx, t_wy1 = geo.ty - t_ity;
    const int t_bu = (int)t_itx + (t_wx1 > 0.f ? 1 : 0), t_bv = (int)t_ity + (t_wy1 > 0.f ? 1 : 0);
    const float t_fx = (t_wx1 > 0.f) ? 1.f - t_wx1 : 0.f, t_fy = (t_wy1 > 0.f) ? 1.f - t_wy1 : 0.f;
    if (!noise && own > 0 && transl && D.h != nullptr) {
        for (int row = wrp; row < own * k; row += DC_THREADS / 32) {
            const int v = Y0 * k + row, v0 = v - t_bv;
            const bool okA = (v0 >= 0 && v0 < nu), okB = (v0 + 1 >= 0 && v0 + 1 < nu);
            const float* hA = D.h + (size_t)(okA ? v0 : 0) * nu;
            const float* hB = D.h + (size_t)(okB ? v0 + 1 : 0) * nu;
            float* dst = fpl + ((v % k) * k) * pst + (v / k - flo) * ld;
#pragma unroll 4
            for (int u = ln; u < nu; u += 32) {
                const int u0 = u - t_bu;
                const bool in0 = (u0 >= 0 && u0 < nu), in1 = (u0 + 1 >= 0 && u0 + 1 < nu);
                const float h00 = (okA && in0) ? __ldg(hA + u0) : 0.f, h01 = (okA && in1) ? __ldg(hA + u0 + 1) : 0.f;
                const float h10 = (okB && in0) ? __ldg(hB + u0) : 0.f, h11 = (okB && in1) ? __ldg(hB + u0 + 1) : 0.f;
                dst[(u % k) * pst + u / k] = (1.f - t_fy) * ((1.f - t_fx) * h00 + t_fx * h01) + t_fy * ((1.f - t_fx) * h10 + t_fx * h11);
            }
        }
    } else if (!noise && own > 0) {
#pragma unroll 4
        for (int i = tid; i < own * k * nu; i += DC_THREADS) {     // (unrolled: the L2 loads of four pixels in flight)
            const int v = Y0 * k + i / nu, u = i % nu;
            float val = 0.f;
            if (D.h != nullptr) {
                float qu, qv;
                geo_src(geo, (float)u, (float)v, qu, qv);
                const float fu0 = floorf(qu), fv0 = floorf(qv);
                const float fu = qu - fu0, fv = qv - fv0;
                const int u0 = (int)fu0, v0 = (int)fv0;
                val = (1.f - fv) * ((1.f - fu) * h_at(D.h, nu, v0, u0) + fu * h_at(D.h, nu, v0, u0 + 1)) +
                      fv * ((1.f - fu) * h_at(D.h, nu, v0 + 1, u0) + fu * h_at(D.h, nu, v0 + 1, u0 + 1));
            }
            fpl[((v % k) * k + (u % k)) * pst + (v / k - flo) * ld + u / k] = val;
        }
    }
    __syncthreads();
    if (!noise && own > 0) {
        for (int m = 0; m < M; ++m) {          // serially per source: windows of different sources may overlap
            for (int i = tid; i < G * G; i += DC_THREADS) {
                const int tv = i / G, tu = i % G;
                const int v = iwin[DC_MMAX + m] + tv, u = iwin[m] + tu;
                if (v >= Y0 * k && v < Y1 * k && u >= 0 && u < nu)
                    fpl[((v % k) * k + (u % k)) * pst + (v / k - flo) * ld + u / k] += par[m] * gy[m * 16 + tv] * gx[m * 16 + tu];
            }
            __syncthreads();
        }
    }
    DC_STAMP(1)
    if (CS > 1) {
        cl.sync();                                // #0: planes cleared and own bands of f complete in every CTA
        DC_STAMP(2)
        if (!noise && own > 0) {
            const int nq = n / 4 + 1;             // float4 per row (the tail reads into the zero gap / shift slack)
            const int nrow = (fhi - flo) - own;
            for (int i = tid; i < kk * nrow * nq; i += DC_THREADS) {
                const int q = i % nq, rr = (i / nq) % nrow, ph = i / (nq * nrow);
                const int Yp = (rr < Y0 - flo) ? flo + rr : Y1 + (rr - (Y0 - flo));
                const int oc = min(Yp / rpc, CS - 1);
                // both rows start at a 16-byte aligned address minus the same shift; copy whole aligned quads
                const float* src = remf[oc] - L.shF + ph * pst + (Yp - remflo[oc]) * ld;
                float* dst = fpl - L.shF + ph * pst + (Yp - flo) * ld;
                reinterpret_cast<float4*>(dst)[q] = reinterpret_cast<const float4*>(src)[q];
            }
        }
        __syncthreads();
    }

    DC_STAMP(3)
    // ---- forward: m = mean + 1/k^2 sum_ph corr(f_ph, S_ph);  r = w (m - d).  A task = XB outputs of one row for
    //      one slice of the NA kernel rows (SPLIT slices on adjacent lanes, summed with shuffles) so that a narrow
    //      band still gives every thread a task.
    lcb_mbar_wait(&s_bar, 0);                     // folded PSF landed in shared memory
    DC_STAMP(4)
    const float dscale = D.cv.mean ? 1.f / (float)kk : 1.f;
    const float* dat = D.data + (size_t)e * n * n;
    const float* wgt = D.weight + (size_t)e * n * n;
    float loss = 0.f, gmean = 0.f;
    const int nxb = (n + DC_XB - 1) / DC_XB;
    // outputs per thread of the warp-split forward pass (0 = the lane-split / unsplit form below): the (row, x-block) tasks of the
    // band must be whole warps, divide the CTA, and the plane of partial sums must fit the scratch it borrows
    int fwd_xb = 0;
    {
        auto ok = [&](int xb) {
            const int ntk = own * ((n + xb - 1) / xb);
            return own > 0 && ntk % 32 == 0 && ntk < DC_THREADS && DC_THREADS % ntk == 0 && (n & 3) == 0 && ld >= n + 4 &&
                   own * (n + 4) <= 2 * DC_MMAX * 4 * DC_EXT;
        };
        if (!noise && !(flags & 8) && (n & 15) == 0 && ok(16)) fwd_xb = 16;
        else if (!noise && ok(DC_XB)) fwd_xb = DC_XB;
    }
    if (noise) {
        for (int i = tid; i < n * n; i += DC_THREADS) {
            const int Y = i / n;
            if (Y >= rlo && Y < bd.rhi) rsm[(Y - rlo) * ld + i % n] = __ldg(wgt + i);
        }
    } else if (fwd_xb != 0) {
        // Narrow band (fewer (row, x-block) tasks than threads): the NA kernel rows are split over WHOLE WARPS, so that every tap
        // load of a warp is one broadcast wavefront (with the slices on lane groups of a warp each quarter-warp fetched its own
        // kernel row: 96 shared-memory wavefronts per 264 FFMA, the forward pass ran at the shared-memory bandwidth).  The slices
        // add their partial sums in slice order into a small plane (the scratch of the pts-source term, dead until later):
        // deterministic.  The epilogue then runs on all threads, coalesced along a row, and the band of r goes to the other CTAs of
        // the cluster as 16-byte stores (it was one 4-byte remote store per output and destination, issued by a quarter of the
        // threads).  16 outputs per thread when the stamp side allows it (see corr_line_x).
        const int ldm = n + 4;
        float* msum = ext;                          // [own][n + 4]
        auto conv_w = [&](auto tag) {
            constexpr int XB = decltype(tag)::value;
            const int ntk = own * ((n + XB - 1) / XB), SPL = DC_THREADS / ntk;
            const int s = tid / ntk, t2 = tid % ntk;
            const int Y = Y0 + t2 % own, X0 = (t2 / own) * XB;
            float acc[XB];
#pragma unroll
            for (int x = 0; x < XB; ++x) acc[x] = 0.f;
            {
                // the kk * NA (phase, kernel row) units are dealt out evenly: slice s takes units [s * upw, (s + 1) * upw)
                const int units = kk * NA, upw = (units + SPL - 1) / SPL;
                const int un0 = s * upw, un1 = min(units, un0 + upw);
                const int lo_ia = max(0, -(Y + A0)), hi_ia = min(NA, n - (Y + A0));   // kernel rows whose f row lies inside the stamp
                const bool by_rows = (flags & 32) != 0;                                // development switch: slices of kernel rows, every phase
                const int ias = (NA + SPL - 1) / SPL;
                for (int ph = by_rows ? 0 : un0 / NA; ph < kk && (by_rows || ph * NA < un1); ++ph) {
                    const int ja = by_rows ? max(lo_ia, s * ias) : max(lo_ia, un0 - ph * NA);
                    const int jb = by_rows ? min(hi_ia, (s + 1) * ias) : min(hi_ia, un1 - ph * NA);
                    const float* pl = fpl + ph * pst + X0 + A0 + (Y + A0 - flo) * ld;
                    const float* Sph = Ssm + ph * NA * NAp;
                    if constexpr (XB == DC_XB) { for (int ia = ja; ia < jb; ++ia) corr_line(pl + ia * ld, Sph + ia * NAp, NA, NA8, acc); }
                    else if (NA8 == 32) { for (int ia = ja; ia < jb; ++ia) corr_line_u<XB, 4>(pl + ia * ld, Sph + ia * NAp, NA - NA8, acc); }
                    else { for (int ia = ja; ia < jb; ++ia) corr_line_x<XB>(pl + ia * ld, Sph + ia * NAp, NA, NA8, acc); }
                }
            }
            float4* mrow = reinterpret_cast<float4*>(msum + (Y - Y0) * ldm + X0);
            for (int turn = 0; turn < SPL; ++turn) {
                if (s == turn) {
#pragma unroll
                    for (int q4 = 0; q4 < XB / 4; ++q4) {
                        float4 v = make_float4(acc[4 * q4], acc[4 * q4 + 1], acc[4 * q4 + 2], acc[4 * q4 + 3]);
                        if (turn > 0) { const float4 a = mrow[q4]; v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w; }
                        mrow[q4] = v;
                    }
                }
                __syncthreads();
            }
        };
        if (fwd_xb == 16) conv_w(std::integral_constant<int, 16>{});
        else conv_w(std::integral_constant<int, DC_XB>{});
        for (int i = tid; i < own * n; i += DC_THREADS) {
            const int Yl = i / n, X = i - Yl * n, Yg = Y0 + Yl;
            const float mval = fmaf(dscale, msum[Yl * ldm + X], mean);
            const float d = __ldg(dat + Yg * n + X), w = __ldg(wgt + Yg * n + X);
            const float diff = mval - d, r = w * diff;
            rsm[(Yg - rlo) * ld + X] = r;
            loss = fmaf(r, diff, loss);
            gmean += r;
            if (want_model) D.model[(size_t)e * n * n + Yg * n + X] = mval;
        }
        if (CS > 1) {
            __syncthreads();
            // rows start at a 16-byte aligned address plus the same shift in every CTA: copy whole aligned quads (the floats before
            // and after the n values of a row are zeros of the gap on both sides)
            const int nq = (n + L.shR + 3) / 4;
            const float* srcb = rsm - L.shR;
            for (int i = tid; i < (CS - 1) * own * nq; i += DC_THREADS) {
                const int q = i % nq, Yl = (i / nq) % own, cc = i / (nq * own);
                const int c = cc + (cc >= crank ? 1 : 0), Yg = Y0 + Yl;
                if (Yg >= remrlo[c] && Yg < remrhi[c])
                    reinterpret_cast<float4*>(remr[c] - L.shR + (Yg - remrlo[c]) * ld)[q] = reinterpret_cast<const float4*>(srcb + (Yg - rlo) * ld)[q];
            }
        }
    } else {
        int SPLIT = 1;
        while (SPLIT < 4 && own * nxb * SPLIT < DC_THREADS) SPLIT *= 2;
        const int ias = (NA + SPLIT - 1) / SPLIT;
        const int ntask = own * nxb * SPLIT, ntask_pad = (ntask + 31) & ~31;
        for (int task = tid; task < ntask_pad; task += DC_THREADS) {
            const int lpg = 32 / SPLIT;           // lanes per slice group
            const int s = (task & 31) / lpg, t2 = (task & ~31) / SPLIT + (task & 31) % lpg;
            const bool valid = t2 < own * nxb;
            const int Y = Y0 + (valid ? t2 % own : 0), X0 = (valid ? t2 / own : 0) * DC_XB;
            float acc[DC_XB];
#pragma unroll
            for (int x = 0; x < DC_XB; ++x) acc[x] = 0.f;
            if (valid) {
                const int ia0 = s * ias, ia1 = min(NA, ia0 + ias);
                // kernel rows whose f row Y + A0 + ia lies inside the stamp: a plain loop without per-row branches
                const int ja = max(ia0, -(Y + A0)), jb = min(ia1, n - (Y + A0));
                for (int ph = 0; ph < kk; ++ph) {
                    const float* pl = fpl + ph * pst + X0 + A0 + (Y + A0 - flo) * ld;
                    const float* Sph = Ssm + ph * NA * NAp;
                    for (int ia = ja; ia < jb; ++ia) corr_line(pl + ia * ld, Sph + ia * NAp, NA, NA8, acc);
                }
            }
            for (int o = 32 / SPLIT; o < 32; o <<= 1) {
#pragma unroll
                for (int x = 0; x < DC_XB; ++x) acc[x] += __shfl_xor_sync(0xffffffffu, acc[x], o);
            }
            if (valid && s == 0) {
#pragma unroll
                for (int x = 0; x < DC_XB; ++x) {
                    const int X = X0 + x;
                    if (X < n) {
                        const float mval = fmaf(dscale, acc[x], mean);
                        const float d = __ldg(dat + Y * n + X), w = __ldg(wgt + Y * n + X);
                        const float diff = mval - d, r = w * diff;
                        for (int c = 0; c < CS; ++c)
                            if (Y >= remrlo[c] && Y < remrhi[c]) remr[c][(Y - remrlo[c]) * ld + X] = r;
                        loss = fmaf(r, diff, loss);
                        gmean += r;
                        if (want_model) D.model[(size_t)e * n * n + Y * n + X] = mval;
                    }
                }
            }
        }
    
    }
    DC_STAMP(5)
    cl.sync();                                    // #1: r complete in every CTA; all CTAs have read the old parameters
    DC_STAMP(6)
    if (upd && crank == 0) {
        ep[tid] = upd_p; D.ep_mu[(size_t)e * np + tid] = upd_mu; D.ep_nu[(size_t)e * np + tid] = upd_nv;
    }
    // ---- adjoint for the own band: dL/df_ph[Y'][X'] = 1/k^2 sum S_ph[av][au] r[Y'-av][X'-au]   (overwrites f)
    const int ntkA16 = kk * own * (n / 16);       // tasks with 16 outputs per thread
    if (!noise && !(flags & 16) && (n & 15) == 0 && ntkA16 > 0 && ntkA16 <= DC_THREADS && ntkA16 % 32 == 0 && DC_THREADS % ntkA16 == 0) {
        // 16 outputs per thread (conv_line_T_x); when that leaves fewer tasks than threads the NA kernel rows are split over whole
        // warps and the slices add into the plane in slice order
        constexpr int XB = 16;
        const int nx16 = n / XB, SPA = DC_THREADS / ntkA16, iasA = (NA + SPA - 1) / SPA;
        const int s = tid / ntkA16, task = tid % ntkA16;
        const int ph = task / (own * nx16), rem = task % (own * nx16), Y = Y0 + rem % own, X0 = (rem / own) * XB;
        float acc[XB];
#pragma unroll
        for (int x = 0; x < XB; ++x) acc[x] = 0.f;
        const float* Sph = Ssm + ph * NA * NAp;
        const float* rrow = rsm + X0 - A0 - (NA8 - 1) + (Y - A0 - rlo) * ld;
        const int ja = max(s * iasA, Y - A0 - (n - 1)), jb = min(min(NA, (s + 1) * iasA), Y - A0 + 1);
        if (NA8 == 32) { for (int ia = ja; ia < jb; ++ia) conv_line_T_u<XB, 4>(rrow - ia * ld, Sph + ia * NAp, NA - NA8, acc); }
        else { for (int ia = ja; ia < jb; ++ia) conv_line_T_x<XB>(rrow - ia * ld, Sph + ia * NAp, NA, NA8, acc); }
        float* dst = fpl + ph * pst + (Y - flo) * ld + X0;
        for (int turn = 0; turn < SPA; ++turn) {
            if (s == turn) {
#pragma unroll
                for (int x = 0; x < XB; ++x) dst[x] = (turn > 0 ? dst[x] : 0.f) + dscale * acc[x];
            }
            if (turn + 1 < SPA) __syncthreads();
        }
    } else
    for (int task = tid; task < kk * own * nxb; task += DC_THREADS) {
        const int ph = task / (own * nxb), rem = task % (own * nxb), Y = Y0 + rem % own, X0 = (rem / own) * DC_XB;
        float acc[DC_XB];
#pragma unroll
        for (int x = 0; x < DC_XB; ++x) acc[x] = 0.f;
        const float* Sph = Ssm + ph * NA * NAp;
        const float* rbase = rsm + X0 - A0 - (NA8 - 1);
        // kernel rows whose r row Y - A0 - ia lies inside the stamp
        const int ja = max(0, Y - A0 - (n - 1)), jb = min(NA, Y - A0 + 1);
        const float* rrow = rbase + (Y - A0 - rlo) * ld;
        if (noise) { for (int ia = ja; ia < jb; ++ia) conv_line_T<true>(rrow - ia * ld, Sph + ia * NAp, NA, NA8, acc); }
        else { for (int ia = ja; ia < jb; ++ia) conv_line_T<false>(rrow - ia * ld, Sph + ia * NAp, NA, NA8, acc); }
#pragma unroll
        for (int x = 0; x < DC_XB; ++x)
            if (X0 + x < n) fpl[ph * pst + (Y - flo) * ld + X0 + x] = (noise ? dscale * dscale : dscale) * acc[x];
    }
    __syncthreads();
    DC_STAMP(7)
    const int vlo = Y0 * k, vhi = Y1 * k;         // rows of f whose dL/df lives in this CTA
    const float sc = (D.cv.half == 0.5f) ? 1.f : 2.f;
    const int warp = tid >> 5, lane = tid & 31;
    float* cp0 = (CS > 1) ? (float*)cl.map_shared_rank(cpart, 0) : cpart;
    // ---- point-source gradients: one warp per source (M <= 8 warps), own rows only
    if (warp < M && !noise) {
        const int m = warp;
        float ga = 0.f, gu = 0.f, gv = 0.f;
        for (int i = lane; i < G * G; i += 32) {
            const int tv = i / G, tu = i % G;
            const int v = iwin[DC_MMAX + m] + tv, u = iwin[m] + tu;
            if (v >= vlo && v < vhi && u >= 0 && u < nu) {
                const float df = fpl[((v % k) * k + (u % k)) * pst + (v / k - flo) * ld + u / k];
                const float gyv = gy[m * 16 + tv], gxv = gx[m * 16 + tu];
                ga = fmaf(df, gyv * gxv, ga);
                gu = fmaf(df, gyv * gx[DC_MMAX * 16 + m * 16 + tu], gu);
                gv = fmaf(df, gy[DC_MMAX * 16 + m * 16 + tv] * gxv, gv);
            }
        }
        ga = warp_sum(ga); gu = warp_sum(gu); gv = warp_sum(gv);
        if (lane == 0) {
            cp0[crank * 32 + 4 + 3 * m] = ga;
            cp0[crank * 32 + 4 + 3 * m + 1] = gu;     // dL/d uc_m / a_m
            cp0[crank * 32 + 4 + 3 * m + 2] = gv;     // dL/d vc_m / a_m
        }
    }
    // ---- shift gradient through the warp: dL/dd = sum_p dL/df[p] * grad h(q(p)) . dq/dd   (own rows)
    float gwx = 0.f, gwy = 0.f;
    if (D.free_d && !noise && D.h != nullptr && transl) {
        for (int row = wrp; row < vhi - vlo; row += DC_THREADS / 32) {
            const int v = vlo + row, v0 = v - t_bv;
            const bool okA = (v0 >= 0 && v0 < nu), okB = (v0 + 1 >= 0 && v0 + 1 < nu);
            const float* hA = D.h + (size_t)(okA ? v0 : 0) * nu;
            const float* hB = D.h + (size_t)(okB ? v0 + 1 : 0) * nu;
            const float* dfrow = fpl + ((v % k) * k) * pst + (v / k - flo) * ld;
#pragma unroll 4
            for (int u = ln; u < nu; u += 32) {
                const int u0 = u - t_bu;
                const bool in0 = (u0 >= 0 && u0 < nu), in1 = (u0 + 1 >= 0 && u0 + 1 < nu);
                const float h00 = (okA && in0) ? __ldg(hA + u0) : 0.f, h01 = (okA && in1) ? __ldg(hA + u0 + 1) : 0.f;
                const float h10 = (okB && in0) ? __ldg(hB + u0) : 0.f, h11 = (okB && in1) ? __ldg(hB + u0 + 1) : 0.f;
                const float dhu = (1.f - t_fy) * (h01 - h00) + t_fy * (h11 - h10);
                const float dhv = (1.f - t_fx) * (h10 - h00) + t_fx * (h11 - h01);
                const float df = dfrow[(u % k) * pst + u / k];
                gwx = fmaf(df, -(float)k * dhu, gwx);      // ca = 1, sa = 0
                gwy = fmaf(df, -(float)k * dhv, gwy);
            }
        }
    } else if (D.free_d && !noise && D.h != nullptr) {
#pragma unroll 4
        for (int i = tid; i < (vhi - vlo) * nu; i += DC_THREADS) {
            const int v = vlo + i / nu, u = i % nu;
            float qu, qv;
            geo_src(geo, (float)u, (float)v, qu, qv);
            const float fu0 = floorf(qu), fv0 = floorf(qv);
            const float fu = qu - fu0, fv = qv - fv0;
            const int u0 = (int)fu0, v0 = (int)fv0;
            const float h00 = h_at(D.h, nu, v0, u0), h01 = h_at(D.h, nu, v0, u0 + 1);
            const float h10 = h_at(D.h, nu, v0 + 1, u0), h11 = h_at(D.h, nu, v0 + 1, u0 + 1);
            const float dhu = (1.f - fv) * (h01 - h00) + fv * (h11 - h10);
            const float dhv = (1.f - fu) * (h10 - h00) + fu * (h11 - h01);
            const float df = fpl[((v % k) * k + (u % k)) * pst + (v / k - flo) * ld + u / k];
            // dq/ddx = -k (ca, -sa),  dq/ddy = -k (sa, ca)
            gwx = fmaf(df, -(float)k * (ca * dhu - sa * dhv), gwx);
            gwy = fmaf(df, -(float)k * (sa * dhu + ca * dhv), gwy);
        }
    }
    {   // four block sums with one pair of barriers (warp order fixed: deterministic)
        __shared__ float red4[(DC_THREADS / 32) * 4];
        loss = warp_sum(loss); gmean = warp_sum(gmean); gwx = warp_sum(gwx); gwy = warp_sum(gwy);
        if (lane == 0) { red4[warp * 4] = loss; red4[warp * 4 + 1] = gmean; red4[warp * 4 + 2] = gwx; red4[warp * 4 + 3] = gwy; }
        __syncthreads();
        if (tid < 4) {
            float s4 = 0.f;
            for (int w = 0; w < DC_THREADS / 32; ++w) s4 += red4[w * 4 + tid];
            cp0[crank * 32 + tid] = s4;
        }
    }

    DC_STAMP(8)
    // ---- pts-source regulariser (rank 0): lam * sum_x W_0[x] |alpha_0(p_e)[x]|, p_e = point-source channel of the
    //      epoch, alpha_0 = first starlet scale.  The Gaussians are separable, so alpha_0(g_m)(v,u) =
    //      g_y(v) g_x(u) - (B g_y)(v) (B g_x)(u) with B the edge-replicated 5-tap B3 filter; by linearity
    //      dR/dtheta = sum_x T[x] alpha_0(dp/dtheta)[x], T = lam W_0 sign(alpha_0(p)).
    const float lam_pts = (noise || (!D.pts_all_epochs && e + D.e0 != 0)) ? 0.f : D.lam_pts;
    const int GEX = G + 4;
    if (crank == 0 && lam_pts != 0.f) {
        if (tid < 2 * M * GEX) {
            const int axis = tid / (M * GEX), m = (tid / GEX) % M, t = tid % GEX;
            const int w0 = iwin[axis * DC_MMAX + m];
            const float* g0 = (axis == 0 ? gx : gy) + m * 16;
            const float* g1 = g0 + DC_MMAX * 16;
            const int u = w0 - 2 + t;
            const float B[5] = {1.f / 16.f, 4.f / 16.f, 6.f / 16.f, 4.f / 16.f, 1.f / 16.f};
            float bg = 0.f, bd = 0.f;
#pragma unroll
            for (int tt = -2; tt <= 2; ++tt) {
                const int uu = min(max(u + tt, 0), nu - 1) - w0;
                if (uu >= 0 && uu < G) { bg = fmaf(B[tt + 2], g0[uu], bg); bd = fmaf(B[tt + 2], g1[uu], bd); }
            }
            const bool in = (t >= 2 && t < G + 2);
            float* ex = ext + (axis * DC_MMAX + m) * 4 * DC_EXT;
            ex[t] = in ? g0[t - 2] : 0.f;
            ex[DC_EXT + t] = bg;
            ex[2 * DC_EXT + t] = in ? g1[t - 2] : 0.f;
            ex[3 * DC_EXT + t] = bd;
        }
        __syncthreads();
        // H warps per source (H * M <= 8): warp w walks every H-th group of 32 pixels of the window of source w % M
        const int H = (M > 0) ? (DC_THREADS / 32) / M : 0;
        if (warp < H * M) {
            const int m = warp % M;
            float pl = 0.f, pa[DC_MMAX], pu[DC_MMAX], pv[DC_MMAX];
#pragma unroll
            for (int q = 0; q < DC_MMAX; ++q) pa[q] = pu[q] = pv[q] = 0.f;
            for (int i = lane + 32 * (warp / M); i < GEX * GEX; i += 32 * H) {
                const int v = iwin[DC_MMAX + m] - 2 + i / GEX, u = iwin[m] - 2 + i % GEX;
                if (v < 0 || v >= nu || u < 0 || u >= nu) continue;
                bool dup = false;                  // pixel already counted by a source of lower index
                for (int q = 0; q < m; ++q) {
                    const int tv = v - (iwin[DC_MMAX + q] - 2), tu = u - (iwin[q] - 2);
                    dup = dup || (tv >= 0 && tv < GEX && tu >= 0 && tu < GEX);
                }
                if (dup) continue;
                float al0 = 0.f, ca_[DC_MMAX], cu_[DC_MMAX], cv_[DC_MMAX];
#pragma unroll
                for (int q = 0; q < DC_MMAX; ++q) {
                    ca_[q] = cu_[q] = cv_[q] = 0.f;
                    if (q < M) {
                        const int tv = v - (iwin[DC_MMAX + q] - 2), tu = u - (iwin[q] - 2);
                        if (tv >= 0 && tv < GEX && tu >= 0 && tu < GEX) {
                            const float* ex = ext + q * 4 * DC_EXT;
                            const float* ey = ext + (DC_MMAX + q) * 4 * DC_EXT;
                            const float gxu = ex[tu], bxu = ex[DC_EXT + tu], dxu = ex[2 * DC_EXT + tu], bdxu = ex[3 * DC_EXT + tu];
                            const float gyv = ey[tv], byv = ey[DC_EXT + tv], dyv = ey[2 * DC_EXT + tv], bdyv = ey[3 * DC_EXT + tv];
                            ca_[q] = gyv * gxu - byv * bxu;
                            cu_[q] = gyv * dxu - byv * bdxu;
                            cv_[q] = dyv * gxu - bdyv * bxu;
                            al0 = fmaf(par[q], ca_[q], al0);
                        }
                    }
                }
                const float lw = lam_pts * (D.W ? __ldg(D.W + (size_t)v * nu + u) : 1.f);
                pl = fmaf(lw, fabsf(al0), pl);
                const float T = (al0 > 0.f) ? lw : (al0 < 0.f) ? -lw : 0.f;
#pragma unroll
                for (int q = 0; q < DC_MMAX; ++q) { pa[q] = fmaf(T, ca_[q], pa[q]); pu[q] = fmaf(T, cu_[q], pu[q]); pv[q] = fmaf(T, cv_[q], pv[q]); }
            }
            pl = warp_sum(pl);
            if (lane == 0) ptsacc[warp * 32] = pl;
#pragma unroll
            for (int q = 0; q < DC_MMAX; ++q) {
                const float s0 = warp_sum(pa[q]), s1 = warp_sum(pu[q]), s2 = warp_sum(pv[q]);
                if (lane == 0 && q < M) { ptsacc[warp * 32 + 1 + 3 * q] = s0; ptsacc[warp * 32 + 2 + 3 * q] = s1; ptsacc[warp * 32 + 3 + 3 * q] = s2; }
            }
        }
    }
    DC_STAMP(9)
    cl.sync();                                    // #2: dL/df bands and partial sums complete
    DC_STAMP(10)
    float* epg = D.ep_g + (size_t)e * np;
    if (crank == 0 && tid == 0 && !noise) {
        float ls = 0.f, gm = 0.f, wx = 0.f, wy = 0.f, ga[DC_MMAX], gu[DC_MMAX], gv[DC_MMAX];
        for (int m = 0; m < M; ++m) ga[m] = gu[m] = gv[m] = 0.f;
        for (int c = 0; c < CS; ++c) {            // fixed order: the result does not depend on scheduling
            const float* q = cpart + c * 32;
            ls += q[0]; gm += q[1]; wx += q[2]; wy += q[3];
            for (int m = 0; m < M; ++m) { ga[m] += q[4 + 3 * m]; gu[m] += q[5 + 3 * m]; gv[m] += q[6 + 3 * m]; }
        }
        float ploss = 0.f;
        const int NWP = (M > 0) ? ((DC_THREADS / 32) / M) * M : 0;   // warps that took part in the pts-source term
        float gdx = sc * wx, gdy = sc * wy;
        for (int m = 0; m < M; ++m) {
            float pa = 0.f, pu = 0.f, pv = 0.f;
            if (lam_pts != 0.f)
                for (int w = 0; w < NWP; ++w) { pa += ptsacc[w * 32 + 1 + 3 * m]; pu += ptsacc[w * 32 + 2 + 3 * m]; pv += ptsacc[w * 32 + 3 + 3 * m]; }
            const float a = par[m];
            const float GU = a * (sc * gu[m] + pu), GV = a * (sc * gv[m] + pv);
            gdx += (float)k * GU; gdy += (float)k * GV;
            D.gc[(size_t)e * 2 * M + m] = (float)k * (ca * GU + sa * GV);
            D.gc[(size_t)e * 2 * M + M + m] = (float)k * (-sa * GU + ca * GV);
            epg[m] = sc * ga[m] + pa;
        }
        if (lam_pts != 0.f) for (int w = 0; w < NWP; ++w) ploss += ptsacc[w * 32];
        epg[M] = gdx; epg[M + 1] = gdy; epg[M + 2] = sc * gm;
        D.eloss[e] = D.cv.half * ls + ploss;
    }
    // ---- transposed warp as an exact gather for the own rows of h: dL/dh[q] = sum_p dL/df[p] * hat(q - q(p));
    //      dL/df of the other bands is read through distributed shared memory
    if (D.free_h || noise) {
        float* Gh = D.Gh + (size_t)e * nu * nu;
        auto dfr = [&](int pv, int pu) -> float {
            if (pv < 0 || pv >= nu || pu < 0 || pu >= nu) return 0.f;
            const int Yp = pv / k, oc = min(Yp / rpc, CS - 1);
            return remf[oc][((pv % k) * k + (pu % k)) * pst + (Yp - remflo[oc]) * ld + pu / k];
        };
        if (al == 0.f) {
            // pure translation t = it + ft: f(p) = ft h[p-it-1] + (1-ft) h[p-it] per axis, so
            // dL/dh[q] = ft dL/df[q+it+1] + (1-ft) dL/df[q+it] (separable, constant weights)
            const float itx = floorf(geo.tx), ity = floorf(geo.ty);
            float wx1 = geo.tx - itx, wy1 = geo.ty - ity, wx0 = 1.f - wx1, wy0 = 1.f - wy1;
            if (noise) { wx0 *= wx0; wx1 *= wx1; wy0 *= wy0; wy1 *= wy1; }
            const int iu = (int)itx, iv = (int)ity;
            // one warp per row of h: the two rows of dL/df it gathers from (own band or a neighbour's, through distributed shared
            // memory) are resolved once per row, the lanes walk the columns with shifts and masks only
            for (int row = wrp; row < vhi - vlo; row += DC_THREADS / 32) {
                const int qv_i = vlo + row, pv = qv_i + iv;
                auto rowptr = [&](int pvv) -> const float* {
                    if (pvv < 0 || pvv >= nu) return nullptr;
                    const int Yp = pvv / k, oc = min(Yp / rpc, CS - 1);
                    return ((oc == crank) ? (const float*)fpl : remf[oc]) + ((pvv % k) * k) * pst + (Yp - remflo[oc]) * ld;
                };
                const float* r0 = rowptr(pv);
                const float* r1 = rowptr(pv + 1);
#pragma unroll 4
                for (int qu_i = ln; qu_i < nu; qu_i += 32) {
                    const int pu = qu_i + iu;
                    const bool in0 = (pu >= 0 && pu < nu), in1 = (pu + 1 >= 0 && pu + 1 < nu);
                    const int o0 = in0 ? (pu % k) * pst + pu / k : 0, o1 = in1 ? ((pu + 1) % k) * pst + (pu + 1) / k : 0;
                    const float d00 = (r0 && in0) ? r0[o0] : 0.f, d01 = (r0 && in1) ? r0[o1] : 0.f;
                    const float d10 = (r1 && in0) ? r1[o0] : 0.f, d11 = (r1 && in1) ? r1[o1] : 0.f;
                    const float acc = wy0 * (wx0 * d00 + wx1 * d01) + wy1 * (wx0 * d10 + wx1 * d11);
                    Gh[qv_i * nu + qu_i] = noise ? acc : sc * acc;
                }
            }
        } else {
            for (int i = tid; i < (vhi - vlo) * nu; i += DC_THREADS) {
                const int qv_i = vlo + i / nu, qu_i = i % nu;
                // forward image of q: p_c = R (q - ctr) + ctr + k d
                const float ru = (float)qu_i - ctr, rv = (float)qv_i - ctr;
                const float pcu = ca * ru - sa * rv + ctr + geo.tx, pcv = sa * ru + ca * rv + ctr + geo.ty;
                const int pu0 = (int)floorf(pcu) - 1, pv0 = (int)floorf(pcv) - 1;
                float acc = 0.f;
                for (int b = 0; b < 4; ++b)
                    for (int a2 = 0; a2 < 4; ++a2) {
                        const int pv = pv0 + b, pu = pu0 + a2;
                        if (pv < 0 || pv >= nu || pu < 0 || pu >= nu) continue;
                        float qu, qv;
                        geo_src(geo, (float)pu, (float)pv, qu, qv);
                        // bilinear weight of tap q for source position (qu,qv): taps are floor(q), floor(q)+1
                        const float fu0 = floorf(qu), fv0 = floorf(qv);
                        float wu = 0.f, wv = 0.f;
                        if ((int)fu0 == qu_i) wu = 1.f - (qu - fu0); else if ((int)fu0 + 1 == qu_i) wu = qu - fu0;
                        if ((int)fv0 == qv_i) wv = 1.f - (qv - fv0); else if ((int)fv0 + 1 == qv_i) wv = qv - fv0;
                        if (noise) { wu *= wu; wv *= wv; }
                        if (wu != 0.f && wv != 0.f) acc = fmaf(wu * wv, dfr(pv, pu), acc);
                    }
                Gh[i + vlo * nu] = noise ? acc : sc * acc;
            }
        }
    }
    DC_STAMP(11)
    cl.sync();                                    // #3: no CTA leaves while its shared memory may still be read
    DC_STAMP(12)
#ifdef LCB_DC_TIMERS
    if (tid == 0 && blockIdx.x < DC_TIM_CTAS) { unsigned long long gt; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt)); g_dc_tim[blockIdx.x][15] = (long long)gt; if (flags & 4) g_dc_tim[blockIdx.x][13] = 0; }
#endif
    if (!(flags & 4)) return;

    // ---- fused reduction over the local epochs (replaces k_deconv_reduce inside lcb_deconv_run), two levels so that no CTA
    //      walks more than ~2 sqrt(E) planes: the last cluster to finish a band of a GROUP of GS consecutive epochs sums the
    //      group (epoch order) into Gp[group]; the last group to finish a band sums the NG partial planes (group order) and
    //      emits them.  Fixed summation order: deterministic, identical on every rank.
    __shared__ int s_last;
    const int grp = e / D.GS, g_lo = grp * D.GS, g_hi = min(g_lo + D.GS, D.E);
    __threadfence();                              // this CTA's rows of Gh (and, rank 0, the per-epoch scalars) are visible device wide
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(D.grp_ctr + grp * DC_CSMAX + crank, 1u) == (unsigned)(g_hi - g_lo - 1)) ? 1 : 0;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const int seq = (D.cm.world > 1) ? ((seq_arg < 0) ? D.ictl[1] : seq_arg) : 0;
    const int Wn = D.cm.world, parity = seq & 1, nu2 = nu * nu;
    auto emit = [&](int i, float v) {
        if (seq == 0) D.red[i] = v;
        else for (int r = 0; r < Wn; ++r) D.cm.slots[r][((size_t)parity * Wn + D.cm.rank) * D.tot_pad + i] = v;
    };
    // four consecutive entries (i a multiple of 4): ONE 16-byte store per peer instead of four 4-byte stores at a 16-byte stride
    auto emit4 = [&](int i, float4 v) {
        if (seq == 0) *reinterpret_cast<float4*>(D.red + i) = v;
        else for (int r = 0; r < Wn; ++r)
            *reinterpret_cast<float4*>(D.cm.slots[r] + ((size_t)parity * Wn + D.cm.rank) * D.tot_pad + i) = v;
    };
    const int cnt = (vhi - vlo) * nu;
    // sums planes src[p * nu2 + i], p = 0 .. np_-1 (in order) over the pixels of the band.  The walk is bound by L2 latency, so
    // every thread keeps 4 quads x 4 planes (16 independent 16-byte loads) in flight; rows that are not a multiple of 4 pixels
    // fall back to scalar loads
    auto band_sum = [&](const float* __restrict__ src, int np_, auto&& put, auto&& put4) {
        if ((nu & 3) == 0) {
            const int cnt4 = cnt >> 2;
            // NQ quads x PE planes = 16 independent 16-byte loads in flight per thread; a narrow band (two quads per thread) takes
            // eight planes per trip instead of four.  The planes of a pixel are added in plane order either way.
            auto walk = [&](auto nq_t, auto pe_t) {
                constexpr int NQ = decltype(nq_t)::value, PE = decltype(pe_t)::value;
                for (int i0 = tid; i0 < cnt4; i0 += NQ * DC_THREADS) {
                    float4 acc[NQ];
#pragma unroll
                    for (int q = 0; q < NQ; ++q) acc[q] = make_float4(0.f, 0.f, 0.f, 0.f);
                    for (int p0 = 0; p0 < np_; p0 += PE) {
                        float4 v[NQ][PE];
#pragma unroll
                        for (int q = 0; q < NQ; ++q)
#pragma unroll
                            for (int pe = 0; pe < PE; ++pe) {
                                const int i = i0 + q * DC_THREADS;
                                v[q][pe] = (i < cnt4 && p0 + pe < np_) ? __ldcg(reinterpret_cast<const float4*>(src + (size_t)(p0 + pe) * nu2) + i)
                                                                       : make_float4(0.f, 0.f, 0.f, 0.f);
                            }
#pragma unroll
                        for (int q = 0; q < NQ; ++q)
#pragma unroll
                            for (int pe = 0; pe < PE; ++pe) {
                                acc[q].x += v[q][pe].x; acc[q].y += v[q][pe].y; acc[q].z += v[q][pe].z; acc[q].w += v[q][pe].w;
                            }
                    }
#pragma unroll
                    for (int q = 0; q < NQ; ++q) {
                        const int i = i0 + q * DC_THREADS;
                        if (i < cnt4) put4(vlo * nu + 4 * i, acc[q]);
                    }
                }
            };
            if (cnt4 <= 2 * DC_THREADS) walk(std::integral_constant<int, 2>{}, std::integral_constant<int, 8>{});
            else walk(std::integral_constant<int, 4>{}, std::integral_constant<int, 4>{});
            return;
        }
        for (int i0 = tid; i0 < cnt; i0 += 4 * DC_THREADS) {
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
            for (int p0 = 0; p0 < np_; p0 += 4) {
                float v[4][4];
#pragma unroll
                for (int q = 0; q < 4; ++q)
#pragma unroll
                    for (int pe = 0; pe < 4; ++pe) {
                        const int i = i0 + q * DC_THREADS;
                        v[q][pe] = (i < cnt && p0 + pe < np_) ? __ldcg(src + (size_t)(p0 + pe) * nu2 + i) : 0.f;
                    }
#pragma unroll
                for (int q = 0; q < 4; ++q)
#pragma unroll
                    for (int pe = 0; pe < 4; ++pe) acc[q] += v[q][pe];
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) { const int i = i0 + q * DC_THREADS; if (i < cnt) put(vlo * nu + i, acc[q]); }
        }
    };
    const bool one_level = (D.NG == 1);
    if (D.free_h) {
        const float* src = D.Gh + (size_t)g_lo * nu2 + (size_t)vlo * nu;
        float* dst = D.Gp + (size_t)grp * nu2;
        if (one_level) band_sum(src, g_hi - g_lo, emit, emit4);
        else band_sum(src, g_hi - g_lo, [&](int i, float v) { dst[i] = v; }, [&](int i, float4 v) { *reinterpret_cast<float4*>(dst + i) = v; });
    } else if (one_level) {
        for (int i = tid; i < cnt; i += DC_THREADS) emit(vlo * nu + i, 0.f);
    }
    if (tid == 0) D.grp_ctr[grp * DC_CSMAX + crank] = 0;       // next use: the next launch of this kernel
    if (!one_level) {
        __threadfence();
        __syncthreads();
        if (tid == 0) s_last = (atomicAdd(D.band_ctr + crank, 1u) == (unsigned)(D.NG - 1)) ? 1 : 0;
        __syncthreads();
        if (!s_last) return;
        __threadfence();
        if (D.free_h) band_sum(D.Gp + (size_t)vlo * nu, D.NG, emit, emit4);
        else for (int i = tid; i < cnt; i += DC_THREADS) emit(vlo * nu + i, 0.f);
    }
    if (crank == 0) {
        // per-epoch scalars (every rank-0 CTA wrote its own before raising its counter): 8 lanes per entry stride the epochs (all
        // loads of an entry in flight at once), fixed-order shuffle tree over the 8 lanes -- one pass for up to 32 entries
        const int sub = tid & 7, iflux = red_flux(D);
        for (int i = nu2 + (tid >> 3); i < ((D.tot - nu2 + 31) & ~31) + nu2; i += DC_THREADS / 8) {
            float sacc = 0.f;
            auto walk = [&](auto term) {           // the terms of 4 epochs per lane are evaluated before the first add
                for (int ee = sub; ee < D.E; ee += 32) {
                    float t[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) t[q] = (ee + 8 * q < D.E) ? term(ee + 8 * q) : 0.f;
#pragma unroll
                    for (int q = 0; q < 4; ++q) sacc += t[q];
                }
            };
            if (i < nu2 + 2 * M) {
                walk([&](int ee) { return __ldcg(D.gc + (size_t)ee * 2 * M + (i - nu2)); });
            } else if (i == nu2 + 2 * M) {
                walk([&](int ee) { return __ldcg(D.eloss + ee); });
            } else if (i == nu2 + 2 * M + 1) {
                walk([&](int ee) {
                    float g2 = 0.f;
                    for (int p = 0; p < np; ++p) {
                        const bool is_free = (p < M) ? D.free_a : (p < M + 2) ? D.free_d : D.free_mean;
                        const float g = __ldcg(D.ep_g + (size_t)ee * np + p);
                        if (is_free) g2 = fmaf(g, g, g2);
                    }
                    return g2;
                });
            } else if (i < iflux + 4 * M) {
                const int q = (i - iflux) / M, m = (i - iflux) % M;
                const float Kf = D.fu[m];
                walk([&](int ee) {
                    const float a = __ldcg(D.ep + (size_t)ee * np + m) - Kf, g = __ldcg(D.ep_g + (size_t)ee * np + m);
                    return (q == 0) ? a : (q == 1) ? a * a : (q == 2) ? g : g * a;
                });
            }
            sacc += __shfl_xor_sync(0xffffffffu, sacc, 1);
            sacc += __shfl_xor_sync(0xffffffffu, sacc, 2);
            sacc += __shfl_xor_sync(0xffffffffu, sacc, 4);
            if (sub == 0 && i < D.tot) emit(i, sacc);
        }
    }
    DC_STAMP(13)
    __threadfence_system();
    __syncthreads();
    if (tid == 0) {
        D.band_ctr[crank] = 0;                      // next use: the next launch of this kernel
        if (atomicAdd(D.band_ctr + DC_CSMAX, 1u) == (unsigned)(CS - 1)) {
            D.band_ctr[DC_CSMAX] = 0;
            __threadfence_system();
            if (seq > 0) for (int r = 0; r < Wn; ++r) *(volatile int*)(D.cm.flags[r] + parity * Wn + D.cm.rank) = seq;
        }
    }
}

// ---------------------------------------------------------------- reduction over the local epochs (+ push to the peers)
// seq == 0: red[] <- local sums.  seq > 0 (multi-GPU): the local sums go to slot [seq&1][my rank] of EVERY rank's
// receive buffer (plain stores into peer memory over NVLink); the last CTA to finish raises flag [seq&1][my rank] = seq
// on every rank (fence - atomic - fence - flag, the threadFenceReduction pattern at system scope).
__global__ void __launch_bounds__(256) k_deconv_reduce(DeconvDev D, int force_h, int seq_arg) {
    // seq_arg < 0: the sequence number is read from D.ictl[1] (CUDA-graph replay), as in k_deconv_epoch
    const int seq = (seq_arg < 0) ? ((D.cm.world > 1) ? D.ictl[1] : 0) : seq_arg;
    // 64 entries of red[] per CTA, 4 threads per entry: thread g sums the epochs e = g, g+4, ... (independent loads in
    // flight), the four partial sums are combined in a fixed order -> deterministic
    __shared__ __align__(16) float part[4][64];
    const int nu2 = D.nu * D.nu, M = D.M, np = D.M + 3;
    const int col = threadIdx.x & 63, grp = threadIdx.x >> 6;
    const int i = blockIdx.x * 64 + col;
    const int iflux = red_flux(D);
    float s = 0.f;
    if (i < nu2) {
        if (D.free_h || force_h) {
            // 16 planes per trip, all loads issued before the first add (the walk is bound by L2 latency: four loads in flight made
            // it 27 us at 100 local epochs); the planes of a thread are added in epoch order, absent ones add +0
            const float* src = D.Gh + i;
            for (int e = grp; e < D.E; e += 64) {
                float a[16];
#pragma unroll
                for (int q = 0; q < 16; ++q) a[q] = (e + 4 * q < D.E) ? __ldcg(src + (size_t)(e + 4 * q) * nu2) : 0.f;
#pragma unroll
                for (int q = 0; q < 16; ++q) s += a[q];
            }
        }
    }
    // Per-epoch scalars (entries nu2 .. tot-1: gradients of c, loss, |g_epoch|^2, flux sums).  The CTA(s) that own them compute them
    // COOPERATIVELY: thread t takes the epochs t, t + 256, ..., loads the 4M + 4 values of an epoch once (all loads in flight), forms
    // the 6M + 2 terms, and the CTA reduces them in a fixed order (shuffle tree per warp, warps in order).  Four threads per entry
    // walking E / 4 epochs with one L2 round trip each made this CTA the last one to finish: the kernel took 8 us + 0.9 us per 4
    // local epochs (53 us at 200).
    __shared__ float sc_red[8][6 * DC_MMAX + 2];
    const int nsc = D.tot - nu2;
    const bool owns_scalars = (int)(blockIdx.x * 64 + 63) >= nu2;      // block-uniform
    if (owns_scalars) {
        float t[6 * DC_MMAX + 2];
#pragma unroll
        for (int k = 0; k < 6 * DC_MMAX + 2; ++k) t[k] = 0.f;
        for (int e = threadIdx.x; e < D.E; e += 256) {
            float gcv[2 * DC_MMAX], gv[DC_MMAX + 3], av[DC_MMAX];
#pragma unroll
            for (int q = 0; q < 2 * DC_MMAX; ++q) gcv[q] = (q < 2 * M) ? __ldcg(D.gc + (size_t)e * 2 * M + q) : 0.f;
#pragma unroll
            for (int q = 0; q < DC_MMAX + 3; ++q) gv[q] = (q < np) ? __ldcg(D.ep_g + (size_t)e * np + q) : 0.f;
#pragma unroll
            for (int q = 0; q < DC_MMAX; ++q) av[q] = (q < M) ? __ldcg(D.ep + (size_t)e * np + q) - D.fu[q] : 0.f;
            const float el = __ldcg(D.eloss + e);
            float g2 = 0.f;
#pragma unroll
            for (int q = 0; q < DC_MMAX + 3; ++q) {
                const bool is_free = (q < M) ? D.free_a : (q < M + 2) ? D.free_d : D.free_mean;
                if (q < np && is_free) g2 = fmaf(gv[q], gv[q], g2);
            }
            // entry order of red[] past the pixels: gc[2M] | loss | |g|^2 | sum (a-K)[M] | sum (a-K)^2[M] | sum g[M] | sum g (a-K)[M]
#pragma unroll
            for (int k = 0; k < 6 * DC_MMAX + 2; ++k) {
                float term = 0.f;
                if (k < 2 * M) term = gcv[k < 2 * DC_MMAX ? k : 0];
                else if (k == 2 * M) term = el;
                else if (k == 2 * M + 1) term = g2;
                else if (k < nsc) {
                    const int qq = (k - 2 * M - 2) / M, m = (k - 2 * M - 2) % M;
                    float a_ = 0.f, g_ = 0.f;
#pragma unroll
                    for (int q = 0; q < DC_MMAX; ++q) if (q == m) { a_ = av[q]; g_ = gv[q]; }
                    term = (qq == 0) ? a_ : (qq == 1) ? a_ * a_ : (qq == 2) ? g_ : g_ * a_;
                }
                t[k] += term;
            }
        }
#pragma unroll
        for (int k = 0; k < 6 * DC_MMAX + 2; ++k) {
            const float w = warp_sum(t[k]);
            if ((threadIdx.x & 31) == 0) sc_red[threadIdx.x >> 5][k] = w;
        }
        __syncthreads();
        if (i >= nu2 && i < D.tot && grp == 0) {
            float acc = 0.f;
            for (int w = 0; w < 8; ++w) acc += sc_red[w][i - nu2];
            s = acc;                                   // (the other three threads of the entry keep s = 0)
        }
    }
    part[grp][col] = s;
    __syncthreads();
    s = (part[0][col] + part[1][col]) + (part[2][col] + part[3][col]);
    const bool writer = (grp == 0) && (i < D.tot);
    if (seq == 0) {
        if (writer) D.red[i] = s;
        return;
    }
    const int W = D.cm.world, par = seq & 1;
    // the 64 sums of the CTA go to every peer as 16 stores of 16 bytes (they were 64 stores of 4 bytes): slots are 16-byte aligned,
    // a CTA starts at a multiple of 64 entries and the slot stride tot_pad is a multiple of 4
    __syncthreads();                                // every thread has read part[][]
    if (grp == 0) part[0][col] = (i < D.tot) ? s : 0.f;
    __syncthreads();
    for (int t = threadIdx.x; t < 16 * W; t += 256) {
        const int r = t >> 4, q = t & 15, i4 = blockIdx.x * 64 + 4 * q;
        if (i4 + 3 < D.tot_pad)
            *reinterpret_cast<float4*>(D.cm.slots[r] + ((size_t)par * W + D.cm.rank) * D.tot_pad + i4) = *reinterpret_cast<const float4*>(&part[0][4 * q]);
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned prev = atomicAdd(D.cm.ctr + par, 1u);
        if (prev == gridDim.x - 1) {
            D.cm.ctr[par] = 0;                      // next use of this parity is two iterations (kernel launches) later
            __threadfence_system();
            for (int r = 0; r < W; ++r) *(volatile int*)(D.cm.flags[r] + par * W + D.cm.rank) = seq;
        }
    }
}

// ---------------------------------------------------------------- shared-parameter update (one cluster)
__device__ __forceinline__ float atrous_g(const float* __restrict__ c, int nu, int v, int u, int Dd, int axis) {
    const float h0 = 1.f / 16.f, h1 = 4.f / 16.f, h2 = 6.f / 16.f;
    if (axis == 0) {
        const float* row = c + v * nu;
        return h0 * (row[max(u - 2 * Dd, 0)] + row[min(u + 2 * Dd, nu - 1)]) + h1 * (row[max(u - Dd, 0)] + row[min(u + Dd, nu - 1)]) + h2 * row[u];
    }
    return h0 * (c[max(v - 2 * Dd, 0) * nu + u] + c[min(v + 2 * Dd, nu - 1) * nu + u]) +
           h1 * (c[max(v - Dd, 0) * nu + u] + c[min(v + Dd, nu - 1) * nu + u]) + h2 * c[v * nu + u];
}

__device__ __forceinline__ float atrous_adj(const float* __restrict__ base, int stride, int nu, int i, int Dd) {
    const float h[5] = {1.f / 16.f, 4.f / 16.f, 6.f / 16.f, 4.f / 16.f, 1.f / 16.f};
    float acc = 0.f;
    if (i > 0 && i < nu - 1) {
#pragma unroll
        for (int t = 0; t < 5; ++t) {
            const int ip = i - (t - 2) * Dd;
            if (ip >= 0 && ip < nu) acc = fmaf(h[t], base[ip * stride], acc);
        }
    } else if (i == 0) {
#pragma unroll
        for (int t = 0; t < 5; ++t) {
            const int hi = min(-(t - 2) * Dd, nu - 1);
            float s = 0.f;
            for (int ip = 0; ip <= hi; ++ip) s += base[ip * stride];
            if (hi >= 0) acc = fmaf(h[t], s, acc);
        }
    } else {
#pragma unroll
        for (int t = 0; t < 5; ++t) {
            const int lo = max(nu - 1 - (t - 2) * Dd, 0);
            float s = 0.f;
            for (int ip = lo; ip < nu; ++ip) s += base[ip * stride];
            if (lo < nu) acc = fmaf(h[t], s, acc);
        }
    }
    return acc;
}

#define DU_THREADS 1024
#define DU_CTAS 8
#define DU_TIMEOUT (5LL << 31)      // ~5 s of SM clocks: a peer that never arrives must not hang the GPU
// The replicated half of an iteration runs on ONE thread-block cluster of 8 CTAs (8192 threads): the nu x nu
// planes live in global memory (L2), every stencil / point-wise pass is spread over the whole cluster and
// passes are separated by cluster.sync() (barrier.cluster arrive.release / wait.acquire, ~0.3 us) instead of
// kernel boundaries.  Cross-CTA sums go through a small global array in a fixed order (deterministic).

__device__ __forceinline__ float cluster_sum(float v, float* red, float* gpart, int slot, int tid, int rank) {
    v = warp_sum(v);
    __syncthreads();
    if ((tid & 31) == 0) red[tid >> 5] = v;
    __syncthreads();
    if (tid == 0) {
        float s = 0.f;
        for (int w = 0; w < DU_THREADS / 32; ++w) s += red[w];
        gpart[slot * DU_CTAS + rank] = s;
    }
    cg::this_cluster().sync();
    float s = 0.f;
    for (int c = 0; c < DU_CTAS; ++c) s += gpart[slot * DU_CTAS + c];
    return s;
}

// Starlet-L1 term of h: value -> ctl[6], gradient -> planes[0] (C0).  It depends on h only, so it runs on a second
// stream CONCURRENTLY with k_deconv_epoch / k_deconv_reduce of the same iteration (8 SMs for ~0.1 ms).
__global__ void __cluster_dims__(DU_CTAS, 1, 1) __launch_bounds__(DU_THREADS) k_deconv_starlet(DeconvDev D) {
    __shared__ float red[DU_THREADS / 32];
    cg::cluster_group cl = cg::this_cluster();
    const int rank = (int)cl.block_rank();
    const int tid = threadIdx.x, gtid = rank * DU_THREADS + tid;
    constexpr int GT_ALL = DU_CTAS * DU_THREADS;
    const int nu = D.nu, pp = nu * nu, J = D.J;
    float* C0 = D.planes;
    float* C1 = C0 + pp;
    float* Tj = C1 + 2 * pp;             // [J][pp]
    float reg = 0.f;
    for (int j = 0; j < J; ++j) {
        const int Dd = 1 << j;
        const float* cur = (j == 0) ? D.h : C0;
        for (int i = gtid; i < pp; i += GT_ALL) C1[i] = atrous_g(cur, nu, i / nu, i % nu, Dd, 0);
        cl.sync();
        const float lam = (j == 0) ? D.lam_hf : D.lam_scales;
        for (int i = gtid; i < pp; i += GT_ALL) {
            const float nxt = atrous_g(C1, nu, i / nu, i % nu, Dd, 1);
            const float al = cur[i] - nxt;
            const float lw = lam * (D.W ? D.W[(size_t)j * pp + i] : 1.f);
            reg = fmaf(lw, fabsf(al), reg);
            Tj[(size_t)j * pp + i] = (al > 0.f) ? lw : (al < 0.f) ? -lw : 0.f;
            C0[i] = nxt;             // in place: cur[i] is only read by the owner of pixel i in this pass
        }
        cl.sync();
    }
    // backward sweep g_j = t_j + H_j^T (g_{j+1} - t_j).  H^T folds the out-of-range taps of the edge-replicated filter
    // onto the two border pixels of every line: those need sums over up to 2 Dd + 1 elements, done by one WARP per
    // (line, side) instead of one thread, so that a pass stays one memory round trip long.
    const int gw = gtid >> 5, lane = tid & 31;
    constexpr int NW = GT_ALL / 32;
    const float h0 = 1.f / 16.f, h1 = 4.f / 16.f, h2 = 6.f / 16.f;
    for (int j = J - 1; j >= 0; --j) {
        const int Dd = 1 << j;
        for (int i = gtid; i < pp; i += GT_ALL) C0[i] = ((j == J - 1) ? 0.f : C0[i]) - Tj[(size_t)j * pp + i];
        cl.sync();
        for (int axis = 1; axis >= 0; --axis) {            // columns (C0 -> C1), then rows (C1 -> C0, + t_j)
            const float* src = axis ? C0 : C1;
            float* dst = axis ? C1 : C0;
            const float* Tp = axis ? nullptr : Tj + (size_t)j * pp;
            for (int i = gtid; i < pp; i += GT_ALL) {
                const int v = i / nu, u = i % nu, pos = axis ? v : u;
                if (pos > 0 && pos < nu - 1) {
                    const float val = axis ? atrous_adj(src + u, nu, nu, v, Dd) : atrous_adj(src + v * nu, 1, nu, u, Dd);
                    dst[i] = (Tp ? Tp[i] : 0.f) + val;
                }
            }
            for (int b = gw; b < 2 * nu; b += NW) {
                const int line = b >> 1, side = b & 1;
                const float* base = axis ? src + line : src + line * nu;
                const int stride = axis ? nu : 1;
                float P1 = 0.f, P2 = 0.f;                   // sums over distances 1..Dd and Dd+1..2Dd from the border
                for (int r = lane + 1; r <= 2 * Dd && r <= nu - 1; r += 32) {
                    const float x = base[(side ? nu - 1 - r : r) * stride];
                    if (r <= Dd) P1 += x; else P2 += x;
                }
                P1 = warp_sum(P1); P2 = warp_sum(P2);
                if (lane == 0) {
                    const float P0 = base[(side ? nu - 1 : 0) * stride];
                    const int o = axis ? (side ? nu - 1 : 0) * nu + line : line * nu + (side ? nu - 1 : 0);
                    dst[o] = (Tp ? Tp[o] : 0.f) + h0 * ((P0 + P1) + P2) + h1 * (P0 + P1) + h2 * P0;
                }
            }
            cl.sync();
        }
    }
    reg = cluster_sum(reg, red, D.gpart + 8 * DU_CTAS, 0, tid, rank);
    if (gtid == 0) D.ctl[6] = reg;
}

// Same term with every plane DISTRIBUTED over the shared memories of the 8 CTAs (CTA c owns rows [c R, (c+1) R) of
// h, the scratch planes and the J sign planes): row passes are local, column passes read the rows of other CTAs
// through DSMEM, and only ONE cluster barrier per scale and direction is needed (scratch planes are double buffered),
// with no global-memory fence inside the loop.  Used when (6 + J) R nu floats fit in shared memory.
__global__ void __cluster_dims__(DU_CTAS, 1, 1) __launch_bounds__(DU_THREADS) k_deconv_starlet_dsm(DeconvDev D, int R) {
    extern __shared__ __align__(16) float ssm[];
    __shared__ float red[DU_THREADS / 32];
    __shared__ float* peer[DU_CTAS];                  // ssm of every CTA
    cg::cluster_group cl = cg::this_cluster();
    const int rank = (int)cl.block_rank();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nu = D.nu, pp = nu * nu, J = D.J;
    const int bs = R * nu;                            // floats per plane band
    const int v0 = rank * R, nrow = max(0, min(R, nu - v0)), cnt = nrow * nu;
    // band-local planes: C0 (current c_j / gradient g), X[2] (row-filtered, double buffered), Q[2] (g - t, double buffered), T[J]
    float* C0 = ssm;
    float* Xb = ssm + bs;
    float* Qb = ssm + 3 * bs;
    float* Tb = ssm + 5 * bs;
    if (tid < DU_CTAS) peer[tid] = (float*)cl.map_shared_rank(ssm, tid);
    for (int i = tid; i < cnt; i += DU_THREADS) C0[i] = D.h[v0 * nu + i];
    cl.sync();
    const float h0 = 1.f / 16.f, h1 = 4.f / 16.f, h2 = 6.f / 16.f;
    auto rd = [&](int plane_off, int v, int u) -> float {      // element (v, u) of a distributed plane
        const int c = v / R;
        return peer[c][plane_off + (v - c * R) * nu + u];
    };
    float reg = 0.f;
    for (int j = 0; j < J; ++j) {
        const int Dd = 1 << j;
        float* X = Xb + (j & 1) * bs;
        const int xo = (int)(X - ssm);
        for (int i = tid; i < cnt; i += DU_THREADS) {           // rows: local
            const int u = i % nu;
            const float* row = C0 + i - u;
            X[i] = h0 * (row[max(u - 2 * Dd, 0)] + row[min(u + 2 * Dd, nu - 1)]) + h1 * (row[max(u - Dd, 0)] + row[min(u + Dd, nu - 1)]) + h2 * row[u];
        }
        cl.sync();
        const float lam = (j == 0) ? D.lam_hf : D.lam_scales;
        for (int i = tid; i < cnt; i += DU_THREADS) {           // columns: other bands through DSMEM
            const int v = v0 + i / nu, u = i % nu;
            const float nxt = h0 * (rd(xo, max(v - 2 * Dd, 0), u) + rd(xo, min(v + 2 * Dd, nu - 1), u)) +
                              h1 * (rd(xo, max(v - Dd, 0), u) + rd(xo, min(v + Dd, nu - 1), u)) + h2 * X[i];
            const float al = C0[i] - nxt;
            const float lw = lam * (D.W ? D.W[(size_t)j * pp + v0 * nu + i] : 1.f);
            reg = fmaf(lw, fabsf(al), reg);
            Tb[j * bs + i] = (al > 0.f) ? lw : (al < 0.f) ? -lw : 0.f;
            C0[i] = nxt;
        }
        __syncthreads();
    }
    // backward sweep g_j = t_j + H_j^T (g_{j+1} - t_j): Q = g - t (local) | barrier | columns (DSMEM) -> X | rows (local) -> C0
    for (int j = J - 1; j >= 0; --j) {
        const int Dd = 1 << j;
        float* Q = Qb + (j & 1) * bs;
        float* X = Xb;
        const int qo = (int)(Q - ssm);
        const float* T = Tb + j * bs;
        for (int i = tid; i < cnt; i += DU_THREADS) Q[i] = ((j == J - 1) ? 0.f : C0[i]) - T[i];
        cl.sync();
        // columns: X[v][u] = sum_t h[t] Q[v - (t-2) Dd][u] for interior rows; rows 0 and nu-1 collect the folded taps
        for (int i = tid; i < cnt; i += DU_THREADS) {
            const int v = v0 + i / nu, u = i % nu;
            if (v > 0 && v < nu - 1) {
                float acc = h2 * Q[i];
                if (v - Dd >= 0) acc = fmaf(h1, rd(qo, v - Dd, u), acc);
                if (v + Dd < nu) acc = fmaf(h1, rd(qo, v + Dd, u), acc);
                if (v - 2 * Dd >= 0) acc = fmaf(h0, rd(qo, v - 2 * Dd, u), acc);
                if (v + 2 * Dd < nu) acc = fmaf(h0, rd(qo, v + 2 * Dd, u), acc);
                X[i] = acc;
            }
        }
        if (rank == 0 || v0 + nrow == nu) {           // bands holding row 0 / row nu-1: sums over <= 2 Dd rows
            for (int side = 0; side < 2; ++side) {
                if ((side == 0 && rank != 0) || (side == 1 && (v0 + nrow != nu || nrow == 0))) continue;
                for (int u = warp; u < nu; u += DU_THREADS / 32) {     // one warp per column
                    float P1 = 0.f, P2 = 0.f;
                    for (int r = lane + 1; r <= 2 * Dd && r <= nu - 1; r += 32) {
                        const float x = rd(qo, side ? nu - 1 - r : r, u);
                        if (r <= Dd) P1 += x; else P2 += x;
                    }
                    P1 = warp_sum(P1); P2 = warp_sum(P2);
                    if (lane == 0) {
                        const int vb = side ? nu - 1 : 0;
                        const float P0 = rd(qo, vb, u);
                        X[(vb - v0) * nu + u] = h0 * ((P0 + P1) + P2) + h1 * (P0 + P1) + h2 * P0;
                    }
                }
            }
        }
        __syncthreads();
        // rows (local): C0 = t_j + H^T X along u; columns 0 and nu-1 of every row collect the folded taps (one warp per row and side)
        for (int i = tid; i < cnt; i += DU_THREADS) {
            const int u = i % nu;
            if (u > 0 && u < nu - 1) {
                const float* row = X + i - u;
                float acc = h2 * row[u];
                if (u - Dd >= 0) acc = fmaf(h1, row[u - Dd], acc);
                if (u + Dd < nu) acc = fmaf(h1, row[u + Dd], acc);
                if (u - 2 * Dd >= 0) acc = fmaf(h0, row[u - 2 * Dd], acc);
                if (u + 2 * Dd < nu) acc = fmaf(h0, row[u + 2 * Dd], acc);
                C0[i] = T[i] + acc;
            }
        }
        for (int b = warp; b < 2 * nrow; b += DU_THREADS / 32) {
            const int rl = b >> 1, side = b & 1;
            const float* row = X + rl * nu;
            float P1 = 0.f, P2 = 0.f;
            for (int r = lane + 1; r <= 2 * Dd && r <= nu - 1; r += 32) {
                const float x = row[side ? nu - 1 - r : r];
                if (r <= Dd) P1 += x; else P2 += x;
            }
            P1 = warp_sum(P1); P2 = warp_sum(P2);
            if (lane == 0) {
                const int ub = side ? nu - 1 : 0;
                const float P0 = row[ub];
                C0[rl * nu + ub] = T[rl * nu + ub] + h0 * ((P0 + P1) + P2) + h1 * (P0 + P1) + h2 * P0;
            }
        }
        __syncthreads();
    }
    for (int i = tid; i < cnt; i += DU_THREADS) D.planes[v0 * nu + i] = C0[i];
    reg = cluster_sum(reg, red, D.gpart + 8 * DU_CTAS, 0, tid, rank);
    if (rank == 0 && tid == 0) D.ctl[6] = reg;
    cl.sync();                                        // nobody leaves while its band may still be read
}

// The same term on ONE SM: the whole nu x nu background (nu = 64 or 128) with its two scratch planes and the packed sign planes
// fits the shared memory of a single CTA (174 KB at nu = 128), so the float4 routine of the PSF fit (lcb_starlet.cuh) applies:
// no DSMEM reads, no cluster barriers, every stencil tap one LDS.128.  Measured at cfg4: 0.13 ms for the 8-CTA DSMEM kernel --
// the critical path of an iteration once the epochs are sharded over 8 GPUs (the epoch kernel takes 0.10 ms there) -- against
// ~0.03 ms here.
#ifndef LCB_SM_STARLET_THREADS
#define LCB_SM_STARLET_THREADS 512
#endif
template <int NUT>
__global__ void __launch_bounds__(NUT == 128 ? LCB_SM_STARLET_THREADS : 256) k_deconv_starlet_sm(DeconvDev D) {
    constexpr int NTH = (NUT == 128) ? LCB_SM_STARLET_THREADS : 256;
    constexpr int PP = NUT * NUT;
    extern __shared__ __align__(16) float ssm[];
    __shared__ float red[NTH / 32];
    const int tid = threadIdx.x;
    float* C0 = ssm;
    float* C1 = C0 + PP;
    float* aux = C1 + PP;                                            // [NUT][NUT/4 + 1] + [NUT][2]
    signed char* sg = reinterpret_cast<signed char*>(aux + NUT * (NUT / 4 + 1) + 2 * NUT);   // [J][PP / 4]
    float reg = starlet_reg_fast4<NUT, NTH>(D.h, C0, C1, sg, aux, D.W, D.lam_hf, D.lam_scales, D.J, tid);
    for (int i = tid; i < PP / 4; i += NTH) reinterpret_cast<float4*>(D.planes)[i] = reinterpret_cast<const float4*>(C0)[i];
    reg = warp_sum(reg);
    if ((tid & 31) == 0) red[tid >> 5] = reg;
    __syncthreads();
    if (tid == 0) {
        float s_ = 0.f;
        for (int w = 0; w < NTH / 32; ++w) s_ += red[w];
        D.ctl[6] = s_;
    }
}

static size_t starlet_sm_bytes(int nu, int J) {
    return ((size_t)2 * nu * nu + (size_t)nu * (nu / 4 + 1) + 2 * nu) * 4 + (size_t)J * nu * nu / 4;
}

// several cluster-wide sums with ONE cluster barrier
template <int N>
__device__ __forceinline__ void cluster_sums(float (&v)[N], float* red, float* gpart, int tid, int rank) {
#pragma unroll
    for (int q = 0; q < N; ++q) v[q] = warp_sum(v[q]);
    __syncthreads();
    if ((tid & 31) == 0) {
#pragma unroll
        for (int q = 0; q < N; ++q) red[q * (DU_THREADS / 32) + (tid >> 5)] = v[q];
    }
    __syncthreads();
    if (tid < N) {
        float s = 0.f;
        for (int w = 0; w < DU_THREADS / 32; ++w) s += red[tid * (DU_THREADS / 32) + w];
        gpart[tid * DU_CTAS + rank] = s;
    }
    cg::this_cluster().sync();
#pragma unroll
    for (int q = 0; q < N; ++q) {
        float s = 0.f;
        for (int c = 0; c < DU_CTAS; ++c) s += gpart[q * DU_CTAS + c];
        v[q] = s;
    }
}

// Replicated half of an iteration.  it < 0: evaluation only (loss + full gradient into the output arrays, no
// update).  seq > 0: red[] is first assembled from the receive slots of all ranks (summed in rank order: bit-identical
// everywhere).  with_reg: planes[0] / ctl[6] hold the starlet gradient / value from k_deconv_starlet.
__global__ void __cluster_dims__(DU_CTAS, 1, 1) __launch_bounds__(DU_THREADS)
k_deconv_update(DeconvDev D, int it_arg, int n_iter, float lr0, int schedule, int seq_arg, int with_reg, float* grad_h_out, float* grad_c_out, float* loss_out) {
    // it_arg == -2: iteration index and sequence number come from D.ictl (identical launch arguments for every iteration, so that
    // one captured CUDA graph of an iteration can be replayed); the kernel advances them at its end
    const bool dev_it = (it_arg == -2);
    const int it = dev_it ? D.ictl[0] : it_arg;
    const int seq = dev_it ? ((D.cm.world > 1) ? D.ictl[1] : 0) : seq_arg;
    __shared__ float red[4 * (DU_THREADS / 32)];
    __shared__ float fus[4 * DC_MMAX];
    cg::cluster_group cl = cg::this_cluster();
    const int rank = (int)cl.block_rank();
    const int tid = threadIdx.x, gtid = rank * DU_THREADS + tid;
    constexpr int GT_ALL = DU_CTAS * DU_THREADS;
    const int nu = D.nu, pp = nu * nu, M = D.M;
    const float* sl = nullptr;
    if (seq > 0) {
        const int W = D.cm.world, par = seq & 1;
        if (tid < W) {
            volatile int* f = D.cm.flags[D.cm.rank] + par * W + tid;
            if (*(volatile float*)(D.ctl + 7) == 0.f) {
                const long long t0 = clock64();
                while (*f < seq) {
                    if (clock64() - t0 > DU_TIMEOUT) { D.ctl[7] = 1.f; break; }
                }
            }
        }
        __threadfence_system();
        __syncthreads();
        sl = D.cm.slots[D.cm.rank] + (size_t)par * W * D.tot_pad;
    }
    // sum over the ranks of entry i of the exchanged buffer, in rank order (bit-identical everywhere); all loads in flight at once
    const int Wn = D.cm.world;
    auto rank_sum = [&](int i) -> float {
        float v[DC_MAXW];
#pragma unroll
        for (int r = 0; r < DC_MAXW; ++r) v[r] = (r < Wn) ? __ldcg(sl + (size_t)r * D.tot_pad + i) : 0.f;
        float t = 0.f;
#pragma unroll
        for (int r = 0; r < DC_MAXW; ++r) t += v[r];
        return t;
    };
    // the scalar tail of red[] (gradients of c, loss, |g_epoch|^2, flux sums) is needed by every CTA: each one assembles its own copy
    // in shared memory, so that no cluster barrier separates the exchange from the gradient pass
    __shared__ float sred[6 * DC_MMAX + 2];
    {
        const int nsc = D.tot - pp;
        if (tid < nsc) {
            const float t = (seq > 0) ? rank_sum(pp + tid) : D.red[pp + tid];
            sred[tid] = t;
            if (seq > 0 && rank == 0) D.red[pp + tid] = t;
        }
        __syncthreads();
    }
    const float* C0 = D.planes;
    float* GT = D.planes + 2 * pp;       // total gradient wrt h
    float v4[4] = {0.f, 0.f, 0.f, 0.f};  // |g|^2, positivity, prior, flux uniformity
    if (D.free_h) {
        for (int i = gtid; i < pp; i += GT_ALL) {
            const float hv = D.h[i];
            const float rs = (seq > 0) ? rank_sum(i) : D.red[i];
            if (seq > 0) D.red[i] = rs;
            float g = rs + (with_reg ? C0[i] : 0.f);
            if (D.lam_pos != 0.f && hv < 0.f) { g -= D.lam_pos; v4[1] -= D.lam_pos * hv; }
            GT[i] = g;
            v4[0] = fmaf(g, g, v4[0]);
            if (grad_h_out) grad_h_out[i] = g;
        }
    } else if (seq > 0) {
        for (int i = gtid; i < pp; i += GT_ALL) D.red[i] = rank_sum(i);
    }
    // c_x, c_y gradients (+ prior): every CTA computes them redundantly (identical), only rank 0 counts them
    __shared__ float gcs[2 * DC_MMAX];
    if (tid < 2 * M) {
        float g = sred[tid];
        if (D.has_prior) {
            const int ax = tid / M, m = tid % M;
            const float mu = D.prior[(2 * ax) * M + m], sg = D.prior[(2 * ax + 1) * M + m];
            const float z = (D.c[tid] - mu) / sg;
            g += z / sg;
            if (rank == 0) v4[2] = 0.5f * z * z;
        }
        gcs[tid] = g;
        if (grad_c_out && rank == 0) grad_c_out[tid] = g;
        if (D.free_c && rank == 0) v4[0] = fmaf(g, g, v4[0]);
    }
    // flux uniformity: lam * sum_m std_e(a_em) [/ |mean_e(a_em)|] from the all-reduced shifted sums; the gradient
    // wrt a_em is A_m (a_em - mean_m) - B_m, applied by the epoch kernel with the pending update
    if (D.lam_fu != 0.f && tid < M) {
        const int o = red_flux(D);
        const float Et = (float)D.E_total, K = D.fu[tid];
        const float S1 = sred[o - pp + tid], S2 = sred[o - pp + M + tid], Sg = sred[o - pp + 2 * M + tid], Sga = sred[o - pp + 3 * M + tid];
        const float dm = S1 / Et, mean = K + dm;
        const float var = fmaxf(S2 / Et - dm * dm, 0.f), sd = sqrtf(var);
        float A = 0.f, B = 0.f, val = 0.f;
        if (sd > 0.f) {
            if (D.fu_relative) {
                const float am = fabsf(mean);
                val = D.lam_fu * sd / am;
                A = D.lam_fu / (Et * sd * am);
                B = D.lam_fu * sd * ((mean > 0.f) ? 1.f : -1.f) / (Et * mean * mean);
            } else {
                val = D.lam_fu * sd;
                A = D.lam_fu / (Et * sd);
            }
        }
        fus[tid] = mean; fus[DC_MMAX + tid] = A; fus[2 * DC_MMAX + tid] = B;
        if (rank == 0) {
            v4[3] = val;
            // |g_a|^2 with the flux-uniformity term: sum_e (g + A (a - mean) - B)^2
            if (D.free_a) v4[0] += A * A * var * Et + Et * B * B + 2.f * A * (Sga - dm * Sg) - 2.f * B * Sg;
        }
    }
    cluster_sums<4>(v4, red, D.gpart, tid, rank);
    const float gn2 = v4[0] + sred[2 * M + 1];
    const float L = sred[2 * M] + (with_reg ? D.ctl[6] : 0.f) + v4[1] + v4[2] + v4[3];
    if (rank == 0 && D.lam_fu != 0.f && tid < M) {      // every read of D.fu above is behind the cluster barrier
        D.fu[tid] = fus[tid]; D.fu[DC_MMAX + tid] = fus[DC_MMAX + tid]; D.fu[2 * DC_MMAX + tid] = fus[2 * DC_MMAX + tid];
    }
    if (gtid == 0) {
        if (loss_out) loss_out[0] = L;
        if (it >= 0 && D.loss_hist) D.loss_hist[it] = L;
        D.ctl[5] = L;
    }
    if (it < 0) return;
    float cs = 1.f, lr = lr0;
    if (schedule) {
        const float gn = sqrtf(gn2);
        cs = (gn < D.cv.clip) ? 1.f : D.cv.clip / gn;
        lr = lr0 * powf(D.cv.decay, (float)it / (float)n_iter);
    }
    const float b1t = powf(D.cv.b1, (float)(it + 1)), b2t = powf(D.cv.b2, (float)(it + 1));
    const BeliefCoef bc = {lr, D.cv.b1, D.cv.b2, 1.f - D.cv.b1, 1.f - D.cv.b2, 1.f / (1.f - b1t), 1.f / (1.f - b2t),
                           D.cv.eps, D.cv.eps_root};
    if (D.free_h) {
        for (int i = gtid; i < pp; i += GT_ALL) {
            float hv = D.h[i], mu = D.h_mu[i], nv = D.h_nu[i];
            belief_update(bc, cs * GT[i], hv, mu, nv);
            D.h[i] = hv; D.h_mu[i] = mu; D.h_nu[i] = nv;
        }
    }
    if (D.free_c && rank == 0 && tid < 2 * M) {
        float cvv = D.c[tid], mu = D.c_mu[tid], nv = D.c_nu[tid];
        belief_update(bc, cs * gcs[tid], cvv, mu, nv);
        D.c[tid] = cvv; D.c_mu[tid] = mu; D.c_nu[tid] = nv;
    }
    if (gtid == 0) { D.ctl[0] = cs; D.ctl[1] = lr; D.ctl[2] = bc.inv_bc1; D.ctl[3] = bc.inv_bc2; D.ctl[4] = 1.f; }
    // (every CTA read ictl before the cluster barrier of cluster_sums: no further barrier is needed before it advances)
    if (dev_it && gtid == 0) { D.ictl[0] = it + 1; D.ictl[1] = D.ictl[1] + 1; }
}

// adds the flux-uniformity gradient to the stored per-epoch gradients (evaluation path only: lcb_deconv_loss_grad)
__global__ void k_deconv_fu_apply(DeconvDev D) {
    const int np = D.M + 3;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= D.E * D.M) return;
    const int e = i / D.M, m = i % D.M;
    D.ep_g[(size_t)e * np + m] += D.fu[DC_MMAX + m] * (D.ep[(size_t)e * np + m] - D.fu[m]) - D.fu[2 * DC_MMAX + m];
}

// ---------------------------------------------------------------- device-side L-BFGS of stage 1 (roi_modelling.py:260-281)
// The reference minimises {dx, dy, a} with scipy's L-BFGS-B (host optimiser, one host round trip per evaluation).  Here the
// optimiser state lives in device memory and ONE CTA drives it between evaluations: projected L-BFGS (m = LB_M pairs, two-loop
// recursion on the gradient with the coordinates that sit on an active bound removed, trial points projected onto the box,
// Armijo backtracking, curvature pairs kept only when s.y > 0) with scipy's stopping rules (relative loss decrease <= ftol,
// projected gradient <= pgtol, maxiter).  The host only enqueues "evaluate ; step" rounds and polls one flag per chunk.
// Vector layout: i = e * (M + 2) + p, p < M: a_em (lower bound a_lo), p = M, M+1: dx_e, dy_e (|.| <= n/2).
#define LB_M 10
#define LB_THREADS 1024
#define LB_FLAT 10
struct LbState {
    float *x, *g, *d, *S, *Y;      // [nv], [nv], [nv], [LB_M][nv], [LB_M][nv]
    float *rho;                    // [LB_M]
    float *sc;                     // [16] 0 f, 1 step t, 2 slope g.d, 3 iterations, 4 pairs stored, 5 converged (1 ftol, 2 pgtol, 3 maxiter,
                                   //      4 line search stalled), 6 evaluations, 7 head of the circular history, 8 |proj grad|_inf, 9 rejected steps in a row,
                                   //      10 accepted steps in a row whose relative decrease was <= ftol
    float *hist;                   // [maxiter] loss after every accepted iteration
    int nv, maxiter;
    float a_lo, d_max, ftol, pgtol;
};

__device__ __forceinline__ float lb_block_sum(float v, float* red, int tid) {
    v = warp_sum(v);
    __syncthreads();
    if ((tid & 31) == 0) red[tid >> 5] = v;
    __syncthreads();
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < LB_THREADS / 32; ++w) s += red[w];
    return s;
}
__device__ __forceinline__ float lb_block_max(float v, float* red, int tid) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();
    if ((tid & 31) == 0) red[tid >> 5] = v;
    __syncthreads();
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < LB_THREADS / 32; ++w) s = fmaxf(s, red[w]);
    return s;
}

__global__ void __launch_bounds__(LB_THREADS) k_deconv_lbfgs_step(DeconvDev D, LbState L, const float* __restrict__ loss_in) {
    __shared__ float red[LB_THREADS / 32];
    __shared__ float alpha[LB_M];
    const int tid = threadIdx.x, nv = L.nv, M = D.M, np = M + 3, npv = M + 2;
    float* sc = L.sc;
    if (sc[5] != 0.f) return;                                   // converged: the remaining rounds of the chunk are no-ops
    auto lo = [&](int i) { return (i % npv < M) ? L.a_lo : -L.d_max; };
    auto hi = [&](int i) { return (i % npv < M) ? INFINITY : L.d_max; };
    auto ep_idx = [&](int i) { return (i / npv) * np + (i % npv); };
    const float f_t = *loss_in;
    const int evals = (int)sc[6];
    bool accept = true;
    const float f_old = sc[0];
    if (evals > 0) {
        // Armijo on the projected step: f_t <= f + c1 g.(x_t - x)
        float part = 0.f;
        for (int i = tid; i < nv; i += LB_THREADS) part = fmaf(L.g[i], D.ep[ep_idx(i)] - L.x[i], part);
        const float gs = lb_block_sum(part, red, tid);
        accept = isfinite(f_t) && (f_t <= f_old + 1e-4f * gs);
    }
    if (!accept) {
        const float t = sc[1] * 0.5f;
        const float rej = sc[9] + 1.f;
        __syncthreads();
        if (tid == 0) { sc[1] = t; sc[9] = rej; sc[6] = (float)(evals + 1); if (rej > 30.f) sc[5] = 4.f; }
        for (int i = tid; i < nv; i += LB_THREADS) D.ep[ep_idx(i)] = fminf(fmaxf(fmaf(t, L.d[i], L.x[i]), lo(i)), hi(i));
        return;
    }
    // ---- accepted (or first evaluation): curvature pair, new point
    int pairs = (int)sc[4], head = (int)sc[7];
    const int iters = (int)sc[3] + (evals > 0 ? 1 : 0);
    if (evals > 0) {
        float sy = 0.f, yy = 0.f;
        float* Sn = L.S + (size_t)head * nv;
        float* Yn = L.Y + (size_t)head * nv;
        for (int i = tid; i < nv; i += LB_THREADS) {
            const float s_ = D.ep[ep_idx(i)] - L.x[i], y_ = D.ep_g[ep_idx(i)] - L.g[i];
            Sn[i] = s_; Yn[i] = y_;
            sy = fmaf(s_, y_, sy); yy = fmaf(y_, y_, yy);
        }
        sy = lb_block_sum(sy, red, tid);
        yy = lb_block_sum(yy, red, tid);
        if (sy > 1e-10f * yy && sy > 0.f) {
            if (tid == 0) L.rho[head] = 1.f / sy;
            head = (head + 1) % LB_M;
            pairs = min(pairs + 1, LB_M);
        }
    }
    float pgmax = 0.f;
    for (int i = tid; i < nv; i += LB_THREADS) {
        const float xi = D.ep[ep_idx(i)], gi = D.ep_g[ep_idx(i)];
        L.x[i] = xi; L.g[i] = gi;
        // projected gradient: x - P(x - g)
        const float pg = xi - fminf(fmaxf(xi - gi, lo(i)), hi(i));
        pgmax = fmaxf(pgmax, fabsf(pg));
    }
    pgmax = lb_block_max(pgmax, red, tid);
    int conv = 0;
    float flat = 0.f;
    if (evals > 0) {
        if (tid == 0 && iters - 1 < L.maxiter) L.hist[iters - 1] = f_t;
        // scipy stops at the first relative decrease <= ftol (2.2e-9 by default).  The loss here is a float32 number: decreases
        // below ~1e-7 of it read as zero although the iteration is still making progress (measured: scipy in float64 gains
        // another 0.3 % over 230 such iterations), so the rule only fires after LB_FLAT accepted steps in a row without a
        // resolvable decrease
        if (f_old - f_t <= L.ftol * fmaxf(fmaxf(fabsf(f_old), fabsf(f_t)), 1.f)) flat = sc[10] + 1.f;
        if (flat >= (float)LB_FLAT) conv = 1;
    }
    if (pgmax <= L.pgtol) conv = 2;
    if (!conv && iters >= L.maxiter) conv = 3;
    __syncthreads();
    // ---- direction: two-loop recursion on the free part of the gradient
    auto is_free = [&](int i, float xi, float gi) { return !((xi <= lo(i) && gi > 0.f) || (xi >= hi(i) && gi < 0.f)); };
    for (int i = tid; i < nv; i += LB_THREADS) { const float xi = L.x[i], gi = L.g[i]; L.d[i] = is_free(i, xi, gi) ? gi : 0.f; }
    __syncthreads();
    for (int q = 0; q < pairs; ++q) {
        const int j = (head - 1 - q + 2 * LB_M) % LB_M;
        float part = 0.f;
        for (int i = tid; i < nv; i += LB_THREADS) part = fmaf(L.S[(size_t)j * nv + i], L.d[i], part);
        const float a_ = L.rho[j] * lb_block_sum(part, red, tid);
        if (tid == 0) alpha[j] = a_;
        for (int i = tid; i < nv; i += LB_THREADS) L.d[i] = fmaf(-a_, L.Y[(size_t)j * nv + i], L.d[i]);
        __syncthreads();
    }
    if (pairs > 0) {
        const int j = (head - 1 + LB_M) % LB_M;
        float yy = 0.f;
        for (int i = tid; i < nv; i += LB_THREADS) { const float y_ = L.Y[(size_t)j * nv + i]; yy = fmaf(y_, y_, yy); }
        yy = lb_block_sum(yy, red, tid);
        const float gam = 1.f / (L.rho[j] * yy);
        for (int i = tid; i < nv; i += LB_THREADS) L.d[i] *= gam;
        __syncthreads();
    }
    for (int q = pairs - 1; q >= 0; --q) {
        const int j = (head - 1 - q + 2 * LB_M) % LB_M;
        float part = 0.f;
        for (int i = tid; i < nv; i += LB_THREADS) part = fmaf(L.Y[(size_t)j * nv + i], L.d[i], part);
        const float b_ = L.rho[j] * lb_block_sum(part, red, tid);
        const float a_ = alpha[j];
        for (int i = tid; i < nv; i += LB_THREADS) L.d[i] = fmaf(a_ - b_, L.S[(size_t)j * nv + i], L.d[i]);
        __syncthreads();
    }
    // d = -H g on the free coordinates; fall back to steepest descent if it is not a descent direction
    float slope = 0.f, gn2 = 0.f;
    for (int i = tid; i < nv; i += LB_THREADS) {
        const float xi = L.x[i], gi = L.g[i];
        const float di = is_free(i, xi, gi) ? -L.d[i] : 0.f;
        L.d[i] = di;
        slope = fmaf(gi, di, slope);
        gn2 += is_free(i, xi, gi) ? gi * gi : 0.f;
    }
    slope = lb_block_sum(slope, red, tid);
    gn2 = lb_block_sum(gn2, red, tid);
    float t = 1.f;
    if (!(slope < 0.f) || pairs == 0) {
        for (int i = tid; i < nv; i += LB_THREADS) { const float xi = L.x[i], gi = L.g[i]; L.d[i] = is_free(i, xi, gi) ? -gi : 0.f; }
        slope = -gn2;
        t = fminf(1.f, rsqrtf(fmaxf(gn2, 1e-30f)));             // scipy / Nocedal: first step of length min(1, 1/|g|)
    }
    __syncthreads();
    if (tid == 0) {
        sc[0] = f_t; sc[1] = t; sc[2] = slope; sc[3] = (float)iters; sc[4] = (float)pairs; sc[5] = (float)conv;
        sc[6] = (float)(evals + 1); sc[7] = (float)head; sc[8] = pgmax; sc[9] = 0.f; sc[10] = flat;
    }
    // next trial point (or, when converged, the accepted point itself: it is already in D.ep)
    if (!conv)
        for (int i = tid; i < nv; i += LB_THREADS) D.ep[ep_idx(i)] = fminf(fmaxf(fmaf(t, L.d[i], L.x[i]), lo(i)), hi(i));
}

// restores the last ACCEPTED point (a rejected trial may be sitting in D.ep when the rounds run out)
__global__ void k_deconv_lbfgs_finish(DeconvDev D, LbState L) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= L.nv || L.sc[6] == 0.f) return;
    const int npv = D.M + 2;
    D.ep[(i / npv) * (D.M + 3) + (i % npv)] = L.x[i];
}

// ================================================================= host side: handle-based ABI
struct DeconvHandle {
    DeconvDev D;
    std::vector<void*> owned;
    std::vector<void*> ipc_open;         // peer mappings to close
    void* comm_buf;                      // own receive buffer (IPC exported)
    cudaStream_t st;
    cudaStream_t st_own;                 // created when the caller passes the NULL stream (the legacy stream cannot be captured into a graph)
    cudaStream_t st2;                    // the starlet term of h runs here, concurrently with the epoch kernel
    cudaEvent_t ev_h, ev_go, ev_reg;     // h final (main stream) -> starlet may start ; starlet dispatched -> epoch may start ; starlet done -> update
    bool reg_pending, starlet_attr;
    int loss_cap;
    size_t smem_epoch;
    int CS;                              // CTAs per epoch (cluster size); 0 = not yet chosen
    int CS_user;
    int n_sm, max_smem, smem_sm;
    int seq;                             // sequence number of the in-kernel all-reduce
    float *gh, *gcx, *ls;                // evaluation outputs (lcb_deconv_loss_grad / _get)
    float* W_spare;                      // weight cube detached by set_reg(W = NULL), reused by the next one
    float* noise_tab;                    // 1-D kernels of the starlet-space noise propagation
    float* lbfgs; size_t lbfgs_floats;   // state of the device-side L-BFGS of stage 1 (lcb_deconv_lbfgs)
};

static int dalloc(DeconvHandle* H, void** p, size_t bytes, bool zero) {
    cudaError_t e = cudaMalloc(p, bytes ? bytes : 4);
    if (e != cudaSuccess) { lcb_set_error("deconv cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e)); return LCB_ERR_NOMEM; }
    H->owned.push_back(*p);
    if (zero) LCB_CUDA(cudaMemsetAsync(*p, 0, bytes ? bytes : 4, H->st));
    return LCB_OK;
}

static int put(DeconvHandle* H, float* dst, const float* src, size_t count, int mem) {
    if (!src) return LCB_OK;
    LCB_CUDA(cudaMemcpyAsync(dst, src, count * 4, mem == LCB_MEM_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, H->st));
    if (mem == LCB_MEM_HOST) LCB_CUDA(cudaStreamSynchronize(H->st));
    return LCB_OK;
}

static int get(DeconvHandle* H, float* dst, const float* src, size_t count, int mem) {
    if (!dst) return LCB_OK;
    LCB_CUDA(cudaMemcpyAsync(dst, src, count * 4, mem == LCB_MEM_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, H->st));
    if (mem == LCB_MEM_HOST) LCB_CUDA(cudaStreamSynchronize(H->st));
    return LCB_OK;
}

// CTAs per epoch: the smallest power of two (<= 8, bands of >= 4 rows) for which two CTAs fit the shared memory of
// an SM and the local epochs give every SM at least one CTA.  Measured on cfg4 shapes (n = 64, P = 64, 148 SMs):
// 200 / 100 / 50 local epochs -> 4, 25 -> 8; small stamps whose planes fit twice per SM anyway -> 1.
static int choose_cluster(const DeconvHandle* H) {
    const DeconvDev& D = H->D;
    const size_t two_per_sm = (size_t)(H->smem_sm / 2 - 1024);
    auto bytes = [&](int c) { return (size_t)dc_layout(D.n, D.k, D.NA, D.A0, c).total * 4; };
    int cs = 1;
    if (H->CS_user > 0) cs = H->CS_user;
    else while (cs < DC_CSMAX && (D.E * cs < H->n_sm || bytes(cs) > two_per_sm)) cs *= 2;
    while (cs > 1 && (D.n + cs - 1) / cs < 4) cs /= 2;
    while (cs < DC_CSMAX && bytes(cs) > (size_t)H->max_smem) cs *= 2;
    return cs;
}

static int launch_epoch(DeconvHandle* H, int flags, int seq = 0) {
    DeconvDev& D = H->D;
    if (H->CS == 0) {
        H->CS = choose_cluster(H);
        H->smem_epoch = (size_t)dc_layout(D.n, D.k, D.NA, D.A0, H->CS).total * 4;
        if (H->smem_epoch > (size_t)H->max_smem) {
            lcb_set_error("deconvolution: n=%d k=%d P=%d needs %zu B of shared memory per CTA (> %d)", D.n, D.k, D.P, H->smem_epoch, H->max_smem);
            H->CS = 0;
            return LCB_ERR_ARG;
        }
        const void* fn = D.k == 1 ? (const void*)k_deconv_epoch<1> : D.k == 2 ? (const void*)k_deconv_epoch<2>
                         : D.k == 3 ? (const void*)k_deconv_epoch<3> : (const void*)k_deconv_epoch<4>;
        LCB_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)H->smem_epoch));
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(D.E * H->CS), 1, 1);
    cfg.blockDim = dim3(DC_THREADS, 1, 1);
    cfg.dynamicSmemBytes = H->smem_epoch;
    cfg.stream = H->st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)H->CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    LcbProfScope ps("k_deconv_epoch", H->st);
    {   // development switches (A/B timing): 8 outputs per thread in the forward / adjoint pass
        static const int dev_flags = (getenv("LCB_DC_FWD8") ? 8 : 0) | (getenv("LCB_DC_ADJ8") ? 16 : 0) | (getenv("LCB_DC_FWDROWS") ? 32 : 0);
        flags |= dev_flags;
    }
    switch (D.k) {
        case 1: LCB_CUDA(cudaLaunchKernelEx(&cfg, k_deconv_epoch<1>, D, flags, H->CS, seq)); break;
        case 2: LCB_CUDA(cudaLaunchKernelEx(&cfg, k_deconv_epoch<2>, D, flags, H->CS, seq)); break;
        case 3: LCB_CUDA(cudaLaunchKernelEx(&cfg, k_deconv_epoch<3>, D, flags, H->CS, seq)); break;
        default: LCB_CUDA(cudaLaunchKernelEx(&cfg, k_deconv_epoch<4>, D, flags, H->CS, seq)); break;
    }
    return LCB_OK;
}

static int launch_reduce(DeconvHandle* H, int force_h, int seq) {
    DeconvDev& D = H->D;
    LcbProfScope ps("k_deconv_reduce", H->st);
    k_deconv_reduce<<<(D.tot + 63) / 64, 256, 0, H->st>>>(D, force_h, seq);
    LCB_CUDA(cudaGetLastError());
    return LCB_OK;
}

static bool with_reg(const DeconvHandle* H) {
    const DeconvDev& D = H->D;
    return D.free_h && (D.lam_scales != 0.f || D.lam_hf != 0.f);
}

// starlet term of the CURRENT h on the second stream; call before launch_epoch of the same evaluation
static int launch_starlet(DeconvHandle* H) {
    H->reg_pending = false;
    if (!with_reg(H)) return LCB_OK;
    LCB_CUDA(cudaEventRecord(H->ev_h, H->st));
    LCB_CUDA(cudaStreamWaitEvent(H->st2, H->ev_h, 0));
    LCB_CUDA(cudaEventRecord(H->ev_go, H->st2));          // the main stream resumes only once the second stream has seen ev_h:
    LCB_CUDA(cudaStreamWaitEvent(H->st, H->ev_go, 0));    // both kernels become ready together, the priority decides
    {
        const DeconvDev& D = H->D;
        const int R = (D.nu + DU_CTAS - 1) / DU_CTAS;
        const size_t dsm = (size_t)(5 + D.J) * R * D.nu * 4;
        LcbProfScope ps("k_deconv_starlet", H->st2);
        const size_t one_sm = starlet_sm_bytes(D.nu, D.J);
        if ((D.nu == 128 || D.nu == 64) && one_sm <= (size_t)H->max_smem && !getenv("LCB_DECONV_STARLET_CLUSTER")) {
            if (!H->starlet_attr) {
                if (D.nu == 128) LCB_CUDA(cudaFuncSetAttribute(k_deconv_starlet_sm<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)one_sm));
                else LCB_CUDA(cudaFuncSetAttribute(k_deconv_starlet_sm<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)one_sm));
                H->starlet_attr = true;
            }
            if (D.nu == 128) k_deconv_starlet_sm<128><<<1, LCB_SM_STARLET_THREADS, one_sm, H->st2>>>(D);
            else k_deconv_starlet_sm<64><<<1, 256, one_sm, H->st2>>>(D);
        } else if (dsm <= (size_t)H->max_smem && !getenv("LCB_DECONV_STARLET_L2")) {
            if (!H->starlet_attr) {
                LCB_CUDA(cudaFuncSetAttribute(k_deconv_starlet_dsm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dsm));
                H->starlet_attr = true;
            }
            k_deconv_starlet_dsm<<<DU_CTAS, DU_THREADS, dsm, H->st2>>>(D, R);
        } else {
            k_deconv_starlet<<<DU_CTAS, DU_THREADS, 0, H->st2>>>(D);
        }
    }
    LCB_CUDA(cudaGetLastError());
    LCB_CUDA(cudaEventRecord(H->ev_reg, H->st2));
    H->reg_pending = true;
    return LCB_OK;
}

static int launch_update(DeconvHandle* H, int it, int n_iter, float lr, int schedule, int seq, float* gh, float* gc, float* ls) {
    const int wr = H->reg_pending ? 1 : 0;
    if (wr) LCB_CUDA(cudaStreamWaitEvent(H->st, H->ev_reg, 0));
    H->reg_pending = false;
    LcbProfScope ps("k_deconv_update", H->st);
    k_deconv_update<<<DU_CTAS, DU_THREADS, 0, H->st>>>(H->D, it, n_iter, lr, schedule, seq, wr, gh, gc, ls);
    LCB_CUDA(cudaGetLastError());
    return LCB_OK;
}

static int next_seq(DeconvHandle* H) { return (H->D.cm.world > 1) ? ++H->seq : 0; }

static int check_peers(DeconvHandle* H) {
    if (H->D.cm.world <= 1) return LCB_OK;
    float flag = 0.f;
    LCB_CUDA(cudaMemcpyAsync(&flag, H->D.ctl + 7, 4, cudaMemcpyDeviceToHost, H->st));
    LCB_CUDA(cudaStreamSynchronize(H->st));
    if (flag != 0.f) { lcb_set_error("joint deconvolution: a peer rank did not deliver its gradient within the time limit"); return LCB_ERR_CUDA; }
    return LCB_OK;
}

extern "C" {

int lcb_deconv_destroy(void* handle) {
    DeconvHandle* H = (DeconvHandle*)handle;
    if (!H) return LCB_OK;
    cudaStreamSynchronize(H->st);
    if (H->st2) { cudaStreamSynchronize(H->st2); cudaStreamDestroy(H->st2); }
    if (H->st_own) cudaStreamDestroy(H->st_own);
    if (H->ev_h) cudaEventDestroy(H->ev_h);
    if (H->ev_go) cudaEventDestroy(H->ev_go);
    if (H->ev_reg) cudaEventDestroy(H->ev_reg);
    for (void* p : H->ipc_open) cudaIpcCloseMemHandle(p);
    if (H->comm_buf) cudaFree(H->comm_buf);
    for (void* p : H->owned) cudaFree(p);
    delete H;
    return LCB_OK;
}

int lcb_deconv_create(const lcb_deconv_problem* p, int mem, void* stream, void** handle) {
    LCB_REQUIRE(p && handle, "lcb_deconv_create: NULL argument");
    LCB_REQUIRE(p->E >= 1 && p->n >= 4 && p->k >= 1 && p->k <= 4 && p->P >= 1 && p->M >= 0 && p->M <= DC_MMAX,
                "lcb_deconv_create: bad sizes E=%d n=%d k=%d P=%d M=%d (M <= %d)", p->E, p->n, p->k, p->P, p->M, DC_MMAX);
    LCB_REQUIRE(p->data && p->weight && p->psf, "lcb_deconv_create: NULL input array");
    if (lcb_device_count() == 0) { lcb_set_error("no CUDA device: liblcb has no CPU fallback"); return LCB_ERR_CUDA; }
    DeconvHandle* H = new DeconvHandle();
    H->st = (cudaStream_t)stream;
    H->st_own = nullptr;
    if (!H->st) {
        // a BLOCKING stream of our own: it keeps the implicit ordering with the legacy default stream (torch's current stream
        // unless the caller changed it) and, unlike the legacy stream, can be captured into a CUDA graph
        if (cudaStreamCreate(&H->st_own) != cudaSuccess) { lcb_set_error("lcb_deconv_create: cannot create a stream"); delete H; return LCB_ERR_CUDA; }
        H->st = H->st_own;
    }
    H->comm_buf = nullptr; H->CS = 0; H->CS_user = 0; H->seq = 0; H->lbfgs = nullptr; H->lbfgs_floats = 0;
    H->st2 = nullptr; H->ev_h = nullptr; H->ev_go = nullptr; H->ev_reg = nullptr; H->reg_pending = false; H->starlet_attr = false; H->noise_tab = nullptr; H->W_spare = nullptr;
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);     // the 8 starlet CTAs must get their SMs before the epoch grid fills the GPU
    if (cudaStreamCreateWithPriority(&H->st2, cudaStreamNonBlocking, prio_hi) != cudaSuccess ||
        cudaEventCreateWithFlags(&H->ev_go, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&H->ev_h, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&H->ev_reg, cudaEventDisableTiming) != cudaSuccess) {
        lcb_set_error("lcb_deconv_create: cannot create the auxiliary stream"); delete H; return LCB_ERR_CUDA;
    }
    if (const char* ev = getenv("LCB_DECONV_CS")) H->CS_user = atoi(ev);
    DeconvDev& D = H->D;
    memset(&D, 0, sizeof(D));
    D.E = p->E; D.n = p->n; D.k = p->k; D.nu = p->n * p->k; D.P = p->P; D.M = p->M;
    D.E_total = p->E; D.e0 = 0;
    D.cm.world = 1; D.cm.rank = 0;
    D.pts_all_epochs = 1; D.fu_relative = 1;
    D.tot = D.nu * D.nu + 6 * D.M + 2;
    D.tot_pad = (D.tot + 3) & ~3;
    D.cv = lcb_devconv();
    D.G = D.cv.G;
    LCB_REQUIRE(D.G <= 16, "gauss_taps must be <= 16");
    const int j0 = (p->P - 1) / 2;
    D.A0 = (j0 - p->P + 1 >= 0) ? (j0 - p->P + 1) / p->k : -((-(j0 - p->P + 1) + p->k - 1) / p->k);
    const int A1 = (j0 + p->k - 1) / p->k;
    D.NA = A1 - D.A0 + 1;
    int J = 0; while ((1 << (J + 1)) <= D.nu) ++J;
    D.J = J;
    const size_t E = D.E, nn = (size_t)D.n * D.n, pp = (size_t)D.nu * D.nu, np = D.M + 3, kk = (size_t)D.k * D.k;
    int rc = 0;
#define AL(field, count, zero) if ((rc = dalloc(H, (void**)&D.field, (count) * 4, zero))) { lcb_deconv_destroy(H); return rc; }
    AL(data, E * nn, false) AL(weight, E * nn, false) AL(S, E * kk * D.NA * (size_t)((D.NA + 7) & ~7), false)
    AL(h, pp, true) AL(h_mu, pp, true) AL(h_nu, pp, true)
    AL(c, 2 * (size_t)DC_MMAX, true) AL(c_mu, 2 * (size_t)DC_MMAX, true) AL(c_nu, 2 * (size_t)DC_MMAX, true)
    AL(ep, E * np, true) AL(ep_mu, E * np, true) AL(ep_nu, E * np, true) AL(ep_g, E * np, true)
    AL(alpha, E, true) AL(Gh, E * pp, true) AL(gc, E * 2 * (size_t)DC_MMAX, true) AL(eloss, E, true)
    AL(red, pp + 6 * DC_MMAX + 2, true) AL(ctl, 8, true) AL(fu, 3 * (size_t)DC_MMAX, true) AL(gpart, 16 * DU_CTAS, true)
    AL(planes, (3 + (size_t)J) * pp, true) AL(model, E * nn, true) AL(prior, 4 * (size_t)DC_MMAX, true)
    AL(band_ctr, (size_t)DC_CSMAX + 1, true) AL(ictl, 2, true)
    D.GS = 1; while (D.GS * D.GS < D.E) ++D.GS;          // ~ sqrt(E) epochs per group
    {   // few local epochs (the 8-GPU shard of cfg4 holds 25): ONE level -- the last cluster of a band walks E planes of 8 KB, which
        // costs less than the second round of counters, fences and L2 round trips
        const char* ev = getenv("LCB_DC_ONELEVEL_MAX");
        const int one_level_max = ev ? atoi(ev) : 32;
        if (D.E <= one_level_max) D.GS = D.E;
    }
    D.NG = (D.E + D.GS - 1) / D.GS;
    AL(grp_ctr, (size_t)D.NG * DC_CSMAX, true) AL(Gp, (size_t)D.NG * pp, true)
#undef AL
    if ((rc = dalloc(H, (void**)&H->gh, pp * 4, true)) || (rc = dalloc(H, (void**)&H->gcx, 2 * DC_MMAX * 4, true)) ||
        (rc = dalloc(H, (void**)&H->ls, 4, true))) { lcb_deconv_destroy(H); return rc; }
    D.W = nullptr;                                    // allocated by lcb_deconv_set_reg when weights are given
    H->loss_cap = 0; D.loss_hist = nullptr;
    // inputs
    float* psf_d = nullptr;
    if ((rc = dalloc(H, (void**)&psf_d, E * (size_t)p->P * p->P * 4, false))) { lcb_deconv_destroy(H); return rc; }
    if ((rc = put(H, D.data, p->data, E * nn, mem)) || (rc = put(H, D.weight, p->weight, E * nn, mem)) ||
        (rc = put(H, psf_d, p->psf, E * (size_t)p->P * p->P, mem))) { lcb_deconv_destroy(H); return rc; }
    k_deconv_fold_psf<<<D.E, 256, 0, H->st>>>(psf_d, D.S, D.E, D.P, D.k, D.NA, D.A0);
    if (cudaGetLastError() != cudaSuccess) { lcb_set_error("k_deconv_fold_psf launch failed"); lcb_deconv_destroy(H); return LCB_ERR_CUDA; }
    D.free_h = D.free_mean = D.free_a = D.free_c = D.free_d = 1;
    H->smem_epoch = 0;
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&H->max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    cudaDeviceGetAttribute(&H->n_sm, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&H->smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev);
    if ((size_t)dc_layout(D.n, D.k, D.NA, D.A0, DC_CSMAX).total * 4 > (size_t)H->max_smem) {
        lcb_set_error("deconvolution: n=%d k=%d P=%d does not fit the shared memory of one SM even with %d CTAs per epoch", D.n, D.k, D.P, DC_CSMAX);
        lcb_deconv_destroy(H);
        return LCB_ERR_ARG;
    }
    *handle = H;
    return LCB_OK;
}

int lcb_deconv_set_cluster(void* handle, int ctas_per_epoch) {
    DeconvHandle* H = (DeconvHandle*)handle;
    LCB_REQUIRE(H, "lcb_deconv_set_cluster: NULL handle");
    LCB_REQUIRE(ctas_per_epoch >= 0 && ctas_per_epoch <= DC_CSMAX,
                "lcb_deconv_set_cluster: CTAs per epoch must be 0 (automatic) or 1 .. %d", DC_CSMAX);
    H->CS_user = ctas_per_epoch; H->CS = 0;
    return LCB_OK;
}

int lcb_deconv_get_cluster(void* handle) {
    DeconvHandle* H = (DeconvHandle*)handle;
    if (!H) return 0;
    return H->CS ? H->CS : choose_cluster(H);
}

// epochs over all ranks, global index of the first local epoch, and (may be NULL) the per-source shift used for
// the flux statistics (any value near the mean flux of each source; identical on every rank)
int lcb_deconv_set_global(void* handle, int E_total, int e0, const float* flux_shift, int mem) {
    DeconvHandle* H = (DeconvHandle*)handle;
    LCB_REQUIRE(H && E_total >= H->D.E && e0 >= 0, "lcb_deconv_set_global: bad arguments");
    H->D.E_total = E_total; H->D.e0 = e0;
    return put(H, H->D.fu, flux_shift, H->D.M, mem);
}

// ---- in-kernel all-reduce over peer memory: every rank allocates a receive buffer, exports it with CUDA IPC
//      (comm_init), the caller exchanges the 64-byte handles (any transport) and every rank maps its peers (comm_connect)
int lcb_deconv_comm_init(void* handle, int rank, int world, void* ipc_handle_out) {
    DeconvHandle* H = (DeconvHandle*)handle;
    LCB_REQUIRE(H && ipc_handle_out && world >= 1 && world <= DC_MAXW && rank >= 0 && rank < world,
                "lcb_deconv_comm_init: bad arguments (world <= %d)", DC_MAXW);
    LCB_REQUIRE(!H->comm_buf, "lcb_deconv_comm_init: already initialised");
    DeconvDev& D = H->D;
    const size_t slots_bytes = ((size_t)2 * world * D.tot_pad * 4 + 255) & ~(size_t)255;
    const size_t bytes = slots_bytes + (size_t)2 * world * 4 + 256;
    LCB_CUDA(cudaMalloc(&H->comm_buf, bytes));
    LCB_CUDA(cudaMemset(H->comm_buf, 0, bytes));
    LCB_CUDA(cudaDeviceSynchronize());
    cudaIpcMemHandle_t hd;
    LCB_CUDA(cudaIpcGetMemHandle(&hd, H->comm_buf));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
    memcpy(ipc_handle_out, &hd, 64);
    D.cm.world = world; D.cm.rank = rank;
    D.cm.slots[rank] = (float*)H->comm_buf;
    D.cm.flags[rank] = (int*)((char*)H->comm_buf + slots_bytes);
    int rc;
    if ((rc = dalloc(H, (void**)&D.cm.ctr, 2 * sizeof(unsigned), true))) return rc;
    LCB_CUDA(cudaStreamSynchronize(H->st));
    if (world == 1) { D.cm.world = 1; }
    return LCB_OK;
}

int lcb_deconv_comm_connect(void* handle, const void* all_handles) {
    DeconvHandle* H = (DeconvHandle*)handle;
    LCB_REQUIRE(H && all_handles && H->comm_buf, "lcb_deconv_comm_connect: call lcb_deconv_comm_init first");
    DeconvDev& D = H->D;
    const size_t slots_bytes = ((size_t)2 * D.cm.world * D.tot_pad * 4 + 255) & ~(size_t)255;
    for (int r = 0; r < D.cm.world; ++r) {
        if (r == D.cm.rank) continue;
        cudaIpcMemHandle_t hd;
        memcpy(&hd, (const char*)all_handles + (size_t)r * 64, 64);
        void* p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, hd, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            lcb_set_error("cudaIpcOpenMemHandle(rank %d): %s", r, cudaGetErrorString(e));
            cudaGetLastError();
            return LCB_ERR_CUDA;
        }
        H->ipc_open.push_back(p);
        D.cm.slots[r] = (float*)p;
        D.cm.flags[r] = (int*)((char*)p + slots_bytes);
    }
    return LCB_OK;
}

int lcb_deconv_set_params(void* handle, const lcb_deconv_params* q, int mem) {
    DeconvHandle* H = (DeconvHandle*)handle;
    LCB_REQUIRE(H && q, "lcb_deconv_set_params: NULL argument");
    DeconvDev& D = H->D;
    const size_t E = D.E, pp = (size_t)D.nu * D.nu, np = D.M + 3;
    int rc;
    if ((rc = put(H, D.h, q->h, pp, mem))) return rc;
    if ((rc = put(H, D.c, q->c_x, D.M, mem)) || (rc = put(H, D.c + D.M, q->c_y, D.M, mem))) return rc;
    if ((rc = put(H, D.alpha, q->alpha, E, mem))) return rc;
    // per-epoch block [E][M+3] is assembled on the host side of the copy: strided device copies
    if (q->a && D.M > 0) LCB_CUDA(cudaMemcpy2DAsync(D.ep, np * 4, q->a, D.M * 4, D.M * 4, E, mem == LCB_MEM_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, H->st));
    const float* cols[3] = {q->dx, q->dy, q->mean};
    for (int c = 0; c < 3; ++c)
        if (cols[c]) LCB_CUDA(cudaMemcpy2DAsync(D.ep + D.M + c, np * 4, cols[c], 4, 4, E, mem == LCB_MEM_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, H->st));
    // restart the optimiser
    LCB_CUDA(cudaMemsetAsync(D.h_mu, 0, pp * 4, H->st)); LCB_CUDA(cudaMemsetAsync(D.h_nu, 0, pp * 4, H->st));
    LCB_CUDA(cudaMemsetAsync(D.c_mu, 0, 2 * DC_MMAX * 4, H->st)); LCB_CUDA(cudaMemsetAsync(D.c_nu, 0, 2 * DC_MMAX * 4, H->st));
    LCB_CUDA(cudaMemsetAsync(D.ep_mu, 0, E * np * 4, H->st)); LCB_CUDA(cudaMemsetAsync(D.ep_nu, 0, E * np * 4, H->st));
    LCB_CUDA(cudaMemsetAsync(D.ctl, 0, 7 * 4, H->st));
    D.free_h = q->free_h; D.free_mean = q->free_mean; D.free_a = q->free_a; D.free_c = q->free_c; D.free_d = q->free_d;
    LCB_CUDA(cudaStreamSynchronize(H->st));
    return LCB_OK;
}

int lcb_deconv_set_reg(void* handle, const lcb_deconv_reg* r, int mem) {
    DeconvHandle* H = (DeconvHandle*)handle;
    LCB_REQUIRE(H && r, "lcb_deconv_set_reg: NULL argument");
    DeconvDev& D = H->D;
    const size_t pp = (size_t)D.nu * D.nu;
    D.lam_scales = r->lam_scales; D.lam_hf = r->lam_hf; D.lam_pos = r->lam_pos;
    D.lam_pts = r->lam_pts; D.lam_fu = r->lam_fu; D.pts_all_epochs = r->pts_all_epochs; D.fu_relative = r->fu_relative;
    int rc;
    if (r->W) {
        if (!D.W) { if (H->W_spare) { D.W = H->W_spare; H->W_spare = nullptr; } else if ((rc = dalloc(H, (void**)&D.W, (size_t)D.J * pp * 4, false))) return rc; }
        if ((rc = put(H, D.W, r->W, (size_t)D.J * pp, mem))) return rc;
    } else { if (D.W) H->W_spare = D.W; D.W = nullptr; }
    D.has_prior = (r->prior_mu_x && r->prior_sig_x && r->prior_mu_y && r->prior_sig_y) ? 1 : 0;
    if (D.has_prior) {
        if ((rc = put(H, D.prior, r->prior_mu_x, D.M, mem)) || (rc = put(H, D.prior + D.M, r->prior_sig_x, D.M, mem)) ||
            (rc = put(H, D.prior + 2 * D.M, r->prior_mu_y, D.M, mem)) || (rc = put(H, D.prior + 3 * D.M, r->prior_sig_y, D.M, mem))) return rc;
    }
    return LCB_OK;
}

// local half of one iteration: per-epoch kernel + reduction over the local epochs into red[]
int lcb_deconv_step_local(void* handle, int want_model) {
    DeconvHandle* H = (DeconvHandle*)handle;
    LCB_REQUIRE(H, "lcb_deconv_step_local: NULL handle");
    int rc;
    if ((rc = launch_starlet(H)) || (rc = launch_epoch(H, want_model ? 1 : 0))) return rc;
    return launch_reduce(H, 0, 0);
}

// device pointer and length of the buffer to all-reduce (sum) across ranks between the two halves
int lcb_deconv_reduce_buffer(void* handle, float** ptr, int* count) {
    DeconvHandle* H = (DeconvHandle*)handle;
    LCB_REQUIRE(H && ptr && count, "lcb_deconv_reduce_buffer: NULL argument");
    *ptr = H->D.red; *count = H->D.tot;
    return LCB_OK;
}

// applies the pending per-epoch AdaBelief update left by the last lcb_deconv_step_update and clears the pending flag
// (ctl[4]): the last call of an external-collective loop (lcb_deconv_run does the same internally)
int lcb_deconv_flush(void* handle) {
    DeconvHandle* H = (DeconvHandle*)handle;
    LCB_REQUIRE(H, "lcb_deconv_flush: NULL handle");
    int rc;
    if ((rc = launch_epoch(H, 0))) return rc;
    LCB_CUDA(cudaMemsetAsync(H->D.ctl + 4, 0, 4, H->st));
    return LCB_OK;
}

// replicated half: regularisers, global norm, AdaBelief on the shared parameters; it < 0 = evaluate only
int lcb_deconv_step_update(void* handle, int it, int n_iter, float lr, int schedule) {
    DeconvHandle* H = (DeconvHandle*)handle;
    LCB_REQUIRE(H, "lcb_deconv_step_update: NULL handle");
    return launch_update(H, it, n_iter, lr, schedule, 0, nullptr, nullptr, nullptr);
}

// n_iter AdaBelief iterations enqueued back to back, no host synchronisation inside the loop.  With a connected
// communicator (lcb_deconv_comm_*) every rank calls this with the same options: the gradient of the shared
// parameters is exchanged inside k_deconv_reduce / k_deconv_update over peer memory.
// Stage 1 of do_modelling_of_roi on the device: minimises the current loss over {dx, dy, a} (whatever the free flags say about
// the other parameters, they stay where they are), at most maxiter iterations.  Single rank only (the sharded stage 1 keeps
// the host optimiser over lcb_deconv_loss_grad).  info: [0] iterations, [1] evaluations, [2] stop reason (1 ftol, 2 pgtol,
// 3 maxiter, 4 line search stalled, 0 evaluation budget exhausted), [3] final loss, [4] |projected gradient|_inf.
int lcb_deconv_lbfgs(void* handle, int maxiter, float a_lower, float ftol, float pgtol, float* loss_hist, float* info, int mem) {
    LcbRange nvtx_range("lcb_deconv_lbfgs");
    DeconvHandle* H = (DeconvHandle*)handle;
    LCB_REQUIRE(H && maxiter >= 1, "lcb_deconv_lbfgs: bad arguments");
    DeconvDev& D = H->D;
    LCB_REQUIRE(D.cm.world <= 1, "lcb_deconv_lbfgs: single-rank handles only (sharded stage 1: host L-BFGS-B over lcb_deconv_loss_grad)");
    const int nv = D.E * (D.M + 2);
    const size_t need = (size_t)nv * (3 + 2 * LB_M) + LB_M + 16 + (size_t)maxiter;
    int rc;
    if (need > H->lbfgs_floats) {
        if ((rc = dalloc(H, (void**)&H->lbfgs, need * 4, false))) return rc;
        H->lbfgs_floats = need;
    }
    LbState L;
    L.x = H->lbfgs; L.g = L.x + nv; L.d = L.g + nv; L.S = L.d + nv; L.Y = L.S + (size_t)LB_M * nv;
    L.rho = L.Y + (size_t)LB_M * nv; L.sc = L.rho + LB_M; L.hist = L.sc + 16;
    L.nv = nv; L.maxiter = maxiter; L.a_lo = a_lower; L.d_max = 0.5f * (float)D.n; L.ftol = ftol; L.pgtol = pgtol;
    LCB_CUDA(cudaMemsetAsync(L.rho, 0, (LB_M + 16 + (size_t)maxiter) * 4, H->st));
    const int max_evals = 20 * maxiter + 20;              // scipy's maxfun as the reference-shaped host path passes it
    int evals = 0;
    float flag = 0.f;
    while (evals < max_evals && flag == 0.f) {
        const int chunk = (max_evals - evals < 16) ? max_evals - evals : 16;
        for (int c = 0; c < chunk; ++c) {
            LCB_CUDA(cudaMemsetAsync(D.ctl + 4, 0, 4, H->st));
            if ((rc = launch_starlet(H)) || (rc = launch_epoch(H, 0)) || (rc = launch_reduce(H, 0, 0)) ||
                (rc = launch_update(H, -1, 1, 0.f, 0, 0, H->gh, H->gcx, H->ls))) return rc;
            if (D.lam_fu != 0.f && D.M > 0) k_deconv_fu_apply<<<(D.E * D.M + 255) / 256, 256, 0, H->st>>>(D);
            { LcbProfScope ps("k_deconv_lbfgs_step", H->st); k_deconv_lbfgs_step<<<1, LB_THREADS, 0, H->st>>>(D, L, H->ls); }
            LCB_CUDA(cudaGetLastError());
        }
        evals += chunk;
        LCB_CUDA(cudaMemcpyAsync(&flag, L.sc + 5, 4, cudaMemcpyDeviceToHost, H->st));     // ONE 4-byte read per 16 evaluations
        LCB_CUDA(cudaStreamSynchronize(H->st));
    }
    k_deconv_lbfgs_finish<<<(nv + 255) / 256, 256, 0, H->st>>>(D, L);
    LCB_CUDA(cudaGetLastError());
    float scl[16];
    LCB_CUDA(cudaMemcpyAsync(scl, L.sc, sizeof(scl), cudaMemcpyDeviceToHost, H->st));
    LCB_CUDA(cudaStreamSynchronize(H->st));
    const int iters = (int)scl[3];
    if (info) {
        const float inf[5] = {scl[3], scl[6], scl[5], scl[0], scl[8]};
        if (mem == LCB_MEM_HOST) memcpy(info, inf, sizeof(inf));
        else LCB_CUDA(cudaMemcpy(info, inf, sizeof(inf), cudaMemcpyHostToDevice));
    }
    if (loss_hist && iters > 0) { if ((rc = get(H, loss_hist, L.hist, (size_t)(iters < maxiter ? iters : maxiter), mem))) return rc; }
    return LCB_OK;
}

// One run = begin (loss buffer, device-resident counters, first iteration eagerly, ONE captured CUDA graph of an iteration),
// n_iter - 1 replays, end (pending per-epoch update, peers, loss history).  lcb_deconv_run drives one handle; lcb_deconv_run_many
// interleaves the replays of several handles (each on its own stream), so that many small joint fits fill the GPU together.
struct DeconvRun {
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t gexec = nullptr;
    bool use_graph = false, multi = false;
    int done = 0, seq0 = 0;
};

// Where the reduction over the local epochs runs: as its own grid-wide kernel (k_deconv_reduce: the default) or in the tail of the
// epoch kernel (last cluster per band, two levels above 32 epochs).  Both push the sums to the peers.  Measured on cfg4 shapes with
// CUDA-graph replays (profiles/README.md, last session of round 2).  Before k_deconv_reduce summed its per-epoch scalars cooperatively
// it took 8 us + 0.9 us per 4 local epochs (53 us at 200) and the routes were within a few per cent of each other: one GPU, 200 epochs,
// fused 2336 vs separate 2314 it/s; two GPUs (100 per rank) 3427 vs 3651; eight GPUs (25 per rank) 9337 vs 9288.  With the reduce kernel
// at 14 us for any shard size the separate route wins on one GPU at every size (200 epochs 2548 vs 2332 it/s, 100: 3711 vs 3485, 50: 6882 vs
// 6398, 25: 10738 vs 10167): a few CTAs walking the planes at the very end of a grid cost more than a kernel boundary.
// LCB_DECONV_REDUCE=fused|separate selects the route explicitly (A/B timing; the fused route stays tested).
static bool fused_reduction(const DeconvHandle*) {
    static const int forced = [] {
        const char* e = getenv("LCB_DECONV_REDUCE");
        return !e ? 0 : (e[0] == 'f' ? 1 : (e[0] == 's' ? 2 : 0));
    }();
    return forced == 1;
}

static int run_iteration(DeconvHandle* H, const lcb_fit_opts* opt, int it_arg, int seq_arg) {
    int r;
    const bool fused = fused_reduction(H);
    if ((r = launch_starlet(H)) || (r = launch_epoch(H, fused ? 4 : 0, seq_arg)) || (!fused && (r = launch_reduce(H, 0, seq_arg))) ||
        (r = launch_update(H, it_arg, opt->n_iter, opt->lr, opt->schedule, seq_arg, nullptr, nullptr, nullptr))) return r;
    return LCB_OK;
}

static void run_release(DeconvRun& R) {
    if (R.gexec) cudaGraphExecDestroy(R.gexec);
    if (R.graph) cudaGraphDestroy(R.graph);
    R.gexec = nullptr; R.graph = nullptr;
}

static int run_begin(DeconvHandle* H, const lcb_fit_opts* opt, DeconvRun& R) {
    DeconvDev& D = H->D;
    int rc;
    if (opt->n_iter > H->loss_cap) {
        if ((rc = dalloc(H, (void**)&D.loss_hist, (size_t)opt->n_iter * 4, true))) return rc;
        H->loss_cap = opt->n_iter;
    }
    // One iteration = starlet term (second stream) | epoch kernel with the fused reduction / NVLink push in its tail | update.
    // The first iteration is launched eagerly (it also sets the function attributes); the remaining ones replay ONE captured
    // CUDA graph of an iteration whose kernels read the iteration index and the exchange sequence number from device memory
    // (identical launch arguments), which removes the per-launch gaps of a launch-bound loop.  With per-kernel profiling
    // enabled (lcb_profile_enable) or LCB_DECONV_GRAPH=0 every iteration is launched eagerly.
    R.multi = D.cm.world > 1;
    R.seq0 = H->seq + 1;
    if (opt->n_iter > 0) {
        const int init[2] = {0, R.seq0};
        LCB_CUDA(cudaMemcpyAsync(D.ictl, init, sizeof(init), cudaMemcpyHostToDevice, H->st));
        LCB_CUDA(cudaStreamSynchronize(H->st));            // init is a stack temporary
    }
    const char* genv = getenv("LCB_DECONV_GRAPH");
    R.use_graph = !(genv && genv[0] == '0') && !lcb_profiling() && opt->n_iter >= 8;
    R.done = 0;
    if (R.use_graph) {
        if ((rc = run_iteration(H, opt, -2, -1))) return rc;   // eager: function attributes, cluster size
        R.done = 1;
        cudaError_t ce = cudaStreamBeginCapture(H->st, cudaStreamCaptureModeThreadLocal);
        bool ok = (ce == cudaSuccess);
        if (ok) {
            rc = run_iteration(H, opt, -2, -1);
            ce = cudaStreamEndCapture(H->st, &R.graph);
            ok = (rc == LCB_OK && ce == cudaSuccess && R.graph != nullptr);
        }
        if (ok) ok = (cudaGraphInstantiate(&R.gexec, R.graph, 0) == cudaSuccess);
        if (!ok) { cudaGetLastError(); run_release(R); }       // capture unavailable: eager launches
        H->reg_pending = false;
    }
    return LCB_OK;
}

// launches up to `count` further iterations (graph replay when available)
static int run_some(DeconvHandle* H, const lcb_fit_opts* opt, DeconvRun& R, int count) {
    int rc;
    for (int i = 0; i < count && R.done < opt->n_iter; ++i, ++R.done) {
        if (R.gexec) {
            if (cudaGraphLaunch(R.gexec, H->st) != cudaSuccess) {
                lcb_set_error("joint deconvolution: cudaGraphLaunch failed: %s", cudaGetErrorString(cudaGetLastError()));
                return LCB_ERR_CUDA;
            }
        } else if ((rc = run_iteration(H, opt, R.use_graph ? -2 : R.done, R.use_graph ? -1 : (R.multi ? R.seq0 + R.done : 0)))) return rc;
    }
    return LCB_OK;
}

static int run_end(DeconvHandle* H, const lcb_fit_opts* opt, DeconvRun& R, float* loss_hist, int mem) {
    DeconvDev& D = H->D;
    int rc;
    run_release(R);
    if (R.multi) H->seq += opt->n_iter;
    // flush the pending per-epoch update so that get() sees the final parameters
    if (opt->n_iter > 0) { if ((rc = launch_epoch(H, 0))) return rc; LCB_CUDA(cudaMemsetAsync(D.ctl + 4, 0, 4, H->st)); }
    if ((rc = check_peers(H))) return rc;
    if (loss_hist) { if ((rc = get(H, loss_hist, D.loss_hist, opt->n_iter, mem))) return rc; }
    return LCB_OK;
}

int lcb_deconv_run(void* handle, const lcb_fit_opts* opt, float* loss_hist, int mem) {
    LcbRange nvtx_range("lcb_deconv_run");
    DeconvHandle* H = (DeconvHandle*)handle;
    LCB_REQUIRE(H && opt && opt->n_iter >= 0, "lcb_deconv_run: bad arguments");
    DeconvRun R;
    int rc;
    if ((rc = run_begin(H, opt, R)) || (rc = run_some(H, opt, R, opt->n_iter))) { run_release(R); return rc; }
    return run_end(H, opt, R, loss_hist, mem);
}

int lcb_deconv_run_many(void* const* handles, int count, const lcb_fit_opts* opt, float* const* loss_hist, int mem) {
    LcbRange nvtx_range("lcb_deconv_run_many");
    LCB_REQUIRE(handles && count >= 0 && opt && opt->n_iter >= 0, "lcb_deconv_run_many: bad arguments");
    std::vector<DeconvRun> R((size_t)count);
    int rc = LCB_OK;
    for (int i = 0; i < count && !rc; ++i) {
        DeconvHandle* H = (DeconvHandle*)handles[i];
        if (!H || H->D.cm.world > 1) { lcb_set_error("lcb_deconv_run_many: handle %d is NULL or sharded over ranks", i); rc = LCB_ERR_ARG; break; }
        rc = run_begin(H, opt, R[i]);
    }
    // round-robin over the handles, a few iterations at a time: every stream always has work queued, no stream's queue overflows
    for (int it = 0; it < opt->n_iter && !rc; it += 4)
        for (int i = 0; i < count && !rc; ++i) rc = run_some((DeconvHandle*)handles[i], opt, R[i], 4);
    for (int i = 0; i < count; ++i) {
        if (rc) { run_release(R[i]); continue; }
        rc = run_end((DeconvHandle*)handles[i], opt, R[i], loss_hist ? loss_hist[i] : nullptr, mem);
    }
    return rc;
}

// loss and gradient at the current parameters (no update); collective when a communicator is connected
int lcb_deconv_loss_grad(void* handle, lcb_deconv_grad* g, int mem) {
    LcbRange nvtx_range("lcb_deconv_loss_grad");
    DeconvHandle* H = (DeconvHandle*)handle;
    LCB_REQUIRE(H && g, "lcb_deconv_loss_grad: NULL argument");
    DeconvDev& D = H->D;
    const size_t E = D.E, pp = (size_t)D.nu * D.nu, np = D.M + 3;
    int rc;
    LCB_CUDA(cudaMemsetAsync(D.ctl + 4, 0, 4, H->st));
    const int seq = next_seq(H);
    if ((rc = launch_starlet(H)) || (rc = launch_epoch(H, 0)) || (rc = launch_reduce(H, 0, seq)) ||
        (rc = launch_update(H, -1, 1, 0.f, 0, seq, H->gh, H->gcx, H->ls))) return rc;
    if (D.lam_fu != 0.f && D.M > 0) {
        k_deconv_fu_apply<<<(D.E * D.M + 255) / 256, 256, 0, H->st>>>(D);
        LCB_CUDA(cudaGetLastError());
    }
    if ((rc = check_peers(H))) return rc;
    if ((rc = get(H, g->loss, H->ls, 1, mem)) || (rc = get(H, g->h, H->gh, pp, mem)) || (rc = get(H, g->c_x, H->gcx, D.M, mem)) ||
        (rc = get(H, g->c_y, H->gcx + D.M, D.M, mem))) return rc;
    const cudaMemcpyKind kd = mem == LCB_MEM_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
    if (g->a && D.M > 0) LCB_CUDA(cudaMemcpy2DAsync(g->a, D.M * 4, D.ep_g, np * 4, D.M * 4, E, kd, H->st));
    float* cols[3] = {g->dx, g->dy, g->mean};
    for (int c = 0; c < 3; ++c)
        if (cols[c]) LCB_CUDA(cudaMemcpy2DAsync(cols[c], 4, D.ep_g + D.M + c, np * 4, 4, E, kd, H->st));
    LCB_CUDA(cudaStreamSynchronize(H->st));
    return LCB_OK;
}

int lcb_deconv_get(void* handle, lcb_deconv_params* q, float* model, float* loss, int mem) {
    DeconvHandle* H = (DeconvHandle*)handle;
    LCB_REQUIRE(H && q, "lcb_deconv_get: NULL argument");
    DeconvDev& D = H->D;
    const size_t E = D.E, pp = (size_t)D.nu * D.nu, np = D.M + 3, nn = (size_t)D.n * D.n;
    int rc;
    if (model || loss) {
        LCB_CUDA(cudaMemsetAsync(D.ctl + 4, 0, 4, H->st));
        if ((rc = launch_epoch(H, 1))) return rc;
        if (loss) {               // LOCAL loss (chi2 of the local epochs + replicated terms): no exchange here
            if ((rc = launch_starlet(H)) || (rc = launch_reduce(H, 0, 0)) || (rc = launch_update(H, -1, 1, 0.f, 0, 0, nullptr, nullptr, H->ls))) return rc;
            if ((rc = get(H, loss, H->ls, 1, mem))) return rc;
        }
        if ((rc = get(H, model, D.model, E * nn, mem))) return rc;
    }
    if ((rc = get(H, (float*)q->h, D.h, pp, mem)) || (rc = get(H, (float*)q->c_x, D.c, D.M, mem)) ||
        (rc = get(H, (float*)q->c_y, D.c + D.M, D.M, mem)) || (rc = get(H, (float*)q->alpha, D.alpha, E, mem))) return rc;
    const cudaMemcpyKind kd = mem == LCB_MEM_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
    if (q->a && D.M > 0) LCB_CUDA(cudaMemcpy2DAsync((float*)q->a, D.M * 4, D.ep, np * 4, D.M * 4, E, kd, H->st));
    float* cols[3] = {(float*)q->dx, (float*)q->dy, (float*)q->mean};
    for (int c = 0; c < 3; ++c)
        if (cols[c]) LCB_CUDA(cudaMemcpy2DAsync(cols[c], 4, D.ep + D.M + c, np * 4, 4, E, kd, H->st));
    LCB_CUDA(cudaStreamSynchronize(H->st));
    return LCB_OK;
}

// Noise weights for the starlet regulariser of h (roi_modelling.py:299, star_photometry.py:108), same
// definition as for the PSF grid: var(p) = sum_e sum_q A_e[q,p]^2 w_e[q] (A_e = warp, PSF, decimation),
// W_j = sqrt(var (*) psi_j^2).  stage 0: local var into the reduce buffer (all-reduce it when sharded);
// stage 1: W from the reduce buffer, installed in the handle (and copied to W_out when non-NULL).
int lcb_deconv_noise_weights(void* handle, int stage, float* W_out, int mem) {
    DeconvHandle* H = (DeconvHandle*)handle;
    LCB_REQUIRE(H, "lcb_deconv_noise_weights: NULL handle");
    DeconvDev& D = H->D;
    const size_t pp = (size_t)D.nu * D.nu;
    int rc;
    if (stage == 0) {
        if ((rc = launch_epoch(H, 2))) return rc;
        return launch_reduce(H, 1, 0);
    }
    float* tabd = H->noise_tab;
    if (!tabd) {                                     // built once per handle
        std::vector<float> tab;
        lcb_build_noise_table(D.nu, D.J, tab);
        if ((rc = dalloc(H, (void**)&tabd, tab.size() * 4, false))) return rc;
        LCB_CUDA(cudaMemcpyAsync(tabd, tab.data(), tab.size() * 4, cudaMemcpyHostToDevice, H->st));
        LCB_CUDA(cudaStreamSynchronize(H->st));      // tab is a host temporary
        H->noise_tab = tabd;
    }
    if (!D.W) { if (H->W_spare) { D.W = H->W_spare; H->W_spare = nullptr; } else if ((rc = dalloc(H, (void**)&D.W, (size_t)D.J * pp * 4, false))) return rc; }
    LCB_CUDA(cudaMemcpyAsync(D.planes, D.red, pp * 4, cudaMemcpyDeviceToDevice, H->st));
    if ((rc = lcb_noise_weights_launch(1, D.nu, D.J, tabd, D.W, D.planes, 2 * pp, H->st))) return rc;
    return get(H, W_out, D.W, (size_t)D.J * pp, mem);
}

}  // extern "C"
