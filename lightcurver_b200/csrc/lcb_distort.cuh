// lcb_distort.cuh -- field distortion of the narrow PSF (starred.psf.psf.apply_distortion / PSF(field_distortion=True) [R];
// call sites lightcurver/processes/psf_modelling.py:169-170, star_photometry.py:293-304, roi_file_preparation.py:169-180).
//
// kwargs_distortion = {dilation_x, dilation_y, shear}: each a first-order polynomial in the star's rescaled frame position
// (X, Y) of utilities/image_coordinates.py:4-25 (origin at the frame centre, so the PSF at the centre is undistorted):
//     ex = t0 X + t1 Y,  ey = t2 X + t3 Y,  sh = t4 X + t5 Y,      theta = (t0 .. t5) per frame
// The PSF seen by a star is the affine resampling of the frame's narrow PSF s about the grid centre c0 = (nu-1)/2:
//     s_i[v][u] = det * bilinear(s; c0 + A (u - c0, v - c0)),   A = [[1 + ex, sh], [sh, 1 + ey]],
// bilinear with zeros outside the grid; det = |A| keeps the integral of the PSF (mode 1) or det = 1 (mode 2).
#pragma once
#include "lcb_common.cuh"

struct LcbAffine { float ex, ey, sh, det; };

__device__ __forceinline__ LcbAffine lcb_affine(const float* __restrict__ th, float X, float Y, int mode) {
    LcbAffine A;
    A.ex = th[0] * X + th[1] * Y;
    A.ey = th[2] * X + th[3] * Y;
    A.sh = th[4] * X + th[5] * Y;
    A.det = (mode == 1) ? (1.f + A.ex) * (1.f + A.ey) - A.sh * A.sh : 1.f;
    return A;
}

struct LcbBilin {
    int i0, j0;            // top-left corner (column, row) of the cell
    float fx, fy;          // fractional position inside the cell
    float s00, s01, s10, s11;   // s[j0][i0], s[j0][i0+1], s[j0+1][i0], s[j0+1][i0+1] (0 outside)
    float rx, ry;          // offsets of the output pixel from the grid centre
};

// s: nu x nu plane with leading dimension lds (shared or global)
__device__ __forceinline__ LcbBilin lcb_bilin(const float* __restrict__ s, int lds, int nu, const LcbAffine& A, int u, int v) {
    LcbBilin b;
    const float c0 = 0.5f * (float)(nu - 1);
    b.rx = (float)u - c0; b.ry = (float)v - c0;
    const float qx = c0 + (1.f + A.ex) * b.rx + A.sh * b.ry;
    const float qy = c0 + A.sh * b.rx + (1.f + A.ey) * b.ry;
    const float fi = floorf(qx), fj = floorf(qy);
    b.fx = qx - fi; b.fy = qy - fj;
    // clamp far-away cells onto a cell that is entirely outside: all four corners read as zero
    b.i0 = (int)fminf(fmaxf(fi, -2.f), (float)nu);
    b.j0 = (int)fminf(fmaxf(fj, -2.f), (float)nu);
    const bool x0 = b.i0 >= 0 && b.i0 < nu, x1 = b.i0 + 1 >= 0 && b.i0 + 1 < nu;
    const bool y0 = b.j0 >= 0 && b.j0 < nu, y1 = b.j0 + 1 >= 0 && b.j0 + 1 < nu;
    b.s00 = (x0 && y0) ? s[b.j0 * lds + b.i0] : 0.f;
    b.s01 = (x1 && y0) ? s[b.j0 * lds + b.i0 + 1] : 0.f;
    b.s10 = (x0 && y1) ? s[(b.j0 + 1) * lds + b.i0] : 0.f;
    b.s11 = (x1 && y1) ? s[(b.j0 + 1) * lds + b.i0 + 1] : 0.f;
    return b;
}

__device__ __forceinline__ float lcb_bilin_value(const LcbBilin& b) {
    return (1.f - b.fy) * ((1.f - b.fx) * b.s00 + b.fx * b.s01) + b.fy * ((1.f - b.fx) * b.s10 + b.fx * b.s11);
}
