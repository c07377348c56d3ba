// Per-pair geometry of continuous_conv: coordinate mapping, filter coordinates, trilinear corner weights and
// the radial window functions.  float32 restatement of SURVEY Appendix A.2/A.3 (open3d.ml continuous_conv) and of
// utils/tools/losses.py:8-44 (windows).
#pragma once
#include "common.cuh"

namespace dmcf {

struct PairGeom {
    int i0;  // x0 | y0 << 8 | z0 << 16  (cell coordinates of the "0" corner)
    int i1;  // x1 | y1 << 8 | z1 << 16
    float wx0, wx1, wy0, wy1, wz0, wz1;  // per-axis corner weights (0 for folded / out-of-range corners)
};

__device__ __forceinline__ float window_value(int window, float fac, float q) {
    switch (window) {
        case DMCF_WIN_POLY6: {  // utils/tools/losses.py:9-12
            const float t = 1.0f - q;
            return fac * fminf(fmaxf(t * t * t, 0.0f), 1.0f);
        }
        case DMCF_WIN_CUBIC: {  // :13-20
            const float s = sqrtf(q);
            float v = 0.0f;
            if (q <= 1.0f) {
                const float u = 1.0f - s;
                v = (s <= 0.5f) ? 6.0f * (s * s * s - q) + 1.0f : 2.0f * u * u * u;
            }
            return fac * 4.0f / 3.0f * v;
        }
        case DMCF_WIN_LINEAR:  // :21-25
            return fac * (1.0f - sqrtf(q));
        case DMCF_WIN_PEAK:  // :26-30
            return fac * (1.0f - 2.0f * sqrtf(q) + q);
        case DMCF_WIN_CUBIC_GRAD: {  // :31-39
            const float s = sqrtf(q);
            float v = 0.0f;
            if (q <= 1.0f) {
                const float u = 1.0f - s;
                v = (s <= 0.5f) ? 18.0f * q - 12.0f * s : -6.0f * u * u;
            }
            return fac * 4.0f / 3.0f * v;
        }
        default:
            return 1.0f;
    }
}

// (x,y,z) = neighbour - centre, returns cube coordinates in [-0.5, 0.5]^3
__device__ __forceinline__ void map_to_cube(int mapping, float inv_extent, float& x, float& y, float& z) {
    if (mapping == DMCF_MAP_IDENTITY) {
        x *= inv_extent; y *= inv_extent; z *= inv_extent;
        return;
    }
    const float s2 = 2.0f * inv_extent;
    x *= s2; y *= s2; z *= s2;
    if (mapping == DMCF_MAP_BALL_TO_CUBE_RADIAL) {
        const float rad = sqrtf(x * x + y * y + z * z);
        const float amax = fmaxf(fabsf(x), fmaxf(fabsf(y), fabsf(z)));
        if (amax < 1e-8f) {
            x = y = z = 0.0f;
        } else {
            const float s = 0.5f * rad / amax;
            x *= s; y *= s; z *= s;
        }
        return;
    }
    // volume preserving: sphere -> cylinder -> cube
    {
        const float xy2 = x * x + y * y;
        const float sq = xy2 + z * z;
        const float n = sqrtf(sq);
        if (sq < 1e-12f) {
            x = y = z = 0.0f;
        } else if (1.25f * z * z > xy2) {
            const float s = sqrtf(3.0f * n / (n + fabsf(z)));
            x *= s; y *= s;
            z = copysignf(n, z);
        } else {
            const float s = n / sqrtf(xy2);
            x *= s; y *= s;
            z *= 1.5f;
        }
    }
    {
        const float sq = x * x + y * y;
        const float n = sqrtf(sq);
        const float four_over_pi = 1.27323954473516f;
        if (sq < 1e-12f) {
            x = y = 0.0f;
        } else if (fabsf(y) <= fabsf(x)) {
            const float t = copysignf(n, x);
            y = t * four_over_pi * atanf(y / x);
            x = t;
        } else {
            const float t = copysignf(n, y);
            x = t * four_over_pi * atanf(x / y);
            y = t;
        }
    }
    x *= 0.5f; y *= 0.5f; z *= 0.5f;
}

__device__ __forceinline__ void axis_weights(int interp, float g, int fs, int& i0, int& i1, float& w0, float& w1) {
    if (interp == DMCF_INTERP_LINEAR) {
        i0 = min(max((int)g, 0), fs - 1);
        i1 = min(i0 + 1, fs - 1);
        float a = fminf(fmaxf(g - (float)i0, 0.0f), 1.0f);
        if (i1 == i0) a = 0.0f;  // fold the clamped corner into corner 0 (same cell, weights sum to 1)
        w0 = 1.0f - a;
        w1 = a;
    } else if (interp == DMCF_INTERP_LINEAR_BORDER) {
        const float f = floorf(g);
        const float a = g - f;
        const int j0 = (int)f, j1 = (int)f + 1;
        w0 = (j0 >= 0 && j0 <= fs - 1) ? 1.0f - a : 0.0f;
        w1 = (j1 >= 0 && j1 <= fs - 1) ? a : 0.0f;
        i0 = min(max(j0, 0), fs - 1);
        i1 = min(max(j1, 0), fs - 1);
    } else {  // nearest neighbour
        i0 = min(max((int)floorf(g + 0.5f), 0), fs - 1);
        i1 = i0;
        w0 = 1.0f;
        w1 = 0.0f;
    }
}

struct GeomParams {
    int kx, ky, kz;
    int mapping, interp, align_corners;
    float inv_extent;
    float offx, offy, offz;
};

__device__ __forceinline__ PairGeom pair_geometry(const GeomParams& gp, float dx, float dy, float dz) {
    map_to_cube(gp.mapping, gp.inv_extent, dx, dy, dz);
    float gx, gy, gz;
    if (gp.align_corners) {
        gx = (dx + 0.5f) * (float)(gp.kx - 1) + gp.offx;
        gy = (dy + 0.5f) * (float)(gp.ky - 1) + gp.offy;
        gz = (dz + 0.5f) * (float)(gp.kz - 1) + gp.offz;
    } else {
        gx = (dx + 0.5f) * (float)gp.kx - 0.5f + gp.offx;
        gy = (dy + 0.5f) * (float)gp.ky - 0.5f + gp.offy;
        gz = (dz + 0.5f) * (float)gp.kz - 0.5f + gp.offz;
    }
    PairGeom r;
    int x0, x1, y0, y1, z0, z1;
    axis_weights(gp.interp, gx, gp.kx, x0, x1, r.wx0, r.wx1);
    axis_weights(gp.interp, gy, gp.ky, y0, y1, r.wy0, r.wy1);
    axis_weights(gp.interp, gz, gp.kz, z0, z1, r.wz0, r.wz1);
    r.i0 = x0 | (y0 << 8) | (z0 << 16);
    r.i1 = x1 | (y1 << 8) | (z1 << 16);
    return r;
}

}  // namespace dmcf
