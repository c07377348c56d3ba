// Pieces shared by the register-patch continuous_conv kernels (cconv_wide.cu, cconv_lean.cu): compile-time filter
// grids, the per-base-cell scatter of the 8 trilinear corner weights into statically indexed registers, and the
// "base form" of a pair (corner block anchored at base <= fs-2 with folded weights).
#pragma once
#include "cconv_common.cuh"

namespace dmcf {

template <int KZ, int KY, int KX>
struct FilterGrid {
    static constexpr int KZ_ = KZ, KY_ = KY, KX_ = KX;
    static constexpr int K = KZ * KY * KX;
    static constexpr int NBX = KX > 1 ? KX - 1 : 1, NBY = KY > 1 ? KY - 1 : 1, NBZ = KZ > 1 ? KZ - 1 : 1;
    static constexpr int NB = NBX * NBY * NBZ;  // distinct "base" cells of the 2x2x2 corner block
};

// corner weights w[c], c = bx + 2*by + 4*bz, packed as wa = (w0..w3), wb = (w4..w7).
// The kernel instance owns the filter z-planes [ZLO, ZLO+NZ): corners on other planes belong to another launch.
template <int KZ, int KY, int KX, int ZLO, int NZ, int B>
__device__ __forceinline__ void scatter_case(float (&acc)[NZ * KY * KX], const float4& wa, const float4& wb, float f) {
    using G = FilterGrid<KZ, KY, KX>;
    if constexpr (B < G::NB) {
        constexpr int x0 = B % G::NBX, y0 = (B / G::NBX) % G::NBY, z0 = B / (G::NBX * G::NBY);
        constexpr int sx = 1, sy = KX, sz = KY * KX;
        if constexpr (z0 >= ZLO && z0 < ZLO + NZ) {
            constexpr int c000 = ((z0 - ZLO) * KY + y0) * KX + x0;
            acc[c000] = fmaf(wa.x, f, acc[c000]);
            if constexpr (KX > 1) acc[c000 + sx] = fmaf(wa.y, f, acc[c000 + sx]);
            if constexpr (KY > 1) acc[c000 + sy] = fmaf(wa.z, f, acc[c000 + sy]);
            if constexpr (KX > 1 && KY > 1) acc[c000 + sx + sy] = fmaf(wa.w, f, acc[c000 + sx + sy]);
        }
        if constexpr (KZ > 1 && z0 + 1 >= ZLO && z0 + 1 < ZLO + NZ) {
            constexpr int c001 = ((z0 + 1 - ZLO) * KY + y0) * KX + x0;
            acc[c001] = fmaf(wb.x, f, acc[c001]);
            if constexpr (KX > 1) acc[c001 + sx] = fmaf(wb.y, f, acc[c001 + sx]);
            if constexpr (KY > 1) acc[c001 + sy] = fmaf(wb.z, f, acc[c001 + sy]);
            if constexpr (KX > 1 && KY > 1) acc[c001 + sx + sy] = fmaf(wb.w, f, acc[c001 + sx + sy]);
        }
        (void)sz;
    }
}

#define DMCF_SC(i) \
    case i:        \
        scatter_case<KZ, KY, KX, ZLO, NZ, i>(acc, wa, wb, f); \
        break;
#define DMCF_SC8(i) DMCF_SC(i) DMCF_SC(i + 1) DMCF_SC(i + 2) DMCF_SC(i + 3) DMCF_SC(i + 4) DMCF_SC(i + 5) DMCF_SC(i + 6) DMCF_SC(i + 7)

template <int KZ, int KY, int KX, int ZLO, int NZ>
__device__ __forceinline__ void scatter_switch(int b, float (&acc)[NZ * KY * KX], const float4& wa, const float4& wb, float f) {
    static_assert(FilterGrid<KZ, KY, KX>::NB <= 64, "too many base cells");
    switch (b) {  // warp-uniform: every lane works on the same pair
        DMCF_SC8(0) DMCF_SC8(8) DMCF_SC8(16) DMCF_SC8(24) DMCF_SC8(32) DMCF_SC8(40) DMCF_SC8(48) DMCF_SC8(56)
        default: break;
    }
}

// one axis of the base form: the corner block always starts at base <= fs-2; a pair clamped onto the last cell
// (i0 == fs-1, folded weights) puts its whole weight on the upper corner.
__device__ __forceinline__ void base_axis(int fs, int i0, float w0, float w1, int& base, float& lo, float& hi) {
    if (fs == 1) {
        base = 0; lo = w0 + w1; hi = 0.0f;
    } else if (i0 > fs - 2) {
        base = fs - 2; lo = 0.0f; hi = w0 + w1;
    } else {
        base = i0; lo = w0; hi = w1;
    }
}


}  // namespace dmcf
