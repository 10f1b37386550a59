// ssim_tile.cuh - shared-memory tiling + separable 13-tap Gaussian passes used by the fused
// photometric kernels (photo.cu) and the stand-alone ssim op (warp_ops.cu).  ssim.py:9-36.
#pragma once
#include "geom.cuh"

namespace ccb {

constexpr int NT = 256;      // threads per CTA
constexpr int TW = 64;       // tile width  (centre pixels)
constexpr int TH = 20;       // tile height
constexpr int PXT = 5;       // centre pixels per thread (one column segment)
constexpr int HP = TW + 1;   // row pitch of the horizontally-blurred buffer (odd: conflict-free)

template <int HALO>
struct Tile {
    static constexpr int RW = TW + 2 * HALO;   // staged region
    static constexpr int RH = TH + 2 * HALO;
    static constexpr int PITCH = RW | 1;       // odd pitch: lanes == rows is conflict-free
    static constexpr int PLANE = RH * PITCH;
};

// ------------------------------------------------------------------------------------------------
// Separable Gaussian, horizontal pass.  lane = region row (RH == 32 for HALO 6), warp = 8-column
// group.  NMAP outputs per column: KIND 0: {x, x*x} (target moments); KIND 1: {y, y*y, x*y};
// KIND 2: three plain planes p0,p1,p2 (backward).
template <int KIND>
__device__ __forceinline__ void hpass(const float* __restrict__ p0, const float* __restrict__ p1,
                                      const float* __restrict__ p2, float* __restrict__ sH,
                                      const float* __restrict__ g) {
    using T = Tile<6>;
    const int row = threadIdx.x & 31, cg = threadIdx.x >> 5;
    if (row >= T::RH) return;
    const float* r0 = p0 + row * T::PITCH + cg * 8;
    float a[20];
#pragma unroll
    for (int k = 0; k < 20; ++k) a[k] = r0[k];
    float* o = sH + row * HP + cg * 8;
    constexpr int PL = T::RH * HP;
    // products are formed once per staged value (20 per row segment), not once per (output, tap)
    if (KIND == 0) {
        float aa[20];
#pragma unroll
        for (int k = 0; k < 20; ++k) aa[k] = a[k] * a[k];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float s0 = 0.f, s1 = 0.f;
#pragma unroll
            for (int k = 0; k < 13; ++k) {
                s0 = fmaf(g[k], a[j + k], s0);
                s1 = fmaf(g[k], aa[j + k], s1);
            }
            o[j] = s0;
            o[PL + j] = s1;
        }
    } else if (KIND == 1) {
        const float* r1 = p1 + row * T::PITCH + cg * 8;
        float b[20], bb[20], ab[20];
#pragma unroll
        for (int k = 0; k < 20; ++k) { b[k] = r1[k]; bb[k] = b[k] * b[k]; ab[k] = a[k] * b[k]; }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float s0 = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll
            for (int k = 0; k < 13; ++k) {
                s0 = fmaf(g[k], b[j + k], s0);
                s1 = fmaf(g[k], bb[j + k], s1);
                s2 = fmaf(g[k], ab[j + k], s2);
            }
            o[j] = s0;
            o[PL + j] = s1;
            o[2 * PL + j] = s2;
        }
    } else {
        const float* r1 = p1 + row * T::PITCH + cg * 8;
        const float* r2 = p2 + row * T::PITCH + cg * 8;
        float b[20], c[20];
#pragma unroll
        for (int k = 0; k < 20; ++k) { b[k] = r1[k]; c[k] = r2[k]; }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float s0 = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll
            for (int k = 0; k < 13; ++k) {
                s0 = fmaf(g[k], a[j + k], s0);
                s1 = fmaf(g[k], b[j + k], s1);
                s2 = fmaf(g[k], c[j + k], s2);
            }
            o[j] = s0;
            o[PL + j] = s1;
            o[2 * PL + j] = s2;
        }
    }
}

// Vertical pass: thread = (col, rg) owns output rows rg*5 .. rg*5+4 of column col.
template <int NMAP>
__device__ __forceinline__ void vpass(const float* __restrict__ sH, const float* __restrict__ g,
                                      float (&out)[NMAP][PXT]) {
    using T = Tile<6>;
    const int col = threadIdx.x & 63, rg = threadIdx.x >> 6;
    constexpr int PL = T::RH * HP;
#pragma unroll
    for (int m = 0; m < NMAP; ++m) {
        const float* p = sH + m * PL + (rg * PXT) * HP + col;
        float a[PXT + 12];
#pragma unroll
        for (int k = 0; k < PXT + 12; ++k) a[k] = p[k * HP];
#pragma unroll
        for (int j = 0; j < PXT; ++j) {
            float s = 0.f;
#pragma unroll
            for (int k = 0; k < 13; ++k) s = fmaf(g[k], a[j + k], s);
            out[m][j] = s;
        }
    }
}

// SSIM value + the three partial derivatives wrt (mu2, Eyy, Exy).  ssim.py:19-36, SURVEY A.4
__device__ __forceinline__ float ssim_point(float mu1, float exx, float mu2, float eyy, float exy,
                                            float& dmu2, float& deyy, float& dexy) {
    const float C1 = 0.0001f, C2 = 0.0009f;
    float mu1_sq = mu1 * mu1, mu2_sq = mu2 * mu2, mu12 = mu1 * mu2;
    float s1 = exx - mu1_sq, s2 = eyy - mu2_sq, s12 = exy - mu12;
    float A1 = 2.f * mu12 + C1, A2 = 2.f * s12 + C2;
    float B1 = mu1_sq + mu2_sq + C1, B2 = s1 + s2 + C2;
    float inv = __fdividef(1.f, B1 * B2);            // B1, B2 >= C1, C2 > 0: fast reciprocal (<= 2 ulp) is safe
    float S = (A1 * A2) * inv;
    dmu2 = 2.f * mu1 * (A2 - A1) * inv - S * 2.f * mu2 * (B2 - B1) * inv;
    deyy = -S * __fdividef(1.f, B2);
    dexy = 2.f * A1 * inv;
    return S;
}

}  // namespace ccb
