// geom.cuh - per-pixel projection / bilinear-sampling device functions shared by the fused
// photometric kernels and the stand-alone warp ops.
//
// Arithmetic recipe (one IEEE op per step, explicit rounding intrinsics, no fast-math) follows
// oracle/np_geometry.py, which is pinned bit-for-bit against the reference's CPU execution of
// inverse_warp.py:31-79 (pixel2cam / cam2pixel) - this is what makes the in-bounds / valid masks
// bit-exact.
#pragma once
#include "ccb_common.cuh"

namespace ccb {

struct Cam {
    float kinv[9];   // (scaled) K^-1, row-major
    float P[12];     // K_s [R|t], row-major 3x4
    float Ks[9];     // scaled K (backward: dT = Ks^T dP)
    float T[12];     // [R|t]
    float rot[3];    // rx, ry, rz  (euler) or qx,qy,qz (quat)
    float w1, h1;    // float(w-1), float(h-1)
};

// 3x3 @ 3xn with the (p0+p1)+p2 rounding of the reference's small bmm (see np_geometry.small_matmul)
__device__ __forceinline__ void mm3_small(const float* A, const float* Bm, int n, float* C) {
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < n; ++c)
            C[r * n + c] = __fadd_rn(__fadd_rn(__fmul_rn(A[r * 3 + 0], Bm[0 * n + c]),
                                               __fmul_rn(A[r * 3 + 1], Bm[1 * n + c])),
                                     __fmul_rn(A[r * 3 + 2], Bm[2 * n + c]));
}

// R = (X @ Y) @ Z, reference inverse_warp.py:82-119
__device__ __forceinline__ void euler_rot(float rx, float ry, float rz, float* R) {
    float sx = sinf(rx), cx = cosf(rx), sy = sinf(ry), cy = cosf(ry), sz = sinf(rz), cz = cosf(rz);
    float X[9] = {1.f, 0.f, 0.f, 0.f, cx, -sx, 0.f, sx, cx};
    float Y[9] = {cy, 0.f, sy, 0.f, 1.f, 0.f, -sy, 0.f, cy};
    float Z[9] = {cz, -sz, 0.f, sz, cz, 0.f, 0.f, 0.f, 1.f};
    float XY[9];
    mm3_small(X, Y, 3, XY);
    mm3_small(XY, Z, 3, R);
}

// reference inverse_warp.py:122-143
__device__ __forceinline__ void quat_rot(float qx, float qy, float qz, float* R) {
    float n = sqrtf(1.f + qx * qx + qy * qy + qz * qz);
    float w = 1.f / n, x = qx / n, y = qy / n, z = qz / n;
    float w2 = w * w, x2 = x * x, y2 = y * y, z2 = z * z;
    float wx = w * x, wy = w * y, wz = w * z, xy = x * y, xz = x * z, yz = y * z;
    R[0] = w2 + x2 - y2 - z2; R[1] = 2 * xy - 2 * wz;   R[2] = 2 * wy + 2 * xz;
    R[3] = 2 * wz + 2 * xy;   R[4] = w2 - x2 + y2 - z2; R[5] = 2 * yz - 2 * wx;
    R[6] = 2 * xz - 2 * wy;   R[7] = 2 * wx + 2 * yz;   R[8] = w2 - x2 - y2 + z2;
}

// Build the per-(batch, ref, level) camera.  downscale = H/h (loss_functions.py:87,91-92); pass 1 for
// pre-scaled intrinsics (stand-alone inverse_warp / pose2flow and the occlusion path).
__device__ __forceinline__ void make_cam(const float* pose6, const float* K, const float* Kinv,
                                         float downscale, int rot_mode, int w, int h, Cam& c) {
    for (int j = 0; j < 6; ++j) c.Ks[j] = (downscale == 1.f) ? K[j] : __fdiv_rn(K[j], downscale);
    for (int j = 6; j < 9; ++j) c.Ks[j] = K[j];
    for (int r = 0; r < 3; ++r) {
        c.kinv[r * 3 + 0] = (downscale == 1.f) ? Kinv[r * 3 + 0] : __fmul_rn(Kinv[r * 3 + 0], downscale);
        c.kinv[r * 3 + 1] = (downscale == 1.f) ? Kinv[r * 3 + 1] : __fmul_rn(Kinv[r * 3 + 1], downscale);
        c.kinv[r * 3 + 2] = Kinv[r * 3 + 2];
    }
    float R[9];
    c.rot[0] = pose6[3]; c.rot[1] = pose6[4]; c.rot[2] = pose6[5];
    if (rot_mode == CCB_ROT_EULER) euler_rot(pose6[3], pose6[4], pose6[5], R);
    else quat_rot(pose6[3], pose6[4], pose6[5], R);
    for (int r = 0; r < 3; ++r) {
        c.T[r * 4 + 0] = R[r * 3 + 0]; c.T[r * 4 + 1] = R[r * 3 + 1]; c.T[r * 4 + 2] = R[r * 3 + 2];
        c.T[r * 4 + 3] = pose6[r];
    }
    mm3_small(c.Ks, c.T, 4, c.P);
    c.w1 = (float)(w - 1);
    c.h1 = (float)(h - 1);
}

struct Proj {
    float c0, c1, c2;     // cam point = depth * ray
    float r0, r1, r2;     // ray = Kinv [x,y,1]
    float X, Y, Z;        // projected (Z clamped)
    float Xn, Yn;         // normalised coords (after the optional OOB->2 rewrite)
    bool xre, yre, zcl;   // x/y rewritten (no grad), z clamped (no grad through Z)
};

// reference inverse_warp.py:40-45 (pixel2cam) + :57-76 (cam2pixel)
__device__ __forceinline__ Proj project(const Cam& cm, float x, float y, float d, bool rewrite) {
    Proj p;
    p.r0 = __fadd_rn(cm.kinv[2], __fmaf_rn(cm.kinv[1], y, __fmul_rn(cm.kinv[0], x)));
    p.r1 = __fadd_rn(cm.kinv[5], __fmaf_rn(cm.kinv[4], y, __fmul_rn(cm.kinv[3], x)));
    p.r2 = __fadd_rn(cm.kinv[8], __fmaf_rn(cm.kinv[7], y, __fmul_rn(cm.kinv[6], x)));
    p.c0 = __fmul_rn(p.r0, d);
    p.c1 = __fmul_rn(p.r1, d);
    p.c2 = __fmul_rn(p.r2, d);
    p.X = __fadd_rn(__fmaf_rn(cm.P[2], p.c2, __fmaf_rn(cm.P[1], p.c1, __fmul_rn(cm.P[0], p.c0))), cm.P[3]);
    p.Y = __fadd_rn(__fmaf_rn(cm.P[6], p.c2, __fmaf_rn(cm.P[5], p.c1, __fmul_rn(cm.P[4], p.c0))), cm.P[7]);
    float Zr = __fadd_rn(__fmaf_rn(cm.P[10], p.c2, __fmaf_rn(cm.P[9], p.c1, __fmul_rn(cm.P[8], p.c0))), cm.P[11]);
    p.zcl = !(Zr >= 1e-3f);
    p.Z = fmaxf(Zr, 1e-3f);
    p.Xn = __fsub_rn(__fdiv_rn(__fmul_rn(2.f, __fdiv_rn(p.X, p.Z)), cm.w1), 1.f);
    p.Yn = __fsub_rn(__fdiv_rn(__fmul_rn(2.f, __fdiv_rn(p.Y, p.Z)), cm.h1), 1.f);
    p.xre = rewrite && ((p.Xn > 1.f) || (p.Xn < -1.f));
    p.yre = rewrite && ((p.Yn > 1.f) || (p.Yn < -1.f));
    if (p.xre) p.Xn = 2.f;
    if (p.yre) p.Yn = 2.f;
    return p;
}

// pose2flow tail, reference inverse_warp.py:217-218
__device__ __forceinline__ void coords_to_flow(const Cam& cm, float Xn, float Yn, float x, float y,
                                               float& u, float& v) {
    // Xn / 2.0 == Xn * 0.5 exactly (power of two): no IEEE-division sequence needed
    u = __fsub_rn(__fmul_rn(cm.w1, __fadd_rn(__fmul_rn(Xn, 0.5f), 0.5f)), x);
    v = __fsub_rn(__fmul_rn(cm.h1, __fadd_rn(__fmul_rn(Yn, 0.5f), 0.5f)), y);
}

// flow_warp grid, reference inverse_warp.py:181-188
__device__ __forceinline__ void flow_coords(float x, float y, float u, float v, float w1, float h1,
                                            float& Xn, float& Yn) {
    Xn = __fmul_rn(2.f, __fsub_rn(__fdiv_rn(__fadd_rn(x, u), w1), 0.5f));
    Yn = __fmul_rn(2.f, __fsub_rn(__fdiv_rn(__fadd_rn(y, v), h1), 0.5f));
}

// occlusion_masks, reference loss_functions.py:343-352 (bw, fw)
__device__ __forceinline__ float occ_mask(float ubw, float vbw, float ufw, float vfw) {
    float mag = __fadd_rn(__fadd_rn(__fmul_rn(ufw, ufw), __fmul_rn(vfw, vfw)),
                          __fadd_rn(__fmul_rn(ubw, ubw), __fmul_rn(vbw, vbw)));
    float s = __fadd_rn(__fadd_rn(ufw, ubw), __fadd_rn(vfw, vbw));
    float th = __fadd_rn(__fmul_rn(0.08f, mag), 1.0f);
    return (s > th) ? 1.f : 0.f;
}

// ---------------------------------------------------------------------------------------------
// Bilinear sampler, grid_sample(mode='bilinear', align_corners=False) as executed by ATen for the
// reference's un-annotated calls.
struct Samp {
    int x0, y0;
    float wx0, wx1, wy0, wy1;
    bool okx0, okx1, oky0, oky1;
    float gmx, gmy;      // d ix / d Xn, d iy / d Yn  (0 where border-clamped)
};

__device__ __forceinline__ Samp make_samp(float Xn, float Yn, int w, int h, int pad_mode) {
    Samp s;
    float ix = (Xn + 1.f) * (0.5f * (float)w) - 0.5f;
    float iy = (Yn + 1.f) * (0.5f * (float)h) - 0.5f;
    s.gmx = 0.5f * (float)w;
    s.gmy = 0.5f * (float)h;
    if (pad_mode == CCB_PAD_BORDER) {
        if (!(ix >= 0.f)) { ix = 0.f; s.gmx = 0.f; }
        else if (!(ix <= (float)(w - 1))) { ix = (float)(w - 1); s.gmx = 0.f; }
        if (!(iy >= 0.f)) { iy = 0.f; s.gmy = 0.f; }
        else if (!(iy <= (float)(h - 1))) { iy = (float)(h - 1); s.gmy = 0.f; }
    }
    // keep the float->int conversion defined for absurd coordinates
    ix = fminf(fmaxf(ix, -4.f), (float)w + 4.f);
    iy = fminf(fmaxf(iy, -4.f), (float)h + 4.f);
    float fx = floorf(ix), fy = floorf(iy);
    s.x0 = (int)fx;
    s.y0 = (int)fy;
    s.wx1 = ix - fx; s.wx0 = 1.f - s.wx1;
    s.wy1 = iy - fy; s.wy0 = 1.f - s.wy1;
    s.okx0 = (s.x0 >= 0) && (s.x0 < w);
    s.okx1 = (s.x0 + 1 >= 0) && (s.x0 + 1 < w);
    s.oky0 = (s.y0 >= 0) && (s.y0 < h);
    s.oky1 = (s.y0 + 1 >= 0) && (s.y0 + 1 < h);
    return s;
}

struct Corners { float v00, v01, v10, v11; };

__device__ __forceinline__ Corners fetch(const float* __restrict__ plane, const Samp& s, int w) {
    Corners c;
    const float* r0 = plane + (s.y0 * w + s.x0);
    const float* r1 = r0 + w;
    c.v00 = (s.oky0 && s.okx0) ? __ldg(r0) : 0.f;
    c.v01 = (s.oky0 && s.okx1) ? __ldg(r0 + 1) : 0.f;
    c.v10 = (s.oky1 && s.okx0) ? __ldg(r1) : 0.f;
    c.v11 = (s.oky1 && s.okx1) ? __ldg(r1 + 1) : 0.f;
    return c;
}

__device__ __forceinline__ float interp(const Corners& c, const Samp& s) {
    return c.v00 * (s.wy0 * s.wx0) + c.v01 * (s.wy0 * s.wx1) + c.v10 * (s.wy1 * s.wx0) + c.v11 * (s.wy1 * s.wx1);
}
// d out / d ix and d out / d iy (per unit upstream grad)
__device__ __forceinline__ float interp_dx(const Corners& c, const Samp& s) {
    return (c.v01 - c.v00) * s.wy0 + (c.v11 - c.v10) * s.wy1;
}
__device__ __forceinline__ float interp_dy(const Corners& c, const Samp& s) {
    return (c.v10 - c.v00) * s.wx0 + (c.v11 - c.v01) * s.wx1;
}

// robust L1 (x^2 + 0.01)^q and its derivative wrt x.  loss_functions.py:18-25
// (values feed means / gradients, not masks: rsqrt.approx (<= 2 ulp) instead of the IEEE sqrt/div sequences)
__device__ __forceinline__ float rl1(float x, float q) {
    float a = x * x + 0.01f;
    return (q == 0.5f) ? a * rsqrtf(a) : powf(a, q);
}
__device__ __forceinline__ float rl1_d(float x, float q) {
    float a = x * x + 0.01f;
    return (q == 0.5f) ? (x * rsqrtf(a)) : (2.f * q * x * powf(a, q - 1.f));
}

// ---------------------------------------------------------------------------------------------
// Backward of the projection chain for one pixel: upstream (gXn, gYn) -> d depth, and the 12 entries
// of dL/dP accumulated into acc[12].
__device__ __forceinline__ float project_bwd(const Cam& cm, const Proj& p, float gXn, float gYn,
                                             float* acc) {
    if (p.xre) gXn = 0.f;
    if (p.yre) gYn = 0.f;
    float iz = 1.f / p.Z;
    float ax = 2.f / cm.w1 * iz, ay = 2.f / cm.h1 * iz;
    float g0 = gXn * ax, g1 = gYn * ay;
    float g2 = p.zcl ? 0.f : -(g0 * p.X + g1 * p.Y) * iz;
    acc[0] += g0 * p.c0; acc[1] += g0 * p.c1; acc[2] += g0 * p.c2;  acc[3] += g0;
    acc[4] += g1 * p.c0; acc[5] += g1 * p.c1; acc[6] += g1 * p.c2;  acc[7] += g1;
    acc[8] += g2 * p.c0; acc[9] += g2 * p.c1; acc[10] += g2 * p.c2; acc[11] += g2;
    float d0 = cm.P[0] * p.r0 + cm.P[1] * p.r1 + cm.P[2] * p.r2;
    float d1 = cm.P[4] * p.r0 + cm.P[5] * p.r1 + cm.P[6] * p.r2;
    float d2 = cm.P[8] * p.r0 + cm.P[9] * p.r1 + cm.P[10] * p.r2;
    return g0 * d0 + g1 * d1 + g2 * d2;
}

// dL/dP (3x4) -> dL/dpose (6): dT = Ks^T dP; dt = dT[:,3]; d angles = <dT[:, :3], dR/d angle>.
__device__ __forceinline__ void pose_grad_from_dP(const Cam& cm, const float* dP, int rot_mode,
                                                  float* dpose /*+= 6*/) {
    float dT[12];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 4; ++c)
            dT[r * 4 + c] = cm.Ks[0 * 3 + r] * dP[0 * 4 + c] + cm.Ks[1 * 3 + r] * dP[1 * 4 + c] +
                            cm.Ks[2 * 3 + r] * dP[2 * 4 + c];
    dpose[0] += dT[3]; dpose[1] += dT[7]; dpose[2] += dT[11];
    float a = cm.rot[0], b = cm.rot[1], c = cm.rot[2];
    if (rot_mode == CCB_ROT_EULER) {
        float sx = sinf(a), cx = cosf(a), sy = sinf(b), cy = cosf(b), sz = sinf(c), cz = cosf(c);
        float X[9] = {1.f, 0.f, 0.f, 0.f, cx, -sx, 0.f, sx, cx};
        float Y[9] = {cy, 0.f, sy, 0.f, 1.f, 0.f, -sy, 0.f, cy};
        float Z[9] = {cz, -sz, 0.f, sz, cz, 0.f, 0.f, 0.f, 1.f};
        float dX[9] = {0.f, 0.f, 0.f, 0.f, -sx, -cx, 0.f, cx, -sx};
        float dY[9] = {-sy, 0.f, cy, 0.f, 0.f, 0.f, -cy, 0.f, -sy};
        float dZ[9] = {-sz, -cz, 0.f, cz, -sz, 0.f, 0.f, 0.f, 0.f};
        float t1[9], t2[9];
        float g[3];
        const float* As[3] = {dX, X, X};
        const float* Bs[3] = {Y, dY, Y};
        const float* Cs[3] = {Z, Z, dZ};
        for (int k = 0; k < 3; ++k) {
            for (int r = 0; r < 3; ++r)
                for (int cc = 0; cc < 3; ++cc)
                    t1[r * 3 + cc] = As[k][r * 3 + 0] * Bs[k][0 * 3 + cc] + As[k][r * 3 + 1] * Bs[k][1 * 3 + cc] +
                                     As[k][r * 3 + 2] * Bs[k][2 * 3 + cc];
            for (int r = 0; r < 3; ++r)
                for (int cc = 0; cc < 3; ++cc)
                    t2[r * 3 + cc] = t1[r * 3 + 0] * Cs[k][0 * 3 + cc] + t1[r * 3 + 1] * Cs[k][1 * 3 + cc] +
                                     t1[r * 3 + 2] * Cs[k][2 * 3 + cc];
            float s = 0.f;
            for (int r = 0; r < 3; ++r)
                for (int cc = 0; cc < 3; ++cc) s += dT[r * 4 + cc] * t2[r * 3 + cc];
            g[k] = s;
        }
        dpose[3] += g[0]; dpose[4] += g[1]; dpose[5] += g[2];
    } else {
        // quaternion: numerical-free analytic chain through nq = [1,q]/|[1,q]|
        float n2 = 1.f + a * a + b * b + c * c, n = sqrtf(n2);
        float q[4] = {1.f / n, a / n, b / n, c / n};
        float w = q[0], x = q[1], y = q[2], z = q[3];
        // dR/dq_k (k = w,x,y,z) contracted with dT
        float G[4];
        const float* d = dT;   // d[r*4+c]
        G[0] = 2.f * (w * d[0] - z * d[1] + y * d[2] + z * d[4] + w * d[5] - x * d[6] - y * d[8] + x * d[9] + w * d[10]);
        G[1] = 2.f * (x * d[0] + y * d[1] + z * d[2] + y * d[4] - x * d[5] - w * d[6] + z * d[8] + w * d[9] - x * d[10]);
        G[2] = 2.f * (-y * d[0] + x * d[1] + w * d[2] + x * d[4] + y * d[5] + z * d[6] - w * d[8] + z * d[9] - y * d[10]);
        G[3] = 2.f * (-z * d[0] - w * d[1] + x * d[2] + w * d[4] - z * d[5] + y * d[6] + x * d[8] + y * d[9] + z * d[10]);
        // q = v / n with v = [1,a,b,c]:  dq_k/dv_j = (delta_kj - q_k q_j) / n
        float dot = G[0] * q[0] + G[1] * q[1] + G[2] * q[2] + G[3] * q[3];
        dpose[3] += (G[1] - dot * q[1]) / n;
        dpose[4] += (G[2] - dot * q[2]) / n;
        dpose[5] += (G[3] - dot * q[3]) / n;
    }
}

}  // namespace ccb
