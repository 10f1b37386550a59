// b2f_ops.cu - the two non-conv operators of Back2Future (models/back2future.py):
//   correlate():  9x9 cost volume (third-party spatial_correlation_sample, kernel_size=1, patch_size=9)
//                 + the reference's channel permutation idx_fwd / idx_bwd, back2future.py:15-25,56-59,173-176
//   Model.warp(): feature warp = grid_sample(border, align_corners=False) of (grid + flow), :287-321
//                 (implemented by the flow_warp kernels in warp_ops.cu with the b2f normalisation)
#include "ccb_common.cuh"

namespace ccb {

// output channel p reads displacement (i, j):  idx_fwd[p] = (80 - p/9) - 9*(p%9);  idx_bwd[p] = idx_fwd[80-p]
__host__ __device__ __forceinline__ int corr_src(int p, int reversed) {
    int q = reversed ? 80 - p : p;
    return (80 - q / 9) - 9 * (q % 9);
}

constexpr int CT = 16;        // 16x16 pixel tile
constexpr int CH = CT + 8;    // + 4 px halo each side

// out[b,p,y,x] = (1/C) sum_c f1[b,c,y,x] * f2[b,c,y+i-4,x+j-4],  (i,j) = divmod(corr_src(p), 9)
__global__ void __launch_bounds__(CT * CT) corr81_fwd_kernel(const float* __restrict__ f1, const float* __restrict__ f2,
                                                             float* __restrict__ out, int B, int C, int h, int w, int reversed) {
    __shared__ float s2[CH][CH + 1];
    const int b = blockIdx.z, x0 = blockIdx.x * CT, y0 = blockIdx.y * CT;
    const int tx = threadIdx.x % CT, ty = threadIdx.x / CT;
    const int x = x0 + tx, y = y0 + ty;
    const long long hw = (long long)h * w;
    float acc[81];
#pragma unroll
    for (int k = 0; k < 81; ++k) acc[k] = 0.f;
    for (int c = 0; c < C; ++c) {
        const float* p2 = f2 + ((long long)b * C + c) * hw;
        __syncthreads();
        for (int idx = threadIdx.x; idx < CH * CH; idx += CT * CT) {
            int ry = idx / CH, rx = idx - ry * CH;
            int gy = y0 - 4 + ry, gx = x0 - 4 + rx;
            s2[ry][rx] = (gy >= 0 && gy < h && gx >= 0 && gx < w) ? __ldg(p2 + (long long)gy * w + gx) : 0.f;
        }
        __syncthreads();
        float a = (y < h && x < w) ? __ldg(f1 + ((long long)b * C + c) * hw + (long long)y * w + x) : 0.f;
#pragma unroll
        for (int i = 0; i < 9; ++i)
#pragma unroll
            for (int j = 0; j < 9; ++j) acc[i * 9 + j] = fmaf(a, s2[ty + i][tx + j], acc[i * 9 + j]);
    }
    if (y < h && x < w) {
        const float inv = 1.f / (float)C;
        for (int p = 0; p < 81; ++p) {
            int src = corr_src(p, reversed);
            out[((long long)b * 81 + p) * hw + (long long)y * w + x] = acc[src] / (float)C;
        }
        (void)inv;
    }
}

// Backward.  d f1[b,c,y,x] = (1/C) sum_p g[b,p,y,x] f2[b,c,y+di,x+dj]        (di,dj) = displacement of channel p
//            d f2[b,c,y,x] = (1/C) sum_p g[b,p,y-di,x-dj] f1[b,c,y-di,x-dj]
// Both are ONE tiled kernel: d[c](y,x) = (1/C) sum_k G[k](y,x) * F[c](y + k/9 - 4, x + k%9 - 4) over the 81 NATURAL
// displacements k, with the thread's 81 G values held in registers and F[c] staged per channel in shared memory (tile
// + 4 px halo), exactly like the forward.  d f1 uses G[k] = g[p(k)] read in place; d f2 uses the mirrored problem
// (corr(f1,f2)[(i,j)](y,x) == corr(f2,f1)[(8-i,8-j)](y+i-4,x+j-4)): G'[80-k](y,x) = g[p(k)](y-di,x-dj), materialised
// by corr81_mirror_kernel (81 shifted copies of g, a streaming pass) and F = f1.
__device__ __forceinline__ int corr_dst(int k, int reversed) {      // inverse of corr_src: natural displacement k -> output channel p
    // corr_src(q) = (80 - q/9) - 9*(q%9)  =>  k = 80 - a - 9 r with a = q/9, r = q%9  =>  r = (80-k)/9, a = (80-k)%9
    const int m = 80 - k;
    const int q = (m % 9) * 9 + m / 9;
    return reversed ? 80 - q : q;
}

__global__ void __launch_bounds__(256) corr81_mirror_kernel(const float* __restrict__ g, float* __restrict__ gm, int B, int h, int w,
                                                            int reversed) {
    const long long hw = (long long)h * w, n = (long long)B * 81 * hw;
    for (long long i0 = (long long)blockIdx.x * 256 + threadIdx.x; i0 < n; i0 += (long long)gridDim.x * 256) {
        const int x = (int)(i0 % w), y = (int)((i0 / w) % h);
        const int kk = (int)((i0 / hw) % 81), b = (int)(i0 / (hw * 81));      // kk = 80 - k: mirrored natural index
        const int k = 80 - kk;
        const int di = k / 9 - 4, dj = k % 9 - 4;
        const int ys = y - di, xs = x - dj;
        float v = 0.f;
        if (ys >= 0 && ys < h && xs >= 0 && xs < w) v = __ldg(g + ((long long)b * 81 + corr_dst(k, reversed)) * hw + (long long)ys * w + xs);
        gm[i0] = v;
    }
}

// natural != 0: G is already in natural displacement order (the mirrored buffer); else G[k] = g[corr_dst(k)]
__global__ void __launch_bounds__(CT * CT) corr81_dgrad_kernel(const float* __restrict__ G, const float* __restrict__ F,
                                                               float* __restrict__ d, int B, int C, int h, int w, int reversed,
                                                               int natural) {
    __shared__ float sf[CH][CH + 1];
    const int b = blockIdx.z, x0 = blockIdx.x * CT, y0 = blockIdx.y * CT;
    const int tx = threadIdx.x % CT, ty = threadIdx.x / CT;
    const int x = x0 + tx, y = y0 + ty;
    const long long hw = (long long)h * w;
    const bool in = (y < h) && (x < w);
    float gk[81];
#pragma unroll
    for (int k = 0; k < 81; ++k) {
        const int p = natural ? k : corr_dst(k, reversed);
        gk[k] = in ? __ldg(G + ((long long)b * 81 + p) * hw + (long long)y * w + x) : 0.f;
    }
    const float invC = 1.f / (float)C;
    for (int c = 0; c < C; ++c) {
        const float* pf = F + ((long long)b * C + c) * hw;
        __syncthreads();
        for (int idx = threadIdx.x; idx < CH * CH; idx += CT * CT) {
            int ry = idx / CH, rx = idx - ry * CH;
            int gy = y0 - 4 + ry, gx = x0 - 4 + rx;
            sf[ry][rx] = (gy >= 0 && gy < h && gx >= 0 && gx < w) ? __ldg(pf + (long long)gy * w + gx) : 0.f;
        }
        __syncthreads();
        float a = 0.f;
#pragma unroll
        for (int i = 0; i < 9; ++i)
#pragma unroll
            for (int j = 0; j < 9; ++j) a = fmaf(gk[i * 9 + j], sf[ty + i][tx + j], a);
        if (in) d[((long long)b * C + c) * hw + (long long)y * w + x] = a * invC;
    }
}

}  // namespace ccb

using namespace ccb;

extern "C" int ccb_corr81_fwd(const float* f1, const float* f2, float* out, int B, int C, int h, int w, int reversed,
                              ccb_stream_t stream) {
    CCB_REQUIRE(f1 && f2 && out && B >= 1 && C >= 1 && h >= 1 && w >= 1, CCB_ERR_ARG, "corr81_fwd: bad argument");
    CCB_LAUNCH(corr81_fwd_kernel, dim3(cdiv(w, CT), cdiv(h, CT), B), dim3(CT * CT), 0, stream, f1, f2, out, B, C, h, w, reversed);
    return check_launch("corr81_fwd");
}

extern "C" int ccb_corr81_bwd(const float* f1, const float* f2, const float* grad_out, float* d_f1, float* d_f2, int B,
                              int C, int h, int w, int reversed, float* work, ccb_stream_t stream) {
    CCB_REQUIRE(f1 && f2 && grad_out && (d_f1 || d_f2), CCB_ERR_ARG, "corr81_bwd: bad argument");
    CCB_REQUIRE(d_f2 == nullptr || work != nullptr, CCB_ERR_ARG, "corr81_bwd: d_f2 needs a workspace of B*81*h*w floats");
    const dim3 grid(cdiv(w, CT), cdiv(h, CT), B);
    if (d_f1) CCB_LAUNCH(corr81_dgrad_kernel, grid, dim3(CT * CT), 0, stream, grad_out, f2, d_f1, B, C, h, w, reversed, 0);
    if (d_f2) {
        const long long n = (long long)B * 81 * h * w;
        long long nb = (n + 255) / 256;
        if (nb > 148 * 16) nb = 148 * 16;
        CCB_LAUNCH(corr81_mirror_kernel, dim3((unsigned)nb), dim3(256), 0, stream, grad_out, work, B, h, w, reversed);
        CCB_LAUNCH(corr81_dgrad_kernel, grid, dim3(CT * CT), 0, stream, (const float*)work, f1, d_f2, B, C, h, w, reversed, 1);
    }
    return check_launch("corr81_bwd");
}
