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

// d f1[b,c,y,x] = (1/C) sum_p g[b,p,y,x] f2[b,c,y+i-4,x+j-4]
// d f2[b,c,y,x] = (1/C) sum_p g[b,p,y-i+4,x-j+4] f1[b,c,y-i+4,x-j+4]
__global__ void __launch_bounds__(256) corr81_bwd_kernel(const float* __restrict__ f1, const float* __restrict__ f2,
                                                         const float* __restrict__ g, float* __restrict__ d1,
                                                         float* __restrict__ d2, int B, int C, int h, int w, int reversed) {
    const long long hw = (long long)h * w;
    long long i0 = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i0 >= (long long)B * C * hw) return;
    int x = (int)(i0 % w), y = (int)((i0 / w) % h);
    int c = (int)((i0 / hw) % C), b = (int)(i0 / (hw * C));
    const float* gb = g + (long long)b * 81 * hw;
    const float* p1 = f1 + ((long long)b * C + c) * hw;
    const float* p2 = f2 + ((long long)b * C + c) * hw;
    float a1 = 0.f, a2 = 0.f;
    for (int p = 0; p < 81; ++p) {
        int src = corr_src(p, reversed);
        int di = src / 9 - 4, dj = src % 9 - 4;
        int yy = y + di, xx = x + dj;
        if (d1 && yy >= 0 && yy < h && xx >= 0 && xx < w)
            a1 = fmaf(__ldg(gb + p * hw + (long long)y * w + x), __ldg(p2 + (long long)yy * w + xx), a1);
        int ys = y - di, xs = x - dj;
        if (d2 && ys >= 0 && ys < h && xs >= 0 && xs < w)
            a2 = fmaf(__ldg(gb + p * hw + (long long)ys * w + xs), __ldg(p1 + (long long)ys * w + xs), a2);
    }
    if (d1) d1[i0] = a1 / (float)C;
    if (d2) d2[i0] = a2 / (float)C;
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
                              int C, int h, int w, int reversed, ccb_stream_t stream) {
    CCB_REQUIRE(f1 && f2 && grad_out && (d_f1 || d_f2), CCB_ERR_ARG, "corr81_bwd: bad argument");
    long long n = (long long)B * C * h * w;
    CCB_LAUNCH(corr81_bwd_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, stream, f1, f2, grad_out, d_f1, d_f2, B, C, h, w,
               reversed);
    return check_launch("corr81_bwd");
}
