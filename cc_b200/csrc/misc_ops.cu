// misc_ops.cu - the small non-GEMM layers of the four networks and the optimiser:
//   BatchNorm2d (train-mode batch statistics) fwd/bwd   models/DispResNet6.py:45-52 (13 layers)
//   bilinear x2 upsample (align_corners=False) fwd/bwd  models/DispResNet6.py:174,180,186; back2future.py:60
//   fused multi-tensor Adam over one flat buffer         train.py:307-310,568
#include "ccb_common.cuh"

namespace ccb {

// ---- BatchNorm (training): one CTA per channel ---------------------------------------------------
// stats[c] = {mean, invstd}; running stats updated with momentum (unbiased variance), like torch.
__global__ void __launch_bounds__(256) bn_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                     const float* __restrict__ beta, float* __restrict__ y,
                                                     float* __restrict__ stats, float* __restrict__ run_mean,
                                                     float* __restrict__ run_var, int B, int C, int plane, float eps,
                                                     float momentum) {
    __shared__ float s_red[32];
    __shared__ float s_mean, s_inv;
    const int c = blockIdx.x;
    const float n = (float)B * (float)plane;
    float v[1] = {0.f};
    for (int b = 0; b < B; ++b) {
        const float* p = x + ((long long)b * C + c) * plane;
        for (int i = threadIdx.x; i < plane; i += 256) v[0] += __ldg(p + i);
    }
    block_sum<1>(v, s_red);
    if (threadIdx.x == 0) s_mean = v[0] / n;
    __syncthreads();
    const float mean = s_mean;
    v[0] = 0.f;
    for (int b = 0; b < B; ++b) {
        const float* p = x + ((long long)b * C + c) * plane;
        for (int i = threadIdx.x; i < plane; i += 256) { float d = __ldg(p + i) - mean; v[0] += d * d; }
    }
    block_sum<1>(v, s_red);
    if (threadIdx.x == 0) {
        float var = v[0] / n;
        s_inv = 1.f / sqrtf(var + eps);
        stats[2 * c] = mean;
        stats[2 * c + 1] = s_inv;
        if (run_mean) {
            run_mean[c] = (1.f - momentum) * run_mean[c] + momentum * mean;
            float unb = (n > 1.f) ? v[0] / (n - 1.f) : var;
            run_var[c] = (1.f - momentum) * run_var[c] + momentum * unb;
        }
    }
    __syncthreads();
    const float inv = s_inv, g = __ldg(gamma + c), bt = __ldg(beta + c);
    for (int b = 0; b < B; ++b) {
        const float* p = x + ((long long)b * C + c) * plane;
        float* q = y + ((long long)b * C + c) * plane;
        for (int i = threadIdx.x; i < plane; i += 256) q[i] = (__ldg(p + i) - mean) * inv * g + bt;
    }
}

// eval mode: y = (x - running_mean) / sqrt(running_var + eps) * gamma + beta
__global__ void __launch_bounds__(256) bn_eval_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                      const float* __restrict__ beta, const float* __restrict__ rm,
                                                      const float* __restrict__ rv, float* __restrict__ y, int B, int C,
                                                      int plane, float eps) {
    const int c = blockIdx.x;
    const float mean = __ldg(rm + c), inv = 1.f / sqrtf(__ldg(rv + c) + eps), g = __ldg(gamma + c), bt = __ldg(beta + c);
    for (int b = 0; b < B; ++b) {
        const float* p = x + ((long long)b * C + c) * plane;
        float* q = y + ((long long)b * C + c) * plane;
        for (int i = threadIdx.x; i < plane; i += 256) q[i] = (__ldg(p + i) - mean) * inv * g + bt;
    }
}

__global__ void __launch_bounds__(256) bn_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                     const float* __restrict__ gamma, const float* __restrict__ stats,
                                                     float* __restrict__ dx, float* __restrict__ dgamma,
                                                     float* __restrict__ dbeta, int B, int C, int plane) {
    __shared__ float s_red[2 * 32];
    __shared__ float s_a, s_b;
    const int c = blockIdx.x;
    const float n = (float)B * (float)plane;
    const float mean = __ldg(stats + 2 * c), inv = __ldg(stats + 2 * c + 1), g = __ldg(gamma + c);
    float v[2] = {0.f, 0.f};
    for (int b = 0; b < B; ++b) {
        const float* p = x + ((long long)b * C + c) * plane;
        const float* q = dy + ((long long)b * C + c) * plane;
        for (int i = threadIdx.x; i < plane; i += 256) {
            float d = __ldg(q + i);
            v[0] += d;
            v[1] += d * (__ldg(p + i) - mean) * inv;
        }
    }
    block_sum<2>(v, s_red);
    if (threadIdx.x == 0) {
        dbeta[c] = v[0];
        dgamma[c] = v[1];
        s_a = v[0] / n;
        s_b = v[1] / n;
    }
    __syncthreads();
    const float ma = s_a, mb = s_b;
    for (int b = 0; b < B; ++b) {
        const float* p = x + ((long long)b * C + c) * plane;
        const float* q = dy + ((long long)b * C + c) * plane;
        float* r = dx + ((long long)b * C + c) * plane;
        for (int i = threadIdx.x; i < plane; i += 256) {
            float xh = (__ldg(p + i) - mean) * inv;
            r[i] = g * inv * (__ldg(q + i) - ma - xh * mb);
        }
    }
}

// ---- bilinear x2 upsample, align_corners=False (F.interpolate scale_factor=2) ---------------------
// out [planes, 2h, 2w]; source index = (o + 0.5)/2 - 0.5 clamped at 0 (ATen area_pixel_compute_source_index)
__device__ __forceinline__ void up2_src(int o, int n, int& i0, int& i1, float& l1) {
    float s = ((float)o + 0.5f) * 0.5f - 0.5f;
    if (s < 0.f) s = 0.f;
    i0 = (int)s;
    i1 = i0 + ((i0 < n - 1) ? 1 : 0);
    l1 = s - (float)i0;
}

__global__ void __launch_bounds__(256) upsample2x_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int planes,
                                                             int h, int w) {
    const int H = 2 * h, W = 2 * w;
    long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= (long long)planes * H * W) return;
    int ox = (int)(i % W), oy = (int)((i / W) % H), p = (int)(i / ((long long)W * H));
    int y0, y1, x0, x1;
    float ly, lx;
    up2_src(oy, h, y0, y1, ly);
    up2_src(ox, w, x0, x1, lx);
    const float* s = x + (long long)p * h * w;
    float hy = 1.f - ly, hx = 1.f - lx;
    y[i] = hy * (hx * __ldg(s + y0 * w + x0) + lx * __ldg(s + y0 * w + x1)) +
           ly * (hx * __ldg(s + y1 * w + x0) + lx * __ldg(s + y1 * w + x1));
}

// gather form of the transpose: each source pixel collects from the <= 4x4 outputs that read it
__global__ void __launch_bounds__(256) upsample2x_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dx, int planes,
                                                             int h, int w) {
    const int H = 2 * h, W = 2 * w;
    long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= (long long)planes * h * w) return;
    int sx = (int)(i % w), sy = (int)((i / w) % h), p = (int)(i / ((long long)w * h));
    const float* g = dy + (long long)p * H * W;
    float acc = 0.f;
    for (int oy = max(0, 2 * sy - 2); oy <= min(H - 1, 2 * sy + 2); ++oy) {
        int y0, y1;
        float ly;
        up2_src(oy, h, y0, y1, ly);
        float wy = ((y0 == sy) ? (1.f - ly) : 0.f) + ((y1 == sy) ? ly : 0.f);
        if (wy == 0.f) continue;
        for (int ox = max(0, 2 * sx - 2); ox <= min(W - 1, 2 * sx + 2); ++ox) {
            int x0, x1;
            float lx;
            up2_src(ox, w, x0, x1, lx);
            float wx = ((x0 == sx) ? (1.f - lx) : 0.f) + ((x1 == sx) ? lx : 0.f);
            if (wx != 0.f) acc += wy * wx * __ldg(g + (long long)oy * W + ox);
        }
    }
    dx[i] = acc;
}

// ---- Adam over a flat parameter / gradient buffer (torch.optim.Adam semantics, no weight decay) ----
// The step counter and bias corrections live in device memory (state[0..2] = step, 1-b1^t, sqrt(1-b2^t))
// so that the whole training step can be captured once in a CUDA graph and replayed.
__global__ void adam_prep_kernel(float* __restrict__ state, float b1, float b2) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        float t = state[0] + 1.f;
        state[0] = t;
        state[1] = (float)(1.0 - pow((double)b1, (double)t));
        state[2] = (float)sqrt(1.0 - pow((double)b2, (double)t));
    }
}

// grad_scale multiplies the gradient first (1/world_size after the NCCL sum).
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                   float* __restrict__ v, long long n, const float* __restrict__ state,
                                                   float lr, float b1, float b2, float eps, float grad_scale) {
    long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const float bc1 = __ldg(state + 1), bc2_sqrt = __ldg(state + 2);
    float gi = __ldg(g + i) * grad_scale;
    float mi = b1 * m[i] + (1.f - b1) * gi;
    float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = p[i] - (lr / bc1) * (mi / denom);
}

}  // namespace ccb

using namespace ccb;

extern "C" int ccb_bn_fwd(const float* x, const float* gamma, const float* beta, float* y, float* stats,
                          float* running_mean, float* running_var, int B, int C, int plane, float eps, float momentum,
                          int training, ccb_stream_t stream) {
    CCB_REQUIRE(x && gamma && beta && y, CCB_ERR_ARG, "bn_fwd: null pointer");
    if (training) {
        CCB_REQUIRE(stats != nullptr, CCB_ERR_ARG, "bn_fwd: stats null in training mode");
        CCB_LAUNCH(bn_fwd_kernel, dim3(C), dim3(256), 0, stream, x, gamma, beta, y, stats, running_mean, running_var, B, C,
                   plane, eps, momentum);
    } else {
        CCB_REQUIRE(running_mean && running_var, CCB_ERR_ARG, "bn_fwd: running stats null in eval mode");
        CCB_LAUNCH(bn_eval_kernel, dim3(C), dim3(256), 0, stream, x, gamma, beta, (const float*)running_mean,
                   (const float*)running_var, y, B, C, plane, eps);
    }
    return check_launch("bn_fwd");
}

extern "C" int ccb_bn_bwd(const float* x, const float* dy, const float* gamma, const float* stats, float* dx,
                          float* dgamma, float* dbeta, int B, int C, int plane, ccb_stream_t stream) {
    CCB_REQUIRE(x && dy && gamma && stats && dx && dgamma && dbeta, CCB_ERR_ARG, "bn_bwd: null pointer");
    CCB_LAUNCH(bn_bwd_kernel, dim3(C), dim3(256), 0, stream, x, dy, gamma, stats, dx, dgamma, dbeta, B, C, plane);
    return check_launch("bn_bwd");
}

extern "C" int ccb_upsample2x_fwd(const float* x, float* y, int planes, int h, int w, ccb_stream_t stream) {
    CCB_REQUIRE(x && y && planes >= 1 && h >= 1 && w >= 1, CCB_ERR_ARG, "upsample2x_fwd: bad argument");
    long long n = (long long)planes * 4 * h * w;
    CCB_LAUNCH(upsample2x_fwd_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, stream, x, y, planes, h, w);
    return check_launch("upsample2x_fwd");
}

extern "C" int ccb_upsample2x_bwd(const float* dy, float* dx, int planes, int h, int w, ccb_stream_t stream) {
    CCB_REQUIRE(dy && dx && planes >= 1 && h >= 1 && w >= 1, CCB_ERR_ARG, "upsample2x_bwd: bad argument");
    long long n = (long long)planes * h * w;
    CCB_LAUNCH(upsample2x_bwd_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, stream, dy, dx, planes, h, w);
    return check_launch("upsample2x_bwd");
}

extern "C" int ccb_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long long n,
                             float* state, float lr, float beta1, float beta2, float eps, float grad_scale,
                             ccb_stream_t stream) {
    CCB_REQUIRE(params && grads && exp_avg && exp_avg_sq && state && n >= 0, CCB_ERR_ARG, "adam_step: bad argument");
    CCB_LAUNCH(adam_prep_kernel, dim3(1), dim3(32), 0, stream, state, beta1, beta2);
    int rc = check_launch("adam_prep");
    if (rc || n == 0) return rc;
    CCB_LAUNCH(adam_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, stream, params, grads, exp_avg, exp_avg_sq, n,
               (const float*)state, lr, beta1, beta2, eps, grad_scale);
    return check_launch("adam_step");
}
