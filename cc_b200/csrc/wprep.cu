// wprep.cu - weight preparation shared by the tensor-core convolution kernels.
//   conv weights [Co][Ci][kh*kw] (Conv2d: mode 0 gathers x, output channel n = co;  DGRAD / ConvTranspose2d:
//   mode 1 gathers dy, output channel n = ci) -> wp[2][N][Kp]: the K order of the consuming kernel, each value
//   split into its tf32 rounding (copy 0) and the remainder (copy 1).
//   A block stages the weights of a few output channels in shared memory with coalesced reads (mode 0: one
//   contiguous run; mode 1: one run of nb*KK floats per gathered channel) and writes full rows of wp.
#include "ccb_common.cuh"

#ifndef CCB_CPU_SIM
namespace ccb {

__device__ __forceinline__ bool wprep_decode(const WPrepDesc& a, int k, int& slot, int& c) {
    if (a.layout == WPREP_TC) {                       // k = slot * cpad + c
        slot = k / a.p0;
        c = k - slot * a.p0;
        return slot < a.ntaps && c < a.Cc;
    }
    if (a.layout == WPREP_TMA) {                      // k = unit * cb + j, unit = slot * cblocks + channel block
        const int unit = k / a.p0, j = k - unit * a.p0;
        if (unit >= a.p2) return false;
        slot = unit / a.p1;
        c = (unit - slot * a.p1) * a.p0 + j;
        return slot < a.ntaps && c < a.Cc;
    }
    // WPREP_SLAB: k-stage -> (channel block, flattened (slot, channel) index inside the block)
    const int kt = k >> 5;
    const int cblock = min(kt / a.p2, a.p1 - 1);
    const int kl = (kt - cblock * a.p2) * 32 + (k & 31);
    const int nch = min(a.p0, a.Cc - cblock * a.p0);
    slot = kl / nch;
    c = cblock * a.p0 + (kl - slot * nch);
    return slot < a.ntaps;
}

__global__ void __launch_bounds__(256) wprep_staged_kernel(const WPrepDesc a, int nb) {
    extern __shared__ float sw[];
    const int n0 = blockIdx.x * nb;
    const int nn = min(nb, a.N - n0);
    const int row = a.Cc * a.KK;                          // staged floats per output channel
    if (a.mode == 0) {
        const float* src = a.w + (long long)n0 * row;     // [nn][Cc][KK] is one contiguous run
        for (int i = threadIdx.x; i < nn * row; i += 256) sw[i] = __ldg(src + i);
    } else {
        const int seg = nn * a.KK;                        // per gathered channel: [nn][KK] contiguous
        for (int i = threadIdx.x; i < a.Cc * seg; i += 256) {
            const int c = i / seg, r = i - c * seg;
            sw[i] = __ldg(a.w + ((long long)c * a.Ci + n0) * a.KK + r);
        }
    }
    __syncthreads();
    const long long plane = (long long)a.N * a.Kp;
    for (int i = threadIdx.x; i < nn * a.Kp; i += 256) {
        const int nl = i / a.Kp, k = i - nl * a.Kp;
        int slot, c;
        float v = 0.f;
        if (wprep_decode(a, k, slot, c)) {
            const int tap = a.tap_index[slot];
            v = (a.mode == 0) ? sw[(nl * a.Cc + c) * a.KK + tap] : sw[(c * nn + nl) * a.KK + tap];
        }
        uint32_t hb;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(v));
        const float h = __uint_as_float(hb);
        const long long o = (long long)(n0 + nl) * a.Kp + k;
        a.wp[o] = h;
        a.wp[plane + o] = v - h;
    }
}

int launch_wprep(const WPrepDesc& d, cudaStream_t st) {
    const long long row_bytes = (long long)d.Cc * d.KK * 4;
    CCB_REQUIRE(row_bytes <= 96 * 1024, CCB_ERR_UNSUPPORTED, "wprep: %d x %d weights per output channel do not fit shared memory", d.Cc, d.KK);
    int nb = (int)((64 * 1024) / (row_bytes > 0 ? row_bytes : 1));
    if (nb > 8) nb = 8;
    if (nb < 1) nb = 1;
    // keep enough blocks in flight
    while (nb > 1 && cdiv(d.N, nb) < 148) nb >>= 1;
    const int smem = (int)(nb * row_bytes);
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(wprep_staged_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
        attr_set = true;
    }
    CCB_LAUNCH(wprep_staged_kernel, dim3((unsigned)cdiv(d.N, nb)), dim3(256), smem, st, d, nb);
    return check_launch("wprep");
}

}  // namespace ccb
#endif
