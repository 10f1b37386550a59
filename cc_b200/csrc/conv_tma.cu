// conv_tma.cu - TMA-fed tcgen05 implicit-GEMM convolution (stride-1 gathers: every stride-1 FPROP, every
// DGRAD parity class, ConvTranspose2d forward).
//
//   one CTA = 128 output pixels (4 image rows x 32 columns) x N <= 128 channels
//   per k-stage (K = 32):
//     A  (activations, MN-major): 4 x (32/cb) TMA boxes  (32 px, 1 row, cb channels) straight from the NCHW
//        tensor at the tap-shifted coordinate - the box IS the im2col tile: zero padding, image borders and
//        channel tails are the TMA's out-of-bounds zero fill, SWIZZLE_128B puts it in the UMMA layout;
//     B  (weights, K-major):     1 (+1) TMA box (32 k, N rows) from the tf32-split weight buffer;
//     3xTF32: eight warps compute lo = x - trunc_tf32(x) smem -> smem (elementwise, layout agnostic); the MMA
//        reads the raw fp32 tile as the hi operand (kind::tf32 ignores the 13 low mantissa bits);
//     one elected thread issues hi*hi (+ lo*hi + hi*lo into a second TMEM accumulator) and commits.
//   No register staging of operands, no per-element index math, producers = ONE thread: the pipeline depth is
//   bounded by shared memory only (3 stages of 64 KB in 3xTF32, 6 stages of 32 KB in single-TF32 mode).
#include "tc_common.cuh"

#ifndef CCB_CPU_SIM

namespace ccb {

constexpr int TM_M = 128;
constexpr int TM_THREADS = 320;            // direct kernel: warp 0 TMA, warp 1 MMA + TMEM, warps 2..9 lo-pass + epilogue
constexpr int SL_THREADS = 576;            // slab kernels: warp 0 TMA, warp 1 MMA + TMEM, two groups of 8 cutter warps
constexpr int TM_MAX_SLOTS = 64;           // tap slots (taps padded to a multiple of 32/cb)
constexpr int TM_A_BYTES = 16384;          // one A operand copy of one stage: 128 px x 32 k x 4 B

struct TmaConvArgs {
    int B, Cin, Hin, Win;
    int Ntot, Hout, Wout, Hc, Wc;
    int out_stride, out_oy, out_ox;
    int ntaps, cb, cblocks, units, ktiles;   // cb: channels per unit (8/16/32); units = ntaps * cblocks
    int tiles_x, tiles_y;
    int splits, kt_per_split;
    long long out_numel;
    float* partial;
    const float* bias;
    const float* res;
    float* out;
    int act;
    float slope;
    int nstages, b_tile_bytes, nalloc, soft, dbg;
    signed char off_y[TM_MAX_SLOTS], off_x[TM_MAX_SLOTS];
};

// A wait that cannot complete is a protocol bug: fail the launch instead of hanging the box.  With the bring-up
// switch (ccb_debug_tc_swap_strides bit 4) the first timeout is recorded in g_tma_status and every wait of the
// launch falls through, so that the probe can report WHO waited on WHAT without losing the CUDA context.
__device__ unsigned int g_tma_status[4];
__device__ __noinline__ void tm_wait_failed(int soft, int role, int it) {
    if (!soft) asm volatile("trap;");
    if (atomicCAS(&g_tma_status[0], 0u, (unsigned)role) == 0u) {
        g_tma_status[1] = (unsigned)it;
        g_tma_status[2] = blockIdx.x;
        g_tma_status[3] = blockIdx.z;
    }
}
__device__ __forceinline__ void tm_mbar_wait(uint64_t* bar, uint32_t parity, int soft, int role, int it) {
    uint32_t spins = 0;
    while (!tm_mbar_try(bar, parity)) {
        ++spins;
        if (soft && (spins & 1023u) == 0 && *(volatile unsigned int*)&g_tma_status[0] != 0u) return;
        if (spins > (soft ? (1u << 20) : (1u << 26))) { tm_wait_failed(soft, role, it); return; }
    }
}
template <bool THREE>
__global__ void __launch_bounds__(TM_THREADS, 1)
conv_tma_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const TmaConvArgs a) {
    CCB_PDL_TRIGGER();
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_addr(smem_raw) & 1023u)) & 1023u);
    const int NST = a.nstages;
    const int stage_bytes = (THREE ? 2 : 1) * (TM_A_BYTES + a.b_tile_bytes);   // [A raw | A lo] [B hi | B lo]
    uint64_t* tma_full = (uint64_t*)(smem + NST * stage_bytes);
    uint64_t* split_full = tma_full + 8;
    uint64_t* empty_bar = split_full + 8;
    uint64_t* accum_bar = empty_bar + 8;
    uint32_t* tmem_slot = (uint32_t*)(accum_bar + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // tile -> (batch, 4-row band, 32-column block)
    int t = blockIdx.x;
    const int per_b = a.tiles_x * a.tiles_y;
    const int b = t / per_b;
    t -= b * per_b;
    const int ty = t / a.tiles_x, tx = t - ty * a.tiles_x;
    const int x0 = tx * 32, y0 = ty * 4;
    const int n0 = blockIdx.y * 128;
    const int ntile = min(128, a.Ntot - n0);
    const int umma_n = (ntile + 15) & ~15;
    const int kt_beg = blockIdx.z * a.kt_per_split;
    const int ktiles = max(0, min(a.ktiles, kt_beg + a.kt_per_split) - kt_beg);
    const int ups = 32 / a.cb;                       // units (tap, channel-block) per stage

    if (tid == 0) {
        for (int s = 0; s < NST; ++s) {
            tm_mbar_init(&tma_full[s], 1);
            tm_mbar_init(&split_full[s], 256);
            tm_mbar_init(&empty_bar[s], 1);
        }
        tm_mbar_init(accum_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(tmem_slot)), "r"(256u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    CCB_PDL_SYNC();                                               // everything above touched no global data

    if (warp == 0) {
        // ===================== TMA producer (one thread) =====================
        if (lane == 0) {
            const uint32_t tx_bytes = (uint32_t)(TM_A_BYTES + (THREE ? 2 : 1) * a.b_tile_bytes);
            for (int it = 0; it < ktiles; ++it) {
                const int s = it % NST;
                if (it >= NST) tm_mbar_wait(&empty_bar[s], ((it / NST) - 1) & 1, a.soft, 1, it);
                unsigned char* st = smem + s * stage_bytes;
                unsigned char* a_raw = st;
                unsigned char* b_hi = st + (THREE ? 2 : 1) * TM_A_BYTES;
                tm_mbar_expect_tx(&tma_full[s], tx_bytes);
                const int kt = kt_beg + it;
                for (int u = 0; u < ups; ++u) {
                    int unit = kt * ups + u;
                    int slot = 0, c0 = 0;
                    if (unit < a.units) { slot = unit / a.cblocks; c0 = (unit - slot * a.cblocks) * a.cb; }
                    // (units beyond the real K range multiply zero weights: any in-range coordinate will do)
                    const int X = x0 + a.off_x[slot];
#pragma unroll
                    for (int mb = 0; mb < 4; ++mb) {
                        const int Y = y0 + mb + a.off_y[slot];
                        tma_load_4d(a_raw + mb * 4096 + u * (a.cb * 128), &map_a, &tma_full[s], X, Y, c0, b);
                    }
                }
                tma_load_2d(b_hi, &map_b, &tma_full[s], kt * 32, n0);
                if (THREE) tma_load_2d(b_hi + a.b_tile_bytes, &map_b, &tma_full[s], kt * 32, a.Ntot + n0);
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (one thread) =====================
        // D fp32, A/B tf32, A MN-major (pixels contiguous), B K-major, N = umma_n, M = 128
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | ((uint32_t)(umma_n >> 3) << 17) |
                               ((uint32_t)(TM_M >> 4) << 24);
        for (int it = 0; it < ktiles; ++it) {
            const int s = it % NST;
            tm_mbar_wait(&tma_full[s], (it / NST) & 1, a.soft, 2, it);
            if (THREE) tm_mbar_wait(&split_full[s], (it / NST) & 1, a.soft, 3, it);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (lane == 0) {
                const uint32_t st = smem_addr(smem + s * stage_bytes);
                const uint32_t a_raw = st, a_lo = st + TM_A_BYTES;
                const uint32_t b_hi = st + (THREE ? 2 : 1) * TM_A_BYTES, b_lo = b_hi + a.b_tile_bytes;
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    // A (MN-major): 32-pixel blocks 4096 B apart (LBO), 4-channel swizzle atoms 512 B apart (SBO)
                    const uint64_t ah = tm_desc(a_raw + ks * 1024, 4096, 512, 1);
                    // B (K-major): 8-row groups 1024 B apart (SBO); a k-step advances 32 B inside the 128 B row
                    const uint64_t bh = tm_desc(b_hi + ks * 32, 16, 1024, 2);
                    tm_umma_tf32(tmem_base, ah, bh, idesc, (it > 0 || ks > 0) ? 1u : 0u);
                    if (THREE) {
                        const uint64_t al = tm_desc(a_lo + ks * 1024, 4096, 512, 1);
                        const uint64_t bl = tm_desc(b_lo + ks * 32, 16, 1024, 2);
                        tm_umma_tf32(tmem_base + 128u, al, bh, idesc, (it > 0 || ks > 0) ? 1u : 0u);
                        tm_umma_tf32(tmem_base + 128u, ah, bl, idesc, 1u);
                    }
                }
                tm_commit(&empty_bar[s]);
                if (it == ktiles - 1) tm_commit(accum_bar);
            }
            __syncwarp();
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    } else {
        // ===================== lo-pass (3xTF32) + epilogue: 8 warps =====================
        const int wt = tid - 64;                       // 0 .. 255
        if (THREE) {
            for (int it = 0; it < ktiles; ++it) {
                const int s = it % NST;
                tm_mbar_wait(&tma_full[s], (it / NST) & 1, a.soft, 4, it);
                float4* raw = (float4*)(smem + s * stage_bytes);
                float4* lo = raw + TM_A_BYTES / 16;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int i = wt + j * 256;            // 1024 float4 per stage
                    lo[i] = tf32_rest4(raw[i]);
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                tm_mbar_arrive(&split_full[s]);
            }
        }
        if (ktiles > 0) tm_mbar_wait(accum_bar, 0, a.soft, 5, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int q4 = warp & 3, colhalf = (warp - 2) >> 2;
        const int oy = y0 + q4, ox = x0 + lane;        // TMEM lane m = 32 * row-in-tile + column-in-tile
        const bool evalid = (oy < a.Hc) && (ox < a.Wc);
        const long long HWout = (long long)a.Hout * a.Wout;
        const long long obase = (long long)b * a.Ntot * HWout + (long long)(oy * a.out_stride + a.out_oy) * a.Wout +
                                (ox * a.out_stride + a.out_ox);
        for (int cg = colhalf; cg * 16 < umma_n; cg += 2) {
            float v[16];
            if (ktiles > 0) {
                tm_ld16(tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(cg * 16), v);
                if (THREE) {
                    float vl[16];
                    tm_ld16(tmem_base + ((uint32_t)(q4 * 32) << 16) + 128u + (uint32_t)(cg * 16), vl);
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] += vl[j];
                }
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = 0.f;
            }
            if (evalid) {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int n = n0 + cg * 16 + j;
                    if (cg * 16 + j < ntile) {
                        float o = v[j];
                        const long long off = obase + (long long)n * HWout;
                        if (a.splits > 1) {
                            a.partial[(long long)blockIdx.z * a.out_numel + off] = o;
                        } else {
                            if (a.bias) o += __ldg(a.bias + n);
                            if (a.res) o += __ldg(a.res + off);
                            a.out[off] = tm_act(o, a.act, a.slope);
                        }
                    }
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
    }
}

void launch_splitk_reduce(const float* work, float* out, const float* bias, const float* res, long long numel, int splits,
                          int plane, int C, int act, float slope, cudaStream_t st);   // conv_ffma.cu

// conv_nhwc.cu: channels-last slab kernel
bool nhwc_applies(const int* off_y, const int* off_x, int ntaps, int in_stride, int Cc, int N, int Hc, int Wc, int three, bool part_of_set);
bool nhwc_prefers_thin();
void nhwc_debug_plan(const int* off_y, const int* off_x, int ntaps, int in_stride, int Cc, int N, int B, int Hc, int Wc, bool part_of_set,
                     int* out16);
long long nhwc_copy_floats(int B, int Cc, int Hin, int Win);
long long nhwc_wp_floats(const int* off_y, const int* off_x, int ntaps, int Cc, int N);
int nhwc_transpose(const float* x, float* xh, int B, int Cc, int Hin, int Win, cudaStream_t st);
int launch_nhwc(const float* xh, int B, int Cc, int Hin, int Win, const float* w, int mode, int N, int KK, int Ci, const int* off_y,
                const int* off_x, const int* tap_index, int ntaps, int Hc, int Wc, int Hout, int Wout, int out_stride, int out_oy,
                int out_ox, const float* bias, const float* res, float* out, int act, float slope, int three, float* work,
                long long wp_floats, int splits, float* partial, long long out_numel, cudaStream_t st);

static int g_tma_enabled = 1, g_tma_soft = 0;
void tma_set_enabled(int v) { g_tma_enabled = v & 1; g_tma_soft = (v >> 1) & 1; }

// Which problems the TMA path takes: stride-1 gathers on tensors whose row pitch is 16-byte aligned.
// One launch: gathered tensor x [B, Cc, Hin, Win] -> output grid (Hc x Wc) written with stride / offset.
static int launch_tma(const float* x, int B, int Cc, int Hin, int Win, const float* w, int mode, int N, int KK, int Ci,
                      const int* off_y, const int* off_x, const int* tap_index, int ntaps, int Hc, int Wc, int Hout, int Wout,
                      int out_stride, int out_oy, int out_ox, const float* bias, const float* res, float* out, int act,
                      float slope, int three, float* work, long long work_floats, int splits, float* partial,
                      long long out_numel, cudaStream_t st) {
    EncodeTiledFn enc = get_encode();
    CCB_REQUIRE(enc != nullptr, CCB_ERR_UNSUPPORTED, "conv_tma: cuTensorMapEncodeTiled unavailable");
    TmaConvArgs a;
    memset(&a, 0, sizeof(a));
    a.B = B; a.Cin = Cc; a.Hin = Hin; a.Win = Win; a.Ntot = N; a.Hout = Hout; a.Wout = Wout; a.Hc = Hc; a.Wc = Wc;
    a.out_stride = out_stride; a.out_oy = out_oy; a.out_ox = out_ox;
    a.cb = Cc >= 32 ? 32 : (Cc > 8 ? 16 : 8);
    a.cblocks = cdiv(Cc, a.cb);
    a.ntaps = ntaps;
    a.units = ntaps * a.cblocks;
    const int Kp = cdiv(a.units * a.cb, 32) * 32 > 0 ? cdiv(a.units * a.cb, 32) * 32 : 32;
    a.ktiles = Kp / 32;
    a.tiles_x = cdiv(Wc, 32); a.tiles_y = cdiv(Hc, 4);
    a.splits = splits; a.kt_per_split = cdiv(a.ktiles, splits);
    a.out_numel = out_numel; a.partial = partial;
    a.bias = bias; a.res = res; a.out = out; a.act = act; a.slope = slope; a.soft = g_tma_soft;
    { const char* e = getenv("CCB_TMA_DBG"); a.dbg = e ? atoi(e) : 0; }
    WPrepDesc p;
    memset(&p, 0, sizeof(p));
    CCB_REQUIRE(ntaps <= TM_MAX_SLOTS, CCB_ERR_ARG, "conv_tma: too many taps");
    for (int t = 0; t < TM_MAX_SLOTS; ++t) {
        a.off_y[t] = (signed char)(t < ntaps ? off_y[t] : 0);
        a.off_x[t] = (signed char)(t < ntaps ? off_x[t] : 0);
        p.tap_index[t] = (signed char)(t < ntaps ? tap_index[t] : 0);
    }
    const long long wp_floats = 2ll * N * Kp;
    CCB_REQUIRE(wp_floats <= work_floats, CCB_ERR_ARG, "conv_tma: workspace too small");
    p.w = w; p.wp = work; p.N = N; p.Cc = Cc; p.KK = KK; p.Ci = Ci; p.mode = mode; p.Kp = Kp; p.ntaps = ntaps;
    p.layout = WPREP_TMA; p.p0 = a.cb; p.p1 = a.cblocks; p.p2 = a.units;
    const float* wpp = nullptr;
    int rc = wprep_get(p, st, &wpp);
    if (rc) return rc;
    const int ntile_max = N < 128 ? N : 128;
    int nalloc = 16;
    while (nalloc < ntile_max) nalloc <<= 1;
    a.nalloc = nalloc;
    a.b_tile_bytes = nalloc * 128;
    const int stage_bytes = (three ? 2 : 1) * (TM_A_BYTES + a.b_tile_bytes);
    a.nstages = (200 * 1024) / stage_bytes;
    if (a.nstages > 8) a.nstages = 8;
    if (a.nstages < 2) a.nstages = 2;
    const int smem = a.nstages * stage_bytes + 2048;
    // tensor maps.  A: x as (W, H, C, B), box (32, 1, cb, 1), SWIZZLE_128B_ATOM_32B, OOB -> 0.  B: wp as (Kp, 2N), box (32, nalloc).
    alignas(64) CUtensorMap map_a, map_b;
    {
        cuuint64_t dims[4] = {(cuuint64_t)Win, (cuuint64_t)Hin, (cuuint64_t)Cc, (cuuint64_t)B};
        cuuint64_t strides[3] = {(cuuint64_t)Win * 4, (cuuint64_t)Win * Hin * 4, (cuuint64_t)Win * Hin * Cc * 4};
        cuuint32_t box[4] = {32, 1, (cuuint32_t)a.cb, 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = enc(&map_a, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)x, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        CCB_REQUIRE(r == CUDA_SUCCESS, CCB_ERR_LAUNCH, "conv_tma: cuTensorMapEncodeTiled(A) failed (%d)", (int)r);
    }
    {
        cuuint64_t dims[2] = {(cuuint64_t)Kp, (cuuint64_t)(2 * N)};
        cuuint64_t strides[1] = {(cuuint64_t)Kp * 4};
        cuuint32_t box[2] = {32, (cuuint32_t)nalloc};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = enc(&map_b, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)wpp, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        CCB_REQUIRE(r == CUDA_SUCCESS, CCB_ERR_LAUNCH, "conv_tma: cuTensorMapEncodeTiled(B) failed (%d)", (int)r);
    }
    dim3 grid(B * a.tiles_x * a.tiles_y, cdiv(N, 128), splits);
    auto kfn = three ? conv_tma_kernel<true> : conv_tma_kernel<false>;
    cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    CCB_LAUNCH(kfn, grid, dim3(TM_THREADS), smem, st, map_a, map_b, a);
    return check_launch("conv_tma");
}


// ---------------------------------------------------------------------------------------------------------------
// Slab kernel: FPROP (stride 1 and 2) and every DGRAD parity class with arbitrary tap offsets.
//   TMA can only start a box on a 16-byte boundary of the innermost (W) dimension, so a tap shifted by one pixel
//   cannot be fetched as a box of its own.  Instead ONE aligned box per channel block - the input patch under the
//   whole 4 x 32 output tile, every tap included, zero padding = OOB fill - lands in shared memory once (the
//   "slab", [channel][row][column], no swizzle), and eight warps cut the per-tap operand tiles out of it:
//   element (pixel, k = (tap, channel)) -> MN-major SWIZZLE_128B_BASE32B tile, written twice: the fp32 value (the
//   tensor core reads its tf32 truncation) and the truncation remainder (3xTF32).
//   K is the flattened (tap, channel) index of the block, padded to 32 once per block: no per-tap padding.
//   MMA per 8-deep k step:  D[:, 0:2n)  += A_hi * [B_hi | B_lo]   (hi*hi and hi*lo share one read of A_hi)
//                           D[:, 256:+n) += A_lo * B_hi
// ---------------------------------------------------------------------------------------------------------------
struct SlabArgs {
    int B, Cin, Hin, Win;
    int Ntot, Hout, Wout, Hc, Wc;
    int out_stride, out_oy, out_ox, in_stride;
    int ntaps;
    int cs, cblocks, kt_full, ktiles;       // channels per slab, slabs, k-stages of a full slab, k-stages in total
    int SW, SH, ox_lo, oy_lo, slab_bytes, slab_tx, nslab;
    int tiles_x, tiles_y, splits, kt_per_split;
    long long out_numel;
    float* partial;
    const float* bias;
    const float* res;
    float* out;
    int act;
    float slope;
    int nstages, nbox, b_tile_bytes, soft, dbg, mt;
    signed char off_y[TM_MAX_SLOTS], off_x[TM_MAX_SLOTS];
};

template <bool THREE>
__global__ void __launch_bounds__(SL_THREADS, 1)
conv_slab_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_b, const SlabArgs a) {
    CCB_PDL_WAIT();
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_addr(smem_raw) & 1023u)) & 1023u);
    const int NST = a.nstages;
    const int stage_bytes = (THREE ? 2 : 1) * TM_A_BYTES + a.b_tile_bytes;     // [A hi | A lo] [B hi ; B lo]
    unsigned char* slab0 = smem + NST * stage_bytes;
    uint64_t* bars = (uint64_t*)(slab0 + a.nslab * a.slab_bytes);
    uint64_t* a_full = bars;             // [8]  the 8 cutter warps of one group
    uint64_t* b_full = bars + 8;         // [8]  TMA weights
    uint64_t* empty_bar = bars + 16;     // [8]  tcgen05.commit
    uint64_t* slab_full = bars + 24;     // [2]  TMA slab
    uint64_t* slab_empty = bars + 26;    // [2]  16 cutter warps
    uint64_t* accum_bar = bars + 28;
    uint32_t* tmem_slot = (uint32_t*)(bars + 29);
    int* toff = (int*)(bars + 32);       // [64] slab offset of each tap
    long long* trace = (long long*)(bars + 64);   // bring-up only (a.dbg & 32): [7][64] time stamps of CTA 0
    const bool tracing = (a.dbg & 32) && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
    const long long t_start = clock64();

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    int t = blockIdx.x;
    const int per_b = a.tiles_x * a.tiles_y;
    const int b = t / per_b;
    t -= b * per_b;
    const int ty = t / a.tiles_x, tx = t - ty * a.tiles_x;
    const int x0 = tx * 32, y0 = ty * 4 * a.mt;                   // a CTA owns mt vertically stacked 4 x 32 tiles (one slab)
    const int n0 = blockIdx.y * 128;
    const int ntile = min(128, a.Ntot - n0);
    const int kt_beg = blockIdx.z * a.kt_per_split;
    const int ktiles = max(0, min(a.ktiles, kt_beg + a.kt_per_split) - kt_beg);
    const int s_in = a.in_stride;
    const int xs = (x0 * s_in + a.ox_lo) & ~3;                    // aligned slab origin (may be negative: OOB fill)
    const int dx0 = x0 * s_in + a.ox_lo - xs;
    const int ys = y0 * s_in + a.oy_lo;
    const int first_block = min(kt_beg / a.kt_full, a.cblocks - 1);

    if (tid >= 64 && tid < 64 + TM_MAX_SLOTS) {
        const int tp = tid - 64;
        toff[tp] = (tp < a.ntaps) ? (a.off_y[tp] - a.oy_lo) * a.SW + (a.off_x[tp] - a.ox_lo) : 0;
    }
    // ---- producer state (thread 0): slabs and weight tiles are requested in k-stage order ----
    const int last_block = (ktiles > 0) ? min((kt_beg + ktiles - 1) / a.kt_full, a.cblocks - 1) : first_block - 1;
    const int nblocks = last_block - first_block + 1;             // slabs this CTA goes through
    int p_it = 0, p_s = 0, p_ph = 0;                               // next k-stage to request, its ring slot and phase
    int p_slab = 0, p_sb = 0, p_sph = 0;                           // next slab to request, its buffer and phase
    int p_blk = 0, p_left = 0;                                     // block of stage p_it, its remaining stages
    const uint32_t b_tx = (uint32_t)((THREE ? 2 : 1) * a.nbox * 128);
    int p_k = 0;                                                   // k-stage of p_it inside its tile
    const int total_stages = a.mt * ktiles;
    auto produce = [&](int limit, bool nowait) {                   // request stages p_it .. limit-1 (nowait: stop at the first wait)
        for (; p_it < limit; ++p_it) {
            if (p_left == 0 && a.mt > 1) p_left = total_stages;   // stacked tiles share the one slab
            if (p_left == 0) {                                     // first stage of a block
                if (p_it > 0) ++p_blk;
                const int cb = first_block + p_blk;
                const int full = (cb < a.cblocks - 1) ? a.kt_full : a.ktiles - (a.cblocks - 1) * a.kt_full;
                p_left = (p_it == 0) ? full - (kt_beg - cb * a.kt_full) : full;
            }
            // the slab of this block and (double buffered) of the next one
            const int want = min(nblocks, p_blk + a.nslab);
            while (p_slab < want) {
                if (p_slab >= a.nslab) {
                    if (nowait) return;
                    tm_mbar_wait(&slab_empty[p_sb], p_sph ^ 1, a.soft, 6, p_it);
                }
                tm_mbar_expect_tx(&slab_full[p_sb], (uint32_t)a.slab_tx);
                tma_load_4d(slab0 + p_sb * a.slab_bytes, &map_x, &slab_full[p_sb], xs, ys, (first_block + p_slab) * a.cs, b);
                ++p_slab;
                if (++p_sb == a.nslab) { p_sb = 0; p_sph ^= 1; }
            }
            if (p_it >= NST) tm_mbar_wait(&empty_bar[p_s], p_ph ^ 1, a.soft, 1, p_it);
            unsigned char* bt = smem + p_s * stage_bytes + (THREE ? 2 : 1) * TM_A_BYTES;
            tm_mbar_expect_tx(&b_full[p_s], b_tx);
            tma_load_2d(bt, &map_b, &b_full[p_s], (kt_beg + p_k) * 32, n0);
            if (THREE) tma_load_2d(bt + a.nbox * 128, &map_b, &b_full[p_s], (kt_beg + p_k) * 32, a.Ntot + n0);
            if (++p_k == ktiles) p_k = 0;
            if (tracing && p_it < 64) trace[0 * 64 + p_it] = clock64() - t_start;
            --p_left;
            if (++p_s == NST) { p_s = 0; p_ph ^= 1; }
        }
    };
    if (tid == 0) {
        tm_prefetch_map(&map_x);
        tm_prefetch_map(&map_b);
        for (int s = 0; s < NST; ++s) {
            tm_mbar_init(&a_full[s], 8);
            tm_mbar_init(&b_full[s], 1);
            tm_mbar_init(&empty_bar[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            tm_mbar_init(&slab_full[s], 1);
            tm_mbar_init(&slab_empty[s], 16);
        }
        tm_mbar_init(accum_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        produce(min(total_stages, NST), true);                           // the first ring of loads needs no consumer: start it before the CTA sync
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    if (tracing && tid == 0) trace[7 * 64] = clock64() - t_start;

    if (warp == 0) {
        // ===================== TMA producer (one thread): the rest of the slabs + weight tiles =====================
        if (lane == 0) produce(total_stages, false);
    } else if (warp == 1) {
        // ===================== MMA issuer: the whole warp runs the loop, lane 0's predicate issues =====================
        const uint32_t idesc_base = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | ((uint32_t)(TM_M >> 4) << 24);
        const uint32_t idesc_n1 = idesc_base | ((uint32_t)(a.nbox >> 3) << 17);
        const uint32_t idesc_n2 = idesc_base | ((uint32_t)((2 * a.nbox) >> 3) << 17);
        const uint32_t leader = (lane == 0) ? 1u : 0u;
        const uint32_t smem16 = smem_addr(smem) >> 4, stage16 = (uint32_t)stage_bytes >> 4;
        int s = 0;
        uint32_t ph = 0;
        for (int m = 0; m < a.mt; ++m) {
        const uint32_t d0 = tmem_base + (uint32_t)(m * (THREE ? 3 : 1) * a.nbox), d1 = d0 + (uint32_t)(2 * a.nbox);
        for (int it = 0; it < ktiles; ++it) {
            tm_mbar_wait(&b_full[s], ph, a.soft, 2, it);
            tm_mbar_wait(&a_full[s], ph, a.soft, 3, it);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (tracing && lane == 0 && it < 64) trace[1 * 64 + it] = clock64() - t_start;
            const uint32_t a_hi16 = smem16 + (uint32_t)s * stage16, a_lo16 = a_hi16 + (TM_A_BYTES >> 4);
            const uint32_t b16 = a_hi16 + (((THREE ? 2 : 1) * TM_A_BYTES) >> 4);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                const uint32_t acc = (it > 0 || ks > 0) ? 1u : 0u;
                tm_umma_tf32_p(d0, (a_hi16 + ks * 64) | DESC_A_MN_LO, DESC_A_MN_HI, (b16 + ks * 2) | DESC_K_LO, DESC_K_HI,
                               THREE ? idesc_n2 : idesc_n1, acc, leader);
                if (THREE)
                    tm_umma_tf32_p(d1, (a_lo16 + ks * 64) | DESC_A_MN_LO, DESC_A_MN_HI, (b16 + ks * 2) | DESC_K_LO, DESC_K_HI,
                                   idesc_n1, acc, leader);
            }
            tm_commit_p(&empty_bar[s], leader);
            if (it == ktiles - 1 && m == a.mt - 1) tm_commit_p(accum_bar, leader);
            if (tracing && lane == 0 && it < 64) trace[2 * 64 + it] = clock64() - t_start;
            if (++s == NST) { s = 0; ph ^= 1; }
        }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    } else {
        // ===================== cutters (2 groups of 8 warps): slab -> operand tiles; then the epilogue =====================
        // The groups take alternate k-stages, so two stages are being cut at any time (one stage is a chain of
        // shared-memory round trips, not a throughput problem).  Inside a group warp w8 owns k rows 4*w8 .. 4*w8+3;
        // a quarter warp writes one 128-byte operand row: lane -> (tile row mb, group of 4 pixels).
        const int grp = (warp - 2) >> 3, w8 = (warp - 2) & 7;
        const int pxg = lane & 7, mb = lane >> 3;
        const int plane = a.SH * a.SW;
        const int thr_off = mb * s_in * a.SW + pxg * 4 * s_in + dx0;           // this thread's corner inside a slab plane
        uint32_t dst[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int k = w8 * 4 + i;
            dst[i] = (uint32_t)(mb * 4096 + k * 128 + (((pxg >> 1) ^ (k & 3)) << 5) + (pxg & 1) * 16);
        }
        int g = 0;                                                              // stages gone by (all tiles of this CTA)
        bool need_slab = true;                                                  // this warp has not looked at the current slab yet
        int sb = 0, s = 0;                                                      // slab buffer / ring slot of the stage ...
        uint32_t sph = 0, ph = 0;                                               // ... and their phases
        for (int m = 0; m < a.mt; ++m) {
        const int thr_m = thr_off + m * 4 * s_in * a.SW;
        int cblock = first_block;
        int kl_stage = kt_beg - cblock * a.kt_full;                             // k-stage inside the block
        int blk_stages = (cblock < a.cblocks - 1) ? a.kt_full : a.ktiles - (a.cblocks - 1) * a.kt_full;
        int nch = min(a.cs, a.Cin - cblock * a.cs);
        int q32 = 32 / nch, r32 = 32 - q32 * nch;
        int tap0 = (kl_stage * 32 + w8 * 4) / nch, c0 = (kl_stage * 32 + w8 * 4) - tap0 * nch;
        for (int it = 0; it < ktiles; ++it, ++g) {
            if ((g & 1) == grp) {
                const float* slab = (const float*)(slab0 + sb * a.slab_bytes);
                if (need_slab) {
                    tm_mbar_wait(&slab_full[sb], sph, a.soft, 4, it);
                    need_slab = false;
                }
                if (g >= NST) tm_mbar_wait(&empty_bar[s], ph ^ 1, a.soft, 7, it);
                unsigned char* st = smem + s * stage_bytes;
                if (tracing && w8 == 0 && lane == 0 && it < 64) trace[3 * 64 + it] = clock64() - t_start;
                int sidx[4];
                bool kval[4];
                {
                    int tap = tap0, c = c0;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        kval[i] = tap < a.ntaps;
                        sidx[i] = thr_m + (kval[i] ? c * plane + toff[tap] : 0);
                        if (++c == nch) { c = 0; ++tap; }
                    }
                }
                float4 v[4];
                if (s_in == 1) {
                    float4 va[4], vb[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) va[i] = *(const float4*)(slab + (sidx[i] & ~3));
#pragma unroll
                    for (int i = 0; i < 4; ++i) vb[i] = *(const float4*)(slab + (sidx[i] & ~3) + 4);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int al = sidx[i] & 3;
                        v[i] = al == 0 ? va[i]
                             : al == 1 ? make_float4(va[i].y, va[i].z, va[i].w, vb[i].x)
                             : al == 2 ? make_float4(va[i].z, va[i].w, vb[i].x, vb[i].y)
                                       : make_float4(va[i].w, vb[i].x, vb[i].y, vb[i].z);
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        v[i].x = slab[sidx[i]]; v[i].y = slab[sidx[i] + s_in];
                        v[i].z = slab[sidx[i] + 2 * s_in]; v[i].w = slab[sidx[i] + 3 * s_in];
                    }
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    if (!kval[i]) v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                    *(float4*)(st + dst[i]) = v[i];
                    if (THREE) *(float4*)(st + TM_A_BYTES + dst[i]) = tf32_rest4(v[i]);
                }
                if (tracing && w8 == 0 && lane == 0 && it < 64) trace[4 * 64 + it] = clock64() - t_start;
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) tm_mbar_arrive(&a_full[s]);
                if (tracing && w8 == 0 && lane == 0 && it < 64) trace[5 * 64 + it] = clock64() - t_start;
            }
            ++kl_stage;
            const bool block_done = (kl_stage == blk_stages);
            if ((block_done || it == ktiles - 1) && lane == 0) tm_mbar_arrive(&slab_empty[sb]);
            if (++s == NST) { s = 0; ph ^= 1; }
            if (block_done && it + 1 < ktiles) {
                ++cblock; kl_stage = 0; need_slab = true;
                if (++sb == a.nslab) { sb = 0; sph ^= 1; }
                blk_stages = (cblock < a.cblocks - 1) ? a.kt_full : a.ktiles - (a.cblocks - 1) * a.kt_full;
                nch = min(a.cs, a.Cin - cblock * a.cs);
                q32 = 32 / nch; r32 = 32 - q32 * nch;
                tap0 = (w8 * 4) / nch; c0 = (w8 * 4) - tap0 * nch;
            } else {
                c0 += r32; tap0 += q32;
                if (c0 >= nch) { c0 -= nch; ++tap0; }
            }
        }
        }
        if (ktiles > 0) tm_mbar_wait(accum_bar, 0, a.soft, 5, 0);
        if (tracing && warp == 2 && lane == 0) trace[6 * 64 + 0] = clock64() - t_start;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int q4 = warp & 3, colq = (warp - 2) >> 2;        // 4 warps per TMEM lane quarter split the columns
        const long long HWout = (long long)a.Hout * a.Wout;
        const int ox = x0 + lane;
        for (int m = 0; m < a.mt; ++m) {
            const int oy = y0 + m * 4 + q4;
            const bool evalid = (oy < a.Hc) && (ox < a.Wc);
            const long long obase = (long long)b * a.Ntot * HWout + (long long)(oy * a.out_stride + a.out_oy) * a.Wout +
                                    (ox * a.out_stride + a.out_ox);
            const uint32_t trow = tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(m * (THREE ? 3 : 1) * a.nbox);
            for (int cg = colq; cg * 16 < ntile; cg += 4) {
                float v[16];
                if (ktiles > 0) {
                    tm_ld16(trow + (uint32_t)(cg * 16), v);
                    if (THREE) {
                        float v2[16];
                        tm_ld16(trow + (uint32_t)(a.nbox + cg * 16), v2);
#pragma unroll
                        for (int j = 0; j < 16; ++j) v[j] += v2[j];
                        tm_ld16(trow + (uint32_t)(2 * a.nbox + cg * 16), v2);
#pragma unroll
                        for (int j = 0; j < 16; ++j) v[j] += v2[j];
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = 0.f;
                }
                if (evalid) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int n = n0 + cg * 16 + j;
                        if (cg * 16 + j < ntile) {
                            float o = v[j];
                            const long long off = obase + (long long)n * HWout;
                            if (a.splits > 1) {
                                a.partial[(long long)blockIdx.z * a.out_numel + off] = o;
                            } else {
                                if (a.bias) o += __ldg(a.bias + n);
                                if (a.res) o += __ldg(a.res + off);
                                a.out[off] = tm_act(o, a.act, a.slope);
                            }
                        }
                    }
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (tracing && tid == 0) {
        printf("slab trace: prologue_done %lld epilogue_start %lld end %lld (cycles)\n", trace[7 * 64], trace[6 * 64], (long long)(clock64() - t_start));
        for (int it = 0; it < ktiles && it < 64; ++it)
            printf("it %2d  tma %6lld  mma_ready %6lld  mma_issued %6lld  cut_start %6lld  cut_stored %6lld  cut_arrived %6lld\n", it,
                   trace[0 * 64 + it], trace[1 * 64 + it], trace[2 * 64 + it], trace[3 * 64 + it], trace[4 * 64 + it], trace[5 * 64 + it]);
    }
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

struct SlabPlan {
    int cs, cblocks, kt_full, ktiles, SW, SH, ox_lo, oy_lo, slab_bytes, nslab, nstages, nbox, b_tile_bytes, smem, mt;
    bool ok;
};
// tap offsets (in gathered-tensor pixels), gather stride, channels, output channels -> tiling of the K dimension
// `stack` > 1 asks for that many vertically stacked tiles per CTA if everything still fits (one channel block only)
static SlabPlan slab_plan(const int* off_y, const int* off_x, int ntaps, int in_stride, int Cc, int N, int three, int stack = 1) {
    SlabPlan p;
    memset(&p, 0, sizeof(p));
    int oxl = 0, oxh = 0, oyl = 0, oyh = 0;
    for (int t = 0; t < ntaps; ++t) {
        if (t == 0 || off_x[t] < oxl) oxl = off_x[t];
        if (t == 0 || off_x[t] > oxh) oxh = off_x[t];
        if (t == 0 || off_y[t] < oyl) oyl = off_y[t];
        if (t == 0 || off_y[t] > oyh) oyh = off_y[t];
    }
    p.ox_lo = oxl; p.oy_lo = oyl;
    p.SW = (31 * in_stride + (oxh - oxl) + 1 + 3 + 3) & ~3;
    p.SH = 3 * in_stride + (oyh - oyl) + 1;
    const int ntile = N < 128 ? N : 128;
    p.nbox = (ntile + 15) & ~15;
    p.b_tile_bytes = (three ? 2 : 1) * p.nbox * 128;
    const int stage = (three ? 2 : 1) * TM_A_BYTES + p.b_tile_bytes;
    const int per_ch = p.SH * p.SW * 4;
    const int nt = ntaps > 0 ? ntaps : 1;
    int g = nt, r = 32;                    // channel granularity that makes cs * ntaps a multiple of 32
    while (r) { int q = g % r; g = r; r = q; }
    const int align = 32 / g;
    const int total = 220 * 1024;
    for (int nst = 3; nst >= 2 && !p.ok; --nst) {
        const int budget = total - nst * stage;
        for (int cb = 1; cb <= Cc && !p.ok; ++cb) {
            int cs = cdiv(Cc, cb);
            if (cb > 1) cs = cdiv(cs, align) * align;
            if (cs > 256 || cs < 1) continue;
            const int cblocks = cdiv(Cc, cs);
            const int nslab = cblocks > 1 ? 2 : 1;
            const int sbytes = cdiv(cs * per_ch, 128) * 128;
            if (nslab * sbytes > budget) continue;
            p.cs = cs; p.cblocks = cblocks; p.nslab = nslab; p.slab_bytes = sbytes;
            p.nstages = (total - nslab * sbytes) / stage;
            if (p.nstages > 6) p.nstages = 6;
            p.ok = true;
        }
    }
    if (!p.ok) return p;
    p.mt = 1;
    if (p.cblocks == 1) {
        for (int mt = stack; mt > 1; --mt) {                     // taller slab, mt accumulators in TMEM
            const int sh = (4 * mt - 1) * in_stride + (oyh - oyl) + 1;
            const int sbytes = cdiv(p.cs * sh * p.SW * 4, 128) * 128;
            if (mt * (three ? 3 : 1) * p.nbox > 512 || sbytes + 3 * stage > total || sh > 256) continue;
            p.mt = mt; p.SH = sh; p.slab_bytes = sbytes;
            p.nstages = (total - sbytes) / stage;
            if (p.nstages > 6) p.nstages = 6;
            break;
        }
    }
    p.kt_full = cdiv(p.cs * nt, 32);
    const int tail = Cc - (p.cblocks - 1) * p.cs;
    p.ktiles = (p.cblocks - 1) * p.kt_full + cdiv(tail * nt, 32);
    p.smem = p.nstages * stage + p.nslab * p.slab_bytes + 1024 + 1024 + 4096;   // + barriers / tap table / bring-up trace
    return p;
}

static int launch_slab(const float* x, int B, int Cc, int Hin, int Win, const float* w, int mode, int N, int KK, int Ci,
                       const int* off_y, const int* off_x, const int* tap_index, int ntaps, int in_stride, int Hc, int Wc, int Hout,
                       int Wout, int out_stride, int out_oy, int out_ox, const float* bias, const float* res, float* out, int act,
                       float slope, int three, float* work, long long wp_floats, int splits, float* partial, long long out_numel,
                       cudaStream_t st) {
    EncodeTiledFn enc = get_encode();
    CCB_REQUIRE(enc != nullptr, CCB_ERR_UNSUPPORTED, "conv_slab: cuTensorMapEncodeTiled unavailable");
    CCB_REQUIRE(ntaps <= TM_MAX_SLOTS, CCB_ERR_ARG, "conv_slab: too many taps");
    // thin layers have thousands of short tiles: stack up to 4 of them per CTA while >= 4 CTAs per SM remain
    int stack = 1;
    if (splits == 1) {
        const long long tiles = (long long)B * cdiv(Wc, 32) * cdiv(Hc, 4) * cdiv(N, 128);
        while (stack < 4 && tiles / (stack * 2) >= 4 * 148 && cdiv(Hc, 4) >= stack * 2) stack *= 2;
    }
    const SlabPlan p = slab_plan(off_y, off_x, ntaps, in_stride, Cc, N, three, stack);
    CCB_REQUIRE(p.ok, CCB_ERR_UNSUPPORTED, "conv_slab: no tiling fits shared memory");
    SlabArgs a;
    memset(&a, 0, sizeof(a));
    a.B = B; a.Cin = Cc; a.Hin = Hin; a.Win = Win; a.Ntot = N; a.Hout = Hout; a.Wout = Wout; a.Hc = Hc; a.Wc = Wc;
    a.out_stride = out_stride; a.out_oy = out_oy; a.out_ox = out_ox; a.in_stride = in_stride;
    a.ntaps = ntaps; a.cs = p.cs; a.cblocks = p.cblocks; a.kt_full = p.kt_full; a.ktiles = p.ktiles;
    a.SW = p.SW; a.SH = p.SH; a.ox_lo = p.ox_lo; a.oy_lo = p.oy_lo; a.slab_bytes = p.slab_bytes; a.slab_tx = p.cs * p.SH * p.SW * 4; a.nslab = p.nslab;
    a.mt = p.mt;
    a.tiles_x = cdiv(Wc, 32); a.tiles_y = cdiv(Hc, 4 * p.mt);
    a.splits = splits; a.kt_per_split = cdiv(a.ktiles, splits);
    a.out_numel = out_numel; a.partial = partial;
    a.bias = bias; a.res = res; a.out = out; a.act = act; a.slope = slope; a.soft = g_tma_soft;
    a.nstages = p.nstages; a.nbox = p.nbox; a.b_tile_bytes = p.b_tile_bytes;
    { const char* e = getenv("CCB_TMA_DBG"); a.dbg = e ? atoi(e) : 0; }
    WPrepDesc pa;
    memset(&pa, 0, sizeof(pa));
    for (int t = 0; t < TM_MAX_SLOTS; ++t) {
        a.off_y[t] = (signed char)(t < ntaps ? off_y[t] : 0);
        a.off_x[t] = (signed char)(t < ntaps ? off_x[t] : 0);
        pa.tap_index[t] = (signed char)(t < ntaps ? tap_index[t] : 0);
    }
    const int Kp = a.ktiles * 32;
    CCB_REQUIRE(2ll * N * Kp <= wp_floats, CCB_ERR_ARG, "conv_slab: workspace too small");
    pa.w = w; pa.wp = work; pa.N = N; pa.Cc = Cc; pa.KK = KK; pa.Ci = Ci; pa.mode = mode; pa.Kp = Kp; pa.ntaps = ntaps;
    pa.layout = WPREP_SLAB; pa.p0 = p.cs; pa.p1 = p.cblocks; pa.p2 = p.kt_full;
    const float* wpp = nullptr;
    int rc = wprep_get(pa, st, &wpp);
    if (rc) return rc;
    alignas(64) CUtensorMap map_x, map_b;
    {
        cuuint64_t dims[4] = {(cuuint64_t)Win, (cuuint64_t)Hin, (cuuint64_t)Cc, (cuuint64_t)B};
        cuuint64_t strides[3] = {(cuuint64_t)Win * 4, (cuuint64_t)Win * Hin * 4, (cuuint64_t)Win * Hin * Cc * 4};
        cuuint32_t box[4] = {(cuuint32_t)p.SW, (cuuint32_t)p.SH, (cuuint32_t)p.cs, 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = enc(&map_x, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)x, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        CCB_REQUIRE(r == CUDA_SUCCESS, CCB_ERR_LAUNCH, "conv_slab: cuTensorMapEncodeTiled(x) failed (%d)", (int)r);
    }
    {
        cuuint64_t dims[2] = {(cuuint64_t)Kp, (cuuint64_t)(2 * N)};
        cuuint64_t strides[1] = {(cuuint64_t)Kp * 4};
        cuuint32_t box[2] = {32, (cuuint32_t)p.nbox};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = enc(&map_b, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)wpp, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        CCB_REQUIRE(r == CUDA_SUCCESS, CCB_ERR_LAUNCH, "conv_slab: cuTensorMapEncodeTiled(w) failed (%d)", (int)r);
    }
    dim3 grid(B * a.tiles_x * a.tiles_y, cdiv(N, 128), splits);
    auto kfn = three ? conv_slab_kernel<true> : conv_slab_kernel<false>;
    cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    CCB_LAUNCH(kfn, grid, dim3(SL_THREADS), p.smem, st, map_x, map_b, a);
    return check_launch("conv_slab");
}


// ---------------------------------------------------------------------------------------------------------------
// Direct kernel (CUDA cores, plain fp32 FMA) for THIN layers: few output channels (<= 32 per block) and a short
// K = channels x taps.  The tensor core's tf32 MMA has K = 8 per instruction and needs 3 passes for fp32 accuracy:
// with N = 16 an MMA carries 16 K MACs and the kernel is issue bound; one fp32 FMA per MAC on the CUDA cores is
// faster there (and exact).  Same interface as the slab kernel: an aligned TMA slab of the input patch (zero
// padding = OOB fill), a tap list, output written with stride / offset (DGRAD parity classes).
//   thread -> 4 pixels (px = qx + 8 e: lanes read consecutive floats) x 8 output channels; block = 32 columns x TH rows
//   x NG channel groups (TH * NG = 32); channels arrive in chunks of CC (one slab per chunk).
// ---------------------------------------------------------------------------------------------------------------
constexpr int DC_THREADS = 256;
struct DirectArgs {
    int B, Cin, Hin, Win;
    int Ntot, Hout, Wout, Hc, Wc;
    int out_stride, out_oy, out_ox, in_stride;
    int ntaps, NG, TH, CC, nchunks;
    int SW, SH, ox_lo, oy_lo, slab_bytes, slab_tx;
    int tiles_x, tiles_y;
    const float* wq;               // [n-block][chunk][tap][CC][NG*8]
    const float* bias;
    const float* res;
    float* out;
    int act;
    float slope;
    signed char off_y[TM_MAX_SLOTS], off_x[TM_MAX_SLOTS];
};

// (d0, d1) += a * (b0, b1) as ONE packed instruction
__device__ __forceinline__ void ffma2(float& d0, float& d1, float a, float b0, float b1) {
    asm("{\n\t"
        ".reg .b64 ra, rb, rd;\n\t"
        "mov.b64 ra, {%2, %2};\n\t"
        "mov.b64 rb, {%3, %4};\n\t"
        "mov.b64 rd, {%0, %1};\n\t"
        "fma.rn.f32x2 rd, ra, rb, rd;\n\t"
        "mov.b64 {%0, %1}, rd;\n\t"
        "}" : "+f"(d0), "+f"(d1) : "f"(a), "f"(b0), "f"(b1));
}

__global__ void __launch_bounds__(DC_THREADS, 2)
conv_direct_kernel(const __grid_constant__ CUtensorMap map_x, const DirectArgs a) {
    CCB_PDL_WAIT();
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    float* slab = (float*)smem_raw;
    const int n8 = a.NG * 8;
    float* wsm = (float*)(smem_raw + a.slab_bytes);                     // [tap][CC][n8]
    int* toff = (int*)(wsm + a.ntaps * a.CC * n8);
    uint64_t* bar = (uint64_t*)(toff + TM_MAX_SLOTS);

    const int tid = threadIdx.x;
    const int qx = tid & 7, rest = tid >> 3;
    const int ng = rest % a.NG, ry = rest / a.NG;
    int t = blockIdx.x;
    const int per_b = a.tiles_x * a.tiles_y;
    const int b = t / per_b;
    t -= b * per_b;
    const int ty = t / a.tiles_x, tx = t - ty * a.tiles_x;
    const int x0 = tx * 32, y0 = ty * a.TH;
    const int nblk = blockIdx.y, n0 = nblk * n8;
    const int s_in = a.in_stride;
    const int xs = (x0 * s_in + a.ox_lo) & ~3;
    const int dx0 = x0 * s_in + a.ox_lo - xs;
    const int ys = y0 * s_in + a.oy_lo;
    const int plane = a.SH * a.SW;

    if (tid < TM_MAX_SLOTS) toff[tid] = (tid < a.ntaps) ? (a.off_y[tid] - a.oy_lo) * a.SW + (a.off_x[tid] - a.ox_lo) : 0;
    if (tid == 0) {
        tm_prefetch_map(&map_x);
        tm_mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    float acc[4][8];
#pragma unroll
    for (int e = 0; e < 4; ++e)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[e][j] = 0.f;

    const int thr = ry * s_in * a.SW + dx0 + qx * s_in;                  // pixel e sits 8 * e * s_in floats further
    const int wchunk = a.ntaps * a.CC * n8;
    const float* wq = a.wq + (long long)nblk * a.nchunks * wchunk;
    for (int ch = 0; ch < a.nchunks; ++ch) {
        if (tid == 0) {
            tm_mbar_expect_tx(bar, (uint32_t)a.slab_tx);
            tma_load_4d(slab, &map_x, bar, xs, ys, ch * a.CC, b);
        }
        {   // this chunk's weights (coalesced float4)
            const float4* src = (const float4*)(wq + (long long)ch * wchunk);
            float4* dst = (float4*)wsm;
            for (int i = tid; i < wchunk / 4; i += DC_THREADS) dst[i] = __ldg(src + i);
        }
        tm_mbar_wait(bar, ch & 1, 0, 8, ch);
        __syncthreads();
        for (int tp = 0; tp < a.ntaps; ++tp) {
            const float* sp = slab + toff[tp] + thr;
            const float* wp = wsm + tp * a.CC * n8 + ng * 8;
#pragma unroll 4
            for (int c = 0; c < a.CC; ++c) {
                const float i0 = sp[0], i1 = sp[8 * s_in], i2 = sp[16 * s_in], i3 = sp[24 * s_in];
                const float4 w0 = *(const float4*)wp, w1 = *(const float4*)(wp + 4);
                const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
                // packed fp32 FMAs (Blackwell fma.rn.f32x2: two IEEE fmas per issue slot, bit-identical to fmaf): the kernel
                // is issue bound (ncu r01: issue active 79 %, FMA pipe 54 %), 16 packed + 4 operand moves replace 32 scalar FMAs
#pragma unroll
                for (int j = 0; j < 8; j += 2) {
                    ffma2(acc[0][j], acc[0][j + 1], i0, wv[j], wv[j + 1]);
                    ffma2(acc[1][j], acc[1][j + 1], i1, wv[j], wv[j + 1]);
                    ffma2(acc[2][j], acc[2][j + 1], i2, wv[j], wv[j + 1]);
                    ffma2(acc[3][j], acc[3][j + 1], i3, wv[j], wv[j + 1]);
                }
                sp += plane;
                wp += n8;
            }
        }
        __syncthreads();                                                  // the next chunk overwrites slab and weights
    }
    const int oy = y0 + ry;
    if (oy < a.Hc) {
        const long long HWout = (long long)a.Hout * a.Wout;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int ox = x0 + qx + 8 * e;
            if (ox >= a.Wc) continue;
            const long long obase = (long long)b * a.Ntot * HWout + (long long)(oy * a.out_stride + a.out_oy) * a.Wout +
                                    (ox * a.out_stride + a.out_ox);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int n = n0 + ng * 8 + j;
                if (n < a.Ntot) {
                    const long long off = obase + (long long)n * HWout;
                    float o = acc[e][j];
                    if (a.bias) o += __ldg(a.bias + n);
                    if (a.res) o += __ldg(a.res + off);
                    a.out[off] = tm_act(o, a.act, a.slope);
                }
            }
        }
    }
}

// wq[n-block][chunk][tap][cc][n8] from the conv weights (mode 0: w[n][c][tap], mode 1: w[c][n][tap]); zero padded
struct DirectPrepArgs {
    const float* w;
    float* wq;
    int N, Cc, KK, Ci, mode, ntaps, CC, nchunks, n8, nblocks;
    signed char tap_index[TM_MAX_SLOTS];
};
__global__ void __launch_bounds__(256) direct_wprep_kernel(const DirectPrepArgs a) {
    CCB_PDL_WAIT();
    const long long total = (long long)a.nblocks * a.nchunks * a.ntaps * a.CC * a.n8;
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= total) return;
    long long r = i;
    const int j = (int)(r % a.n8); r /= a.n8;
    const int cc = (int)(r % a.CC); r /= a.CC;
    const int tp = (int)(r % a.ntaps); r /= a.ntaps;
    const int ch = (int)(r % a.nchunks); r /= a.nchunks;
    const int n = (int)r * a.n8 + j, c = ch * a.CC + cc;
    float v = 0.f;
    if (n < a.N && c < a.Cc) {
        const int tap = a.tap_index[tp];
        v = (a.mode == 0) ? __ldg(a.w + ((long long)n * a.Cc + c) * a.KK + tap) : __ldg(a.w + ((long long)c * a.Ci + n) * a.KK + tap);
    }
    a.wq[i] = v;
}

struct DirectPlan {
    int NG, TH, CC, nchunks, SW, SH, ox_lo, oy_lo, slab_bytes, smem, n8, nblocks;
    bool ok;
};
static DirectPlan direct_plan(const int* off_y, const int* off_x, int ntaps, int in_stride, int Cc, int N) {
    DirectPlan p;
    memset(&p, 0, sizeof(p));
    if (ntaps < 1) return p;
    int oxl = 0, oxh = 0, oyl = 0, oyh = 0;
    for (int t = 0; t < ntaps; ++t) {
        if (t == 0 || off_x[t] < oxl) oxl = off_x[t];
        if (t == 0 || off_x[t] > oxh) oxh = off_x[t];
        if (t == 0 || off_y[t] < oyl) oyl = off_y[t];
        if (t == 0 || off_y[t] > oyh) oyh = off_y[t];
    }
    p.ox_lo = oxl; p.oy_lo = oyl;
    p.NG = N > 16 ? 4 : (N > 8 ? 2 : 1);
    p.TH = 32 / p.NG;
    p.n8 = p.NG * 8;
    p.nblocks = cdiv(N, p.n8);
    p.SW = (31 * in_stride + (oxh - oxl) + 1 + 3 + 3) & ~3;
    p.SH = (p.TH - 1) * in_stride + (oyh - oyl) + 1;
    if (p.SW > 256 || p.SH > 256) return p;
    const int per_ch = p.SH * p.SW * 4;
    int cc = (44 * 1024) / per_ch;
    if (cc > Cc) cc = Cc;
    if (cc > 32) cc = 32;
    if (cc < 1) return p;
    p.CC = cc;
    p.nchunks = cdiv(Cc, cc);
    p.slab_bytes = cdiv(cc * per_ch, 128) * 128;
    p.smem = p.slab_bytes + ntaps * cc * p.n8 * 4 + TM_MAX_SLOTS * 4 + 64;
    p.ok = p.smem <= 100 * 1024;
    return p;
}
// thin problem: few output channels, short reduction
// measured (tools/tma_probe.py): ~17-20 TFLOP/s whatever the shape; the tensor-core kernels pass that at N = 32 with K >= 288
static bool direct_profitable(int N, int Cc, int ntaps) {
    const long long K = (long long)Cc * ntaps;
    return ntaps >= 1 && K <= 1024 && (N <= 24 || K <= 160);
}

static long long direct_wq_floats(const DirectPlan& p, int ntaps) { return (long long)p.nblocks * p.nchunks * ntaps * p.CC * p.n8; }

static int launch_direct(const float* x, int B, int Cc, int Hin, int Win, const float* w, int mode, int N, int KK, int Ci,
                         const int* off_y, const int* off_x, const int* tap_index, int ntaps, int in_stride, int Hc, int Wc, int Hout,
                         int Wout, int out_stride, int out_oy, int out_ox, const float* bias, const float* res, float* out, int act,
                         float slope, float* work, long long work_floats, cudaStream_t st) {
    EncodeTiledFn enc = get_encode();
    CCB_REQUIRE(enc != nullptr, CCB_ERR_UNSUPPORTED, "conv_direct: cuTensorMapEncodeTiled unavailable");
    const DirectPlan p = direct_plan(off_y, off_x, ntaps, in_stride, Cc, N);
    CCB_REQUIRE(p.ok, CCB_ERR_UNSUPPORTED, "conv_direct: no tiling fits shared memory");
    const long long wqf = direct_wq_floats(p, ntaps);
    CCB_REQUIRE(work && wqf <= work_floats, CCB_ERR_ARG, "conv_direct: workspace too small");
    DirectPrepArgs pa;
    memset(&pa, 0, sizeof(pa));
    DirectArgs a;
    memset(&a, 0, sizeof(a));
    for (int t = 0; t < TM_MAX_SLOTS; ++t) {
        a.off_y[t] = (signed char)(t < ntaps ? off_y[t] : 0);
        a.off_x[t] = (signed char)(t < ntaps ? off_x[t] : 0);
        pa.tap_index[t] = (signed char)(t < ntaps ? tap_index[t] : 0);
    }
    pa.w = w; pa.wq = work; pa.N = N; pa.Cc = Cc; pa.KK = KK; pa.Ci = Ci; pa.mode = mode; pa.ntaps = ntaps; pa.CC = p.CC;
    pa.nchunks = p.nchunks; pa.n8 = p.n8; pa.nblocks = p.nblocks;
    CCB_LAUNCH(direct_wprep_kernel, dim3((unsigned)((wqf + 255) / 256)), dim3(256), 0, st, pa);
    int rc = check_launch("conv_direct wprep");
    if (rc) return rc;
    a.B = B; a.Cin = Cc; a.Hin = Hin; a.Win = Win; a.Ntot = N; a.Hout = Hout; a.Wout = Wout; a.Hc = Hc; a.Wc = Wc;
    a.out_stride = out_stride; a.out_oy = out_oy; a.out_ox = out_ox; a.in_stride = in_stride;
    a.ntaps = ntaps; a.NG = p.NG; a.TH = p.TH; a.CC = p.CC; a.nchunks = p.nchunks;
    a.SW = p.SW; a.SH = p.SH; a.ox_lo = p.ox_lo; a.oy_lo = p.oy_lo; a.slab_bytes = p.slab_bytes; a.slab_tx = p.CC * p.SH * p.SW * 4;
    a.tiles_x = cdiv(Wc, 32); a.tiles_y = cdiv(Hc, p.TH);
    a.wq = work; a.bias = bias; a.res = res; a.out = out; a.act = act; a.slope = slope;
    alignas(64) CUtensorMap map_x;
    {
        cuuint64_t dims[4] = {(cuuint64_t)Win, (cuuint64_t)Hin, (cuuint64_t)Cc, (cuuint64_t)B};
        cuuint64_t strides[3] = {(cuuint64_t)Win * 4, (cuuint64_t)Win * Hin * 4, (cuuint64_t)Win * Hin * Cc * 4};
        cuuint32_t box[4] = {(cuuint32_t)p.SW, (cuuint32_t)p.SH, (cuuint32_t)p.CC, 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = enc(&map_x, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)x, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        CCB_REQUIRE(r == CUDA_SUCCESS, CCB_ERR_LAUNCH, "conv_direct: cuTensorMapEncodeTiled(x) failed (%d)", (int)r);
    }
    dim3 grid(B * a.tiles_x * a.tiles_y, p.nblocks, 1);
    cudaFuncSetAttribute(conv_direct_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    CCB_LAUNCH(conv_direct_kernel, grid, dim3(DC_THREADS), p.smem, st, map_x, a);
    return check_launch("conv_direct");
}

// tap lists: offsets are in pixels of the gathered tensor, tap_index addresses the kh*kw weight plane
static int fprop_taps(const ccb_conv_desc* d, int* oy, int* ox, int* tix) {
    for (int ky = 0; ky < d->kh; ++ky)
        for (int kx = 0; kx < d->kw; ++kx) {
            const int t = ky * d->kw + kx;
            oy[t] = ky - d->pad; ox[t] = kx - d->pad; tix[t] = t;
        }
    return d->kh * d->kw;
}
static int dgrad_taps(const ccb_conv_desc* d, int py, int px, int* oy, int* ox, int* tix) {
    const int s = d->stride;
    int nt = 0;
    for (int ky = 0; ky < d->kh; ++ky) {
        if ((py + d->pad - ky) % s != 0) continue;
        for (int kx = 0; kx < d->kw; ++kx) {
            if ((px + d->pad - kx) % s != 0) continue;
            oy[nt] = (py + d->pad - ky) / s; ox[nt] = (px + d->pad - kx) / s; tix[nt] = ky * d->kw + kx;
            ++nt;
        }
    }
    return nt;
}
// a tap list whose column offsets are all multiples of 4 pixels can be fetched tap by tap (conv_tma_kernel);
// anything else goes through the slab kernel
static bool taps_aligned(const int* ox, int nt, int in_stride, int Cc) {
    if (in_stride != 1 || Cc < 8) return false;
    for (int t = 0; t < nt; ++t)
        if (ox[t] & 3) return false;
    return true;
}

// TMA needs 16-byte aligned rows.  A small gathered tensor whose width is not a multiple of 4 (26, 13 ...) is first
// copied into rows padded with zeros (exactly what the convolution's own zero padding would read there).
__global__ void __launch_bounds__(256) tma_pad_rows_kernel(const float* __restrict__ src, float* __restrict__ dst, long long rows, int W,
                                                           int Wp) {
    CCB_PDL_WAIT();
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= rows * Wp) return;
    const long long r = i / Wp;
    const int x = (int)(i - r * Wp);
    dst[i] = (x < W) ? __ldg(src + r * W + x) : 0.f;
}
// floats of the padded copy of the gathered tensor (0: rows are aligned already, -1: too large to copy)
static long long tma_pad_floats(const ccb_conv_desc* d, int op) {
    const int W = (op == CCB_CONV_FPROP) ? d->Wi : d->Wo;
    if (W % 4 == 0) return 0;
    const long long f = (op == CCB_CONV_FPROP) ? (long long)d->B * d->Ci * d->Hi * ((W + 3) & ~3) : (long long)d->B * d->Co * d->Ho * ((W + 3) & ~3);
    return f <= (4ll << 20) ? f : -1;
}

// Which problems the TMA-fed kernels take: rows of the gathered tensor 16-byte aligned (or small enough to pad), output
// rows that fill most of a 32-column tile.
bool tma_conv_supported(const ccb_conv_desc* d, int op) {
    if (!g_tma_enabled || get_encode() == nullptr) return false;
    if (d->kh != d->kw || d->kh * d->kw > TM_MAX_SLOTS) return false;
    if (op != CCB_CONV_FPROP && op != CCB_CONV_DGRAD) return false;
    if (tma_pad_floats(d, op) < 0) return false;
    if (op == CCB_CONV_FPROP) return (d->stride == 1 || d->stride == 2) && d->Wo >= 20;
    return cdiv(d->Wi, d->stride) >= 20;
}

// does the direct (CUDA-core) kernel take this launch?  `tiles32` = number of 32-column output tiles x rows / 8 (coarse CTA count)
static bool direct_applies(const int* oy, const int* ox, int nt, int in_stride, int Cc, int N, long long out_px) {
    if (!direct_profitable(N, Cc, nt) || out_px < 148ll * 1024) return false;
    return direct_plan(oy, ox, nt, in_stride, Cc, N).ok;
}

static long long tma_wp_floats(const int* oy, const int* ox, int nt, int in_stride, int Cc, int N, long long out_px) {
    if (direct_applies(oy, ox, nt, in_stride, Cc, N, out_px)) return direct_wq_floats(direct_plan(oy, ox, nt, in_stride, Cc, N), nt);
    if (taps_aligned(ox, nt, in_stride, Cc)) {
        const int cb = Cc >= 32 ? 32 : (Cc > 8 ? 16 : 8);
        const int Kp = cdiv((nt > 0 ? nt : 1) * cdiv(Cc, cb) * cb, 32) * 32;
        return 2ll * N * Kp;
    }
    const SlabPlan p = slab_plan(oy, ox, nt, in_stride, Cc, N, 1);
    return p.ok ? 2ll * N * p.ktiles * 32 : -1;
}

// FPROP shapes the direct kernel takes (the dispatcher otherwise keeps thin FPROPs on the register-gather kernel)
bool tma_direct_fprop(const ccb_conv_desc* d) {
    int oy[TM_MAX_SLOTS], ox[TM_MAX_SLOTS], tix[TM_MAX_SLOTS];
    if (d->kh * d->kw > TM_MAX_SLOTS) return false;
    const int nt = fprop_taps(d, oy, ox, tix);
    return direct_applies(oy, ox, nt, d->stride, d->Ci, d->Co, (long long)d->B * d->Ho * d->Wo);
}

// Does the channels-last kernel take the whole op (FPROP: the one gather; DGRAD: every parity class)?  wpf = prepared
// weight floats (max over the classes), tiles = CTAs of one launch before split-K.
static bool nhwc_takes(const ccb_conv_desc* d, int op, long long* wpf_out, long long* tiles_out) {
    int oy[TM_MAX_SLOTS], ox[TM_MAX_SLOTS], tix[TM_MAX_SLOTS];
    if (d->kh * d->kw > TM_MAX_SLOTS) return false;
    long long wpf = 0, tiles = 0;
    if (op == CCB_CONV_FPROP) {
        if (d->stride != 1) return false;
        const int nt = fprop_taps(d, oy, ox, tix);
        // thin layers stay on the CUDA-core direct kernel (measured: 223 vs 248 us for 16 -> 16 at 256x832), except few output
        // channels over >= 32 input channels, where the tensor cores win (32 -> 16 at 128x416: 69 vs 89 us)
        if (!nhwc_prefers_thin() && d->Ci < 32 && direct_applies(oy, ox, nt, 1, d->Ci, d->Co, (long long)d->B * d->Ho * d->Wo)) return false;
        if (!nhwc_applies(oy, ox, nt, 1, d->Ci, d->Co, d->Ho, d->Wo, 1, false)) return false;
        wpf = nhwc_wp_floats(oy, ox, nt, d->Ci, d->Co);
        tiles = (long long)d->B * cdiv(d->Wo, 8) * cdiv(d->Ho, 16) * cdiv(d->Co, 64);
    } else {
        const int s = d->stride;
        for (int py = 0; py < s && py < d->Hi; ++py)
            for (int px = 0; px < s && px < d->Wi; ++px) {
                const int nt = dgrad_taps(d, py, px, oy, ox, tix);
                const int Hc = (d->Hi - py + s - 1) / s, Wc = (d->Wi - px + s - 1) / s;
                if (nt < 1) return false;
                if (!nhwc_prefers_thin() && direct_applies(oy, ox, nt, 1, d->Co, d->Ci, (long long)d->B * Hc * Wc)) return false;
                if (!nhwc_applies(oy, ox, nt, 1, d->Co, d->Ci, Hc, Wc, 1, s > 1)) return false;
                const long long f = nhwc_wp_floats(oy, ox, nt, d->Co, d->Ci);
                if (f > wpf) wpf = f;
            }
        tiles = (long long)d->B * cdiv(cdiv(d->Wi, s), 8) * cdiv(cdiv(d->Hi, s), 16) * cdiv(d->Ci, 64);
    }
    if (wpf < 0) return false;
    *wpf_out = wpf; *tiles_out = tiles;
    return true;
}
bool tma_nhwc_takes(const ccb_conv_desc* d, int op) {
    long long a, b;
    return nhwc_takes(d, op, &a, &b);
}
static int nhwc_plan_splits(long long tiles, int cblocks, long long out_numel, long long part_floats) {
    // splitting below two CTAs per SM costs nothing measurable (33.95 vs 34.0 ms per step without it) and the fp32 sum of the
    // partials shortens the tensor core's accumulation chains: without it 8 instead of 5 Back2Future gradient tensors
    // exceeded 1e-3 at full size
    if (tiles >= 2 * 148 || cblocks < 4) return 1;
    long long s = (3 * 148 + tiles - 1) / tiles;
    if (s > cblocks / 2) s = cblocks / 2;
    if (s > 8) s = 8;
    if (out_numel > 0 && s * out_numel > part_floats) s = part_floats / out_numel;
    return s < 2 ? 1 : (int)s;
}

long long tma_workspace_floats(const ccb_conv_desc* d, int op) {
    int oy[TM_MAX_SLOTS], ox[TM_MAX_SLOTS], tix[TM_MAX_SLOTS];
    long long wpf = 0, tiles, out_numel;
    {
        long long nwpf, ntiles;
        if (nhwc_takes(d, op, &nwpf, &ntiles)) {
            const long long copyf = (op == CCB_CONV_FPROP) ? nhwc_copy_floats(d->B, d->Ci, d->Hi, d->Wi) : nhwc_copy_floats(d->B, d->Co, d->Ho, d->Wo);
            const long long on = (op == CCB_CONV_FPROP) ? (long long)d->B * d->Co * d->Ho * d->Wo : (long long)d->B * d->Ci * d->Hi * d->Wi;
            return copyf + nwpf + (ntiles < 2 * 148 ? 8 * on : 0);
        }
    }
    if (op == CCB_CONV_FPROP) {
        const int nt = fprop_taps(d, oy, ox, tix);
        wpf = tma_wp_floats(oy, ox, nt, d->stride, d->Ci, d->Co, (long long)d->B * d->Ho * d->Wo);
        tiles = (long long)d->B * cdiv(d->Wo, 32) * cdiv(d->Ho, 4) * cdiv(d->Co, 128);
        out_numel = (long long)d->B * d->Co * d->Ho * d->Wo;
    } else {
        const int s = d->stride;
        for (int py = 0; py < s && py < d->Hi; ++py)
            for (int px = 0; px < s && px < d->Wi; ++px) {
                const int nt = dgrad_taps(d, py, px, oy, ox, tix);
                const long long f = tma_wp_floats(oy, ox, nt, 1, d->Co, d->Ci, (long long)d->B * cdiv(d->Hi, s) * cdiv(d->Wi, s));
                if (f < 0) return -1;
                if (f > wpf) wpf = f;
            }
        tiles = (long long)d->B * cdiv(cdiv(d->Wi, s), 32) * cdiv(cdiv(d->Hi, s), 4) * cdiv(d->Ci, 128);
        out_numel = (long long)d->B * d->Ci * d->Hi * d->Wi;
    }
    if (wpf < 0) return -1;
    const long long padf = tma_pad_floats(d, op);
    if (padf < 0) return -1;
    return padf + wpf + (tiles < 148 ? 8 * out_numel : 0);      // padded copy + weights + room for up to 8 split-K partials
}

static int tma_plan_splits(long long tiles, int ktiles, long long out_numel, long long part_floats) {
    if (tiles >= 148 || ktiles < 4) return 1;
    long long s = (2 * 148 + tiles - 1) / tiles;
    if (s > ktiles / 2) s = ktiles / 2;
    if (s > 8) s = 8;
    if (out_numel > 0 && s * out_numel > part_floats) s = part_floats / out_numel;
    return s < 2 ? 1 : (int)s;
}

// one launch of whichever kernel fits the tap list
static int launch_any(const float* x, int B, int Cc, int Hin, int Win, const float* w, int mode, int N, int KK, int Ci,
                      const int* oy, const int* ox, const int* tix, int nt, int in_stride, int Hc, int Wc, int Hout, int Wout,
                      int out_stride, int out_oy, int out_ox, const float* bias, const float* res, float* out, int act, float slope,
                      int three, float* work, long long wp_floats, int splits, float* partial, long long out_numel, cudaStream_t st) {
    if (splits == 1 && direct_applies(oy, ox, nt, in_stride, Cc, N, (long long)B * Hc * Wc))
        return launch_direct(x, B, Cc, Hin, Win, w, mode, N, KK, Ci, oy, ox, tix, nt, in_stride, Hc, Wc, Hout, Wout, out_stride, out_oy,
                             out_ox, bias, res, out, act, slope, work, wp_floats, st);
    if (taps_aligned(ox, nt, in_stride, Cc))
        return launch_tma(x, B, Cc, Hin, Win, w, mode, N, KK, Ci, oy, ox, tix, nt, Hc, Wc, Hout, Wout, out_stride, out_oy, out_ox, bias,
                          res, out, act, slope, three, work, wp_floats, splits, partial, out_numel, st);
    return launch_slab(x, B, Cc, Hin, Win, w, mode, N, KK, Ci, oy, ox, tix, nt, in_stride, Hc, Wc, Hout, Wout, out_stride, out_oy,
                       out_ox, bias, res, out, act, slope, three, work, wp_floats, splits, partial, out_numel, st);
}

int tma_fprop(const ccb_conv_desc* d, const float* x, const float* w, const float* bias, const float* res, float* y, float* work,
              long long work_floats, int three, cudaStream_t st) {
    int oy[TM_MAX_SLOTS], ox[TM_MAX_SLOTS], tix[TM_MAX_SLOTS];
    const int nt = fprop_taps(d, oy, ox, tix);
    const long long out_numel = (long long)d->B * d->Co * d->Ho * d->Wo;
    {
        long long nwpf, ntiles;
        if (nhwc_takes(d, CCB_CONV_FPROP, &nwpf, &ntiles)) {                      // channels-last slab kernel
            const long long copyf = nhwc_copy_floats(d->B, d->Ci, d->Hi, d->Wi);
            CCB_REQUIRE(copyf + nwpf <= work_floats, CCB_ERR_ARG, "conv_nhwc fprop: workspace too small");
            int rc = nhwc_transpose(x, work, d->B, d->Ci, d->Hi, d->Wi, st);
            if (rc) return rc;
            float* wk = work + copyf;
            const int splits = nhwc_plan_splits(ntiles, cdiv(d->Ci, 32), out_numel, work_floats - copyf - nwpf);
            rc = launch_nhwc(work, d->B, d->Ci, d->Hi, d->Wi, w, 0, d->Co, nt, d->Ci, oy, ox, tix, nt, d->Ho, d->Wo, d->Ho, d->Wo, 1, 0, 0, bias,
                             res, y, d->act, d->slope, three, wk, nwpf, splits, wk + nwpf, out_numel, st);
            if (rc || splits == 1) return rc;
            launch_splitk_reduce(wk + nwpf, y, bias, res, out_numel, splits, d->Ho * d->Wo, d->Co, d->act, d->slope, st);
            return check_launch("conv_nhwc splitk reduce");
        }
    }
    const long long wpf = tma_wp_floats(oy, ox, nt, d->stride, d->Ci, d->Co, (long long)d->B * d->Ho * d->Wo);
    const long long padf = tma_pad_floats(d, CCB_CONV_FPROP);
    CCB_REQUIRE(wpf >= 0 && padf >= 0 && padf + wpf <= work_floats, CCB_ERR_ARG, "conv_tma fprop: workspace too small");
    int Wi = d->Wi;
    if (padf > 0) {                                              // narrow unaligned map: padded copy first
        Wi = (d->Wi + 3) & ~3;
        CCB_LAUNCH(tma_pad_rows_kernel, dim3((unsigned)((padf + 255) / 256)), dim3(256), 0, st, x, work, (long long)d->B * d->Ci * d->Hi,
                   d->Wi, Wi);
        x = work; work += padf; work_floats -= padf;
    }
    const long long tiles = (long long)d->B * cdiv(d->Wo, 32) * cdiv(d->Ho, 4) * cdiv(d->Co, 128);
    const int splits = tma_plan_splits(tiles, (int)(wpf / (64ll * d->Co)), out_numel, work_floats - wpf);
    int rc = launch_any(x, d->B, d->Ci, d->Hi, Wi, w, 0, d->Co, nt, d->Ci, oy, ox, tix, nt, d->stride, d->Ho, d->Wo, d->Ho, d->Wo, 1,
                        0, 0, bias, res, y, d->act, d->slope, three, work, wpf, splits, work + wpf, out_numel, st);
    if (rc || splits == 1) return rc;
    launch_splitk_reduce(work + wpf, y, bias, res, out_numel, splits, d->Ho * d->Wo, d->Co, d->act, d->slope, st);
    return check_launch("conv_tma splitk reduce");
}

int tma_dgrad(const ccb_conv_desc* d, const float* dy, const float* w, const float* bias, const float* res, float* dx, float* work,
              long long work_floats, int three, cudaStream_t st) {
    const int s = d->stride;
    const long long out_numel = (long long)d->B * d->Ci * d->Hi * d->Wi;
    int oy[TM_MAX_SLOTS], ox[TM_MAX_SLOTS], tix[TM_MAX_SLOTS];
    {
        long long nwpf, ntiles;
        if (nhwc_takes(d, CCB_CONV_DGRAD, &nwpf, &ntiles)) {                      // channels-last slab kernel, one copy of dy for all classes
            const long long copyf = nhwc_copy_floats(d->B, d->Co, d->Ho, d->Wo);
            CCB_REQUIRE(copyf + nwpf <= work_floats, CCB_ERR_ARG, "conv_nhwc dgrad: workspace too small");
            int rc = nhwc_transpose(dy, work, d->B, d->Co, d->Ho, d->Wo, st);
            if (rc) return rc;
            float* wk = work + copyf;
            // split-K partials of different parity classes would alias: split only stride-1 problems
            const int splits = (s == 1) ? nhwc_plan_splits(ntiles, cdiv(d->Co, 32), out_numel, work_floats - copyf - nwpf) : 1;
            for (int py = 0; py < s && py < d->Hi; ++py)
                for (int px = 0; px < s && px < d->Wi; ++px) {
                    const int nt = dgrad_taps(d, py, px, oy, ox, tix);
                    const int Hc = (d->Hi - py + s - 1) / s, Wc = (d->Wi - px + s - 1) / s;
                    rc = launch_nhwc(work, d->B, d->Co, d->Ho, d->Wo, w, 1, d->Ci, d->kh * d->kw, d->Ci, oy, ox, tix, nt, Hc, Wc, d->Hi, d->Wi, s,
                                     py, px, bias, res, dx, d->act, d->slope, three, wk, nwpf, splits, wk + nwpf, out_numel, st);
                    if (rc) return rc;
                }
            if (splits > 1) {
                launch_splitk_reduce(wk + nwpf, dx, bias, res, out_numel, splits, d->Hi * d->Wi, d->Ci, d->act, d->slope, st);
                return check_launch("conv_nhwc dgrad splitk reduce");
            }
            return CCB_OK;
        }
    }
    long long wpf_max = 0;
    for (int py = 0; py < s && py < d->Hi; ++py)
        for (int px = 0; px < s && px < d->Wi; ++px) {
            const int nt = dgrad_taps(d, py, px, oy, ox, tix);
            const long long f = tma_wp_floats(oy, ox, nt, 1, d->Co, d->Ci, (long long)d->B * cdiv(d->Hi, s) * cdiv(d->Wi, s));
            CCB_REQUIRE(f >= 0, CCB_ERR_UNSUPPORTED, "conv_tma dgrad: no tiling");
            if (f > wpf_max) wpf_max = f;
        }
    const long long padf = tma_pad_floats(d, CCB_CONV_DGRAD);
    CCB_REQUIRE(padf >= 0 && padf + wpf_max <= work_floats, CCB_ERR_ARG, "conv_tma dgrad: workspace too small");
    int Wo = d->Wo;
    if (padf > 0) {
        Wo = (d->Wo + 3) & ~3;
        CCB_LAUNCH(tma_pad_rows_kernel, dim3((unsigned)((padf + 255) / 256)), dim3(256), 0, st, dy, work, (long long)d->B * d->Co * d->Ho,
                   d->Wo, Wo);
        dy = work; work += padf; work_floats -= padf;
    }
    const long long tiles = (long long)d->B * cdiv(cdiv(d->Wi, s), 32) * cdiv(cdiv(d->Hi, s), 4) * cdiv(d->Ci, 128);
    const int splits = tma_plan_splits(tiles, (int)(wpf_max / (64ll * d->Ci)), out_numel, work_floats - wpf_max);
    for (int py = 0; py < s && py < d->Hi; ++py)
        for (int px = 0; px < s && px < d->Wi; ++px) {
            const int nt = dgrad_taps(d, py, px, oy, ox, tix);
            const int Hc = (d->Hi - py + s - 1) / s, Wc = (d->Wi - px + s - 1) / s;
            int rc = launch_any(dy, d->B, d->Co, d->Ho, Wo, w, 1, d->Ci, d->kh * d->kw, d->Ci, oy, ox, tix, nt, 1, Hc, Wc, d->Hi, d->Wi,
                                s, py, px, bias, res, dx, d->act, d->slope, three, work, wpf_max, splits, work + wpf_max, out_numel, st);
            if (rc) return rc;
        }
    if (splits > 1) {
        launch_splitk_reduce(work + wpf_max, dx, bias, res, out_numel, splits, d->Hi * d->Wi, d->Ci, d->act, d->slope, st);
        return check_launch("conv_tma dgrad splitk reduce");
    }
    return CCB_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// WGRAD: dw[co][ci][ky][kx] = sum over (b, oy, ox) of dy[b,co,oy,ox] * x[b,ci,oy*s+ky-p, ox*s+kx-p]
//   GEMM with K = output pixels, one k-stage = 32 pixels of one output row:
//     B: one TMA box (32 px, 1 row, n channels) of dy - K-major SWIZZLE_128B as it lands; the cutters add its
//        tf32 remainder right behind it, so that [B_hi | B_lo] is one 2n-wide operand;
//     A: rows = (tap, input channel) pairs of this CTA's M tile.  The input rows under those taps arrive as one
//        aligned slab box; the cutters copy the tap-shifted 32-pixel runs into the K-major operand tile (value and
//        remainder).  Zero padding and row tails are the TMA's OOB fill.
//   D[(tap, ci)][co] accumulates in TMEM over this CTA's share of the pixels; split partials are summed in a fixed
//   order by tma_splitk_sum_kernel.
// ---------------------------------------------------------------------------------------------------------------
struct SlabWgradArgs {
    int B, Ci, Co, Ho, Wo, KK, kw;
    int stride, pad;
    int cwid, cblocks, tpt, tgroups;       // channels per M tile, channel blocks, taps per M tile, tap groups
    int SW, SH, slab_bytes, slab_tx, dx0;
    int segs, stages, per_split, splits;
    int nbox, nstages, soft;
    long long numel;
    float* out;
};

template <bool THREE>
__global__ void __launch_bounds__(SL_THREADS, 1)
conv_slab_wgrad_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_dy, const SlabWgradArgs a) {
    CCB_PDL_WAIT();
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_addr(smem_raw) & 1023u)) & 1023u);
    const int NST = a.nstages;
    const int b_bytes = (THREE ? 2 : 1) * a.nbox * 128;
    const int ab_bytes = (THREE ? 2 : 1) * TM_A_BYTES + b_bytes;              // [A hi | A lo] [B hi ; B lo]
    const int stage_bytes = ab_bytes + a.slab_bytes;                           // ... [slab]
    uint64_t* bars = (uint64_t*)(smem + NST * stage_bytes);
    uint64_t* tma_full = bars;           // [8] slab + dy box
    uint64_t* a_full = bars + 8;         // [8] 256 cutter threads
    uint64_t* empty_bar = bars + 16;     // [8] tcgen05.commit
    uint64_t* accum_bar = bars + 24;
    uint32_t* tmem_slot = (uint32_t*)(bars + 25);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tg = blockIdx.x / a.cblocks, cblock = blockIdx.x - tg * a.cblocks;
    const int n0 = blockIdx.y * 128;
    const int ntile = min(128, a.Co - n0);
    const int st_beg = blockIdx.z * a.per_split;
    const int nst = max(0, min(a.stages, st_beg + a.per_split) - st_beg);
    const int tap0 = tg * a.tpt;
    const int ntap = min(a.tpt, a.KK - tap0);                                  // taps of this tile
    const int c0 = cblock * a.cwid;
    const int nch = min(a.cwid, a.Ci - c0);
    const int ky_lo = tap0 / a.kw;

    if (tid == 0) {
        tm_prefetch_map(&map_x);
        tm_prefetch_map(&map_dy);
        for (int s = 0; s < NST; ++s) {
            tm_mbar_init(&tma_full[s], 1);
            tm_mbar_init(&a_full[s], 8);
            tm_mbar_init(&empty_bar[s], 1);
        }
        tm_mbar_init(accum_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // operand rows without a (tap, channel) behind them stay zero for the whole launch
    for (int s = 0; s < NST; ++s) {
        float4* p = (float4*)(smem + s * stage_bytes);
        for (int i = tid; i < (THREE ? 2048 : 1024); i += SL_THREADS) p[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            const uint32_t tx_bytes = (uint32_t)(a.slab_tx + a.nbox * 128);
            const int per_img = a.Ho * a.segs;
            int g = st_beg;
            int b = g / per_img;
            int r = g - b * per_img;
            int oy = r / a.segs, seg = r - oy * a.segs;
            int s = 0;
            uint32_t ph = 0;
            for (int it = 0; it < nst; ++it) {
                if (it >= NST) tm_mbar_wait(&empty_bar[s], ph ^ 1, a.soft, 1, it);
                unsigned char* st = smem + s * stage_bytes;
                tm_mbar_expect_tx(&tma_full[s], tx_bytes);
                tma_load_4d(st + ab_bytes, &map_x, &tma_full[s], seg * 32 * a.stride - a.pad - a.dx0, oy * a.stride + ky_lo - a.pad, c0, b);
                tma_load_4d(st + (THREE ? 2 : 1) * TM_A_BYTES, &map_dy, &tma_full[s], seg * 32, oy, n0, b);
                if (++seg == a.segs) { seg = 0; if (++oy == a.Ho) { oy = 0; ++b; } }
                if (++s == NST) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        // the whole warp runs the loop, lane 0's predicate issues (both operands K-major SWIZZLE_128B)
        const uint32_t idesc_base = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TM_M >> 4) << 24);
        const uint32_t idesc_n1 = idesc_base | ((uint32_t)(a.nbox >> 3) << 17);
        const uint32_t idesc_n2 = idesc_base | ((uint32_t)((2 * a.nbox) >> 3) << 17);
        const uint32_t leader = (lane == 0) ? 1u : 0u;
        const uint32_t smem16 = smem_addr(smem) >> 4, stage16 = (uint32_t)stage_bytes >> 4;
        int s = 0;
        uint32_t ph = 0;
        for (int it = 0; it < nst; ++it) {
            tm_mbar_wait(&a_full[s], ph, a.soft, 3, it);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t a_hi16 = smem16 + (uint32_t)s * stage16, a_lo16 = a_hi16 + (TM_A_BYTES >> 4);
            const uint32_t b16 = a_hi16 + (((THREE ? 2 : 1) * TM_A_BYTES) >> 4);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                const uint32_t acc = (it > 0 || ks > 0) ? 1u : 0u;
                tm_umma_tf32_p(tmem_base, (a_hi16 + ks * 2) | DESC_K_LO, DESC_K_HI, (b16 + ks * 2) | DESC_K_LO, DESC_K_HI,
                               THREE ? idesc_n2 : idesc_n1, acc, leader);
                if (THREE)
                    tm_umma_tf32_p(tmem_base + 256u, (a_lo16 + ks * 2) | DESC_K_LO, DESC_K_HI, (b16 + ks * 2) | DESC_K_LO, DESC_K_HI,
                                   idesc_n1, acc, leader);
            }
            tm_commit_p(&empty_bar[s], leader);
            if (it == nst - 1) tm_commit_p(accum_bar, leader);
            if (++s == NST) { s = 0; ph ^= 1; }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    } else {
        // two groups of 8 cutter warps take alternate k-stages.  Inside a group warp w8 owns operand rows
        // 16*w8 .. 16*w8+15; a quarter warp writes one 128-byte row (32 pixels): lane -> (row within a group of 4,
        // group of 4 pixels).  Row -> (tap, channel) never changes.
        const int grp = (warp - 2) >> 3, w8 = (warp - 2) & 7, wt = (tid - 64) & 255;
        const int pxg = lane & 7, sub = lane >> 3;
        const int plane = a.SH * a.SW;
        int ridx[4];
        uint32_t rdst[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int r = w8 * 16 + j * 4 + sub;
            const int tl = r / a.cwid, c = r - tl * a.cwid;
            const int tap = tap0 + tl;
            const int ky = tap / a.kw, kx = tap - ky * a.kw;
            ridx[j] = (tl < ntap && c < nch) ? c * plane + (ky - ky_lo) * a.SW + a.dx0 + kx + pxg * 4 * a.stride : -1;
            rdst[j] = (uint32_t)(r * 128 + ((pxg ^ (r & 7)) << 4));
        }
        const int b4 = a.nbox * 8;                                             // float4 of the dy tile
        int s = grp % NST;
        uint32_t ph = (grp / NST) & 1;
        for (int it = grp; it < nst; it += 2) {
            tm_mbar_wait(&tma_full[s], ph, a.soft, 4, it);
            unsigned char* st = smem + s * stage_bytes;
            const float* slab = (const float*)(st + ab_bytes);
            float4 v[4];
            if (a.stride == 1) {
                float4 va[4], vb[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) va[j] = *(const float4*)(slab + (max(ridx[j], 0) & ~3));
#pragma unroll
                for (int j = 0; j < 4; ++j) vb[j] = *(const float4*)(slab + (max(ridx[j], 0) & ~3) + 4);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int al = ridx[j] & 3;
                    v[j] = al == 0 ? va[j]
                         : al == 1 ? make_float4(va[j].y, va[j].z, va[j].w, vb[j].x)
                         : al == 2 ? make_float4(va[j].z, va[j].w, vb[j].x, vb[j].y)
                                   : make_float4(va[j].w, vb[j].x, vb[j].y, vb[j].z);
                }
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int i0 = max(ridx[j], 0);
                    v[j].x = slab[i0]; v[j].y = slab[i0 + a.stride]; v[j].z = slab[i0 + 2 * a.stride]; v[j].w = slab[i0 + 3 * a.stride];
                }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (ridx[j] >= 0) {
                    *(float4*)(st + rdst[j]) = v[j];
                    if (THREE) *(float4*)(st + TM_A_BYTES + rdst[j]) = tf32_rest4(v[j]);
                }
            }
            if (THREE) {
                float4* braw = (float4*)(st + 2 * TM_A_BYTES);
                float4* blo = braw + b4;
                for (int i = wt; i < b4; i += 256) blo[i] = tf32_rest4(braw[i]);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) tm_mbar_arrive(&a_full[s]);
            s += 2;
            if (s >= NST) { s -= NST; ph ^= 1; }
        }
        if (nst > 0) tm_mbar_wait(accum_bar, 0, a.soft, 5, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int q4 = warp & 3, colq = (warp - 2) >> 2;
        const int m = q4 * 32 + lane;                       // TMEM lane = operand row = (tap, channel)
        const int tl = m / a.cwid, cl = m - tl * a.cwid;
        const int tap = tap0 + tl, c = c0 + cl;
        const bool rvalid = (tl < ntap) && (cl < nch);
        float* outp = a.out + (long long)blockIdx.z * a.numel;
        const uint32_t trow = tmem_base + ((uint32_t)(q4 * 32) << 16);
        for (int cg = colq; cg * 16 < ntile; cg += 4) {
            float v[16];
            if (nst > 0) {
                tm_ld16(trow + (uint32_t)(cg * 16), v);
                if (THREE) {
                    float v2[16];
                    tm_ld16(trow + (uint32_t)(a.nbox + cg * 16), v2);
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] += v2[j];
                    tm_ld16(trow + 256u + (uint32_t)(cg * 16), v2);
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] += v2[j];
                }
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = 0.f;
            }
            if (rvalid) {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int co = n0 + cg * 16 + j;
                    if (cg * 16 + j < ntile) outp[((long long)co * a.Ci + c) * a.KK + tap] = v[j];
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

__global__ void __launch_bounds__(256) tma_splitk_sum_kernel(const float* __restrict__ work, float* __restrict__ out, long long numel,
                                                             int splits) {
    CCB_PDL_WAIT();
    long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= numel) return;
    float v = 0.f;
    for (int s = 0; s < splits; ++s) v += __ldg(work + (long long)s * numel + i);
    out[i] = v;
}

static bool wgrad_plan(const ccb_conv_desc* d, int three, SlabWgradArgs& a, int& smem) {
    const int KK = d->kh * d->kw;
    a.cwid = d->Ci < 128 ? d->Ci : 128;
    a.cblocks = cdiv(d->Ci, a.cwid);
    const int tpt_max = 128 / a.cwid;
    a.tgroups = cdiv(KK, tpt_max);
    a.tpt = cdiv(KK, a.tgroups);                           // balanced tap groups
    // input rows under one tap group
    int sh = 1;
    for (int g = 0; g < a.tgroups; ++g) {
        const int t0 = g * a.tpt, t1 = (t0 + a.tpt < KK ? t0 + a.tpt : KK) - 1;
        const int span = t1 / d->kw - t0 / d->kw + 1;
        if (span > sh) sh = span;
    }
    a.SH = sh;
    a.dx0 = ((-d->pad) % 4 + 4) % 4;                       // column of tap kx = 0, pixel 0 inside the aligned slab
    a.SW = (31 * d->stride + d->kw + a.dx0 + 3) & ~3;
    a.slab_tx = a.cwid * a.SH * a.SW * 4;
    a.slab_bytes = cdiv(a.slab_tx, 1024) * 1024;
    const int ntile = d->Co < 128 ? d->Co : 128;
    a.nbox = (ntile + 15) & ~15;
    const int stage = (three ? 2 : 1) * (TM_A_BYTES + a.nbox * 128) + a.slab_bytes;
    a.nstages = (224 * 1024) / stage;
    if (a.nstages > 6) a.nstages = 6;
    if (a.nstages < 2 || a.SW > 256 || a.cwid > 256) return false;
    smem = a.nstages * stage + 512 + 1024;
    a.segs = cdiv(d->Wo, 32);
    a.stages = d->B * d->Ho * a.segs;
    const int tiles = a.tgroups * a.cblocks * cdiv(d->Co, 128);
    int splits = tiles >= 148 ? 1 : 148 / tiles;
    if (splits > a.stages / 8) splits = a.stages / 8;
    if (splits < 1) splits = 1;
    a.per_split = cdiv(a.stages, splits);
    a.splits = cdiv(a.stages, a.per_split);
    return true;
}

bool tma_wgrad_supported(const ccb_conv_desc* d) {
    if (!g_tma_enabled || get_encode() == nullptr) return false;
    if (d->kh != d->kw || (d->stride != 1 && d->stride != 2)) return false;
    if ((d->Wi % 4) || (d->Wo % 4) || d->Wo < 20) return false;
    SlabWgradArgs a;
    int smem;
    return wgrad_plan(d, 1, a, smem);
}
long long tma_wgrad_workspace_floats(const ccb_conv_desc* d) {
    SlabWgradArgs a;
    int smem;
    if (!wgrad_plan(d, 1, a, smem)) return 0;
    return a.splits > 1 ? (long long)a.splits * d->Co * d->Ci * d->kh * d->kw : 0;
}

int tma_wgrad(const ccb_conv_desc* d, const float* x, const float* dy, float* dw, float* work, long long work_floats, int three,
              cudaStream_t st) {
    EncodeTiledFn enc = get_encode();
    CCB_REQUIRE(enc != nullptr, CCB_ERR_UNSUPPORTED, "conv_tma: cuTensorMapEncodeTiled unavailable");
    SlabWgradArgs a;
    memset(&a, 0, sizeof(a));
    int smem = 0;
    CCB_REQUIRE(wgrad_plan(d, three, a, smem), CCB_ERR_UNSUPPORTED, "conv_tma wgrad: no tiling fits shared memory");
    a.B = d->B; a.Ci = d->Ci; a.Co = d->Co; a.Ho = d->Ho; a.Wo = d->Wo; a.KK = d->kh * d->kw; a.kw = d->kw;
    a.stride = d->stride; a.pad = d->pad; a.soft = g_tma_soft;
    a.numel = (long long)d->Co * d->Ci * a.KK;
    if (a.splits > 1) {
        CCB_REQUIRE(work && (long long)a.splits * a.numel <= work_floats, CCB_ERR_ARG, "conv_tma wgrad: workspace too small");
        a.out = work;
    } else {
        a.out = dw;
    }
    alignas(64) CUtensorMap map_x, map_dy;
    {
        cuuint64_t dims[4] = {(cuuint64_t)d->Wi, (cuuint64_t)d->Hi, (cuuint64_t)d->Ci, (cuuint64_t)d->B};
        cuuint64_t strides[3] = {(cuuint64_t)d->Wi * 4, (cuuint64_t)d->Wi * d->Hi * 4, (cuuint64_t)d->Wi * d->Hi * d->Ci * 4};
        cuuint32_t box[4] = {(cuuint32_t)a.SW, (cuuint32_t)a.SH, (cuuint32_t)a.cwid, 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = enc(&map_x, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)x, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        CCB_REQUIRE(r == CUDA_SUCCESS, CCB_ERR_LAUNCH, "conv_tma wgrad: cuTensorMapEncodeTiled(x) failed (%d)", (int)r);
    }
    {
        cuuint64_t dims[4] = {(cuuint64_t)d->Wo, (cuuint64_t)d->Ho, (cuuint64_t)d->Co, (cuuint64_t)d->B};
        cuuint64_t strides[3] = {(cuuint64_t)d->Wo * 4, (cuuint64_t)d->Wo * d->Ho * 4, (cuuint64_t)d->Wo * d->Ho * d->Co * 4};
        cuuint32_t box[4] = {32, 1, (cuuint32_t)a.nbox, 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = enc(&map_dy, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)dy, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        CCB_REQUIRE(r == CUDA_SUCCESS, CCB_ERR_LAUNCH, "conv_tma wgrad: cuTensorMapEncodeTiled(dy) failed (%d)", (int)r);
    }
    dim3 grid(a.tgroups * a.cblocks, cdiv(d->Co, 128), a.splits);
    auto kfn = three ? conv_slab_wgrad_kernel<true> : conv_slab_wgrad_kernel<false>;
    cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    CCB_LAUNCH(kfn, grid, dim3(SL_THREADS), smem, st, map_x, map_dy, a);
    int rc = check_launch("conv_slab_wgrad");
    if (rc || a.splits == 1) return rc;
    CCB_LAUNCH(tma_splitk_sum_kernel, dim3((unsigned)((a.numel + 255) / 256)), dim3(256), 0, st, (const float*)work, dw, a.numel, a.splits);
    return check_launch("conv_slab_wgrad_reduce");
}

}  // namespace ccb

// Host-side tiling decisions of the TMA-fed family for one problem, without launching anything (and without needing a
// driver): lets the CPU test-suite sweep every layer shape of the four networks and check the invariants the kernels
// rely on (shared memory, TMEM columns, K coverage, TMA box limits).
//   op FPROP / DGRAD (parity class py, px): out = {kind, cs, cblocks, kt_full, ktiles, SW, SH, slab_bytes, nslab, nstages,
//       nbox, smem, mt, ntaps, span_x, span_y}   kind: 2 slab, 3 aligned per-tap TMA, 4 direct (then cs = CC, cblocks = nchunks,
//       kt_full = NG, ktiles = TH, nbox = n8, mt = n-blocks), -1 no tiling
//   op 16 + FPROP / 16 + DGRAD: out = {6 / -1, cblocks, SH, SW, mt, slab_bytes, slab_tx, nslab, nstages, nbox, ntile_w, tmem_cols, smem, Kp,
//       span_x << 16 | span_y, CTAs}: the channels-last kernel's tiling (conv_nhwc.cu) when the shape qualifies
//   op WGRAD: out = {kind 5 / -1, cwid, cblocks, tpt, tgroups, SW, SH, slab_bytes, 1, nstages, nbox, smem, splits, KK, dx0, stages}
extern "C" int ccb_debug_conv_plan(const ccb_conv_desc* d, int op, int py, int px, int* out16) {
    using namespace ccb;
    if (!d || !out16 || d->kh != d->kw || d->kh * d->kw > TM_MAX_SLOTS) return CCB_ERR_ARG;
    for (int i = 0; i < 16; ++i) out16[i] = 0;
    if (op == CCB_CONV_WGRAD) {
        SlabWgradArgs a;
        memset(&a, 0, sizeof(a));
        int smem = 0;
        if (!wgrad_plan(d, 1, a, smem)) { out16[0] = -1; return CCB_OK; }
        const int v[16] = {5, a.cwid, a.cblocks, a.tpt, a.tgroups, a.SW, a.SH, a.slab_bytes, 1, a.nstages, a.nbox, smem, a.splits,
                           d->kh * d->kw, a.dx0, a.stages};
        for (int i = 0; i < 16; ++i) out16[i] = v[i];
        return CCB_OK;
    }
    int oy[TM_MAX_SLOTS], ox[TM_MAX_SLOTS], tix[TM_MAX_SLOTS];
    int nt, in_stride, Cc, N;
    long long out_px;
    if (op >= 16) {                                              // 16 + FPROP / DGRAD: the channels-last plan of this gather (kind 6) or -1
        const int s = d->stride;
        if (op - 16 == CCB_CONV_FPROP) {
            nt = fprop_taps(d, oy, ox, tix);
            nhwc_debug_plan(oy, ox, nt, d->stride, d->Ci, d->Co, d->B, d->Ho, d->Wo, false, out16);
        } else {
            nt = dgrad_taps(d, py, px, oy, ox, tix);
            if (nt < 1) { out16[0] = -1; return CCB_OK; }
            nhwc_debug_plan(oy, ox, nt, 1, d->Co, d->Ci, d->B, (d->Hi - py + s - 1) / s, (d->Wi - px + s - 1) / s, s > 1, out16);
        }
        int sx = 0, sy = 0;
        for (int t = 0; t < nt; ++t)
            for (int u = 0; u < nt; ++u) {
                if (ox[t] - ox[u] > sx) sx = ox[t] - ox[u];
                if (oy[t] - oy[u] > sy) sy = oy[t] - oy[u];
            }
        if (out16[0] == 6) { out16[14] = sx * 65536 + sy; }      // spans replace Cp (the test derives Cp itself)
        return CCB_OK;
    }
    if (op == CCB_CONV_FPROP) {
        nt = fprop_taps(d, oy, ox, tix); in_stride = d->stride; Cc = d->Ci; N = d->Co;
        out_px = (long long)d->B * d->Ho * d->Wo;
    } else {
        nt = dgrad_taps(d, py, px, oy, ox, tix); in_stride = 1; Cc = d->Co; N = d->Ci;
        out_px = (long long)d->B * cdiv(d->Hi, d->stride) * cdiv(d->Wi, d->stride);
    }
    int sx = 0, sy = 0;
    for (int t = 0; t < nt; ++t)
        for (int u = 0; u < nt; ++u) {
            if (ox[t] - ox[u] > sx) sx = ox[t] - ox[u];
            if (oy[t] - oy[u] > sy) sy = oy[t] - oy[u];
        }
    out16[13] = nt; out16[14] = sx; out16[15] = sy;
    if (direct_applies(oy, ox, nt, in_stride, Cc, N, out_px)) {
        const DirectPlan p = direct_plan(oy, ox, nt, in_stride, Cc, N);
        const int v[13] = {4, p.CC, p.nchunks, p.NG, p.TH, p.SW, p.SH, p.slab_bytes, 1, 1, p.n8, p.smem, p.nblocks};
        for (int i = 0; i < 13; ++i) out16[i] = v[i];
        return CCB_OK;
    }
    if (taps_aligned(ox, nt, in_stride, Cc)) { out16[0] = 3; return CCB_OK; }
    const SlabPlan p = slab_plan(oy, ox, nt, in_stride, Cc, N, 1, 4);
    if (!p.ok) { out16[0] = -1; return CCB_OK; }
    const int v[13] = {2, p.cs, p.cblocks, p.kt_full, p.ktiles, p.SW, p.SH, p.slab_bytes, p.nslab, p.nstages, p.nbox, p.smem, p.mt};
    for (int i = 0; i < 13; ++i) out16[i] = v[i];
    return CCB_OK;
}

extern "C" int ccb_debug_tma_status(unsigned int* out4) {
    if (!out4) return CCB_ERR_ARG;
    unsigned int zero[4] = {0, 0, 0, 0};
    if (cudaDeviceSynchronize() != cudaSuccess) return CCB_ERR_LAUNCH;
    if (cudaMemcpyFromSymbol(out4, ccb::g_tma_status, sizeof(zero)) != cudaSuccess) return CCB_ERR_LAUNCH;
    if (cudaMemcpyToSymbol(ccb::g_tma_status, zero, sizeof(zero)) != cudaSuccess) return CCB_ERR_LAUNCH;
    return CCB_OK;
}

#else

namespace ccb {
void tma_set_enabled(int) {}
bool tma_conv_supported(const ccb_conv_desc*, int) { return false; }
bool tma_direct_fprop(const ccb_conv_desc*) { return false; }
bool tma_nhwc_takes(const ccb_conv_desc*, int) { return false; }
long long tma_workspace_floats(const ccb_conv_desc*, int) { return 0; }
int tma_fprop(const ccb_conv_desc*, const float*, const float*, const float*, const float*, float*, float*, long long, int,
              cudaStream_t) { return CCB_ERR_UNSUPPORTED; }
int tma_dgrad(const ccb_conv_desc*, const float*, const float*, const float*, const float*, float*, float*, long long, int,
              cudaStream_t) { return CCB_ERR_UNSUPPORTED; }
bool tma_wgrad_supported(const ccb_conv_desc*) { return false; }
long long tma_wgrad_workspace_floats(const ccb_conv_desc*) { return 0; }
int tma_wgrad(const ccb_conv_desc*, const float*, const float*, float*, float*, long long, int, cudaStream_t) {
    return CCB_ERR_UNSUPPORTED;
}
}  // namespace ccb
extern "C" int ccb_debug_conv_plan(const ccb_conv_desc*, int, int, int, int* out16) {
    if (out16) out16[0] = -1;
    return CCB_OK;
}
extern "C" int ccb_debug_tma_status(unsigned int* out4) {
    if (out4) out4[0] = out4[1] = out4[2] = out4[3] = 0;
    return CCB_OK;
}

#endif
