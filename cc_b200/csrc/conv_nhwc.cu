// conv_nhwc.cu - channels-last slab kernel: TMA-fed tcgen05 implicit-GEMM convolution with NO per-tap data movement.
//
//   The stride-1 gathers (every stride-1 FPROP, every DGRAD parity class, ConvTranspose2d forward) first get a
//   channels-last copy xh[B][H][W][Cp] of the gathered tensor (one streaming transposition, nchw_to_nhwc_kernel).  In that
//   layout a pixel is one 128-byte row of 32 channels, so
//     * the input patch under a 16 x 8 output tile, halo included, is ONE TMA box per 32-channel block
//       (32 ch x SW px x SH rows, SW = 8 + halo, SWIZZLE_128B, zero padding / image borders / channel tails = TMA out-of-bounds fill):
//       the "slab", [SH][SW][32 ch], 128 bytes per pixel;
//     * the K-major A operand of filter tap (dy, dx) IS the slab at byte offset ((dy - oy_lo) * SW + (dx - ox_lo)) * 128:
//       UMMA rows m = 8 * (tile row) + (tile column) -> 8-row groups one slab row (SW x 128 B) apart (SBO), rows 128 B apart.  A tap is a
//       descriptor start address - nothing is copied, cut or re-laid-out per tap.  (Measured on the B200,
//       tools/nhwc_probe.py: the tensor core applies the 128-byte swizzle to the ABSOLUTE shared-memory address, exactly
//       like the TMA that wrote the slab, so a start address that is only 128-byte aligned needs NO base-offset field;
//       setting the field to (addr >> 7) & 7 as the PTX text suggests computes garbage for every unaligned tap.)
//     * 3xTF32: lo = x - trunc_tf32(x) is computed ONCE per slab (elementwise, layout agnostic) by eight warps, not per tap.
//   Per k-stage (one tap x 32 channels) the issuing thread runs 4 k-steps of
//        D[:, 0:2n)   += A_hi * [B_hi | B_lo]        D[:, 2n:3n) += A_lo * B_hi
//   with the weight tiles (K-major, hi | lo) arriving by TMA in a ring.  The epilogue adds the three partial sums in
//   fp32, applies bias / residual / activation and writes NCHW.
//
//   warp 0: TMA producer (one thread)   warp 1: MMA issuer + TMEM owner   warps 2-9: lo pass, then epilogue
#include "tc_common.cuh"

#ifndef CCB_CPU_SIM
namespace ccb {

constexpr int NH_THREADS = 320;
constexpr int NH_TH = 16, NH_TW = 8, NH_SW_MAX = 16;  // output tile rows x columns; slab columns = 8 + halo <= 16
constexpr int NH_MAX_TAPS = 64;

struct NhwcArgs {
    int B, Hin, Win;
    int Ntot, Hout, Wout, Hc, Wc;
    int out_stride, out_oy, out_ox;
    int ntaps, cblocks;
    int SH, SW, mt, ox_lo, oy_lo, slab_bytes, slab_tx, nslab;   // slab rows / columns (pixels), stacked 16-row tiles per CTA
    int tiles_x, tiles_y, splits, cb_per_split;
    long long out_numel;
    float* partial;
    const float* bias;
    const float* res;
    float* out;
    int act;
    float slope;
    int nstages, nbox, ntile_w, tmem_cols, soft, dbg;    // ntile_w: output channels per CTA (64 or 128)
    signed char off_y[NH_MAX_TAPS], off_x[NH_MAX_TAPS];
};

// first barrier time-out of a launch: {role, iteration, blockIdx.x, blockIdx.z}; role 0 = none
__device__ unsigned int g_nhwc_status[4];
__device__ __noinline__ void nh_wait_failed(int soft, int role, int it) {
    if (!soft) asm volatile("trap;");
    if (atomicCAS(&g_nhwc_status[0], 0u, (unsigned)role) == 0u) {
        g_nhwc_status[1] = (unsigned)it;
        g_nhwc_status[2] = blockIdx.x;
        g_nhwc_status[3] = blockIdx.z;
    }
}
// every wait is bounded: a protocol bug fails the launch (or, with the bring-up switch, is recorded) instead of hanging the GPU
__device__ __forceinline__ void nh_wait(uint64_t* bar, uint32_t parity, int soft, int role, int it) {
    uint32_t spins = 0;
    while (!tm_mbar_try(bar, parity)) {
        ++spins;
        if (soft && (spins & 1023u) == 0 && *(volatile unsigned int*)&g_nhwc_status[0] != 0u) return;
        if (spins > (soft ? (1u << 20) : (1u << 26))) { nh_wait_failed(soft, role, it); return; }
    }
}

template <bool THREE>
__global__ void __launch_bounds__(NH_THREADS, 2)
conv_nhwc_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_b, const NhwcArgs a) {
    CCB_PDL_TRIGGER();
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_addr(smem_raw) & 1023u)) & 1023u);
    const int NST = a.nstages;
    const int b_half = a.nbox * 128;                              // one weight copy of one stage: nbox rows x 32 k
    const int b_stage = (THREE ? 2 : 1) * b_half;
    unsigned char* slab_hi = smem;                                // [nslab][slab_bytes]
    unsigned char* slab_lo = slab_hi + a.nslab * a.slab_bytes;    // [nslab][slab_bytes] (3xTF32 only)
    unsigned char* b_ring = slab_lo + (THREE ? a.nslab * a.slab_bytes : 0);
    uint64_t* bars = (uint64_t*)(b_ring + NST * b_stage);
    uint64_t* b_full = bars;              // [8] TMA weights
    uint64_t* b_empty = bars + 8;         // [8] tcgen05.commit
    uint64_t* slab_full = bars + 16;      // [2] TMA slab
    uint64_t* lo_full = bars + 18;        // [2] 8 lo-pass warps
    uint64_t* slab_empty = bars + 20;     // [2] tcgen05.commit
    uint64_t* accum_bar = bars + 22;
    uint32_t* tmem_slot = (uint32_t*)(bars + 23);
    int* toff = (int*)(bars + 24);        // [64] slab pixel offset of each tap

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    int t = blockIdx.x;
    const int per_b = a.tiles_x * a.tiles_y;
    const int b = t / per_b;
    t -= b * per_b;
    const int ty = t / a.tiles_x, tx = t - ty * a.tiles_x;
    const int x0 = tx * NH_TW, y0 = ty * NH_TH * a.mt;             // a CTA owns mt vertically stacked 16 x 8 tiles (one slab)
    const int n0 = blockIdx.y * a.ntile_w;
    const int ntile = min(a.ntile_w, a.Ntot - n0);
    const int cb_beg = blockIdx.z * a.cb_per_split;
    const int cb_end = min(a.cblocks, cb_beg + a.cb_per_split);
    const int nblocks = max(0, cb_end - cb_beg);

    if (tid >= 64 && tid < 64 + NH_MAX_TAPS) {
        const int tp = tid - 64;
        toff[tp] = (tp < a.ntaps) ? (a.off_y[tp] - a.oy_lo) * a.SW + (a.off_x[tp] - a.ox_lo) : 0;
    }
    if (tid == 0) {
        tm_prefetch_map(&map_x);
        tm_prefetch_map(&map_b);
        for (int s = 0; s < NST; ++s) {
            tm_mbar_init(&b_full[s], 1);
            tm_mbar_init(&b_empty[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            tm_mbar_init(&slab_full[s], 1);
            tm_mbar_init(&lo_full[s], 8);
            tm_mbar_init(&slab_empty[s], 1);
        }
        tm_mbar_init(accum_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(tmem_slot)), "r"((uint32_t)a.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    CCB_PDL_SYNC();                                               // everything above touched no global data

    if (warp == 0) {
        // ===================== TMA producer (one thread) =====================
        if (lane == 0 && nblocks > 0) {
            int s = 0, ph = 0, issued = 0;                         // weight ring slot / phase / stages requested so far
            int sl_next = 0;                                        // slabs requested so far
            auto load_slab = [&](int it) {
                const int sb = sl_next % a.nslab;
                if (sl_next >= a.nslab) nh_wait(&slab_empty[sb], ((sl_next / a.nslab) - 1) & 1, a.soft, 6, it);
                tm_mbar_expect_tx(&slab_full[sb], (uint32_t)a.slab_tx);
                tma_load_4d(slab_hi + sb * a.slab_bytes, &map_x, &slab_full[sb], (cb_beg + sl_next) * 32, x0 + a.ox_lo, y0 + a.oy_lo, b);
                ++sl_next;
            };
            load_slab(0);
            // the slab of the NEXT channel block is requested a few weight tiles into the current block: the wait for its
            // buffer (the block before this one) then never starves the MMA warp of weight tiles.  With one slab buffer
            // the request has to follow the block's last weight tile (its MMAs free the buffer).
            const int t_pre = (a.nslab > 1) ? min(a.ntaps - 1, NST - 1) : a.ntaps;
            for (int blk = 0; blk < nblocks; ++blk) {
                const int cb = cb_beg + blk;
                for (int tp = 0; tp <= a.ntaps; ++tp) {
                    if (tp == t_pre && blk + 1 < nblocks) load_slab(issued);
                    if (tp == a.ntaps) break;
                    if (issued >= NST) nh_wait(&b_empty[s], ph ^ 1, a.soft, 1, issued);
                    unsigned char* bt = b_ring + s * b_stage;
                    tm_mbar_expect_tx(&b_full[s], (uint32_t)b_stage);
                    const int k0 = (cb * a.ntaps + tp) * 32;
                    tma_load_2d(bt, &map_b, &b_full[s], k0, n0);
                    if (THREE) tma_load_2d(bt + b_half, &map_b, &b_full[s], k0, a.Ntot + n0);
                    ++issued;
                    if (++s == NST) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer: the whole warp runs the loop, lane 0's predicate issues =====================
        // D fp32, A/B tf32, both K-major, M = 128
        const uint32_t idesc_base = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t idesc_n1 = idesc_base | ((uint32_t)(a.nbox >> 3) << 17);
        const uint32_t idesc_n2 = idesc_base | ((uint32_t)((2 * a.nbox) >> 3) << 17);
        const uint32_t leader = (lane == 0) ? 1u : 0u;
        // A: 8-row groups one slab row apart; B: 8-row groups 1024 B apart; both SWIZZLE_128B, K-major
        const uint32_t a_desc_hi = ((uint32_t)(a.SW * 128) >> 4) | (1u << 14) | (2u << 29);
        const uint32_t tile16 = (uint32_t)(NH_TH * a.SW * 8);       // one stacked tile further down the slab, in 16-byte units
        int s = 0;
        uint32_t ph = 0;
        int it = 0;
        for (int blk = 0; blk < nblocks; ++blk) {
            const int sb = blk % a.nslab;
            const uint32_t sph = (uint32_t)((blk / a.nslab) & 1);
            nh_wait(&slab_full[sb], sph, a.soft, 2, it);
            if (THREE) nh_wait(&lo_full[sb], sph, a.soft, 3, it);
            const uint32_t hi16 = smem_addr(slab_hi + sb * a.slab_bytes) >> 4;
            const uint32_t lo16 = smem_addr(slab_lo + sb * a.slab_bytes) >> 4;
            for (int tp = 0; tp < a.ntaps; ++tp, ++it) {
                nh_wait(&b_full[s], ph, a.soft, 4, it);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t pix = (uint32_t)toff[tp];
                // no base-offset field: the swizzle is a function of the absolute address (bring-up bit 0 sets it: must break)
                const uint32_t a_hi_word = a_desc_hi | ((a.dbg & 1) ? ((pix & 7u) << 17) : 0u);
                const uint32_t b16 = smem_addr(b_ring + s * b_stage) >> 4;
                for (int m = 0; m < a.mt; ++m) {
                    const uint32_t d0 = tmem_base + (uint32_t)(m * (THREE ? 3 : 1) * a.nbox), d1 = d0 + (uint32_t)(2 * a.nbox);
                    const uint32_t ah = hi16 + pix * 8u + m * tile16, al = lo16 + pix * 8u + m * tile16;
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        const uint32_t acc = (it > 0 || ks > 0) ? 1u : 0u;
                        tm_umma_tf32_p(d0, ((ah + ks * 2u) & 0x3FFFu) | DESC_K_LO, a_hi_word, (b16 + ks * 2) | DESC_K_LO, DESC_K_HI,
                                       THREE ? idesc_n2 : idesc_n1, acc, leader);
                        if (THREE)
                            tm_umma_tf32_p(d1, ((al + ks * 2u) & 0x3FFFu) | DESC_K_LO, a_hi_word, (b16 + ks * 2) | DESC_K_LO, DESC_K_HI,
                                           idesc_n1, acc, leader);
                    }
                }
                tm_commit_p(&b_empty[s], leader);
                if (++s == NST) { s = 0; ph ^= 1; }
            }
            tm_commit_p(&slab_empty[sb], leader);
        }
        if (nblocks > 0) tm_commit_p(accum_bar, leader);
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    } else {
        // ===================== lo pass (once per slab), then the epilogue: 8 warps =====================
        const int wt = tid - 64;                                   // 0 .. 255
        if (THREE) {
            const int nvec = a.slab_tx >> 4;
            for (int blk = 0; blk < nblocks; ++blk) {
                const int sb = blk % a.nslab;
                nh_wait(&slab_full[sb], (uint32_t)((blk / a.nslab) & 1), a.soft, 5, blk);
                const float4* raw = (const float4*)(slab_hi + sb * a.slab_bytes);
                float4* lo = (float4*)(slab_lo + sb * a.slab_bytes);
                for (int i = wt; i < nvec; i += 256) lo[i] = tf32_rest4(raw[i]);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) tm_mbar_arrive(&lo_full[sb]);
            }
        }
        if (nblocks > 0) nh_wait(accum_bar, 0, a.soft, 7, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // TMEM lane m = 8 * (tile row) + (tile column): a warp's lane quarter is 4 tile rows x 8 columns
        const int q4 = warp & 3, colh = (warp - 2) >> 2;           // 2 warps per TMEM lane quarter split the column groups
        const int ox = x0 + (lane & 7);
        const long long HWout = (long long)a.Hout * a.Wout;
        for (int m = 0; m < a.mt; ++m) {
        const int oy = y0 + m * NH_TH + q4 * 4 + (lane >> 3);
        const bool evalid = (oy < a.Hc) && (ox < a.Wc);
        const long long obase = (long long)b * a.Ntot * HWout + (long long)(oy * a.out_stride + a.out_oy) * a.Wout +
                                (ox * a.out_stride + a.out_ox);
        const uint32_t trow = tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(m * (THREE ? 3 : 1) * a.nbox);
        for (int cg = colh; cg * 16 < ntile; cg += 2) {
            float v[16];
            if (nblocks > 0) {
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
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)a.tmem_cols) : "memory");
    }
}

// NCHW [B][C][HW] -> channels-last [B][HW][Cp] (Cp >= C, a multiple of 4: channel tails are zero), 32 x 32 tiles through
// shared memory: reads coalesced along pixels, writes coalesced along channels
__global__ void __launch_bounds__(256) nchw_to_nhwc_kernel(const float* __restrict__ x, float* __restrict__ xh, int C, int Cp, int HW) {
    CCB_PDL_WAIT();
    __shared__ float tile[32][33];
    const int b = blockIdx.z, c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const float* xb = x + (long long)b * C * HW;
    float* ob = xh + (long long)b * HW * Cp;
#pragma unroll
    for (int j = ty; j < 32; j += 8) {
        const int c = c0 + j, p = p0 + tx;
        tile[j][tx] = (c < C && p < HW) ? __ldg(xb + (long long)c * HW + p) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int j = ty; j < 32; j += 8) {
        const int p = p0 + j, c = c0 + tx;
        if (p < HW && c < Cp) ob[(long long)p * Cp + c] = tile[tx][j];
    }
}

static int g_nhwc_enabled = 1, g_nhwc_soft = 0, g_nhwc_dbg = 0;

struct NhwcPlan {
    int Cp, cblocks, SH, SW, mt, ox_lo, oy_lo, slab_bytes, slab_tx, nslab, nstages, nbox, ntile_w, tmem_cols, smem, Kp;
    bool ok;
};
static int g_nhwc_mt = 0;          // bring-up: force the number of stacked tiles (0 = planned)
static int g_nhwc_wgrad_on = 0;    // dbg bit 7 sends stride-1 weight gradients through conv_wgrad_nhwc_kernel (off by default, see there)
void nhwc_set_debug(int enabled, int soft, int dbg) { g_nhwc_enabled = enabled; g_nhwc_soft = soft; g_nhwc_dbg = dbg & 15; g_nhwc_mt = (dbg >> 4) & 7; g_nhwc_wgrad_on = (dbg >> 7) & 1; }
// tap offsets (pixels of the gathered tensor), channels, output channels, output grid -> tiling; !ok: the path does not take
// this problem.
//   * Two CTAs per SM (<= 111 KB of shared memory, <= 256 TMEM columns each) whenever the problem allows: prologue
//     (barriers, TMEM, first slab) and epilogue (TMEM -> registers -> NCHW stores) of one CTA then hide under the main loop
//     of the other; measured (ncu, 128 -> 128 3x3): tensor pipe 45 % with one CTA per SM, MMA floor 14 us of a 37 us tile.
//     More than 64 output channels are cut into 64-wide CTAs for it (grid.y).
//   * Stacked tiles (mt > 1, one taller slab, mt accumulator sets in TMEM) reuse every weight tile mt times; used for
//     thin layers on large maps when everything still fits.
static NhwcPlan nhwc_plan(const int* off_y, const int* off_x, int ntaps, int Cc, int N, int three, int Hc = 0, long long tiles1 = 0) {
    NhwcPlan p;
    memset(&p, 0, sizeof(p));
    if (ntaps < 1 || ntaps > NH_MAX_TAPS) return p;
    int oxl = off_x[0], oxh = off_x[0], oyl = off_y[0], oyh = off_y[0];
    for (int t = 1; t < ntaps; ++t) {
        oxl = off_x[t] < oxl ? off_x[t] : oxl; oxh = off_x[t] > oxh ? off_x[t] : oxh;
        oyl = off_y[t] < oyl ? off_y[t] : oyl; oyh = off_y[t] > oyh ? off_y[t] : oyh;
    }
    p.SW = NH_TW + (oxh - oxl);
    if (p.SW > NH_SW_MAX) return p;                               // halo wider than the widest slab
    p.ox_lo = oxl; p.oy_lo = oyl;
    p.Cp = (Cc + 3) & ~3;
    p.cblocks = cdiv(Cc, 32);
    p.Kp = p.cblocks * ntaps * 32;
    int mt_max = 1;
    if (g_nhwc_mt > 0) mt_max = g_nhwc_mt;
    else
        for (int mt = 2; mt <= 4; mt *= 2)
            if (Hc >= NH_TH * mt && tiles1 / mt >= 4 * 148) mt_max = mt;
    // candidate configurations in order of preference: two CTAs per SM first
    for (int pass = (g_nhwc_dbg & 8) ? 1 : 0; pass < 2 && !p.ok; ++pass) {
        const bool two = (pass == 0);
        const int ntile_w = two ? 64 : 128;
        const int ntile = N < ntile_w ? N : ntile_w;
        const int nbox = (ntile + 15) & ~15;
        const int b_stage = (three ? 2 : 1) * nbox * 128;
        const int total = (two ? 111 : 224) * 1024 - 2048;        // barriers + tap table + alignment slack
        for (int mt = mt_max; mt >= 1 && !p.ok; mt >>= 1) {
            if (mt * (three ? 3 : 1) * nbox > (two ? 256 : 512)) continue;
            const int SH = NH_TH * mt + (oyh - oyl);
            if (SH > 256) continue;
            const int tx = SH * p.SW * 128;
            const int sbytes = (tx + 1023) & ~1023;               // slab bases stay 1024-byte aligned (TMA swizzle atom)
            for (int nslab = (p.cblocks > 1 ? 2 : 1); nslab >= 1 && !p.ok; --nslab) {
                const int slabs = (three ? 2 : 1) * nslab * sbytes;
                int nst = (total - slabs) / b_stage;
                if (nst > 8) nst = 8;
                if (nst >= 3 || (nst >= 2 && !two && mt == 1)) {
                    p.mt = mt; p.SH = SH; p.slab_tx = tx; p.slab_bytes = sbytes; p.nslab = nslab; p.nstages = nst; p.nbox = nbox;
                    p.ntile_w = ntile_w; p.tmem_cols = two ? 256 : 512;
                    p.smem = slabs + nst * b_stage + 2048 + 1024;
                    p.ok = true;
                }
            }
        }
    }
    return p;
}

// Does the channels-last kernel take this gather?  (measured dispatch: everything with >= 32 gathered channels whose
// output grid fills 16 x 8 tiles reasonably)
// shape conditions of the channels-last kernel (no driver needed: the CPU plan tests call this too)
static bool nhwc_shape_ok(const int* off_y, const int* off_x, int ntaps, int in_stride, int Cc, int N, int Hc, int Wc, bool part_of_set) {
    if (!g_nhwc_enabled || in_stride != 1) return false;
    if (Cc < ((g_nhwc_dbg & 2) ? 16 : 32) || N < 16) return false;
    if (ntaps == 1 && off_x[0] == 0 && off_y[0] == 0 && !part_of_set && !(g_nhwc_dbg & 4)) return false;   // 1x1: the aligned NCHW TMA kernel needs no copy
    // maps of 8 rows would half-fill the 16-row tile: measured (512 -> 512 3x3 at 8x26) 89 us here against 72 us on the NCHW
    // slab kernel with split-K, 179 against 126 us for 1024 -> 512: they stay there
    if (Hc < 12 || Wc < 7) return false;
    if ((long long)cdiv(Hc, NH_TH) * NH_TH * cdiv(Wc, NH_TW) * NH_TW * 10 > (long long)Hc * Wc * 14) return false;   // > 40 % tile waste
    return nhwc_plan(off_y, off_x, ntaps, Cc, N, 1).ok;
}
// `part_of_set`: one parity class of a strided DGRAD whose other classes need the channels-last copy anyway
bool nhwc_applies(const int* off_y, const int* off_x, int ntaps, int in_stride, int Cc, int N, int Hc, int Wc, int three, bool part_of_set) {
    (void)three;
    return get_encode() != nullptr && nhwc_shape_ok(off_y, off_x, ntaps, in_stride, Cc, N, Hc, Wc, part_of_set);
}
// host-side tiling for the plan tests: out16 = {6, cblocks, SH, SW, mt, slab_bytes, slab_tx, nslab, nstages, nbox, ntile_w, tmem_cols,
// smem, Kp, Cp, grid CTAs}; out16[0] = -1 when the shape is not taken
void nhwc_debug_plan(const int* off_y, const int* off_x, int ntaps, int in_stride, int Cc, int N, int B, int Hc, int Wc, bool part_of_set,
                     int* out16) {
    out16[0] = -1;
    if (!nhwc_shape_ok(off_y, off_x, ntaps, in_stride, Cc, N, Hc, Wc, part_of_set)) return;
    const long long tiles1 = (long long)B * cdiv(Wc, NH_TW) * cdiv(Hc, NH_TH) * cdiv(N, 128);
    const NhwcPlan p = nhwc_plan(off_y, off_x, ntaps, Cc, N, 1, Hc, tiles1);
    if (!p.ok) return;
    const int v[16] = {6, p.cblocks, p.SH, p.SW, p.mt, p.slab_bytes, p.slab_tx, p.nslab, p.nstages, p.nbox, p.ntile_w, p.tmem_cols, p.smem, p.Kp,
                       p.Cp, (int)(B * cdiv(Wc, NH_TW) * cdiv(Hc, NH_TH * p.mt) * cdiv(N, p.ntile_w))};
    for (int i = 0; i < 16; ++i) out16[i] = v[i];
}
bool nhwc_prefers_thin() { return (g_nhwc_dbg & 2) != 0; }
long long nhwc_copy_floats(int B, int Cc, int Hin, int Win) { return (long long)B * Hin * Win * ((Cc + 3) & ~3); }
long long nhwc_wp_floats(const int* off_y, const int* off_x, int ntaps, int Cc, int N) {
    const NhwcPlan p = nhwc_plan(off_y, off_x, ntaps, Cc, N, 1);
    return p.ok ? 2ll * N * p.Kp : -1;
}

int nhwc_transpose(const float* x, float* xh, int B, int Cc, int Hin, int Win, cudaStream_t st) {
    const int Cp = (Cc + 3) & ~3, HW = Hin * Win;
    CCB_REQUIRE(B <= 65535 && cdiv(Cp, 32) <= 65535, CCB_ERR_ARG, "nhwc_transpose: grid too large");
    CCB_LAUNCH(nchw_to_nhwc_kernel, dim3(cdiv(HW, 32), cdiv(Cp, 32), B), dim3(256), 0, st, x, xh, Cc, Cp, HW);
    return check_launch("nchw_to_nhwc");
}

// One launch: channels-last gathered tensor xh [B][Hin][Win][Cp] -> output grid (Hc x Wc) written NCHW with stride / offset.
int launch_nhwc(const float* xh, int B, int Cc, int Hin, int Win, const float* w, int mode, int N, int KK, int Ci, const int* off_y,
                const int* off_x, const int* tap_index, int ntaps, int Hc, int Wc, int Hout, int Wout, int out_stride, int out_oy,
                int out_ox, const float* bias, const float* res, float* out, int act, float slope, int three, float* work,
                long long wp_floats, int splits, float* partial, long long out_numel, cudaStream_t st) {
    EncodeTiledFn enc = get_encode();
    CCB_REQUIRE(enc != nullptr, CCB_ERR_UNSUPPORTED, "conv_nhwc: cuTensorMapEncodeTiled unavailable");
    const long long tiles1 = (long long)B * cdiv(Wc, NH_TW) * cdiv(Hc, NH_TH) * cdiv(N, 128);
    const NhwcPlan p = nhwc_plan(off_y, off_x, ntaps, Cc, N, three, splits > 1 ? 0 : Hc, tiles1);
    CCB_REQUIRE(p.ok, CCB_ERR_UNSUPPORTED, "conv_nhwc: no tiling");
    CCB_REQUIRE((((uintptr_t)xh) & 15) == 0, CCB_ERR_ARG, "conv_nhwc: unaligned channels-last copy");
    NhwcArgs a;
    memset(&a, 0, sizeof(a));
    a.B = B; a.Hin = Hin; a.Win = Win; a.Ntot = N; a.Hout = Hout; a.Wout = Wout; a.Hc = Hc; a.Wc = Wc;
    a.out_stride = out_stride; a.out_oy = out_oy; a.out_ox = out_ox;
    a.ntaps = ntaps; a.cblocks = p.cblocks; a.SH = p.SH; a.SW = p.SW; a.mt = p.mt; a.ox_lo = p.ox_lo; a.oy_lo = p.oy_lo;
    a.slab_bytes = p.slab_bytes; a.slab_tx = p.slab_tx; a.nslab = p.nslab;
    a.tiles_x = cdiv(Wc, NH_TW); a.tiles_y = cdiv(Hc, NH_TH * p.mt);
    if (splits > p.cblocks) splits = p.cblocks;
    if (splits < 1) splits = 1;
    a.splits = splits; a.cb_per_split = cdiv(p.cblocks, splits);
    a.out_numel = out_numel; a.partial = partial;
    a.bias = bias; a.res = res; a.out = out; a.act = act; a.slope = slope; a.soft = g_nhwc_soft; a.dbg = g_nhwc_dbg;
    a.nstages = p.nstages; a.nbox = p.nbox; a.ntile_w = p.ntile_w; a.tmem_cols = p.tmem_cols;
    WPrepDesc pa;
    memset(&pa, 0, sizeof(pa));
    for (int t = 0; t < NH_MAX_TAPS; ++t) {
        a.off_y[t] = (signed char)(t < ntaps ? off_y[t] : 0);
        a.off_x[t] = (signed char)(t < ntaps ? off_x[t] : 0);
        pa.tap_index[t] = (signed char)(t < ntaps ? tap_index[t] : 0);
    }
    CCB_REQUIRE(2ll * N * p.Kp <= wp_floats, CCB_ERR_ARG, "conv_nhwc: workspace too small");
    pa.w = w; pa.wp = work; pa.N = N; pa.Cc = Cc; pa.KK = KK; pa.Ci = Ci; pa.mode = mode; pa.Kp = p.Kp; pa.ntaps = ntaps;
    pa.layout = WPREP_NHWC; pa.p0 = 32; pa.p1 = p.cblocks; pa.p2 = ntaps;
    const float* wpp = nullptr;
    int rc = wprep_get(pa, st, &wpp);
    if (rc) return rc;
    alignas(64) CUtensorMap map_x, map_b;
    {
        // xh as (C, W, H, B); box (32 ch, 16 px, SH rows, 1): out-of-range coordinates (padding, borders, channel tail) read zeros
        cuuint64_t dims[4] = {(cuuint64_t)p.Cp, (cuuint64_t)Win, (cuuint64_t)Hin, (cuuint64_t)B};
        cuuint64_t strides[3] = {(cuuint64_t)p.Cp * 4, (cuuint64_t)Win * p.Cp * 4, (cuuint64_t)Hin * Win * p.Cp * 4};
        cuuint32_t box[4] = {32, (cuuint32_t)p.SW, (cuuint32_t)p.SH, 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = enc(&map_x, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)xh, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        CCB_REQUIRE(r == CUDA_SUCCESS, CCB_ERR_LAUNCH, "conv_nhwc: cuTensorMapEncodeTiled(x) failed (%d)", (int)r);
    }
    {
        cuuint64_t dims[2] = {(cuuint64_t)p.Kp, (cuuint64_t)(2 * N)};
        cuuint64_t strides[1] = {(cuuint64_t)p.Kp * 4};
        cuuint32_t box[2] = {32, (cuuint32_t)p.nbox};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = enc(&map_b, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)wpp, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        CCB_REQUIRE(r == CUDA_SUCCESS, CCB_ERR_LAUNCH, "conv_nhwc: cuTensorMapEncodeTiled(w) failed (%d)", (int)r);
    }
    dim3 grid(B * a.tiles_x * a.tiles_y, cdiv(N, p.ntile_w), splits);
    auto kfn = three ? conv_nhwc_kernel<true> : conv_nhwc_kernel<false>;
    cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    CCB_LAUNCH(kfn, grid, dim3(NH_THREADS), p.smem, st, map_x, map_b, a);
    return check_launch("conv_nhwc");
}


// ---------------------------------------------------------------------------------------------------------------
// Channels-last WEIGHT GRADIENT:  dw[co][ci][tap] = sum over pixels of dz[px][co] * x[px + tap][ci]   (stride 1)
//   GEMM with K = pixels.  In NHWC both operands are "K rows of 128 bytes" (a pixel x 32 channels): MN-major operands
//   exactly in the layout the aligned NCHW kernel (conv_tma_kernel) already feeds the tensor core with
//   (SWIZZLE_128B_ATOM_32B, 4-row atoms 512 B apart, 32-wide MN blocks 4096 B apart) - and, because a CTA owns whole
//   (tap, 32-channel block) pairs, the tap shift is a TMA COORDINATE, not a shared-memory offset: no cutter warps at all
//   (the NCHW slab kernel spends its time cutting the (tap, ci) x 32 px operand out of a slab, tensor pipe 30 %).
//     M = 128 = 4 (tap, ci-block) pairs: four TMA boxes (32 ci, 8 px, 4 rows) of x at the tap-shifted coordinate
//     N = the co tile (<= 128, in 32-wide boxes of dz), doubled by the tf32 remainder: D[:, 0:2n) += A_hi * [B_hi | B_lo],
//         D[:, 2n:3n) += A_lo * B_hi
//     K = 32 pixels (4 rows x 8 columns) per stage, 4 k-steps of 8; image borders / padding / channel tails = TMA zero fill
//   grid = (pairs / 4, co tiles, pixel splits); split partials are summed in a fixed order.
//   warp 0: TMA producer   warp 1: MMA issuer   warps 2-9: remainder pass + epilogue
//   MEASURED (B200, 128 -> 128 3x3, b4 64x208; profiles/r02_ncu_wgrad_nhwc_128.txt): correct to 2e-6, tensor pipe 42.7 % active -
//   but 193 us against 168 us for the NCHW slab kernel: MN-major tf32 MMAs keep the tensor pipe busy 2.1 x the nominal
//   M x N x K / 2048 cycles (82 us of pipe time for 39 us of nominal work), so removing the cutters does not pay here.
//   The dispatcher therefore keeps weight gradients on conv_slab_wgrad_kernel (K-major operands: pixels contiguous in
//   NCHW); this kernel stays selectable (ccb_debug_nhwc dbg bit 7) and is covered by the parity tests.
// ---------------------------------------------------------------------------------------------------------------
struct NhwcWgradArgs {
    int B, Ci, Co, Ho, Wo, KK, kw, pad;
    int npairs, ciblocks;
    int tiles_x, tiles_y, stages, per_split, splits;
    int nboxp, nblkN;                    // co tile rounded to 32, number of 32-wide dz boxes
    int nstages, soft;
    long long numel;
    float* out;
};
constexpr int NW_BLK = 4096;             // one 32-wide MN block of one stage: 32 px x 128 B

template <bool THREE>
__global__ void __launch_bounds__(NH_THREADS, 1)
conv_wgrad_nhwc_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_dz, const NhwcWgradArgs a) {
    CCB_PDL_WAIT();
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_addr(smem_raw) & 1023u)) & 1023u);
    // Two rings: the TMA-filled operand tiles (A_hi | B_hi, NST deep: the loads run far enough ahead to hide the L2 latency)
    // and the tf32 remainders (A_lo | B_lo, NSTL = 2 deep: produced by the eight remainder warps just ahead of the MMAs)
    const int NST = a.nstages;
    constexpr int NSTL = 2;
    const int a_bytes = 4 * NW_BLK, b_bytes = a.nblkN * NW_BLK;
    const int hi_bytes = a_bytes + b_bytes;                                   // [A_hi][B_hi]
    unsigned char* lo_ring = smem + NST * hi_bytes;                            // [A_lo][B_lo] x NSTL
    uint64_t* bars = (uint64_t*)(lo_ring + (THREE ? NSTL * hi_bytes : 0));
    uint64_t* full = bars;                // [8] TMA
    uint64_t* empty = bars + 8;           // [8] tcgen05.commit: operand stage consumed
    uint64_t* lo_full = bars + 16;        // [2] 8 remainder warps
    uint64_t* lo_empty = bars + 18;       // [2] tcgen05.commit: remainder stage consumed
    uint64_t* accum_bar = bars + 20;
    uint32_t* tmem_slot = (uint32_t*)(bars + 21);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int p0 = blockIdx.x * 4;                                            // first (tap, ci-block) pair of this CTA
    const int n0 = blockIdx.y * 128;
    const int st_beg = blockIdx.z * a.per_split;
    const int nst = max(0, min(a.stages, st_beg + a.per_split) - st_beg);
    const int per_b = a.tiles_x * a.tiles_y;

    if (tid == 0) {
        tm_prefetch_map(&map_x);
        tm_prefetch_map(&map_dz);
        for (int s = 0; s < NST; ++s) {
            tm_mbar_init(&full[s], 1);
            tm_mbar_init(&empty[s], 1);
        }
        for (int s = 0; s < NSTL; ++s) {
            tm_mbar_init(&lo_full[s], 8);
            tm_mbar_init(&lo_empty[s], 1);
        }
        tm_mbar_init(accum_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer (one thread) =====================
        if (lane == 0) {
            const int npv = min(4, a.npairs - p0);                             // valid pairs: the other rows are never read back
            int tapy[4], tapx[4], cch[4];
            for (int j = 0; j < 4; ++j) {
                const int p = min(p0 + j, a.npairs - 1);
                const int tap = p / a.ciblocks;
                cch[j] = (p - tap * a.ciblocks) * 32;
                tapy[j] = tap / a.kw - a.pad;
                tapx[j] = tap - (tap / a.kw) * a.kw - a.pad;
            }
            const uint32_t tx_bytes = (uint32_t)(npv * NW_BLK + b_bytes);
            int s = 0, ph = 0;
            for (int it = 0; it < nst; ++it) {
                if (it >= NST) nh_wait(&empty[s], ph ^ 1, a.soft, 1, it);
                int t = st_beg + it;
                const int b = t / per_b;
                t -= b * per_b;
                const int ty = t / a.tiles_x, tx = t - ty * a.tiles_x;
                const int x0 = tx * 8, y0 = ty * 4;
                unsigned char* st = smem + s * hi_bytes;
                tm_mbar_expect_tx(&full[s], tx_bytes);
                for (int j = 0; j < npv; ++j) tma_load_4d(st + j * NW_BLK, &map_x, &full[s], cch[j], x0 + tapx[j], y0 + tapy[j], b);
                for (int j = 0; j < a.nblkN; ++j) tma_load_4d(st + a_bytes + j * NW_BLK, &map_dz, &full[s], n0 + j * 32, x0, y0, b);
                if (++s == NST) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        // D fp32, A/B tf32, BOTH MN-major, M = 128.  hi*hi and hi*lo are two MMAs here (B_hi and B_lo live in different
        // rings), lo*hi the third; hi*lo and lo*hi share one accumulator, the main product keeps its own.
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(128 >> 4) << 24) |
                               ((uint32_t)(a.nboxp >> 3) << 17);
        const uint32_t leader = (lane == 0) ? 1u : 0u;
        const uint32_t d0 = tmem_base, d1 = tmem_base + (uint32_t)a.nboxp;
        int s = 0, sl = 0;
        uint32_t ph = 0, phl = 0;
        for (int it = 0; it < nst; ++it) {
            nh_wait(&full[s], ph, a.soft, 2, it);
            if (THREE) nh_wait(&lo_full[sl], phl, a.soft, 3, it);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t a_hi16 = smem_addr(smem + s * hi_bytes) >> 4, b_hi16 = a_hi16 + (a_bytes >> 4);
            const uint32_t a_lo16 = smem_addr(lo_ring + sl * hi_bytes) >> 4, b_lo16 = a_lo16 + (a_bytes >> 4);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                const uint32_t acc = (it > 0 || ks > 0) ? 1u : 0u;
                // a k-step = 8 pixel rows = 1024 B; 32-wide MN blocks 4096 B apart (LBO), 4-row atoms 512 B apart (SBO)
                tm_umma_tf32_p(d0, (a_hi16 + ks * 64) | DESC_A_MN_LO, DESC_A_MN_HI, (b_hi16 + ks * 64) | DESC_A_MN_LO, DESC_A_MN_HI, idesc, acc, leader);
                if (THREE) {
                    tm_umma_tf32_p(d1, (a_hi16 + ks * 64) | DESC_A_MN_LO, DESC_A_MN_HI, (b_lo16 + ks * 64) | DESC_A_MN_LO, DESC_A_MN_HI, idesc, acc, leader);
                    tm_umma_tf32_p(d1, (a_lo16 + ks * 64) | DESC_A_MN_LO, DESC_A_MN_HI, (b_hi16 + ks * 64) | DESC_A_MN_LO, DESC_A_MN_HI, idesc, 1u, leader);
                }
            }
            tm_commit_p(&empty[s], leader);
            if (THREE) tm_commit_p(&lo_empty[sl], leader);
            if (++s == NST) { s = 0; ph ^= 1; }
            if (++sl == NSTL) { sl = 0; phl ^= 1; }
        }
        if (nst > 0) tm_commit_p(accum_bar, leader);
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    } else {
        // ===================== remainder pass, then the epilogue: 8 warps =====================
        const int wt = tid - 64;
        if (THREE) {
            const int nv = hi_bytes >> 4;
            int s = 0, sl = 0;
            uint32_t ph = 0, phl = 0;
            for (int it = 0; it < nst; ++it) {
                nh_wait(&full[s], ph, a.soft, 5, it);
                if (it >= NSTL) nh_wait(&lo_empty[sl], phl ^ 1, a.soft, 6, it);
                const float4* hi = (const float4*)(smem + s * hi_bytes);
                float4* lo = (float4*)(lo_ring + sl * hi_bytes);
                for (int i = wt; i < nv; i += 256) lo[i] = tf32_rest4(hi[i]);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) tm_mbar_arrive(&lo_full[sl]);
                if (++s == NST) { s = 0; ph ^= 1; }
                if (++sl == NSTL) { sl = 0; phl ^= 1; }
            }
        }
        if (nst > 0) nh_wait(accum_bar, 0, a.soft, 7, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int q4 = warp & 3, colh = (warp - 2) >> 2;
        const int p = p0 + q4;                                                 // TMEM lane quarter = one (tap, ci-block) pair
        const int tap = min(p, a.npairs - 1) / a.ciblocks;
        const int ci = (min(p, a.npairs - 1) - tap * a.ciblocks) * 32 + lane;
        const bool rvalid = (p < a.npairs) && (ci < a.Ci);
        const int ntile = min(128, a.Co - n0);
        float* outp = a.out + (long long)blockIdx.z * a.numel;
        const uint32_t trow = tmem_base + ((uint32_t)(q4 * 32) << 16);
        for (int cg = colh; cg * 16 < ntile; cg += 2) {
            float v[16];
            if (nst > 0) {
                tm_ld16(trow + (uint32_t)(cg * 16), v);
                if (THREE) {
                    float v2[16];
                    tm_ld16(trow + (uint32_t)(a.nboxp + cg * 16), v2);
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
                    if (cg * 16 + j < ntile) outp[((long long)co * a.Ci + ci) * a.KK + tap] = v[j];
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

__global__ void __launch_bounds__(256) nhwc_wgrad_sum_kernel(const float* __restrict__ work, float* __restrict__ out, long long numel, int splits) {
    CCB_PDL_WAIT();
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= numel) return;
    float v = 0.f;
    for (int s = 0; s < splits; ++s) v += __ldg(work + (long long)s * numel + i);
    out[i] = v;
}

static bool nhwc_wgrad_plan(const ccb_conv_desc* d, int three, NhwcWgradArgs& a, int& smem) {
    memset(&a, 0, sizeof(a));
    a.B = d->B; a.Ci = d->Ci; a.Co = d->Co; a.Ho = d->Ho; a.Wo = d->Wo; a.KK = d->kh * d->kw; a.kw = d->kw; a.pad = d->pad;
    a.ciblocks = cdiv(d->Ci, 32);
    a.npairs = a.KK * a.ciblocks;
    a.tiles_x = cdiv(d->Wo, 8); a.tiles_y = cdiv(d->Ho, 4);
    a.stages = d->B * a.tiles_x * a.tiles_y;
    const int ntile = d->Co < 128 ? d->Co : 128;
    a.nblkN = cdiv(ntile, 32);
    a.nboxp = a.nblkN * 32;
    const int hi = 4 * NW_BLK + a.nblkN * NW_BLK;               // one operand stage; the remainder ring holds 2 more of this size
    a.nstages = (222 * 1024 - (three ? 2 * hi : 0)) / hi;
    if (a.nstages > 8) a.nstages = 8;
    if (a.nstages < 3) return false;
    smem = (a.nstages + (three ? 2 : 0)) * hi + 1024 + 1024;
    const int base = cdiv(a.npairs, 4) * cdiv(d->Co, 128);
    int splits = base >= 148 ? 1 : cdiv(148, base);
    if (splits > a.stages / 6) splits = a.stages / 6;
    if (splits < 1) splits = 1;
    a.per_split = cdiv(a.stages, splits);
    a.splits = cdiv(a.stages, a.per_split);
    a.numel = (long long)d->Co * d->Ci * a.KK;
    return true;
}
// Which problems: stride 1, enough channels on both sides to fill 32-wide boxes, maps whose 4 x 8 pixel tiles are mostly full
bool nhwc_wgrad_takes(const ccb_conv_desc* d) {
    if (!g_nhwc_enabled || !g_nhwc_wgrad_on || get_encode() == nullptr) return false;
    if (d->stride != 1 || d->kh != d->kw || d->Ci < 24 || d->Co < 24) return false;
    if (d->Ho < 4 || d->Wo < 8) return false;
    if ((long long)cdiv(d->Ho, 4) * 4 * cdiv(d->Wo, 8) * 8 * 10 > (long long)d->Ho * d->Wo * 14) return false;
    if ((long long)d->B * d->Ho * d->Wo < 2048) return false;
    NhwcWgradArgs a;
    int smem;
    return nhwc_wgrad_plan(d, 1, a, smem);
}
long long nhwc_wgrad_workspace_floats(const ccb_conv_desc* d) {
    NhwcWgradArgs a;
    int smem;
    if (!nhwc_wgrad_plan(d, 1, a, smem)) return -1;
    return nhwc_copy_floats(d->B, d->Ci, d->Hi, d->Wi) + nhwc_copy_floats(d->B, d->Co, d->Ho, d->Wo) + (a.splits > 1 ? a.splits * a.numel : 0);
}
int nhwc_wgrad(const ccb_conv_desc* d, const float* x, const float* dz, float* dw, float* work, long long work_floats, int three,
               cudaStream_t st) {
    EncodeTiledFn enc = get_encode();
    CCB_REQUIRE(enc != nullptr, CCB_ERR_UNSUPPORTED, "conv_nhwc wgrad: cuTensorMapEncodeTiled unavailable");
    NhwcWgradArgs a;
    int smem = 0;
    CCB_REQUIRE(nhwc_wgrad_plan(d, three, a, smem), CCB_ERR_UNSUPPORTED, "conv_nhwc wgrad: no tiling");
    a.soft = g_nhwc_soft;
    const long long xf = nhwc_copy_floats(d->B, d->Ci, d->Hi, d->Wi), zf = nhwc_copy_floats(d->B, d->Co, d->Ho, d->Wo);
    CCB_REQUIRE(work && xf + zf + (a.splits > 1 ? a.splits * a.numel : 0) <= work_floats, CCB_ERR_ARG, "conv_nhwc wgrad: workspace too small");
    float* xh = work;
    float* zh = work + xf;
    int rc = nhwc_transpose(x, xh, d->B, d->Ci, d->Hi, d->Wi, st);
    if (rc) return rc;
    rc = nhwc_transpose(dz, zh, d->B, d->Co, d->Ho, d->Wo, st);
    if (rc) return rc;
    a.out = a.splits > 1 ? work + xf + zf : dw;
    const int Cip = (d->Ci + 3) & ~3, Cop = (d->Co + 3) & ~3;
    alignas(64) CUtensorMap map_x, map_dz;
    {
        cuuint64_t dims[4] = {(cuuint64_t)Cip, (cuuint64_t)d->Wi, (cuuint64_t)d->Hi, (cuuint64_t)d->B};
        cuuint64_t strides[3] = {(cuuint64_t)Cip * 4, (cuuint64_t)d->Wi * Cip * 4, (cuuint64_t)d->Hi * d->Wi * Cip * 4};
        cuuint32_t box[4] = {32, 8, 4, 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = enc(&map_x, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)xh, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        CCB_REQUIRE(r == CUDA_SUCCESS, CCB_ERR_LAUNCH, "conv_nhwc wgrad: cuTensorMapEncodeTiled(x) failed (%d)", (int)r);
    }
    {
        cuuint64_t dims[4] = {(cuuint64_t)Cop, (cuuint64_t)d->Wo, (cuuint64_t)d->Ho, (cuuint64_t)d->B};
        cuuint64_t strides[3] = {(cuuint64_t)Cop * 4, (cuuint64_t)d->Wo * Cop * 4, (cuuint64_t)d->Ho * d->Wo * Cop * 4};
        cuuint32_t box[4] = {32, 8, 4, 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = enc(&map_dz, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)zh, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        CCB_REQUIRE(r == CUDA_SUCCESS, CCB_ERR_LAUNCH, "conv_nhwc wgrad: cuTensorMapEncodeTiled(dz) failed (%d)", (int)r);
    }
    dim3 grid(cdiv(a.npairs, 4), cdiv(d->Co, 128), a.splits);
    auto kfn = three ? conv_wgrad_nhwc_kernel<true> : conv_wgrad_nhwc_kernel<false>;
    cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    CCB_LAUNCH(kfn, grid, dim3(NH_THREADS), smem, st, map_x, map_dz, a);
    rc = check_launch("conv_nhwc_wgrad");
    if (rc || a.splits == 1) return rc;
    CCB_LAUNCH(nhwc_wgrad_sum_kernel, dim3((unsigned)((a.numel + 255) / 256)), dim3(256), 0, st, (const float*)a.out, dw, a.numel, a.splits);
    return check_launch("conv_nhwc_wgrad_reduce");
}

}  // namespace ccb

using namespace ccb;

// bring-up: enabled (0 turns the channels-last kernel off), soft (barrier time-outs are recorded, not trapped),
// dbg bit 0: descriptors WITH the base-offset field (wrong results for unaligned taps), bit 1: thin layers (16 .. 31 gathered
// channels) take this path instead of the CUDA-core direct kernel, bit 2: 1x1 convolutions take this path
// too, bit 3: one CTA per SM (the 128-wide configuration) only, bits 4-6: force that many stacked tiles per CTA, bit 7:
// stride-1 weight gradients through the channels-last wgrad kernel
extern "C" void ccb_debug_nhwc(int enabled, int soft, int dbg) { nhwc_set_debug(enabled, soft, dbg); }
extern "C" int ccb_debug_nhwc_status(unsigned int* out4) {
    cudaDeviceSynchronize();
    unsigned int h[4] = {0, 0, 0, 0}, z[4] = {0, 0, 0, 0};
    cudaMemcpyFromSymbol(h, g_nhwc_status, sizeof(h));
    cudaMemcpyToSymbol(g_nhwc_status, z, sizeof(z));
    cudaGetLastError();
    for (int i = 0; i < 4; ++i) out4[i] = h[i];
    return CCB_OK;
}
#else
namespace ccb {
bool nhwc_applies(const int*, const int*, int, int, int, int, int, int, int, bool) { return false; }
bool nhwc_prefers_thin() { return false; }
long long nhwc_copy_floats(int, int, int, int) { return 0; }
bool nhwc_wgrad_takes(const ccb_conv_desc*) { return false; }
long long nhwc_wgrad_workspace_floats(const ccb_conv_desc*) { return -1; }
int nhwc_wgrad(const ccb_conv_desc*, const float*, const float*, float*, float*, long long, int, cudaStream_t) { return CCB_ERR_UNSUPPORTED; }
long long nhwc_wp_floats(const int*, const int*, int, int, int) { return -1; }
}
extern "C" void ccb_debug_nhwc(int, int, int) {}
extern "C" int ccb_debug_nhwc_status(unsigned int* out4) { out4[0] = out4[1] = out4[2] = out4[3] = 0; return 0; }
#endif
