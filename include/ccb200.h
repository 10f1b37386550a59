/* ccb200.h - C ABI of libccb200.so: the B200 (sm_100a) implementation of the Competitive-Collaboration
 * training step's dense per-pixel path (anuragranj/cc @ 2b4e362).
 *
 * Conventions:
 *   - every pointer is a DEVICE pointer to contiguous fp32 NCHW data unless a comment says "host";
 *   - the caller (torch.empty on the Python side) owns every buffer, including workspaces and the
 *     buffers saved for backward; the library owns no tensors and keeps no state between calls (the only process-wide
 *     variables are the launch counter and the bring-up switches of ccb200_debug.h; a weight cache is an explicit handle
 *     the caller creates, passes in ccb_conv_desc and destroys)
 *     (the reference caches a module-global pixel grid, inverse_warp.py:10 - we do not);
 *   - every entry point is asynchronous on `stream` (a cudaStream_t), never synchronises the host,
 *     never throws, and returns CCB_OK or a negative ccb_status; ccb_last_error_string() explains it;
 *   - shape errors that the reference reports as Python AssertionError (inverse_warp.py:23-28) are
 *     raised by the Python mirror (cc_b200/*.py) before the call; the ABI re-checks what it needs.
 *
 * Each entry point names the reference interface it replaces (file:line relative to the reference).
 */
#ifndef CCB200_H
#define CCB200_H

#ifdef __cplusplus
extern "C" {
#endif

#define CCB_MAX_LEVELS 8
#define CCB_MAX_REFS 4
#define CCB_SSIM_TAPS 13

typedef void* ccb_stream_t; /* cudaStream_t */

typedef enum ccb_status {
    CCB_OK = 0,
    CCB_ERR_ARG = -1,         /* bad size / null pointer / unsupported combination */
    CCB_ERR_LAUNCH = -2,      /* CUDA launch or runtime error */
    CCB_ERR_UNSUPPORTED = -3
} ccb_status;

const char* ccb_last_error_string(void);
int ccb_version(void);
/* 1 only for the CPU execution-model simulator build used by the GPU-less unit tests (tests/sim). */
int ccb_is_simulator(void);

/* ------------------------------------------------------------------------------------------------
 * Image pyramid: level l = exact 2^l x 2^l box mean of the full-resolution planes.
 * Replaces the 15 adaptive_avg_pool2d calls per level per step (loss_functions.py:36-37,89-90,
 * 163-165,315).  out_levels: HOST array of nlevels-1 device pointers (levels 1..nlevels-1),
 * each [planes, H>>l, W>>l].  H and W must be divisible by 2^(nlevels-1).
 * ---------------------------------------------------------------------------------------------- */
int ccb_image_pyramid(const float* img, int planes, int H, int W, int nlevels,
                      float* const* out_levels, ccb_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Fused multi-scale photometric loss (one launch covers every pyramid level).
 *   mode CCB_PHOTO_RIGID : photometric_reconstruction_loss  (loss_functions.py:80-128)
 *        = pixel2cam -> pose_vec2mat -> cam2pixel -> bilinear sample (inverse_warp.py:250-283)
 *          + depth_occlusion_masks (loss_functions.py:132-137,343-352) + SSIM 13x13 (ssim.py:19-36)
 *          + robust L1 (loss_functions.py:18-21) + oob normalisation.
 *   mode CCB_PHOTO_FLOW  : photometric_flow_loss             (loss_functions.py:27-77)
 *        = flow_warp (inverse_warp.py:164-192) + occlusion_masks + SSIM + robust L1.
 *   mode CCB_PHOTO_CONSENSUS : consensus_exp_masks targets   (loss_functions.py:160-202), R = 3
 *        "refs": (ref_fwd,cam_flow_fwd) (ref_bwd,cam_flow_bwd) (ref_fwd,flow_fwd); no gradient.
 * ---------------------------------------------------------------------------------------------- */
enum { CCB_PHOTO_RIGID = 0, CCB_PHOTO_FLOW = 1, CCB_PHOTO_CONSENSUS = 2 };
enum { CCB_ROT_EULER = 0, CCB_ROT_QUAT = 1 };
enum { CCB_PAD_ZEROS = 0, CCB_PAD_BORDER = 1, CCB_PAD_NONE = 2 };

typedef struct ccb_photo_desc {
    int mode;
    int B, R;                 /* batch, number of reference frames (<= CCB_MAX_REFS) */
    int H, W;                 /* full-resolution size: downscale_l = H / h[l] (loss_functions.py:87) */
    int nlevels;
    int h[CCB_MAX_LEVELS], w[CCB_MAX_LEVELS];
    int has_mask;             /* explainability mask given (mask[l] != NULL for all l) */
    int has_occ;              /* apply occlusion masks (always 1 in the reference paths) */
    int rotation_mode;        /* CCB_ROT_* (rigid) */
    int padding_mode;         /* CCB_PAD_ZEROS | CCB_PAD_BORDER (rigid) */
    float wssim, qch, lambda_oob, wrig;
    float one_minus_wssim;    /* (1 - wssim) evaluated in double by the caller, as the reference does */
    float taps[CCB_SSIM_TAPS]; /* fp32 Gaussian taps exactly as ssim.py:9-11 builds them */
    /* inputs */
    const float* tgt[CCB_MAX_LEVELS];                 /* [B,3,h,w] pooled target frame */
    const float* ref[CCB_MAX_LEVELS][CCB_MAX_REFS];   /* [B,3,h,w] pooled reference frames */
    const float* depth[CCB_MAX_LEVELS];               /* rigid: [B,1,h,w] */
    const float* flow[CCB_MAX_LEVELS][CCB_MAX_REFS];  /* flow/consensus: [B,2,h,w] */
    const float* mask[CCB_MAX_LEVELS];                /* [B,R,h,w] or NULL */
    const float* pose;                                /* rigid: [B,R,6] */
    const float* K;                                   /* rigid: [B,3,3] full-res intrinsics */
    const float* Kinv;                                /* rigid: [B,3,3] */
    /* saved for backward (written by fwd, read by bwd) */
    float* dmaps[CCB_MAX_LEVELS];    /* [B,R,9,h,w] gamma * dS/d(mu2,Eyy,Exy); unused when wssim == 0 */
    float* gmask[CCB_MAX_LEVELS];    /* [B,R,h,w]  unscaled d loss / d mask (has_mask only) */
    float* vo[CCB_MAX_LEVELS];       /* [B,R,h,w]  valid * (1 - occ) */
    float* scal;                     /* [nlevels,R,4] : c_l, oob, sum_valid, level-ref loss */
    /* forward outputs / workspace */
    float* partials;                 /* [ccb_photo_partials_floats()] */
    float* loss;                     /* [1] */
    float* target[CCB_MAX_LEVELS];   /* consensus: [B,1,h,w] 0/1 */
    /* backward inputs / outputs / workspace */
    const float* grad_out;           /* [1] d L / d loss */
    float* d_depth[CCB_MAX_LEVELS];                 /* rigid: [B,1,h,w] */
    float* d_flow[CCB_MAX_LEVELS][CCB_MAX_REFS];    /* flow:  [B,2,h,w] */
    float* d_mask[CCB_MAX_LEVELS];                  /* [B,R,h,w] (has_mask) */
    float* d_pose;                                  /* rigid: [B,R,6] */
    float* pose_partials;            /* [ccb_photo_pose_partials_floats()] */
} ccb_photo_desc;

long long ccb_photo_partials_floats(const ccb_photo_desc* d);
long long ccb_photo_pose_partials_floats(const ccb_photo_desc* d);
int ccb_photo_loss_fwd(const ccb_photo_desc* d, ccb_stream_t stream);
int ccb_photo_loss_bwd(const ccb_photo_desc* d, ccb_stream_t stream);
int ccb_consensus_targets(const ccb_photo_desc* d, ccb_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Stand-alone warp layer (the functions train.py:22 imports by name).
 * ---------------------------------------------------------------------------------------------- */
/* inverse_warp (inverse_warp.py:250-283): img [B,3,h,w], depth [B,h,w], pose [B,6] with row stride
 * pose_stride floats, K/Kinv [B,3,3] (already scaled by the caller) -> out [B,3,h,w]. */
int ccb_inverse_warp_fwd(const float* img, const float* depth, const float* pose, int pose_stride,
                         const float* K, const float* Kinv, int B, int h, int w, int rotation_mode,
                         int padding_mode, float* out, ccb_stream_t stream);
/* grads wrt depth [B,h,w] and pose [B,6] (contiguous); pose_partials: B*ntiles*12 floats. */
int ccb_inverse_warp_bwd(const float* img, const float* depth, const float* pose, int pose_stride,
                         const float* K, const float* Kinv, int B, int h, int w, int rotation_mode,
                         int padding_mode, const float* grad_out, float* d_depth, float* d_pose,
                         float* pose_partials, ccb_stream_t stream);
long long ccb_warp_pose_partials_floats(int B, int h, int w);
/* flow_warp (inverse_warp.py:164-192): img [B,C,h,w], flow [B,2,h,w]; padding zeros|border. */
int ccb_flow_warp_fwd(const float* img, const float* flow, int B, int C, int h, int w,
                      int padding_mode, float* out, ccb_stream_t stream);
/* d_flow [B,2,h,w] (may be NULL), d_img [B,C,h,w] (may be NULL; must be zero-filled by the caller). */
int ccb_flow_warp_bwd(const float* img, const float* flow, int B, int C, int h, int w,
                      int padding_mode, const float* grad_out, float* d_flow, float* d_img,
                      ccb_stream_t stream);
/* pose2flow (inverse_warp.py:195-220): -> flow [B,2,h,w]; padding_mode CCB_PAD_NONE | CCB_PAD_ZEROS. */
int ccb_pose2flow_fwd(const float* depth, const float* pose, int pose_stride, const float* K,
                      const float* Kinv, int B, int h, int w, int rotation_mode, int padding_mode,
                      float* flow, ccb_stream_t stream);
int ccb_pose2flow_bwd(const float* depth, const float* pose, int pose_stride, const float* K,
                      const float* Kinv, int B, int h, int w, int rotation_mode, int padding_mode,
                      const float* grad_flow, float* d_depth, float* d_pose, float* pose_partials,
                      ccb_stream_t stream);

/* ssim map (ssim.py:68-76, window 13, sigma 1.5, zero padding): img1,img2,out [planes,h,w]. */
int ccb_ssim_fwd(const float* img1, const float* img2, int planes, int h, int w, const float* taps_host,
                 float* out, ccb_stream_t stream);
/* d_img1/d_img2 may be NULL; work: 5*planes*h*w floats. */
int ccb_ssim_bwd(const float* img1, const float* img2, int planes, int h, int w, const float* taps_host,
                 const float* grad_out, float* d_img1, float* d_img2, float* work, ccb_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Smoothness (loss_functions.py:287-341) for a list of predictions [B,C,h_l,w_l].
 *   kind CCB_SMOOTH_EDGE   : edge_aware_smoothness_loss(img, pred): needs img[l] = pooled tgt [B,3,h,w]
 *   kind CCB_SMOOTH_SECOND : smooth_loss(pred), level weight 1/2.3^l
 * ---------------------------------------------------------------------------------------------- */
enum { CCB_SMOOTH_EDGE = 0, CCB_SMOOTH_SECOND = 1 };
typedef struct ccb_smooth_desc {
    int kind, B, C, nlevels;
    int h[CCB_MAX_LEVELS], w[CCB_MAX_LEVELS];
    const float* img[CCB_MAX_LEVELS];
    const float* pred[CCB_MAX_LEVELS];
    float* partials;          /* [ccb_smooth_partials_floats()] */
    float* loss;              /* [1] */
    const float* grad_out;    /* [1] */
    float* d_pred[CCB_MAX_LEVELS];
} ccb_smooth_desc;
long long ccb_smooth_partials_floats(const ccb_smooth_desc* d);
int ccb_smooth_fwd(const ccb_smooth_desc* d, ccb_stream_t stream);
int ccb_smooth_bwd(const ccb_smooth_desc* d, ccb_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Mask cross-entropies.
 *   kind CCB_BCE_ONES     : explainability_loss (loss_functions.py:148-155): BCE(mask, 1) per level
 *   kind CCB_BCE_CONSENSUS: consensus_depth_flow_mask + weighted_binary_cross_entropy
 *                           (loss_functions.py:221-261)
 * ---------------------------------------------------------------------------------------------- */
enum { CCB_BCE_ONES = 0, CCB_BCE_CONSENSUS = 1 };
typedef struct ccb_bce_desc {
    int kind, B, C, nlevels;  /* C = mask channels (4) */
    int h[CCB_MAX_LEVELS], w[CCB_MAX_LEVELS];
    float thresh, wbce;
    const float* mask[CCB_MAX_LEVELS];        /* [B,C,h,w] */
    const float* census_bwd[CCB_MAX_LEVELS];  /* [B,2,h,w] |cam_flow_bwd - flow_bwd| */
    const float* census_fwd[CCB_MAX_LEVELS];  /* [B,2,h,w] */
    const float* target_bwd[CCB_MAX_LEVELS];  /* [B,1,h,w] */
    const float* target_fwd[CCB_MAX_LEVELS];  /* [B,1,h,w] */
    float* partials;
    float* loss;
    const float* grad_out;
    float* d_mask[CCB_MAX_LEVELS];
} ccb_bce_desc;
long long ccb_bce_partials_floats(const ccb_bce_desc* d);
int ccb_bce_fwd(const ccb_bce_desc* d, ccb_stream_t stream);
int ccb_bce_bwd(const ccb_bce_desc* d, ccb_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Convolutions of the four networks (replace the cuDNN calls behind nn.Conv2d / nn.ConvTranspose2d /
 * nn.BatchNorm2d in models/DispResNet6.py:9-94, PoseNetB6.py:10-21, MaskNet6.py:5-16,
 * back2future.py:27-48).  NCHW fp32, weights [Co,Ci,kh,kw], square stride/pad.
 *   y = act(conv(x, w) + bias + res)                               ccb_conv2d_fprop
 *   dx = act(conv_transpose(dy, w) + bias + res)                   ccb_conv2d_dgrad
 *        (plain data-gradient when bias/res are NULL and act is NONE; with them it is the
 *         nn.ConvTranspose2d forward of a layer whose torch weight [Cin_t,Cout_t,k,k] is this w)
 *   dw = d/dw, db = sum dy                                         ccb_conv2d_wgrad
 * impl: CCB_CONV_IMPL_AUTO picks tcgen05 tensor-core tiles where the shape allows and the FFMA
 *       kernels otherwise; _FFMA / _TC force one (tests).
 * ---------------------------------------------------------------------------------------------- */
enum { CCB_ACT_NONE = 0, CCB_ACT_RELU = 1, CCB_ACT_LEAKY = 2, CCB_ACT_SIGMOID = 3 };
enum { CCB_CONV_FPROP = 0, CCB_CONV_DGRAD = 1, CCB_CONV_WGRAD = 2 };
enum { CCB_CONV_IMPL_AUTO = 0, CCB_CONV_IMPL_FFMA = 1, CCB_CONV_IMPL_TC = 2 /* tcgen05 3xTF32 (fp32 parity) */,
       CCB_CONV_IMPL_TC_TF32 = 3 /* tcgen05 single TF32 (cuDNN's default math for the reference) */ };
typedef struct ccb_conv_desc {
    int B, Ci, Hi, Wi;      /* input  [B,Ci,Hi,Wi] */
    int Co, Ho, Wo;         /* output [B,Co,Ho,Wo]; Ho = (Hi + 2 pad - kh) / stride + 1 */
    int kh, kw, stride, pad;
    int act;                /* CCB_ACT_* fused into the epilogue */
    float slope;            /* LeakyReLU negative slope */
    int impl;               /* CCB_CONV_IMPL_* */
    void* wcache;           /* weight cache handle (ccb_wcache_create) or NULL: prepared weight copies are then made per call */
} ccb_conv_desc;
/* Weight cache.  The tensor-core kernels read weights from a prepared copy (tf32 hi | lo split, K order of the kernel).
 * Without a cache every conv call prepares its copy into `work`.  With one (a trainer owns it; the weights then only
 * change in the optimiser step): while the cache is RECORDING, conv calls note which prepared layouts they need (and still
 * prepare on the spot); ccb_wcache_plan_floats / ccb_wcache_table_bytes size the caller-allocated persistent buffer and
 * device table, ccb_wcache_commit binds them, ccb_wcache_refresh re-prepares EVERY copy in one launch (call it after each
 * weight update), and conv calls whose (weights, layout) are recorded skip their preparation launch.
 * The caller must refresh after ANY change of the weights (optimizer step, load_state_dict). */
void* ccb_wcache_create(void);
void ccb_wcache_destroy(void* cache);
long long ccb_wcache_plan_floats(void* cache);
long long ccb_wcache_table_bytes(void* cache);
int ccb_wcache_commit(void* cache, float* buf, long long buf_floats, void* table, long long table_bytes, ccb_stream_t stream);
int ccb_wcache_refresh(void* cache, ccb_stream_t stream);
/* out4 = {recorded layouts, cache hits, misses after commit, state (0 recording, 1 committed)} */
void ccb_wcache_stats(void* cache, long long* out4);
long long ccb_conv_workspace_floats(const ccb_conv_desc* d, int op);
int ccb_conv2d_fprop(const ccb_conv_desc* d, const float* x, const float* w, const float* bias,
                     const float* res, float* y, float* work, long long work_floats, ccb_stream_t stream);
int ccb_conv2d_dgrad(const ccb_conv_desc* d, const float* dy, const float* w, const float* bias,
                     const float* res, float* dx, float* work, long long work_floats, ccb_stream_t stream);
int ccb_conv2d_wgrad(const ccb_conv_desc* d, const float* x, const float* dy, float* dw, float* db,
                     float* work, long long work_floats, ccb_stream_t stream);
/* dz = dy * act'(.) expressed through the activation OUTPUT y (in place allowed: dz == dy). */
int ccb_act_bwd(const float* dy, const float* y, float* dz, long long numel, int act, float slope,
                ccb_stream_t stream);
/* fused: dz = dy * act'(y) (not touched when act == CCB_ACT_NONE; in place allowed) and, when db != NULL,
 * db[c] = sum over (b, pixel) of dz - one pass over the gradient.  work: ccb_act_bwd_bias_workspace_floats() floats. */
long long ccb_act_bwd_bias_workspace_floats(int B, int C, int plane);
int ccb_act_bwd_bias(const float* dy, const float* y, float* dz, float* db, int B, int C, int plane, int act, float slope,
                     float* work, long long work_floats, ccb_stream_t stream);
/* db[c] = sum over (b, pixel) of dy; `work` (ccb_bias_grad_workspace_floats) lets large planes be reduced in two
 * deterministic stages, without it one block per channel does the whole sum */
long long ccb_bias_grad_workspace_floats(int B, int C, int plane);
int ccb_bias_grad(const float* dy, float* db, int B, int C, int plane, float* work, long long work_floats,
                  ccb_stream_t stream);
/* bring-up / diagnostic entry points (ccb_debug_*): include/ccb200_debug.h - not part of the drop-in boundary */

/* Back2Future operators (models/back2future.py).
 * corr81: cost volume of correlate() :15-25 (third-party spatial_correlation_sample, kernel 1, patch 9,
 * zero padded, divided by C) with the reference's channel permutation baked in (reversed=0: idx_fwd,
 * 1: idx_bwd, :56-59).  f1,f2 [B,C,h,w] -> out [B,81,h,w].   d_f1 / d_f2 may be NULL.
 * featwarp: Model.warp :287-321 = grid_sample(x, grid+flow, padding border, align_corners False). */
/* work: ccb_corr81_fwd_workspace_floats() floats (per-channel-chunk partial sums; 0 when one CTA per tile sums all
 * channels) */
long long ccb_corr81_fwd_workspace_floats(int B, int C, int h, int w);
int ccb_corr81_fwd(const float* f1, const float* f2, float* out, int B, int C, int h, int w, int reversed,
                   float* work, long long work_floats, ccb_stream_t stream);
/* work: B*81*h*w floats (the mirrored gradient planes), required when d_f2 != NULL */
int ccb_corr81_bwd(const float* f1, const float* f2, const float* grad_out, float* d_f1, float* d_f2, int B,
                   int C, int h, int w, int reversed, float* work, ccb_stream_t stream);
int ccb_featwarp_fwd(const float* x, const float* flow, int B, int C, int h, int w, float* out,
                     ccb_stream_t stream);
/* d_x must be zero-filled by the caller (scatter-add); d_flow / d_x may be NULL */
int ccb_featwarp_bwd(const float* x, const float* flow, int B, int C, int h, int w, const float* grad_out,
                     float* d_flow, float* d_x, ccb_stream_t stream);

/* BatchNorm2d over [B,C,plane] (DispResNet6.py:45-52).  training: batch statistics, stats[C][2] =
 * {mean, invstd} saved for backward, running stats updated in place (momentum, unbiased var). */
long long ccb_bn_workspace_floats(int B, int C, int plane);   /* `work` size for both calls */
int ccb_bn_fwd(const float* x, const float* gamma, const float* beta, float* y, float* stats,
               float* running_mean, float* running_var, int B, int C, int plane, float eps, float momentum,
               int training, float* work, ccb_stream_t stream);
int ccb_bn_bwd(const float* x, const float* dy, const float* gamma, const float* stats, float* dx,
               float* dgamma, float* dbeta, int B, int C, int plane, float* work, ccb_stream_t stream);
/* bilinear x2 upsample, align_corners=False (DispResNet6.py:174; back2future.py:60): [planes,h,w] -> [planes,2h,2w] */
int ccb_upsample2x_fwd(const float* x, float* y, int planes, int h, int w, ccb_stream_t stream);
int ccb_upsample2x_bwd(const float* dy, float* dx, int planes, int h, int w, ccb_stream_t stream);
/* torch.optim.Adam step (train.py:307-310,568) on one flat fp32 buffer; grad_scale pre-multiplies the
 * gradient (1/world_size after the NCCL all-reduce sum).  state: 3 device floats {step count, 1-b1^t,
 * sqrt(1-b2^t)}, zero-initialised by the caller; the call increments the step on the device, so a
 * captured CUDA graph of the training step replays correctly. */
int ccb_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long long n,
                  float* state, float lr, float beta1, float beta2, float eps, float grad_scale,
                  ccb_stream_t stream);
/* ---- callers either side of the step -------------------------------------------------------
 * Validation metrics, reference loss_functions.py:355-467, as fused masked reductions (deterministic, no host sync).
 * ccb_flow_metrics: gt [B,nc,Hg,Wg] (nc 3: third channel = valid mask; nc 2: plain mean), predictions [B,2,hp,wp]
 *   bilinearly resized to the ground truth (F.upsample(size=gt) = align_corners False) and rescaled by Wg/wp, Hg/hp.
 *   pred_nonrigid / rigidity_mask NULL: out4 = {compute_epe :368-387, -, -, outlier_err :389-407 with tau = (tau0, tau1)}.
 *   Otherwise compute_all_epes :409-427 (mask [B,1,hm,wm], composite by mask > thresh at prediction resolution,
 *   ground truth split at its own): out4 = {all, rigid, non-rigid EPE, outliers}.
 *   epe_map (optional, [B,Hg,Wg]) receives flow_diff :355-365 of the (composited) prediction.
 *   work: ccb_flow_metrics_workspace_bytes(B, Hg, Wg) bytes, 8-byte aligned.
 * ccb_depth_errors: compute_errors :430-467 on gt, pred [B,H,W]: valid = 0 < gt < 80 (inside the Garg crop when
 *   crop != 0), pred clamped to [1e-3, 80] and scaled by median(gt)/median(pred) per sample (lower medians, by radix
 *   select on the device), out6 = batch means of {abs_diff, abs_rel, sq_rel, a1, a2, a3}. */
long long ccb_flow_metrics_workspace_bytes(int B, int Hg, int Wg);
int ccb_flow_metrics(const float* gt, const float* pred_rigid, const float* pred_nonrigid, const float* rigidity_mask,
                     int B, int nc, int Hg, int Wg, int hp, int wp, int hm, int wm, float thresh, float tau0,
                     float tau1, float* epe_map, void* work, float* out4, ccb_stream_t stream);
long long ccb_depth_errors_workspace_bytes(int B, int H, int W);
int ccb_depth_errors(const float* gt, const float* pred, int B, int H, int W, int crop, void* work, float* out6,
                     ccb_stream_t stream);
/* Input pipeline on the device (train.py:448-451 H2D + custom_transforms.py:21-30,47-118): uint8 HWC frames
 * src [B,F,Hs,Ws,3] -> F normalised fp32 NCHW tensors dst[f] [B,3,H,W] = (v/255 - .5)/.5, per sample horizontally
 * flipped (params[b][0] != 0) and scale-cropped: resized by (params[b][1], params[b][2]) = (scaled_w/Ws, scaled_h/Hs)
 * with a half-pixel-centre bilinear lookup, then cropped at offs[b] = (x0, y0).  params [B,4] floats, offs [B,2] ints,
 * device memory; dst = host array of F device pointers. */
int ccb_prep_frames(const unsigned char* src_u8, float* const* dst, const float* params, const int* offs, int B, int F,
                    int Hs, int Ws, int H, int W, ccb_stream_t stream);
/* number of kernel launches issued through this library by the calling process so far */
long long ccb_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* CCB200_H */
