/*
 * pof.h — C ABI of libpof.so, the sm_100a (B200) implementation of the DR-SPAAM
 * per-point scan hot path of huzjkevin/planar_optical_flow.
 *
 * The reference has no FFI / plugin layer: its boundary for this path is four
 * Python callables.  Each entry point below states which one
 * it replaces (paths relative to the reference root).  The Python host code in
 * planar_optical_flow_b200/ binds these with ctypes and re-exposes the
 * reference's own signatures; INTEGRATION.md shows the stub a reference
 * maintainer would add.
 *
 * Conventions (all entry points)
 *   - every data pointer is a DEVICE pointer into caller-owned memory, except
 *     where an argument is explicitly described as a host pointer;
 *   - tensors are dense, row-major, in the layouts written beside each
 *     argument; float = IEEE binary32, double = binary64, int = int32;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default
 *     stream); work is enqueued, never synchronised, by the device entry points;
 *   - return 0 on success, a NEGATIVE pof_status on a bad argument, a POSITIVE
 *     cudaError_t if the CUDA runtime failed; pof_last_error() then returns a
 *     thread-local, human readable message;
 *   - no global mutable state, no allocation that outlives a call (scratch is
 *     passed in as `ws`), no CPU fallback: without a CUDA device every compute
 *     entry point fails with a CUDA error.
 */
#ifndef POF_H_
#define POF_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define POF_ABI_VERSION 2

#if defined(__GNUC__)
#define POF_API __attribute__((visibility("default")))
#else
#define POF_API
#endif

typedef enum pof_status {
    POF_OK = 0,
    POF_ERR_NULL_POINTER = -1,
    POF_ERR_BAD_SHAPE = -2,
    POF_ERR_BAD_PARAM = -3,
    POF_ERR_WORKSPACE = -4,
    POF_ERR_UNSUPPORTED = -5
} pof_status;

POF_API int pof_abi_version(void);
POF_API const char* pof_last_error(void);

/* Number of SMs / compute capability of the current device (for the host side's
 * sanity checks: it refuses to run on anything that is not sm_100).            */
POF_API int pof_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ------------------------------------------------------------------------- *
 * 1. Distance-adaptive polar cutout
 *    replaces  scans_to_cutout        src/utils/utils.py:259-334
 *    (and the unused torch port scans_to_cutout_torch, :337-420, whose results
 *     follow :259-334's mixed fp32/fp64 arithmetic here, not the all-fp32 port)
 *
 *    One call = B independent invocations of the reference function (one per
 *    sequence sample; the area-mode oversampling factor `s_area`, utils.py:308,
 *    is a per-invocation global and is therefore reduced per b).
 *
 *    scans   [B, S, N] float        ranges; S scans per sample, newest last
 *    phi     [N] float or double    beam angles (`phi_is_f64` selects);
 *                                   uniform pitch, phi[1]-phi[0] != 0
 *    out     [B, M, S, P] float     M = ceil(N / stride)   (utils.py:332-334,
 *                                   batched as dataset_dr_spaam.py:464-468)
 *    s_area_out [B] int or NULL     the factor each sample used (0 = no point
 *                                   of that sample was area-resampled)
 *    half_alpha_in  [B, S, M] float or NULL   window half-angles to use INSTEAD of
 *                                   atan(0.5*window_width / max(d, 1e-2)) (utils.py:279).
 *                                   NumPy's float32 arctan is a platform-specific SIMD
 *                                   kernel (1-2 ulp from correctly rounded), the one
 *                                   step no other machine can reproduce bit for bit;
 *                                   parity tests feed the reference's own values here
 *                                   to prove every other operation exact.
 *    half_alpha_out [B, S, M] float or NULL   the half-angles this call used
 *    ws      pof_cutout_ws_bytes(B) bytes of device scratch
 *
 *    window_width, window_depth, padding_val are doubles because the reference
 *    receives Python floats and rounds them at specific places (utils.py:279,
 *    326-330).  `fixed`, `centered`, `area_mode` as in the reference; window_depth > 0.
 *    `numerics` selects POF_CUTOUT_EXACT or POF_CUTOUT_FAST (same algorithm, cheaper
 *    arithmetic; see csrc/pof_cutout.cu).
 * ------------------------------------------------------------------------- */
#define POF_CUTOUT_EXACT 0   /* every rounding of the reference reproduced (modulo the arctangent) */
#define POF_CUTOUT_FAST 1    /* fixed-point index line + float32 blend: within ~3e-6 of EXACT        */
#define POF_CUTOUT_EXACT_PIECES 2   /* EXACT on the first (piece-per-thread) kernel: same bits, kept as a cross-check */

POF_API size_t pof_cutout_ws_bytes(int B);

POF_API int pof_cutout_fwd(const float* scans, const void* phi, int phi_is_f64,
                   int B, int S, int N, int stride, int P,
                   double window_width, double window_depth, double padding_val,
                   int fixed, int centered, int area_mode, int numerics,
                   float* out, int* s_area_out,
                   const float* half_alpha_in, float* half_alpha_out,
                   void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------- *
 * 2. Auto-regressive spatial-attention memory update
 *    replaces  _SpatialAttention.forward  src/depracted/model/dr_spaam.py:183-215
 *    (everything after the two embedding convolutions of :176-181, which stay
 *     on cuDNN) with the neighbour table of _generate_neighbor_mask, :145-160.
 *
 *    x, tmpl      [B, N, CL] float  current / remembered per-point features
 *                                   (CL = 256*14 = 3584 for DR-SPAAM), CL % 4 == 0,
 *                                   16-byte aligned
 *    emb_x, emb_t [B, N, E] float   their similarity embeddings, E % 4 == 0
 *    W                              window size 2*hw+1 (odd, 1..15)
 *    out_tmpl     [B, N, CL] float  alpha*x + (1-alpha)*sum_j w_ij tmpl_j   (:210-215)
 *                                   must NOT alias x or tmpl (template rows are
 *                                   re-read by neighbouring points)
 *    feat_fused   [B, N, W] float   raw similarities at CLAMPED neighbour indices (:187)
 *    attn_w       [B, N, W] float or NULL   softmax weights over the UNIQUE in-range
 *                                   neighbours (0 on clamped duplicates); saved for bwd
 *    out_split    [B, N, CL / split_channels, 2 * split_channels] binary16 or NULL
 *                                   the new memory again as the [hi | lo] operand rows of
 *                                   pof_conv_tc_f16_fwd (channels-last memory rows of
 *                                   split_channels channels), written in the same pass
 *    status       device int or NULL   set to 32 if a staged copy never landed (results
 *                                   invalid), 16 if a value of out_split left the binary16
 *                                   range; the caller zeroes it
 * ------------------------------------------------------------------------- */
POF_API int pof_spaam_gate_fwd(const float* x, const float* tmpl,
                       const float* emb_x, const float* emb_t,
                       int B, int N, int CL, int E, int W, float alpha,
                       float* out_tmpl, float* feat_fused, float* attn_w,
                       void* out_split, int split_channels, int* status,
                       void* stream);

/*    Backward of the above for training (autograd through the sequential gate
 *    loop of SpatialDROW.forward, dr_spaam.py:266-273).
 *    g_out [B,N,CL], g_feat [B,N,W] (or NULL = zero)  ->
 *    g_x [B,N,CL], g_tmpl [B,N,CL], g_emb_x [B,N,E], g_emb_t [B,N,E]
 *    ws: pof_spaam_gate_bwd_ws_bytes(B,N,W) bytes of device scratch.           */
POF_API size_t pof_spaam_gate_bwd_ws_bytes(int B, int N, int W);

POF_API int pof_spaam_gate_bwd(const float* tmpl, const float* emb_x, const float* emb_t,
                       const float* attn_w, const float* g_out, const float* g_feat,
                       int B, int N, int CL, int E, int W, float alpha,
                       float* g_x, float* g_tmpl, float* g_emb_x, float* g_emb_t,
                       void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------- *
 * 3. Predicted-centre vote / group / NMS
 *    replaces  nms_predicted_center   src/utils/utils.py:535-571
 *    (with canonical_to_global :109-116 and rphi_to_xy :47-48)
 *
 *    One call = B independent scans.
 *    scan   [B, N] float or double  (`scan_is_f64`)   ranges
 *    phi    [N]    float or double  (`phi_is_f64`)    beam angles
 *    cls    [B, N] float            post-sigmoid confidences (one class)
 *    reg    [B, N, 2] float         canonical (dx, dy) votes
 *    Arithmetic follows NumPy's promotion for those dtypes (see oracle/nms.py).
 *
 *    order         [B, N] int       point indices by descending confidence
 *                                   (ties: higher index first)
 *    keep_idx      [B, N] int       first n_keep[b] entries: ORIGINAL indices of
 *                                   the surviving centres, descending confidence
 *    n_keep        [B] int
 *    instance_mask [B, N] int       1-based id of the last kept centre within
 *                                   min_dist of each point (utils.py:565)
 *    det_xy        [B, N, 2] double first n_keep[b] rows valid (utils.py:568)
 *    det_cls       [B, N] float     first n_keep[b] valid (utils.py:569)
 *    ws            pof_nms_ws_bytes(B, N) bytes of device scratch
 * ------------------------------------------------------------------------- */
POF_API size_t pof_nms_ws_bytes(int B, int N);

POF_API int pof_nms_centers(const void* scan, int scan_is_f64, const void* phi, int phi_is_f64,
                    const float* cls, const float* reg, int B, int N, double min_dist,
                    int* order, int* keep_idx, int* n_keep, int* instance_mask,
                    double* det_xy, float* det_cls,
                    void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------- *
 * 4. Backbone glue - used by the streaming engine only.
 *    The convolutions of DROW / SpatialDROW (src/depracted/model/dr_spaam.py:8-12,
 *    49-59,87-114) stay on cuDNN; these fuse what runs between them in eval mode
 *    (bias of the BN-folded convolution, LeakyReLU :12, max_pool1d(2) :81) into one
 *    pass over channels-last activations, and optionally emit the [hi | lo | hi]
 *    TF32 operand split that makes tensor-core convolutions fp32-accurate.
 *
 *    split_parts = 3: out_split rows are [hi | lo | hi], lo = x - hi exact (operand of a cuDNN TF32
 *    convolution against [w_hi | w_hi | w_lo]); split_parts = 2: [hi | lo], lo rounded to TF32 (operand
 *    of pof_conv_tc_fwd); split_parts = POF_SPLIT_F16: [hi | lo] as binary16 (operand of pof_conv_tc_f16_fwd).
 *
 *    pof_act_fwd         y [rows_in, C] -> out_plain [rows_in/pool, C] and/or
 *                        out_split [rows_in/pool, split_parts*C]; bias [C] or NULL; pool in {1,2}
 *                        (pool = 2 takes the max of consecutive row pairs: rows of one
 *                        cutout are its L positions, L even); slope = 1 -> identity.
 *    pof_conv_first_fwd  cutouts [M, P] (x) weight [C, 3], bias [C] -> [M*P, C] / [M*P, 3C]:
 *                        the 1 -> C, k = 3, zero-padded first layer + LeakyReLU (+ split).
 *    `status` (device int or NULL): with split_parts = POF_SPLIT_F16 an activation beyond 65504 sets it to 16.
 * ------------------------------------------------------------------------- */
#define POF_SPLIT_F16 16   /* split_parts: out_split rows are [hi | lo] in binary16 (operand of pof_conv_tc_f16_fwd) */
POF_API int pof_act_fwd(const float* y, const float* bias, long long rows_in, int C, int pool,
                        float slope, float* out_plain, void* out_split, int split_parts, int* status,
                        void* stream);

POF_API int pof_conv_first_fwd(const float* cutouts, const float* weight, const float* bias,
                               long long M, int P, int C, float slope,
                               float* out_plain, void* out_split, int split_parts, int* status, void* stream);

/*    pof_head_fwd        the tail of DROW._forward_fused_cutout (dr_spaam.py:110-114): y [M, L, C] raw
 *                        output of the last convolution -> +bias, LeakyReLU -> avg_pool1d over L ->
 *                        H 1x1-convolution heads (w_head [H, C], b_head [H]; conv_cls rows first, then
 *                        conv_reg) -> sigmoid on the first n_sigmoid heads -> out [M, H].  H <= 8.
 *                        With out_rest != NULL the heads are written as two matrices instead: the first
 *                        n_sigmoid to out [M, n_sigmoid], the others to out_rest [M, H - n_sigmoid] (the
 *                        pred_cls / pred_reg tensors of dr_spaam.py:116-121).                            */
POF_API int pof_head_fwd(const float* y, const float* bias, long long M, int L, int C, float slope,
                         const float* w_head, const float* b_head, int H, int n_sigmoid,
                         float* out, float* out_rest, void* stream);

/*    pof_conv_tc_fwd     fp32-accurate convolution / whole-row GEMM on tcgen05 (3xTF32 split products,
 *                        chained accumulation promoted to fp32 registers):
 *                          out[m, l, n] = sum_{t < taps, c < Cin} A[m, l + t - pad, c] * W[t][n, c]
 *                        a_split [Mcut, LA, 2 Cin] = [hi | lo] rows (pof_act_fwd split_parts = 2);
 *                        w_split [taps, 2, Cout, Cin] = (hi, lo) per tap; rows outside 0 <= l' < LA are zero.
 *                        Conv1d(k=3,p=1): Lout = LA, taps = 3, pad = 1.  Gate embedding (dr_spaam.py:130-133,
 *                        Conv1d(k=LA)): Lout = 1, taps = LA, pad = 0.  Epilogue: + bias, LeakyReLU(slope),
 *                        max over `pool` consecutive rows, -> out_plain [Mcut*Lout/pool, Cout] and/or
 *                        out_split [.., 2 Cout].  Cin % 16 == 0; Cout in {64, 128, 256 k}.
 *                        `status` is a device int the caller zeroes: non-zero after the launch means an
 *                        internal pipeline wait timed out (results invalid).
 *                        `chain_channels`: input channels (of one tap) accumulated inside the tensor core
 *                        before the partial sum is promoted to an fp32 register accumulator; 0 = default (64),
 *                        multiples of 16.  Longer chains are faster and less accurate (tuning and tests).          */
#define POF_CONV_TC_SINGLE_CTA 0x10000   /* OR into chain_channels: one CTA per tile instead of an SM pair (cta_group::2) */
#define POF_CONV_TC_STREAM_W   0x20000   /* OR into chain_channels: never keep the weights resident in shared memory (tuning) */
#define POF_CONV_TC_NO_DEBIAS  0x40000   /* OR into chain_channels: pof_conv_tc_f16_fwd does not compensate the tensor core's truncating accumulation (tests) */
#define POF_CONV_TC_NO_SPLIT_TILE 0x80000 /* OR into chain_channels: never split a cutout between the two CTAs of a pair (tuning / tests) */
#define POF_CONV_TC_HALO 0x100000 /* OR into chain_channels: where it applies (64 input channels, k = 3), load a tile ONCE with its halo rows and run the taps on row-shifted views (measured: no faster; tests) */
POF_API int pof_conv_tc_fwd(const float* a_split, const float* w_split, const float* bias,
                            long long Mcut, int LA, int Lout, int Cin, int Cout, int taps, int pad,
                            int pool, float slope, float* out_plain, float* out_split,
                            int* status, int chain_channels, void* stream);

/*    pof_conv_tc_f16_fwd the same convolution with binary16 hi / lo parts (tcgen05 kind::f16: twice the
 *                        tensor-core rate, half the operand bytes, the same 22 significant bits).
 *                        a_split [Mcut, LA, 2 Cin] and w_split [taps, 2, Cout, Cin] are binary16
 *                        (pof_act_fwd / pof_conv_first_fwd with split_parts = POF_SPLIT_F16); the caller
 *                        multiplies a layer's weights by a power of two before splitting them (so that hi
 *                        and lo are normal binary16 numbers) and passes the inverse power as `out_scale`,
 *                        which the epilogue applies to the sum before the bias.  out_split is binary16
 *                        [.., 2 Cout]; an activation beyond 65504 sets *status = 16.  Cin % 32 == 0;
 *                        chain_channels: multiples of 32, 0 = default (128).                                    */
POF_API int pof_conv_tc_f16_fwd(const void* a_split, const void* w_split, const float* bias,
                                long long Mcut, int LA, int Lout, int Cin, int Cout, int taps, int pad,
                                int pool, float slope, float out_scale, float* out_plain, void* out_split,
                                int* status, int chain_channels, void* stream);

/*    Training-mode BatchNorm + LeakyReLU (+ max-pool over row pairs) of one conv layer (dr_spaam.py:8-12 `_conv`, :81-97
 *    max_pool1d(2)) on channels-last activations y [rows, C] (the convolution's output), as one operator with its own
 *    backward - what the training step of bin/train_dr_spaam.py spends most of its time on (profiles/r2_train_launch_summary.txt).
 *      pof_bn_act_stats  sums [2, C] double <- per-channel sum and sum of squares of y (zeroed by the call)
 *      pof_bn_act_fwd    z [rows / pool, C] = max over `pool` consecutive rows of lrelu((y - mean) * invstd * gamma + beta),
 *                        batch statistics from `sums` (biased variance); writes mean / invstd [C] (saved for the backward) and,
 *                        when given, updates running_mean / running_var with `momentum` (unbiased variance), as
 *                        torch.nn.functional.batch_norm(training=True) does
 *      pof_bn_act_bwd    dz [rows / pool, C] -> dx [rows, C], dgamma [C], dbeta [C]; recomputes the pre-activations from y
 *                        (nothing else is saved); a pooled pair's gradient goes to its first maximum; `sums` [2, C] double scratch
 *    `groups`: the rows are G consecutive blocks of rows / G rows (the S scans of a training sample, whose layers the
 *    reference runs as S separate batch_norm calls, dr_spaam.py:264-273); each block is normalised with ITS OWN batch
 *    statistics, the running statistics receive one update per block in block order, gamma / beta gradients add up.
 *    sums [G, 2, C], mean / invstd [G, C].  C % 4 == 0 with C / 4 dividing 256; pool in {1, 2}.                              */
POF_API int pof_bn_act_stats(const float* y, long long rows, int C, int groups, double* sums, void* stream);
/*    pof_conv_first_wgrad   weight gradient of the first layer, Conv1d(1 -> C, k = 3, p = 1) (dr_spaam.py:49), from
 *                        dy [M * P, C] (channels last) and the cutouts [M, P]: dw [C, 3]; `sums` [3, C] double scratch.
 *                        (The forward of that layer in training is pof_conv_first_fwd with slope = 1 and a zero bias.)      */
POF_API int pof_conv_first_wgrad(const float* dy, const float* cutouts, long long M, int P, int C,
                                 double* sums, float* dw, void* stream);
POF_API int pof_bn_act_fwd(const float* y, const double* sums, const float* gamma, const float* beta,
                           long long rows, int C, int groups, int pool, float eps, float slope, float momentum,
                           float* z, float* mean, float* invstd, float* running_mean, float* running_var, void* stream);
POF_API int pof_bn_act_bwd(const float* y, const float* dz, const float* mean, const float* invstd,
                           const float* gamma, const float* beta, long long rows, int C, int groups, int pool, float slope,
                           double* sums, float* dx, float* dgamma, float* dbeta, void* stream);

/* ------------------------------------------------------------------------- *
 * 5. Windowed patch correlation of the scan-pair flow prototype
 *    replaces Prototype._fusion, src/depracted/model/prototype.py:118-156 (dense [N, N] patch
 *    correlation followed by a +-max_displacement gather):
 *      out[b, d + D, i] = sum_c sum_{k=-h..h} f1[b, c, clamp(i+k)] * f2[b, c, clamp(clamp(i+d)+k)]
 *    feat1, feat2 [B, C, N]; out / grad_out [B, 2D+1, N]; h = kernel_size / 2 (odd kernel_size),
 *    2D+1 <= 32.  The backward is deterministic (no atomics).
 * ------------------------------------------------------------------------- */
POF_API int pof_patch_corr_fwd(const float* feat1, const float* feat2, int B, int C, int N,
                               int kernel_size, int max_displacement, float* out, void* stream);
POF_API int pof_patch_corr_bwd(const float* feat1, const float* feat2, const float* grad_out,
                               int B, int C, int N, int kernel_size, int max_displacement,
                               float* grad_feat1, float* grad_feat2, void* stream);

/* ------------------------------------------------------------------------- *
 * 6. Legacy preprocessing
 *    pof_cutout_original_fwd  replaces scans_to_cutout_original, src/utils/utils.py:423-489 (integer beam
 *        window resampled with cv2.resize: INTER_AREA when shrinking, INTER_LINEAR otherwise); selected by
 *        configs without `area_mode` (src/utils/dataset_dr_spaam.py:440-443).  scans [B, S, N] -> out
 *        [B, N, S, P], each b one reference call.  `angle_incre_is_f32`: the reference divides by
 *        scan_phi[1] - scan_phi[0] in that value's own dtype.
 *    pof_polar_grid_fwd       replaces scans_to_polar_grid, utils.py:492-531: scans [S, N] -> out
 *        [S, R, N], R = int((max_range - min_range) / range_bin_size) + 1.
 * ------------------------------------------------------------------------- */
POF_API int pof_cutout_original_fwd(const float* scans, int B, int S, int N, double angle_incre,
                                    int angle_incre_is_f32, int P, double window_width, double window_depth,
                                    double padding_val, int fixed, int centered, float* out, void* stream);
POF_API int pof_polar_grid_fwd(const float* scans, int S, int N, double min_range, double max_range,
                               double range_bin_size, double tsdf_clip, int normalize, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* POF_H_ */
