/*
 * eamm_b200 -- C ABI of the B200-native EAMM generation hot path.
 *
 * The reference (jixinya/EAMM) is pure Python/PyTorch and has no FFI: its operator API for this
 * path is the nn.Module pair modules/generator.py:8 (OcclusionAwareGenerator) and
 * modules/dense_motion.py:7 (DenseMotionNetwork).  The drop-in Python classes in
 * eamm_b200/modules/ keep that API and call the entry points below through ctypes; every entry
 * point names the reference lines it replaces.  All pointers are DEVICE pointers owned by the
 * caller, nothing is allocated or cached inside, every call is asynchronous on `stream`
 * (a cudaStream_t passed as void*), and the return value is 0 on success, a negative
 * EAMM_ERR_* for a rejected argument, or a positive cudaError_t from the launch.
 *
 * Activation layout ("eamm_act"): NHWC with the channel axis optionally holding 2 bf16 planes
 * (hi, lo) so that value = hi + lo carries 16 mantissa bits; see DESIGN.md "Data layout in HBM".
 * EAMM_F16 with planes == 2 is the mixed operand format of the fp16 + fp8 convolution scheme: plane 0 holds
 * c_buf fp16 values hi = fp16(v * 2^scale_exp), plane 1 (the same 2*c_buf bytes) holds c_buf e4m3 bytes
 * lo8 = e4m3((v * 2^scale_exp - hi) * 64) followed by c_buf e4m3 bytes hi8 = e4m3(hi / 64);
 * value = (hi + lo8 / 64) * 2^-scale_exp (15-16 significant bits inside a 13-octave window below the maximum).
 */
#ifndef EAMM_B200_H
#define EAMM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EAMM_ABI_VERSION 2

enum { EAMM_F32 = 0, EAMM_BF16 = 1, EAMM_F16 = 2 };

enum {
  EAMM_ERR_ARG = -1,       /* null pointer / non-positive size */
  EAMM_ERR_SHAPE = -2,     /* shapes of the operands do not agree */
  EAMM_ERR_DTYPE = -3,     /* dtype / plane combination not supported by this entry point */
  EAMM_ERR_ALIGN = -4,     /* pointer or stride not aligned as the kernel requires */
  EAMM_ERR_UNSUPPORTED = -5
};

/* NHWC activation view.  Element (n, y, x, plane p, channel ch) lives at
 *   data[n*n_stride + (y*w + x)*planes*c_buf + p*c_buf + c_off + ch]        (in elements)
 * A view selects channels [c_off, c_off+c) of a wider buffer, which is how the hourglass skip
 * concatenations (util.py:982-987) are formed without a copy.  n_stride == 0 broadcasts one image
 * over the whole batch (shared source). */
typedef struct {
  void* data;
  int32_t dtype;     /* EAMM_F32 (planes must be 1), EAMM_BF16 (planes 1 or 2) or EAMM_F16 (planes 1, or 2 = fp16 + 2 x e4m3) */
  int32_t n, h, w;
  int32_t c;
  int32_t c_off;
  int32_t c_buf;
  int32_t planes;
  int64_t n_stride;
  int32_t scale_exp; /* EAMM_F16 only: stored = value * 2^scale_exp (per-tensor power-of-two pre-scale chosen by the host
                        from a calibration pass so that the e4m3 planes sit inside their exponent range); 0 otherwise */
  int32_t reserved;
} eamm_act;

/* Keypoints of one side (driving or source): value [n,K,2] fp32 (x,y) and optionally
 * jacobian [n,K,2,2] fp32 (NULL when the caller's dict has no 'jacobian' key,
 * dense_motion.py:55).  n_stride_* == 0 broadcasts one keypoint set over the batch. */
typedef struct {
  const float* value;
  const float* jacobian;
  int64_t value_stride;     /* elements between batch items, K*2 or 0 */
  int64_t jacobian_stride;  /* K*4 or 0 */
} eamm_kp;

/* ---- conv kinds / epilogue flags for eamm_conv_* -------------------------------------------- */
enum {
  EAMM_CONV_3X3 = 0,       /* 3x3, padding 1 (util.py:865-866, 889-890, 909-910)            */
  EAMM_CONV_7X7 = 1,       /* 7x7, padding 3 (generator.py:25,46; dense_motion.py:18,21)     */
  EAMM_CONV_UP2_3X3 = 2,   /* F.interpolate(x2, nearest) then 3x3 pad 1 (util.py:895-897),
                              executed as four 2x2 convolutions on the low-res input, one per
                              output-pixel parity class; weights are pre-combined by the host  */
  EAMM_CONV_ROW7_PACKED = 3 /* 7x7 pad 3 over a <=3-channel image packed by eamm_pack_image
                              (generator.py:25 `first`): K = (ky) x [8 pixels x 8 channels], the
                              kx taps and both bf16 planes live inside one 64-wide K window;
                              eamm_conv_tc only                                                 */
};
enum {
  EAMM_EPI_RELU = 1,       /* y = max(y, 0) after the folded-BN bias                          */
  EAMM_EPI_POOL2 = 2,      /* 2x2 average pool after the activation (util.py:913,919)         */
  EAMM_EPI_SIGMOID = 4     /* y = sigmoid(y) (generator.py:93) -- only with out_nchw          */
};

typedef struct {
  int32_t kind;            /* EAMM_CONV_*                                                      */
  int32_t flags;           /* EAMM_EPI_*                                                       */
  int32_t cin, cout;
  const eamm_act* in;      /* input view, c == cin                                             */
  const void* weight;      /* layout depends on the implementation, see each entry point       */
  const float* bias;       /* [cout] fp32, conv bias with eval-BN folded in                    */
  const eamm_act* residual;/* optional: y += residual (util.py:879), same shape as out         */
  const eamm_act* out;     /* optional NHWC output view (c == cout)                            */
  const eamm_act* out2;    /* optional second NHWC output: relu(y*scale2 + shift2), i.e. the
                              next ResBlock2d's norm1+relu (util.py:873-874) fused here         */
  const float* scale2;     /* [cout] fp32                                                      */
  const float* shift2;     /* [cout] fp32                                                      */
  float* out_nchw;         /* optional fp32 NCHW output [n, out_nchw_c, H, W] (final prediction)*/
  int32_t out_nchw_c;      /* channels written to out_nchw (<= cout; cout may be padded)       */
  float* out_nhwc_f32;     /* optional fp32 NHWC raw output [n, H, W, cout] (mask/occ logits)  */
  int32_t pack_passes;     /* EAMM_CONV_ROW7_PACKED only: 1 (bf16) or 2 (hi/lo split) weight passes */
  int32_t weight_fold;     /* eamm_conv_tc: the scheme `weight` was packed for, must equal
                              eamm_conv_tc_fold(...) for this layer (0 when not folded)         */
  uint8_t* out_u8_nhwc;    /* optional, only together with out_nchw: [n, H, W, out_nchw_c] uint8 frames,
                              clip(rint(y * 255), 0, 255) of the value written to out_nchw -- skimage
                              img_as_ubyte of a float image in [0,1] (demo.py:281,507) */
  const float* acc_scale;  /* optional [cout] fp32, eamm_conv_tc only: the accumulator is multiplied by acc_scale[co] before the
                              bias is added.  With EAMM_F16 operands the host passes 2^-(in.scale_exp + weight exponent of
                              row co); NULL = 1                                                                          */
  float* amax_out;         /* optional device float, eamm_conv_tc only: atomically raised to max |value| written to `out`
                              (the value before the 2^scale_exp pre-scale): the host's running calibration statistic   */
  float* amax_out2;        /* same for `out2`                                                                          */
  void* splitk_ws;         /* optional, eamm_conv_tc only: device workspace for split-K on layers whose output has too
                              few tiles to occupy the chip (hourglass 8x8 ... 2x2 maps).  The first 4 KiB are arrival
                              counters: zero them ONCE after allocation, the kernel leaves them zero.  Launches that
                              may run concurrently must not share a workspace.  NULL = never split                    */
  int64_t splitk_ws_bytes; /* size of splitk_ws; EAMM_SPLITK_WS_BYTES always suffices                               */
} eamm_conv_args;
#define EAMM_SPLITK_WS_BYTES (4096 + 160ll * 128 * 256 * 4)

/* ---- library info --------------------------------------------------------------------------- */
int eamm_abi_version(void);
/* 1 when the visible device is sm_100 and the kernels in this library can run on it. */
int eamm_device_ok(int device);

/* ---- a3: AntiAliasInterpolation2d (util.py:1044-1052; call site dense_motion.py:83) ---------
 * src [n,3,H,W] fp32 NCHW -> dst [n,H/step,W/step,4] fp32 (RGB0 per pixel).  Zero pad taps/2, taps x taps
 * Gaussian (normalised), subsample ::step (step = int(1/scale_factor)).  g1 is the `taps`-tap 1-D factor (the 2-D
 * buffer of the reference is its normalised outer product); the reference hard-codes sigma = 1.5, i.e. 13 taps, at
 * every scale (util.py:1011-1013); any odd taps <= 13 is accepted.  taps = 1, g1 = {1}, step = 1 is the plain
 * NCHW -> RGB0 copy used when scale_factor == 1 (dense_motion.py:28-30,82: no `down` module then). */
int eamm_aa_downsample(const float* src, int64_t src_n_stride, float* dst, int n, int H, int W,
                       int step, const float* g1, int taps, void* stream);

/* Same filter, written as channels [0,4) = (R,G,B,0) of an NHWC activation view (the keypoint detector's
 * hourglass input, keypoint_detector.py:78-79); remaining channels of the view are left untouched. */
int eamm_aa_downsample_act(const float* src, int64_t src_n_stride, int n, int H, int W, int step,
                           const float* g1, int taps, const eamm_act* dst, void* stream);

/* ---- keypoint heads (keypoint_detector.py:40-50, 82-103, 183-203) --------
 * logits [n,h,w,ldl] fp32 NHWC of a SAME-padded 7x7 conv over the detector's feature map: channels [0,K) =
 * `kp` logits, [K, K+4J) = `jacobian` maps (J = K, 1 for single_jacobian_map, 0 without).  The reference
 * convolves with padding `pad` (0 in every config); its (h-6+2pad) x (w-6+2pad) output is the window of the
 * same-padded map at offset 3-pad.  Per (image, keypoint): heatmap = softmax(logit / temperature) over the
 * window; value = sum heatmap * make_coordinate_grid(window); jacobian = sum heatmap * map.
 * Outputs: heatmap [n,K,hh,ww] (may be NULL), value [n,K,2], jacobian [n,K,2,2] (NULL when J == 0). */
int eamm_kp_head(const float* logits, int ldl, int n, int h, int w, int num_kp, int num_jac_maps, int pad,
                 float temperature, float* heatmap, float* value, float* jacobian, void* stream);

/* ---- per-clip keypoint glue between the detector and the generator ------
 * One-Euro smoothing of the detector outputs over the T frames of a clip (filter1.py:14-47 as driven by
 * demo.py:231-248), emotion-row accumulation (demo.py:263-271: kp row emo_rows[2r] += gain[r] * emotion row
 * emo_rows[2r+1]) and normalize_kp with relative movement/jacobian (demo.py:112-132; `movement_scale` is the
 * ConvexHull ratio of :114-117, computed on the host once per clip).  All arrays fp32 on the device:
 * drv_value [T,K,2], drv_jac [T,K,2,2], emo_* [T,Ke,...] or NULL, src_/init_ [K,...]; outputs [T,K,2], [T,K,2,2];
 * emo_scratch [T*Ke*6].  Removes the per-frame device<->host round trips of demo.py:235-248 and lets the
 * generator run batched over T. */
typedef struct { float mincutoff, beta, dcutoff, freq, scale; } eamm_one_euro;
int eamm_kp_clip(const float* drv_value, const float* drv_jac, const float* emo_value, const float* emo_jac,
                 int T, int K, int Ke, const eamm_one_euro* f_kp, const eamm_one_euro* f_emo,
                 const int32_t* emo_rows, const float* emo_gain, int n_emo_rows, const float* src_value,
                 const float* src_jac, const float* init_value, const float* init_jac, float movement_scale,
                 int relative, float* out_value, float* out_jac, float* emo_scratch, void* stream);

/* ---- a4+a5+a6: heatmaps, sparse motions, deformed source (dense_motion.py:32-79, util.py:815-855)
 * small [n,h,w,4] fp32 from eamm_aa_downsample (small_n_stride 0 = shared source);
 * writes the (K+1)*4-channel hourglass input (channel order [hm_k,R_k,G_k,B_k], dense_motion.py:93-94,
 * remaining channels of the view zero-filled) and sparse_deformed [n,K+1,3,h,w] fp32 NCHW.
 * status (device int32, may be NULL) gets bit 0 set when a driving Jacobian is singular
 * (torch.inverse would raise, dense_motion.py:56). */
int eamm_kp_stage(const float* small_img, int64_t small_n_stride, const eamm_kp* kp_driving,
                  const eamm_kp* kp_source, int num_kp, float kp_variance, const eamm_act* hg_in,
                  float* sparse_deformed, int32_t* status, void* stream);

/* ---- a8 epilogue: softmax over K+1 mask logits, flow combine, sigmoid occlusion
 * (dense_motion.py:98-111).  logits [n,h,w,ldl] fp32 NHWC with channels [0,K] = mask logits and
 * channel K+1 = occlusion logit (has_occ).  Sparse motions are recomputed from the keypoints.
 * Outputs: mask [n,K+1,h,w] fp32, deformation [n,h,w,2] fp32, occlusion [n,1,h,w] fp32. */
int eamm_flow_combine(const float* logits, int ldl, const eamm_kp* kp_driving, const eamm_kp* kp_source,
                      int num_kp, int has_occ, int n, int h, int w, float* mask, float* deformation,
                      float* occlusion, void* stream);

/* ---- a9-i: grid_sample(features, deformation) * occlusion (generator.py:57,79-84), fused with
 * the first ResBlock2d's norm1+relu (out2, optional).  feat/out/out2 are NHWC views with equal
 * h,w,c; deformation [n,fh,fw,2]; occlusion [n,1,fh,fw] or NULL.  When the motion grid (fh, fw) differs from the
 * feature grid, flow and occlusion are resized on the fly exactly as generator.py:53-56 / :82-83 do
 * (F.interpolate bilinear, align_corners=False); fh = fw = 0 means "same as feat".  deformation == NULL (a generator
 * built with dense_motion_params=None, generator.py:67): no warp, out = feat.  amax_out2 (device float, may be NULL) is atomically
 * raised to max |out2 value|: the calibration statistic of an EAMM_F16 second output (see eamm_act.scale_exp). */
int eamm_warp_occlude(const eamm_act* feat, const float* deformation, const float* occlusion, int fh, int fw,
                      const eamm_act* out, const eamm_act* out2, const float* scale2,
                      const float* shift2, float* amax_out2, void* stream);

/* ---- a9-ii: 'deformed' = grid_sample(source, bilinear_upsample(deformation)) (generator.py:50-57,86)
 * src [n,C,H,W] fp32 NCHW, deformation [n,h,w,2] fp32 -> dst [n,C,H,W] fp32 NCHW. */
int eamm_warp_image(const float* src, int64_t src_n_stride, const float* deformation, float* dst,
                    int n, int C, int H, int W, int h, int w, void* stream);

/* ---- source image NCHW fp32 -> NHWC activation view (channels beyond C zero-filled) ---------- */
int eamm_nchw_to_act(const float* src, int n, int C, int H, int W, const eamm_act* dst, void* stream);

/* ---- convolutions (a1, a2, a7, a8-conv, a10, a11, a12) --------------------------------------
 * eamm_conv_simt: fp32 CUDA-core implicit GEMM; weight fp32 [taps][cin][cout] (UP2: [4 classes][4 taps]).
 *                 Accepts F32 and BF16 (1 or 2 planes) activations.  Exact-fp32 parity path. */
int eamm_conv_simt(const eamm_conv_args* args, void* stream);

/* eamm_conv_tc:   tcgen05/TMEM/TMA implicit GEMM (bf16 / fp16 / fp16+fp8 operands, fp32 accumulate).  Activations
 *                 must be EAMM_BF16 or EAMM_F16 with c_buf, c_off and cin multiples of 64 and power-of-two h, w; cout a
 *                 multiple of 16.  weight is bf16 [classes*cout][passes*taps*cin] (K contiguous),
 *                 K ordered (pass, tap, channel); passes = 1 for single-plane inputs, 3 for hi/lo
 *                 inputs (weight planes lo, hi, hi against activation planes hi, lo, hi: the two
 *                 cross terms first, the dominant hi*hi term last).  UP2 has
 *                 4 classes of 4 taps, class c occupying rows [c*cout, (c+1)*cout).
 *                 EAMM_F16 single-plane inputs: the same with fp16 weights (one pass).
 *                 EAMM_F16 two-plane (mixed) inputs, cin a multiple of 128: `weight` is a byte matrix
 *                 [classes*cout][taps*cin*4]: e4m3 lo8 [tap][cin] | e4m3 hi8 [tap][cin] | fp16 hi [tap][cin], of
 *                 w * 2^e_row split like the activations (include/eamm_b200.h top).  The K loop runs
 *                 a_hi8 x w_lo8 and a_lo8 x w_hi8 as tcgen05.mma.kind::f8f6f4 steps (K = 32, twice the fp16 rate) and
 *                 a_hi x w_hi as kind::f16 steps into one TMEM accumulator: two pass-equivalents instead of the three
 *                 bf16 passes of hi/lo inputs.  acc_scale[co] = 2^-(in.scale_exp + e_row) undoes the pre-scales.
 *                 cin == 64 is accepted when the view is the whole 64-channel buffer (c_off 0, c_buf 64): the pixel's
 *                 [lo8 | hi8] bytes are then one 128-byte K chunk and the weight bytes of a tap are
 *                 e4m3 [hi8 x 64 | lo8 x 64], all taps, followed by fp16 hi [tap][64].
 *                 3x3 / UP2 kinds only (no 7x7 schemes, no fold). */
int eamm_conv_tc(const eamm_conv_args* args, void* stream);

/* Which scheme eamm_conv_tc uses for a 7x7 layer, i.e. which weight matrix it expects:
 *   0  one TMA tile per tap           bf16 [cout][passes * 49 taps * cin]
 *   1  halo row, kx-shifted views     bf16 [7 kx * cout][passes * 7 ky * cin]      (w % 128 == 0, cout <= 32)
 *   2  kx taps in the GEMM N axis     bf16 [32 = kx*4 + co][passes * 7 ky * cin]   (w % 128 == 0, only
 *      out_nchw with <= 4 channels; out_nchw_c = 0 when the call has any NHWC output)
 * eamm_conv_tc_query may also report (EAMM_TC_KXW switch, see conv_tc.cu):
 *   3  as 2 with four output rows per tile: bf16 [112 = dr*28 + kx*4 + co][passes * 10 input rows * cin],
 *      row block j of output row dr = w[ky = j - dr] (zero outside the filter)         (h % 4 == 0)
 *   4  kx in N, full-width tiles      bf16 [112 = kx*16 + co][passes * 7 ky * cin]    (cout == 16, w <= 128,
 *      only out_nhwc_f32, no flags) */
int eamm_conv_tc_uses_halo(int kind, int w, int cout, int out_nchw_c);

/* Split (hi/lo) layers with cout <= 128 (and the kx-in-N 7x7 scheme) stack the two weight planes along N:
 *   0  not folded: K = (pass, tap, channel) as described at eamm_conv_tc
 *   1  rows [hi block (classes*cout) | lo block], K = (tap, channel); per (tap, 64 channels) the kernel runs
 *      a_lo x b_hi (N = bn; all of them first) and a_hi x [b_hi; b_lo] (N = 2*bn); the epilogue adds the halves
 *   2  EAMM_CONV_ROW7_PACKED with pack_passes == 2: rows [w_hi vs (a_hi, a_lo) | w_lo vs a_hi], K = (ky, 64)
 * `split` = input has hi/lo planes (or pack_passes == 2); `halo_scheme` = eamm_conv_tc_uses_halo(...). */
int eamm_conv_tc_fold(int kind, int split, int cout, int halo_scheme);

/* Planning dry run of eamm_conv_tc for `args` (weight_fold ignored, nothing launched):
 * out[0] = N tile, out[1] = 7x7 scheme (eamm_conv_tc_uses_halo), out[2] = fold (eamm_conv_tc_fold),
 * out[3] = K chunks (halo-tile scheme: filter taps) per pipeline stage, out[4] = bit 0 CTA pairs (cta_group::2), bit 1
 * halo-tile scheme (3x3 / UP2 with EAMM_F16 inputs on maps whose 8 x 16-pixel tiles fill the chip: one 10 x 18 halo tile per
 * K chunk in shared memory, the taps are descriptor views of it), bits 8.. = split-K factor (1 = none; planned as if a
 * workspace were given),
 * out[5] = pipeline stages (`out` has room for 6 ints).  The host packs the weights for out[1]/out[2]. */
int eamm_conv_tc_query(const eamm_conv_args* args, int* out);

/* ---- source image for EAMM_CONV_ROW7_PACKED: src [n,C<=3,H,W] fp32 NCHW -> dst bf16
 * [n][H+6][W+8][8] with channels [hi0,hi1,hi2,0,lo0,lo1,lo2,0] (lo = bf16(v-hi); zero when
 * split == 0; split == 2 writes fp16 values instead of bf16, lo zero: the EAMM_F16 single-plane path).  The zero border (3 rows top/bottom, 3 columns left, 5 right) must already be
 * zero in dst (allocate it zeroed once); only the interior is written. */
int eamm_pack_image(const float* src, int n, int C, int H, int W, int split, void* dst, void* stream);

/* ---- AT_net2, the per-clip audio -> motion-feature network (replaces the ATen ops behind
 * /root/reference/modules/util.py:580-613; its convolutions run through eamm_conv_simt: DownBlock2d as 3x3+POOL2,
 * ConvTranspose2d(k=4, s=2, p=1) as EAMM_CONV_UP2_3X3 with parity-class weights W[cin][cout][3-a-2ty][3-b-2tx]).
 *
 * eamm_linear: y[m][n] = scale * act(sum_k x[m][k] w[k][n] + bias[n] + add_rows[m / add_period][n])
 *   nn.Linear (util.py:532-556), the LSTM input projections (util.py:557) and the 1x1 -> 4x4 ConvTranspose2d
 *   (util.py:560).  x [M][ldx] fp32, w [K][N] fp32 (the transpose of nn.Linear.weight), bias [N] or NULL,
 *   add_rows [ceil(M/add_period)][N] or NULL, y [M][ldy]; N and ldy multiples of 4; relu != 0 applies ReLU. */
int eamm_linear(const float* x, int ldx, const float* w, const float* bias, const float* add_rows, int add_period,
                float* y, int ldy, int M, int K, int N, int relu, float scale, void* stream);

/* eamm_maxpool: nn.MaxPool2d(k, stride=(stride_y, stride_x)), no padding, floor mode (util.py:543,547). */
int eamm_maxpool(const eamm_act* in, const eamm_act* out, int k, int stride_y, int stride_x, void* stream);

/* eamm_act_copy: dst(n, y < h, x < w, :) = src(n, y, x, :) between two views of any map size / storage format (same n and c);
 * zero_rest != 0 also clears every other pixel of dst (src == dst: only that).  No reference counterpart: AT_net2's MFCC
 * convolutions (util.py:541-546, 28x12 and 26x5 maps) run on tensor cores over maps padded to powers of two; this moves
 * activations into / out of the padded buffers and restores the zero padding between layers. */
int eamm_act_copy(const eamm_act* src, const eamm_act* dst, int h, int w, int zero_rest, void* stream);

/* eamm_lstm_layer: the recurrence of one nn.LSTM layer (util.py:557,597; gate order i,f,g,o; zero initial state)
 * over whole sequences.  gates_x [B][T][4*hidden] = W_ih x_t + b_ih + b_hh (eamm_linear), w_hh [4*hidden][hidden]
 * (nn.LSTM.weight_hh_l*, as stored), h_out [B][T][hidden].  hidden must be 256.  One 8-CTA cluster per sequence. */
int eamm_lstm_layer(const float* gates_x, const float* w_hh, float* h_out, int B, int T, int hidden, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* EAMM_B200_H */
