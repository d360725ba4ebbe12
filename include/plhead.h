/*
 * plhead.h — C ABI of libplhead.so: the B200-native (sm_100a) PixelLink / EAST
 * per-pixel text-detection head.
 *
 * Drop-in boundary for the hot path of BowieHsu/tensorflow_ocr.
 * The reference has no native code; its operator boundary is "Python callable
 * handed numpy arrays" (tf.py_func, tool/pixellink_fn.py:114,157) plus plain
 * Python functions building TF ops.  Each entry point below names the reference
 * function (file:line under the reference root) whose arithmetic it replaces.
 *
 * Conventions
 *  - All tensor pointers are DEVICE pointers, NHWC, C-contiguous, fp32 unless
 *    noted, 16-byte aligned.  The caller owns every buffer; the library
 *    allocates nothing and keeps no state between calls.
 *  - `stream` is a cudaStream_t passed as void*.  Calls only enqueue work; they
 *    never synchronise the host.  Re-entrant for distinct (stream, workspace).
 *  - Return value: 0 ok; negative = argument / shape / alignment / workspace
 *    error (PLH_E_*); positive = cudaError_t from a launch.  Nothing throws.
 *  - Outputs are written in full (never accumulated into).
 *  - NaN produced by empty link classes is DATA, not an error (nets/model.py:252-253).
 *
 * Numerical contract (what tests/ assert against the reference-executed goldens and the oracle)
 *  - Decisions are bit-exact: OHEM masks, integer normalisers, threshold flags, component labels
 *    (minimum pixel index), integer box corners, restore_rectangle's row order.
 *  - Loss scalars: |a - b| <= 1e-5 * |b|.
 *  - Gradients, ELEMENTWISE: |a - b| <= 1e-5 * |b| + 4 * 2^-24 * w, where w is that element's weight
 *    (gradient = w * (softmax - onehot); w = 2*M/n_seg_pos for pixel logits, M*(1/sum_Wp or 1/sum_Wn) for
 *    link logits).  The absolute term is the fp32 rounding of `softmax - onehot` that the reference itself
 *    (TF autodiff) carries on saturated pixels; the kernels evaluate the same quantity as a sigmoid of the
 *    logit difference, which has no cancellation.  tests/util.py:grad_close states this check; the
 *    max-norm check |a - b|_inf <= 1e-5 * |b|_inf (tests/util.py:rel_err) is asserted as well.
 *  - One host thread per device at a time per (stream, workspace); several devices may be driven from one
 *    process (kernel attributes are set per device).
 */
#ifndef PLHEAD_H_
#define PLHEAD_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PLH_VERSION 100 /* 0.1.0 */

#if defined(__GNUC__)
#define PLH_API __attribute__((visibility("default")))
#else
#define PLH_API
#endif

/* error codes */
#define PLH_OK 0
#define PLH_E_NULL (-1)      /* required pointer is NULL */
#define PLH_E_SHAPE (-2)     /* B/H/W/N out of range */
#define PLH_E_ALIGN (-3)     /* pointer not 16-byte aligned */
#define PLH_E_WORKSPACE (-4) /* workspace too small / NULL */
#define PLH_E_PARAM (-5)     /* bad enum / parameter value */
#define PLH_E_DEVICE (-6)    /* not an sm_100 device, or no CUDA device */

/* ---- ops, for plh_workspace_bytes ---- */
#define PLH_OP_LOSS 0
#define PLH_OP_DECODE 1
#define PLH_OP_DICE 2
#define PLH_OP_EAST_LOSS 3
#define PLH_OP_RESTORE 4
#define PLH_OP_LOSS_DECODE 5 /* fused loss + decode sharing one logits read */

/* ---- loss variants (which reference function the weights follow) ---- */
#define PLH_VARIANT_MODEL 0     /* nets/model.py:204-261 loss(): OHEM 3:1, link CE x OHEM mask   */
#define PLH_VARIANT_POS_ONLY 1  /* nets/model_vgg_16.py:243-282 ohem_loss(): positives only      */
#define PLH_VARIANT_PIXELLINK 2 /* nets/pixellink.py:88-263 build_loss(): mean pixel CE, guarded */
/* per-pixel term */
#define PLH_TERM_CE 0    /* softmax cross entropy */
#define PLH_TERM_FOCAL 1 /* focal loss, Lin et al. 2017 (NOT in the reference; parity unpinned) */

typedef struct plh_loss_params {
  int32_t variant;       /* PLH_VARIANT_* */
  int32_t term;          /* PLH_TERM_* */
  int32_t neg_pos_ratio; /* 3 (nets/model.py:171; config.max_neg_pos_ratio in nets/pixellink.py:116) */
  float focal_alpha;     /* 0.25 */
  float focal_gamma;     /* 2.0 */
  int32_t reserved[3];   /* reserved[0] bit 0: run only the main pass, on a workspace that a previous call with the
                            same inputs prepared (skips selection, mask and normalisers): lets bench.py time
                            the bandwidth-bound pass alone.  Bit 1: scheduling hint, results identical — keep
                            the mask + normaliser pass a separate kernel instead of fusing it into the
                            selection kernel.  Bit 2: scheduling hint — the kernel that precedes this call on
                            the stream is one of this library's (plh_decode phase), so the first kernel here
                            may launch programmatically under its tail.  Other bits must be 0. */
} plh_loss_params;

/* Layout of the `stats` output (device floats).  PLH_STATS_FLOATS + B entries. */
#define PLH_STATS_FLOATS 64
#define PLH_ST_TOTAL 0       /* returned loss: link + 2*pixel (model.py:261)               */
#define PLH_ST_L_PIX 1       /* classification_loss (model.py:233) / pixel_cls_loss (pixellink.py:160) */
#define PLH_ST_L_LINK 2      /* [8] per-direction link loss (model.py:254)                 */
#define PLH_ST_N_SEG_POS 10  /* n_seg_pos (model.py:221; pixellink.py:155 = sum of selected) */
#define PLH_ST_SUM_WP 11     /* [8] link_pos_n (model.py:249)                              */
#define PLH_ST_SUM_WN 19     /* [8] link_neg_n (model.py:250)                              */
#define PLH_ST_S_PIX 27      /* sum(term * selected_mask)                                  */
#define PLH_ST_S_POS 28      /* [8] sum(term * W_link_pos)                                 */
#define PLH_ST_S_NEG 36      /* [8] sum(term * W_link_neg)                                 */
#define PLH_ST_LINK_TOTAL 44 /* weight_link_loss (model.py:256)                            */
#define PLH_ST_N_SELECTED 45 /* number of pixels in the OHEM mask                          */
#define PLH_ST_THR 64        /* [B] per-image OHEM threshold score, NaN = none selected    */

/*
 * PixelLink loss forward + backward, one fused pipeline.
 * Replaces (per `variant`): nets/model.py:204-261 `loss` incl. OHNM_batch :186-197,
 * OHNM_single_image :161-184, get_pos_and_neg_masks :199-202, slim.softmax :216;
 * nets/model_vgg_16.py:243-282 `ohem_loss` + cal_link_loss :227-241;
 * nets/pixellink.py:88-263 `PixelLinkNet.build_loss`; and TF autodiff of those.
 *
 *  pix_logits [B,H,W,2]  link_logits [B,H,W,16]  pix_lab [B,H,W(,1)]  link_lab [B,H,W,8]
 *  train_mask: accepted, never read (nets/model.py ignores it), may be NULL.
 *  stats      [PLH_STATS_FLOATS + B] floats (required)
 *  grad_pix   [B,H,W,2]  / grad_link [B,H,W,16]: d loss / d logits for upstream
 *             gradient 1.0; both NULL = forward only; otherwise both required.
 *  ohem_mask  [B,H,W] uint8, optional, 16-byte aligned (the main pass fetches it with bulk copies):
 *             pixel_selected_mask (model.py:220).
 *  decode_flags [B,H,W] uint16, optional: by-product for plh_decode_from_flags —
 *             bit d (0..7) = link_d score > link_thresh, bit 8 = pixel score >
 *             pixel_thresh (thresholds from `dp`, which may be NULL iff
 *             decode_flags is NULL).
 */
struct plh_decode_params;
PLH_API int plh_pixellink_loss(const float* pix_logits, const float* link_logits, const float* pix_lab,
                       const float* link_lab, const float* train_mask, int B, int H, int W,
                       const plh_loss_params* p, float* stats, float* grad_pix, float* grad_link,
                       uint8_t* ohem_mask, uint16_t* decode_flags, const struct plh_decode_params* dp,
                       void* workspace, size_t workspace_bytes, void* stream);

/*
 * OHEM selection only (nets/model.py:186-197 OHNM_batch with :161-184, or the
 * nets/pixellink.py:106-150 variant when variant == PLH_VARIANT_PIXELLINK).
 *  scores  [B,N] softmax probability of the NEGATIVE class
 *  pos/neg [B,N] uint8 masks
 *  n_pos   [B] int32 optional (device): overrides the count of pos_mask, as in
 *          OHNM_single_image(scores, n_pos, neg_mask) where n_pos is an argument
 *  selected_mask [B,N] float: pos + selected negatives;  thr [B] float.
 *  workspace: plh_workspace_bytes(PLH_OP_LOSS, B, 1, N, 0).
 */
PLH_API int plh_ohnm_batch(const float* scores, const uint8_t* pos_mask, const uint8_t* neg_mask,
                           const int32_t* n_pos, int B, int N, int variant, int neg_pos_ratio,
                           float* selected_mask, float* thr, void* workspace, size_t workspace_bytes,
                           void* stream);

/*
 * nets/model.py:145-159 == nets/model_vgg_16.py:179-193 `dice_coefficient`:
 * one scalar over the whole tensor; predictions are PROBABILITIES.
 *  y_true / y_pred / mask [M] floats.  out [4]: loss, loss, I, U.
 *  grad [M] optional: d loss / d y_pred = -2 m (t U - I) / U^2.
 */
PLH_API int plh_dice(const float* y_true, const float* y_pred, const float* mask, long long M, float* out, float* grad,
             void* workspace, size_t workspace_bytes, void* stream);

/*
 * nets/model_vgg_16.py:196-225 `loss`: 2*dice(pixel) + sum_d dice(link_d).
 *  t_pix / p_pix [M,1], t_link / p_link [M,8], mask [M] (broadcast over channels).
 *  out [28]: total, then (dice, I, U) for the pixel channel and the 8 link channels.
 *  grad_pix [M,1] / grad_link [M,8] optional (both or neither).
 */
PLH_API int plh_dice_head(const float* t_pix, const float* p_pix, const float* t_link, const float* p_link,
                  const float* mask, long long M, float* out, float* grad_pix, float* grad_link, void* workspace,
                  size_t workspace_bytes, void* stream);

/* ---- decode ---- */
typedef struct plh_decode_params {
  float pixel_thresh; /* 0.8  (test_pixellink_fast.py:12) strict > */
  float link_thresh;  /* 0.9  (test_pixellink_fast.py:13) strict > */
  int32_t min_size;   /* 10   (test_pixellink_fast.py:174; 200 in test_pixellink.py:177): keep size > min_size */
  int32_t max_boxes;  /* capacity K of the per-image box list */
  double scale_x;     /* 4.0  = 1280/320 (test_pixellink_fast.py:196); must be >= 1 */
  double scale_y;     /* 3.75 = 720/192  (test_pixellink_fast.py:197); must be >= 1 */
  int32_t reserved[2]; /* reserved[0], phase selection on one workspace (results identical to the whole call):
                          bit 0: components + label map only (no boxes); bit 1: boxes only, from the workspace
                          of a previous bit-0 call; bit 2: only the first kernel (thresholds + labelling inside
                          32x16 tiles); bit 3: everything after it, on the workspace of a bit-2 call.  Lets a
                          caller interleave the decode with another pipeline (tensorflow_ocr_b200/head.py).
                          Bit 5: use the resident form of the component labelling (one 8-CTA cluster per image,
                          the map as bit planes in distributed shared memory; PLH_E_SHAPE if a strip of H/8 rows
                          does not fit in 227 KB) instead of the default tiled form.  Same results either way
                          (the tests run both); bit 4 is accepted and means the default. */
} plh_decode_params;

/*
 * PixelLink decode: test_pixellink_fast.py:110-202 (thresholds, directed
 * 8-neighbour link graph from interior pixels, connected components, size
 * filter, per-component cv2.minAreaRect -> cv2.boxPoints -> np.int0).
 * Components are the weakly-connected components of the reference's edge set
 * (SURVEY.md §8a D2), labelled by their minimum linear pixel index.
 *
 *  labels  [B,H,W] int32: -1 background / filtered, else min linear index y*W+x
 *  boxes   [B,K,8] int32: x0,y0,...,x3,y3 in cv2.boxPoints order, components in
 *          ascending label order; only the first min(n_boxes[b],K) rows are written
 *  n_boxes [B] int32: number of components with size > min_size (may exceed K:
 *          then only K are written)
 *  rects   [B,K,5] float optional: cx, cy, w, h, angle of cv2.minAreaRect
 *  comp    [B,K,2] int32 optional: label (min pixel index) and pixel count per box
 * Limits: H <= 1024 rows, W <= 2048 columns, W*scale_x < 32768, H*scale_y < 32768.
 */
PLH_API int plh_decode(const float* pix_logits, const float* link_logits, int B, int H, int W,
               const plh_decode_params* p, int32_t* labels, int32_t* boxes, int32_t* n_boxes, float* rects,
               int32_t* comp, void* workspace, size_t workspace_bytes, void* stream);

/* Same, starting from the uint16 flags map emitted by plh_pixellink_loss. */
PLH_API int plh_decode_from_flags(const uint16_t* flags, int B, int H, int W, const plh_decode_params* p,
                          int32_t* labels, int32_t* boxes, int32_t* n_boxes, float* rects, int32_t* comp,
                          void* workspace, size_t workspace_bytes, void* stream);

/*
 * Polygon rasterisation of the ground-truth generators: tool/pixellink_fn.py:66-79 (cv2.fillPoly of every
 * quadrilateral in order, then cv2.resize(INTER_NEAREST) to (w/4, h/4)) and datasets/icdar.py:486-514 (cv2.fillPoly
 * at full resolution, training mask cleared inside flagged polygons).  Bit-identical to cv2 4.13 (outline by
 * cv::line, scan-line interior, clipping of polygons that leave the canvas); nothing is drawn at full resolution,
 * every output pixel looks at the source pixel it samples.
 *  quads     [sum n_b, 4, 2] int32 (x, y) canvas coordinates (the caller truncates the float vertices like
 *            np.array(points, np.int32) does); quad_off [B+1] int32 (device): first polygon of image b
 *  zero_flags [sum n_b] uint8 optional: polygons that clear the training mask
 *  H, W      canvas; Ho, Wo output grid (Wo <= 2048)
 *  mode 0    output (oy, ox) = canvas (oy*stride, ox*stride)  (stride 1: the canvas itself; 4: [::4, ::4])
 *  mode 1    output = cv2.resize(canvas, (Wo, Ho), INTER_NEAREST)
 *  last_ids  [B,Ho,Wo] int32: 1-based index of the LAST polygon covering the pixel (0 = none) — what cv2 leaves
 *  first_ids [B,Ho,Wo] int32 optional: index of the FIRST polygon covering it (plh_link_labels_icdar)
 *  ids_u8    [B,Ho,Wo] uint8 optional: min(last, 255), the reference's uint8 poly_mask (plh_link_labels)
 *  score     [B,Ho,Wo] float optional: 1.0 where any polygon covers the pixel
 *  training_mask [B,Ho,Wo] uint8 optional: 0 where a flagged polygon covers the pixel, else 1
 */
PLH_API int plh_fill_quads(const int32_t* quads, const int32_t* quad_off, const uint8_t* zero_flags, int B, int H, int W,
                   int Ho, int Wo, int mode, int stride, int32_t* last_ids, int32_t* first_ids, uint8_t* ids_u8,
                   float* score, uint8_t* training_mask, void* stream);

/*
 * Link labels of the EAST-fork generator, datasets/icdar.py:83-105 valid_link + :486-539 generate_rbox (the one
 * train.sh uses), with its quirks kept (Q17: direction names move the other axis, x is tested against h-1,
 * index -1 wraps, a link asks "is the neighbour text so far", polygons filled in order).
 *  last_ids / first_ids [B,H,W] int32: 1-based index of the last / first polygon covering the pixel (0 =
 *            background); H == W (the reference indexes out of range otherwise)
 *  stride    1, or 4 for the [::4, ::4] subsample of icdar.py:632-634 (outputs are [B, ceil(H/stride), ceil(W/stride), .])
 *  link_lab  [B,Ho,Wo,8] float: left, left_down, left_up, right, right_down, right_up, up, down
 *  score     [B,Ho,Wo] float optional: score_map (1 where some polygon covers the pixel)
 */
PLH_API int plh_link_labels_icdar(const int32_t* last_ids, const int32_t* first_ids, int B, int H, int W, int stride,
                          float* link_lab, float* score, void* stream);

/*
 * Contour path of the decode, test.py:182-218: cv2.findContours(mask, RETR_TREE, CHAIN_APPROX_SIMPLE) — hole
 * borders included (quirk Q14) — then per contour cv2.minAreaRect -> cv2.boxPoints -> np.int0, x4, /ratio_w,
 * /ratio_h (assigned into an integer array: truncation, test.py:193-200) and order_points (test.py:24-35, :217).
 *  mask      [B,H,W] uint8 (device), nonzero = text (the output of plh_pixel_detect)
 *  boxes     [B,K,4,2] int32: the ordered box of contour slot k (tl, tr, br, bl)
 *  raw_boxes [B,K,4,2] int32 optional: the box before order_points (what test.py draws)
 *  info      [B,K,6] int32 per slot: scan position y*W+x at which OpenCV's raster scan finds the border, hole
 *            flag, own key, parent key (-1 = top level), first point in `points`, number of points.  Slots are
 *            in arrival order; OpenCV's output order is the pre-order of the border tree with siblings in
 *            reverse scan order (tensorflow_ocr_b200/decode.py:contour_boxes puts them in that order).
 *  n_contours [B] int32: number of borders found (may exceed K: then only K slots are written)
 *  points    [B, 2*H*W, 2] int32 optional: the CHAIN_APPROX_SIMPLE points (x, y) of every border (else they
 *            live in the workspace).  Limits: H <= 1020, W <= 2044; a border with more than 2048 points gets
 *            the sentinel box INT32_MIN.
 *  workspace: plh_contour_workspace_bytes(B, H, W).
 */
PLH_API size_t plh_contour_workspace_bytes(int B, int H, int W);
PLH_API int plh_contour_boxes(const uint8_t* mask, int B, int H, int W, double ratio_w, double ratio_h, int K,
                      int32_t* boxes, int32_t* raw_boxes, int32_t* info, int32_t* n_contours, int32_t* points,
                      void* workspace, size_t workspace_bytes, void* stream);

/*
 * The head's logit producer: one level of the feature fusion that ends both networks,
 * nets/pixellink.py:37-38,56-67 and nets/model.py:14-15,129-141 —
 *     y = bilinear_x2(prev) + act_a(scale_a * (xa Wa) + shift_a) [+ act_b(scale_b * (xb Wb) + shift_b)]
 * optionally followed by  y <- y w_out + b_out  (the last 1x1 convolutions), written either as [.,18] for the next
 * level or split into the pixel and link tensors the loss / decode entry points read.  The 2 pixel and 16 link channels
 * go through together: 18 columns, pixel first.
 *  xa, xb    [B,H,W,Ka] / [B,H,W,Kb] float NHWC feature maps (xb optional: fc7 + conv5_3 share a level); K % 4 == 0
 *  wa, wb    [K,18] float: the 1x1 convolutions' weights, pixel columns 0-1, link columns 2-17
 *  scale, shift [18] optional: per-channel affine after the convolution (shift alone = bias; both = batch norm in
 *            its inference form); relu != 0: ReLU after it (nets/model.py:103-107 arg_scope)
 *  prev      [B,H/2,W/2,18] optional: the previous level, upsampled like tf.image.resize_bilinear(align_corners =
 *            False) does for an exact factor 2
 *  w_out [18,18] (in, out), b_out [18] optional: y <- y w_out + b_out
 *  output    either y18 [B,H,W,18] (a level that feeds the next one) or pix_logits [B,H,W,2] + link_logits
 *            [B,H,W,16] (the last level), independently of w_out: a caller whose fuse convolutions carry no
 *            activation can fold the output matrix into the weights (x (W w_out) + up2(prev w_out)) and apply w_out
 *            one level earlier, on a quarter of the pixels
 *  flags [B,H,W] uint16 optional (last level only, with flag_params = the decode's thresholds): the word
 *            plh_decode_flags would compute from the logits just produced (bit-identical: same logit-space
 *            thresholds on the same fp32 values), so that plh_decode_from_flags can start without reading the 72 B
 *            per pixel of logits again
 * fp32 in, fp32 out, fp32-accurate accumulation (K % 32 == 0: tcgen05 tensor cores, TF32 products with the 3xTF32
 * split, fp32 accumulator in tensor memory; other K: fp32 FMAs): within 1e-5 of an fp64 evaluation relative to the
 * largest logit (the contract tests/test_gpu_headfuse.py states).  wa / wb must be 16-byte aligned like the feature maps.
 */
PLH_API int plh_head_fuse_level(const float* xa, int Ka, const float* wa, const float* scale_a, const float* shift_a, int relu_a,
                        const float* xb, int Kb, const float* wb, const float* scale_b, const float* shift_b, int relu_b,
                        const float* prev, const float* w_out, const float* b_out, int B, int H, int W, float* y18,
                        float* pix_logits, float* link_logits, const plh_decode_params* flag_params, uint16_t* flags,
                        void* stream);

/*
 * Detection evaluation.
 * plh_quad_jaccard — tool/bboxes.py:252-282 np_bboxes_jaccard for every (detection, ground truth) pair of every
 * image: both quadrilaterals rasterised as cv2.drawContours(thickness = -1) draws them (outline by cv::line +
 * scan-line interior, cv2 4.13; no mask is materialised) and iou = (float)((double)|A & B| / (double)|A | B|).
 *  dets      [sum D_b, 4, 2] int32 (x, y), gts [sum G_b, 4, 2] int32; |coordinate| < 2^20 (else the pair gets NaN).
 *            The reference's mask starts at (0, 0): parts of a box at negative coordinates are clipped the way
 *            cv2 clips them (cv::clipLine on the outline, edge slopes from the clipped segments)
 *  det_off / gt_off [B+1] int32 (device): first detection / ground truth of image b
 *  pair_off  [B+1] int64 (device): pair_off[b] = sum_{i<b} D_i * G_i; total_pairs = pair_off[B]
 *  jaccard   [total_pairs] float: image b's D_b x G_b matrix, row-major, at pair_off[b]
 * plh_bboxes_matching — tool/bboxes.py:158-246: detections of an image in the given (score) order, each against
 * the FIRST maximum of its Jaccard row: match = iou > matching_threshold; true positive = matched, not ignored,
 * ground truth not matched before; false positive = not ignored and (unmatched or already matched).
 *  gignored  [sum G_b] uint8; gmatch [sum G_b] uint8 scratch; tp / fp [sum D_b] uint8;
 *  n_gbboxes [B] int32 = number of not-ignored ground truths.  An image without ground truth gets tp = fp = 0
 *            (the reference raises on it).
 */
PLH_API int plh_quad_jaccard(const int32_t* dets, const int32_t* gts, const int32_t* det_off, const int32_t* gt_off,
                     const int64_t* pair_off, int B, long long total_pairs, float* jaccard, void* stream);
PLH_API int plh_bboxes_matching(const float* jaccard, const int32_t* det_off, const int32_t* gt_off, const int64_t* pair_off,
                        int B, const uint8_t* gignored, float matching_threshold, uint8_t* gmatch, uint8_t* tp,
                        uint8_t* fp, int32_t* n_gbboxes, void* stream);

/*
 * The threshold pass of the decode alone (test_pixellink_fast.py:120-128): flags[b,y,x] = bit 8: pixel
 * score > pixel_thresh, bits 0..7: link d score > link_thresh, for plh_decode_from_flags.  Lets a caller
 * schedule this bandwidth-bound pass separately from the latency-bound component labelling.
 */
PLH_API int plh_decode_flags(const float* pix_logits, const float* link_logits, int B, int H, int W,
                             const plh_decode_params* p, uint16_t* flags, void* stream);

/*
 * cv2.minAreaRect -> cv2.boxPoints -> np.int0 for explicit point lists
 * (test_pixellink_fast.py:199-200, test.py:190-191), one CTA per list.
 *  pts [total,2] int32 (x,y) in the caller's order; offsets [n_sets+1] int32
 *  boxes [n_sets,8] int32; rects [n_sets,5] float optional.  Each set <= 2048 points.
 */
PLH_API int plh_min_area_boxes(const int32_t* pts, const int32_t* offsets, int n_sets, int32_t* boxes, float* rects,
                       void* stream);

/*
 * tool/pixellink_fn.py:120-154 `pixel_detect`: mask = score > thr_p AND, for all
 * 8 directions, link score >= thr_l.   Inputs are softmax PROBABILITIES:
 *  score [H,W] (score_map[0,:,:,0]);  link [8,H,W,2] (geo_map[:,0]); out [H,W] uint8.
 */
PLH_API int plh_pixel_detect(const float* score, const float* link, int H, int W, float score_map_thresh,
                     float link_thresh, uint8_t* out, void* stream);

/*
 * datasets/icdar.py:410-483 `restore_rectangle_rbox`.
 *  origin [N,2] fp32, geometry [N,5] fp32 (top,right,bottom,left,theta)
 *  out [N,4,2] fp64, rows ordered theta>=0 first then theta<0 (icdar.py:479),
 *  out_index [N] int32 optional: input row of each output row.
 */
PLH_API int plh_restore_rectangle(const float* origin, const float* geometry, int N, double* out, int32_t* out_index,
                          void* workspace, size_t workspace_bytes, void* stream);
/* Same with either input in float64 (numpy computes the distance sums and cos/sin in geometry's dtype and the
 * rest in float64; a float64 origin enters the final translation only).  fp32/fp32 is bit-exact against
 * numpy; with a float64 geometry the device cos/sin may differ from the host libm in the last bit. */
PLH_API int plh_restore_rectangle_ex(const void* origin, int origin_is_f64, const void* geometry, int geometry_is_f64,
                             int N, double* out, int32_t* out_index, void* workspace, size_t workspace_bytes,
                             void* stream);


/*
 * EAST RBOX loss fwd+bwd — NOT in the reference; restated
 * from upstream argman/EAST model.loss.  score [M] prob, geo [M,5].
 *  out [8]: total, dice, L_g mean, I, U, sum(L_aabb*w), sum(L_theta*w), reserved
 */
PLH_API int plh_east_loss(const float* score_gt, const float* score_pred, const float* geo_gt, const float* geo_pred,
                  const float* mask, long long M, float* out, float* grad_score, float* grad_geo,
                  void* workspace, size_t workspace_bytes, void* stream);

/*
 * Locality-aware NMS — NOT in the reference; restated from upstream argman/EAST
 * locality_aware_nms.nms_locality: row-major fold merging consecutive overlapping boxes by
 * score-weighted averaging, then standard NMS.  Parity unpinned (checked against oracle/east.py).
 *  polys [total,9] fp64 (8 coords + score) — the boxes of image b are rows offsets[b]..offsets[b+1]
 *  out   [total,9] fp64: survivors of image b start at row offsets[b], n_out[b] of them, by descending score
 *  workspace >= align256(total*72) + total*4 + 256 bytes
 */
PLH_API int plh_lanms(const double* polys, const int32_t* offsets, int B, int total, double thres, double* out,
                      int32_t* n_out, void* workspace, size_t workspace_bytes, void* stream);

/* ---- housekeeping ---- */
/* `K` = plh_decode_params.max_boxes for the decode ops, ignored otherwise */
PLH_API size_t plh_workspace_bytes(int op, int B, int H, int W, int K);
PLH_API int plh_version(void);
PLH_API const char* plh_strerror(int code);
/*
 * Link (and pixel) labels from a polygon-id map — the GPU part of generate_rbox
 * (tool/pixellink_fn.py:81-111 with valid_link :9-47).
 *  ids      [B,H,W] uint8: 0 background, i = polygon i (what cv2.fillPoly(poly_mask, poly, idx+1) and the
 *           INTER_NEAREST resize to (w/4, h/4) leave, :76-79)
 *  link_lab [B,H,W,8] float: for a pixel of polygon v: 1 in all directions on the map border (:10-11), else
 *           1 where the neighbour in that direction has id v; 0 for background.  Channel order as the
 *           reference: left, left_down, left_up, right, right_down, right_up, up, down.
 *  pix_lab  [B,H,W] float, optional: 1 where ids != 0.
 */
PLH_API int plh_link_labels(const uint8_t* ids, int B, int H, int W, float* link_lab, float* pix_lab, void* stream);

/*
 * Measurement hooks (bench.py roofline): between begin and end, every plh_pixellink_loss call
 * records CUDA events around its dominant kernel (loss_main) on the launching stream;
 * end() returns the summed device time and the number of launches.  Not for use under
 * CUDA-graph capture.  This is the only state the library keeps, and only while enabled.
 */
PLH_API int plh_profile_begin(int max_launches);
PLH_API int plh_profile_end(float* total_ms, int* n_launches);
/* device-side view of the same launches: sum of (last CTA end - first CTA start), %globaltimer; call it
   before plh_profile_end.  Event time minus this is launch/drain latency, not kernel work. */
PLH_API int plh_profile_kernel_window(float* total_ms);
/* number of kernels the library has launched in this process (for bench.py's gpu_launches) */
PLH_API long long plh_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* PLHEAD_H_ */
