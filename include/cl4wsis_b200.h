/*
 * cl4wsis_b200.h — C ABI of the B200-native (sm_100a) pseudo-label hot path of CL4WSIS.
 *
 * This is the drop-in boundary.  The reference has no native code and therefore
 * no FFI of its own ("zero native code"); each entry point below
 * replaces the body of one reference Python function, cited per function
 * (paths relative to the reference checkout).  INTEGRATION.md shows the ctypes
 * binding a reference maintainer would add.
 *
 * Conventions:
 *   - plain pointers and sizes only; every tensor is contiguous NCHW fp32 unless
 *     stated otherwise; all data pointers are DEVICE pointers on the current
 *     CUDA device; the caller owns every buffer, including scratch;
 *   - every call enqueues work on `stream` (a cudaStream_t passed as void*) and
 *     returns immediately; nothing here synchronises or allocates;
 *   - return value: CL4_OK (0) or a negative error code; never throws;
 *     cl4_last_error() gives a thread-local message for the last failure.
 */
#ifndef CL4WSIS_B200_H_
#define CL4WSIS_B200_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CL4_OK 0
#define CL4_EINVAL (-1)       /* bad shape / null pointer / bad parameter            */
#define CL4_EUNSUPPORTED (-2) /* valid in the reference but outside this build's range */
#define CL4_ECUDA (-3)        /* a CUDA runtime call or launch failed                  */
#define CL4_ESCRATCH (-4)     /* scratch buffer too small                              */

#define CL4_MAX_DILATIONS 8

typedef void* cl4_stream_t; /* cudaStream_t */
typedef void* cl4_event_t;  /* cudaEvent_t  */

/* ABI version (bumped on any signature change) and last error text. */
int cl4_abi_version(void);
const char* cl4_last_error(void);

/* ------------------------------------------------------------------------- *
 * PAMR — wss/modules.py:122-152 (class PAMR), stencils :17-119.
 * ------------------------------------------------------------------------- */

/* Stand-alone forwards of the helper stencils PAMR is built from (the fused PAMR entry
 * points below never materialise these tensors).  x [planes,H,W] (planes = B*K):
 *   cl4_local_affinity, mode 0: LocalAffinity.forward      x - shift_p(x)    wss/modules.py:47-62
 *                       mode 1: LocalAffinityAbs.forward   |x - shift_p(x)|  wss/modules.py:115-119
 *                       mode 2: LocalAffinityCopy.forward  shift_p(x)        wss/modules.py:65-83
 *     -> out [planes,8*D,H,W], p = dilation_index*8 + tap (tap order :30-40), replicate padding :57;
 *   cl4_local_stdev: LocalStDev.forward, unbiased std over the 9*D samples   wss/modules.py:86-112
 *     -> out [planes,1,H,W]. */
int cl4_local_affinity(const float* x, float* out, int planes, int H, int W, const int* dilations, int D,
                       int mode, cl4_stream_t stream);
int cl4_local_stdev(const float* x, float* out, int planes, int H, int W, const int* dilations, int D,
                    cl4_stream_t stream);

/* smoothing(heat, kernel=3) — wss/utils.py:28-32 (applied to the CAM before peak_extract,
 * train.py:429): avg_pool2d(kernel, stride 1, zero padding (kernel-1)//2, padded cells counted).
 * x [planes,H,W] -> out [planes,H,W]; kernel odd; out must not alias x. */
int cl4_smoothing(const float* x, float* out, int planes, int H, int W, int kernel, cl4_stream_t stream);

/* F.interpolate(mask, size=(H,W), mode="bilinear", align_corners=True), the first
 * line of PAMR.forward (wss/modules.py:134).  in [planes,h,w] -> out [planes,H,W]. */
int cl4_resize_bilinear_ac(const float* in, float* out, int planes, int h, int w, int H, int W,
                           cl4_stream_t stream);

/* Affinity weights, wss/modules.py:141-145:
 *   w = softmax_p( mean_k( -|x_k - shift_p x_k| / (1e-8 + 0.1 * std_k) ) )
 * img [B,K,H,W] -> w [B,8*D,H,W] (p = dilation_index*8 + tap, taps in the
 * LocalAffinity order wss/modules.py:30-40; replicate padding :57). */
int cl4_pamr_weights(const float* img, float* w, int B, int K, int H, int W, const int* dilations, int D,
                     cl4_stream_t stream);

/* One propagation sweep, wss/modules.py:148-149:
 *   mask_out[b,c] = sum_p w[b,p] * shift_p(mask_in[b,c])
 * mask_in must not alias mask_out. */
int cl4_pamr_sweep(const float* w, const float* mask_in, float* mask_out, int B, int C, int H, int W,
                   const int* dilations, int D, cl4_stream_t stream);

/* Whole PAMR.forward after the resize: weights + num_iter sweeps.
 * img [B,K,H,W], mask_in [B,C,H,W] -> mask_out [B,C,H,W].  scratch must hold
 * cl4_pamr_scratch_bytes(...) bytes (weights [B,8D,H,W] + up to two replicate-padded
 * mask buffers [B*C,H+48,W+48]).  mask_in is not modified; mask_out may not alias it.
 * img, mask_in, mask_out and scratch must be 16-byte aligned (128-bit loads, TMA), else CL4_EINVAL.
 * The _timed variant additionally records the two (nullable) events on `stream`
 * around the num_iter propagation sweeps, for per-kernel timing. */
size_t cl4_pamr_scratch_bytes(int B, int K, int C, int H, int W, int D, int num_iter);
int cl4_pamr_forward(const float* img, const float* mask_in, float* mask_out, void* scratch,
                     size_t scratch_bytes, int B, int K, int C, int H, int W, const int* dilations, int D,
                     int num_iter, cl4_stream_t stream);
int cl4_pamr_forward_timed(const float* img, const float* mask_in, float* mask_out, void* scratch,
                           size_t scratch_bytes, int B, int K, int C, int H, int W, const int* dilations, int D,
                           int num_iter, cl4_stream_t stream, cl4_event_t ev_sweeps_begin,
                           cl4_event_t ev_sweeps_end);

/* ------------------------------------------------------------------------- *
 * peak_extract — wss/utils.py:3-25.
 *   keep = (max_pool2d(heat,k,1,(k-1)//2) == heat); peak = heat*keep;
 *   top-K of peak per (b,c) plane, sorted by (score desc, flat index asc);
 *   ys = int(float(idx)/W), xs = idx % W.
 * heat [B,C,H,W] -> scores f32 [B,C,K], ys i32 [B,C,K], xs i32 [B,C,K] (device).
 * kernel must be odd (an even kernel raises in the reference: shape mismatch at
 * wss/utils.py:11); 1 <= K <= H*W (K > 256 is selected in rounds of 256).
 * ------------------------------------------------------------------------- */
size_t cl4_peak_extract_scratch_bytes(int B, int C, int H, int W, int kernel, int K);
int cl4_peak_extract(const float* heat, float* scores, int* ys, int* xs, void* scratch, size_t scratch_bytes,
                     int B, int C, int H, int W, int kernel, int K, cl4_stream_t stream);
/* The phase-2 chain in front of peak_extract (train.py:426-436):
 *   _, cam = peakgenerator(int_masks, l1h)   -> PeakGenerator.cam_normalize, wss/modules.py:425-434
 *   cam = smoothing(cam)                     -> cl4_smoothing (wss/utils.py:28-32)
 *   cam = F.interpolate(cam, images.shape[-2:], mode="bilinear", align_corners=False)
 *   peak_extract(cam, kernel=15)
 * cl4_cam_normalize: cam [B,C,h,w], label [B,C] -> out [B,C,hs,ws] = relu(cam) * label, bilinearly resized
 *   (align_corners=False; hs == h and ws == w, the trainer's call, is the identity), divided by
 *   (plane maximum + 1e-5).
 * cl4_peak_extract_upsampled: peak_extract of F.interpolate(small [B,C,h,w], (H,W), bilinear,
 *   align_corners=False) WITHOUT materialising the [B,C,H,W] map: the tile loader evaluates ATen's bilinear
 *   formula on the fly from the small map.  Same outputs, scratch and limits as cl4_peak_extract(…, H, W, …). */
int cl4_cam_normalize(const float* cam, const float* label, float* out, int B, int C, int h, int w, int hs, int ws,
                      cl4_stream_t stream);
int cl4_peak_extract_upsampled(const float* small, int h, int w, float* scores, int* ys, int* xs, void* scratch,
                               size_t scratch_bytes, int B, int C, int H, int W, int kernel, int K, cl4_stream_t stream);

/* ------------------------------------------------------------------------- *
 * find_instance_center — modules/utils.py:463-502 (twin: dataset/utils.py:623-661),
 * batched over N independent [H,W] heat-maps.
 *   t = x > thr ? x : -1; p = max_pool2d(t,k,1,(k-1)//2); centre iff t == p and t > min_value
 *   (min_value = 0 reproduces :492; the reference's degenerate top_k branch
 *   :500-502 re-runs the same test with another min_value).
 * ctr_out [N,max_out,2] int64 (y,x) in torch.nonzero (row-major) order;
 * count_out [N] int32 = TOTAL number of centres per map (may exceed max_out; only
 * the first max_out are written).  heat is not modified.
 * ------------------------------------------------------------------------- */
size_t cl4_center_nms_scratch_bytes(int N, int H, int W);
int cl4_center_nms(const float* heat, float threshold, float min_value, int kernel, int N, int H, int W,
                   long long* ctr_out, int* count_out, int max_out, void* scratch, size_t scratch_bytes,
                   cl4_stream_t stream);

/* ------------------------------------------------------------------------- *
 * group_pixels — modules/utils.py:505-542, with the `(fg * ins_seg).long()` of
 * get_instance_segmentation (:606) folded in when fg != NULL; batched over N.
 *   id = 1 + argmin_k sqrt_rn(fma_rn(dx,dx,rn(dy*dy))), first minimum wins,
 *   dy = float(ctr_y) - (float(y) + off_y), dx likewise.
 * ctr [N,ctr_stride,2] int64 (y,x); the number of centres of map n is
 * count_dev[n] (device int32, clamped to ctr_stride) when count_dev != NULL,
 * else Kc for every map.  offsets [N,2,H,W]; fg [N,H,W] uint8 (0/1) or NULL;
 * ids [N,H,W] int64.  A map with zero centres gets all-zero ids when
 * empty_mode == 0 (ignore=True, :597-598) or fg (as 0/1) when empty_mode == 1 (:599-600).
 * ------------------------------------------------------------------------- */
int cl4_group_pixels(const long long* ctr, const int* count_dev, int Kc, int ctr_stride, const float* offsets,
                     const unsigned char* fg, long long* ids, int N, int H, int W, int empty_mode,
                     cl4_stream_t stream);

/* ------------------------------------------------------------------------- *
 * Connected components for cluster_peaks — modules/utils.py:608-632 (the centre
 * clustering of get_instance_segmentation, :567-594).
 *   weak = (sqrt(off_x^2 + off_y^2) < thresh) & fg
 *   cv2.connectedComponentsWithStats(weak, connectivity=4); keep area_lo < area < area_hi
 * offsets [2,H,W] fp32 (dy,dx), fg [H,W] uint8.  roots_out [max_out,2] int64: first
 * pixel (y,x) of every kept component in OpenCV's label order (raster order of the
 * first pixel); stats_out [max_out+1,3] int64: row 0 = OpenCV's label 0 (all other
 * pixels): area, sum x, sum y; rows 1.. = the kept components; count_out [1] int32 =
 * number of kept components (may exceed max_out).  centroid = sum / area in double,
 * exactly OpenCV's.
 * ------------------------------------------------------------------------- */
size_t cl4_ccl4_scratch_bytes(int H, int W);
int cl4_ccl4_components(const float* offsets, const unsigned char* fg, float thresh, float area_lo, float area_hi,
                        int H, int W, long long* roots_out, long long* stats_out, int* count_out, int max_out,
                        void* scratch, size_t scratch_bytes, cl4_stream_t stream);

/* ------------------------------------------------------------------------- *
 * refine_label_generation — modules/utils.py:257-385, the phase-2 caller of the
 * post-processing (train.py:492-500), for a whole batch with no host round trip.
 *
 * cl4_contours8: the cv2.connectedComponentsWithStats(seg == cls+1, connectivity=8)
 * of :305-307 for every (image, valid class) at once.  gt_seg [B,H,W] int64 labels
 * (0 = background, cls+1 otherwise), label [B,C] (non-zero = class present, :299).
 * comp_out [B,H,W] int32: contour slot of each pixel, -1 for background, invalid
 * classes and contours smaller than min_area (:313); info_out
 * [B,cl4_refine_max_contours(),5] int32 per slot: first pixel index, class,
 * int(centroid x), int(centroid y), area; ncomp_out [B]; status_out [1] (bit 0:
 * more contours than slots).  Slots are numbered in no particular order (the
 * reference's results do not depend on OpenCV's label order).
 *
 * cl4_refine_labels: seg_logits [B,C+1,H,W], center [B,C,H,W], offsets [B,2,H,W]
 * -> out_center [B,C,H,W], out_offset [B,2,H,W], out_weight [B,1,H,W] (all fp32).
 * gauss: the (6*sigma+3)^2 bump of modules/utils.py:49-59 as fp32; refine_thresh,
 * nms_kernel, beta, sigma = args.refine_thresh / kernel / beta / sigma; min_area =
 * MINIMUM_MASK_SIZE (20), max_inst = MAXIMUM_NUM_INST (5); top_k < 0 = None.
 * status_out [1] device int32: 0 = done; any bit set = a capacity limit was hit
 * (1: > 1024 contours per image, 2: > 64 centres in a contour, 4: > 4096 centres or
 * cluster components per image, 8: the reference's degenerate top_k branch would
 * run) and the outputs must be recomputed contour by contour.
 * ------------------------------------------------------------------------- */
int cl4_refine_max_contours(void);
/* pseudo_label_generation — modules/utils.py:179-253 as driven by train.py:451-466, for a batch:
 * the peaks of wss.utils.peak_extract (peak_conf f32, peak_y/peak_x i32, all [B,C,K], score-sorted)
 * with conf >= pseudo_thresh are matched to the 8-connected contours of seg_gt [B,H,W]; a contour of
 * >= min_area pixels holding exactly ONE peak of its class becomes an instance: gaussian max-splat at
 * the contour's centroid into out_center [B,C,H,W], weight 1 and centroid offsets on its pixels
 * (out_weight [B,1,H,W], out_offset [B,2,H,W]); total_match [B] int32 counts the accepted contours.
 * cls_label [B,C]: non-zero = class considered.  Scratch: cl4_refine_scratch_bytes(B,H,W). */
int cl4_pseudo_labels(const long long* seg_gt, const float* cls_label, const float* peak_conf,
                      const int* peak_y, const int* peak_x, int K, float pseudo_thresh, const float* gauss,
                      int sigma, int min_area, float* out_center, float* out_offset, float* out_weight,
                      int* total_match, int* status_out, int B, int C, int H, int W, void* scratch,
                      size_t scratch_bytes, cl4_stream_t stream);
size_t cl4_refine_scratch_bytes(int B, int H, int W);
int cl4_contours8(const long long* gt_seg, const float* label, int min_area, int B, int C, int H, int W,
                  int* comp_out, int* info_out, int* ncomp_out, int* status_out, void* scratch,
                  size_t scratch_bytes, cl4_stream_t stream);
int cl4_refine_labels(const float* seg_logits, const float* center, const float* offsets, const float* label,
                      const long long* gt_seg, const float* gauss, int sigma, double refine_thresh,
                      int nms_kernel, float beta, int min_area, int max_inst, long long top_k,
                      float* out_center, float* out_offset, float* out_weight, int* status_out, int B, int C,
                      int H, int W, void* scratch, size_t scratch_bytes, cl4_stream_t stream);

/* refine_label_generation_with_point — modules/utils.py:388-460 (point supervision; not reached by train.py), for a batch
 * in one launch.  gt_seg [B,H,W] int64 labels; label [B,C]; points [B,C,M,2] int64 (y,x) with keep [B,C,M] bytes: non-zero =
 * the point survives the reference's filter `gt_y != 0 and gt_x != 0` (:436, applied to the caller's values before the
 * int32 conversion of :437); offsets [B,2,H,W].  Every pixel whose gt class is valid (:431) and has a kept point gets
 * out_weight [B,1,H,W] = 1 and out_offset [B,2,H,W] = nearest kept point - pixel, nearest by group_pixels' arithmetic with
 * the first minimum winning (:444); all other pixels 0. */
int cl4_refine_labels_with_point(const long long* gt_seg, const float* label, const long long* points,
                                 const unsigned char* keep, const float* offsets, float* out_offset, float* out_weight,
                                 int B, int C, int M, int H, int W, cl4_stream_t stream);

/* ------------------------------------------------------------------------- *
 * get_ins_map — dataset/utils.py:795-902, the validation post-processing (Trainer.validate,
 * train.py:622), for one image with no host round trip per class / contour / instance.
 *
 * cl4_ins_map: seg_logits [Bf,C+1,H,W], center [Bf,C,H,W] with Bf = 2 when `flip` (args.val_flip: the
 * second view is mirrored and averaged in, :823-825) else 1; offset0 [2,H,W] = out['offset'][0], rescaled
 * IN PLACE by scale_y = target_h / H and scale_x = target_w / W as the reference does (:831-832);
 * cls_label [C] or NULL (args.val_clean, :835).  val_thresh / val_kernel / beta / ignore =
 * args.val_thresh / val_kernel / beta / val_ignore; min_area = MINIMUM_MASK_SIZE of dataset/utils.py:147 (50).
 * Outputs (device): seg_map_out [H,W] int64 (the argmax, :837); inst_map_out [H,W] int32 = index of the
 * pixel's instance in the output lists or -1; label_out / score_out [cl4_ins_map_max_instances()] =
 * pred_label / pred_score in the reference's order (class ascending, contours in OpenCV's label order,
 * instance id ascending, empty ids skipped); n_out [1]; status_out [1]: 0 = done, otherwise a capacity
 * limit was hit (1: > 1024 contours of >= min_area pixels; 2: > 64 accepted CLUSTER centres in one contour;
 * 4: > 4096 NMS centres / cluster components / instances in the image).  NMS centres per contour are not limited.
 * cl4_ins_masks: inst_map -> pred_mask [n,H,W] bytes (0/1).
 * ------------------------------------------------------------------------- */
int cl4_ins_map_max_instances(void);
size_t cl4_ins_map_scratch_bytes(int C, int H, int W);
int cl4_ins_map(const float* seg_logits, const float* center, float* offset0, const float* cls_label, int flip,
                float scale_y, float scale_x, float val_thresh, int val_kernel, float beta, int ignore, int min_area,
                long long* seg_map_out, int* inst_map_out, int* label_out, double* score_out, int* n_out,
                int* status_out, int C, int H, int W, void* scratch, size_t scratch_bytes, cl4_stream_t stream);
int cl4_ins_masks(const int* inst_map, int n, int H, int W, unsigned char* masks_out, cl4_stream_t stream);

/* ------------------------------------------------------------------------- *
 * Producers and consumers of PAMR inside the phase-1 step (train.py:372-385), SURVEY 8f rank 2.
 *
 * cl4_denorm: utils/utils.py:26-41 -- out = x * std[k] + mean[k] per RGB channel (two roundings, as
 *   Tensor.mul_().add_()); images [planes = B*3][HW].  mean, std: HOST arrays of three floats.
 * cl4_denorm_resize_ac: denorm followed by F.interpolate(..., (h, w), "bilinear", align_corners=True)
 *   (train.py:376-378) in one pass; images [B,K,Hi,Wi] -> out [B,K,h,w].  mean == std == NULL: resize only.
 * cl4_softmax_channels: int_masks.softmax(dim=1) (train.py:372-373); x, out [B,C,HW].
 * cl4_pseudo_gtmask: `int_masks_soft[:, 1:] *= l1h[:, :, None, None]` (train.py:382; labels [B,C-1], NULL = no
 *   gating; the gated mask goes to gated_out [B,C,HW], which may alias mask) followed by
 *   wss/single_stage.py:18-40: per (b, c) plane maximum, scaled by cutoff_bkg (c = 0) / cutoff_top, floored at
 *   cutoff_low; pseudo_out = (mask > threshold) as 0/1 floats, pixels claimed by more than one class cleared
 *   when `ambiguous`.  thr_scratch: B*C floats.
 * ------------------------------------------------------------------------- */
/* The whole of train.py:372-385 for feature-resolution maps (h, w <= 64; 1..6 dilations, each <= 24; num_iter >= 1) in TWO
 * launches: (1) per (image, 32x32 tile, 8-row slab): denorm + align-corners shrink of the image into a shared-memory window,
 * affinity weights from it; class softmax in CTAs of its own; (2) all num_iter PAMR sweeps on-chip followed by label gating,
 * plane maxima, thresholds and pseudo_gtmask(ambiguous=True) -- every CTA writes the labels of its planes and counts its claims
 * per pixel, the last CTA of an image to finish clears the pixels claimed more than once.
 * images [B,3,Hi,Wi], int_masks [B,C,h,w] logits, l1h [B,C-1] or NULL, mean / std: HOST arrays of three floats (NULL: no
 * denorm) -> soft_out [B,C,h,w] (int_masks_soft after gating), pseudo_out [B,C,h,w] (0/1).  CL4_EUNSUPPORTED outside the
 * stated range: use the separate entry points above. */
size_t cl4_phase1_scratch_bytes(int B, int C, int h, int w, int D);
int cl4_phase1_pseudo_labels(const float* images, const float* int_masks, const float* l1h, const float* mean,
                             const float* std, const int* dilations, int D, int num_iter, float cutoff_top,
                             float cutoff_bkg, float cutoff_low, float* soft_out, float* pseudo_out, void* scratch,
                             size_t scratch_bytes, int B, int C, int Hi, int Wi, int h, int w, cl4_stream_t stream);
int cl4_denorm(const float* images, float* out, int planes, int K, long long HW, const float* mean, const float* std,
               cl4_stream_t stream);
int cl4_denorm_resize_ac(const float* images, float* out, int B, int K, int Hi, int Wi, int h, int w, const float* mean,
                         const float* std, cl4_stream_t stream);
int cl4_softmax_channels(const float* x, float* out, int B, int C, long long HW, cl4_stream_t stream);
int cl4_pseudo_gtmask(const float* mask, const float* labels, float* gated_out, float* pseudo_out, float* thr_scratch,
                      int B, int C, int HW, float cutoff_top, float cutoff_bkg, float cutoff_low, int ambiguous,
                      cl4_stream_t stream);

/* Layout introspection of the default propagation kernel (host only): the thread (0..127 of warp group `group`, 0 = dilations
 * {4,8,12}, 1 = {1,2,24}) and pixel slot (0..7) that hold the affinity weights of pixel (y, x) of a 32 x 32 tile; the weight
 * of tap t (0..23 within the group) sits at float index ((slot*24 + t)/4 * 256 + group*128 + thread)*4 + t%4 of the tile's
 * 49152-float block in the scratch buffer.  No reference counterpart (wss/modules.py keeps weights as a [B,1,8D,H,W] tensor). */
int cl4_lattice_owner(int group, int y, int x, int* thread_out, int* slot_out);

#ifdef __cplusplus
}
#endif
#endif /* CL4WSIS_B200_H_ */
