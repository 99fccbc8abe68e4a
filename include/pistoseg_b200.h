/*
 * pistoseg_b200.h -- C ABI of libpistoseg_b200.so: the sm_100a (B200) implementation of PistoSeg's
 * dense-prediction post-processing hot path and of the mosaic dataset-synthesis gather.
 *
 * The reference (Vison307/PistoSeg) is pure Python and has no FFI; the boundary it offers is its Python
 * surface.  Every entry point below names the reference code it replaces
 * (file:line in the reference tree); pistoseg_b200/*.py binds these symbols with ctypes and mirrors the
 * reference's function / class names on top of them.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller unless the name ends in _host;
 *   - the library never frees or retains caller memory beyond the call;
 *   - every device entry point is asynchronous on the given stream (pisto_stream_t == cudaStream_t);
 *   - return value: PISTO_OK (0) or an error code; pisto_last_error() gives a thread-local message;
 *   - no global state: scratch / constant tables hang off an explicit per-device handle.
 */
#ifndef PISTOSEG_B200_H
#define PISTOSEG_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PISTO_ABI_VERSION 1

typedef struct pisto_ctx* pisto_handle_t;
typedef void* pisto_stream_t; /* cudaStream_t */

enum {
  PISTO_OK = 0,
  PISTO_ERR_INVALID = 1,     /* bad argument (message says which) */
  PISTO_ERR_CUDA = 2,        /* a CUDA runtime call failed */
  PISTO_ERR_UNSUPPORTED = 3, /* shape outside the compiled range */
  PISTO_ERR_NO_DEVICE = 4    /* no sm_100 device: there is NO CPU fallback */
};

/* how the V views are merged */
enum {
  PISTO_FUSE_LOGIT_MEAN = 0, /* (sum_v up_v) / V          ttach Merger('mean'), infer_pseudo_masks.py:96 */
  PISTO_FUSE_PROB_MEAN = 1   /* (sum_v softmax_c(up_v)) / V   segmentation_test.py:150,173 */
};

/* what is done with the per-tile class-presence vector `present[N][C]` */
enum {
  PISTO_MASK_NONE = 0,    /* present ignored                                         loss.py:55-60 */
  PISTO_MASK_FILL = 1,    /* absent -> -1e10; exactly one class present -> that class everywhere
                             (no scores read)                                        infer_pseudo_masks.py:69-83 */
  PISTO_MASK_NEG_INF = 2, /* absent -> -inf                       OEEM/classification/utils/generate_CAM.py:91-97 */
  PISTO_MASK_MULTIPLY = 3 /* score * present (absent -> exactly 0, may win)          infer_revise_masks.py:137-139 */
};

/* how the label is decided from the (masked) fused scores */
enum {
  PISTO_DECIDE_SOFTMAX = 0, /* argmax_c softmax_c(x), lowest index on ties   infer_pseudo_masks.py:80-82, loss.py:59-60 */
  PISTO_DECIDE_RAW = 1      /* argmax_c x                                    loss.py:57, segmentation_test.py:209 */
};

/* One test-time-augmentation view: the logits the backbone produced for the augmented input. */
typedef struct {
  const float* logits; /* [N][C][h][w], tile n starts at logits + n * tile_stride */
  int64_t tile_stride; /* elements between consecutive tiles; 0 means dense (C*h*w) */
  int32_t h, w;        /* resolution of this view (stride-8 multi-scale: 21/28/35 for a 224 tile) */
  int32_t xform;       /* de-augmentation k + 4*hflip: deaug(y) = flip_w_if_hflip(rot90(y, k)) (ttach d4) */
  int32_t reserved;
} pisto_view_t;

/* -------------------------------------------------------------------------------------------------- */
/* lifetime                                                                                           */
/* -------------------------------------------------------------------------------------------------- */
int pisto_abi_version(void);
const char* pisto_last_error(void);
/* Fails with PISTO_ERR_NO_DEVICE when `device` is not an sm_100 GPU. */
int pisto_create(pisto_handle_t* out, int device);
int pisto_destroy(pisto_handle_t h);
/* number of kernel launches issued through this handle so far (bench.py's gpu_launches) */
int64_t pisto_launch_count(pisto_handle_t h);
/* Data-dependence record of the filtered fusion kernels (the label fast path trusts a pixel only above an error-bound margin and
 * re-evaluates the others exactly -- in whole 4-pixel groups, by the shape-specialised kernel's fixer warp): out_host[1] = multi-label tiles processed, out_host[2] = pixels that went through the exact
 * pass, out_host[3] = tiles evaluated exactly as a whole (queue overflow, non-finite / absurd logits, empty presence vector);
 * out_host[0] = mosaic cells still rejected by the "background < 80 %" test at the last of max_tries draws (pisto_mosaic_plan_cells
 * accepts them -- the reference would loop for ever -- but never silently).  Synchronises the device; reset != 0 zeroes the counters afterwards. */
int pisto_filter_stats(pisto_handle_t h, unsigned long long* out_host /* [4] */, int reset);

/* -------------------------------------------------------------------------------------------------- */
/* confusion matrix: replaces mIoUMask._generate_matrix / add_batch (loss.py:17-31)                   */
/*   conf[gt*C + pred] += 1 for every pixel with gt < C (rows = ground truth, cols = prediction).     */
/*   pred >= C at a counted pixel is an error in the reference (bincount overflow); here such pixels  */
/*   are counted in *bad_pred (device, may be NULL) and skipped.                                      */
/* -------------------------------------------------------------------------------------------------- */
int pisto_confusion_accumulate(pisto_handle_t h, const uint8_t* pred, const uint8_t* gt, int64_t n_px, int C,
                               unsigned long long* conf /* [C*C], accumulated */,
                               unsigned long long* bad_pred /* [1] or NULL */, pisto_stream_t stream);

/* -------------------------------------------------------------------------------------------------- */
/* the fused hot path: replaces                                                                       */
/*   tta.SegmentationTTAWrapper(...,'mean') merge            infer_pseudo_masks.py:96,121            */
/*   interpolate_tensor (bilinear, align_corners=False)       infer_pseudo_masks.py:89-90,126         */
/*   get_mask_pred_and_entropy                                infer_pseudo_masks.py:69-87             */
/*   mIoUMask.forward (softmax, argmax, confusion)            loss.py:55-67                           */
/*   revise-mask argmax + background                          infer_revise_masks.py:137-155           */
/* For each tile n and output pixel (y, x):                                                           */
/*   s_c   = fuse_v( bilinear(deaug(view_v))[c][y][x] )                 (fp32, view order, one divide)  */
/*   x_c   = mask(s_c, present[n][c])                                                                 */
/*   lab   = decide(x)                          (lowest index on ties)                                */
/*   conf[gt*C + lab] += 1  if gt != NULL and gt < C       (BEFORE the background overwrite)          */
/*   label_out = (bg != NULL and bg[n][y][x] == bg_match) ? bg_label : lab                            */
/* Optional outputs: fused_out = s (unmasked), entropy_out = -sum p*log(p+1e-10) of the masked        */
/* softmax, lowres_out = bilinear(s -> [low_h][low_w]) (the 32x32 logits of infer_pseudo_masks.py:126;*/
/* computed in-kernel when T/low is an odd integer (224 -> 32 is the gather [3::7]); otherwise        */
/* fused_out must be given and the library runs pisto_upsample_bilinear on it).                       */
/* -------------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t N, C, T_h, T_w;
  int32_t fuse_mode;   /* PISTO_FUSE_*   */
  int32_t mask_mode;   /* PISTO_MASK_*   */
  int32_t decide_mode; /* PISTO_DECIDE_* */
  int32_t bg_match;    /* pixel is background iff bg[..] == bg_match (tissue==0: 0; gt==3: 3) */
  int32_t bg_label;    /* label written on background pixels (len(patch_label) / 3) */
  int32_t low_h, low_w;
  int32_t impl;        /* 0 = auto, 1 = generic one-thread-per-pixel kernel, 2 = exact streaming kernel, 3 / 4 = filtered streaming kernel with 2 / 4 columns per thread, 5 = block-tiled filtered kernel (large tiles), 6 = shape-specialised filtered kernel (224x224 tiles, 21/28/35 px views), 7 = the same with two CTAs per SM, 8 = the same with 2 columns per thread (26 warps per SM) */
  const uint8_t* present; /* [N][C] 0/1 or NULL */
  const uint8_t* bg;      /* [N][T_h][T_w] or NULL */
  const uint8_t* gt;      /* [N][T_h][T_w] or NULL */
  uint8_t* label_out;     /* [N][T_h][T_w] or NULL */
  float* fused_out;       /* [N][C][T_h][T_w] or NULL */
  float* entropy_out;     /* [N][T_h][T_w] or NULL */
  float* lowres_out;      /* [N][C][low_h][low_w] or NULL */
  unsigned long long* conf; /* [C*C] accumulated, or NULL */
  uint8_t* label_raw_out;   /* [N][T_h][T_w] or NULL: the labels of the same scores decided with PISTO_DECIDE_RAW (argmax of the fused logits),
                             * written beside label_out in the same pass, with the same background overwrite.  segmentation_test.py:137-139,182
                             * needs both for BCSS -- softmax-argmax for the confusion matrix, logit-argmax for the PNG -- and read the logits
                             * twice.  Served by the one-view full-resolution kernel and the generic kernel (impl 0 / 1). */
} pisto_fuse_args_t;

int pisto_fuse_argmax_confusion(pisto_handle_t h, const pisto_view_t* views_host, int V, const pisto_fuse_args_t* args_host,
                                pisto_stream_t stream);

/* Same call with HOST buffers (pinned or pageable): uploads the views / masks, runs the kernel, downloads the outputs,
 * in `chunk` tile chunks double-buffered over internal streams, and returns when everything is on the host.
 * This is the call the end-to-end ("e2e") benchmark times.  conf_host is int64[C*C], accumulated. */
int pisto_fuse_argmax_confusion_host(pisto_handle_t h, const pisto_view_t* views_host_ptrs, int V,
                                     const pisto_fuse_args_t* args_host_ptrs, int chunk);
/* device-clock duration (CUDA events, first H2D start -> last D2H end) of the last *_host call on this handle */
double pisto_last_pipeline_ms(pisto_handle_t h);

/* Self-test hook: counts the float32 inputs a (all 2^32 bit patterns) for which the library's a / V shortcut differs
 * from IEEE division; adds the count to *mismatches_dev (device).  Expected: 0. */
int pisto_selftest_div(pisto_handle_t h, int V, unsigned long long* mismatches_dev, pisto_stream_t stream);

/* -------------------------------------------------------------------------------------------------- */
/* revise-mask tail: PIL mode-'P' resize (= NEAREST, ImagingScaleAffine) to the original size, then   */
/* mask[background > 0] = bg_value AT THE ORIGINAL RESOLUTION (infer_revise_masks.py:152-155,164-174). */
/* in [n_tiles][sh][sw] u8 label maps (device); one descriptor per output image: which tile, its      */
/* original (h, w), offsets of its row / column source-index tables in index_pool (int32, built by    */
/* the caller exactly as PIL accumulates them), of its background mask in bg_pool (-1: none) and of    */
/* its output in out_pool.  All pools and the descriptors are device memory.                           */
/* -------------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t tile;            /* index into in[] */
  int32_t h, w;            /* original size */
  int32_t reserved;
  int64_t iy_off, ix_off;  /* int32 elements into index_pool: h row indices, w column indices */
  int64_t bg_off;          /* bytes into bg_pool ([h][w] u8), or -1 */
  int64_t out_off;         /* bytes into out_pool ([h][w] u8) */
} pisto_resize_desc_t;

int pisto_resize_nearest_bg(pisto_handle_t h, const uint8_t* in, int n_tiles, int sh, int sw, const pisto_resize_desc_t* desc, int n_desc,
                            const int32_t* index_pool, const uint8_t* bg_pool, uint8_t* out_pool, int bg_value, pisto_stream_t stream);

/* -------------------------------------------------------------------------------------------------- */
/* bilinear resize, align_corners=False: replaces interpolate_tensor / F.interpolate(mode='bilinear') */
/*   infer_pseudo_masks.py:89-90, segmentation_test.py:88-89,197, prepare_seg_inputs.py:116,131,137   */
/*   dtype: 0 = float32, 1 = float64.  in [NC][hi][wi] -> out [NC][ho][wo]                            */
/* -------------------------------------------------------------------------------------------------- */
int pisto_upsample_bilinear(pisto_handle_t h, const void* in, void* out, int64_t NC, int hi, int wi, int ho, int wo,
                            int dtype, pisto_stream_t stream);

/* -------------------------------------------------------------------------------------------------- */
/* overlap-add stitching of tile scores into a float64 canvas: replaces                               */
/*   segmentation_test.py:145-174 (softmax of the cropped tile, canvas[y:y+h, x:x+w, :] += p, cnt += 1)*/
/*   prepare_seg_inputs.py:120-128 (sum_cam[:, y:y+side, x:x+side] += crop; counter += 1)            */
/* tiles [n][C][th][tw] f32 (row pitch tw); tile k covers canvas rows y_k.. of height crop_h_k (clipped */
/* to the canvas).  canvas is [C][H][W] f64 (planar), count [H][W] f64.  softmax != 0 applies a channel */
/* softmax (fp32, as torch) to each pixel first.  Tiles may overlap: float64 atomics are NOT used --  */
/* each canvas pixel is owned by one thread that visits the covering tiles in index order, so the sum  */
/* order (and therefore the bits) equal the reference's sequential loop.                              */
/* -------------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t y, x;           /* top-left of the tile on the canvas */
  int32_t crop_h, crop_w; /* valid rows / cols of the tile (original_h / original_w) */
} pisto_tile_pos_t;

int pisto_stitch_accumulate(pisto_handle_t h, const float* tiles, const pisto_tile_pos_t* pos /* device [n] */, int n, int C,
                            int th, int tw, int softmax, double* canvas, double* count, int H, int W,
                            pisto_stream_t stream);

/* canvas[c][p] /= max(count[p], min_count) in place (segmentation_test.py:188,204; prepare_seg_inputs.py:128-130
 * uses min_count = 1; the reference test script divides by the raw count -> min_count = 0 gives nan on uncovered px) */
int pisto_canvas_normalize(pisto_handle_t h, double* canvas, const double* count, int C, int64_t HW, double min_count,
                           pisto_stream_t stream);

/* out[c][p] += in[c][p] * scale     (segmentation_test.py:198, prepare_seg_inputs.py:134-136) */
int pisto_canvas_axpy(pisto_handle_t h, double* out, const double* in, int64_t n, double scale, pisto_stream_t stream);

/* argmax over C of a planar float64 [C][HW] score map (np.argmax, lowest index on ties; present/-inf masking as
 * generate_CAM.py:91-99), confusion against gt (before bg), bg overwrite -- segmentation_test.py:207-211 */
int pisto_argmax_f64(pisto_handle_t h, const double* scores, int C, int64_t HW, const uint8_t* present /* [C] host or NULL */,
                     const uint8_t* gt, int bg_match, int bg_label, uint8_t* pred_out, uint8_t* label_out,
                     unsigned long long* conf, pisto_stream_t stream);

/* -------------------------------------------------------------------------------------------------- */
/* mosaic synthesis: replaces CropAndConcatDataset.__getitem__                                        */
/*   create_dataset.ipynb:273-372 [cell 9], create_dataset_bcss.ipynb:257-340 [cell 8]                */
/* Plan in -> pixels out, bit-exact with cv2.flip / cv2.warpAffine(INTER_LINEAR | INTER_NEAREST,      */
/* BORDER_REFLECT_101) / albumentations PadIfNeeded + RandomCrop given the same decisions.            */
/* -------------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t flip;   /* 0 none, 1 = cv2.flip code 0 (rows), 2 = code 1 (cols), 3 = code -1 (both) */
  int32_t warp;   /* 0 / 1: ShiftScaleRotate applied */
  int32_t crop_y, crop_x; /* RandomCrop origin on the patch_num*patch_size composite */
  double minv[6]; /* INVERSE affine map (dst -> src), float64, as cv::warpAffine derives it from M */
} pisto_mosaic_quad_t;

typedef struct {
  int32_t split_h, split_w; /* create_mosaic h, w (even) */
  int32_t reserved[2];
  pisto_mosaic_quad_t quad[4];
} pisto_mosaic_plan_t;

typedef struct {
  int32_t tile;   /* index into the tile pool */
  int16_t cy, cx; /* crop origin in PADDED tile coordinates (PadIfNeeded to >= patch_size, centred, REFLECT_101) */
} pisto_mosaic_cell_t;

int pisto_mosaic_gather(pisto_handle_t h, const uint8_t* pool_img /* HWC u8, tile t at 3*pool_off[t] */,
                        const uint8_t* pool_bg /* HW u8 at pool_off[t], >0 = background, or NULL (BCSS) */,
                        const int64_t* pool_off /* [P] pixel offsets */, const int32_t* pool_hw /* [P][2] */,
                        const uint8_t* pool_label /* [P] */, const pisto_mosaic_plan_t* plans /* [N] */,
                        const pisto_mosaic_cell_t* cells /* [N][4][patch_num^2] */, int N, int patch_num, int patch_size,
                        int bg_label, uint8_t* img_out /* [N][S][S][3] */, uint8_t* mask_out /* [N][S][S] */,
                        pisto_stream_t stream);

/* The same gather from a PACKED pool: one 32-bit word per source pixel, r | g << 8 | b << 16 | (bg > 0) << 24, built once per
 * pool by pisto_mosaic_pack_pool.  Same bytes per pixel as image + mask (4), but every bilinear tap is one aligned load. */
int pisto_mosaic_pack_pool(pisto_handle_t h, const uint8_t* pool_img, const uint8_t* pool_bg /* or NULL */, int64_t n_px,
                           uint32_t* pool_rgba, pisto_stream_t stream);
int pisto_mosaic_gather_packed(pisto_handle_t h, const uint32_t* pool_rgba, const int64_t* pool_off, const int32_t* pool_hw,
                               const uint8_t* pool_label, const pisto_mosaic_plan_t* plans, const pisto_mosaic_cell_t* cells, int N,
                               int patch_num, int patch_size, int bg_label, uint8_t* img_out, uint8_t* mask_out, pisto_stream_t stream);

/* Planning of the per-cell decisions on the device (create_dataset.ipynb:291-321: source tile, RandomCrop origin, the
 * "background area is smaller than 80 %" rejection loop).  Counter-based: cell (q, c) of mosaic i is a pure function of
 * (seed, i, q, c) through Philox4x32-10, identical for any GPU count; mosaic k of the call has index
 * first_index + k * index_stride (rank r of G: first_index = r, stride = G, as create_dataset.ipynb:554).
 * integral / integral_off (both NULL: no rejection, BCSS) is the 16-bit summed-area table of the padded background masks
 * built by pisto_mosaic_bg_integral: tile t occupies (ph+1)*(pw+1) entries at integral_off[t], ph = max(h, patch_size). */
int pisto_mosaic_bg_integral(pisto_handle_t h, const uint8_t* pool_bg, const int64_t* pool_off, const int32_t* pool_hw,
                             const int64_t* integral_off, int P, int patch_size, uint16_t* integral, pisto_stream_t stream);
int pisto_mosaic_plan_cells(pisto_handle_t h, uint64_t seed, int64_t first_index, int64_t index_stride, int N, int patch_num,
                            int patch_size, int P, const int32_t* pool_hw, const uint16_t* integral, const int64_t* integral_off,
                            int bg_label, int max_tries, pisto_mosaic_cell_t* cells /* [N][4][patch_num^2] */, pisto_stream_t stream);
/* The per-quadrant decisions of the same mosaics (create_dataset.ipynb:324-354: split, Flip, ShiftScaleRotate -> inverse affine in
 * float64, RandomCrop origin), also a pure function of (seed, i): Philox draws, explicit round-to-nearest float64 arithmetic and a
 * fixed cos / sin polynomial, bit-identical to pistoseg_b200/mosaic.py::MosaicPlanner.quad_plans on the host. */
int pisto_mosaic_plan_quads(pisto_handle_t h, uint64_t seed, int64_t first_index, int64_t index_stride, int N, int patch_num, int patch_size,
                            double p_flip, double p_warp, double shift_limit, double scale_limit, double rotate_limit,
                            pisto_mosaic_plan_t* plans, pisto_stream_t stream);

/* normalise + resize + sum over scales in one pass (segmentation_test.py:187-199, prepare_seg_inputs.py:128-134):
 *   out[c] (+)= bilinear_f64( canvas[c] / max(count, min_count if > 0) ) resized from (hi, wi) to (ho, wo);
 *   accumulate = 0 stores (first scale: `pred = m.copy()`), 1 adds (`pred + m`).  Operation order per value as in the reference. */
int pisto_canvas_resize_accumulate(pisto_handle_t h, const double* canvas /* [C][hi][wi] */, const double* count /* [hi][wi] */, int C,
                                   int hi, int wi, double min_count, double* out /* [C][ho][wo] */, int ho, int wo, int accumulate,
                                   pisto_stream_t stream);

/* -------------------------------------------------------------------------------------------------- */
/* background mask of RGB tiles: replaces utils.get_background (utils.py:155-163) and                 */
/* BaseDataset._get_background (dataset.py:100-109):                                                  */
/*   gray = cv2.cvtColor(RGB2GRAY) (8-bit fixed point); binary = gray > thresh (200);                 */
/*   4-connected components with fewer than min_size (50) pixels are cleared                          */
/*   (skimage.morphology.remove_small_objects(connectivity=1)); mask = 255 where kept, else 0.        */
/*   rgb [N][H][W][3] u8, mask_out [N][H][W] u8, scratch int32 [2*N*H*W] (labels + component sizes).   */
/* -------------------------------------------------------------------------------------------------- */
int pisto_get_background(pisto_handle_t h, const uint8_t* rgb, int N, int H, int W, int thresh, int min_size,
                         int32_t* scratch, uint8_t* mask_out, pisto_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* PISTOSEG_B200_H */
