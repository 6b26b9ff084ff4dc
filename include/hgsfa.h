/*
 * hgsfa.h -- C ABI of the B200-native HiGSFA sliding-window hot path (libhgsfa.so).
 *
 * This is the drop-in boundary for the data-parallel hot path of AlbertoEsc/PyFaceAnalysis
 * (SURVEY.md section 8b).  The reference is pure Python; every entry point below replaces the Python
 * call named beside it, and is what a ctypes (or cffi / CPython-extension) binding on the reference
 * side would bind -- see INTEGRATION.md for the stub.
 *
 * Conventions
 *   - plain C: pointers, sizes, ints.  No torch / numpy types cross this boundary.
 *   - every function returns 0 on success, non-zero on failure; hgsfa_last_error() then holds a
 *     message for the calling thread.  Dimension mismatches are errors, never truncation
 *     (MDP raises in _pre_execution_checks; SURVEY.md 8b "Errors").
 *   - handles are not thread-safe; one handle per (object, device) (SURVEY.md 8b "Threading").
 *   - "host" entry points take host pointers and do the H2D / D2H copies themselves;
 *     "_device" entry points take device pointers on the handle's device.
 *   - stream arguments are cudaStream_t passed as void*.  For "_device" entry points NULL is CUDA's
 *     default stream; for host entry points NULL selects a stream owned by the handle (the call
 *     returns only after its device-to-host copies have completed either way).
 */
#ifndef HGSFA_H_
#define HGSFA_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HGSFA_VERSION 1

/* element types of caller buffers */
enum { HGSFA_U8 = 0, HGSFA_F32 = 1, HGSFA_F64 = 2 };

/* layouts of a (n_windows x dim) matrix
 *   ROWMAJOR : element (w, f) at [w * ld + f]                      (what numpy hands over)
 *   TILED    : element (w, f) at [((w / 128) * dim + f) * 128 + w % 128]
 *              ("window-minor" tiles of 128 windows: the layout every flow kernel reads and writes,
 *              so that a warp touching one feature of 128 consecutive windows moves 512 contiguous bytes)
 */
enum { HGSFA_ROWMAJOR = 0, HGSFA_TILED = 1 };
#define HGSFA_TILE 128

/* resampling filters, numerically equal to Pillow's Image.NEAREST / Image.BILINEAR / Image.BICUBIC (the
 * alternatives the reference lists for interpolation_formats, FaceDetectUpdated.py:125) */
enum { HGSFA_NEAREST = 0, HGSFA_BILINEAR = 2, HGSFA_BICUBIC = 3 };

typedef struct hgsfa_plan_s*  hgsfa_plan_t;
typedef struct hgsfa_gauss_s* hgsfa_gauss_t;

const char* hgsfa_last_error(void);
int hgsfa_version(void);
int hgsfa_device_count(int* count);

/* ------------------------------------------------------------------------------------------------
 * Flow forward: replaces  sl = networks[i].execute(subimages_arr, benchmark=benchmark)
 *   reference call sites: FaceDetectUpdated.py:699, face_analysis.py:1064, face_analysis.py:1257
 *   (mdp.Flow.execute -> per-node execute of hinet.Switchboard / Layer / CloneLayer, SFANode,
 *   PCANode, WhiteningNode, cuicuilco GeneralExpansionNode and iGSFANode).
 *
 * A plan is built once per flow from a "plan blob": the flow's node graph lowered by
 * pyfaceanalysis_b200/plan.py into fused layer operations (receptive-field gather + mean
 * subtraction + expansion term table + projection), format documented in DESIGN.md section 4.
 * ---------------------------------------------------------------------------------------------- */
int hgsfa_plan_create(const void* blob, size_t nbytes, int device, hgsfa_plan_t* plan);
int hgsfa_plan_destroy(hgsfa_plan_t plan);
int hgsfa_plan_info(hgsfa_plan_t plan, int64_t* input_dim, int64_t* output_dim, int64_t* n_ops);

/* Work and traffic model of one forward pass over n windows:
 *   algorithmic_flops : sum over the reference's linear maps of 2*d_in*d_out, + d per mean
 *                       subtraction, + one per expansion term (as stored in the node graph)
 *   executed_flops    : what the fused kernels execute (folded maps, padded tiles)
 *   min_bytes         : input read once (in x_dtype) + y_cols outputs of 4 bytes per window */
int hgsfa_plan_flops(hgsfa_plan_t plan, int64_t n, int x_dtype,
                     double* algorithmic_flops, double* executed_flops, double* min_bytes);

/* x: host, row-major (n x input_dim, leading dimension ld elements), dtype u8 / f32 / f64.
 * y: host, row-major (n x y_cols), dtype f32 / f64; y_cols <= output_dim keeps the first y_cols
 *    features (the caller's sl[:, 0:D] slice).  Copies are chunked and overlapped with compute. */
int hgsfa_plan_execute(hgsfa_plan_t plan, const void* x, int x_dtype, int64_t n, int64_t ld,
                       void* y, int y_dtype, int64_t y_cols, void* stream);

/* Same with device buffers.  x_layout HGSFA_TILED is only valid for u8 / f32 and must hold
 * ceil(n/128)*128 windows (what hgsfa_crop_extent_device writes). y is row-major n x y_cols. */
int hgsfa_plan_execute_device(hgsfa_plan_t plan, const void* d_x, int x_dtype, int x_layout,
                              int64_t n, int64_t ld, void* d_y, int y_dtype, int64_t y_cols,
                              void* stream);

/* number of kernel launches issued by this plan since creation (bench.py's gpu_launches) and the
 * accumulated device time of the most recent execute (CUDA events on the plan's stream; ms) */
int hgsfa_plan_stats(hgsfa_plan_t plan, int64_t* launches, double* last_ms);

/* Per-operation device times (bench.py's roofline; not a reference interface).  hgsfa_plan_profile(plan, 1)
 * makes every following execute bracket each layer launch with CUDA events; hgsfa_plan_op_stats waits for
 * the work, adds the elapsed times to the per-op totals and returns them.  Arrays hold `capacity` entries
 * (one per op, see hgsfa_plan_info); engine: 0 = FFMA kernel, 1 = tensor-core kernel; flops are per window. */
int hgsfa_plan_profile(hgsfa_plan_t plan, int enable);
int hgsfa_plan_op_stats(hgsfa_plan_t plan, int64_t capacity, double* ms, int32_t* engine, double* alg_flops,
                        double* exe_flops);

/* tuning knobs (0 keeps the default): windows per front-segment chunk / back-segment chunk */
int hgsfa_plan_set_chunks(hgsfa_plan_t plan, int64_t front_chunk, int64_t back_chunk);

/* ------------------------------------------------------------------------------------------------
 * Window extraction: replaces  load_network_subimages(images, idx, coords, angles, w, h, interp, False)
 *   reference: face_analysis.py:775-800 -> cuicuilco extract_subimages_rotate + images_asarray ->
 *   Pillow Image.transform(out_size, EXTENT, box, filter)   (SURVEY.md row a-4).
 *
 * img    : uint8 (H x W), row-major, leading dimension W.
 * boxes  : n x 4 float64 (x0, y0, x1, y1); angles : n float64 current face angles in degrees or NULL
 *          (the patch is extracted with delta_ang = -angle like the reference).
 * out    : n x (ow*oh), row-major patches; out_dtype u8 / f32 / f64 (the reference returns f64 0..255).
 * angle == 0 windows follow Pillow's NEAREST index rule bit-exactly (sequential double accumulation).
 * ---------------------------------------------------------------------------------------------- */
int hgsfa_crop_extent(const uint8_t* img, int H, int W, const double* boxes, const double* angles,
                      int64_t n, int ow, int oh, int filter, void* out, int out_dtype, int device,
                      void* stream);

/* device version: d_img / d_boxes / d_angles / d_out are device pointers; out_layout ROWMAJOR or
 * TILED (u8 / f32 only; padded to a multiple of 128 windows, padding windows are zero). */
int hgsfa_crop_extent_device(const uint8_t* d_img, int H, int W, const double* d_boxes,
                             const double* d_angles, int64_t n, int ow, int oh, int filter,
                             void* d_out, int out_dtype, int out_layout, void* stream);

/* batch version: window w is extracted from image d_img_index[w]; d_img_ptrs[k] is the device address of
 * image k (uint8, row-major, leading dimension = its width) and d_img_hw[2k], d_img_hw[2k+1] its height and
 * width -- the windows of every image and scale of a detection batch in ONE launch
 * (the reference loops over windows, FaceDetectUpdated.py:686 -> face_analysis.py:781-786). */
int hgsfa_crop_extent_batch_device(const uint8_t* const* d_img_ptrs, const int32_t* d_img_hw,
                                   const int32_t* d_img_index, const double* d_boxes,
                                   const double* d_angles, int64_t n, int ow, int oh, int filter,
                                   void* d_out, int out_dtype, int out_layout, void* stream);

/* Per-patch contrast normalisation of TILED float32 patches, in place: the
 * contrast_enhance="AgeContrastEnhancement_Avg_Std", obj_avg, obj_std arguments of cuicuilco's
 * extract_subimages_rotate as the eye stage passes them (face_analysis.py:1042-1045):
 * v = x / 255;  y = (v - mean(v)) / (std(v) + 1e-8) * obj_std + obj_avg   (definition: oracle/crop.py). */
int hgsfa_contrast_avg_std_device(float* d_patches_tiled, int64_t n, int64_t dim, double obj_avg,
                                  double obj_std, void* stream);

/* Age-stage crop of n faces: replaces face_normalization_tools.normalize_image (face_normalization_tools.py:111-329,
 * called at face_analysis.py:1212) followed by the 96 x 96 sub-sampling of load_image_data_monoprocessor
 * (face_analysis.py:1230-1246) -- integer crop -> BICUBIC rotation -> BICUBIC EXTENT to 256 x 260 -> NEAREST
 * sub-sampling -- evaluated per output sample without materialising the intermediate images.
 * d_params: n x 16 doubles per face (pyfaceanalysis_b200/normalize.py: crop origin x, y, crop width, height, rotate
 * flag, rotation matrix a0..a5, EXTENT affine xs, x0, ys, y0); d_xtab / d_ytab: ow / oh NEAREST source indices into the
 * 256 x 260 image (-1 = outside).  Output: TILED float32 patches with values 0..255 (feed
 * hgsfa_contrast_avg_std_device, then the age flow). */
int hgsfa_age_crop_device(const uint8_t* const* d_img_ptrs, const int32_t* d_img_hw, const int32_t* d_img_index,
                          const double* d_params, int64_t n, const int32_t* d_xtab, const int32_t* d_ytab, int ow,
                          int oh, float* d_out_tiled, void* stream);

/* row-major <-> tiled conversion of a window matrix on the device (u8 or f32 elements) */
int hgsfa_tile_windows_device(const void* d_src, int dtype, int64_t n, int64_t dim, int64_t ld,
                              void* d_dst_tiled, int dst_dtype, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Gaussian classifier head: replaces
 *   reg_out = classifiers[i].regression(sl[:, 0:D], avg_labels [, estimate_std=True])
 *   classifiers[i].label(x) / class_probabilities(x)
 *   reference call sites: FaceDetectUpdated.py:709-719, face_analysis.py:1068-1071, 1261-1287
 *   (mdp.nodes.GaussianClassifier + cuicuilco's regression patch).
 *
 * means: C x D, inv_covs: C x D x D, sqrt_det: C (= _sqrt_def_covs), priors: C (= p), all float64.
 * Evaluated in float64 in the reference's operation order, so that the all-classes-underflow case
 * yields NaN exactly where the reference does.
 * ---------------------------------------------------------------------------------------------- */
int hgsfa_gauss_create(const double* means, const double* inv_covs, const double* sqrt_det,
                       const double* priors, int C, int D, int device, hgsfa_gauss_t* h);
int hgsfa_gauss_destroy(hgsfa_gauss_t h);

/* x: host n x D (leading dimension ld), dtype f32 / f64.  Any output pointer may be NULL.
 *   value : n float64   posterior-weighted mean of avg_labels  (regression)
 *   std   : n float64   sqrt(sum_c P_c (avg_label_c - value)^2) (estimate_std=True)
 *   winner: n int32     argmax_c P(c | x)                        (label)
 *   probs : n x C float64 class_probabilities                                           */
int hgsfa_gauss_regress(hgsfa_gauss_t h, const void* x, int x_dtype, int64_t n, int64_t ld,
                        const double* avg_labels, double* value, double* std, int32_t* winner,
                        double* probs, void* stream);
int hgsfa_gauss_regress_device(hgsfa_gauss_t h, const void* d_x, int x_dtype, int64_t n, int64_t ld,
                               const double* d_avg_labels, double* d_value, double* d_std,
                               int32_t* d_winner, double* d_probs, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Cascade controller on the device: replaces, for a batch of windows,
 *   update_current_subimage_coordinates(network_type, coords, angles, reg_out, ...)  face_analysis.py:803-840
 *   identify_patches_to_discard(network_type, ...)                                   face_analysis.py:842-887
 *   the order-preserving compaction block                                            FaceDetectUpdated.py:739-759
 *
 * type: 0 Disc, 1 PosX, 2 PosY, 3 PAng, 4 Scale.  coords (n x 4) and angles (n) are updated in place,
 * keep[i] = !new_wrong_images[i]; for Disc stages conf[i] = reg_out[i] (curr_confidence) when conf != NULL.
 * orig_coords / orig_angles / patch_wh ([.][2] = patch_width, patch_height of the window's scale) are
 * indexed through orig_index, like orig_subimage_coordinates[curr_orig_index] in the reference.
 * params12 = {net_Dx, net_Dy, net_Dang, regression_width, regression_height, min_scale_radio,
 *             max_scale_radio, tolerance_posxy, tolerance_scale, tolerance_angle, desired_sampling,
 *             cut_off_face}.  float64 throughout, reference operation order, no FMA contraction.
 * ---------------------------------------------------------------------------------------------- */
int hgsfa_cascade_update_device(int type, double* d_coords, double* d_angles, const double* d_reg_out,
                                const double* d_orig_coords, const double* d_orig_angles,
                                const int32_t* d_orig_index, const double* d_patch_wh, int64_t n,
                                const double* params12, uint8_t* d_keep, double* d_conf, void* stream);

/* src_index[j] = index of the j-th kept element (ascending), *d_count = number kept (device int64).
 * d_scratch: at least ceil(n / 1024) int32. */
int hgsfa_compact_index_device(const uint8_t* d_keep, int64_t n, int32_t* d_src_index, int64_t* d_count,
                               int32_t* d_scratch, int64_t scratch_ints, void* stream);

/* dst[j] = src[index[j]] for rows of row_bytes bytes (multiple of 4): curr_x = curr_x[new_wrong_images == 0] */
int hgsfa_gather_rows_device(const void* d_src, void* d_dst, const int32_t* d_index, int64_t n_out,
                             int64_t row_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HGSFA_H_ */
