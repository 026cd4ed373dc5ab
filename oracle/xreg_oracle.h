/*
 * xreg_oracle.h -- CPU restatement of xReg's RayCasterLineIntCPU and
 * ImgSimMetric2D{NCC,GradNCC,PatchNCC,PatchGradNCC}CPU arithmetic.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, the smoke()
 * check in __graft_entry__.py and bench.py's cpu_baseline / --impl reference
 * legs may load it.  The shipped library (libxreg_cuda.so) never links or
 * calls anything in oracle/.
 *
 * PINNING.  The reference ships no golden vectors, known-answer tests or
 * fixtures for this path, and as a whole it cannot
 * be compiled here (ITK / Eigen / OpenCV / TBB are absent).
 *   DRR (xo_drr): PINNED TO THE REFERENCE'S OWN SOURCE.  oracle/ref_pin/ compiles
 *     RayRectIntersect, CameraModel::ind_pt_to_phys_det_pt and the line-integral
 *     kernels + ComputeLineInts<Kernel> from /root/reference where they lie, over
 *     functional stand-ins for the Eigen / ITK / TBB types (our restatement of the
 *     un-vendored dependencies, conventions stated in ref_pin_prelude.h), into
 *     oracle/_ref/libxreg_refslice.so; tests/test_oracle_ref_slice.py requires
 *     xo_drr to equal that code bit for bit on every pixel of 48 random scenes
 *     (both kernels, all frame types, REPLACE and ACCUM).
 *   Patch NCC (xo_patch_weights, xo_patch_ncc; the core of patch gradient-NCC,
 *     which the reference composes from two patch-NCC objects on the gradient
 *     images): PINNED TO THE REFERENCE'S OWN SOURCE the same way --
 *     ImgSimMetric2DPatchCommon::{setup_patches, compute_weights,
 *     patch_indices_to_use}, ImgSimMetric2DPatchNCCCPU::{allocate_resources,
 *     compute, process_mask} and detail::ComputePatchMeanStdDev compiled over
 *     cv::Mat / itk::Image stand-ins that carry no arithmetic
 *     (oracle/_ref/libxreg_refslice_metric.so); patch grid, weights, per-patch
 *     values and image scores agree bit for bit on 40 random cases x every option
 *     combination (tests/test_oracle_ref_slice.py).
 *   NCC (xo_ncc; gradient-NCC is two of them on the gradient images): the class
 *     code is pinned the same way (libxreg_refslice_ncc.so).  Masked: plain scalar
 *     loops in the reference -> bit for bit, no convention.  Unmasked: three Eigen
 *     reductions, for which stand-in and oracle follow the same documented Eigen 3.3
 *     SSE reduction shape -> pinned up to that convention.
 *   Gradient-NCC and patch gradient-NCC (xo_grad_ncc, xo_patch_grad_ncc): the
 *     reference's GradImgCPU / GradNCCCPU / PatchGradNCCCPU classes compile the same
 *     way (libxreg_refslice_grad.so) with cv::GaussianBlur / cv::Sobel as call-outs:
 *     bit for bit with the oracle's filters installed, <= 2e-6 with the real OpenCV
 *     (cv2) installed.
 *   HU -> linear attenuation (xo_hu_to_lin_att): pinned to
 *     HUToLinAttFilter::GenerateData (libxreg_refslice_hu.so).
 *   SSD (xo_ssd): class code pinned like unmasked NCC (one Eigen reduction under
 *     the stated convention).
 *   Gaussian / Sobel gradient images (xo_gauss_blur, xo_sobel): PARITY UNPINNED by
 *     the reference (OpenCV's filters); pinned instead by OpenCV's Python binding
 *     (Sobel bit-exact, Gaussian <= 1 ulp) and known answers
 *     (tests/test_oracle_metrics.py).
 *
 * Every function cites the reference file:line it follows (paths relative to
 * the reference checkout).  Arithmetic is single precision wherever the
 * reference is single precision and double where ITK is double; the evaluation
 * order written here is the frozen definition of "bit-exact" for the ray/box
 * masks and step counts (the reference's true Eigen/SSE order is unknowable
 * without Eigen).  Build with -ffp-contract=off and without -march flags so
 * no FMA contraction happens (mirrors the reference's plain Release build).
 */
#ifndef XREG_ORACLE_H
#define XREG_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Same POD layout as xrc_cam in include/xreg_cuda.h
 * (CameraModel, lib/transforms/xregPerspectiveXform.h:108-173). */
typedef struct xo_cam
{
  uint32_t rows;
  uint32_t cols;
  float intrins_inv[9];   /* row-major 3x3 */
  float extrins_inv[12];  /* row-major 3x4 */
  float pinhole[3];
  float focal_len;
  int32_t frame_type;     /* 0 = DET_POS_Z, 1 = DET_NEG_Z, 2 = ORIGIN_ON_DETECTOR */
} xo_cam;

enum { XO_KERNEL_SUM = 0, XO_KERNEL_MAX = 1 };
enum { XO_STORE_REPLACE = 0, XO_STORE_ACCUM = 1 };

/* ---- small f32 geometry helpers (frozen evaluation order) ---- */
void xo_affine_inverse(const float a[12], float out[12]);
void xo_affine_compose(const float a[12], const float b[12], float out[12]);
void xo_mat3_inverse(const float m[9], float out[9]);

/* MakeNaiveIntrins + CameraModel::setup(focal_len, nr, nc, rs, cs)
 * (lib/transforms/xregPerspectiveXform.cpp:200-254). extrinsic = identity. */
void xo_cam_setup_naive(xo_cam* cam, float focal_len, uint32_t rows, uint32_t cols,
                        float row_spacing, float col_spacing, int32_t frame_type);

/* CameraModel::setup(intrins, extrins, ...) (xregPerspectiveXform.cpp:302-334).
 * intrins row-major 3x3, extrins row-major 4x4 (rigid). */
void xo_cam_setup(xo_cam* cam, const float intrins[9], const float extrins[16],
                  uint32_t rows, uint32_t cols, float row_spacing, float col_spacing,
                  int32_t frame_type);

/* RayCaster::distribute_xforms_among_cam_models (xregRayCastInterface.cpp:97-114) */
void xo_distribute_xforms(const float* poses, uint32_t n_poses, uint32_t n_cams,
                          float* out_poses, uint32_t* out_cam_idx);

/* RayCasterCPU::pre_compute (xregRayCastBaseCPU.cpp:128-158).
 * bg_projs: n_cams pointers or NULL. */
void xo_pre_compute(float* buf, uint32_t n_projs, uint32_t rows, uint32_t cols,
                    const uint32_t* cam_idx, const float* const* bg_projs,
                    int store_method, float default_bg);

/* RayCasterLineIntCPU::compute / ComputeLineInts (xregRayCastLineIntCPU.cpp:105-349),
 * linear interpolation, no anti-aliasing.  buf is read-modify-written
 * (buf = K(buf, val)); call xo_pre_compute first.
 * Optional outputs (may be NULL): hit_mask[n_projs*rows*cols] (1 = ray clipped
 * to the volume and marched), num_steps_out (num_steps+1 for hit rays, else 0),
 * total_samples = S of SURVEY 8(d).  n_threads <= 0: all cores. */
int xo_drr(const float* vol, const uint64_t dims[3], const float idx_to_phys[12],
           const xo_cam* cams, uint32_t n_cams,
           const float* poses, const uint32_t* cam_idx, uint32_t n_projs,
           float step_size, int kernel_id,
           float* buf, uint8_t* hit_mask, uint32_t* num_steps_out,
           uint64_t* total_samples, int n_threads);

/* the same with the interpolator selectable (RayCaster::InterpMethod, xregRayCastInterface.h: kRAY_CAST_INTERP_LINEAR = 0,
 * kRAY_CAST_INTERP_NN = 1; sinc / B-spline are not restated: -3) */
#define XO_INTERP_LINEAR 0
#define XO_INTERP_NN 1
int xo_drr_interp(const float* vol, const uint64_t dims[3], const float idx_to_phys[12],
                  const xo_cam* cams, uint32_t n_cams,
                  const float* poses, const uint32_t* cam_idx, uint32_t n_projs,
                  float step_size, int kernel_id, int interp,
                  float* buf, uint8_t* hit_mask, uint32_t* num_steps_out,
                  uint64_t* total_samples, int n_threads);

/* RayCasterDepthCPU::compute (lib/ray_cast/xregRayCastDepthCPU.cpp:42-272): depth of the first sample >= collision_thresh
 * along every ray, refined by num_backtracking_steps step halvings; buf = min(buf, depth) (initialise buf with
 * kRAY_CAST_MAX_DEPTH = 1e37 via xo_pre_compute, as the class does). */
#define XO_RAY_CAST_MAX_DEPTH 1.0e37f
int xo_depth(const float* vol, const uint64_t dims[3], const float idx_to_phys[12],
             const xo_cam* cams, uint32_t n_cams,
             const float* poses, const uint32_t* cam_idx, uint32_t n_projs,
             float step_size, int interp, float collision_thresh, uint32_t num_backtracking_steps,
             float* buf, int n_threads);

/* Log remap of a projection: ImageIntensLogTransFilter::GenerateData (lib/image/xregImageIntensLogTrans.cpp:55-144), and
 * the ITK smoothing behind its default I0 (itk::DiscreteGaussianImageFilter, restated; PARITY UNPINNED for that piece). */
int xo_itk_gaussian_coeffs(double variance, double max_error, int max_width, double* coeffs);
void xo_itk_discrete_gaussian_2d(const float* img, uint32_t rows, uint32_t cols, double variance, float* out);
void xo_log_remap(const float* img, uint32_t rows, uint32_t cols, int normalize_zero_one, int use_max_intensity_as_I0,
                  float I0, const float* smoothed, float* out, float* I0_used);

/* DownsampleImage (lib/itk/xregITKResampleUtils.h:49-112, cubic B-spline default :181-188), what DownsampleProjData
 * (lib/image/xregProjData.cpp:40-99) applies to a projection: ITK's Gaussian smoothing + B-spline resampling restated
 * (PARITY UNPINNED: all of the arithmetic is ITK's).  sigma < 0: the default 0.5 / factor.  out: xo_downsample_size. */
void xo_downsample_size(uint32_t rows, uint32_t cols, double factor, uint32_t* out_rows, uint32_t* out_cols);
void xo_downsample_image(const float* img, uint32_t rows, uint32_t cols, double factor, double sigma, float* out);

/* ITK LinearInterpolateImageFunction::EvaluateOptimized(Dispatch<3>) restated. */
double xo_interp_linear(const float* vol, const uint64_t dims[3], const float x[3]);
/* ITK NearestNeighborInterpolateImageFunction::EvaluateAtContinuousIndex restated. */
double xo_interp_nn(const float* vol, const uint64_t dims[3], const float x[3]);

/* ---- similarity metrics ---- */

/* ImgSimMetric2DNCCCPU (xregImgSimMetric2DNCCCPU.cpp:52-236).  mov is
 * overwritten with the zero-mean images exactly like the reference.
 * mask may be NULL. */
void xo_hu_to_lin_att(const float* hu, float* att, uint64_t n, float hu_lower);
void xo_ssd(const float* fixed, const uint8_t* mask, uint32_t rows, uint32_t cols,
            float* mov, uint32_t n_imgs, float* sims, int n_threads);
void xo_ncc(const float* fixed, const uint8_t* mask, uint32_t rows, uint32_t cols,
            float* mov, uint32_t n_imgs, float* sims, int n_threads);

/* cv::getGaussianKernel(k, 0, CV_32F) */
int xo_gauss_kernel(int width, float* coeffs);

/* cv::GaussianBlur(k x k, sigma 0) then cv::Sobel dx / dy, BORDER_REFLECT_101
 * (xregImgSimMetric2DGradImgCPU.cpp:32-102).  width 0 disables smoothing. */
void xo_gauss_blur(const float* img, uint32_t rows, uint32_t cols, int width, float* out);
void xo_sobel(const float* img, uint32_t rows, uint32_t cols, float* gx, float* gy);
void xo_grad_imgs(const float* img, uint32_t rows, uint32_t cols, int gauss_width,
                  float* gx, float* gy);

/* ImgSimMetric2DGradNCCCPU (xregImgSimMetric2DGradNCCCPU.cpp:29-65) */
void xo_grad_ncc(const float* fixed, const uint8_t* mask, uint32_t rows, uint32_t cols,
                 int gauss_width, const float* mov, uint32_t n_imgs, float* sims,
                 int n_threads);

/* Patch options (ImgSimMetric2DPatchCommon.h:141-166 defaults in comments) */
typedef struct xo_patch_opts
{
  uint32_t radius;                    /* 5 */
  uint32_t stride;                    /* 1 */
  int32_t compute_mean_of_patch_sims; /* 0 */
  int32_t weight_patch_sims;          /* 1 */
  int32_t use_mask_for_weighting;     /* 1 */
  int32_t use_mask_for_patch_stats;   /* 0 */
  int32_t normalize_weights_as_prob;  /* 1 */
} xo_patch_opts;

/* number of patches of the grid (xregImgSimMetric2DPatchCommon.cpp:269-292) */
uint64_t xo_num_patches(uint32_t rows, uint32_t cols, uint32_t radius, uint32_t stride);

/* ImgSimMetric2DPatchCommon::compute_weights (PatchCommon.cpp:309-410);
 * wgt_img may be NULL, mask may be NULL.  weights[num_patches] out. */
void xo_patch_weights(uint32_t rows, uint32_t cols, const xo_patch_opts* o,
                      const uint8_t* mask, const float* wgt_img, float* weights);

/* ImgSimMetric2DPatchNCCCPU (xregImgSimMetric2DPatchNCCCPU.cpp:74-300,332-441,558-619).
 * weights: per-patch (as returned by xo_patch_weights) or NULL = all 1.
 * patch_sims (optional, n_imgs x num_patches): 1 - acc_k per patch. */
void xo_patch_ncc(const float* fixed, const uint8_t* mask, uint32_t rows, uint32_t cols,
                  const xo_patch_opts* o, const float* weights,
                  const float* mov, uint32_t n_imgs, float* sims, float* patch_sims,
                  int n_threads);

/* ImgSimMetric2DPatchGradNCCCPU (xregImgSimMetric2DPatchGradNCCCPU.cpp:34-253) */
void xo_patch_grad_ncc(const float* fixed, const uint8_t* mask, uint32_t rows, uint32_t cols,
                       int gauss_width, const xo_patch_opts* o, const float* weights,
                       const float* mov, uint32_t n_imgs, float* sims, int n_threads);
/* the same metrics over a patch subset (set_patches_to_use / random patches; local order = subset order) */
void xo_patch_ncc_subset(const float* fixed, const uint8_t* mask, uint32_t rows, uint32_t cols,
                         const xo_patch_opts* o, const float* weights, const uint64_t* subset, uint64_t n_subset,
                         const float* mov, uint32_t n_imgs, float* sims, int n_threads);
void xo_patch_grad_ncc_subset(const float* fixed, const uint8_t* mask, uint32_t rows, uint32_t cols, int gauss_width,
                              const xo_patch_opts* o, const float* weights, const uint64_t* subset, uint64_t n_subset,
                              const float* mov, uint32_t n_imgs, float* sims, int n_threads);

/* ImgSimMetric2DCombineMean (xregImgSimMetric2DCombine.cpp:67-86) */
void xo_combine_mean(const float* view_sims, uint32_t n_views, uint32_t n_poses, float* out);

int xo_num_threads(void);
/* omp_set_num_threads: overrides an inherited OMP_NUM_THREADS (torchrun exports 1 to every rank) */
void xo_set_num_threads(int n);

#ifdef __cplusplus
}
#endif
#endif
