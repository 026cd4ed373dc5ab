/*
 * xreg_cuda.h -- C ABI of the B200 (sm_100a) DRR ray caster and 2D similarity
 * metrics that replace xReg's OpenCL backend on the intensity-registration
 * hot path (the ...OCL classes of lib/ray_cast and lib/regi/sim_metrics_2d, and
 * all of lib/opencl).
 *
 * Plain C: opaque handles, POD arguments, int status codes.  No torch, ITK,
 * Eigen or OpenCV types cross this boundary.  Paths in comments are relative
 * to the reference checkout (rg2/xreg 2021.09.19.1).
 *
 * Semantics shared by all entry points
 *   - One context = one CUDA device + one stream.  All work of the ray casters
 *     and metrics created from a context is enqueued on that stream in call
 *     order; entry points that return host-visible results
 *     (xrc_rc_read_projs, xrc_sm_read_sims, xrc_eval_batch, ...) synchronise.
 *     Like the reference classes, objects are not thread safe.
 *   - Return value 0 = XRC_OK; anything else is an error and
 *     xrc_last_error() returns a thread-local description.  The C++ adapters
 *      rethrow these as the reference's exception types
 *     (lib/common/xregExceptionUtils.h:36-63, xregRayCastInterface.h:59).
 *   - There is no CPU fallback: every compute entry point fails with
 *     XRC_ERR_CUDA when no sm_100-class device is usable.
 *   - Matrices are row-major.  A rigid / affine transform is the 3x4 top of
 *     the 4x4 matrix (12 floats: r00 r01 r02 tx r10 ...).
 */
#ifndef XREG_CUDA_H
#define XREG_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define XRC_VERSION 100

enum xrc_status
{
  XRC_OK = 0,
  XRC_ERR_INVALID = 1,      /* bad argument / call order: xregASSERT failures in the reference */
  XRC_ERR_UNSUPPORTED = 2,  /* RayCaster::UnsupportedOperationException (xregRayCastInterface.h:59) */
  XRC_ERR_CUDA = 3,         /* CUDA runtime / driver failure, or no usable device */
  XRC_ERR_NOMEM = 4
};

typedef struct xrc_ctx xrc_ctx;
typedef struct xrc_rc xrc_rc;
typedef struct xrc_sm xrc_sm;

/* CameraModel (lib/transforms/xregPerspectiveXform.h:108-173): exactly the
 * members RayCasterLineIntCPU reads (xregRayCastLineIntCPU.cpp:185-251 via
 * ind_pt_to_phys_det_pt, xregPerspectiveXform.cpp:391-414). */
typedef struct xrc_cam
{
  uint32_t rows;          /* num_det_rows */
  uint32_t cols;          /* num_det_cols */
  float intrins_inv[9];   /* 3x3 */
  float extrins_inv[12];  /* 3x4 camera -> camera-world */
  float pinhole[3];       /* pinhole_pt */
  float focal_len;
  int32_t frame_type;     /* CameraCoordFrame: 0 DET_POS_Z, 1 DET_NEG_Z, 2 ORIGIN_ON_DETECTOR */
} xrc_cam;

/* RayCaster::InterpMethod (xregRayCastInterface.h:61-67).  LINEAR (the default and the optimised path) and NN
 * (nearest neighbour: the voxel at floor(x + 0.5) per axis, ITK's ConvertContinuousIndexToNearestIndex; sums
 * bit-identical to the CPU class, xregRayCastLineIntCPU.cpp:128-130; plain xrc_rc_compute only, not tile-sharded) are
 * implemented; SINC and BSPLINE return XRC_ERR_UNSUPPORTED (the reference's OpenCL backend supports linear only,
 * xregRayCastBaseOCL.cpp:338-341). */
enum { XRC_INTERP_LINEAR = 0, XRC_INTERP_NN = 1, XRC_INTERP_SINC = 2, XRC_INTERP_BSPLINE = 3 };
/* RayCaster::ProjPixelStoreMethod (xregRayCastInterface.h:71-75) */
enum { XRC_STORE_REPLACE = 0, XRC_STORE_ACCUM = 1 };
/* RayCastLineIntKernel (xregRayCastInterface.h:575-579) */
enum { XRC_KERNEL_SUM = 0, XRC_KERNEL_MAX = 1 };
/* Metric kinds: ImgSimMetric2D{NCC,GradNCC,PatchNCC,PatchGradNCC}CPU/OCL, and
 * ImgSimMetric2DSSDCPU/OCL: sum over ALL pixels of (fixed - moving)^2 / num_pixels with both images zeroed outside
 * the mask (xregImgSimMetric2DSSDCPU.cpp:62-88) */
enum { XRC_SM_NCC = 0, XRC_SM_GRAD_NCC = 1, XRC_SM_PATCH_NCC = 2, XRC_SM_PATCH_GRAD_NCC = 3, XRC_SM_SSD = 4 };
/* Volume layouts in HBM (DESIGN.md "Data layout").  All fetch exact f32 voxels and take the same samples;
 * LINEAR..TEX lerp x, y, z and agree bit for bit, PAX (the default: one padded XY-quad stack per principal
 * ray axis) lerps mid axis, slow axis, fast axis and differs from them in the last ulp only. */
enum { XRC_LAYOUT_DEFAULT = -1, XRC_LAYOUT_LINEAR = 0, XRC_LAYOUT_QUAD = 1, XRC_LAYOUT_TEX_QUAD = 2,
       XRC_LAYOUT_OCT = 3, XRC_LAYOUT_TEX = 4, XRC_LAYOUT_PAX = 5 };

const char* xrc_last_error(void);
int xrc_version(void);
/* number of kernels this library has launched in this process (for bench.py's gpu_launches) */
uint64_t xrc_launch_count(void);

/* ---- context: replaces lib/opencl/xregOpenCLSys.cpp:31-84 (device pick) and
 * ProgOpts::selected_ocl_ctx_queue (lib/common/xregProgOptUtils.cpp:1712-1725) ---- */
int xrc_ctx_create(int device, xrc_ctx** out);
/* Use a caller-owned cudaStream_t (e.g. torch's current stream) instead of a private one. */
int xrc_ctx_create_on_stream(int device, void* cuda_stream, xrc_ctx** out);
int xrc_ctx_destroy(xrc_ctx* ctx);
int xrc_ctx_synchronize(xrc_ctx* ctx);
int xrc_ctx_device(const xrc_ctx* ctx, int* device);
/* the cudaStream_t all work of this context is ordered on */
int xrc_ctx_stream(const xrc_ctx* ctx, void** cuda_stream);

/* ---- ray caster: replaces RayCasterOCL / RayCasterLineIntOCL
 * (lib/ray_cast/xregRayCastBaseOCL.cpp, xregRayCastLineIntOCL.cpp) behind
 * RayCaster (lib/ray_cast/xregRayCastInterface.h:43-434) ---- */
int xrc_rc_create(xrc_ctx* ctx, xrc_rc** out);
int xrc_rc_destroy(xrc_rc* rc);

/* Tuning knob, call before xrc_rc_set_volumes.  Default: XRC_LAYOUT_DEFAULT. */
int xrc_rc_set_layout(xrc_rc* rc, int layout);
/* Tuning knob: CTA launch order, 0 = projection fastest (default), 1 = detector tile fastest. */
int xrc_rc_set_cta_order(xrc_rc* rc, int order);

/* Device memory the ray caster's volume representation occupies right now (all volumes; payload stacks, the f32 source
 * kept for stacks built on demand, the empty-space maps).  Reported in bench.py's config. */
int xrc_rc_volume_bytes(const xrc_rc* rc, uint64_t* bytes);
/* The layout volume vol_idx really uses (XRC_LAYOUT_*): with XRC_LAYOUT_DEFAULT the library chooses the principal-axis
 * stacks and falls back to one XY-quad stack when the record index does not fit 32 bits or device memory runs out. */
int xrc_rc_volume_layout(const xrc_rc* rc, uint32_t vol_idx, int* layout);
/* Log remap of a projection (pre-processing): ImageIntensLogTransFilter
 * (lib/image/xregImageIntensLogTrans.{h,cpp}; what ProjPreProc applies to every fluoroscopic image,
 * lib/image/xregProjPreProc.cpp:63-84) on the device: out = -log(x / I0) for x > 1e-6, and the value of the smallest such
 * pixel for the others.  normalize_zero_one: the image is scaled by 1 / max first (SetNormalizeZeroOne);
 * use_max_intensity_as_I0 (the filter's default): I0 = the maximum of the image smoothed by a discrete Gaussian of
 * variance 2 (itk::DiscreteGaussianImageFilter -- ITK is an un-vendored dependency of the reference, its published
 * algorithm is restated), or 1 after normalisation; else the given I0 (SetI0).  host_img / host_out:
 * rows x cols floats (may alias); I0_used (optional) returns the I0 of the map.  Synchronises. */
int xrc_log_remap(xrc_ctx* ctx, const float* host_img, uint32_t rows, uint32_t cols, int normalize_zero_one,
                  int use_max_intensity_as_I0, float I0, float* host_out, float* I0_used);
/* Down-sampling of a projection image (pre-processing): DownsampleImage
 * (lib/itk/xregITKResampleUtils.h:49-112 with the cubic B-spline default of :181-188), the image half of DownsampleProjData
 * (lib/image/xregProjData.cpp:40-99; the camera half is DownsampleCameraModel, plain host arithmetic) that
 * MultiLevelMultiObjRegi applies to every fixed image at every level: Gaussian smoothing with sigma (< 0: 0.5 / factor;
 * |sigma| <= 1e-6 or factor >= 1: none), then resampling at the continuous input indices i / factor with a cubic B-spline,
 * 0 outside the buffer; out_rows x out_cols = xrc_downsample_size (size * factor + 0.5, truncated).  All of the
 * arithmetic is ITK 5.1.1's (DiscreteGaussianImageFilter, BSplineDecompositionImageFilter,
 * BSplineInterpolateImageFunction, ResampleImageFilter) -- un-vendored: restated from the published algorithms,
 * DESIGN.md section 4.8.  Synchronises. */
int xrc_downsample_size(uint32_t rows, uint32_t cols, double factor, uint32_t* out_rows, uint32_t* out_cols);
int xrc_downsample_image(xrc_ctx* ctx, const float* host_img, uint32_t rows, uint32_t cols, double factor, double sigma,
                         float* host_out);
/* RayCasterDepthCPU::compute (lib/ray_cast/xregRayCastDepthCPU.cpp:42-272) on this ray caster's
 * volumes, cameras and poses: per pixel the depth -- distance from the pinhole, in the camera frame -- of the first sample
 * along the (unlimited) ray whose interpolated value is >= collision_thresh, refined by num_backtracking_steps halvings
 * of the step; out = min(initial value, depth) with the initial value as for xrc_rc_compute (the class's default
 * background is kRAY_CAST_MAX_DEPTH: xrc_rc_set_params(..., default_bg = XRC_RAY_CAST_MAX_DEPTH)); rays that never reach
 * the threshold keep the initial value.  Linear (ITK's arithmetic itself: f64 lerps of the f32 voxels, so that every
 * threshold decision is the CPU class's) or nearest-neighbour interpolation.  Bit-identical to the CPU class.
 * The class defaults are collision_thresh 150, num_backtracking_steps 0 (xregRayCastInterface.cpp:427-428). */
#define XRC_RAY_CAST_MAX_DEPTH 1.0e37f
int xrc_rc_compute_depth(xrc_rc* rc, uint32_t vol_idx, float collision_thresh, uint32_t num_backtracking_steps);
/* Empty-space trimming, default on.  Samples whose 8 corner voxels are all zero add +0 to the
 * sequential f32 sum of xregRayCastLineIntCPU.cpp:270-279, so the sum kernel does not fetch the leading
 * and trailing samples of a ray that a per-volume block map proves to be zero (air around the body,
 * everything outside the bone mask).  Results are bit-identical with trimming on or off; 0 turns it off
 * (measurement), 1 (or 2) = on.  3 = on, and the kernel also skips runs of empty samples INSIDE a ray's range -- the air
 * between two separated structures along the view direction -- warp by warp, marching in segments of 16 samples between
 * looks at the map (same bits again; pays when such gaps are long: marching in segments costs ~5 % where there are none,
 * so it is a request, not the default).  The max kernel never trims. */
int xrc_rc_set_skip_empty(xrc_rc* rc, int enable);

/* RayCaster::set_volumes (xregRayCastInterface.h:90) + vols_changed (:425).
 * host_ptrs[i]: x-fastest float volume of dims[i] = {nx, ny, nz}; copied to the
 * device (the caller's memory is not referenced afterwards).
 * idx_to_phys[i]: ITKImagePhysicalPointTransformsAsEigen
 * (lib/itk/xregITKBasicImageUtils.h:131-168): M = Dir * diag(spacing), t = origin,
 * already cast to float. */
int xrc_rc_set_volumes(xrc_rc* rc, uint32_t n, const float* const* host_ptrs,
                       const uint64_t (*dims)[3], const float (*idx_to_phys)[12]);
/* the volumes are in Hounsfield units; convert them to linear attenuation on the device
 * while loading (HUToLinAtt, lib/image/xregHUToLinAtt.cpp:45-69: max(hu * (mu_water - mu_air) / 1000 + mu_water -
 * mu_lower, 0) in double, mu_lower = the value of hu_lower, reference default -1000) -- what the registration
 * apps do on the host before set_volumes.  Bit-identical to converting on the host first. */
int xrc_rc_set_volumes_hu(xrc_rc* rc, uint32_t n, const float* const* host_ptrs, const uint64_t (*dims)[3],
                          const float (*idx_to_phys)[12], float hu_lower);
/* same, but the source volume already lives on this context's device */
int xrc_rc_set_volumes_device(xrc_rc* rc, uint32_t n, const float* const* dev_ptrs,
                              const uint64_t (*dims)[3], const float (*idx_to_phys)[12]);

/* RayCaster::set_camera_models (xregRayCastInterface.h:110); all cameras must
 * share rows/cols (xregRayCastBaseCPU.cpp:60-70). */
int xrc_rc_set_cameras(xrc_rc* rc, uint32_t n, const xrc_cam* cams);

/* RayCaster::set_num_projs + allocate_resources (xregRayCastInterface.h:158,256):
 * capacity is max_projs. */
int xrc_rc_allocate(xrc_rc* rc, uint32_t max_projs);
/* RayCaster::set_num_projs after allocation (n <= capacity, xregRayCastInterface.cpp:131-139) */
int xrc_rc_set_num_projs(xrc_rc* rc, uint32_t n);
int xrc_rc_num_projs(const xrc_rc* rc, uint32_t* n);
/* RayCaster::max_num_projs_possible (xregRayCastBaseOCL.cpp:235-246): bounded by free HBM */
int xrc_rc_max_projs_possible(const xrc_rc* rc, uint64_t* n);

/* RayCaster::set_xforms_cam_to_itk_phys + set_camera_model_proj_associations
 * (xregRayCastInterface.h:173-212).  cam_to_phys: n x 12, cam_idx: n (NULL = all 0).
 * n must equal the current num_projs. */
int xrc_rc_set_poses(xrc_rc* rc, uint32_t n, const float* cam_to_phys, const uint32_t* cam_idx);
/* RayCaster::distribute_xforms_among_cam_models (xregRayCastInterface.cpp:97-114):
 * n_poses * n_cams must equal num_projs; camera-major replication. */
int xrc_rc_distribute_poses(xrc_rc* rc, uint32_t n_poses, const float* cam_to_phys);
/* Poses that already live on the device (e.g. written by a GPU-resident optimiser, or
 * a torch tensor): the ray caster reads dev_cam_to_phys (n x 12 floats) and dev_cam_idx
 * (n x uint32, NULL = camera 0 for all) in place at every compute() until the next
 * xrc_rc_set_poses / xrc_rc_distribute_poses call.  No copy, no synchronisation. */
int xrc_rc_set_poses_device(xrc_rc* rc, uint32_t n, const float* dev_cam_to_phys,
                            const uint32_t* dev_cam_idx);
/* Same, with a host copy of the same values (n x 12 floats, optional n camera ids): the device arrays are what the
 * kernels read; the mirror only lets the host see which principal-axis stacks of the volume these poses need, so that
 * it builds just those (without a mirror, device poses make the ray caster build all three stacks: 12x instead of
 * 4x + 1x the f32 volume for a single view).  A mirror that disagrees with the device arrays costs speed and
 * last-bit reproducibility, never correctness. */
int xrc_rc_set_poses_device_mirrored(xrc_rc* rc, uint32_t n, const float* dev_cam_to_phys, const uint32_t* dev_cam_idx,
                                     const float* host_cam_to_phys, const uint32_t* host_cam_idx);

/* set_ray_step_size / set_interp_method / RayCastLineIntParamInterface::set_kernel_id /
 * set_proj_store_method / set_default_bg_pixel_val (xregRayCastInterface.h:140-350,581-591) */
int xrc_rc_set_params(xrc_rc* rc, float step_size, int interp, int kernel_id,
                      int store_method, float default_bg);
/* set_use_bg_projs / set_bg_projs (xregRayCastInterface.h:330-350): one host image per camera */
int xrc_rc_set_bg_projs(xrc_rc* rc, const float* const* host_imgs, int use_bg);

/* RayCaster::compute(vol_idx) (xregRayCastInterface.h:259; CPU semantics of
 * xregRayCastLineIntCPU.cpp:294-349 incl. pre_compute).  Asynchronous on the
 * context stream. */
int xrc_rc_compute(xrc_rc* rc, uint32_t vol_idx);

/* Device pointer of the projection buffer: what RayCastSyncOCLBufFromOCL hands
 * to a same-device metric (lib/ray_cast/xregRayCastSyncBuf.cpp:155-159). */
int xrc_rc_device_buf(xrc_rc* rc, float** dev_ptr);
/* RayCaster::proj / raw_host_pixel_buf / to_host_buf()->sync()
 * (xregRayCastBaseCPU.cpp:90-126, xregRayCastSyncBuf.cpp:60-110): D2H of
 * projections [first, first+count). Synchronises. */
int xrc_rc_read_projs(xrc_rc* rc, uint32_t first, uint32_t count, float* host_dst);
/* RayCaster::use_other_proj_buf (xregRayCastInterface.h:318) */
int xrc_rc_use_other_proj_buf(xrc_rc* rc, xrc_rc* other);

/* Parity / roofline instrumentation.  Runs the same ray set-up code as compute()
 * for the current poses and returns, per ray, the clip mask (1 = marched) and
 * num_steps+1 (0 for missed rays); either pointer may be NULL.  total = S of
 * SURVEY 8(d).  Synchronises. */
int xrc_rc_ray_info(xrc_rc* rc, uint32_t vol_idx, uint8_t* host_mask, uint32_t* host_steps,
                    uint64_t* total_samples);

/* Instrumentation: the number of trilinear samples compute() fetches for the current poses
 * (<= total_samples of xrc_rc_ray_info; equal when trimming is off).  Projections are not touched.
 * Synchronises. */
int xrc_rc_fetched_samples(xrc_rc* rc, uint32_t vol_idx, uint64_t* fetched);

/* ---- similarity metrics: replace ImgSimMetric2D*OCL
 * (lib/regi/sim_metrics_2d/xregImgSimMetric2D{NCC,GradImg,GradNCC,PatchNCC,PatchGradNCC}OCL.cpp)
 * behind ImgSimMetric2D (xregImgSimMetric2D.h:42-156) ---- */
int xrc_sm_create(xrc_ctx* ctx, int kind, xrc_sm** out);
int xrc_sm_destroy(xrc_sm* sm);

/* set_fixed_image (:69): row-major rows x cols floats, copied. */
int xrc_sm_set_fixed(xrc_sm* sm, const float* host_img, uint32_t rows, uint32_t cols);
/* set_mask (:130): uint8 rows x cols or NULL; may be called again later
 * (process_updated_mask semantics). */
int xrc_sm_set_mask(xrc_sm* sm, const uint8_t* host_mask);
/* ImgSimMetric2DGradImgParamInterface::set_smooth_img_before_sobel_kernel_radius
 * (xregImgSimMetric2DGradImgParamInterface.h:31-39): really the kernel WIDTH
 * (odd, 0 = off, default 5). */
int xrc_sm_set_grad_params(xrc_sm* sm, uint32_t gauss_width);
/* ImgSimMetric2DPatchCommon parameters (xregImgSimMetric2DPatchCommon.h:67-127).
 * weights: one float per patch of the full grid in row-major centre order (the
 * values compute_weights() would leave in patch_infos_[k].weight), or NULL for
 * all-ones.  Host patch-weight logic stays in the adapter. */
int xrc_sm_set_patch_params(xrc_sm* sm, uint32_t radius, uint32_t stride,
                            int compute_mean_of_patch_sims, int weight_patch_sims_in_combine,
                            int use_mask_for_patch_stats, const float* weights,
                            uint64_t n_weights);

/* How the patch metrics combine the per-patch values into the image score.  The reference adds them with a
 * sequential f32 loop and divides by an f32 total weight accumulated the same way
 * (xregImgSimMetric2DPatchNCCCPU.cpp:262-287); that sum's rounding error is part of its result (up to ~3e-5 of the
 * value with mask-coverage weights, ~5e-6 rms at 200 000 unweighted patches).
 *   XRC_COMBINE_REFERENCE (default): the same sequential f32 sum, reproduced bit for bit by a parallel kernel
 *       (sim.cu: patch_seqsum_kernel); agrees with the CPU class to ~1e-7.
 *   XRC_COMBINE_REFERENCE_SERIAL: the literal one-thread loop (verification of the parallel emulation; slow).
 *   XRC_COMBINE_F64: f64 sums and divisor: closest to exact arithmetic, not to the reference; no per-patch buffer.
 * Patch kinds only; ignored by the others.  May be changed between computes. */
enum { XRC_COMBINE_REFERENCE = 0, XRC_COMBINE_REFERENCE_SERIAL = 1, XRC_COMBINE_F64 = 2 };
int xrc_sm_set_combine_mode(xrc_sm* sm, int mode);
/* Instrumentation: the kernel behind XRC_COMBINE_REFERENCE on caller data.  host_vals: n_seq sequences of n floats;
 * host_out[s] = (((0 + v[s][0]) + v[s][1]) + ...) with one f32 rounding per addition.  serial: 0 the parallel
 * emulation (binades predicted from a float64 prefix sum, one cheap dependent pass), 1 the literal one-thread loop,
 * 2 the round-to-round chained emulation (what sequences too long for the first fall back to), 10 + c the first with
 * a cluster of c = 1, 2, 4 or 8 CTAs per sequence instead of the automatic choice.  Synchronises. */
int xrc_seqsum_f32(xrc_ctx* ctx, const float* host_vals, uint32_t n_seq, uint64_t n, int serial, float* host_out);

/* ImgSimMetric2DPatchCommon::set_patches_to_use / reset_patches_to_use and the random patches of
 * patch_indices_to_use (xregImgSimMetric2DPatchCommon.cpp:231-241, 413-493): the metric is evaluated over the
 * LOCAL patch list patch_inds[0 .. n) of global indices into the (strided) patch grid, in list order, repeats allowed --
 * per-patch values and weights are those of the whole grid, the sequential f32 sum, the mean's divisor (n) and the
 * weighted divisor (sequential f32 sum of the listed weights) follow xregImgSimMetric2DPatchNCCCPU.cpp:97-101, 204,
 * 262-285.  n == 0 restores the whole grid.  Patch kinds only; may be changed between computes (the adapter of a
 * random-patch metric calls it with the reference's own draw before every compute).  The host picks the indices. */
int xrc_sm_set_patch_subset(xrc_sm* sm, const uint64_t* patch_inds, uint64_t n);

/* set_mov_imgs_buf_from_ray_caster (:119): zero-copy device hand-off. Re-callable
 * with a new offset; a different ray caster than the first is an error
 * (xregImgSimMetric2DCPU.cpp:45-70). */
int xrc_sm_bind_ray_caster(xrc_sm* sm, xrc_rc* rc, uint32_t proj_offset);
/* set_mov_imgs_host_buf (:122): images are copied H2D at every compute(). */
int xrc_sm_bind_host(xrc_sm* sm, const float* host_buf, uint32_t proj_offset);
/* moving images already on the device (e.g. a torch tensor) */
int xrc_sm_bind_device(xrc_sm* sm, const float* dev_buf, uint32_t proj_offset);

/* set_num_moving_images + allocate_resources (:81,93) */
int xrc_sm_allocate(xrc_sm* sm, uint32_t max_imgs);
int xrc_sm_set_num_imgs(xrc_sm* sm, uint32_t n);
/* compute() (:88): asynchronous on the context stream */
int xrc_sm_compute(xrc_sm* sm);
/* sim_vals() (:101-112): D2H of n floats; synchronises */
int xrc_sm_read_sims(xrc_sm* sm, float* host_dst, uint32_t n);
/* device pointer of the per-image similarity values (for an on-device gather) */
int xrc_sm_device_sims(xrc_sm* sm, float** dev_ptr);
/* gradient images of moving image `img` (debug / parity of the Sobel stage).
 * Valid after compute() for the GRAD / PATCH_GRAD kinds. */
int xrc_sm_read_grads(xrc_sm* sm, uint32_t img, float* host_gx, float* host_gy);

/* ---- fused batch evaluation: what Intensity2D3DRegi::obj_fn does per iteration
 * (lib/regi/interfaces_2d_3d/xregIntensity2D3DRegi.cpp:571-696): one
 * RayCaster::compute(vol_idx) followed by every view's ImgSimMetric2D::compute()
 * and one gather of views x pop floats.  sims_out[v * n_per_view + p].
 * Metrics must be bound to rc.  Synchronises once. */
int xrc_eval_batch(xrc_rc* rc, uint32_t vol_idx, xrc_sm* const* sms, uint32_t n_views,
                   uint32_t n_per_view, float* sims_out);
/* same without the final read-back / synchronise (results via xrc_sm_device_sims) */
int xrc_eval_batch_async(xrc_rc* rc, uint32_t vol_idx, xrc_sm* const* sms, uint32_t n_views);


/* ---- the whole objective in one call (caller-side glue): what
 * Intensity2D3DRegi::obj_fn(frame_xforms, ...) does for one moving volume
 * (xregIntensity2D3DRegi.cpp:571-696): size the ray caster / metrics for the population
 * (set_num_projs, set_num_moving_images, view-major offsets, :63-94), replicate the n_poses
 * poses over the views camera-major (RayCaster::distribute_xforms_among_cam_models,
 * xregRayCastInterface.cpp:97-114), ray cast, evaluate every view's metric, gather, and average
 * over views (ImgSimMetric2DCombineMean, xregImgSimMetric2DCombine.cpp:67-86).
 * cam_to_phys: n_poses x 12 host floats.  sims_out: n_poses.  per_view_out: optional
 * n_views x n_poses.  n_views must equal the number of camera models.  One synchronisation. */
int xrc_obj_fn(xrc_rc* rc, uint32_t vol_idx, xrc_sm* const* sms, uint32_t n_views, uint32_t n_poses,
               const float* cam_to_phys, float* sims_out, float* per_view_out);
/* The objective for several moving objects (Intensity2D3DRegi::obj_fn's loop over volumes,
 * xregIntensity2D3DRegi.cpp:594-629): object j (volume vol_idx[j], poses cam_to_phys[j * n_poses .. (j + 1) * n_poses))
 * is ray cast into the same n_views x n_poses projections -- the first object with the REPLACE store method (on top of
 * the per-camera background projections when use_bg_projs != 0, the reference's has_a_static_vol_ case), the others
 * with ACCUM -- then every view's metric, gather, mean over views.  The ray caster's own store method / background
 * settings are untouched on return. */
int xrc_obj_fn_objects(xrc_rc* rc, uint32_t n_objs, const uint32_t* vol_idx, xrc_sm* const* sms, uint32_t n_views,
                       uint32_t n_poses, const float* cam_to_phys, int use_bg_projs, float* sims_out, float* per_view_out);
/* The same objective spread over several GPUs from ONE host thread (the reference's optimiser loops are single
 * threaded, SURVEY 8(b) "Threading"): device d owns rcs[d] and the n_views metrics
 * sms[d * n_views .. d * n_views + n_views), each configured exactly like the single-device objects (same volume,
 * cameras, fixed images, parameters; one xrc_ctx per device).  The camera-major list of n_views x n_poses projections
 * (unit u = v * n_poses + p, the reference's global projection index) is cut into contiguous balanced chunks that may
 * straddle views (the first n_units % n_dev devices take one unit more): with one view this splits the
 * population (100 poses on 8 devices: 13 13 13 13 12 12 12 12), with several views and few poses it puts the views
 * on different devices (three views, one pose: one view each).  Every device's work is enqueued before any is waited
 * for, and only the n_views x n_poses scalars come back.  Results equal xrc_obj_fn's bit for bit (a pose's value does
 * not depend on its batch).  Each rcs[d] must be allocated for ceil(n_views * n_poses / n_dev) projections and each
 * metric for min(n_poses, that many) images. */
int xrc_obj_fn_multi(uint32_t n_dev, xrc_rc* const* rcs, xrc_sm* const* sms, uint32_t vol_idx, uint32_t n_views,
                     uint32_t n_poses, const float* cam_to_phys, float* sims_out, float* per_view_out);
/* One device's part of a sharded objective, for callers that shard across PROCESSES (one rank per GPU: SURVEY 8(e)):
 * evaluates only the units [first_unit, first_unit + n_units) of the camera-major list u = view * n_poses + pose
 * (Intensity2D3DRegi::setup's projection order, xregIntensity2D3DRegi.cpp:63-94) and writes their per-view similarity
 * values to unit_sims_out[0 .. n_units).  cam_to_phys holds all n_poses poses.  The ranks exchange these scalars
 * (all-gather) and average over views themselves (ImgSimMetric2DCombineMean).  Values equal xrc_obj_fn's per-view
 * values bit for bit.  rc must be allocated for n_units projections, each metric for the poses of its view in range. */
int xrc_obj_fn_units(xrc_rc* rc, uint32_t vol_idx, xrc_sm* const* sms, uint32_t n_views, uint32_t n_poses,
                     const float* cam_to_phys, uint32_t first_unit, uint32_t n_units, float* unit_sims_out);
/* xrc_obj_fn_units without the final synchronise / read-back: hands over the poses and enqueues the ray cast and the
 * metrics of the units in range on the context stream.  View v's values land in the first entries of its metric's
 * device result vector (xrc_sm_device_sims) and host-mapped copy, for callers that gather on the device (NCCL
 * all-gather straight from that vector) and synchronise once after their collective. */
int xrc_obj_fn_units_enqueue(xrc_rc* rc, uint32_t vol_idx, xrc_sm* const* sms, uint32_t n_views, uint32_t n_poses,
                             const float* cam_to_phys, uint32_t first_unit, uint32_t n_units);
/* ---- Tile-sharded objective for one process per GPU.  Sharding the POSES gives every
 * GPU pop / N projections to ray cast, and a tile's beams are then shared in L2 by pop / N poses only; sharding the
 * detector TILES instead keeps all pop poses of a tile on one GPU (the single-GPU access pattern) and balances the ray
 * casting by measured work.  Every rank ray casts its tiles of ALL n_views x n_poses projections and stores each
 * projection straight into the buffer of the rank that owns it -- the camera-major (view, pose) list cut into contiguous
 * balanced chunks exactly like xrc_obj_fn_units -- at the projection's global index, through a peer-mapped address
 * (NVLink stores issued by the ray-casting kernel itself; no staging copy, no all-to-all).  After a barrier across the
 * ranks, ordered on their streams, every rank scores the projections it owns and the ranks all-gather the scalars.
 * Results are bitwise those of one GPU (a pixel does not depend on which CTA, GPU or batch computed it).
 *   xrc_rc_peer_export   cudaIpcGetMemHandle of this ray caster's own projection buffer (after xrc_rc_allocate, which
 *                        must have room for ALL projections on every rank)
 *   xrc_rc_peer_attach   the 64-byte handles of all n_ranks <= 8 ranks (this rank's own entry is ignored); opens the
 *                        peers' buffers (cudaIpcOpenMemHandle, peer access enabled lazily).  Re-allocation detaches.
 *   xrc_rc_plan_tiles    which tiles each rank ray casts: contiguous ranges of the row-major tile list (neighbouring tiles
 *                        share their beams in L2), cut where the work measured by one count-only pass of the CURRENT
 *                        projections balances; exact integer counts, so every rank derives the same plan on its own.
 *                        Done implicitly by the first xrc_rc_compute_tiles after an attach; call it again when the pose
 *                        distribution moves (a new registration level).  xrc_rc_tile_plan reads the n_ranks + 1 bounds.
 *   xrc_rc_plan_tiles_timed  feedback from the clock: rank_ms[r] = what rank r's ray-casting kernel took under the current
 *                        plan (the same n_ranks numbers on every rank, e.g. all-gathered CUDA-event times); the tiles of a
 *                        rank that took longer than its planned share get a larger cost multiplier and the ranges are cut
 *                        again (samples are only a proxy of a tile's cost).  Two or three rounds settle within ~1 %.
 *   xrc_rc_compute_tiles RayCaster::compute for this rank's tiles of all current projections, written to their owners
 *   xrc_rc_tile_samples  instrumentation: trilinear samples of this rank's tiles (algorithmic / fetched), for rooflines */
#define XRC_IPC_HANDLE_BYTES 64
int xrc_rc_peer_export(xrc_rc* rc, uint8_t handle[XRC_IPC_HANDLE_BYTES]);
int xrc_rc_peer_attach(xrc_rc* rc, uint32_t n_ranks, uint32_t rank, const uint8_t* handles);
int xrc_rc_peer_detach(xrc_rc* rc);
int xrc_rc_plan_tiles(xrc_rc* rc, uint32_t vol_idx);
int xrc_rc_plan_tiles_timed(xrc_rc* rc, uint32_t vol_idx, const float* rank_ms);
int xrc_rc_tile_plan(const xrc_rc* rc, uint32_t* tile_begin);
int xrc_rc_compute_tiles(xrc_rc* rc, uint32_t vol_idx);
int xrc_rc_tile_samples(xrc_rc* rc, uint32_t vol_idx, uint64_t* algorithmic, uint64_t* fetched);
/* The two halves of the tile-sharded objective, both asynchronous on the context stream: (1) hand over all n_poses
 * poses (replicated over the views camera-major) and ray cast this rank's tiles; (2) -- after the caller's barrier --
 * every view's metric over the units [first_unit, first_unit + n_units) this rank owns, read at their global indices.
 * View v's values land in the first entries of its metric's result vector, as with xrc_obj_fn_units_enqueue. */
int xrc_obj_fn_tiles_enqueue_drr(xrc_rc* rc, uint32_t vol_idx, uint32_t n_views, uint32_t n_poses, const float* cam_to_phys);
int xrc_obj_fn_units_enqueue_metrics(xrc_rc* rc, xrc_sm* const* sms, uint32_t n_views, uint32_t n_poses, uint32_t first_unit,
                                     uint32_t n_units);
/* The same step without NCCL: the barrier and the all-gather of the scalars are done by one small kernel per rank over the
 * peer mappings (an exchange block at the tail of every rank's projection buffer: system-scope flags and NVLink stores).
 *   xrc_rc_peer_barrier              enqueue: signal every rank, wait for every rank (stream-ordered: what this stream did
 *                                    before -- the ray casting into the peers' buffers -- is visible to a rank that passes)
 *   xrc_obj_fn_tiles_enqueue_gather  enqueue: every view's metric over the units this rank owns (rank r owns chunk r of the
 *                                    camera-major list cut into n_ranks contiguous balanced chunks), their values stored into
 *                                    EVERY rank's gathered vector, barrier, all n_views x n_poses values to host-mapped memory
 *   xrc_obj_fn_tiles_finish          synchronise; per_view_out[v * n_poses + p] (optional) and the mean over the views
 *   xrc_obj_fn_tiles                 the whole evaluation in one call: enqueue_drr, barrier, enqueue_gather, finish.  Every
 *                                    rank calls it with the same poses and gets the same values, bitwise those of one GPU.
 * All ranks must make the same sequence of these calls (the barriers count epochs), with buffers allocated for the same
 * detector and max_projs.  A rank that does not arrive within 30 s makes the others return XRC_ERR_CUDA instead of hanging. */
int xrc_rc_peer_barrier(xrc_rc* rc);
int xrc_obj_fn_tiles_enqueue_gather(xrc_rc* rc, xrc_sm* const* sms, uint32_t n_views, uint32_t n_poses);
int xrc_obj_fn_tiles_finish(xrc_rc* rc, uint32_t n_views, uint32_t n_poses, float* sims_out, float* per_view_out);
int xrc_obj_fn_tiles(xrc_rc* rc, uint32_t vol_idx, xrc_sm* const* sms, uint32_t n_views, uint32_t n_poses,
                     const float* cam_to_phys, float* sims_out, float* per_view_out);
/* The partition xrc_obj_fn_multi uses (host only, needs no device): of view `view`, device `dev` evaluates the poses
 * [*first_pose, *first_pose + *count).  For sizing the per-device objects and for callers that shard by themselves. */
int xrc_obj_fn_multi_share(uint32_t n_dev, uint32_t n_views, uint32_t n_poses, uint32_t dev, uint32_t view,
                           uint32_t* first_pose, uint32_t* count);
/* Same from optimiser variables: pose_p = pre * ExpSE3(params_p) * post with
 * SE3OptVarsLieAlg (lib/regi/xregSE3OptVars.cpp:128-137; params = [w_x w_y w_z v_x v_y v_z]) and the
 * intermediate-frame composition of apply_inter_transforms_for_obj_fn (xregIntensity2D3DRegi.cpp:1049-1071).
 * pre12 / post12: row-major 3x4, NULL = identity.  All f32, evaluated on the host. */
int xrc_obj_fn_se3(xrc_rc* rc, uint32_t vol_idx, xrc_sm* const* sms, uint32_t n_views, uint32_t n_poses,
                   const float* params, const float* pre12, const float* post12, float* sims_out,
                   float* per_view_out);
/* The regulariser of Intensity2D3DRegi::obj_fn (xregIntensity2D3DRegi.cpp:653-688) for the penalty the reference's
 * multi-object apps use, Regi2D3DPenaltyFnSE3Mag with FoldNormDist densities
 * (lib/regi/penalty_fns_2d_3d/xregRegi2D3DPenaltyFnSE3Mag.cpp:32-117, lib/basic_math/xregFoldNormDist.cpp; e.g.
 * apps/hip_surgery/pao/frag_multi_view_regi_2d_3d/...main.cpp:303-309): per pose, the rotation angle and translation
 * magnitude of  inter^-1 * init * cur * inter  (ComputeRotAngTransMag, xregRigidUtils.cpp:247-251) against folded
 * normal densities, reg = (log Z_rot - log p_rot) + (log Z_trans - log p_trans).  All f32, host only. */
typedef struct xrc_se3_penalty
{
  float rot_mean, rot_std;       /* FoldNormDist(m, s) of the rotation angle, radians */
  float trans_mean, trans_std;   /* FoldNormDist(m, s) of the translation magnitude, mm */
  int32_t use_coeffs;            /* set_img_sim_penalty_coefs was called: sim * img_sim_coeff + reg * penalty_coeff */
  float img_sim_coeff, penalty_coeff;
  int32_t inter_wrt_vol;         /* intermediate_frames_wrt_vol[obj] */
  float inter_frame[12];         /* intermediate_frames[obj], row-major 3x4 */
  float init_cam_to_vol[12];     /* regi_xform_guesses[obj], row-major 3x4 */
} xrc_se3_penalty;
/* reg_vals_out[p] for the n poses cam_wrt_obj (n x 12, the transforms handed to the ray caster).  Needs no device. */
int xrc_se3_mag_penalty(const xrc_se3_penalty* pen, uint32_t n, const float* cam_wrt_obj, float* reg_vals_out);
/* xrc_obj_fn_se3 plus the regulariser, as one call: sims_out[p] = sim[p] (* img_sim_coeff) + reg[p] (* penalty_coeff).
 * The host evaluates the regulariser while the device ray casts.  penalty_out (optional): the unscaled reg values.
 * pen == NULL: plain xrc_obj_fn_se3. */
int xrc_obj_fn_se3_pen(xrc_rc* rc, uint32_t vol_idx, xrc_sm* const* sms, uint32_t n_views, uint32_t n_poses,
                       const float* params, const float* pre12, const float* post12, const xrc_se3_penalty* pen,
                       float* sims_out, float* per_view_out, float* penalty_out);
/* ExpSE3(Pt6) (lib/transforms/xregRigidUtils.cpp:40-85) in f32; host only, needs no device. */
void xrc_exp_se3(const float params[6], float out12[12]);

#ifdef __cplusplus
}
#endif
#endif
