/*
 * ImgSimMetric2D*CUDA -- drop-in replacements of ImgSimMetric2D{NCC,GradNCC,PatchNCC,PatchGradNCC}OCL
 * (lib/regi/sim_metrics_2d/) that forward to libxreg_cuda.so (include/xreg_cuda.h).
 * Source-only here; compiles inside an xReg checkout.
 *
 * The patch classes inherit ImgSimMetric2DPatchCommon and the gradient classes
 * ImgSimMetric2DGradImgParamInterface because the apps dynamic_cast to those mix-ins
 * (apps/hip_surgery/pelvis_single_view_regi_2d_3d/...main.cpp:259-270).  The patch grid /
 * weight logic (ImgSimMetric2DPatchCommon::setup_patches / compute_weights) keeps running on
 * the host exactly as in the reference; only the resulting per-patch weights cross the ABI.
 */
#ifndef XREGIMGSIMMETRIC2DCUDA_H_
#define XREGIMGSIMMETRIC2DCUDA_H_

#include "xregImgSimMetric2D.h"
#include "xregImgSimMetric2DGradImgParamInterface.h"
#include "xregImgSimMetric2DPatchCommon.h"
#include "xregRayCastSyncBuf.h"

#include "xreg_cuda.h"

namespace xreg
{

class ImgSimMetric2DCUDA : public ImgSimMetric2D
{
public:
  ImgSimMetric2DCUDA(xrc_ctx* ctx, const int kind);

  ~ImgSimMetric2DCUDA() override;

  void allocate_resources() override;

  void compute() override;

  /// Zero-copy when the ray caster is a RayCasterLineIntCUDA on the same context; any other
  /// ray caster is read through its host sync buffer (to_host_buf()) and uploaded per compute().
  void set_mov_imgs_buf_from_ray_caster(RayCaster* ray_caster, const size_type proj_offset = 0) override;

  void set_mov_imgs_host_buf(Scalar* mov_imgs_buf, const size_type proj_offset = 0) override;

  xrc_sm* handle() { return sm_; }

protected:
  void process_mask() override;

  /// hook for the patch / gradient parameters, called before allocation and on mask updates
  virtual void push_params() { }

  /// hook run at the start of every compute() (the patch metrics choose their patch list here)
  virtual void pre_compute_params() { }

  xrc_ctx* ctx_ = nullptr;
  xrc_sm* sm_ = nullptr;

  RayCastSyncHostBuf* sync_host_buf_ = nullptr;  // non-CUDA ray caster
  size_type proj_off_ = 0;
  bool sm_allocated_ = false;
};

class ImgSimMetric2DNCCCUDA : public ImgSimMetric2DCUDA
{
public:
  explicit ImgSimMetric2DNCCCUDA(xrc_ctx* ctx) : ImgSimMetric2DCUDA(ctx, XRC_SM_NCC) { }
};

/// Replaces ImgSimMetric2DSSDOCL (SSDSimMetricFromProgOpts, xregImgSimMetric2DProgOpts.cpp)
class ImgSimMetric2DSSDCUDA : public ImgSimMetric2DCUDA
{
public:
  explicit ImgSimMetric2DSSDCUDA(xrc_ctx* ctx) : ImgSimMetric2DCUDA(ctx, XRC_SM_SSD) { }
};

class ImgSimMetric2DGradNCCCUDA : public ImgSimMetric2DCUDA, public ImgSimMetric2DGradImgParamInterface
{
public:
  explicit ImgSimMetric2DGradNCCCUDA(xrc_ctx* ctx) : ImgSimMetric2DCUDA(ctx, XRC_SM_GRAD_NCC) { }

  size_type smooth_img_before_sobel_kernel_radius() const override { return smooth_img_kernel_rad_; }

  void set_smooth_img_before_sobel_kernel_radius(const size_type r) override;

private:
  size_type smooth_img_kernel_rad_ = 5;
};

class ImgSimMetric2DPatchNCCCUDA : public ImgSimMetric2DCUDA, public ImgSimMetric2DPatchCommon
{
public:
  explicit ImgSimMetric2DPatchNCCCUDA(xrc_ctx* ctx) : ImgSimMetric2DCUDA(ctx, XRC_SM_PATCH_NCC) { }

  void allocate_resources() override;

protected:
  void push_params() override;

  void pre_compute_params() override;
};

class ImgSimMetric2DPatchGradNCCCUDA : public ImgSimMetric2DCUDA,
                                       public ImgSimMetric2DPatchCommon,
                                       public ImgSimMetric2DGradImgParamInterface
{
public:
  explicit ImgSimMetric2DPatchGradNCCCUDA(xrc_ctx* ctx) : ImgSimMetric2DCUDA(ctx, XRC_SM_PATCH_GRAD_NCC) { }

  void allocate_resources() override;

  size_type smooth_img_before_sobel_kernel_radius() const override { return smooth_img_kernel_rad_; }

  void set_smooth_img_before_sobel_kernel_radius(const size_type r) override;

protected:
  void push_params() override;

  void pre_compute_params() override;

private:
  size_type smooth_img_kernel_rad_ = 5;
};

}  // namespace xreg

#endif
