// Compile check (tests/test_adapters_compile.py): every adapter class is concrete (all pure virtuals of the reference's
// RayCaster / ImgSimMetric2D are overridden with matching signatures) and is reachable through the reference's base
// classes and parameter mix-ins the way the backend factories and the apps use them
// (lib/ray_cast/xregRayCastProgOpts.cpp:47-70, lib/regi/sim_metrics_2d/xregImgSimMetric2DProgOpts.cpp:52-94,
// apps/hip_surgery/pelvis_single_view_regi_2d_3d/...main.cpp:259-270).
#include <memory>

#include "xregImgSimMetric2DCUDA.h"
#include "xregRayCastLineIntCUDA.h"

namespace
{

std::shared_ptr<xreg::RayCaster> MakeLineIntRayCaster(xrc_ctx* ctx)
{
  auto rc = std::make_shared<xreg::RayCasterLineIntCUDA>(ctx);
  rc->set_kernel_id(xreg::kRAY_CAST_LINE_INT_SUM_KERNEL);  // RayCastLineIntParamInterface
  return rc;
}

// the depth ray caster behind the reference's base class and its collision-parameter mix-in
// (RayCasterDepthCPU : RayCasterCPU, RayCasterCollisionParamInterface)
std::shared_ptr<xreg::RayCaster> MakeDepthRayCaster(xrc_ctx* ctx)
{
  auto rc = std::make_shared<xreg::RayCasterDepthCUDA>(ctx);
  rc->set_render_thresh(200);
  rc->set_num_backtracking_steps(8);
  if (auto* coll = dynamic_cast<xreg::RayCasterCollisionParamInterface*>(rc.get()))
  {
    coll->set_render_thresh(coll->render_thresh() + 1);
  }
  return rc;
}

std::shared_ptr<xreg::ImgSimMetric2D> MakeSimMetric(xrc_ctx* ctx, const int which)
{
  switch (which)
  {
    case 0: return std::make_shared<xreg::ImgSimMetric2DNCCCUDA>(ctx);
    case 1: return std::make_shared<xreg::ImgSimMetric2DGradNCCCUDA>(ctx);
    case 2: return std::make_shared<xreg::ImgSimMetric2DPatchNCCCUDA>(ctx);
    default: return std::make_shared<xreg::ImgSimMetric2DPatchGradNCCCUDA>(ctx);
  }
}

}  // namespace

float UseThroughTheReferenceInterfaces(xrc_ctx* ctx, xreg::RayCaster::VolPtr vol, const xreg::CameraModel& cam,
                                       xreg::ImgSimMetric2D::ImagePtr fixed, xreg::ImgSimMetric2D::ImageMaskPtr mask,
                                       const xreg::FrameTransformList& poses)
{
  {
    auto depth = MakeDepthRayCaster(ctx);
    depth->set_volume(vol);
    depth->set_camera_model(cam);
    depth->set_num_projs(poses.size());
    depth->allocate_resources();
    depth->set_xforms_cam_to_itk_phys(poses);
    depth->compute();
  }
  auto rc = MakeLineIntRayCaster(ctx);
  rc->set_volume(vol);
  rc->set_camera_model(cam);
  rc->set_num_projs(poses.size());
  rc->allocate_resources();

  auto sm = MakeSimMetric(ctx, 3);
  if (auto* patch = dynamic_cast<xreg::ImgSimMetric2DPatchCommon*>(sm.get()))
  {
    patch->set_patch_radius(13);
    patch->set_patch_stride(1);
  }
  if (auto* grad = dynamic_cast<xreg::ImgSimMetric2DGradImgParamInterface*>(sm.get()))
  {
    grad->set_smooth_img_before_sobel_kernel_radius(5);
  }
  sm->set_num_moving_images(poses.size());
  sm->set_fixed_image(fixed);
  sm->set_mov_imgs_buf_from_ray_caster(rc.get());
  sm->set_mask(mask);
  sm->allocate_resources();

  rc->distribute_xforms_among_cam_models(poses);
  rc->use_proj_store_replace_method();
  rc->compute();
  sm->compute();

  xreg::RayCaster::ProjPtr p = rc->proj(0);
  cv::Mat m = rc->proj_ocv(0);
  (void)p;
  (void)m;
  return sm->sim_val(0) + rc->raw_host_pixel_buf()[0];
}
