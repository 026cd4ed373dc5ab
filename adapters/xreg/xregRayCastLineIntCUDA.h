/*
 * RayCasterLineIntCUDA -- drop-in replacement of RayCasterLineIntOCL
 * (lib/ray_cast/xregRayCastLineIntOCL.{h,cpp}) that forwards to libxreg_cuda.so
 * (include/xreg_cuda.h).  Source-only in this repository: it compiles inside an xReg
 * checkout (needs the reference's ITK / Eigen / OpenCV based headers), see INTEGRATION.md.
 *
 * The class keeps no numerical code: RayCaster's protected state (vols_, camera_models_,
 * xforms_cam_to_itk_phys_, cam_model_for_proj_, ... xregRayCastInterface.h:360-417) is
 * flattened into the POD arguments of the C ABI right before compute().
 */
#ifndef XREGRAYCASTLINEINTCUDA_H_
#define XREGRAYCASTLINEINTCUDA_H_

#include <vector>

#include "xregRayCastInterface.h"
#include "xregRayCastSyncBuf.h"

#include "xreg_cuda.h"

namespace xreg
{

/// Lazy device -> host hand-off of the projection buffer: the CUDA analogue of
/// RayCastSyncHostBufFromOCL (lib/ray_cast/xregRayCastSyncBuf.cpp:60-110).  A CPU
/// similarity metric (or proj()) pulls DRRs through this object; a CUDA metric never does.
class RayCastSyncHostBufFromCUDA : public RayCastSyncHostBuf
{
public:
  void set_ray_caster(xrc_rc* rc, const size_type num_pix_per_proj);

  void sync() override;   ///< D2H of [range_start_, range_end_) when modified_
  void alloc() override;  ///< sizes the host vector to the ray caster's capacity

  HostBuf& host_buf() override;

  void set_external_host_buf(BufElem* buf);

private:
  xrc_rc* rc_ = nullptr;
  size_type num_pix_per_proj_ = 0;
  HostVec host_vec_;
  BufElem* ext_buf_ = nullptr;
  HostBuf host_buf_;
};

class RayCasterLineIntCUDA : public RayCaster, public RayCastLineIntParamInterface
{
public:
  /// \param ctx  context (device + stream) shared with the similarity metrics, the
  ///             analogue of the (boost::compute::context, command_queue) pair that
  ///             RayCasterOCL takes (xregRayCastBaseOCL.h:60-75)
  explicit RayCasterLineIntCUDA(xrc_ctx* ctx);

  ~RayCasterLineIntCUDA() override;

  void set_num_projs(const size_type num_projs) override;

  void allocate_resources() override;

  void compute(const size_type vol_idx = 0) override;

  ProjPtr proj(const size_type proj_idx) override;

  cv::Mat proj_ocv(const size_type proj_idx) override;

  PixelScalar2D* raw_host_pixel_buf() override;

  void use_external_host_pixel_buf(void* buf) override;

  size_type max_num_projs_possible() const override;

  void use_other_proj_buf(RayCaster* other_ray_caster) override;

  RayCastSyncHostBuf* to_host_buf() override;

  /// Handle for the CUDA metrics (zero-copy device hand-off, replaces to_ocl_buf())
  xrc_rc* handle() { return rc_; }

  xrc_ctx* ctx() { return ctx_; }

protected:
  void vols_changed() override;

  void camera_models_changed() override;

protected:
  /// flattens the RayCaster state into the library (parameters, poses, background flag): what compute() does first
  void push_params_and_poses();

  RayCastSyncHostBufFromCUDA& sync_to_host_buf() { return sync_to_host_; }

private:
  xrc_ctx* ctx_ = nullptr;
  xrc_rc* rc_ = nullptr;

  RayCastSyncHostBufFromCUDA sync_to_host_;

  std::vector<float> tmp_poses_;      // num_projs x 12, row-major
  std::vector<uint32_t> tmp_cam_idx_;
};

/// RayCasterDepthCUDA -- the CUDA counterpart of RayCasterDepthCPU (lib/ray_cast/xregRayCastDepthCPU.{h,cpp}): the depth
/// of the first sample along every ray whose interpolated value reaches render_thresh(), refined by
/// num_backtracking_steps() halvings of the step, min-combined with the background (kRAY_CAST_MAX_DEPTH by default).
/// Volumes, cameras, poses, store methods and the host hand-off are the line-integral adapter's.
class RayCasterDepthCUDA : public RayCasterLineIntCUDA, public RayCasterCollisionParamInterface
{
public:
  explicit RayCasterDepthCUDA(xrc_ctx* ctx);

  void compute(const size_type vol_idx = 0) override;
};

}  // namespace xreg

#endif
