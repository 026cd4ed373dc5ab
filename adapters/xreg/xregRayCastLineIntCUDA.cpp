#include "xregRayCastLineIntCUDA.h"

#include "xregAssert.h"
#include "xregExceptionUtils.h"
#include "xregITKBasicImageUtils.h"

namespace
{

using namespace xreg;

/// xrc_status -> the reference's exception types
/// (lib/common/xregExceptionUtils.h:36-63, xregRayCastInterface.h:59)
void CheckXRC(const int status)
{
  if (status == XRC_OK)
  {
    return;
  }
  if (status == XRC_ERR_UNSUPPORTED)
  {
    throw RayCaster::UnsupportedOperationException();
  }
  xregThrow("xreg_cuda: %s", xrc_last_error());
}

xrc_cam ToXrcCam(const CameraModel& cam)
{
  xrc_cam c;
  c.rows = static_cast<uint32_t>(cam.num_det_rows);
  c.cols = static_cast<uint32_t>(cam.num_det_cols);
  for (int r = 0; r < 3; ++r)
  {
    for (int col = 0; col < 3; ++col)
    {
      c.intrins_inv[3 * r + col] = cam.intrins_inv(r, col);
    }
    for (int col = 0; col < 4; ++col)
    {
      c.extrins_inv[4 * r + col] = cam.extrins_inv.matrix()(r, col);
    }
    c.pinhole[r] = cam.pinhole_pt(r);
  }
  c.focal_len = cam.focal_len;
  c.frame_type = static_cast<int32_t>(cam.coord_frame_type);
  return c;
}

}  // namespace

void xreg::RayCastSyncHostBufFromCUDA::set_ray_caster(xrc_rc* rc, const size_type num_pix_per_proj)
{
  rc_ = rc;
  num_pix_per_proj_ = num_pix_per_proj;
}

void xreg::RayCastSyncHostBufFromCUDA::alloc()
{
  uint32_t n = 0;
  CheckXRC(xrc_rc_num_projs(rc_, &n));
  if (!ext_buf_)
  {
    host_vec_.resize(static_cast<size_type>(n) * num_pix_per_proj_);
    host_buf_ = HostBuf(host_vec_.data(), host_vec_.size());
  }
  else
  {
    host_buf_ = HostBuf(ext_buf_, static_cast<size_type>(n) * num_pix_per_proj_);
  }
}

void xreg::RayCastSyncHostBufFromCUDA::sync()
{
  if (this->modified_)
  {
    uint32_t n = 0;
    CheckXRC(xrc_rc_num_projs(rc_, &n));
    // range_* are element offsets into the projection buffer (set by set_num_projs)
    const size_type first = this->range_start_ / num_pix_per_proj_;
    const size_type last = (this->range_end_ == kRANGE_AT_BUF_END) ? n : (this->range_end_ / num_pix_per_proj_);
    CheckXRC(xrc_rc_read_projs(rc_, static_cast<uint32_t>(first), static_cast<uint32_t>(last - first),
                               host_buf_.buf + (first * num_pix_per_proj_)));
    this->modified_ = false;
  }
}

xreg::RayCastSyncBuf::HostBuf& xreg::RayCastSyncHostBufFromCUDA::host_buf()
{
  return host_buf_;
}

void xreg::RayCastSyncHostBufFromCUDA::set_external_host_buf(BufElem* buf)
{
  ext_buf_ = buf;
}

xreg::RayCasterLineIntCUDA::RayCasterLineIntCUDA(xrc_ctx* ctx) : ctx_(ctx)
{
  CheckXRC(xrc_rc_create(ctx_, &rc_));
}

xreg::RayCasterLineIntCUDA::~RayCasterLineIntCUDA()
{
  xrc_rc_destroy(rc_);
}

void xreg::RayCasterLineIntCUDA::vols_changed()
{
  // RayCasterOCL::vols_changed (xregRayCastBaseOCL.cpp:440-470) aliases host memory with
  // use_host_ptr; here every volume is copied to (and repacked on) the device once.
  const size_type n = this->vols_.size();
  std::vector<const float*> ptrs(n);
  std::vector<uint64_t> dims(3 * n);
  std::vector<float> xf(12 * n);
  for (size_type i = 0; i < n; ++i)
  {
    auto* v = this->vols_[i].GetPointer();
    ptrs[i] = v->GetBufferPointer();
    const auto sz = v->GetLargestPossibleRegion().GetSize();
    dims[3 * i] = sz[0];
    dims[3 * i + 1] = sz[1];
    dims[3 * i + 2] = sz[2];
    const FrameTransform idx_to_phys = ITKImagePhysicalPointTransformsAsEigen(v);
    for (int r = 0; r < 3; ++r)
    {
      for (int c = 0; c < 4; ++c)
      {
        xf[12 * i + 4 * r + c] = idx_to_phys.matrix()(r, c);
      }
    }
  }
  CheckXRC(xrc_rc_set_volumes(rc_, static_cast<uint32_t>(n), ptrs.data(),
                              reinterpret_cast<const uint64_t(*)[3]>(dims.data()),
                              reinterpret_cast<const float(*)[12]>(xf.data())));
}

void xreg::RayCasterLineIntCUDA::camera_models_changed()
{
  std::vector<xrc_cam> cams;
  cams.reserve(this->camera_models_.size());
  for (const auto& c : this->camera_models_)
  {
    cams.push_back(ToXrcCam(c));
  }
  CheckXRC(xrc_rc_set_cameras(rc_, static_cast<uint32_t>(cams.size()), cams.data()));
}

void xreg::RayCasterLineIntCUDA::set_num_projs(const size_type num_projs)
{
  RayCaster::set_num_projs(num_projs);
  if (this->resources_allocated_)
  {
    CheckXRC(xrc_rc_set_num_projs(rc_, static_cast<uint32_t>(num_projs)));
    const size_type num_pix = this->camera_models_[0].num_det_rows * this->camera_models_[0].num_det_cols;
    sync_to_host_.set_range(0, num_projs * num_pix);  // as RayCasterCPU::set_num_projs
  }
}

void xreg::RayCasterLineIntCUDA::allocate_resources()
{
  RayCaster::allocate_resources();
  CheckXRC(xrc_rc_allocate(rc_, static_cast<uint32_t>(this->num_projs_)));
  const size_type num_pix = this->camera_models_[0].num_det_rows * this->camera_models_[0].num_det_cols;
  sync_to_host_.set_ray_caster(rc_, num_pix);
  sync_to_host_.alloc();
  this->resources_allocated_ = true;
}

void xreg::RayCasterLineIntCUDA::push_params_and_poses()
{
  CheckXRC(xrc_rc_set_params(rc_, this->ray_step_size_, static_cast<int>(this->interp_method_),
                             static_cast<int>(this->kernel_id()), static_cast<int>(this->proj_store_meth_),
                             this->default_bg_pixel_val_));
  if (this->use_bg_projs_ && this->bg_projs_updated_)
  {
    std::vector<const float*> bgs;
    for (auto& p : this->bg_projs_for_each_cam_)
    {
      bgs.push_back(p->GetBufferPointer());
    }
    CheckXRC(xrc_rc_set_bg_projs(rc_, bgs.data(), 1));
    this->bg_projs_updated_ = false;
  }
  else
  {
    // set_use_bg_projs() only flips the flag in the reference, and Intensity2D3DRegi::obj_fn toggles it around every
    // static-volume evaluation (true -> compute(vol 0) -> false -> compute(vol 1), xregIntensity2D3DRegi.cpp:594-629):
    // push the flag on every compute; NULL images re-enable the background already resident on the device
    CheckXRC(xrc_rc_set_bg_projs(rc_, nullptr, this->use_bg_projs_ ? 1 : 0));
  }

  const size_type n = this->num_projs_;
  tmp_poses_.resize(12 * n);
  tmp_cam_idx_.resize(n);
  for (size_type i = 0; i < n; ++i)
  {
    const auto& m = this->xforms_cam_to_itk_phys_[i].matrix();
    for (int r = 0; r < 3; ++r)
    {
      for (int c = 0; c < 4; ++c)
      {
        tmp_poses_[12 * i + 4 * r + c] = m(r, c);
      }
    }
    tmp_cam_idx_[i] = static_cast<uint32_t>(this->cam_model_for_proj_[i]);
  }
  CheckXRC(xrc_rc_set_poses(rc_, static_cast<uint32_t>(n), tmp_poses_.data(), tmp_cam_idx_.data()));
}

void xreg::RayCasterLineIntCUDA::compute(const size_type vol_idx)
{
  xregASSERT(this->resources_allocated_);
  push_params_and_poses();
  CheckXRC(xrc_rc_compute(rc_, static_cast<uint32_t>(vol_idx)));
  sync_to_host_.set_modified();
}

xreg::RayCasterDepthCUDA::RayCasterDepthCUDA(xrc_ctx* ctx) : RayCasterLineIntCUDA(ctx)
{
  this->default_bg_pixel_val_ = kRAY_CAST_MAX_DEPTH;  // as RayCasterDepthCPU::RayCasterDepthCPU (xregRayCastDepthCPU.cpp:231-234)
}

void xreg::RayCasterDepthCUDA::compute(const size_type vol_idx)
{
  xregASSERT(this->resources_allocated_);
  push_params_and_poses();
  CheckXRC(xrc_rc_compute_depth(handle(), static_cast<uint32_t>(vol_idx), this->render_thresh(),
                                static_cast<uint32_t>(this->num_backtracking_steps())));
  sync_to_host_buf().set_modified();
}

xreg::RayCaster::ProjPtr xreg::RayCasterLineIntCUDA::proj(const size_type proj_idx)
{
  sync_to_host_.sync();

  const auto& cam = this->camera_models_[this->cam_model_for_proj_[proj_idx]];
  const size_type num_dets = cam.num_det_rows * cam.num_det_cols;

  auto img = Proj::New();
  auto px = Proj::PixelContainer::New();
  px->SetImportPointer(sync_to_host_.host_buf().buf + (num_dets * proj_idx), num_dets, false);
  img->SetPixelContainer(px);

  Proj::RegionType region;
  region.SetIndex(0, 0);
  region.SetIndex(1, 0);
  region.SetSize(0, cam.num_det_cols);
  region.SetSize(1, cam.num_det_rows);
  img->SetRegions(region);

  const CoordScalar spacings[2] = {cam.det_col_spacing, cam.det_row_spacing};
  img->SetSpacing(spacings);
  return img;
}

cv::Mat xreg::RayCasterLineIntCUDA::proj_ocv(const size_type proj_idx)
{
  sync_to_host_.sync();
  const auto& cam = this->camera_models_[this->cam_model_for_proj_[proj_idx]];
  return cv::Mat(cam.num_det_rows, cam.num_det_cols, cv::DataType<PixelScalar2D>::type,
                 sync_to_host_.host_buf().buf + (cam.num_det_rows * cam.num_det_cols * proj_idx));
}

xreg::RayCaster::PixelScalar2D* xreg::RayCasterLineIntCUDA::raw_host_pixel_buf()
{
  sync_to_host_.sync();
  return sync_to_host_.host_buf().buf;
}

void xreg::RayCasterLineIntCUDA::use_external_host_pixel_buf(void* buf)
{
  sync_to_host_.set_external_host_buf(static_cast<PixelScalar2D*>(buf));
  if (this->resources_allocated_)
  {
    sync_to_host_.alloc();
    sync_to_host_.set_modified();
  }
}

xreg::size_type xreg::RayCasterLineIntCUDA::max_num_projs_possible() const
{
  uint64_t n = 0;
  CheckXRC(xrc_rc_max_projs_possible(rc_, &n));
  return static_cast<size_type>(n);
}

void xreg::RayCasterLineIntCUDA::use_other_proj_buf(RayCaster* other_ray_caster)
{
  auto* other = dynamic_cast<RayCasterLineIntCUDA*>(other_ray_caster);
  if (!other)
  {
    throw UnsupportedOperationException();
  }
  CheckXRC(xrc_rc_use_other_proj_buf(rc_, other->handle()));
}

xreg::RayCastSyncHostBuf* xreg::RayCasterLineIntCUDA::to_host_buf()
{
  return &sync_to_host_;
}
