#include "xregImgSimMetric2DCUDA.h"

#include <vector>

#include "xregAssert.h"
#include "xregExceptionUtils.h"
#include "xregITKOpenCVUtils.h"
#include "xregRayCastLineIntCUDA.h"

namespace
{

using namespace xreg;

void CheckXRC(const int status)
{
  if (status == XRC_OK)
  {
    return;
  }
  if (status == XRC_ERR_UNSUPPORTED)
  {
    throw ImgSimMetric2D::UnsupportedOperationException();
  }
  xregThrow("xreg_cuda: %s", xrc_last_error());
}

/// Flattens ImgSimMetric2DPatchCommon's state into xrc_sm_set_patch_params.  compute_weights()
/// (xregImgSimMetric2DPatchCommon.cpp:309-410) has already filled patch_infos[k].weight.
template <class tPatchSim>
void PushPatchParams(tPatchSim* self, xrc_sm* sm, const bool weights_are_trivial)
{
  std::vector<float> w;
  if (!weights_are_trivial)
  {
    w.reserve(self->patch_infos().size());
    for (const auto& p : self->patch_infos())
    {
      w.push_back(p.weight);
    }
  }
  CheckXRC(xrc_sm_set_patch_params(sm, static_cast<uint32_t>(self->patch_radius()),
                                   static_cast<uint32_t>(self->patch_stride()),
                                   self->compute_mean_of_patch_sims() ? 1 : 0,
                                   self->weight_patch_sims_in_combine() ? 1 : 0,
                                   self->use_mask_for_patch_stats() ? 1 : 0,
                                   w.empty() ? nullptr : w.data(), w.size()));
}

/// The local patch list of this compute() call, exactly as ImgSimMetric2DPatchNCCCPU::compute chooses it
/// (xregImgSimMetric2DPatchNCCCPU.cpp:97-101): a fresh patch_indices_to_use() draw unless set_patches_to_use() pinned
/// the list; the whole grid in natural order needs no list.
template <class tPatchSim>
void PushPatchSubset(tPatchSim* self, xrc_sm* sm, const bool pinned, const std::vector<xreg::size_type>& pinned_inds,
                     std::vector<xreg::size_type>* drawn)
{
  const std::vector<xreg::size_type>* inds = nullptr;
  if (pinned)
  {
    inds = &pinned_inds;
  }
  else if (self->choose_rand_patches())
  {
    inds = drawn;
  }
  if (inds)
  {
    std::vector<uint64_t> tmp(inds->begin(), inds->end());
    CheckXRC(xrc_sm_set_patch_subset(sm, tmp.data(), tmp.size()));
  }
  else
  {
    CheckXRC(xrc_sm_set_patch_subset(sm, nullptr, 0));
  }
}

}  // namespace

xreg::ImgSimMetric2DCUDA::ImgSimMetric2DCUDA(xrc_ctx* ctx, const int kind) : ctx_(ctx)
{
  CheckXRC(xrc_sm_create(ctx_, kind, &sm_));
}

xreg::ImgSimMetric2DCUDA::~ImgSimMetric2DCUDA()
{
  xrc_sm_destroy(sm_);
}

void xreg::ImgSimMetric2DCUDA::set_mov_imgs_buf_from_ray_caster(RayCaster* ray_caster, const size_type proj_offset)
{
  proj_off_ = proj_offset;
  if (auto* cuda_rc = dynamic_cast<RayCasterLineIntCUDA*>(ray_caster))
  {
    // same-device hand-off, the analogue of RayCastSyncOCLBufFromOCL (xregRayCastSyncBuf.cpp:155-159)
    sync_host_buf_ = nullptr;
    CheckXRC(xrc_sm_bind_ray_caster(sm_, cuda_rc->handle(), static_cast<uint32_t>(proj_offset)));
  }
  else
  {
    sync_host_buf_ = ray_caster->to_host_buf();
    if (sm_allocated_)
    {
      CheckXRC(xrc_sm_bind_host(sm_, sync_host_buf_->host_buf().buf, static_cast<uint32_t>(proj_offset)));
    }
  }
}

void xreg::ImgSimMetric2DCUDA::set_mov_imgs_host_buf(Scalar* mov_imgs_buf, const size_type proj_offset)
{
  xregASSERT(!sync_host_buf_);
  proj_off_ = proj_offset;
  CheckXRC(xrc_sm_bind_host(sm_, mov_imgs_buf, static_cast<uint32_t>(proj_offset)));
}

void xreg::ImgSimMetric2DCUDA::allocate_resources()
{
  ImgSimMetric2D::allocate_resources();

  const auto sz = this->fixed_img_->GetLargestPossibleRegion().GetSize();
  CheckXRC(xrc_sm_set_fixed(sm_, this->fixed_img_->GetBufferPointer(), static_cast<uint32_t>(sz[1]),
                            static_cast<uint32_t>(sz[0])));
  if (sync_host_buf_)
  {
    sync_host_buf_->alloc();
    CheckXRC(xrc_sm_bind_host(sm_, sync_host_buf_->host_buf().buf, static_cast<uint32_t>(proj_off_)));
  }
  this->process_updated_mask();
  push_params();
  CheckXRC(xrc_sm_allocate(sm_, static_cast<uint32_t>(this->num_mov_imgs_)));
  sm_allocated_ = true;
}

void xreg::ImgSimMetric2DCUDA::process_mask()
{
  CheckXRC(xrc_sm_set_mask(sm_, this->mask_ ? this->mask_->GetBufferPointer() : nullptr));
  if (sm_allocated_)
  {
    push_params();
  }
}

void xreg::ImgSimMetric2DCUDA::compute()
{
  xregASSERT(sm_allocated_);
  if (sync_host_buf_)
  {
    sync_host_buf_->sync();
  }
  this->process_updated_mask();
  pre_compute_params();
  CheckXRC(xrc_sm_set_num_imgs(sm_, static_cast<uint32_t>(this->num_mov_imgs_)));
  CheckXRC(xrc_sm_compute(sm_));
  CheckXRC(xrc_sm_read_sims(sm_, this->sim_vals_.data(), static_cast<uint32_t>(this->num_mov_imgs_)));
}

void xreg::ImgSimMetric2DGradNCCCUDA::set_smooth_img_before_sobel_kernel_radius(const size_type r)
{
  smooth_img_kernel_rad_ = r;
  CheckXRC(xrc_sm_set_grad_params(sm_, static_cast<uint32_t>(r)));
}

void xreg::ImgSimMetric2DPatchNCCCUDA::allocate_resources()
{
  const auto sz = this->fixed_img_->GetLargestPossibleRegion().GetSize();
  cv::Mat ocv_mask;
  if (this->mask_)
  {
    ocv_mask = ShallowCopyItkToOpenCV(this->mask_.GetPointer());
  }
  this->setup_patches(sz[1], sz[0], this->mask_ ? &ocv_mask : nullptr, this->num_mov_imgs_);
  ImgSimMetric2DCUDA::allocate_resources();
}

void xreg::ImgSimMetric2DPatchNCCCUDA::push_params()
{
  cv::Mat ocv_mask;
  if (this->mask_)
  {
    ocv_mask = ShallowCopyItkToOpenCV(this->mask_.GetPointer());
  }
  this->need_to_recompute_weights_ = true;
  this->compute_weights(this->mask_ ? &ocv_mask : nullptr);
  const bool trivial = !this->wgt_img_ && !(this->use_mask_for_weighting_ && this->mask_);
  PushPatchParams(this, sm_, trivial);
}

void xreg::ImgSimMetric2DPatchNCCCUDA::pre_compute_params()
{
  if (!this->do_not_update_patch_inds_to_use_)
  {
    this->patch_inds_to_use_ = this->patch_indices_to_use();
  }
  PushPatchSubset(this, sm_, this->do_not_update_patch_inds_to_use_, this->patch_inds_to_use_, &this->patch_inds_to_use_);
}

void xreg::ImgSimMetric2DPatchGradNCCCUDA::pre_compute_params()
{
  // one list for both gradient directions (xregImgSimMetric2DPatchGradNCCCPU.cpp:126-135)
  if (!this->do_not_update_patch_inds_to_use_)
  {
    this->patch_inds_to_use_ = this->patch_indices_to_use();
  }
  PushPatchSubset(this, sm_, this->do_not_update_patch_inds_to_use_, this->patch_inds_to_use_, &this->patch_inds_to_use_);
}

void xreg::ImgSimMetric2DPatchGradNCCCUDA::allocate_resources()
{
  const auto sz = this->fixed_img_->GetLargestPossibleRegion().GetSize();
  cv::Mat ocv_mask;
  if (this->mask_)
  {
    ocv_mask = ShallowCopyItkToOpenCV(this->mask_.GetPointer());
  }
  this->setup_patches(sz[1], sz[0], this->mask_ ? &ocv_mask : nullptr, this->num_mov_imgs_);
  CheckXRC(xrc_sm_set_grad_params(sm_, static_cast<uint32_t>(smooth_img_kernel_rad_)));
  ImgSimMetric2DCUDA::allocate_resources();
}

void xreg::ImgSimMetric2DPatchGradNCCCUDA::set_smooth_img_before_sobel_kernel_radius(const size_type r)
{
  smooth_img_kernel_rad_ = r;
  CheckXRC(xrc_sm_set_grad_params(sm_, static_cast<uint32_t>(r)));
}

void xreg::ImgSimMetric2DPatchGradNCCCUDA::push_params()
{
  cv::Mat ocv_mask;
  if (this->mask_)
  {
    ocv_mask = ShallowCopyItkToOpenCV(this->mask_.GetPointer());
  }
  this->need_to_recompute_weights_ = true;
  this->compute_weights(this->mask_ ? &ocv_mask : nullptr);
  const bool trivial = !this->wgt_img_ && !(this->use_mask_for_weighting_ && this->mask_);
  PushPatchParams(this, sm_, trivial);
}
