#!/usr/bin/env python
"""TEST INFRASTRUCTURE.  Builds oracle/_ref/libxreg_refslice.so from the reference's OWN source lines, where they lie
under /root/reference, for the geometric core of the DRR path:

  lib/spatial/xregSpatialPrimitives.cpp     xreg::RayRectIntersect
  lib/transforms/xregPerspectiveXform.cpp   xreg::CameraModel::ind_pt_to_phys_det_pt (both overloads)
  lib/ray_cast/xregRayCastLineIntCPU.cpp    the un-named namespace: AccumLineIntKernel, MaxLineIntKernel,
                                            LineIntParams, ComputeLineInts<Kernel>

The reference as a whole cannot be compiled here (ITK, Eigen, OpenCV, TBB, Boost are absent: DESIGN.md section 1), but
these functions only need a handful of vector / matrix / interpolator types.  This script cuts the functions out of the
reference files BY ANCHOR (function signatures, not line numbers) into a generated translation unit under oracle/_ref/
(git-ignored: no reference source enters the repository), puts oracle/ref_pin/ref_pin_prelude.h (our functional
stand-ins for the Eigen / ITK / TBB types, with the stated arithmetic conventions) in front and a C ABI behind, and
compiles it with the oracle's flags (g++ -O2 -ffp-contract=off, no -march).  tests/test_oracle_ref_slice.py then
requires the oracle's restatement to agree with it bit for bit.

Run from build() in __graft_entry__.py when /root/reference exists; the .so travels to the GPU box with the snapshot.
"""
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(os.path.dirname(HERE), "_ref")
REF = os.environ.get("XREG_REFERENCE", "/root/reference")
LIB = os.path.join(OUT_DIR, "libxreg_refslice.so")


def _lines(rel):
    with open(os.path.join(REF, rel)) as f:
        return f.read().split("\n")


def _cut_function(lines, first_line_regex, n_closing=1):
    """Lines from the one matching first_line_regex through the n-th following line that is exactly '}'."""
    start = next(i for i, ln in enumerate(lines) if re.search(first_line_regex, ln))
    i, seen = start, 0
    while seen < n_closing:
        i += 1
        if lines[i].rstrip() == "}":
            seen += 1
    return start, i


def slices():
    out = []
    # RayRectIntersect: the return type sits on the line before the qualified name
    ln = _lines("lib/spatial/xregSpatialPrimitives.cpp")
    s, e = _cut_function(ln, r"^xreg::RayRectIntersect\(")
    assert ln[s - 1].startswith("std::tuple<bool,"), ln[s - 1]
    out.append(("lib/spatial/xregSpatialPrimitives.cpp", s - 1, e, ln[s - 1:e + 1]))
    # ind_pt_to_phys_det_pt(const Pt2&) and (const Pt3&): two consecutive definitions
    ln = _lines("lib/transforms/xregPerspectiveXform.cpp")
    s, e = _cut_function(ln, r"^xreg::Pt3 xreg::CameraModel::ind_pt_to_phys_det_pt\(const Pt2& ind_pt\) const", n_closing=2)
    assert any("ind_pt_to_phys_det_pt(const Pt3& ind_pt) const" in x for x in ln[s:e + 1])
    out.append(("lib/transforms/xregPerspectiveXform.cpp", s, e, ln[s:e + 1]))
    # the un-named namespace of the CPU line-integral ray caster
    ln = _lines("lib/ray_cast/xregRayCastLineIntCPU.cpp")
    k = next(i for i, x in enumerate(ln) if x.startswith("struct AccumLineIntKernel"))
    s = max(i for i in range(k) if ln[i].strip() == "namespace")
    e = next(i for i in range(k, len(ln)) if re.match(r"^\}\s*//\s*un-named", ln[i]))
    body = ln[s:e + 1]
    assert any("void ComputeLineInts(" in x for x in body) and any("struct LineIntParams" in x for x in body)
    out.append(("lib/ray_cast/xregRayCastLineIntCPU.cpp", s, e, body))
    # the un-named namespace of the CPU depth ray caster: RayCastDepthFn
    ln = _lines("lib/ray_cast/xregRayCastDepthCPU.cpp")
    k = next(i for i, x in enumerate(ln) if x.startswith("struct RayCastDepthFn"))
    s = max(i for i in range(k) if ln[i].strip().startswith("namespace"))
    e = next(i for i in range(k, len(ln)) if re.match(r"^\}\s*//\s*un-named", ln[i]))
    body = ln[s:e + 1]
    assert any("collision_thresh" in x for x in body) and any("num_backtracking_steps" in x for x in body)
    out.append(("lib/ray_cast/xregRayCastDepthCPU.cpp", s, e, body))
    # RayCaster::distribute_xforms_among_cam_models and RayCasterCPU::pre_compute
    ln = _lines("lib/ray_cast/xregRayCastInterface.cpp")
    s, e = _cut_function(ln, r"^void xreg::RayCaster::distribute_xforms_among_cam_models\(")
    out.append(("lib/ray_cast/xregRayCastInterface.cpp", s, e, ln[s:e + 1]))
    ln = _lines("lib/ray_cast/xregRayCastBaseCPU.cpp")
    s, e = _cut_function(ln, r"^void xreg::RayCasterCPU::pre_compute\(\)")
    out.append(("lib/ray_cast/xregRayCastBaseCPU.cpp", s, e, ln[s:e + 1]))
    return out


WRAPPER = r'''
// ---- C ABI over the reference's functions (ours) --------------------------------------------------------------------
struct xref_cam
{
  uint32_t rows, cols;
  float intrins_inv[9];   // row-major
  float extrins_inv[12];  // row-major 3x4
  float pinhole[3];
  float focal_len;
  int32_t frame_type;
};

static xreg::FrameTransform affine_from12(const float* a)
{
  xreg::FrameTransform t;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 4; ++j)
      t.m[i][j] = a[4 * i + j];
  return t;
}

static xreg::CameraModel cam_from(const xref_cam& c)
{
  xreg::CameraModel m;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      m.intrins_inv(i, j) = c.intrins_inv[3 * i + j];
  m.extrins_inv = affine_from12(c.extrins_inv);
  for (int i = 0; i < 3; ++i)
    m.pinhole_pt(i) = c.pinhole[i];
  m.focal_len = c.focal_len;
  m.num_det_rows = c.rows;
  m.num_det_cols = c.cols;
  m.coord_frame_type = static_cast<xreg::CameraModel::CameraCoordFrame>(c.frame_type);
  return m;
}

extern "C" int xref_ray_rect_intersect(const float mn[3], const float mx[3], const float p[3], const float d[3],
                                       int limit_to_segment, float* t_start, float* t_stop)
{
  xreg::Pt3 a, b, c, e;
  for (int i = 0; i < 3; ++i)
  {
    a(i) = mn[i];
    b(i) = mx[i];
    c(i) = p[i];
    e(i) = d[i];
  }
  bool hit = false;
  std::tie(hit, *t_start, *t_stop) = xreg::RayRectIntersect(a, b, c, e, limit_to_segment != 0);
  return hit ? 1 : 0;
}

extern "C" void xref_ind_pt_to_phys_det_pt(const xref_cam* cam, float col, float row, float out[3])
{
  const xreg::CameraModel m = cam_from(*cam);
  xreg::Pt2 ind;
  ind(0) = col;
  ind(1) = row;
  const xreg::Pt3 r = m.ind_pt_to_phys_det_pt(ind);
  for (int i = 0; i < 3; ++i)
    out[i] = r(i);
}

// distribute_xforms_among_cam_models: n_poses poses -> n_cams * n_poses (pose, camera index) pairs
extern "C" void xref_distribute_xforms(const float* poses, uint32_t n_poses, uint32_t n_cams, float* out_poses,
                                       uint32_t* out_cam_idx)
{
  xreg::RayCaster rc;
  rc.camera_models_.resize(n_cams);
  rc.num_projs_ = (std::size_t)n_poses * n_cams;
  rc.xforms_cam_to_itk_phys_.resize(rc.num_projs_);
  rc.cam_model_for_proj_.assign(rc.num_projs_, ~(std::size_t)0);
  xreg::FrameTransformList in;
  for (uint32_t p = 0; p < n_poses; ++p)
    in.push_back(affine_from12(poses + 12 * (std::size_t)p));
  rc.distribute_xforms_among_cam_models(in);
  for (std::size_t g = 0; g < rc.num_projs_; ++g)
  {
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 4; ++j)
        out_poses[12 * g + 4 * i + j] = rc.xforms_cam_to_itk_phys_[g].m[i][j];
    out_cam_idx[g] = (uint32_t)rc.cam_model_for_proj_[g];
  }
}

// RayCasterCPU::pre_compute on a caller-owned projection buffer
extern "C" void xref_pre_compute(float* buf, uint32_t n_projs, uint32_t rows, uint32_t cols, const uint32_t* cam_idx,
                                 float* const* bg_projs, uint32_t n_cams, int store_method, float default_bg)
{
  xreg::RayCasterCPU rc;
  rc.camera_models_.resize(n_cams);
  for (auto& c : rc.camera_models_)
  {
    c.num_det_rows = rows;
    c.num_det_cols = cols;
  }
  rc.num_projs_ = n_projs;
  rc.cam_model_for_proj_.assign(cam_idx, cam_idx + n_projs);
  rc.buf_ = buf;
  rc.proj_store_meth_ = static_cast<xreg::RayCaster::ProjPixelStoreMethod>(store_method);
  rc.default_bg_pixel_val_ = default_bg;
  std::vector<xreg::Proj2D> bgs(n_cams);
  if (bg_projs)
  {
    rc.use_bg_projs_ = true;
    for (uint32_t c = 0; c < n_cams; ++c)
    {
      bgs[c].buf = bg_projs[c];
      xreg::Proj2D::Pointer p;
      p.p = &bgs[c];
      rc.bg_projs_for_each_cam_.push_back(p);
    }
  }
  rc.pre_compute();
}

// RayCasterLineIntCPU::compute's call of ComputeLineInts (xregRayCastLineIntCPU.cpp, after pre_compute): the caller
// passes the already inverted physical-point -> index transform (Eigen's .inverse() is third-party arithmetic) and an
// initialised projection buffer (REPLACE: zeros / background, ACCUM: the previous content).
extern "C" int xref_compute_line_ints_interp(const float* vol, const uint64_t dims[3], const float phys_to_idx[12],
                                             const xref_cam* cams, uint32_t n_cams, const float* poses,
                                             const uint32_t* cam_idx, uint32_t n_projs, float step_size, int kernel_id,
                                             int interp, float* proj_buf);
extern "C" int xref_compute_line_ints(const float* vol, const uint64_t dims[3], const float phys_to_idx[12],
                                      const xref_cam* cams, uint32_t n_cams, const float* poses,
                                      const uint32_t* cam_idx, uint32_t n_projs, float step_size, int kernel_id,
                                      float* proj_buf)
{
  return xref_compute_line_ints_interp(vol, dims, phys_to_idx, cams, n_cams, poses, cam_idx, n_projs, step_size, kernel_id,
                                       (int)xreg::RayCaster::kRAY_CAST_INTERP_LINEAR, proj_buf);
}

// RayCasterDepthCPU::compute's call of RayCastDepthFn over the whole projection range (xregRayCastDepthCPU.cpp:254-268);
// proj_buf initialised by the caller (pre_compute: kRAY_CAST_MAX_DEPTH or the previous content)
extern "C" int xref_compute_depth(const float* vol, const uint64_t dims[3], const float phys_to_idx[12],
                                  const xref_cam* cams, uint32_t n_cams, const float* poses, const uint32_t* cam_idx,
                                  uint32_t n_projs, float step_size, int interp, float collision_thresh,
                                  uint32_t num_backtracking_steps, float* proj_buf)
{
  if (!n_cams || !n_projs)
    return 0;
  xreg::RayCaster::Vol img;
  img.data = vol;
  for (int i = 0; i < 3; ++i)
    img.size[i] = (std::size_t)dims[i];
  xreg::RayCaster::CameraModelList cam_list;
  for (uint32_t c = 0; c < n_cams; ++c)
    cam_list.push_back(cam_from(cams[c]));
  xreg::FrameTransformList xforms;
  xreg::RayCaster::CamModelAssocList assoc;
  for (uint32_t p = 0; p < n_projs; ++p)
  {
    xforms.push_back(affine_from12(poses + 12 * (std::size_t)p));
    assoc.push_back(cam_idx ? cam_idx[p] : 0);
  }
  xreg::Pt3 bb_min, bb_max;
  for (int i = 0; i < 3; ++i)
  {
    bb_min(i) = 0;
    bb_max(i) = static_cast<float>(dims[i] - 1);
  }
  RayCastDepthFn fn = {&img, bb_min, bb_max, affine_from12(phys_to_idx), n_projs, cam_list, xforms, assoc, step_size,
                       static_cast<xreg::RayCaster::InterpMethod>(interp), proj_buf, collision_thresh,
                       (xreg::size_type)num_backtracking_steps};
  fn(xreg::RangeType(0, (std::size_t)n_projs * cam_list[0].num_det_rows * cam_list[0].num_det_cols));
  return 0;
}

// interp: RayCaster::InterpMethod (0 linear, 1 nearest neighbour; the stand-ins for sinc / B-spline throw)
extern "C" int xref_compute_line_ints_interp(const float* vol, const uint64_t dims[3], const float phys_to_idx[12],
                                             const xref_cam* cams, uint32_t n_cams, const float* poses,
                                             const uint32_t* cam_idx, uint32_t n_projs, float step_size, int kernel_id,
                                             int interp, float* proj_buf)
{
  if (!n_cams || !n_projs)
    return 0;
  xreg::RayCaster::Vol img;
  img.data = vol;
  for (int i = 0; i < 3; ++i)
    img.size[i] = (std::size_t)dims[i];
  xreg::RayCaster::CameraModelList cam_list;
  for (uint32_t c = 0; c < n_cams; ++c)
    cam_list.push_back(cam_from(cams[c]));
  xreg::FrameTransformList xforms;
  xreg::RayCaster::CamModelAssocList assoc;
  for (uint32_t p = 0; p < n_projs; ++p)
  {
    xforms.push_back(affine_from12(poses + 12 * (std::size_t)p));
    assoc.push_back(cam_idx ? cam_idx[p] : 0);
  }
  xreg::Pt3 bb_min, bb_max;  // ITKImageIndexBoundsAsEigen: [0, size - 1]
  for (int i = 0; i < 3; ++i)
  {
    bb_min(i) = 0;
    bb_max(i) = static_cast<float>(dims[i] - 1);
  }
  const LineIntParams params = {0, &img, bb_min, bb_max, affine_from12(phys_to_idx), n_projs, cam_list, xforms, assoc,
                                step_size, static_cast<xreg::RayCaster::InterpMethod>(interp)};
  const xreg::RangeType full_range(0, (std::size_t)n_projs * cam_list[0].num_det_rows * cam_list[0].num_det_cols);
  if (kernel_id == 0)
    ComputeLineInts<AccumLineIntKernel>(params, proj_buf, full_range);
  else
    ComputeLineInts<MaxLineIntKernel>(params, proj_buf, full_range);
  return 0;
}
'''


METRIC_WRAPPER = r'''
// ---- C ABI over the reference's patch-NCC class (ours) --------------------------------------------------------------
struct xref_patch_opts   // layout of oracle/xreg_oracle.h: xo_patch_opts
{
  uint32_t radius, stride;
  int32_t compute_mean_of_patch_sims, weight_patch_sims, use_mask_for_weighting, use_mask_for_patch_stats,
      normalize_weights_as_prob;
};

extern "C" void xref_patch_mean_std(const float* img, const uint8_t* mask, uint32_t rows, uint32_t cols, uint32_t r0,
                                    uint32_t c0, uint32_t d, int use_mask_for_stats, float* mean, float* sd, uint64_t* n)
{
  cv::Mat im(rows, cols, cv::DataType<float>::type, const_cast<float*>(img));
  cv::Mat mk(rows, cols, cv::DataType<unsigned char>::type, const_cast<uint8_t*>(mask));
  cv::Rect roi;
  roi.x = (int)c0;
  roi.y = (int)r0;
  roi.width = roi.height = (int)d;
  const cv::Mat p = im(roi), m = mk(roi);
  xreg::size_type cnt = 0;
  std::tie(*mean, *sd, cnt) = xreg::detail::ComputePatchMeanStdDev(p, mask ? &m : nullptr, use_mask_for_stats != 0);
  *n = cnt;
}

// set_fixed_image / set_mask / set_mov_imgs_host_buf / patch parameters, allocate_resources(), compute(), sim_vals()
extern "C" int xref_patch_ncc(const float* fixed, const uint8_t* mask, uint32_t rows, uint32_t cols,
                              const xref_patch_opts* o, const float* wgt_img, const float* mov, uint32_t n_imgs,
                              float* sims_out, float* weights_out, float* patch_sims_out,
                              const uint64_t* subset = nullptr, uint64_t n_subset = 0)
{
  using Sim = xreg::ImgSimMetric2DPatchNCCCPU;
  Sim sm;
  sm.fixed_img_.p = std::make_shared<Sim::Image>();
  sm.fixed_img_.p->buf = const_cast<float*>(fixed);
  sm.fixed_img_.p->sz.s[0] = cols;
  sm.fixed_img_.p->sz.s[1] = rows;
  if (mask)
  {
    sm.mask_.p = std::make_shared<Sim::ImageMask>();
    sm.mask_.p->buf = const_cast<uint8_t*>(mask);
    sm.mask_.p->sz.s[0] = cols;
    sm.mask_.p->sz.s[1] = rows;
  }
  if (wgt_img)
  {
    Sim::WgtImgPtr w;
    w.p = std::make_shared<Sim::WgtImg>();
    w.p->buf = const_cast<float*>(wgt_img);
    w.p->sz.s[0] = cols;
    w.p->sz.s[1] = rows;
    sm.set_wgt_img(w);
  }
  std::vector<float> mov_copy(mov, mov + (std::size_t)n_imgs * rows * cols);   // the metric may modify its buffer
  sm.num_mov_imgs_ = n_imgs;
  sm.mov_imgs_buf_ = mov_copy.data();
  sm.patch_radius_ = o->radius;
  sm.patch_stride_ = o->stride;
  sm.compute_mean_of_patch_sims_ = o->compute_mean_of_patch_sims != 0;
  sm.weight_patch_sims_in_combine_ = o->weight_patch_sims != 0;
  sm.use_mask_for_weighting_ = o->use_mask_for_weighting != 0;
  sm.use_mask_for_patch_stats_ = o->use_mask_for_patch_stats != 0;
  sm.normalize_weights_as_prob_ = o->normalize_weights_as_prob != 0;
  sm.save_all_per_patch_scores_ = patch_sims_out != nullptr;
  if (n_subset)
  {
    Sim::PatchIndexList inds(subset, subset + n_subset);
    sm.set_patches_to_use(inds);   // the reference's own lines (xregImgSimMetric2DPatchCommon.cpp:231-235)
  }
  sm.allocate_resources();
  sm.compute();
  for (uint32_t i = 0; i < n_imgs; ++i)
    sims_out[i] = sm.sim_vals_[i];
  const std::size_t np = sm.patch_infos_.size();
  if (weights_out)
    for (std::size_t k = 0; k < np; ++k)
      weights_out[k] = sm.patch_infos_[k].weight;
  if (patch_sims_out)
    for (uint32_t i = 0; i < n_imgs; ++i)
      for (std::size_t k = 0; k < sm.sim_vals_for_each_patch_.size(); ++k)
        patch_sims_out[(std::size_t)i * sm.sim_vals_for_each_patch_.size() + k] = sm.sim_vals_for_each_patch_[k][i];
  return (int)np;
}
'''


def _with_prev(lines, regex, n_prev, n_closing=1):
    s, e = _cut_function(lines, regex, n_closing)
    return s - n_prev, e


def metric_slices():
    out = []
    rel = "lib/regi/sim_metrics_2d/xregImgSimMetric2DPatchCommon.cpp"
    ln = _lines(rel)
    for regex, n_prev in ((r"^xreg::ImgSimMetric2DPatchCommon::PatchInfo::center_row_col\(\) const", 1),
                          (r"^xreg::ImgSimMetric2DPatchCommon::PatchInfo::ocv_roi\(\) const", 1),
                          (r"^xreg::size_type xreg::ImgSimMetric2DPatchCommon::num_patches\(\) const", 0),
                          (r"^void xreg::ImgSimMetric2DPatchCommon::setup_patches\(", 0),
                          (r"^bool xreg::ImgSimMetric2DPatchCommon::compute_weights\(", 0),
                          (r"^void xreg::ImgSimMetric2DPatchCommon::set_patches_to_use\(", 0),
                          (r"^xreg::ImgSimMetric2DPatchCommon::patch_indices_to_use\(\)", 1)):
        s, e = _with_prev(ln, regex, n_prev)
        out.append((rel, s, e, ln[s:e + 1]))
    rel = "lib/regi/sim_metrics_2d/xregImgSimMetric2DPatchNCCCPU.cpp"
    ln = _lines(rel)
    for regex, n_prev in ((r"^void xreg::ImgSimMetric2DPatchNCCCPU::allocate_resources\(\)", 0),
                          (r"^void xreg::ImgSimMetric2DPatchNCCCPU::compute\(\)", 0),
                          (r"^void xreg::ImgSimMetric2DPatchNCCCPU::process_mask\(\)", 0),
                          (r"^xreg::detail::ComputePatchMeanStdDev\(", 3)):
        s, e = _with_prev(ln, regex, n_prev)
        out.append((rel, s, e, ln[s:e + 1]))
    assert out[-1][3][0].startswith("std::tuple<"), out[-1][3][0]
    return out


HU_WRAPPER = r'''
// HUToLinAtt(hu_vol, hu_lower): SetInput, SetHULower, Update (= GenerateData), GetOutput
extern "C" void xref_hu_to_lin_att(const float* hu, float* att, uint64_t n, float hu_lower)
{
  xreg::HUToLinAttFilter::Vol in;
  in.in = hu;
  in.n = (std::size_t)n;
  xreg::HUToLinAttFilter f;
  f.input = &in;
  f.hu_lower_ = hu_lower;   // SetHULower(const double hul)
  f.GenerateData();
  for (uint64_t i = 0; i < n; ++i)
    att[i] = f.output.out[i];
}
'''


def hu_slices():
    rel = "lib/image/xregHUToLinAtt.cpp"
    ln = _lines(rel)
    s, e = _cut_function(ln, r"^void xreg::HUToLinAttFilter::GenerateData\(\)")
    return [(rel, s, e, ln[s:e + 1])]


NCC_WRAPPER = r'''
// set_fixed_image / set_mask / set_mov_imgs_host_buf, allocate_resources(), compute(), sim_vals()
extern "C" void xref_ncc(const float* fixed, const uint8_t* mask, uint32_t rows, uint32_t cols, const float* mov,
                         uint32_t n_imgs, float* sims_out)
{
  using Sim = xreg::ImgSimMetric2DNCCCPU;
  Sim sm;
  sm.fixed_img_.p = std::make_shared<Sim::Image>();
  sm.fixed_img_.p->buf = const_cast<float*>(fixed);
  sm.fixed_img_.p->sz.s[0] = cols;
  sm.fixed_img_.p->sz.s[1] = rows;
  if (mask)
  {
    sm.mask_.p = std::make_shared<Sim::ImageMask>();
    sm.mask_.p->buf = const_cast<uint8_t*>(mask);
    sm.mask_.p->sz.s[0] = cols;
    sm.mask_.p->sz.s[1] = rows;
  }
  std::vector<float> mov_copy(mov, mov + (std::size_t)n_imgs * rows * cols);   // NCC zero-means its buffer in place
  sm.num_mov_imgs_ = n_imgs;
  sm.mov_imgs_buf_ = mov_copy.data();
  sm.allocate_resources();
  sm.compute();
  for (uint32_t i = 0; i < n_imgs; ++i)
    sims_out[i] = sm.sim_vals_[i];
}

extern "C" void xref_ssd(const float* fixed, const uint8_t* mask, uint32_t rows, uint32_t cols, const float* mov,
                         uint32_t n_imgs, float* sims_out)
{
  using Sim = xreg::ImgSimMetric2DSSDCPU;
  Sim sm;
  std::vector<float> fixed_copy(fixed, fixed + (std::size_t)rows * cols);   // the mask is applied to the fixed image in place
  sm.fixed_img_.p = std::make_shared<Sim::Image>();
  sm.fixed_img_.p->buf = fixed_copy.data();
  sm.fixed_img_.p->sz.s[0] = cols;
  sm.fixed_img_.p->sz.s[1] = rows;
  if (mask)
  {
    sm.mask_.p = std::make_shared<Sim::ImageMask>();
    sm.mask_.p->buf = const_cast<uint8_t*>(mask);
    sm.mask_.p->sz.s[0] = cols;
    sm.mask_.p->sz.s[1] = rows;
  }
  std::vector<float> mov_copy(mov, mov + (std::size_t)n_imgs * rows * cols);
  sm.num_mov_imgs_ = n_imgs;
  sm.mov_imgs_buf_ = mov_copy.data();
  sm.allocate_resources();
  sm.compute();
  for (uint32_t i = 0; i < n_imgs; ++i)
    sims_out[i] = sm.sim_vals_[i];
}

// ImgSimMetric2DCombineMean / Addition over n_views metrics holding view_sims[v * n_poses + p]
struct HeldSims : xreg::ImgSimMetric2DCPU
{
  void compute() override {}
};
extern "C" void xref_combine(const float* view_sims, uint32_t n_views, uint32_t n_poses, int mean, float* out)
{
  std::vector<HeldSims> held(n_views);
  xreg::ImgSimMetric2DCombineMean cm;
  xreg::ImgSimMetric2DCombineAddition ca;
  xreg::ImgSimMetric2DCombine* c = mean ? static_cast<xreg::ImgSimMetric2DCombine*>(&cm) : &ca;
  c->num_sim_metrics_ = n_views;
  c->num_projs_per_sim_metric_ = n_poses;
  for (uint32_t v = 0; v < n_views; ++v)
  {
    held[v].sim_vals_.assign(view_sims + (std::size_t)v * n_poses, view_sims + (std::size_t)(v + 1) * n_poses);
    c->sim_objs_.push_back(&held[v]);
  }
  c->compute();
  for (uint32_t p = 0; p < n_poses; ++p)
    out[p] = c->sim_vals_[p];
}
'''


def ncc_slices():
    rel = "lib/regi/sim_metrics_2d/xregImgSimMetric2DNCCCPU.cpp"
    ln = _lines(rel)
    out = []
    k = next(i for i, x in enumerate(ln) if "ComputeLenFromMask(" in x)
    s = max(i for i in range(k) if ln[i].startswith("namespace"))
    e = next(i for i in range(k, len(ln)) if re.match(r"^\}\s*//\s*un-named", ln[i]))
    out.append((rel, s, e, ln[s:e + 1]))
    for regex in (r"^void xreg::ImgSimMetric2DNCCCPU::allocate_resources\(\)", r"^void xreg::ImgSimMetric2DNCCCPU::compute\(\)",
                  r"^void xreg::ImgSimMetric2DNCCCPU::process_mask\(\)"):
        s, e = _cut_function(ln, regex)
        out.append((rel, s, e, ln[s:e + 1]))
    rel = "lib/regi/sim_metrics_2d/xregImgSimMetric2DSSDCPU.cpp"
    ln = _lines(rel)
    k = next(i for i, x in enumerate(ln) if "void ApplyMaskToEigenMatInPlace(" in x)
    s = max(i for i in range(k) if ln[i].startswith("namespace"))
    e = next(i for i in range(k, len(ln)) if re.match(r"^\}\s*//\s*un-named", ln[i]))
    out.append((rel, s, e, ln[s:e + 1]))
    for regex in (r"^void xreg::ImgSimMetric2DSSDCPU::allocate_resources\(\)", r"^void xreg::ImgSimMetric2DSSDCPU::compute\(\)",
                  r"^void xreg::ImgSimMetric2DSSDCPU::process_mask\(\)"):
        s, e = _cut_function(ln, regex)
        out.append((rel, s, e, ln[s:e + 1]))
    rel = "lib/regi/sim_metrics_2d/xregImgSimMetric2DCombine.cpp"
    ln = _lines(rel)
    for regex in (r"^void xreg::ImgSimMetric2DCombineAddition::compute\(\)", r"^void xreg::ImgSimMetric2DCombineMean::compute\(\)"):
        s, e = _cut_function(ln, regex)
        out.append((rel, s, e, ln[s:e + 1]))
    return out


GRAD_WRAPPER = r'''
cv::gauss_fn cv::g_gauss = nullptr;
cv::sobel_fn cv::g_sobel = nullptr;

// which Gaussian / Sobel the classes call: the real OpenCV (through cv2) or the oracle's restatement
extern "C" void xref_set_cv(cv::gauss_fn g, cv::sobel_fn s)
{
  cv::g_gauss = g;
  cv::g_sobel = s;
}

struct xref_patch_opts   // layout of oracle/xreg_oracle.h: xo_patch_opts
{
  uint32_t radius, stride;
  int32_t compute_mean_of_patch_sims, weight_patch_sims, use_mask_for_weighting, use_mask_for_patch_stats,
      normalize_weights_as_prob;
};

template <class Sim>
static void set_images(Sim& sm, const float* fixed, const uint8_t* mask, uint32_t rows, uint32_t cols, float* mov, uint32_t n)
{
  sm.fixed_img_.p = std::make_shared<typename Sim::Image>();
  sm.fixed_img_.p->buf = const_cast<float*>(fixed);
  sm.fixed_img_.p->sz.s[0] = cols;
  sm.fixed_img_.p->sz.s[1] = rows;
  if (mask)
  {
    sm.mask_.p = std::make_shared<typename Sim::ImageMask>();
    sm.mask_.p->buf = const_cast<uint8_t*>(mask);
    sm.mask_.p->sz.s[0] = cols;
    sm.mask_.p->sz.s[1] = rows;
  }
  sm.num_mov_imgs_ = n;
  sm.mov_imgs_buf_ = mov;
}

extern "C" void xref_grad_ncc(const float* fixed, const uint8_t* mask, uint32_t rows, uint32_t cols, int gauss_width,
                              const float* mov, uint32_t n_imgs, float* sims_out)
{
  xreg::ImgSimMetric2DGradNCCCPU sm;
  std::vector<float> mov_copy(mov, mov + (std::size_t)n_imgs * rows * cols);
  set_images(sm, fixed, mask, rows, cols, mov_copy.data(), n_imgs);
  sm.smooth_img_kernel_rad_ = (std::size_t)gauss_width;   // set_smooth_img_before_sobel_kernel_radius
  sm.allocate_resources();
  sm.compute();
  for (uint32_t i = 0; i < n_imgs; ++i)
    sims_out[i] = sm.sim_vals_[i];
}

extern "C" int xref_patch_grad_ncc(const float* fixed, const uint8_t* mask, uint32_t rows, uint32_t cols, int gauss_width,
                                   const xref_patch_opts* o, const float* mov, uint32_t n_imgs, float* sims_out)
{
  xreg::ImgSimMetric2DPatchGradNCCCPU sm;
  std::vector<float> mov_copy(mov, mov + (std::size_t)n_imgs * rows * cols);
  set_images(sm, fixed, mask, rows, cols, mov_copy.data(), n_imgs);
  sm.smooth_img_kernel_rad_ = (std::size_t)gauss_width;
  sm.patch_radius_ = o->radius;
  sm.patch_stride_ = o->stride;
  sm.compute_mean_of_patch_sims_ = o->compute_mean_of_patch_sims != 0;
  sm.weight_patch_sims_in_combine_ = o->weight_patch_sims != 0;
  sm.use_mask_for_weighting_ = o->use_mask_for_weighting != 0;
  sm.use_mask_for_patch_stats_ = o->use_mask_for_patch_stats != 0;
  sm.normalize_weights_as_prob_ = o->normalize_weights_as_prob != 0;
  sm.allocate_resources();
  sm.compute();
  for (uint32_t i = 0; i < n_imgs; ++i)
    sims_out[i] = sm.sim_vals_[i];
  return (int)sm.patch_infos_.size();
}
'''


def grad_slices():
    out = [x for x in ncc_slices() if "Combine" not in x[0] and "SSDCPU" not in x[0]] + metric_slices()
    d = "lib/regi/sim_metrics_2d/"
    more = (
        (d + "xregImgSimMetric2DPatchCommon.cpp", (r"^void xreg::ImgSimMetric2DPatchCommon::set_from_other\(",
                                                   r"^void xreg::ImgSimMetric2DPatchCommon::set_weights_from_other\(")),
        (d + "xregImgSimMetric2DPatchNCCCPU.cpp", (r"^void xreg::ImgSimMetric2DPatchNCCCPU::set_use_fixed_img_patch_variances_as_wgts\(",
                                                   r"^void xreg::ImgSimMetric2DPatchNCCCPU::set_use_mov_img_patch_variances_as_wgts\(",
                                                   r"^void xreg::ImgSimMetric2DPatchNCCCPU::set_other_mov_img_patch_vars\(")),
        (d + "xregImgSimMetric2DGradImgCPU.cpp", (r"^void xreg::ImgSimMetric2DGradImgCPU::allocate_resources\(\)",
                                                  r"^void xreg::ImgSimMetric2DGradImgCPU::compute_sobel_grads\(\)")),
        (d + "xregImgSimMetric2DGradNCCCPU.cpp", (r"^void xreg::ImgSimMetric2DGradNCCCPU::allocate_resources\(\)",
                                                  r"^void xreg::ImgSimMetric2DGradNCCCPU::compute\(\)",
                                                  r"^void xreg::ImgSimMetric2DGradNCCCPU::process_mask\(\)")),
        (d + "xregImgSimMetric2DPatchGradNCCCPU.cpp", (r"^void xreg::ImgSimMetric2DPatchGradNCCCPU::allocate_resources\(\)",
                                                       r"^void xreg::ImgSimMetric2DPatchGradNCCCPU::compute\(\)",
                                                       r"^void xreg::ImgSimMetric2DPatchGradNCCCPU::process_mask\(\)")),
    )
    for rel, regexes in more:
        ln = _lines(rel)
        for regex in regexes:
            s, e = _cut_function(ln, regex)
            out.append((rel, s, e, ln[s:e + 1]))
    # two accessors whose return type sits on the line before the qualified name
    ln = _lines(d + "xregImgSimMetric2DPatchNCCCPU.cpp")
    s, e = _with_prev(ln, r"^xreg::ImgSimMetric2DPatchNCCCPU::aux_info\(\)|aux_info\(\)$", 0)
    if not ln[s].startswith("std::shared_ptr"):
        s -= 1
    out.append((d + "xregImgSimMetric2DPatchNCCCPU.cpp", s, e, ln[s:e + 1]))
    ln = _lines(d + "xregImgSimMetric2DPatchCommon.cpp")
    s, e = _cut_function(ln, r"^xreg::ImgSimMetric2DPatchCommon::sim_vals_for_each_patch\(\) const")
    out.append((d + "xregImgSimMetric2DPatchCommon.cpp", s - 1, e, ln[s - 1:e + 1]))
    return out


SE3_WRAPPER = r'''
// ExpSE3(Pt6): [w_x w_y w_z v_x v_y v_z] -> row-major 3x4
extern "C" void xref_exp_se3(const float x[6], float out12[12])
{
  xreg::Pt6 p;
  for (int i = 0; i < 6; ++i)
    p(i) = x[i];
  const xreg::Mat4x4 T = xreg::ExpSE3(p);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 4; ++j)
      out12[4 * i + j] = T(i, j);
}
'''


def se3_slices():
    out = []
    rel = "lib/transforms/xregRotUtils.cpp"
    ln = _lines(rel)
    for regex in (r"^xreg::Mat3x3 xreg::SkewMatrix\(const Pt3& v\)", r"^xreg::Pt3 xreg::WedgeSkew\(const Mat3x3& W\)",
                  r"^xreg::Mat3x3 xreg::ExpSO3\(const Pt3& x\)", r"^xreg::Mat3x3 xreg::ExpSO3\(const Mat3x3& W\)"):
        s, e = _cut_function(ln, regex)
        out.append((rel, s, e, ln[s:e + 1]))
    rel = "lib/transforms/xregRigidUtils.cpp"
    ln = _lines(rel)
    for regex in (r"^xreg::Mat4x4 xreg::ExpSE3\(const Mat4x4& M\)", r"^xreg::Mat4x4 xreg::ExpSE3\(const Pt6& x\)"):
        s, e = _cut_function(ln, regex)
        out.append((rel, s, e, ln[s:e + 1]))
    return out


CAM_WRAPPER = r'''
struct xref_cam_out
{
  uint32_t rows, cols;
  float intrins_inv[9];
  float extrins_inv[12];
  float pinhole[3];
  float focal_len;
  int32_t frame_type;
};

static void export_cam(const xreg::CameraModel& m, xref_cam_out* o, float* intrins9, float* spacing2)
{
  o->rows = (uint32_t)m.num_det_rows;
  o->cols = (uint32_t)m.num_det_cols;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
    {
      o->intrins_inv[3 * i + j] = m.intrins_inv(i, j);
      if (intrins9)
        intrins9[3 * i + j] = m.intrins(i, j);
    }
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 4; ++j)
      o->extrins_inv[4 * i + j] = m.extrins_inv.matrix()(i, j);
  for (int i = 0; i < 3; ++i)
    o->pinhole[i] = m.pinhole_pt(i);
  o->focal_len = m.focal_len;
  o->frame_type = (int32_t)m.coord_frame_type;
  if (spacing2)
  {
    spacing2[0] = m.det_row_spacing;
    spacing2[1] = m.det_col_spacing;
  }
}

extern "C" void xref_cam_setup_naive(xref_cam_out* o, float focal_len, uint32_t rows, uint32_t cols, float row_spacing,
                                     float col_spacing, int32_t frame_type)
{
  xreg::CameraModel m;
  m.coord_frame_type = static_cast<xreg::CameraModel::CameraCoordFrame>(frame_type);
  m.setup(focal_len, rows, cols, row_spacing, col_spacing);
  export_cam(m, o, nullptr, nullptr);
}

static xreg::CameraModel make_cam(const float intrins[9], const float extrins[16], uint32_t rows, uint32_t cols,
                                  float row_spacing, float col_spacing, int32_t frame_type)
{
  xreg::CameraModel m;
  m.coord_frame_type = static_cast<xreg::CameraModel::CameraCoordFrame>(frame_type);
  xreg::Mat3x3 K;
  xreg::Mat4x4 E;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      K(i, j) = intrins[3 * i + j];
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j)
      E(i, j) = extrins[4 * i + j];
  m.setup(K, E, rows, cols, row_spacing, col_spacing);
  return m;
}

extern "C" void xref_cam_setup(xref_cam_out* o, const float intrins[9], const float extrins[16], uint32_t rows,
                               uint32_t cols, float row_spacing, float col_spacing, int32_t frame_type)
{
  export_cam(make_cam(intrins, extrins, rows, cols, row_spacing, col_spacing, frame_type), o, nullptr, nullptr);
}

// DownsampleCameraModel of a camera made by setup(intrins, extrins): also returns the new intrinsics and spacings
extern "C" void xref_cam_downsample(xref_cam_out* o, float intrins_out[9], float spacing_out[2], const float intrins[9],
                                    const float extrins[16], uint32_t rows, uint32_t cols, float row_spacing,
                                    float col_spacing, int32_t frame_type, float ds_factor, int force_even_dims)
{
  const xreg::CameraModel src = make_cam(intrins, extrins, rows, cols, row_spacing, col_spacing, frame_type);
  export_cam(xreg::DownsampleCameraModel(src, ds_factor, force_even_dims != 0), o, intrins_out, spacing_out);
}
'''


def cam_slices():
    out = []
    rel = "lib/transforms/xregRigidUtils.cpp"
    ln = _lines(rel)
    s, e = _cut_function(ln, r"^xreg::Mat4x4 xreg::SE3Inv\(const Mat4x4& T\)")
    out.append((rel, s, e, ln[s:e + 1]))
    rel = "lib/transforms/xregPerspectiveXform.cpp"
    ln = _lines(rel)
    for regex, n_closing in ((r"^xreg::CoordScalar xreg::FocalLenFromIntrins\(", 1), (r"^xreg::Mat3x3 xreg::MakeNaiveIntrins\(", 1),
                             (r"^void xreg::CameraModel::setup\(const CoordScalar focal_len_arg", 1),
                             (r"^void xreg::CameraModel::setup\(const Mat3x3& intrins_mat, const Mat4x4& extrins_mat", 1),
                             (r"^xreg::CameraModel xreg::DownsampleCameraModel\(", 1)):
        s, e = _cut_function(ln, regex, n_closing)
        out.append((rel, s, e, ln[s:e + 1]))
    return out


ITK_WRAPPER = r'''
// the two helpers RayCasterLineIntCPU::compute starts with, for an image given by its meta data
extern "C" void xref_itk_volume_geometry(const uint64_t dims[3], const double origin[3], const double spacing[3],
                                         const double direction[9], float aabb_min[3], float aabb_max[3],
                                         float idx_to_phys12[12])
{
  itk::Image<float, 3> img;
  for (int i = 0; i < 3; ++i)
  {
    img.size.s[i] = (std::size_t)dims[i];
    img.origin.p[i] = origin[i];
    img.spacing.p[i] = spacing[i];
    for (int j = 0; j < 3; ++j)
      img.direction.d[i][j] = direction[3 * i + j];
  }
  Eigen::Matrix<float, 3, 1> mn, mx;
  std::tie(mn, mx) = xreg::ITKImageIndexBoundsAsEigen(&img);
  const Eigen::Transform<float, 3, Eigen::Affine> t = xreg::ITKImagePhysicalPointTransformsAsEigen(&img);
  for (int i = 0; i < 3; ++i)
  {
    aabb_min[i] = mn[i];
    aabb_max[i] = mx[i];
    for (int j = 0; j < 4; ++j)
      idx_to_phys12[4 * i + j] = t.matrix()(i, j);
  }
}
'''


def itk_slices():
    """Two function templates defined inside `namespace xreg { }` of a header: re-wrapped in that namespace."""
    rel = "lib/itk/xregITKBasicImageUtils.h"
    ln = _lines(rel)
    out = []
    for name in ("ITKImageIndexBoundsAsEigen", "ITKImagePhysicalPointTransformsAsEigen"):
        k = next(i for i, x in enumerate(ln) if x.startswith(name + "("))
        s = max(i for i in range(k) if ln[i].startswith("template <"))
        e = next(i for i in range(k, len(ln)) if ln[i].rstrip() == "}")
        out.append((rel, s, e, ["namespace xreg", "{"] + ln[s:e + 1] + ["}"]))
    return out


LOG_WRAPPER = r"""
// set_* + Update() of the filter on a caller's image; the DiscreteGaussianImageFilter inside is the installed call-out
extern "C" void xref_log_remap_set_gaussian(xref_gaussian_fn fn) { g_xref_gaussian = fn; }

extern "C" void xref_log_remap(const float* img, uint32_t rows, uint32_t cols, int normalize_zero_one,
                               int use_max_intensity_as_I0, float I0, float* out)
{
  itk::Image<float, 2> in;
  in.region.rows = rows;
  in.region.cols = cols;
  in.ext = img;
  xreg::ImageIntensLogTransFilter f;
  f.in = &in;
  f.normalize_zero_one_ = normalize_zero_one != 0;
  f.use_max_intensity_as_I0_ = use_max_intensity_as_I0 != 0;
  f.I0_ = I0;
  f.GenerateData();
  const float* o = static_cast<const itk::Image<float, 2>&>(f.out).GetBufferPointer();
  std::copy(o, o + (std::size_t)rows * cols, out);
}
"""


def log_slices():
    rel = "lib/image/xregImageIntensLogTrans.cpp"
    ln = _lines(rel)
    s, e = _cut_function(ln, r"^void xreg::ImageIntensLogTransFilter::GenerateData\(\)")
    body = ln[s:e + 1]
    assert any("min_pos" in x for x in body) and any("DiscreteGaussianImageFilter" in x for x in body)
    return [(rel, s, e, ["#include <cstdint>"] + body)]


UNITS = (
    # (library, prelude header, slice list function, C ABI wrapper)
    ("libxreg_refslice.so", "ref_pin_prelude.h", slices, WRAPPER),
    ("libxreg_refslice_metric.so", "ref_pin_metric_prelude.h", metric_slices, METRIC_WRAPPER),
    ("libxreg_refslice_hu.so", "ref_pin_hu_prelude.h", hu_slices, HU_WRAPPER),
    ("libxreg_refslice_ncc.so", "ref_pin_ncc_prelude.h", ncc_slices, NCC_WRAPPER),
    ("libxreg_refslice_grad.so", "ref_pin_grad_prelude.h", grad_slices, GRAD_WRAPPER),
    ("libxreg_refslice_se3.so", "ref_pin_se3_prelude.h", se3_slices, SE3_WRAPPER),
    ("libxreg_refslice_cam.so", "ref_pin_cam_prelude.h", cam_slices, CAM_WRAPPER),
    ("libxreg_refslice_itk.so", "ref_pin_itk_prelude.h", itk_slices, ITK_WRAPPER),
    ("libxreg_refslice_log.so", "ref_pin_log_prelude.h", log_slices, LOG_WRAPPER),
)
LOG_LIB = os.path.join(OUT_DIR, "libxreg_refslice_log.so")
ITK_LIB = os.path.join(OUT_DIR, "libxreg_refslice_itk.so")
CAM_LIB = os.path.join(OUT_DIR, "libxreg_refslice_cam.so")
SE3_LIB = os.path.join(OUT_DIR, "libxreg_refslice_se3.so")
GRAD_LIB = os.path.join(OUT_DIR, "libxreg_refslice_grad.so")
NCC_LIB = os.path.join(OUT_DIR, "libxreg_refslice_ncc.so")
HU_LIB = os.path.join(OUT_DIR, "libxreg_refslice_hu.so")
METRIC_LIB = os.path.join(OUT_DIR, "libxreg_refslice_metric.so")


def build(verbose=False):
    """Compile every unit; returns the DRR library path, or None without a reference checkout."""
    if not os.path.isdir(REF):
        return None
    os.makedirs(OUT_DIR, exist_ok=True)
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    for lib_name, prelude, slice_fn, wrapper in UNITS:
        gen = os.path.join(OUT_DIR, lib_name.replace(".so", "_gen.cpp"))
        parts = ["// GENERATED by oracle/ref_pin/build_ref_slice.py -- not tracked; the slices below are the reference's own lines",
                 '#include "%s"' % prelude, ""]
        cut = slice_fn()
        for rel, s, e, body in cut:
            parts.append("// ---- %s:%d-%d " % (rel, s + 1, e + 1) + "-" * 40)
            parts.extend(body)
            parts.append("")
        parts.append(wrapper)
        with open(gen, "w") as f:
            f.write("\n".join(parts))
        cmd = [cxx, "-std=c++11", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-I", HERE, gen, "-o",
               os.path.join(OUT_DIR, lib_name)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if not os.environ.get("XREG_KEEP_REF_SLICE_SOURCE"):
            os.remove(gen)   # the generated unit holds reference source lines: only the binary stays (and only untracked)
        if r.returncode != 0:
            raise RuntimeError("reference slice %s failed to compile:\n%s" % (lib_name, r.stderr[-6000:]))
        if verbose:
            for rel, s, e, _ in cut:
                print("%s:%d-%d" % (rel, s + 1, e + 1))
            print(os.path.join(OUT_DIR, lib_name))
    return LIB


if __name__ == "__main__":
    if build(verbose=True) is None:
        print("reference checkout not found at %s: nothing built" % REF)
        sys.exit(0)
