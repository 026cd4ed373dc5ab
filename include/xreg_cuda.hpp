/*
 * xreg_cuda.hpp -- header-only C++14 host layer over the C ABI (xreg_cuda.h).
 *
 * The reference is C++ (rg2/xreg), but its headers need ITK / Eigen / OpenCV / Boost, none of which
 * exist where this library is built and tested.  This header is the dependency-free C++ mirror of the
 * two reference interfaces the library replaces, with the same class names, method names, argument
 * meaning, call-order contract and error behaviour:
 *
 *   xreg::RayCaster + RayCastLineIntParamInterface      lib/ray_cast/xregRayCastInterface.h:43-434,575-591
 *   xreg::ImgSimMetric2D                                 lib/regi/sim_metrics_2d/xregImgSimMetric2D.h:42-156
 *   xreg::ImgSimMetric2DGradImgParamInterface            .../xregImgSimMetric2DGradImgParamInterface.h:31-39
 *   xreg::ImgSimMetric2DPatchCommon                      .../xregImgSimMetric2DPatchCommon.{h,cpp}
 *   xreg::ImgSimMetric2DCombine{Addition,Mean}           .../xregImgSimMetric2DCombine.{h,cpp}
 *   xreg::CameraModel                                    lib/transforms/xregPerspectiveXform.{h,cpp}
 *
 * ITK images become plain views (Volume, Image2D: pointer + sizes + ITK metadata), Eigen transforms a
 * row-major 4x4 float matrix (FrameTransform).  Everything else reads like the reference, so the C++
 * parity test (tests/cpp/host_mirror_test.cpp) is written the way a test inside an xReg checkout
 * would be.  The classes that additionally derive from the real xreg:: base classes -- what a maintainer
 * compiles inside an xReg checkout -- are adapters/xreg/ (INTEGRATION.md).
 *
 * Errors: the reference throws (xregASSERT / xregThrow / UnsupportedOperationException); so does this
 * layer: XregCudaError for XRC_ERR_INVALID / _CUDA / _NOMEM, UnsupportedOperationException for
 * XRC_ERR_UNSUPPORTED.  There is no CPU fallback anywhere in this header: every compute call goes to
 * libxreg_cuda.so and fails loudly without a device.
 */
#ifndef XREG_CUDA_HPP
#define XREG_CUDA_HPP

#include <cmath>
#include <cstddef>
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "xreg_cuda.h"

namespace xreg_b200
{

using size_type = std::size_t;
using CoordScalar = float;           // lib/common/xregCommon.h:47
using RayCastPixelScalar = float;    // lib/common/xregCommon.h:128

/// xregThrow / xregASSERT failures (lib/common/xregExceptionUtils.h:36-63, xregAssert.h)
class XregCudaError : public std::runtime_error
{
public:
  XregCudaError(const int status, const std::string& msg) : std::runtime_error(msg), status_(status) {}
  int status() const { return status_; }

private:
  int status_;
};

/// RayCaster::UnsupportedOperationException (xregRayCastInterface.h:59)
class UnsupportedOperationException : public XregCudaError
{
public:
  explicit UnsupportedOperationException(const std::string& msg) : XregCudaError(XRC_ERR_UNSUPPORTED, msg) {}
};

namespace detail
{
inline void Check(const int status)
{
  if (status == XRC_OK)
  {
    return;
  }
  const char* m = xrc_last_error();
  const std::string msg = std::string("xreg_cuda: ") + (m ? m : "unknown error");
  if (status == XRC_ERR_UNSUPPORTED)
  {
    throw UnsupportedOperationException(msg);
  }
  throw XregCudaError(status, msg);
}

inline void Assert(const bool cond, const char* what)
{
  if (!cond)
  {
    throw XregCudaError(XRC_ERR_INVALID, std::string("assertion failed: ") + what);
  }
}

/// f32 dot product in the left-to-right order of an un-vectorised Eigen fixed-size product
inline float Dot3(const float a0, const float a1, const float a2, const float b0, const float b1, const float b2)
{
  return ((a0 * b0) + (a1 * b1)) + (a2 * b2);
}

/// 3x3 inverse by cofactors, inv(i,j) = cof(j,i) / det (the shape of Eigen's fixed-size 3x3 inverse;
/// call site xregPerspectiveXform.cpp:247)
inline void Inverse3x3(const float m[9], float out[9])
{
  auto cof = [&](const int i, const int j) {
    const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
    return (m[3 * i1 + j1] * m[3 * i2 + j2]) - (m[3 * i1 + j2] * m[3 * i2 + j1]);
  };
  const float det = ((cof(0, 0) * m[0]) + (cof(1, 0) * m[3])) + (cof(2, 0) * m[6]);
  const float inv_det = 1.0f / det;
  for (int i = 0; i < 3; ++i)
  {
    for (int j = 0; j < 3; ++j)
    {
      out[3 * i + j] = cof(j, i) * inv_det;
    }
  }
}
}  // namespace detail

/// Eigen::Transform<float,3,Affine> stand-in: row-major 4x4, last row (0 0 0 1).
struct FrameTransform
{
  float m[16];

  FrameTransform()
  {
    for (int i = 0; i < 16; ++i)
    {
      m[i] = (i % 5 == 0) ? 1.0f : 0.0f;
    }
  }

  static FrameTransform Identity() { return FrameTransform(); }

  /// from the 12 floats of a row-major 3x4
  static FrameTransform From3x4(const float* a)
  {
    FrameTransform t;
    for (int i = 0; i < 12; ++i)
    {
      t.m[i] = a[i];
    }
    return t;
  }

  float& operator()(const int r, const int c) { return m[4 * r + c]; }
  const float& operator()(const int r, const int c) const { return m[4 * r + c]; }

  /// affine * affine: linear = A B, translation = (A b_t) + a_t
  FrameTransform operator*(const FrameTransform& b) const
  {
    FrameTransform o;
    for (int r = 0; r < 3; ++r)
    {
      for (int c = 0; c < 3; ++c)
      {
        o(r, c) = detail::Dot3(m[4 * r], m[4 * r + 1], m[4 * r + 2], b(0, c), b(1, c), b(2, c));
      }
      o(r, 3) = detail::Dot3(m[4 * r], m[4 * r + 1], m[4 * r + 2], b(0, 3), b(1, 3), b(2, 3)) + m[4 * r + 3];
    }
    return o;
  }

  /// SE3Inv (lib/transforms/xregRigidUtils.cpp:29-38): R^T, -1 * R^T * t
  FrameTransform rigid_inverse() const
  {
    FrameTransform o;
    for (int r = 0; r < 3; ++r)
    {
      for (int c = 0; c < 3; ++c)
      {
        o(r, c) = (*this)(c, r);
      }
    }
    for (int r = 0; r < 3; ++r)
    {
      o(r, 3) = detail::Dot3(-1.0f * o(r, 0), -1.0f * o(r, 1), -1.0f * o(r, 2), m[3], m[7], m[11]);
    }
    return o;
  }

  void to3x4(float* out) const
  {
    for (int i = 0; i < 12; ++i)
    {
      out[i] = m[i];
    }
  }
};

using FrameTransformList = std::vector<FrameTransform>;

/// ExpSE3(Pt6) (lib/transforms/xregRigidUtils.cpp:40-85), [w_x w_y w_z v_x v_y v_z]; evaluated by the library
inline FrameTransform ExpSE3(const float params[6])
{
  float a[12];
  xrc_exp_se3(params, a);
  return FrameTransform::From3x4(a);
}

/// CameraModel (lib/transforms/xregPerspectiveXform.h:108-273): the members the ray caster reads.
struct CameraModel
{
  enum CameraCoordFrame  // xregPerspectiveXform.h:128-133
  {
    kORIGIN_AT_FOCAL_PT_DET_POS_Z = 0,
    kORIGIN_AT_FOCAL_PT_DET_NEG_Z = 1,
    kORIGIN_ON_DETECTOR = 2
  };

  CameraCoordFrame coord_frame_type = kORIGIN_AT_FOCAL_PT_DET_NEG_Z;

  float intrins[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  float intrins_inv[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  FrameTransform extrins;
  FrameTransform extrins_inv;
  float pinhole_pt[3] = {0, 0, 0};
  CoordScalar focal_len = 0;
  size_type num_det_rows = 0;
  size_type num_det_cols = 0;
  CoordScalar det_row_spacing = 0;
  CoordScalar det_col_spacing = 0;

  /// setup(focal_len, nr, nc, rs, cs) with MakeNaiveIntrins (xregPerspectiveXform.cpp:200-254):
  /// principal point at the detector centre, identity extrinsic.
  void setup(const CoordScalar fl, const size_type nr, const size_type nc, const CoordScalar rs, const CoordScalar cs)
  {
    detail::Assert(fl > 1.0e-8f && nr && nc && rs > 1.0e-8f && cs > 1.0e-8f, "CameraModel::setup arguments");
    float K[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    K[0] = fl / cs;
    K[4] = fl / rs;
    if (coord_frame_type == kORIGIN_AT_FOCAL_PT_DET_NEG_Z)
    {
      K[0] *= -1;
      K[4] *= -1;
    }
    K[2] = static_cast<float>(static_cast<double>(nc - 1) * 0.5);
    K[5] = static_cast<float>(static_cast<double>(nr - 1) * 0.5);
    for (int i = 0; i < 9; ++i)
    {
      intrins[i] = K[i];
    }
    detail::Inverse3x3(K, intrins_inv);
    extrins = FrameTransform();
    extrins_inv = FrameTransform();
    pinhole_pt[0] = pinhole_pt[1] = pinhole_pt[2] = 0;  // Pt3::Zero() for every frame type (:253)
    focal_len = fl;
    num_det_rows = nr;
    num_det_cols = nc;
    det_row_spacing = rs;
    det_col_spacing = cs;
  }

  /// setup(intrins, extrins, nr, nc, rs, cs) (xregPerspectiveXform.cpp:302-334)
  void setup(const float K[9], const FrameTransform& ext, const size_type nr, const size_type nc, const CoordScalar rs,
             const CoordScalar cs)
  {
    num_det_rows = nr;
    num_det_cols = nc;
    det_row_spacing = rs;
    det_col_spacing = cs;
    for (int i = 0; i < 9; ++i)
    {
      intrins[i] = K[i];
    }
    detail::Inverse3x3(K, intrins_inv);
    // FocalLenFromIntrins (xregPerspectiveXform.cpp:186-190)
    focal_len = (std::fabs(K[0] * cs) + std::fabs(K[4] * ((rs < 0) ? cs : rs))) / 2.0f;
    extrins = ext;
    extrins_inv = ext.rigid_inverse();
    if (coord_frame_type == kORIGIN_ON_DETECTOR)
    {
      for (int r = 0; r < 3; ++r)
      {
        pinhole_pt[r] = detail::Dot3(extrins_inv(r, 0), extrins_inv(r, 1), extrins_inv(r, 2), 0.0f, 0.0f, focal_len) +
                        extrins_inv(r, 3);
      }
    }
    else
    {
      for (int r = 0; r < 3; ++r)
      {
        pinhole_pt[r] = extrins_inv(r, 3);
      }
    }
  }

  xrc_cam to_xrc() const
  {
    xrc_cam c;
    c.rows = static_cast<uint32_t>(num_det_rows);
    c.cols = static_cast<uint32_t>(num_det_cols);
    for (int i = 0; i < 9; ++i)
    {
      c.intrins_inv[i] = intrins_inv[i];
    }
    extrins_inv.to3x4(c.extrins_inv);
    for (int i = 0; i < 3; ++i)
    {
      c.pinhole[i] = pinhole_pt[i];
    }
    c.focal_len = focal_len;
    c.frame_type = static_cast<int32_t>(coord_frame_type);
    return c;
  }
};

/// itk::Image<float,3> stand-in: a non-owning view of x-fastest voxels plus the ITK metadata (doubles).
/// Like the reference (ITK smart pointers retained in vols_), the memory must stay valid until
/// set_volumes() returns; the library copies it to the device there.
struct Volume
{
  const float* data = nullptr;
  uint64_t size[3] = {0, 0, 0};  // nx, ny, nz
  double spacing[3] = {1, 1, 1};
  double origin[3] = {0, 0, 0};
  double direction[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};  // row-major

  /// ITKImagePhysicalPointTransformsAsEigen (lib/itk/xregITKBasicImageUtils.h:131-168):
  /// M(r,c) = float(Dir(r,c) * spacing[c]), t(r) = float(origin[r])
  void idx_to_phys(float out[12]) const
  {
    for (int r = 0; r < 3; ++r)
    {
      for (int c = 0; c < 3; ++c)
      {
        out[4 * r + c] = static_cast<float>(direction[3 * r + c] * spacing[c]);
      }
      out[4 * r + 3] = static_cast<float>(origin[r]);
    }
  }
};

/// itk::Image<T,2> stand-in: non-owning row-major view.
template <class T>
struct Image2D
{
  T* data = nullptr;
  size_type rows = 0;
  size_type cols = 0;

  Image2D() = default;
  Image2D(T* d, const size_type r, const size_type c) : data(d), rows(r), cols(c) {}
  explicit operator bool() const { return data != nullptr; }
};

/// One CUDA device + stream: replaces the (boost::compute::context, command_queue) pair the OpenCL classes
/// take (lib/ray_cast/xregRayCastBaseOCL.h:60-75, lib/common/xregProgOptUtils.cpp:1712-1725).
class Context
{
public:
  explicit Context(const int device = 0) { detail::Check(xrc_ctx_create(device, &ctx_)); }
  Context(const int device, void* cuda_stream) { detail::Check(xrc_ctx_create_on_stream(device, cuda_stream, &ctx_)); }
  ~Context() { xrc_ctx_destroy(ctx_); }
  Context(const Context&) = delete;
  Context& operator=(const Context&) = delete;

  void synchronize() { detail::Check(xrc_ctx_synchronize(ctx_)); }
  xrc_ctx* handle() const { return ctx_; }

private:
  xrc_ctx* ctx_ = nullptr;
};

/// xreg::RayCaster (xregRayCastInterface.h:43-434): state and bookkeeping exactly as the reference base class;
/// device work is in the derived class.
class RayCaster
{
public:
  using PixelScalar2D = RayCastPixelScalar;
  using PixelScalar3D = RayCastPixelScalar;
  using VolList = std::vector<Volume>;
  using Proj = Image2D<PixelScalar2D>;
  using ProjList = std::vector<Proj>;
  using CameraModelList = std::vector<CameraModel>;
  using CamModelAssocList = std::vector<size_type>;

  enum InterpMethod  // :61-67
  {
    kRAY_CAST_INTERP_LINEAR = 0,
    kRAY_CAST_INTERP_NN,
    kRAY_CAST_INTERP_SINC,
    kRAY_CAST_INTERP_BSPLINE
  };

  enum ProjPixelStoreMethod  // :71-75
  {
    kRAY_CAST_PIXEL_REPLACE = 0,
    kRAY_CAST_PIXEL_ACCUM
  };

  RayCaster() = default;
  virtual ~RayCaster() = default;
  RayCaster(const RayCaster&) = delete;
  RayCaster& operator=(const RayCaster&) = delete;

  void set_volume(const Volume& vol) { set_volumes(VolList(1, vol)); }

  void set_volumes(const VolList& vols)
  {
    vols_ = vols;
    vols_changed();
  }

  size_type num_vols() const { return vols_.size(); }

  size_type num_camera_models() const { return camera_models_.size(); }

  void set_camera_model(const CameraModel& cam) { set_camera_models(CameraModelList(1, cam)); }

  void set_camera_models(const CameraModelList& cams)
  {
    camera_models_ = cams;
    camera_models_changed();
  }

  const CameraModelList& camera_models() const { return camera_models_; }

  const CameraModel& camera_model(const size_type cam_idx = 0) const { return camera_models_.at(cam_idx); }

  void set_proj_cam_model(const size_type proj_idx, const size_type cam_idx)
  {
    cam_model_for_proj_.at(proj_idx) = cam_idx;
    poses_dirty_ = true;
  }

  const CamModelAssocList& camera_model_proj_associations() const { return cam_model_for_proj_; }

  void set_camera_model_proj_associations(const CamModelAssocList& a)
  {
    detail::Assert(a.size() == num_projs_, "cam_model_for_proj.size() == num_projs_");  // xregRayCastInterface.cpp:91
    cam_model_for_proj_ = a;
    poses_dirty_ = true;
  }

  /// xregRayCastInterface.cpp:97-114: camera-major replication, projection c * n + p <- (pose p, camera c)
  void distribute_xforms_among_cam_models(const FrameTransformList& xforms)
  {
    const size_type n = xforms.size();
    const size_type num_cams = num_camera_models();
    detail::Assert((n * num_cams) == num_projs(), "(num_passed_xforms * num_cams) == num_projs()");
    detail::Assert(xforms_cam_to_itk_phys_.size() >= num_projs_, "resources allocated");
    size_type g = 0;
    for (size_type cam_idx = 0; cam_idx < num_cams; ++cam_idx)
    {
      for (size_type p = 0; p < n; ++p, ++g)
      {
        xforms_cam_to_itk_phys_[g] = xforms[p];
        cam_model_for_proj_[g] = cam_idx;
      }
    }
    poses_dirty_ = true;
  }

  void distribute_xform_among_cam_models(const FrameTransform& xform)
  {
    distribute_xforms_among_cam_models(FrameTransformList(1, xform));
  }

  void set_ray_step_size(const CoordScalar& step_size)
  {
    ray_step_size_ = step_size;
    params_dirty_ = true;
  }

  CoordScalar ray_step_size() const { return ray_step_size_; }

  /// xregRayCastInterface.cpp:131-139: before allocation this sizes the capacity, afterwards n <= capacity
  virtual void set_num_projs(const size_type num_projs)
  {
    detail::Assert(!max_num_projs_ || (num_projs <= max_num_projs_), "num_projs <= max_num_projs_");
    num_projs_ = num_projs;
    cam_model_for_proj_.resize(num_projs_);
    poses_dirty_ = true;
  }

  size_type num_projs() const { return num_projs_; }

  void set_interp_method(const InterpMethod& m)
  {
    interp_method_ = m;
    params_dirty_ = true;
  }

  InterpMethod interp_method() const { return interp_method_; }
  void use_linear_interp() { set_interp_method(kRAY_CAST_INTERP_LINEAR); }
  void use_nn_interp() { set_interp_method(kRAY_CAST_INTERP_NN); }
  void use_sinc_interp() { set_interp_method(kRAY_CAST_INTERP_SINC); }
  void use_bspline_interp() { set_interp_method(kRAY_CAST_INTERP_BSPLINE); }

  void set_xforms_cam_to_itk_phys(const FrameTransformList& xforms)
  {
    detail::Assert(xforms.size() == num_projs_, "xforms.size() == num_projs_");
    xforms_cam_to_itk_phys_ = xforms;
    poses_dirty_ = true;
  }

  const FrameTransformList& xforms_cam_to_itk_phys() const { return xforms_cam_to_itk_phys_; }

  /// mutable reference like the reference accessor (:215); the pose list is re-sent at the next compute()
  FrameTransform& xform_cam_to_itk_phys(const size_type proj_idx)
  {
    poses_dirty_ = true;
    return xforms_cam_to_itk_phys_.at(proj_idx);
  }

  const FrameTransform& xform_cam_to_itk_phys(const size_type proj_idx) const { return xforms_cam_to_itk_phys_.at(proj_idx); }

  void post_multiply_all_xforms(const FrameTransform& post_xform)
  {
    for (size_type i = 0; i < num_projs_; ++i)
    {
      xforms_cam_to_itk_phys_[i] = xforms_cam_to_itk_phys_[i] * post_xform;
    }
    poses_dirty_ = true;
  }

  void pre_multiply_all_xforms(const FrameTransform& pre_xform)
  {
    for (size_type i = 0; i < num_projs_; ++i)
    {
      xforms_cam_to_itk_phys_[i] = pre_xform * xforms_cam_to_itk_phys_[i];
    }
    poses_dirty_ = true;
  }

  /// xregRayCastInterface.cpp:262-275
  virtual void allocate_resources()
  {
    detail::Assert(num_projs_ != 0, "num_projs_");
    detail::Assert(!camera_models_.empty(), "!camera_models_.empty()");
    max_num_projs_ = num_projs_;
    xforms_cam_to_itk_phys_.resize(num_projs_);
    cam_model_for_proj_.resize(num_projs_, 0);
    resources_allocated_ = true;
  }

  virtual void compute(const size_type vol_idx = 0) = 0;

  /// non-owning view of projection proj_idx in the host buffer, invalidated by the next compute() (:263-266)
  virtual Proj proj(const size_type proj_idx) = 0;

  virtual PixelScalar2D* raw_host_pixel_buf() = 0;

  virtual void use_external_host_pixel_buf(void* buf) = 0;

  virtual size_type max_num_projs_possible() const = 0;

  void set_proj_store_method(const ProjPixelStoreMethod m)
  {
    proj_store_meth_ = m;
    params_dirty_ = true;
  }

  ProjPixelStoreMethod proj_store_method() const { return proj_store_meth_; }
  void use_proj_store_replace_method() { set_proj_store_method(kRAY_CAST_PIXEL_REPLACE); }
  void use_proj_store_accum_method() { set_proj_store_method(kRAY_CAST_PIXEL_ACCUM); }

  virtual void use_other_proj_buf(RayCaster* other_ray_caster) = 0;

  void set_use_bg_projs(const bool use_bg_projs)
  {
    use_bg_projs_ = use_bg_projs;
    bg_projs_changed(false);
  }

  bool use_bg_projs() const { return use_bg_projs_; }

  void set_bg_proj(const Proj& proj, const bool use_bg_projs = true) { set_bg_projs(ProjList(1, proj), use_bg_projs); }

  void set_bg_projs(const ProjList& projs, const bool use_bg_projs = true)
  {
    bg_projs_ = projs;
    use_bg_projs_ = use_bg_projs;
    bg_projs_changed(true);
  }

  size_type max_num_projs() const { return max_num_projs_; }

  PixelScalar2D default_bg_pixel_val() const { return default_bg_pixel_val_; }

  void set_default_bg_pixel_val(const PixelScalar2D bg_val)
  {
    default_bg_pixel_val_ = bg_val;
    params_dirty_ = true;
  }

protected:
  virtual void vols_changed() {}
  virtual void camera_models_changed() {}
  virtual void bg_projs_changed(const bool /*new_images*/) {}

  VolList vols_;
  CameraModelList camera_models_;
  FrameTransformList xforms_cam_to_itk_phys_;
  CamModelAssocList cam_model_for_proj_;
  size_type num_projs_ = 0;
  size_type max_num_projs_ = 0;
  CoordScalar ray_step_size_ = 1;  // :390
  InterpMethod interp_method_ = kRAY_CAST_INTERP_LINEAR;
  ProjPixelStoreMethod proj_store_meth_ = kRAY_CAST_PIXEL_REPLACE;
  PixelScalar2D default_bg_pixel_val_ = 0;
  bool use_bg_projs_ = false;
  ProjList bg_projs_;
  bool resources_allocated_ = false;
  bool poses_dirty_ = true;
  bool params_dirty_ = true;
};

enum RayCastLineIntKernel  // xregRayCastInterface.h:575-579
{
  kRAY_CAST_LINE_INT_SUM_KERNEL = 0,
  kRAY_CAST_LINE_INT_MAX_KERNEL
};

/// xregRayCastInterface.h:581-591
class RayCastLineIntParamInterface
{
public:
  RayCastLineIntKernel kernel_id() const { return kernel_id_; }

  void set_kernel_id(const RayCastLineIntKernel k)
  {
    kernel_id_ = k;
    kernel_dirty_ = true;
  }

protected:
  RayCastLineIntKernel kernel_id_ = kRAY_CAST_LINE_INT_SUM_KERNEL;
  bool kernel_dirty_ = true;
};

/// Replaces RayCasterLineIntOCL (lib/ray_cast/xregRayCastLineIntOCL.{h,cpp}); results follow
/// RayCasterLineIntCPU (xregRayCastLineIntCPU.cpp:105-349).
class RayCasterLineIntCUDA : public RayCaster, public RayCastLineIntParamInterface
{
public:
  explicit RayCasterLineIntCUDA(Context& ctx) : ctx_(ctx) { detail::Check(xrc_rc_create(ctx.handle(), &rc_)); }

  ~RayCasterLineIntCUDA() override { xrc_rc_destroy(rc_); }

  void set_num_projs(const size_type num_projs) override
  {
    RayCaster::set_num_projs(num_projs);
    if (resources_allocated_)
    {
      detail::Check(xrc_rc_set_num_projs(rc_, static_cast<uint32_t>(num_projs)));
    }
    host_valid_ = false;
  }

  void allocate_resources() override
  {
    RayCaster::allocate_resources();
    detail::Check(xrc_rc_allocate(rc_, static_cast<uint32_t>(num_projs_)));
    const CameraModel& cam = camera_models_[0];
    num_pix_per_proj_ = cam.num_det_rows * cam.num_det_cols;
    if (!ext_host_buf_)
    {
      host_buf_.assign(max_num_projs_ * num_pix_per_proj_, 0.0f);
    }
    poses_dirty_ = params_dirty_ = true;
    host_valid_ = false;
  }

  /// blocking like the reference: "results must be complete on return" holds for everything a caller can
  /// observe (proj(), raw_host_pixel_buf(), a metric's sim_vals()); the kernel itself is asynchronous on
  /// the context stream and the observers synchronise.
  void compute(const size_type vol_idx = 0) override
  {
    detail::Assert(resources_allocated_, "resources_allocated_ (xregRayCastLineIntCPU.cpp:296)");
    flush();
    detail::Check(xrc_rc_compute(rc_, static_cast<uint32_t>(vol_idx)));
    host_valid_ = false;
  }

  Proj proj(const size_type proj_idx) override
  {
    detail::Assert(proj_idx < num_projs_, "proj_idx < num_projs_");
    sync_host();
    return Proj(host_ptr() + proj_idx * num_pix_per_proj_, camera_models_[cam_model_for_proj_[proj_idx]].num_det_rows,
                camera_models_[cam_model_for_proj_[proj_idx]].num_det_cols);
  }

  PixelScalar2D* raw_host_pixel_buf() override
  {
    sync_host();
    return host_ptr();
  }

  /// xregRayCastBaseCPU.cpp:120-126: the caller's buffer receives the projections (num_projs x rows x cols)
  void use_external_host_pixel_buf(void* buf) override
  {
    ext_host_buf_ = static_cast<PixelScalar2D*>(buf);
    host_valid_ = false;
  }

  size_type max_num_projs_possible() const override
  {
    uint64_t n = 0;
    detail::Check(xrc_rc_max_projs_possible(rc_, &n));
    return static_cast<size_type>(n);
  }

  void use_other_proj_buf(RayCaster* other_ray_caster) override
  {
    auto* o = dynamic_cast<RayCasterLineIntCUDA*>(other_ray_caster);
    if (!o)
    {
      throw UnsupportedOperationException("use_other_proj_buf: the other ray caster is not a CUDA ray caster");
    }
    detail::Check(xrc_rc_use_other_proj_buf(rc_, o->rc_));
  }

  /// Hounsfield-unit volumes converted on the device while loading (lib/image/xregHUToLinAtt.cpp:45-69)
  void set_volumes_hu(const VolList& vols, const float hu_lower = -1000.0f)
  {
    vols_ = vols;
    push_volumes(true, hu_lower);
  }

  void set_skip_empty(const bool enable) { detail::Check(xrc_rc_set_skip_empty(rc_, enable ? 1 : 0)); }

  /// parity instrumentation: clip mask, samples per ray and their total for the current poses
  uint64_t ray_info(uint8_t* mask, uint32_t* steps, const size_type vol_idx = 0)
  {
    flush();
    uint64_t total = 0;
    detail::Check(xrc_rc_ray_info(rc_, static_cast<uint32_t>(vol_idx), mask, steps, &total));
    return total;
  }

  /// device hand-off to the CUDA metrics (replaces to_ocl_buf(), xregRayCastInterface.h:324)
  xrc_rc* handle() { return rc_; }
  Context& context() { return ctx_; }

  /// send pending parameters to the library
  void flush_params()
  {
    if (params_dirty_ || kernel_dirty_)
    {
      detail::Check(xrc_rc_set_params(rc_, ray_step_size_, static_cast<int>(interp_method_), static_cast<int>(kernel_id_),
                                      static_cast<int>(proj_store_meth_), default_bg_pixel_val_));
      params_dirty_ = kernel_dirty_ = false;
    }
  }

  /// send pending parameters and poses to the library (what compute() does first)
  void flush()
  {
    flush_params();
    if (poses_dirty_ && num_projs_)
    {
      tmp_poses_.resize(12 * num_projs_);
      tmp_cam_idx_.resize(num_projs_);
      for (size_type i = 0; i < num_projs_; ++i)
      {
        xforms_cam_to_itk_phys_[i].to3x4(&tmp_poses_[12 * i]);
        tmp_cam_idx_[i] = static_cast<uint32_t>(cam_model_for_proj_[i]);
      }
      detail::Check(xrc_rc_set_poses(rc_, static_cast<uint32_t>(num_projs_), tmp_poses_.data(), tmp_cam_idx_.data()));
      poses_dirty_ = false;
    }
  }

protected:
  void vols_changed() override { push_volumes(false, 0.0f); }

  void camera_models_changed() override
  {
    std::vector<xrc_cam> cams;
    cams.reserve(camera_models_.size());
    for (const auto& c : camera_models_)
    {
      cams.push_back(c.to_xrc());
    }
    detail::Check(xrc_rc_set_cameras(rc_, static_cast<uint32_t>(cams.size()), cams.data()));
  }

  void bg_projs_changed(const bool new_images) override
  {
    if (use_bg_projs_ && new_images)
    {
      detail::Assert(bg_projs_.size() == num_camera_models(), "one background projection per camera model");
      std::vector<const float*> ptrs;
      for (const auto& p : bg_projs_)
      {
        ptrs.push_back(p.data);
      }
      detail::Check(xrc_rc_set_bg_projs(rc_, ptrs.data(), 1));
    }
    else
    {
      detail::Check(xrc_rc_set_bg_projs(rc_, nullptr, use_bg_projs_ ? 1 : 0));
    }
  }

protected:
  friend class Intensity2D3DObjFn;

  /// xrc_obj_fn sized the library's ray caster for num_projs projections and distributed its own poses
  void library_resized(const size_type num_projs)
  {
    num_projs_ = num_projs;
    cam_model_for_proj_.resize(num_projs_);
    poses_dirty_ = true;
    host_valid_ = false;
  }

  void push_volumes(const bool hu, const float hu_lower)
  {
    const size_type n = vols_.size();
    std::vector<const float*> ptrs(n);
    std::vector<uint64_t> dims(3 * n);
    std::vector<float> xf(12 * n);
    for (size_type i = 0; i < n; ++i)
    {
      ptrs[i] = vols_[i].data;
      for (int k = 0; k < 3; ++k)
      {
        dims[3 * i + k] = vols_[i].size[k];
      }
      vols_[i].idx_to_phys(&xf[12 * i]);
    }
    auto* d = reinterpret_cast<const uint64_t(*)[3]>(dims.data());
    auto* x = reinterpret_cast<const float(*)[12]>(xf.data());
    if (hu)
    {
      detail::Check(xrc_rc_set_volumes_hu(rc_, static_cast<uint32_t>(n), ptrs.data(), d, x, hu_lower));
    }
    else
    {
      detail::Check(xrc_rc_set_volumes(rc_, static_cast<uint32_t>(n), ptrs.data(), d, x));
    }
  }

  PixelScalar2D* host_ptr() { return ext_host_buf_ ? ext_host_buf_ : host_buf_.data(); }

  /// lazy device -> host copy: the analogue of RayCastSyncHostBufFromOCL::sync()
  /// (lib/ray_cast/xregRayCastSyncBuf.cpp:60-110), only when a host consumer asks
  void sync_host()
  {
    if (!host_valid_ && num_projs_)
    {
      detail::Assert(resources_allocated_, "resources_allocated_");
      detail::Check(xrc_rc_read_projs(rc_, 0, static_cast<uint32_t>(num_projs_), host_ptr()));
      host_valid_ = true;
    }
  }

  Context& ctx_;
  xrc_rc* rc_ = nullptr;
  size_type num_pix_per_proj_ = 0;
  std::vector<PixelScalar2D> host_buf_;
  PixelScalar2D* ext_host_buf_ = nullptr;
  bool host_valid_ = false;
  std::vector<float> tmp_poses_;
  std::vector<uint32_t> tmp_cam_idx_;
};

/// xreg::RayCasterDepthCPU (lib/ray_cast/xregRayCastDepthCPU.{h,cpp}) with RayCasterCollisionParamInterface
/// (xregRayCastInterface.h:436-475): per pixel the depth of the first sample >= render_thresh() along the ray, refined
/// by num_backtracking_steps() halvings of the step, min-combined on top of the background (kRAY_CAST_MAX_DEPTH by
/// default, as the class's constructor sets it).  Everything else is the line-integral ray caster's state.
constexpr float kRAY_CAST_MAX_DEPTH = XRC_RAY_CAST_MAX_DEPTH;  // xregRayCastInterface.h:601

class RayCasterDepthCUDA : public RayCasterLineIntCUDA
{
public:
  explicit RayCasterDepthCUDA(Context& ctx) : RayCasterLineIntCUDA(ctx) { set_default_bg_pixel_val(kRAY_CAST_MAX_DEPTH); }

  void set_render_thresh(const PixelScalar3D t) { render_thresh_ = t; }
  PixelScalar3D render_thresh() const { return render_thresh_; }
  void set_num_backtracking_steps(const size_type n) { num_backtracking_steps_ = n; }
  size_type num_backtracking_steps() const { return num_backtracking_steps_; }

  void compute(const size_type vol_idx = 0) override
  {
    detail::Assert(resources_allocated_, "resources_allocated_ (xregRayCastDepthCPU.cpp:238)");
    flush();
    detail::Check(xrc_rc_compute_depth(rc_, static_cast<uint32_t>(vol_idx), render_thresh_,
                                       static_cast<uint32_t>(num_backtracking_steps_)));
    host_valid_ = false;
  }

private:
  PixelScalar3D render_thresh_ = 150;          // xregRayCastInterface.cpp:427-428
  size_type num_backtracking_steps_ = 0;
};

/// xreg::ImgSimMetric2D (xregImgSimMetric2D.h:42-156) over the xrc_sm_* entry points.
class ImgSimMetric2D
{
public:
  using Scalar = RayCastPixelScalar;
  using Image = Image2D<const Scalar>;
  using ScalarList = std::vector<Scalar>;
  using MaskScalar = unsigned char;
  using ImageMask = Image2D<const MaskScalar>;

  virtual ~ImgSimMetric2D() { xrc_sm_destroy(sm_); }
  ImgSimMetric2D(const ImgSimMetric2D&) = delete;
  ImgSimMetric2D& operator=(const ImgSimMetric2D&) = delete;

  /// the pixels are copied to the device here; the view need not outlive the call
  void set_fixed_image(const Image& fixed_img)
  {
    detail::Assert(bool(fixed_img), "fixed image");
    fixed_rows_ = fixed_img.rows;
    fixed_cols_ = fixed_img.cols;
    detail::Check(xrc_sm_set_fixed(sm_, fixed_img.data, static_cast<uint32_t>(fixed_img.rows),
                                   static_cast<uint32_t>(fixed_img.cols)));
  }

  /// xregImgSimMetric2D.cpp: sizes sim_vals_ too; re-callable after allocation with n <= capacity
  void set_num_moving_images(const size_type n)
  {
    num_mov_imgs_ = n;
    if (allocated_)
    {
      detail::Check(xrc_sm_set_num_imgs(sm_, static_cast<uint32_t>(n)));
    }
    sim_vals_.assign(n, 0);
  }

  size_type num_moving_images() const { return num_mov_imgs_; }

  virtual void allocate_resources()
  {
    detail::Assert(num_mov_imgs_ != 0, "num_mov_imgs_");
    pre_allocate();
    detail::Check(xrc_sm_allocate(sm_, static_cast<uint32_t>(num_mov_imgs_)));
    sim_vals_.assign(num_mov_imgs_, 0);
    allocated_ = true;
  }

  /// blocking: the similarity values are valid on return
  virtual void compute()
  {
    detail::Assert(allocated_, "resources allocated");
    if (rc_)
    {
      rc_->flush();
    }
    detail::Check(xrc_sm_compute(sm_));
    if (num_mov_imgs_)
    {
      detail::Check(xrc_sm_read_sims(sm_, sim_vals_.data(), static_cast<uint32_t>(num_mov_imgs_)));
    }
  }

  Scalar& sim_val(const size_type mov_img_idx) { return sim_vals_.at(mov_img_idx); }
  const Scalar& sim_val(const size_type mov_img_idx) const { return sim_vals_.at(mov_img_idx); }
  ScalarList& sim_vals() { return sim_vals_; }
  const ScalarList& sim_vals() const { return sim_vals_; }

  /// zero-copy device hand-off; re-callable with a new offset, a different ray caster is an error
  /// (xregImgSimMetric2DCPU.cpp:45-70)
  virtual void set_mov_imgs_buf_from_ray_caster(RayCaster* ray_caster, const size_type proj_offset = 0)
  {
    auto* rc = dynamic_cast<RayCasterLineIntCUDA*>(ray_caster);
    if (!rc)
    {
      throw UnsupportedOperationException("set_mov_imgs_buf_from_ray_caster: not a CUDA ray caster");
    }
    detail::Check(xrc_sm_bind_ray_caster(sm_, rc->handle(), static_cast<uint32_t>(proj_offset)));
    rc_ = rc;
  }

  /// caller-owned host images, copied to the device at every compute(); must outlive the metric
  virtual void set_mov_imgs_host_buf(Scalar* mov_imgs_buf, const size_type proj_offset = 0)
  {
    detail::Check(xrc_sm_bind_host(sm_, mov_imgs_buf, static_cast<uint32_t>(proj_offset)));
    rc_ = nullptr;
  }

  /// uint8 mask of the fixed image's size, or an empty view to remove it; may be changed between computes
  void set_mask(const ImageMask& mask)
  {
    if (mask)
    {
      detail::Assert(mask.rows == fixed_rows_ && mask.cols == fixed_cols_, "mask size == fixed image size");
      mask_.assign(mask.data, mask.data + mask.rows * mask.cols);
      detail::Check(xrc_sm_set_mask(sm_, mask_.data()));
    }
    else
    {
      mask_.clear();
      detail::Check(xrc_sm_set_mask(sm_, nullptr));
    }
    mask_changed();
  }

  bool has_mask() const { return !mask_.empty(); }

  size_type num_pix_per_proj() const { return fixed_rows_ * fixed_cols_; }

  xrc_sm* handle() { return sm_; }

protected:
  friend class Intensity2D3DObjFn;

  ImgSimMetric2D(Context& ctx, const int kind) { detail::Check(xrc_sm_create(ctx.handle(), kind, &sm_)); }

  virtual void pre_allocate() {}
  virtual void mask_changed() {}

  xrc_sm* sm_ = nullptr;
  RayCasterLineIntCUDA* rc_ = nullptr;
  size_type fixed_rows_ = 0;
  size_type fixed_cols_ = 0;
  size_type num_mov_imgs_ = 0;
  ScalarList sim_vals_;
  std::vector<MaskScalar> mask_;
  bool allocated_ = false;
};

/// xregImgSimMetric2DGradImgParamInterface.h:31-39 -- the "radius" is the Gaussian kernel WIDTH
/// (cv::GaussianBlur ksize, xregImgSimMetric2DGradImgCPU.cpp:54-55), 0 = no smoothing, default 5
class ImgSimMetric2DGradImgParamInterface
{
public:
  virtual ~ImgSimMetric2DGradImgParamInterface() = default;
  virtual size_type smooth_img_before_sobel_kernel_radius() const = 0;
  virtual void set_smooth_img_before_sobel_kernel_radius(const size_type r) = 0;
};

/// ImgSimMetric2DPatchCommon (xregImgSimMetric2DPatchCommon.{h,cpp}): the patch grid and the per-patch
/// weights are host logic, exactly as in the reference; the library receives them as arrays.
class ImgSimMetric2DPatchCommon
{
public:
  using Scalar = ImgSimMetric2D::Scalar;
  using MaskScalar = ImgSimMetric2D::MaskScalar;
  using WgtImg = Image2D<const Scalar>;

  virtual ~ImgSimMetric2DPatchCommon() = default;

  size_type patch_radius() const { return patch_radius_; }
  void set_patch_radius(const size_type r) { patch_radius_ = r; }
  void set_patch_stride(const size_type s) { patch_stride_ = s; }
  size_type patch_stride() const { return patch_stride_; }
  void set_compute_mean_of_patch_sims(const bool b) { compute_mean_of_patch_sims_ = b; }
  bool compute_mean_of_patch_sims() const { return compute_mean_of_patch_sims_; }
  void set_weight_patch_sims_in_combine(const bool b) { weight_patch_sims_in_combine_ = b; }
  bool weight_patch_sims_in_combine() const { return weight_patch_sims_in_combine_; }
  void set_use_mask_for_patch_weighting(const bool u) { use_mask_for_weighting_ = u; }
  bool use_mask_for_patch_weighting() const { return use_mask_for_weighting_; }
  void set_use_mask_for_patch_stats(const bool u) { use_mask_for_patch_stats_ = u; }
  bool use_mask_for_patch_stats() const { return use_mask_for_patch_stats_; }
  void set_normalize_weights_as_prob(const bool n) { normalize_weights_as_prob_ = n; }
  bool normalize_weights_as_prob() const { return normalize_weights_as_prob_; }

  void set_choose_rand_patches(const bool b)
  {
    if (b)
    {
      throw UnsupportedOperationException("random patch subsets are not supported by the CUDA metrics");
    }
  }

  bool choose_rand_patches() const { return false; }

  /// weight image of the fixed image's size (copied)
  void set_wgt_img(const WgtImg& w)
  {
    if (w)
    {
      wgt_img_.assign(w.data, w.data + w.rows * w.cols);
      wgt_rows_ = w.rows;
      wgt_cols_ = w.cols;
    }
    else
    {
      wgt_img_.clear();
    }
    weights_changed();
  }

  /// number of patches of the grid (xregImgSimMetric2DPatchCommon.cpp:269-292)
  static size_type NumPatches(const size_type rows, const size_type cols, const size_type r, const size_type s)
  {
    const size_type d = 2 * r + 1;
    if (d > rows || d > cols || !s)
    {
      return 0;
    }
    return ((rows - d) / s + 1) * ((cols - d) / s + 1);
  }

  /// compute_weights (xregImgSimMetric2DPatchCommon.cpp:309-410).  Returns false (weights untouched) when every
  /// weight stays 1; otherwise one weight per patch in row-major centre order.
  bool compute_weights(const size_type rows, const size_type cols, const MaskScalar* mask, std::vector<Scalar>* wgts) const
  {
    const bool use_mask_wgts = use_mask_for_weighting_ && mask;
    const bool use_img_wgts = !wgt_img_.empty();
    if (!(use_img_wgts || use_mask_wgts))
    {
      return false;
    }
    const size_type r = patch_radius_, s = patch_stride_, d = 2 * r + 1;
    detail::Assert(d <= rows && d <= cols && s, "patch_diam_ <= image size");
    if (use_img_wgts)
    {
      detail::Assert(wgt_rows_ == rows && wgt_cols_ == cols, "weight image size == fixed image size");
    }
    wgts->clear();
    wgts->reserve(NumPatches(rows, cols, r, s));
    // integral image of the mask: a (2r+1)^2 count per patch without the reference's d^2 loop (same integers)
    std::vector<uint32_t> ii;
    if (!use_img_wgts)
    {
      ii.assign((rows + 1) * (cols + 1), 0);
      for (size_type y = 0; y < rows; ++y)
      {
        uint32_t run = 0;
        for (size_type x = 0; x < cols; ++x)
        {
          run += mask[y * cols + x] ? 1u : 0u;
          ii[(y + 1) * (cols + 1) + (x + 1)] = ii[y * (cols + 1) + (x + 1)] + run;
        }
      }
    }
    for (size_type cr = r; cr + r <= rows - 1; cr += s)
    {
      for (size_type cc = r; cc + r <= cols - 1; cc += s)
      {
        Scalar w;
        if (use_img_wgts)
        {
          w = wgt_img_[cr * cols + cc];
          if (use_mask_wgts && !mask[cr * cols + cc])
          {
            w = 0;
          }
        }
        else
        {
          const size_type r0 = cr - r, c0 = cc - r, W = cols + 1;
          const uint32_t cnt = ii[(r0 + d) * W + (c0 + d)] - ii[r0 * W + (c0 + d)] - ii[(r0 + d) * W + c0] + ii[r0 * W + c0];
          w = static_cast<Scalar>(cnt) / static_cast<Scalar>(d * d);
        }
        wgts->push_back(w);
      }
    }
    if (normalize_weights_as_prob_)
    {
      Scalar wgt_sum = 0;
      for (const Scalar w : *wgts)
      {
        wgt_sum += w;
      }
      for (Scalar& w : *wgts)
      {
        w /= wgt_sum;
      }
    }
    return true;
  }

protected:
  virtual void weights_changed() {}

  void push_patch_params(xrc_sm* sm, const size_type rows, const size_type cols, const MaskScalar* mask)
  {
    std::vector<Scalar> w;
    const bool have = compute_weights(rows, cols, mask, &w);
    detail::Check(xrc_sm_set_patch_params(sm, static_cast<uint32_t>(patch_radius_), static_cast<uint32_t>(patch_stride_),
                                          compute_mean_of_patch_sims_ ? 1 : 0, weight_patch_sims_in_combine_ ? 1 : 0,
                                          use_mask_for_patch_stats_ ? 1 : 0, have ? w.data() : nullptr,
                                          have ? static_cast<uint64_t>(w.size()) : 0));
  }

  size_type patch_radius_ = 5;  // xregImgSimMetric2DPatchCommon.h:141-166
  size_type patch_stride_ = 1;
  bool compute_mean_of_patch_sims_ = false;
  bool weight_patch_sims_in_combine_ = true;
  bool use_mask_for_weighting_ = true;
  bool use_mask_for_patch_stats_ = false;
  bool normalize_weights_as_prob_ = true;
  std::vector<Scalar> wgt_img_;
  size_type wgt_rows_ = 0;
  size_type wgt_cols_ = 0;
};

/// ImgSimMetric2DSSDCPU / OCL (xregImgSimMetric2DSSDCPU.cpp:62-110)
class ImgSimMetric2DSSDCUDA : public ImgSimMetric2D
{
public:
  explicit ImgSimMetric2DSSDCUDA(Context& ctx) : ImgSimMetric2D(ctx, XRC_SM_SSD) {}
};

/// ImgSimMetric2DNCCCPU / OCL (xregImgSimMetric2DNCCCPU.cpp:52-236).  The moving-image buffer is left
/// untouched (the CPU class overwrites it with the zero-mean images).
class ImgSimMetric2DNCCCUDA : public ImgSimMetric2D
{
public:
  explicit ImgSimMetric2DNCCCUDA(Context& ctx) : ImgSimMetric2D(ctx, XRC_SM_NCC) {}
};

/// ImgSimMetric2DGradNCCCPU / OCL (xregImgSimMetric2DGradNCCCPU.cpp:29-65)
class ImgSimMetric2DGradNCCCUDA : public ImgSimMetric2D, public ImgSimMetric2DGradImgParamInterface
{
public:
  explicit ImgSimMetric2DGradNCCCUDA(Context& ctx) : ImgSimMetric2D(ctx, XRC_SM_GRAD_NCC) {}

  size_type smooth_img_before_sobel_kernel_radius() const override { return smooth_img_kernel_rad_; }

  void set_smooth_img_before_sobel_kernel_radius(const size_type r) override
  {
    smooth_img_kernel_rad_ = r;
    detail::Check(xrc_sm_set_grad_params(sm_, static_cast<uint32_t>(r)));
  }

  void read_grads(const size_type img, float* gx, float* gy)
  {
    detail::Check(xrc_sm_read_grads(sm_, static_cast<uint32_t>(img), gx, gy));
  }

private:
  size_type smooth_img_kernel_rad_ = 5;  // xregImgSimMetric2DGradImgCPU.h:81
};

/// ImgSimMetric2DPatchNCCCPU / OCL (xregImgSimMetric2DPatchNCCCPU.cpp:74-300,332-441,558-619)
class ImgSimMetric2DPatchNCCCUDA : public ImgSimMetric2D, public ImgSimMetric2DPatchCommon
{
public:
  explicit ImgSimMetric2DPatchNCCCUDA(Context& ctx) : ImgSimMetric2D(ctx, XRC_SM_PATCH_NCC) {}

  /// XRC_COMBINE_REFERENCE (default: the reference's sequential f32 sum of the per-patch values, bit for bit),
  /// XRC_COMBINE_REFERENCE_SERIAL (literal loop, verification) or XRC_COMBINE_F64 (double; closest to exact)
  void set_combine_mode(const int mode) { detail::Check(xrc_sm_set_combine_mode(sm_, mode)); }

protected:
  ImgSimMetric2DPatchNCCCUDA(Context& ctx, const int kind) : ImgSimMetric2D(ctx, kind) {}

  void pre_allocate() override { push(); }

  void mask_changed() override
  {
    if (allocated_)
    {
      push();
    }
  }

  void weights_changed() override
  {
    if (allocated_)
    {
      push();
    }
  }

private:
  void push() { push_patch_params(sm_, fixed_rows_, fixed_cols_, mask_.empty() ? nullptr : mask_.data()); }
};

/// ImgSimMetric2DPatchGradNCCCPU / OCL (xregImgSimMetric2DPatchGradNCCCPU.cpp:34-253)
class ImgSimMetric2DPatchGradNCCCUDA : public ImgSimMetric2DPatchNCCCUDA, public ImgSimMetric2DGradImgParamInterface
{
public:
  explicit ImgSimMetric2DPatchGradNCCCUDA(Context& ctx) : ImgSimMetric2DPatchNCCCUDA(ctx, XRC_SM_PATCH_GRAD_NCC) {}

  size_type smooth_img_before_sobel_kernel_radius() const override { return smooth_img_kernel_rad_; }

  void set_smooth_img_before_sobel_kernel_radius(const size_type r) override
  {
    smooth_img_kernel_rad_ = r;
    detail::Check(xrc_sm_set_grad_params(sm_, static_cast<uint32_t>(r)));
  }

  void read_grads(const size_type img, float* gx, float* gy)
  {
    detail::Check(xrc_sm_read_grads(sm_, static_cast<uint32_t>(img), gx, gy));
  }

private:
  size_type smooth_img_kernel_rad_ = 5;
};

/// ImgSimMetric2DCombine (xregImgSimMetric2DCombine.{h,cpp}): host combination of the views' values
class ImgSimMetric2DCombine
{
public:
  using Scalar = ImgSimMetric2D::Scalar;
  using ScalarList = ImgSimMetric2D::ScalarList;

  ImgSimMetric2DCombine() = default;
  virtual ~ImgSimMetric2DCombine() = default;
  ImgSimMetric2DCombine(const ImgSimMetric2DCombine&) = delete;
  ImgSimMetric2DCombine& operator=(const ImgSimMetric2DCombine&) = delete;

  void allocate_resources()
  {
    sim_objs_.resize(num_sim_metrics_);
    sim_vals_.resize(num_projs_per_sim_metric_);
  }

  void set_num_sim_metrics(const size_type n) { num_sim_metrics_ = n; }
  void set_num_projs_per_sim_metric(const size_type n) { num_projs_per_sim_metric_ = n; }

  void set_sim_metric(const size_type sim_idx, ImgSimMetric2D* sim)
  {
    sim_objs_.at(sim_idx) = sim;
    detail::Assert(sim->num_moving_images() == num_projs_per_sim_metric_, "sim->num_moving_images() == num_projs_per_sim_metric_");
  }

  const ScalarList& sim_vals() const { return sim_vals_; }

  Scalar sim_val(const size_type proj_idx) const
  {
    detail::Assert(proj_idx < num_projs_per_sim_metric_, "proj_idx < num_projs_per_sim_metric_");
    return sim_vals_[proj_idx];
  }

  virtual void compute() = 0;

protected:
  void add_all()
  {
    sim_vals_.assign(num_projs_per_sim_metric_, 0);
    for (size_type s = 0; s < num_sim_metrics_; ++s)
    {
      for (size_type p = 0; p < num_projs_per_sim_metric_; ++p)
      {
        sim_vals_[p] += sim_objs_[s]->sim_val(p);
      }
    }
  }

  size_type num_sim_metrics_ = 0;
  size_type num_projs_per_sim_metric_ = 0;
  std::vector<ImgSimMetric2D*> sim_objs_;
  ScalarList sim_vals_;
};

class ImgSimMetric2DCombineAddition : public ImgSimMetric2DCombine
{
public:
  void compute() override { add_all(); }
};

/// xregImgSimMetric2DCombine.cpp:67-86
class ImgSimMetric2DCombineMean : public ImgSimMetric2DCombine
{
public:
  void compute() override
  {
    add_all();
    for (Scalar& v : sim_vals_)
    {
      v /= static_cast<Scalar>(num_sim_metrics_);
    }
  }
};

/// What Intensity2D3DRegi::setup + obj_fn do around the two interfaces for one moving volume
/// (lib/regi/interfaces_2d_3d/xregIntensity2D3DRegi.cpp:43-133,571-696): view-major projection buffer,
/// metric v bound at offset v * pop, one library call per objective evaluation (xrc_obj_fn).
class Intensity2D3DObjFn
{
public:
  /// the ray caster must have its volumes and cameras set; the metrics their fixed images, masks and
  /// parameters.  Sizes and allocates everything for populations of up to max_pop poses.
  Intensity2D3DObjFn(RayCasterLineIntCUDA* rc, const std::vector<ImgSimMetric2D*>& sims, const size_type max_pop)
      : rc_(rc), sims_(sims), max_pop_(max_pop)
  {
    detail::Assert(rc->num_camera_models() == sims.size(), "one similarity metric per view");
    rc_->set_num_projs(max_pop * sims.size());
    rc_->allocate_resources();
    for (size_type v = 0; v < sims.size(); ++v)
    {
      sims_[v]->set_num_moving_images(max_pop);
      sims_[v]->set_mov_imgs_buf_from_ray_caster(rc_, max_pop * v);
      sims_[v]->allocate_resources();
      handles_.push_back(sims_[v]->handle());
    }
  }

  /// poses: cam -> volume physical, one per population member; returns the mean over views per pose
  const std::vector<float>& operator()(const FrameTransformList& poses, const size_type vol_idx = 0)
  {
    const size_type n = poses.size();
    detail::Assert(n <= max_pop_, "population <= allocated capacity");
    out_.assign(n, 0);
    if (!n)
    {
      return out_;
    }
    tmp_.resize(12 * n);
    for (size_type i = 0; i < n; ++i)
    {
      poses[i].to3x4(&tmp_[12 * i]);
    }
    per_view_.resize(n * sims_.size());
    rc_->flush_params();
    // the library sizes the ray caster and the metrics for this population and distributes the poses
    detail::Check(xrc_obj_fn(rc_->handle(), static_cast<uint32_t>(vol_idx), handles_.data(), static_cast<uint32_t>(sims_.size()),
                             static_cast<uint32_t>(n), tmp_.data(), out_.data(), per_view_.data()));
    rc_->library_resized(n * sims_.size());
    for (size_type v = 0; v < sims_.size(); ++v)
    {
      sims_[v]->num_mov_imgs_ = n;
      sims_[v]->sim_vals_.assign(per_view_.begin() + v * n, per_view_.begin() + (v + 1) * n);
    }
    return out_;
  }

  /// per-view values of the last evaluation, view-major
  const std::vector<float>& per_view() const { return per_view_; }

private:
  RayCasterLineIntCUDA* rc_;
  std::vector<ImgSimMetric2D*> sims_;
  size_type max_pop_;
  std::vector<xrc_sm*> handles_;
  std::vector<float> tmp_, out_, per_view_;
};

// ---- projection pre-processing ------------------------------------------------------------------
/// xreg::ImageIntensLogTransFilter (lib/image/xregImageIntensLogTrans.{h,cpp}) as one call: SetNormalizeZeroOne /
/// SetUseMaxIntensityAsI0 / SetI0 are the arguments, Update() is the call.  In place when dst == src.  Returns the I0 used.
inline float LogRemap(Context& ctx, const Image2D<const float>& src, float* dst, const bool normalize_zero_one = false,
                      const bool use_max_intensity_as_I0 = true, const float I0 = 1.0f)
{
  detail::Assert(bool(src) && dst, "LogRemap: null image");
  float used = 0.0f;
  detail::Check(xrc_log_remap(ctx.handle(), src.data, static_cast<uint32_t>(src.rows), static_cast<uint32_t>(src.cols),
                              normalize_zero_one ? 1 : 0, use_max_intensity_as_I0 ? 1 : 0, I0, dst, &used));
  return used;
}

/// xreg::DownsampleImage (lib/itk/xregITKResampleUtils.h:49-112, cubic B-spline default): the image comes back in `dst`
/// (resized), its size in rows / cols.  sigma < 0: the default smoothing 0.5 / factor.
inline void DownsampleImage(Context& ctx, const Image2D<const float>& src, const double factor, std::vector<float>* dst,
                            size_type* rows, size_type* cols, const double sigma = -1.0)
{
  detail::Assert(bool(src) && dst && rows && cols, "DownsampleImage: null argument");
  uint32_t r = 0, c = 0;
  detail::Check(xrc_downsample_size(static_cast<uint32_t>(src.rows), static_cast<uint32_t>(src.cols), factor, &r, &c));
  dst->assign(static_cast<size_t>(r) * c, 0.0f);
  detail::Check(xrc_downsample_image(ctx.handle(), src.data, static_cast<uint32_t>(src.rows), static_cast<uint32_t>(src.cols),
                                     factor, sigma, dst->data()));
  *rows = r;
  *cols = c;
}

}  // namespace xreg_b200

#endif
