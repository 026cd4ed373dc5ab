/* TEST INFRASTRUCTURE.  Functional stand-ins for the third-party types the reference's DRR path is written in, just
 * wide enough to compile the reference's OWN source lines for
 *   - RayRectIntersect                      (lib/spatial/xregSpatialPrimitives.cpp)
 *   - CameraModel::ind_pt_to_phys_det_pt    (lib/transforms/xregPerspectiveXform.cpp)
 *   - the line-integral kernels, LineIntParams and ComputeLineInts<Kernel>  (lib/ray_cast/xregRayCastLineIntCPU.cpp)
 * where they lie under /root/reference (oracle/ref_pin/build_ref_slice.py cuts those line ranges into a generated
 * translation unit under oracle/_ref/; nothing of the reference is copied into this repository).
 *
 * What this header restates is NOT the reference but its un-vendored dependencies, with the arithmetic conventions
 * DESIGN.md section 1 lists:
 *   Eigen 3.3.4 fixed-size float vectors / matrices / Transform<float,3,Affine>: element-wise ops in f32, products as
 *     left-to-right sums over the inner index without FMA (this file is compiled with -ffp-contract=off, no -march),
 *     norm() = sqrt(sum of squares, left to right), normalized() = v / norm(),
 *     Affine * Affine = affine(3x4) * matrix(4x4) coefficient-wise (the 4th term multiplies the constant last row),
 *     Affine * vector = linear * v + translation;
 *   ITK 5.1.1 Image<float,3>, ContinuousIndex, Vector and LinearInterpolateImageFunction::EvaluateOptimized(Dispatch<3>)
 *     (f64 lerps with f32 distances, base index clamped to the start index, neighbours beyond the end index dropped);
 *   TBB: ParallelFor runs the body once over the whole range (the reference's own XREG_NO_TBB fallback does the same).
 * The test that uses the resulting library (tests/test_oracle_ref_slice.py) therefore pins the oracle's restatement of
 * the REFERENCE's control flow and expression order (clip test, nudge, step length, step count, sample loop, kernels,
 * scaling, store) bit for bit, under these stated conventions for the libraries underneath.
 */
#ifndef XREG_REF_PIN_PRELUDE_H
#define XREG_REF_PIN_PRELUDE_H

#include <algorithm>
#include <array>
#include <cassert>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <limits>
#include <memory>
#include <random>
#include <stdexcept>
#include <tuple>
#include <vector>

namespace Eigen
{

template <int N>
struct Vec
{
  float v[N];
  Vec()
  {
    for (int i = 0; i < N; ++i)
      v[i] = 0.0f;
  }
  Vec(float a, float b)   /* Pt2{col, row} (xregRayCastDepthCPU.cpp:126) */
  {
    static_assert(N == 2, "two-component initialiser");
    v[0] = a;
    v[1] = b;
  }
  static Vec Zero() { return Vec(); }
  float& operator()(int i) { return v[i]; }
  const float& operator()(int i) const { return v[i]; }
  float& operator[](int i) { return v[i]; }
  const float& operator[](int i) const { return v[i]; }
  Vec& operator+=(const Vec& o)
  {
    for (int i = 0; i < N; ++i)
      v[i] = v[i] + o.v[i];
    return *this;
  }
  float squaredNorm() const
  {
    float s = v[0] * v[0];
    for (int i = 1; i < N; ++i)
      s = s + (v[i] * v[i]);
    return s;
  }
  float norm() const { return std::sqrt(squaredNorm()); }
  Vec normalized() const
  {
    const float n = norm();
    Vec r;
    for (int i = 0; i < N; ++i)
      r.v[i] = v[i] / n;
    return r;
  }
};

template <int N>
inline Vec<N> operator+(const Vec<N>& a, const Vec<N>& b)
{
  Vec<N> r;
  for (int i = 0; i < N; ++i)
    r.v[i] = a.v[i] + b.v[i];
  return r;
}
template <int N>
inline Vec<N> operator-(const Vec<N>& a, const Vec<N>& b)
{
  Vec<N> r;
  for (int i = 0; i < N; ++i)
    r.v[i] = a.v[i] - b.v[i];
  return r;
}
template <int N>
inline Vec<N> operator*(float s, const Vec<N>& a)
{
  Vec<N> r;
  for (int i = 0; i < N; ++i)
    r.v[i] = s * a.v[i];
  return r;
}
template <int N>
inline Vec<N> operator*(const Vec<N>& a, float s)
{
  Vec<N> r;
  for (int i = 0; i < N; ++i)
    r.v[i] = a.v[i] * s;
  return r;
}

struct Mat3
{
  float m[3][3];
  Mat3()
  {
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j)
        m[i][j] = (i == j) ? 1.0f : 0.0f;
  }
  float& operator()(int i, int j) { return m[i][j]; }
  const float& operator()(int i, int j) const { return m[i][j]; }
};
inline Vec<3> operator*(const Mat3& a, const Vec<3>& x)
{
  Vec<3> r;
  for (int i = 0; i < 3; ++i)
    r.v[i] = ((a.m[i][0] * x.v[0]) + (a.m[i][1] * x.v[1])) + (a.m[i][2] * x.v[2]);
  return r;
}

/* the 4x4 behind a Transform; block(0,0,3,3) is the only accessor the slices use */
struct Mat4View
{
  const float (*m)[4];
  Mat3 block(int r0, int c0, int nr, int nc) const
  {
    assert(r0 == 0 && c0 == 0 && nr == 3 && nc == 3);
    (void)r0, (void)c0, (void)nr, (void)nc;
    Mat3 b;
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j)
        b.m[i][j] = m[i][j];
    return b;
  }
};

struct Affine3
{
  float m[4][4];
  Affine3()
  {
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 4; ++j)
        m[i][j] = (i == j) ? 1.0f : 0.0f;
  }
  static Affine3 Identity() { return Affine3(); }
  Mat4View matrix() const { return Mat4View{m}; }
  /* Transform<float,3,Affine>::inverse(): linear^-1 by cofactors (Eigen 3.3 compute_inverse_size3 shape: inv(i,j) =
   * cof<j,i> / det, det = (cof<0,0> m00 + cof<1,0> m10) + cof<2,0> m20) and -(linear^-1) t -- third-party arithmetic,
   * the convention the oracle states too (xo_affine_inverse); used by the depth ray caster's slice only */
  Affine3 inverse() const
  {
    auto cof = [this](int i, int j) {
      const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
      return (m[i1][j1] * m[i2][j2]) - (m[i1][j2] * m[i2][j1]);
    };
    const float c0 = cof(0, 0), c1 = cof(1, 0), c2 = cof(2, 0);
    const float det = ((c0 * m[0][0]) + (c1 * m[1][0])) + (c2 * m[2][0]);
    const float invdet = 1.0f / det;
    Affine3 r;
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j)
        r.m[i][j] = cof(j, i) * invdet;
    for (int i = 0; i < 3; ++i)
      r.m[i][3] = -(((r.m[i][0] * m[0][3]) + (r.m[i][1] * m[1][3])) + (r.m[i][2] * m[2][3]));
    return r;
  }
};
/* Transform * Transform, Affine mode: res.affine() = lhs.affine() * rhs.matrix() (sum over all four inner indices, in
 * order), last row copied */
inline Affine3 operator*(const Affine3& a, const Affine3& b)
{
  Affine3 r;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 4; ++j)
      r.m[i][j] = (((a.m[i][0] * b.m[0][j]) + (a.m[i][1] * b.m[1][j])) + (a.m[i][2] * b.m[2][j])) + (a.m[i][3] * b.m[3][j]);
  for (int j = 0; j < 4; ++j)
    r.m[3][j] = b.m[3][j];
  return r;
}
/* Transform * vector, Affine mode: linear * v + translation */
inline Vec<3> operator*(const Affine3& a, const Vec<3>& x)
{
  Vec<3> r;
  for (int i = 0; i < 3; ++i)
    r.v[i] = (((a.m[i][0] * x.v[0]) + (a.m[i][1] * x.v[1])) + (a.m[i][2] * x.v[2])) + a.m[i][3];
  return r;
}

}  // namespace Eigen

namespace itk
{

template <class T, unsigned N>
struct Image
{
  const T* data = nullptr;
  std::size_t size[N];
  T GetPixelAt(std::size_t i, std::size_t j, std::size_t k) const { return data[i + size[0] * (j + size[1] * k)]; }
};

template <class T, unsigned N>
struct Vector
{
  T v[N];
  T& operator[](unsigned i) { return v[i]; }
  const T& operator[](unsigned i) const { return v[i]; }
  Vector& operator*=(const T& s)   /* itk::Vector::operator*=(const ValueType&): the scalar is converted to T first */
  {
    for (unsigned i = 0; i < N; ++i)
      v[i] = static_cast<T>(v[i] * s);
    return *this;
  }
  Vector operator-() const
  {
    Vector r;
    for (unsigned i = 0; i < N; ++i)
      r.v[i] = -v[i];
    return r;
  }
};

template <class T, unsigned N>
struct ContinuousIndex
{
  T v[N];
  T& operator[](unsigned i) { return v[i]; }
  const T& operator[](unsigned i) const { return v[i]; }
  ContinuousIndex& operator+=(const Vector<T, N>& d)
  {
    for (unsigned i = 0; i < N; ++i)
      v[i] = v[i] + d.v[i];
    return *this;
  }
  ContinuousIndex& operator-=(const Vector<T, N>& d)
  {
    for (unsigned i = 0; i < N; ++i)
      v[i] = v[i] - d.v[i];
    return *this;
  }
};

namespace Function
{
template <unsigned R>
struct LanczosWindowFunction
{
};
}  // namespace Function
template <class I>
struct ConstantBoundaryCondition
{
};

/* smart-pointer stand-in: New() returns one, assignment between related function types works */
template <class F>
struct FnPtr
{
  std::shared_ptr<F> p;
  FnPtr() {}
  template <class G>
  FnPtr(const FnPtr<G>& o) : p(o.p)
  {
  }
  template <class G>
  FnPtr& operator=(const FnPtr<G>& o)
  {
    p = o.p;
    return *this;
  }
  F* operator->() const { return p.get(); }
};

template <class TImage, class TCoord>
struct InterpolateImageFunction
{
  using Pointer = FnPtr<InterpolateImageFunction>;
  using ContinuousIndexType = ContinuousIndex<TCoord, 3>;
  const TImage* img = nullptr;
  virtual ~InterpolateImageFunction() {}
  void SetInputImage(const TImage* i) { img = i; }
  virtual double EvaluateAtContinuousIndex(const ContinuousIndexType& x) const = 0;
};

/* itk::LinearInterpolateImageFunction<Image<float,3>,float>::EvaluateOptimized(Dispatch<3>, index), ITK 5.1.1 */
template <class TImage, class TCoord>
struct LinearInterpolateImageFunction : InterpolateImageFunction<TImage, TCoord>
{
  using Base = InterpolateImageFunction<TImage, TCoord>;
  using Pointer = FnPtr<LinearInterpolateImageFunction>;
  static Pointer New()
  {
    Pointer p;
    p.p = std::make_shared<LinearInterpolateImageFunction>();
    return p;
  }
  double EvaluateAtContinuousIndex(const typename Base::ContinuousIndexType& x) const override
  {
    const TImage& im = *this->img;
    long b[3], n[3];
    double d[3];
    for (int k = 0; k < 3; ++k)
    {
      const long end = (long)im.size[k] - 1;
      long bk = (long)std::floor(x[k]);   /* Math::Floor<IndexValueType>(index[k]) */
      if (bk < 0)
        bk = 0;                           /* basei[k] < m_StartIndex[k] -> m_StartIndex[k] */
      if (bk > end)
        bk = end;                         /* never read out of bounds (the ray caster's nudge keeps ITK inside) */
      TCoord dist = x[k] - static_cast<TCoord>(bk);   /* distance in the coordinate type (float) */
      long nk = bk + 1;
      if (dist <= 0)
      {
        dist = 0;                         /* that axis is not interpolated */
        nk = bk;
      }
      if (nk > end)
      {
        nk = bk;                          /* neighbour beyond m_EndIndex dropped */
        dist = 0;
      }
      b[k] = bk;
      n[k] = nk;
      d[k] = dist;
    }
    const double v000 = im.GetPixelAt(b[0], b[1], b[2]), v100 = im.GetPixelAt(n[0], b[1], b[2]);
    const double v010 = im.GetPixelAt(b[0], n[1], b[2]), v110 = im.GetPixelAt(n[0], n[1], b[2]);
    const double v001 = im.GetPixelAt(b[0], b[1], n[2]), v101 = im.GetPixelAt(n[0], b[1], n[2]);
    const double v011 = im.GetPixelAt(b[0], n[1], n[2]), v111 = im.GetPixelAt(n[0], n[1], n[2]);
    const double vx00 = v000 + (v100 - v000) * d[0];
    const double vx10 = v010 + (v110 - v010) * d[0];
    const double vxx0 = vx00 + (vx10 - vx00) * d[1];
    const double vx01 = v001 + (v101 - v001) * d[0];
    const double vx11 = v011 + (v111 - v011) * d[0];
    const double vxx1 = vx01 + (vx11 - vx01) * d[1];
    return vxx0 + (vxx1 - vxx0) * d[2];
  }
};

/* the other interpolators are outside the scope of this repository (the CUDA path rejects them as unsupported) */
template <class F, class Base>
struct UnsupportedInterp : Base
{
  using Pointer = FnPtr<F>;
  static Pointer New() { throw std::runtime_error("interpolation method outside the pinned path"); }
  double EvaluateAtContinuousIndex(const typename Base::ContinuousIndexType&) const override { return 0.0; }
};
/* itk::NearestNeighborInterpolateImageFunction<Image<float,3>,float>::EvaluateAtContinuousIndex, ITK 5.1.1: the pixel at
 * ConvertContinuousIndexToNearestIndex(x) = Math::RoundHalfIntegerUp per axis, stated as floor(x + 0.5) evaluated
 * exactly; the index is clamped (ITK would read out of bounds, the ray caster's nudge keeps it inside) */
template <class TImage, class TCoord>
struct NearestNeighborInterpolateImageFunction : InterpolateImageFunction<TImage, TCoord>
{
  using Base = InterpolateImageFunction<TImage, TCoord>;
  using Pointer = FnPtr<NearestNeighborInterpolateImageFunction>;
  static Pointer New()
  {
    Pointer p;
    p.p = std::make_shared<NearestNeighborInterpolateImageFunction>();
    return p;
  }
  double EvaluateAtContinuousIndex(const typename Base::ContinuousIndexType& x) const override
  {
    const TImage& im = *this->img;
    long b[3];
    for (int k = 0; k < 3; ++k)
    {
      const long end = (long)im.size[k] - 1;
      long bk = (long)std::floor(static_cast<double>(x[k]) + 0.5);
      if (bk < 0)
        bk = 0;
      if (bk > end)
        bk = end;
      b[k] = bk;
    }
    return im.GetPixelAt(b[0], b[1], b[2]);
  }
};
template <class TImage, class TCoord = double>
struct BSplineInterpolateImageFunction
  : UnsupportedInterp<BSplineInterpolateImageFunction<TImage, TCoord>, InterpolateImageFunction<TImage, TCoord>>
{
  void SetSplineOrder(unsigned) {}
};
template <class TImage, unsigned R, class W, class B, class TCoord>
struct WindowedSincInterpolateImageFunction
  : UnsupportedInterp<WindowedSincInterpolateImageFunction<TImage, R, W, B, TCoord>, InterpolateImageFunction<TImage, TCoord>>
{
};

}  // namespace itk

/* what the slices need from the reference's own headers: type names and data members only
 * (xregCommon.h:42-93, xregPerspectiveXform.h:108-173, xregRayCastInterface.h:43-77, xregRayCastBaseCPU.h:37,
 * xregTBBUtils.h:69-101) */
#define xregASSERT(x) assert(x)

namespace xreg
{

using size_type = std::size_t;
using CoordScalar = float;
using Pt2 = Eigen::Vec<2>;
using Pt3 = Eigen::Vec<3>;
using Mat3x3 = Eigen::Mat3;
using FrameTransform = Eigen::Affine3;
using FrameTransformList = std::vector<FrameTransform>;

struct CameraModel
{
  enum CameraCoordFrame
  {
    kORIGIN_AT_FOCAL_PT_DET_POS_Z,
    kORIGIN_AT_FOCAL_PT_DET_NEG_Z,
    kORIGIN_ON_DETECTOR
  };
  Mat3x3 intrins_inv;
  FrameTransform extrins_inv;
  Pt3 pinhole_pt;
  CoordScalar focal_len = 0;
  size_type num_det_rows = 0;
  size_type num_det_cols = 0;
  CameraCoordFrame coord_frame_type = kORIGIN_AT_FOCAL_PT_DET_NEG_Z;

  Pt3 ind_pt_to_phys_det_pt(const Pt2& ind_pt) const;
  Pt3 ind_pt_to_phys_det_pt(const Pt3& ind_pt) const;
};

/* a background projection: only GetBufferPointer() is used (through the smart pointer) */
struct Proj2D
{
  float* buf = nullptr;
  float* GetBufferPointer() { return buf; }
  struct Pointer
  {
    Proj2D* p = nullptr;
    Proj2D* operator->() const { return p; }
  };
};

/* RayCaster / RayCasterCPU: type names, enums and the data members that distribute_xforms_among_cam_models
 * (xregRayCastInterface.cpp) and RayCasterCPU::pre_compute (xregRayCastBaseCPU.cpp) touch; their definitions are the
 * reference's lines */
struct RayCaster
{
  using PixelScalar2D = float;
  using PixelScalar3D = float;
  using Vol = itk::Image<float, 3>;
  using CameraModelList = std::vector<CameraModel>;
  using CamModelAssocList = std::vector<size_type>;
  using ProjList = std::vector<Proj2D::Pointer>;
  enum InterpMethod
  {
    kRAY_CAST_INTERP_LINEAR = 0,
    kRAY_CAST_INTERP_NN,
    kRAY_CAST_INTERP_SINC,
    kRAY_CAST_INTERP_BSPLINE
  };
  enum ProjPixelStoreMethod
  {
    kRAY_CAST_PIXEL_REPLACE = 0,
    kRAY_CAST_PIXEL_ACCUM
  };
  size_type num_camera_models() const { return camera_models_.size(); }
  size_type num_projs() const { return num_projs_; }
  void distribute_xforms_among_cam_models(const FrameTransformList& xforms_cam_to_itk_phys);

  CameraModelList camera_models_;
  FrameTransformList xforms_cam_to_itk_phys_;
  CamModelAssocList cam_model_for_proj_;
  size_type num_projs_ = 0;
  ProjPixelStoreMethod proj_store_meth_ = kRAY_CAST_PIXEL_REPLACE;
  bool use_bg_projs_ = false;
  ProjList bg_projs_for_each_cam_;
  PixelScalar2D default_bg_pixel_val_ = 0;
};

struct RayCasterCPU : RayCaster
{
  constexpr static CoordScalar kVOL_BB_STEP_INC_TOL = 1.0e-3;
  PixelScalar2D* buf_ = nullptr;
  PixelScalar2D* pixel_buf_to_use() { return buf_; }
  void pre_compute();
};

struct RayCasterLineIntCPU : RayCasterCPU
{
};
struct RayCasterDepthCPU : RayCasterCPU   /* the constant RayCastDepthFn names */
{
};

struct RangeType
{
  size_type begin_, end_;
  RangeType(const size_type b, const size_type e) : begin_(b), end_(e) {}
  size_type begin() const { return begin_; }
  size_type end() const { return end_; }
};

template <class Fn>
void ParallelFor(Fn& fn_obj, const RangeType& r)
{
  fn_obj(r);
}

std::tuple<bool, CoordScalar, CoordScalar> RayRectIntersect(const Pt3& min_rect_corner, const Pt3& max_rect_corner,
                                                            const Pt3& line_start_pt, const Pt3& line_vec,
                                                            const bool limit_to_segment);

}  // namespace xreg

#endif
