/* TEST INFRASTRUCTURE.  On top of ref_pin_se3_prelude.h: what the reference's OWN camera set-up lines need --
 * FocalLenFromIntrins, MakeNaiveIntrins, CameraModel::setup (naive and intrinsics + extrinsics), DownsampleCameraModel
 * (lib/transforms/xregPerspectiveXform.cpp) and SE3Inv (lib/transforms/xregRigidUtils.cpp).
 * Conventions added here (Eigen, un-vendored): Matrix3f::inverse() by cofactors with
 * det = (cof00 m00 + cof10 m10) + cof20 m20 and inv(i,j) = cof(j,i) * (1 / det) (Eigen 3.3 compute_inverse_size3 shape);
 * `-1 * A * b` evaluated left to right; Transform * vector = linear * v + translation.  The same conventions are stated in
 * xreg_oracle.c; what the comparison pins is the reference's control flow around them (frame types, which spacing goes
 * where, the double-precision principal point, rounding of the down-sampled size, the pinhole point). */
#ifndef XREG_REF_PIN_CAM_PRELUDE_H
#define XREG_REF_PIN_CAM_PRELUDE_H

#include "ref_pin_eigen_small.h"

#include <cstdint>
#include <cstdlib>
#include <tuple>

namespace Eigen
{
/* run-time sized temporary (<= 4x4) for arithmetic on block views */
struct Dyn
{
  float a[4][4];
  int r = 0, c = 0;
  template <int R, int C>
  operator M<R, C>() const
  {
    assert(R == r && C == c);
    M<R, C> m;
    for (int i = 0; i < R; ++i)
      for (int j = 0; j < C; ++j)
        m.a[i][j] = a[i][j];
    return m;
  }
};
template <class Blk>
inline Dyn dyn_of(const Blk& b)
{
  Dyn d;
  d.r = b.nr;
  d.c = b.nc;
  for (int i = 0; i < b.nr; ++i)
    for (int j = 0; j < b.nc; ++j)
      d.a[i][j] = b.cm->a[b.r0 + i][b.c0 + j];
  return d;
}
inline Dyn transpose_of(const Dyn& x)
{
  Dyn d;
  d.r = x.c;
  d.c = x.r;
  for (int i = 0; i < x.r; ++i)
    for (int j = 0; j < x.c; ++j)
      d.a[j][i] = x.a[i][j];
  return d;
}
inline Dyn operator*(float s, const Dyn& x)
{
  Dyn d = x;
  for (int i = 0; i < x.r; ++i)
    for (int j = 0; j < x.c; ++j)
      d.a[i][j] = s * x.a[i][j];
  return d;
}
inline Dyn operator*(const Dyn& x, const Dyn& y)
{
  assert(x.c == y.r);
  Dyn d;
  d.r = x.r;
  d.c = y.c;
  for (int i = 0; i < x.r; ++i)
    for (int j = 0; j < y.c; ++j)
    {
      float s = x.a[i][0] * y.a[0][j];
      for (int k = 1; k < x.c; ++k)
        s = s + x.a[i][k] * y.a[k][j];
      d.a[i][j] = s;
    }
  return d;
}

/* 4x4 with the block arithmetic SE3Inv uses */
struct M44 : M<4, 4>
{
  M44() {}
  M44(const M<4, 4>& m) : M<4, 4>(m) {}
  static M44 Identity() { return M44(M<4, 4>::Identity()); }
  struct Blk
  {
    M44* m;
    const M44* cm;
    int r0, c0, nr, nc;
    Dyn transpose() const { return transpose_of(dyn_of(*this)); }
    Blk& operator=(const Dyn& d)
    {
      assert(m && d.r == nr && d.c == nc);
      for (int i = 0; i < nr; ++i)
        for (int j = 0; j < nc; ++j)
          m->a[r0 + i][c0 + j] = d.a[i][j];
      return *this;
    }
    template <int R, int C>
    operator M<R, C>() const
    {
      return dyn_of(*this);
    }
  };
  Blk block(int r0, int c0, int nr, int nc) { return Blk{this, this, r0, c0, nr, nc}; }
  Blk block(int r0, int c0, int nr, int nc) const { return Blk{nullptr, this, r0, c0, nr, nc}; }
};
inline Dyn operator*(float s, const M44::Blk& b) { return s * dyn_of(b); }
inline Dyn operator*(const Dyn& x, const M44::Blk& b) { return x * dyn_of(b); }

struct M33 : M<3, 3>
{
  M33() {}
  M33(const M<3, 3>& m) : M<3, 3>(m) {}
  static M33 Identity() { return M33(M<3, 3>::Identity()); }
  M33 inverse() const
  {
    auto cof = [this](int i, int j) {
      const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
      return (a[i1][j1] * a[i2][j2]) - (a[i1][j2] * a[i2][j1]);
    };
    const float c0 = cof(0, 0), c1 = cof(1, 0), c2 = cof(2, 0);
    const float det = ((c0 * a[0][0]) + (c1 * a[1][0])) + (c2 * a[2][0]);
    const float invdet = 1.0f / det;
    M33 r;
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j)
        r.a[i][j] = cof(j, i) * invdet;
    return r;
  }
};
inline M<3, 1> operator*(const M33& x, const M<3, 1>& v) { return static_cast<const M<3, 3>&>(x) * v; }

struct Aff
{
  M44 mat = M44::Identity();
  static Aff Identity() { return Aff(); }
  M44& matrix() { return mat; }
  const M44& matrix() const { return mat; }
  Aff& operator=(const M44& m)
  {
    mat = m;
    return *this;
  }
};
inline M<3, 1> operator*(const Aff& t, const M<3, 1>& v)
{
  M<3, 1> r;
  for (int i = 0; i < 3; ++i)
    r.a[i][0] = (((t.mat.a[i][0] * v.a[0][0]) + (t.mat.a[i][1] * v.a[1][0])) + (t.mat.a[i][2] * v.a[2][0])) + t.mat.a[i][3];
  return r;
}
}  // namespace Eigen

namespace xreg
{
using size_type = std::size_t;
using CoordScalar = float;
using Pt3 = Eigen::M<3, 1>;
using Mat3x3 = Eigen::M33;
using Mat4x4 = Eigen::M44;
using FrameTransform = Eigen::Aff;

Mat4x4 SE3Inv(const Mat4x4& T);
CoordScalar FocalLenFromIntrins(const Mat3x3& K, CoordScalar xps, CoordScalar yps);
Mat3x3 MakeNaiveIntrins(const CoordScalar focal_len, const unsigned long num_rows, const unsigned long num_cols,
                        const CoordScalar pixel_row_spacing, const CoordScalar pixel_col_spacing, const bool z_is_neg);

/* xregPerspectiveXform.h:108-173: the data members */
struct CameraModel
{
  enum CameraCoordFrame
  {
    kORIGIN_AT_FOCAL_PT_DET_POS_Z,
    kORIGIN_AT_FOCAL_PT_DET_NEG_Z,
    kORIGIN_ON_DETECTOR
  };
  Mat3x3 intrins = Mat3x3(Mat3x3::Identity());
  Mat3x3 intrins_inv = Mat3x3(Mat3x3::Identity());
  FrameTransform extrins = FrameTransform::Identity();
  FrameTransform extrins_inv = FrameTransform::Identity();
  Pt3 pinhole_pt = Pt3(Pt3::Zero());
  CoordScalar focal_len = 0;
  size_type num_det_rows = 0;
  size_type num_det_cols = 0;
  CoordScalar det_row_spacing = 0;
  CoordScalar det_col_spacing = 0;
  CameraCoordFrame coord_frame_type = kORIGIN_AT_FOCAL_PT_DET_NEG_Z;

  void setup(const CoordScalar focal_len_arg, const size_type nr, const size_type nc, const CoordScalar rs,
             const CoordScalar cs);
  void setup(const Mat3x3& intrins_mat, const Mat4x4& extrins_mat, const size_type nr, const size_type nc,
             const CoordScalar rs, const CoordScalar cs);
};

CameraModel DownsampleCameraModel(const CameraModel& src_cam, const CoordScalar ds_factor, const bool force_even_dims);
}  // namespace xreg

#endif
