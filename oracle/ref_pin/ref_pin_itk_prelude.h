/* TEST INFRASTRUCTURE.  Stand-ins (accessors only, no arithmetic) for compiling the reference's OWN
 * ITKImageIndexBoundsAsEigen and ITKImagePhysicalPointTransformsAsEigen (lib/itk/xregITKBasicImageUtils.h): an itk::Image
 * that only carries its meta data (size, origin, spacing, direction as doubles) and the two Eigen types the functions
 * fill. */
#ifndef XREG_REF_PIN_ITK_PRELUDE_H
#define XREG_REF_PIN_ITK_PRELUDE_H

#include <cstddef>
#include <cstdint>
#include <tuple>

namespace Eigen
{
enum { Affine = 2 };
template <class T, int R, int C>
struct Matrix
{
  T v[R * C];
  T& operator[](int i) { return v[i]; }
  const T& operator[](int i) const { return v[i]; }
  T& operator()(int r, int c) { return v[r * C + c]; }
  const T& operator()(int r, int c) const { return v[r * C + c]; }
  void setIdentity()
  {
    for (int r = 0; r < R; ++r)
      for (int c = 0; c < C; ++c)
        v[r * C + c] = (r == c) ? T(1) : T(0);
  }
};
template <class T, int N, int Mode>
struct Transform
{
  using MatrixType = Matrix<T, N + 1, N + 1>;
  MatrixType m;
  static Transform Identity()
  {
    Transform t;
    t.m.setIdentity();
    return t;
  }
  MatrixType& matrix() { return m; }
  const MatrixType& matrix() const { return m; }
};
}  // namespace Eigen

namespace itk
{
template <class T, unsigned N>
struct Image
{
  struct SizeType
  {
    std::size_t s[N];
    std::size_t operator[](unsigned i) const { return s[i]; }
  };
  struct RegionType
  {
    SizeType sz;
    SizeType GetSize() const { return sz; }
  };
  struct PointType
  {
    double p[N];
    double operator[](unsigned i) const { return p[i]; }
  };
  using SpacingType = PointType;
  struct DirectionType
  {
    double d[N][N];
    double operator()(unsigned r, unsigned c) const { return d[r][c]; }
  };
  SizeType size;
  PointType origin;
  SpacingType spacing;
  DirectionType direction;
  RegionType GetLargestPossibleRegion() const { return RegionType{size}; }
  PointType GetOrigin() const { return origin; }
  SpacingType GetSpacing() const { return spacing; }
  DirectionType GetDirection() const { return direction; }
};
}  // namespace itk

namespace xreg
{
using CoordScalar = float;
}

#endif
