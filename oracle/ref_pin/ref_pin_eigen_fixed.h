/* TEST INFRASTRUCTURE.  Eigen::Matrix<T, R, C> for small fixed sizes (element access, difference, norm): the only Eigen
 * use of the patch-metric slices (distance between patch centres in the random-patch branch).  The NCC stand-in header
 * specialises the same template for dynamic row vectors. */
#ifndef XREG_REF_PIN_EIGEN_FIXED_H
#define XREG_REF_PIN_EIGEN_FIXED_H

#include <cmath>

namespace Eigen
{
/* only Matrix<double,2,1> is used (distance between patch centres in the random-patch branch) */
template <class T, int R, int C>
struct Matrix
{
  T v[R * C];
  T& operator[](int i) { return v[i]; }
  const T& operator[](int i) const { return v[i]; }
  Matrix operator-(const Matrix& o) const
  {
    Matrix r;
    for (int i = 0; i < R * C; ++i)
      r.v[i] = v[i] - o.v[i];
    return r;
  }
  T norm() const
  {
    T s = 0;
    for (int i = 0; i < R * C; ++i)
      s += v[i] * v[i];
    return std::sqrt(s);
  }
};
}  // namespace Eigen

#endif
