/* TEST INFRASTRUCTURE.  Eigen stand-in for small fixed-size float matrices / vectors with block views (element-wise f32,
 * products as left-to-right sums; compile with -ffp-contract=off).  Shared by ref_pin_se3_prelude.h and
 * ref_pin_cam_prelude.h. */
#ifndef XREG_REF_PIN_EIGEN_SMALL_H
#define XREG_REF_PIN_EIGEN_SMALL_H

#include <cassert>
#include <cmath>
#include <cstddef>

#ifndef xregASSERT
#define xregASSERT(x) assert(x)
#endif

namespace Eigen
{
template <int R, int C>
struct M
{
  float a[R][C];
  M()
  {
    for (int i = 0; i < R; ++i)
      for (int j = 0; j < C; ++j)
        a[i][j] = 0.0f;
  }
  static M Zero() { return M(); }
  static M Identity()
  {
    M m;
    for (int i = 0; i < (R < C ? R : C); ++i)
      m.a[i][i] = 1.0f;
    return m;
  }
  float& operator()(int i, int j) { return a[i][j]; }
  const float& operator()(int i, int j) const { return a[i][j]; }
  /* vectors (C == 1) */
  float& operator()(int i) { return a[i][0]; }
  const float& operator()(int i) const { return a[i][0]; }
  float& operator[](int i) { return a[i][0]; }
  const float& operator[](int i) const { return a[i][0]; }
  M<3, 1> head(int n) const
  {
    assert(n == 3 && C == 1);
    (void)n;
    M<3, 1> h;
    for (int i = 0; i < 3; ++i)
      h.a[i][0] = a[i][0];
    return h;
  }
  M<C, R> transpose() const
  {
    M<C, R> t;
    for (int i = 0; i < R; ++i)
      for (int j = 0; j < C; ++j)
        t.a[j][i] = a[i][j];
    return t;
  }
  float norm() const
  {
    float s = 0.0f;
    for (int j = 0; j < C; ++j)
      for (int i = 0; i < R; ++i)
        s = s + a[i][j] * a[i][j];
    return std::sqrt(s);
  }
  M& operator+=(const M& o)
  {
    for (int i = 0; i < R; ++i)
      for (int j = 0; j < C; ++j)
        a[i][j] = a[i][j] + o.a[i][j];
    return *this;
  }
  /* block views */
  template <int BR, int BC>
  struct Block
  {
    M* m;
    int r0, c0;
    Block& operator=(const M<BR, BC>& v)
    {
      for (int i = 0; i < BR; ++i)
        for (int j = 0; j < BC; ++j)
          m->a[r0 + i][c0 + j] = v.a[i][j];
      return *this;
    }
  };
  struct AnyBlock
  {
    M* m;
    const M* cm;
    int r0, c0, nr, nc;
    template <int BR, int BC>
    operator M<BR, BC>() const
    {
      assert(BR == nr && BC == nc);
      M<BR, BC> v;
      for (int i = 0; i < BR; ++i)
        for (int j = 0; j < BC; ++j)
          v.a[i][j] = cm->a[r0 + i][c0 + j];
      return v;
    }
    template <int BR, int BC>
    AnyBlock& operator=(const M<BR, BC>& v)
    {
      assert(BR == nr && BC == nc && m);
      for (int i = 0; i < BR; ++i)
        for (int j = 0; j < BC; ++j)
          m->a[r0 + i][c0 + j] = v.a[i][j];
      return *this;
    }
  };
  AnyBlock block(int r0, int c0, int nr, int nc) { return AnyBlock{this, this, r0, c0, nr, nc}; }
  AnyBlock block(int r0, int c0, int nr, int nc) const { return AnyBlock{nullptr, this, r0, c0, nr, nc}; }
};

template <int R, int C>
inline M<R, C> operator+(const M<R, C>& x, const M<R, C>& y)
{
  M<R, C> r;
  for (int i = 0; i < R; ++i)
    for (int j = 0; j < C; ++j)
      r.a[i][j] = x.a[i][j] + y.a[i][j];
  return r;
}
template <int R, int C>
inline M<R, C> operator*(float s, const M<R, C>& x)
{
  M<R, C> r;
  for (int i = 0; i < R; ++i)
    for (int j = 0; j < C; ++j)
      r.a[i][j] = s * x.a[i][j];
  return r;
}
template <int R, int C>
inline M<R, C> operator/(const M<R, C>& x, float s)
{
  M<R, C> r;
  for (int i = 0; i < R; ++i)
    for (int j = 0; j < C; ++j)
      r.a[i][j] = x.a[i][j] / s;
  return r;
}
template <int R, int K, int C>
inline M<R, C> operator*(const M<R, K>& x, const M<K, C>& y)
{
  M<R, C> r;
  for (int i = 0; i < R; ++i)
    for (int j = 0; j < C; ++j)
    {
      float s = x.a[i][0] * y.a[0][j];
      for (int k = 1; k < K; ++k)
        s = s + x.a[i][k] * y.a[k][j];
      r.a[i][j] = s;
    }
  return r;
}
}  // namespace Eigen

#endif
