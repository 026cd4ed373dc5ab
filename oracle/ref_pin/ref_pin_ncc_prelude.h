/* TEST INFRASTRUCTURE.  Stand-ins that let the reference's OWN NCC source lines compile outside an xReg build:
 *   lib/regi/sim_metrics_2d/xregImgSimMetric2DNCCCPU.cpp   the un-named namespace (ComputeLenFromMask,
 *       ComputeImage2DMeanStdDev, ComputeZeroMeanImageAndStats, ...WithMask), allocate_resources, compute, process_mask
 * (cut by anchor by build_ref_slice.py).  The class is declared with the reference's members (xregImgSimMetric2DNCCCPU.h).
 *
 * Eigen stand-in: dynamic row vectors, Map, MatrixBase<Derived> (CRTP) with element access, size(), array(), dot().
 * Element-wise pieces carry no convention.  The three REDUCTIONS the unmasked path uses -- array().mean(),
 * (array() - m).square().sum(), a.dot(b) -- do: they follow the documented shape of Eigen 3.3's linear vectorised
 * reduction for SSE floats (Packet4f, two packet accumulators over 8-element strides, one extra packet, horizontal
 * add as (p0 + p2) + (p1 + p3), scalar tail; start taken as aligned), the same convention xreg_oracle.c states.
 * Consequence: the MASKED NCC path (plain scalar loops in the reference) is pinned with no convention involved; the
 * unmasked path is pinned up to that reduction convention.
 */
#ifndef XREG_REF_PIN_NCC_PRELUDE_H
#define XREG_REF_PIN_NCC_PRELUDE_H

#include "ref_pin_sim_base.h"
#include "ref_pin_eigen_fixed.h"

namespace Eigen
{
const int Dynamic = -1;

namespace redux
{
/* sum over i of f(i), f given as a functor, in the vectorised order described above */
template <class F>
inline float run(std::ptrdiff_t n, F f)
{
  const std::ptrdiff_t ps = 4, a2 = (n / (2 * ps)) * (2 * ps), a1 = (n / ps) * ps;
  float res;
  if (a1)
  {
    float p0[4], p1[4];
    for (int l = 0; l < 4; ++l)
      p0[l] = f(l);
    if (a1 > ps)
    {
      for (int l = 0; l < 4; ++l)
        p1[l] = f(ps + l);
      for (std::ptrdiff_t i = 2 * ps; i < a2; i += 2 * ps)
        for (int l = 0; l < 4; ++l)
        {
          p0[l] = p0[l] + f(i + l);
          p1[l] = p1[l] + f(i + ps + l);
        }
      for (int l = 0; l < 4; ++l)
        p0[l] = p0[l] + p1[l];
      if (a1 > a2)
        for (int l = 0; l < 4; ++l)
          p0[l] = p0[l] + f(a2 + l);
    }
    res = (p0[0] + p0[2]) + (p0[1] + p0[3]);
    for (std::ptrdiff_t i = a1; i < n; ++i)
      res = res + f(i);
  }
  else
  {
    res = f(0);
    for (std::ptrdiff_t i = 1; i < n; ++i)
      res = res + f(i);
  }
  return res;
}
}  // namespace redux

template <class D>
struct scalar_of;

template <class S>
struct SqDiffExpr
{
  const S* p;
  std::ptrdiff_t n;
  S c;
  S sum() const
  {
    const S* q = p;
    const S cc = c;
    return redux::run(n, [q, cc](std::ptrdiff_t i) { return (q[i] - cc) * (q[i] - cc); });
  }
};
template <class S>
struct DiffExpr
{
  const S* p;
  std::ptrdiff_t n;
  S c;
  SqDiffExpr<S> square() const { return SqDiffExpr<S>{p, n, c}; }
};
template <class S>
struct ConstArrayView
{
  const S* p;
  std::ptrdiff_t n;
  S mean() const
  {
    const S* q = p;
    return redux::run(n, [q](std::ptrdiff_t i) { return q[i]; }) / S(n);   // sum() / Scalar(size())
  }
  DiffExpr<S> operator-(S c) const { return DiffExpr<S>{p, n, c}; }
};
template <class S>
struct ArrayView
{
  S* p;
  std::ptrdiff_t n;
  ArrayView& operator-=(S c)
  {
    for (std::ptrdiff_t i = 0; i < n; ++i)
      p[i] = p[i] - c;
    return *this;
  }
};

/* (a - b).array().square().sum(): the SSD reduction, same vectorised order */
template <class S>
struct VecSqDiffSum
{
  const S* a;
  const S* b;
  std::ptrdiff_t n;
  const VecSqDiffSum& array() const { return *this; }
  const VecSqDiffSum& square() const { return *this; }
  S sum() const
  {
    const S* p = a;
    const S* q = b;
    return redux::run(n, [p, q](std::ptrdiff_t i) { return (p[i] - q[i]) * (p[i] - q[i]); });
  }
};

template <class D>
struct MatrixBase
{
  using Scalar = typename scalar_of<D>::type;
  const D& derived() const { return *static_cast<const D*>(this); }
  D& derived() { return *static_cast<D*>(this); }
  std::ptrdiff_t size() const { return derived().size_(); }
  Scalar operator()(std::ptrdiff_t i) const { return derived().ptr_()[i]; }
  Scalar& operator()(std::ptrdiff_t i) { return derived().ptr_()[i]; }
  ConstArrayView<Scalar> array() const { return ConstArrayView<Scalar>{derived().ptr_(), size()}; }
  ArrayView<Scalar> array() { return ArrayView<Scalar>{derived().ptr_(), size()}; }
  /* row vectors: 1 x n */
  unsigned long rows() const { return 1; }
  unsigned long cols() const { return (unsigned long)size(); }
  Scalar operator()(unsigned long, unsigned long c) const { return derived().ptr_()[c]; }
  Scalar& operator()(unsigned long, unsigned long c) { return derived().ptr_()[c]; }
  template <class O>
  VecSqDiffSum<Scalar> operator-(const MatrixBase<O>& o) const
  {
    return VecSqDiffSum<Scalar>{derived().ptr_(), o.derived().ptr_(), size()};
  }
  template <class O>
  Scalar dot(const MatrixBase<O>& o) const
  {
    const Scalar* a = derived().ptr_();
    const Scalar* b = o.derived().ptr_();
    return redux::run(size(), [a, b](std::ptrdiff_t i) { return a[i] * b[i]; });
  }
};

template <class M>
struct Map;

/* dynamic row vector: specialisation of the fixed-size template of ref_pin_eigen_fixed.h */
template <class S>
struct Matrix<S, 1, Dynamic> : MatrixBase<Matrix<S, 1, Dynamic>>
{
  using Scalar = S;
  std::vector<S> v;
  void resize(std::size_t n) { v.resize(n); }
  std::ptrdiff_t size_() const { return (std::ptrdiff_t)v.size(); }
  const S* ptr_() const { return v.data(); }
  S* ptr_() { return v.data(); }
  Matrix& operator=(const Map<Matrix>& m);
};
template <class S>
struct scalar_of<Matrix<S, 1, Dynamic>>
{
  using type = S;
};

template <class M>
struct Map : MatrixBase<Map<M>>
{
  using Scalar = typename scalar_of<M>::type;
  Scalar* p;
  std::ptrdiff_t n;
  Map(Scalar* ptr, std::size_t len) : p(ptr), n((std::ptrdiff_t)len) {}
  std::ptrdiff_t size_() const { return n; }
  const Scalar* ptr_() const { return p; }
  Scalar* ptr_() { return p; }
};
template <class M>
struct scalar_of<Map<M>>
{
  using type = typename scalar_of<M>::type;
};

template <class S>
Matrix<S, 1, Dynamic>& Matrix<S, 1, Dynamic>::operator=(const Map<Matrix<S, 1, Dynamic>>& m)
{
  v.assign(m.p, m.p + m.n);
  return *this;
}

}  // namespace Eigen

namespace itk
{
/* GetBufferPointer() on the smart pointer target is all process_mask needs (provided by Image2) */
}

namespace xreg
{

class ImgSimMetric2DNCCCPU : public ImgSimMetric2DCPU
{
public:
  void allocate_resources() override;
  void compute() override;
  void process_mask() override;

  using ImageVec = Eigen::Matrix<Scalar, 1, Eigen::Dynamic>;
  using ImageMaskVec = Eigen::Matrix<MaskScalar, 1, Eigen::Dynamic>;
  using MappedImageMaskVec = Eigen::Map<ImageMaskVec>;

  size_type img_num_rows_ = 0;
  size_type img_num_cols_ = 0;
  size_type img_num_pix_ = 0;
  ImageVec zero_mean_fixed_vec_;
  Scalar fixed_img_mean_ = 0;
  Scalar fixed_img_stddev_ = 0;
  ImageMaskVec mask_vec_;
  size_type mask_len_ = 0;
};

/* xregImgSimMetric2DSSDCPU.h:36-75 */
class ImgSimMetric2DSSDCPU : public ImgSimMetric2DCPU
{
public:
  void allocate_resources() override;
  void compute() override;
  void process_mask() override;
  using ImageVec = Eigen::Matrix<Scalar, 1, Eigen::Dynamic>;
  using ImageMaskVec = Eigen::Matrix<MaskScalar, 1, Eigen::Dynamic>;
  using MappedImageVec = Eigen::Map<ImageVec>;
  using MappedImageMaskVec = Eigen::Map<ImageMaskVec>;
  ImageVec fixed_img_vec_;
  ImageMaskVec mask_vec_;
};

/* xregImgSimMetric2DCombine.h:36-100: the view combiners (members as in the reference; compute() bodies are its lines) */
using ImgSimMetric2D = ImgSimMetric2DCPU;
class ImgSimMetric2DCombine
{
public:
  using Scalar = ImgSimMetric2D::Scalar;
  using ScalarList = ImgSimMetric2D::ScalarList;
  virtual ~ImgSimMetric2DCombine() {}
  virtual void compute() = 0;
  size_type num_sim_metrics_ = 0;
  size_type num_projs_per_sim_metric_ = 0;
  ScalarList sim_vals_;
  std::vector<ImgSimMetric2D*> sim_objs_;
};
class ImgSimMetric2DCombineAddition : public ImgSimMetric2DCombine
{
public:
  void compute();
};
class ImgSimMetric2DCombineMean : public ImgSimMetric2DCombine
{
public:
  void compute();
};

}  // namespace xreg

#endif
