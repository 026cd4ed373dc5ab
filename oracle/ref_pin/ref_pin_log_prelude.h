/* TEST INFRASTRUCTURE.  Stand-ins in front of the reference's ImageIntensLogTransFilter::GenerateData
 * (lib/image/xregImageIntensLogTrans.cpp:55-144), compiled by oracle/ref_pin/build_ref_slice.py into
 * oracle/_ref/libxreg_refslice_log.so.  They carry NO arithmetic of the filter: an owning 2-D float image, the
 * ImageToImageFilter plumbing GenerateData uses (GetInput / GetOutput / SetRegions / Allocate), a serial ParallelTransform,
 * and itk::DiscreteGaussianImageFilter as a CALL-OUT (ITK is an un-vendored dependency: the test installs the oracle's
 * restatement of it, so everything but that smoothing is the reference's own lines). */
#pragma once
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <memory>
#include <type_traits>
#include <vector>

extern "C" {
typedef void (*xref_gaussian_fn)(const float* img, unsigned rows, unsigned cols, double variance, float* out);
}
static xref_gaussian_fn g_xref_gaussian = nullptr;

namespace itk
{
struct Region2
{
  std::size_t cols = 0, rows = 0;
  std::size_t GetNumberOfPixels() const { return cols * rows; }
};

template <class T, unsigned N>
struct Image
{
  static_assert(N == 2, "2-D images only");
  Region2 region;
  std::vector<T> own;
  const T* ext = nullptr;   /* a caller's buffer (inputs) */
  Region2 GetLargestPossibleRegion() const { return region; }
  void SetRegions(const Region2& r) { region = r; }
  void Allocate() { own.assign(region.GetNumberOfPixels(), T(0)); }
  const T* GetBufferPointer() const { return ext ? ext : own.data(); }
  T* GetBufferPointer() { return own.data(); }
};

template <class TIn, class TOut>
struct ImageToImageFilter
{
  using InputImageType = TIn;
  using InputImagePixelType = float;
  TIn* in = nullptr;
  TOut out;
  virtual ~ImageToImageFilter() {}
  const TIn* GetInput() const { return in; }
  TOut* GetOutput() { return &out; }
  void Modified() {}
};

/* itk::DiscreteGaussianImageFilter: every call of GenerateData's use (:97-103) forwarded to the installed call-out */
template <class TIn, class TOut>
struct DiscreteGaussianImageFilter
{
  struct Ptr
  {
    std::shared_ptr<DiscreteGaussianImageFilter> p;
    DiscreteGaussianImageFilter* operator->() const { return p.get(); }
  };
  static Ptr New()
  {
    Ptr q;
    q.p = std::make_shared<DiscreteGaussianImageFilter>();
    return q;
  }
  const TIn* in = nullptr;
  double variance = 0;
  bool use_spacing = true;
  TOut out;
  void SetInput(const TIn* i) { in = i; }
  void SetUseImageSpacing(bool b) { use_spacing = b; }
  void SetVariance(double v) { variance = v; }
  void Update()
  {
    out.SetRegions(in->GetLargestPossibleRegion());
    out.Allocate();
    g_xref_gaussian(in->GetBufferPointer(), (unsigned)in->region.rows, (unsigned)in->region.cols, variance, out.GetBufferPointer());
  }
  TOut* GetOutput() { return &out; }
};
}  // namespace itk

namespace xreg
{
template <class InputIt, class OutputIt, class UnaryOp>
OutputIt ParallelTransform(InputIt begin_in, InputIt end_in, OutputIt begin_out, UnaryOp op)
{
  return std::transform(begin_in, end_in, begin_out, op);   /* lib/common/xregTBBUtils.h:161-180, XREG_NO_TBB branch */
}

class ImageIntensLogTransFilter : public itk::ImageToImageFilter<itk::Image<float, 2>, itk::Image<float, 2>>
{
public:
  using Superclass = itk::ImageToImageFilter<itk::Image<float, 2>, itk::Image<float, 2>>;
  using InputImagePixelType = Superclass::InputImagePixelType;
  void GenerateData();
  bool normalize_zero_one_ = false;
  bool use_max_intensity_as_I0_ = true;
  InputImagePixelType I0_ = 1;
};
}  // namespace xreg
