// compile-check stand-in (see tests/shim/README.md)
#ifndef XRC_SHIM_ITK_INTENSITY_WINDOWING
#define XRC_SHIM_ITK_INTENSITY_WINDOWING
#include <itkProcessShim.h>
namespace itk
{
template <class TIn, class TOut = TIn>
class IntensityWindowingImageFilter : public ImageToImageFilter<TIn, TOut>
{
public:
  using Pointer = SmartPointer<IntensityWindowingImageFilter>;
  static Pointer New();
  void SetWindowMinimum(typename TIn::PixelType);
  void SetWindowMaximum(typename TIn::PixelType);
  void SetOutputMinimum(typename TOut::PixelType);
  void SetOutputMaximum(typename TOut::PixelType);
};
}  // namespace itk
#endif
