// compile-check stand-in (see tests/shim/README.md)
#ifndef XRC_SHIM_ITK_MINMAX
#define XRC_SHIM_ITK_MINMAX
#include <itkImage.h>
namespace itk
{
template <class TImage>
class MinimumMaximumImageCalculator
{
public:
  using Pointer = SmartPointer<MinimumMaximumImageCalculator>;
  static Pointer New();
  void SetImage(const TImage*);
  void Compute();
  void ComputeMinimum();
  void ComputeMaximum();
  typename TImage::PixelType GetMinimum() const;
  typename TImage::PixelType GetMaximum() const;
  typename TImage::IndexType GetIndexOfMinimum() const;
  typename TImage::IndexType GetIndexOfMaximum() const;
};
}  // namespace itk
#endif
