// compile-check stand-in (see tests/shim/README.md): the shape of an ITK image-to-image filter
#ifndef XRC_SHIM_ITK_PROCESS
#define XRC_SHIM_ITK_PROCESS
#include <itkImage.h>
namespace itk
{
template <class TIn, class TOut>
class ImageToImageFilter
{
public:
  void SetInput(const TIn*);
  TOut* GetOutput();
  void Update();
};
}  // namespace itk
#define XRC_SHIM_ITK_FILTER(Name)                                   \
  namespace itk                                                     \
  {                                                                 \
  template <class TIn, class TOut = TIn>                            \
  class Name : public ImageToImageFilter<TIn, TOut>                 \
  {                                                                 \
  public:                                                           \
    using Pointer = SmartPointer<Name>;                             \
    static Pointer New();                                           \
  };                                                                \
  }
#endif
