// compile-check stand-in (see tests/shim/README.md)
#ifndef XRC_SHIM_ITK_RGBPIXEL
#define XRC_SHIM_ITK_RGBPIXEL
#include <itkImage.h>
namespace itk
{
template <class T = unsigned short>
class RGBPixel : public FixedArray<T, 3>
{
public:
  using ComponentType = T;
  RGBPixel();
  void Set(T, T, T);
  void SetRed(T);
  void SetGreen(T);
  void SetBlue(T);
  const T& GetRed() const;
  const T& GetGreen() const;
  const T& GetBlue() const;
};
}  // namespace itk
#endif
