// compile-check stand-in (see tests/shim/README.md)
#ifndef XRC_SHIM_itkImageRegionConstIterator
#define XRC_SHIM_itkImageRegionConstIterator
#include <itkImage.h>
#ifndef XRC_SHIM_ITK_ITERATORS
#define XRC_SHIM_ITK_ITERATORS
namespace itk
{
template <class TImage>
class ImageRegionConstIterator
{
public:
  ImageRegionConstIterator(const TImage*, const typename TImage::RegionType&);
  void GoToBegin();
  bool IsAtEnd() const;
  ImageRegionConstIterator& operator++();
  const typename TImage::PixelType& Get() const;
  const typename TImage::PixelType& Value() const;
  typename TImage::IndexType GetIndex() const;
};
template <class TImage>
class ImageRegionIterator : public ImageRegionConstIterator<TImage>
{
public:
  ImageRegionIterator(TImage*, const typename TImage::RegionType&);
  void Set(const typename TImage::PixelType&) const;
  typename TImage::PixelType& Value();
};
}  // namespace itk
#endif
#endif
