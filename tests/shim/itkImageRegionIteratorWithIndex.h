// compile-check stand-in (see tests/shim/README.md)
#ifndef XRC_SHIM_ITK_ITER_WITH_INDEX
#define XRC_SHIM_ITK_ITER_WITH_INDEX
#include <itkImageRegionConstIterator.h>
namespace itk
{
template <class TImage>
class ImageRegionConstIteratorWithIndex : public ImageRegionConstIterator<TImage>
{
public:
  ImageRegionConstIteratorWithIndex(const TImage*, const typename TImage::RegionType&);
};
template <class TImage>
class ImageRegionIteratorWithIndex : public ImageRegionIterator<TImage>
{
public:
  ImageRegionIteratorWithIndex(TImage*, const typename TImage::RegionType&);
};
}  // namespace itk
#endif
