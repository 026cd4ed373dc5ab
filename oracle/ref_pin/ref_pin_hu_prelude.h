/* TEST INFRASTRUCTURE.  Stand-ins (no arithmetic) for the ITK filter scaffolding around the reference's
 * HUToLinAttFilter::GenerateData (lib/image/xregHUToLinAtt.cpp), which build_ref_slice.py cuts from /root/reference:
 * a flat itk::Image<float,3>, region iterators over it, and the filter class with the reference's data members and
 * default values (xregHUToLinAtt.h: mu_water_, mu_air_, hu_lower_). */
#ifndef XREG_REF_PIN_HU_PRELUDE_H
#define XREG_REF_PIN_HU_PRELUDE_H

#include <algorithm>
#include <cstddef>
#include <cstdint>
#include <vector>

namespace itk
{
template <class T, unsigned N>
struct Image
{
  using RegionType = std::size_t;   // number of voxels is all the slice needs from a region
  const T* in = nullptr;            // input view
  std::vector<T> out;               // output storage
  std::size_t n = 0;
  RegionType GetLargestPossibleRegion() const { return n; }
  void SetRegions(RegionType r) { n = r; }
  void Allocate() { out.assign(n, T()); }
};
template <class I>
struct ImageRegionConstIterator
{
  const I* img;
  std::size_t i = 0, n;
  ImageRegionConstIterator(const I* im, std::size_t region) : img(im), n(region) {}
  void GoToBegin() { i = 0; }
  bool IsAtEnd() const { return i >= n; }
  ImageRegionConstIterator& operator++()
  {
    ++i;
    return *this;
  }
  float Get() const { return img->in[i]; }
};
template <class I>
struct ImageRegionIterator
{
  I* img;
  std::size_t i = 0, n;
  ImageRegionIterator(I* im, std::size_t region) : img(im), n(region) {}
  void GoToBegin() { i = 0; }
  ImageRegionIterator& operator++()
  {
    ++i;
    return *this;
  }
  void Set(float v) { img->out[i] = v; }
};
}  // namespace itk

namespace xreg
{
class HUToLinAttFilter
{
public:
  using Vol = itk::Image<float, 3>;
  const Vol* input = nullptr;
  Vol output;
  const Vol* GetInput() const { return input; }
  Vol* GetOutput() { return &output; }
  void GenerateData();
  double mu_water_ = 0.02683 * 1.0;
  double mu_air_ = 0.02485 * 0.0001;
  double hu_lower_ = -1000;
};
}  // namespace xreg

#endif
