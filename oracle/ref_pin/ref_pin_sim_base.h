/* TEST INFRASTRUCTURE.  Shared by the metric stand-in headers of oracle/ref_pin/: itk::Image<T,2> behind a smart
 * pointer, the serial ParallelFor, and the reference's ImgSimMetric2D -> ImgSimMetric2DCPU classes flattened to the data
 * members the sliced functions touch (declarations; no arithmetic). */
#ifndef XREG_REF_PIN_SIM_BASE_H
#define XREG_REF_PIN_SIM_BASE_H

#include <algorithm>
#include <array>
#include <cassert>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <limits>
#include <memory>
#include <numeric>
#include <random>
#include <string>
#include <tuple>
#include <vector>

#ifndef xregASSERT
#define xregASSERT(x) assert(x)
#endif

namespace itk
{
template <class T, unsigned N>
struct Image2
{
  struct Size
  {
    std::size_t s[N];
    std::size_t operator[](unsigned i) const { return s[i]; }
  };
  struct Region
  {
    Size sz;
    Size GetSize() const { return sz; }
  };
  std::vector<T> own;
  T* buf = nullptr;
  Size sz;
  Region GetLargestPossibleRegion() const { return Region{sz}; }
  T* GetBufferPointer() { return buf; }
  /* itk::SmartPointer: GetPointer(), implicit conversion to the raw pointer (hence to bool), -> */
  struct Pointer
  {
    std::shared_ptr<Image2> p;
    Image2* GetPointer() const { return p.get(); }
    operator Image2*() const { return p.get(); }
    Image2* operator->() const { return p.get(); }
  };
};
}  // namespace itk

namespace H5
{
class Group;
}

namespace xreg
{

using size_type = std::size_t;

struct RangeType
{
  size_type begin_, end_;
  RangeType(const size_type b, const size_type e) : begin_(b), end_(e) {}
  size_type begin() const { return begin_; }
  size_type end() const { return end_; }
};
template <class Fn>
void ParallelFor(Fn& fn_obj, const RangeType& r)
{
  fn_obj(r);
}

struct H5ReadWriteInterface
{
  virtual ~H5ReadWriteInterface() {}
};

/* ImgSimMetric2D + ImgSimMetric2DCPU, flattened: the members the sliced functions touch */
class ImgSimMetric2DCPU
{
public:
  using Scalar = float;
  using MaskScalar = unsigned char;
  using ScalarList = std::vector<Scalar>;
  using Image = itk::Image2<Scalar, 2>;
  using ImagePtr = Image::Pointer;
  using ImageMask = itk::Image2<MaskScalar, 2>;
  using ImageMaskPtr = ImageMask::Pointer;

  virtual ~ImgSimMetric2DCPU() {}
  virtual void allocate_resources() { sim_vals_.assign(num_mov_imgs_, 0); }
  virtual void compute() = 0;

  /* ImgSimMetric2D::process_updated_mask (xregImgSimMetric2D.cpp:110-118) */
  void process_updated_mask()
  {
    if (mask_updated_)
    {
      process_mask();
      mask_updated_ = false;
    }
  }
  /* ImgSimMetric2DCPU::pre_compute (xregImgSimMetric2DCPU.cpp:90-98) with a host buffer: nothing to sync */
  void pre_compute() { process_updated_mask(); }

  Scalar sim_val(const size_type i) const { return sim_vals_[i]; }
  /* ImgSimMetric2D / ImgSimMetric2DCPU setters (xregImgSimMetric2D.cpp, xregImgSimMetric2DCPU.cpp:72-88): assignments */
  void set_num_moving_images(const size_type n) { num_mov_imgs_ = n; }
  void set_mov_imgs_host_buf(Scalar* buf, const size_type proj_offset = 0)
  {
    mov_imgs_buf_ = buf;
    (void)proj_offset;
  }
  void set_fixed_image(ImagePtr img) { fixed_img_ = img; }
  void set_mask(ImageMaskPtr m)
  {
    mask_ = m;
    mask_updated_ = true;
  }
  void set_save_aux_info(const bool b) { save_aux_info_ = b; }

  ImagePtr fixed_img_;
  ImageMaskPtr mask_;
  bool mask_updated_ = true;
  size_type num_mov_imgs_ = 0;
  Scalar* mov_imgs_buf_ = nullptr;
  ScalarList sim_vals_;
  bool save_aux_info_ = false;

protected:
  virtual void process_mask() {}
};

}  // namespace xreg

#endif
