/* TEST INFRASTRUCTURE.  Stand-ins that let the reference's OWN patch-NCC source lines compile outside an xReg build:
 *   lib/regi/sim_metrics_2d/xregImgSimMetric2DPatchCommon.cpp   PatchInfo::center_row_col / ocv_roi, num_patches,
 *                                                               setup_patches, compute_weights, patch_indices_to_use
 *   lib/regi/sim_metrics_2d/xregImgSimMetric2DPatchNCCCPU.cpp   allocate_resources, compute, process_mask,
 *                                                               detail::ComputePatchMeanStdDev
 * (cut by anchor into a generated translation unit by build_ref_slice.py; nothing of the reference is copied here).
 *
 * Third-party types are replaced by minimal functional ones -- cv::Mat as a strided view (constructor over user data,
 * ROI by cv::Rect, at<T>(r, c)), cv::Rect, cv::DataType, itk::Image<T,2> behind a smart pointer, a serial ParallelFor --
 * none of which carries arithmetic: every floating-point operation of the patch metric is in the reference's lines.
 * The reference's class hierarchy (ImgSimMetric2D -> ImgSimMetric2DCPU, ImgSimMetric2DPatchCommon ->
 * ImgSimMetric2DPatchNCCCPU) is declared here with the data members those functions touch, under the reference's
 * names and types (xregImgSimMetric2D.h:42-156, xregImgSimMetric2DCPU.h, xregImgSimMetric2DPatchCommon.h:40-170,
 * xregImgSimMetric2DPatchNCCCPU.h:36-115): declarations only, the definitions come from the reference.
 */
#ifndef XREG_REF_PIN_METRIC_PRELUDE_H
#define XREG_REF_PIN_METRIC_PRELUDE_H

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

#include "ref_pin_sim_base.h"

namespace cv
{
struct Rect
{
  int x = 0, y = 0, width = 0, height = 0;
};
template <class T>
struct DataType;
template <>
struct DataType<float>
{
  enum { type = 5 };
};
template <>
struct DataType<unsigned char>
{
  enum { type = 0 };
};
struct Size
{
  int width = 0, height = 0;
  Size() {}
  Size(std::size_t w, std::size_t h) : width((int)w), height((int)h) {}
};
struct Mat
{
  int rows = 0, cols = 0;
  int type_ = 0;
  unsigned char* data = nullptr;
  std::size_t step = 0;  // bytes per row
  std::shared_ptr<std::vector<unsigned char>> own;   // storage of Mat::zeros
  Size size() const { return Size(cols, rows); }
  int type() const { return type_; }
  static Mat zeros(Size s, int type)
  {
    Mat m;
    m.rows = s.height;
    m.cols = s.width;
    m.type_ = type;
    m.step = (std::size_t)s.width * elem(type);
    m.own = std::make_shared<std::vector<unsigned char>>(m.step * (std::size_t)s.height, (unsigned char)0);
    m.data = m.own->data();
    return m;
  }
  static std::size_t elem(int type) { return type == 5 ? 4u : 1u; }
  Mat() {}
  Mat(std::size_t r, std::size_t c, int type, void* p)
    : rows((int)r), cols((int)c), type_(type), data(static_cast<unsigned char*>(p)), step(c * elem(type))
  {
  }
  Mat operator()(const Rect& roi) const
  {
    Mat m;
    m.rows = roi.height;
    m.cols = roi.width;
    m.type_ = type_;
    m.step = step;
    m.data = data + (std::size_t)roi.y * step + (std::size_t)roi.x * elem(type_);
    return m;
  }
  template <class T>
  T& at(int r, int c)
  {
    return *reinterpret_cast<T*>(data + (std::size_t)r * step + (std::size_t)c * sizeof(T));
  }
  template <class T>
  const T& at(int r, int c) const
  {
    return *reinterpret_cast<const T*>(data + (std::size_t)r * step + (std::size_t)c * sizeof(T));
  }
};
}  // namespace cv

#include "ref_pin_eigen_fixed.h"


namespace xreg
{

template <class T>
cv::Mat ShallowCopyItkToOpenCV(itk::Image2<T, 2>* img)
{
  return cv::Mat(img->sz[1], img->sz[0], cv::DataType<T>::type, img->buf);
}

template <class T>
typename itk::Image2<T, 2>::Pointer MakeITK2DVol(const size_type num_cols, const size_type num_rows, const T val = T())
{
  typename itk::Image2<T, 2>::Pointer p;
  p.p = std::make_shared<itk::Image2<T, 2>>();
  p.p->own.assign(num_cols * num_rows, val);
  p.p->buf = p.p->own.data();
  p.p->sz.s[0] = num_cols;
  p.p->sz.s[1] = num_rows;
  return p;
}

template <class tPixelType>
std::vector<cv::Mat> AllocContiguousBufferForOpenCVImages(const size_type num_rows, const size_type num_cols,
                                                          const size_type num_imgs, std::vector<tPixelType>* pix_buf)
{
  pix_buf->assign(num_rows * num_cols * num_imgs, tPixelType(0));
  std::vector<cv::Mat> imgs(num_imgs);
  for (size_type i = 0; i < num_imgs; ++i)
    imgs[i] = cv::Mat(num_rows, num_cols, cv::DataType<tPixelType>::type, &pix_buf->operator[](num_rows * num_cols * i));
  return imgs;
}

inline void SeedRNGEngWithRandDev(std::mt19937* eng)
{
  std::random_device rd;
  eng->seed(rd());
}

class ImgSimMetric2DPatchCommon
{
public:
  using Scalar = ImgSimMetric2DCPU::Scalar;
  using MaskScalar = ImgSimMetric2DCPU::MaskScalar;
  using ListOfSimScalarLists = std::vector<std::vector<Scalar>>;
  using WgtImg = itk::Image2<Scalar, 2>;
  using WgtImgPtr = WgtImg::Pointer;

  struct PatchInfo
  {
    size_type start_row;
    size_type start_col;
    size_type stop_row;
    size_type stop_col;
    Scalar weight;
    std::array<size_type, 2> center_row_col() const;
    cv::Rect ocv_roi() const;
  };
  using PatchInfoList = std::vector<PatchInfo>;
  using PatchIndexList = std::vector<size_type>;

  size_type num_patches() const;
  void set_from_other(const ImgSimMetric2DPatchCommon& other);
  void set_weights_from_other(const ImgSimMetric2DPatchCommon& other);
  void set_patches_to_use(const PatchIndexList& patch_inds);
  const ListOfSimScalarLists& sim_vals_for_each_patch() const;
  void set_wgt_img(WgtImgPtr wgt_img)
  {
    wgt_img_ = wgt_img;
    need_to_recompute_weights_ = true;
  }

  void setup_patches(const size_type img_num_rows, const size_type img_num_cols, cv::Mat* mask,
                     const size_type num_mov_imgs, const bool seed_rng = true);
  bool compute_weights(cv::Mat* mask);
  PatchIndexList patch_indices_to_use();

  PatchInfoList patch_infos_;
  size_type patch_radius_ = 5;
  size_type patch_stride_ = 1;
  size_type patch_diam_ = 0;
  bool compute_mean_of_patch_sims_ = false;
  bool weight_patch_sims_in_combine_ = true;
  bool use_mask_for_weighting_ = true;
  bool use_mask_for_patch_stats_ = false;
  bool save_all_per_patch_scores_ = false;
  bool normalize_weights_as_prob_ = true;
  bool choose_rand_patches_ = false;
  size_type num_rand_patches_ = 100;
  double rand_patch_min_pixels_sep_ = -1;
  bool patches_setup_ = false;
  ListOfSimScalarLists sim_vals_for_each_patch_;
  std::mt19937 rng_eng_;
  std::discrete_distribution<size_type> patch_idx_dist_;
  bool do_not_update_patch_inds_to_use_ = false;
  PatchIndexList patch_inds_to_use_;
  WgtImgPtr wgt_img_;
  bool need_to_recompute_weights_ = true;
};

class ImgSimMetric2DPatchNCCCPU : public ImgSimMetric2DCPU, public ImgSimMetric2DPatchCommon
{
public:
  using Scalar = ImgSimMetric2DCPU::Scalar;
  using MaskScalar = ImgSimMetric2DCPU::MaskScalar;

  void allocate_resources() override;
  void compute() override;

  struct SimAux : public H5ReadWriteInterface
  {
    std::vector<PatchInfoList> patch_infos_per_compute_call;
    std::vector<PatchIndexList> patch_indices_per_compute_call;
  };

  void process_mask() override;
  std::shared_ptr<H5ReadWriteInterface> aux_info();
  void set_use_fixed_img_patch_variances_as_wgts(const bool use_vars_as_wgts);
  void set_use_mov_img_patch_variances_as_wgts(const bool use_vars_as_wgts);
  void set_other_mov_img_patch_vars(const std::vector<ScalarList>* other_vars);

  size_type img_num_rows_ = 0;
  size_type img_num_cols_ = 0;
  std::vector<Scalar> fixed_scaled_buf_;
  std::vector<cv::Mat> fixed_scaled_patches_;
  ScalarList cur_mov_img_patch_ncc_vals_;
  bool use_fixed_img_patch_variances_as_wgts_ = false;
  bool use_mov_img_patch_variances_as_wgts_ = false;
  const std::vector<ScalarList>* other_mov_img_patch_vars_ = nullptr;
  std::shared_ptr<SimAux> sim_aux_;
  bool init_fixed_img_stats_computed_ = false;
  cv::Mat fixed_ocv_img_;
};

namespace detail
{
std::tuple<ImgSimMetric2DPatchNCCCPU::Scalar, ImgSimMetric2DPatchNCCCPU::Scalar, size_type>
ComputePatchMeanStdDev(const cv::Mat& p, const cv::Mat* m, const bool use_mask_for_stats);
}  // namespace detail

}  // namespace xreg

#endif
