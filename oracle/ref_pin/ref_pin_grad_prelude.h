/* TEST INFRASTRUCTURE.  Stand-ins for compiling the reference's OWN gradient-metric classes on top of the NCC and
 * patch-NCC units:
 *   lib/regi/sim_metrics_2d/xregImgSimMetric2DGradImgCPU.cpp        allocate_resources, compute_sobel_grads
 *   lib/regi/sim_metrics_2d/xregImgSimMetric2DGradNCCCPU.cpp        allocate_resources, compute, process_mask
 *   lib/regi/sim_metrics_2d/xregImgSimMetric2DPatchGradNCCCPU.cpp   allocate_resources, compute, process_mask
 *   (+ the PatchCommon / PatchNCCCPU setters they call)
 * cv::GaussianBlur and cv::Sobel -- OpenCV, un-vendored -- are CALL-OUTS here: the test installs either the real OpenCV
 * (the cv2 Python binding) or the oracle's restatement of the two filters, so the class code runs over a chosen filter
 * implementation and carries none of its own.
 */
#ifndef XREG_REF_PIN_GRAD_PRELUDE_H
#define XREG_REF_PIN_GRAD_PRELUDE_H

#include "ref_pin_metric_prelude.h"
#include "ref_pin_ncc_prelude.h"

namespace cv
{
typedef void (*gauss_fn)(const float* src, int rows, int cols, int ksize, float* dst);
typedef void (*sobel_fn)(const float* src, int rows, int cols, int dx, int dy, float* dst);
extern gauss_fn g_gauss;
extern sobel_fn g_sobel;
inline void GaussianBlur(const Mat& src, Mat& dst, Size ksize, double, double)
{
  g_gauss(reinterpret_cast<const float*>(src.data), src.rows, src.cols, ksize.width, reinterpret_cast<float*>(dst.data));
}
inline void Sobel(const Mat& src, Mat& dst, int, int dx, int dy)
{
  g_sobel(reinterpret_cast<const float*>(src.data), src.rows, src.cols, dx, dy, reinterpret_cast<float*>(dst.data));
}
}  // namespace cv

namespace xreg
{

template <class T>
typename itk::Image2<T, 2>::Pointer ShallowCopyOpenCVToItk(cv::Mat& m)
{
  typename itk::Image2<T, 2>::Pointer p;
  p.p = std::make_shared<itk::Image2<T, 2>>();
  p.p->buf = reinterpret_cast<T*>(m.data);
  p.p->sz.s[0] = (std::size_t)m.cols;
  p.p->sz.s[1] = (std::size_t)m.rows;
  return p;
}

/* xregImgSimMetric2DGradImgCPU.h:36-100 */
class ImgSimMetric2DGradImgCPU : public ImgSimMetric2DCPU
{
public:
  void allocate_resources() override;
  using PixelBuffer = std::vector<Scalar>;
  using cvMatList = std::vector<cv::Mat>;
  void compute_sobel_grads();
  cv::Mat fixed_grad_img_x_;
  cv::Mat fixed_grad_img_y_;
  PixelBuffer grad_x_mov_imgs_buf_;
  PixelBuffer grad_y_mov_imgs_buf_;
  cvMatList mov_grad_imgs_x_;
  cvMatList mov_grad_imgs_y_;
  size_type smooth_img_kernel_rad_ = 5;
  cv::Mat tmp_smooth_img_;
};

/* xregImgSimMetric2DGradNCCCPU.h:36-80 */
class ImgSimMetric2DGradNCCCPU : public ImgSimMetric2DGradImgCPU
{
public:
  void allocate_resources() override;
  void compute() override;
  void process_mask() override;
  ImgSimMetric2DNCCCPU ncc_sim_x_;
  ImgSimMetric2DNCCCPU ncc_sim_y_;
};

/* xregImgSimMetric2DPatchGradNCCCPU.h:36-120 */
class ImgSimMetric2DPatchGradNCCCPU : public ImgSimMetric2DGradImgCPU, public ImgSimMetric2DPatchCommon
{
public:
  using Scalar = ImgSimMetric2DGradImgCPU::Scalar;
  using MaskScalar = ImgSimMetric2DGradImgCPU::MaskScalar;
  void allocate_resources() override;
  void compute() override;
  void process_mask() override;
  struct SimAux : public H5ReadWriteInterface
  {
    std::shared_ptr<H5ReadWriteInterface> sim_aux_x;
    std::shared_ptr<H5ReadWriteInterface> sim_aux_y;
  };
  ImgSimMetric2DPatchNCCCPU patch_ncc_x_;
  ImgSimMetric2DPatchNCCCPU patch_ncc_y_;
  bool enforce_same_patches_in_both_x_and_y_ = true;
  bool use_fixed_img_patch_variances_as_wgts_ = false;
  bool use_mov_img_patch_variances_as_wgts_ = false;
  bool use_variances_in_grad_imgs_as_wgts_ = false;
  std::shared_ptr<SimAux> sim_aux_;
  std::vector<ScalarList> mov_img_patch_vars_;
  std::vector<bool> do_not_use_scores_from_sub_objs_;
};

}  // namespace xreg

#endif
