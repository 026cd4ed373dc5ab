// compile-check stand-in for <opencv2/imgproc/imgproc.hpp> (see tests/shim/README.md)
#ifndef XRC_SHIM_OPENCV_IMGPROC
#define XRC_SHIM_OPENCV_IMGPROC
#include <opencv2/core/core.hpp>
namespace cv
{
void GaussianBlur(const Mat&, Mat&, Size, double, double = 0, int = 4);
void Sobel(const Mat&, Mat&, int, int, int, int = 3, double = 1, double = 0, int = 4);
void Canny(const Mat&, Mat&, double, double, int = 3, bool = false);
}  // namespace cv
#endif
