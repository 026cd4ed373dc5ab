// compile-check stand-in for <opencv2/core/core.hpp> (see tests/shim/README.md): declarations only
#ifndef XRC_SHIM_OPENCV_CORE
#define XRC_SHIM_OPENCV_CORE
#include <cstddef>
#define CV_8U 0
#define CV_8S 1
#define CV_16U 2
#define CV_16S 3
#define CV_32S 4
#define CV_32F 5
#define CV_64F 6
#define CV_8UC1 0
#define CV_8SC1 1
#define CV_16UC1 2
#define CV_16SC1 3
#define CV_32SC1 4
#define CV_8UC3 16
#define CV_32FC1 5
#define CV_64FC1 6
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn)-1) << 3))
namespace cv
{
typedef unsigned char uchar;
template <class T>
struct Size_
{
  Size_();
  Size_(T, T);
  T width, height;
};
typedef Size_<int> Size;
template <class T>
struct Rect_
{
  Rect_();
  Rect_(T, T, T, T);
  T x, y, width, height;
};
typedef Rect_<int> Rect;
template <class T>
struct Point_
{
  Point_();
  Point_(T, T);
  T x, y;
};
typedef Point_<int> Point;
template <class T, int N>
struct Vec
{
  T val[N];
  T& operator[](int);
  const T& operator[](int) const;
};
typedef Vec<uchar, 3> Vec3b;
struct Scalar
{
  Scalar();
  Scalar(double);
  Scalar(double, double, double, double = 0);
};
template <class T>
struct DataType
{
  enum { type = CV_32F, depth = CV_32F, channels = 1 };
};
template <class T>
struct DataDepth
{
  enum { value = CV_32F };
};
class Mat
{
public:
  Mat();
  Mat(int rows, int cols, int type);
  Mat(int rows, int cols, int type, void* data, std::size_t step = 0);
  Mat(Size, int type);
  Mat(const Mat&, const Rect&);
  Mat operator()(const Rect&) const;
  Mat clone() const;
  void copyTo(Mat&) const;
  void create(int rows, int cols, int type);
  Mat& setTo(const Scalar&);
  template <class T>
  T& at(int, int);
  template <class T>
  const T& at(int, int) const;
  template <class T>
  T* ptr(int = 0);
  template <class T>
  const T* ptr(int = 0) const;
  bool isContinuous() const;
  std::size_t total() const;
  std::size_t elemSize() const;
  bool empty() const;
  int type() const;
  int depth() const;
  int channels() const;
  Size size() const;
  int rows, cols;
  unsigned char* data;
};
}  // namespace cv
#endif
