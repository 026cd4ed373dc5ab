// compile-check stand-in (see tests/shim/README.md)
#ifndef XRC_SHIM_ITK_AFFINE
#define XRC_SHIM_ITK_AFFINE
#include <itkImage.h>
namespace itk
{
template <class T = double, unsigned int N = 3>
class AffineTransform
{
public:
  using Pointer = SmartPointer<AffineTransform>;
  using MatrixType = Matrix<T, N, N>;
  using OutputVectorType = Vector<T, N>;
  using InputPointType = Point<T, N>;
  using OutputPointType = Point<T, N>;
  static Pointer New();
  void SetMatrix(const MatrixType&);
  const MatrixType& GetMatrix() const;
  void SetTranslation(const OutputVectorType&);
  const OutputVectorType& GetTranslation() const;
  void SetOffset(const OutputVectorType&);
  const OutputVectorType& GetOffset() const;
  void SetIdentity();
  OutputPointType TransformPoint(const InputPointType&) const;
};
}  // namespace itk
#endif
