// compile-check stand-in for <itkImage.h> (see tests/shim/README.md): declarations only
#ifndef XRC_SHIM_ITK_IMAGE
#define XRC_SHIM_ITK_IMAGE
// the real header pulls these in; the reference's headers rely on that
#include <algorithm>
#include <cstddef>
#include <memory>
#include <string>
#include <vector>
namespace itk
{
typedef unsigned long SizeValueType;
typedef long IndexValueType;
template <class T>
class SmartPointer
{
public:
  SmartPointer();
  SmartPointer(T*);
  SmartPointer(const SmartPointer&);
  SmartPointer& operator=(T*);
  SmartPointer& operator=(const SmartPointer&);
  T* operator->() const;
  T& operator*() const;
  T* GetPointer() const;
  operator T*() const;
  bool IsNull() const;
  bool IsNotNull() const;
};
template <unsigned int N>
struct Size
{
  SizeValueType m_Size[N];
  SizeValueType& operator[](unsigned int);
  const SizeValueType& operator[](unsigned int) const;
  void Fill(SizeValueType);
};
template <unsigned int N>
struct Index
{
  IndexValueType m_Index[N];
  IndexValueType& operator[](unsigned int);
  const IndexValueType& operator[](unsigned int) const;
  void Fill(IndexValueType);
};
template <class T, unsigned int N>
class FixedArray
{
public:
  FixedArray();
  T& operator[](unsigned int);
  const T& operator[](unsigned int) const;
  void Fill(const T&);
  T* GetDataPointer();
  const T* GetDataPointer() const;
};
template <class T, unsigned int N>
class Vector : public FixedArray<T, N>
{
};
template <class T, unsigned int N>
class Point : public FixedArray<T, N>
{
};
template <class T, unsigned int N>
class ContinuousIndex : public FixedArray<T, N>
{
};
template <class T, unsigned int R, unsigned int C>
class Matrix
{
public:
  T& operator()(unsigned int, unsigned int);
  const T& operator()(unsigned int, unsigned int) const;
  T* operator[](unsigned int);
  const T* operator[](unsigned int) const;
  void SetIdentity();
  void Fill(const T&);
};
template <unsigned int N>
class ImageRegion
{
public:
  using SizeType = Size<N>;
  using IndexType = Index<N>;
  ImageRegion();
  ImageRegion(const IndexType&, const SizeType&);
  void SetSize(const SizeType&);
  void SetSize(unsigned int, SizeValueType);
  void SetIndex(const IndexType&);
  void SetIndex(unsigned int, IndexValueType);
  const SizeType& GetSize() const;
  SizeValueType GetSize(unsigned int) const;
  const IndexType& GetIndex() const;
  SizeValueType GetNumberOfPixels() const;
};
template <class ElemId, class T>
class ImportImageContainer
{
public:
  using Pointer = SmartPointer<ImportImageContainer>;
  static Pointer New();
  void SetImportPointer(T* ptr, ElemId num, bool let_container_manage_memory = false);
  T* GetImportPointer();
  T* GetBufferPointer();
  ElemId Size() const;
  void Reserve(ElemId);
  void ContainerManageMemoryOn();
  void ContainerManageMemoryOff();
};
template <class T, unsigned int N = 2>
class Image
{
public:
  using Self = Image;
  using Pointer = SmartPointer<Self>;
  using ConstPointer = SmartPointer<const Self>;
  using PixelType = T;
  using InternalPixelType = T;
  using ValueType = T;
  using SizeType = Size<N>;
  using IndexType = Index<N>;
  using RegionType = ImageRegion<N>;
  using SpacingType = Vector<double, N>;
  using PointType = Point<double, N>;
  using DirectionType = Matrix<double, N, N>;
  using PixelContainer = ImportImageContainer<SizeValueType, T>;
  using PixelContainerPointer = typename PixelContainer::Pointer;
  static constexpr unsigned int ImageDimension = N;
  static Pointer New();
  void SetRegions(const RegionType&);
  void SetRegions(const SizeType&);
  void SetLargestPossibleRegion(const RegionType&);
  void SetBufferedRegion(const RegionType&);
  void SetRequestedRegion(const RegionType&);
  const RegionType& GetLargestPossibleRegion() const;
  const RegionType& GetBufferedRegion() const;
  const RegionType& GetRequestedRegion() const;
  void Allocate(bool = false);
  void FillBuffer(const T&);
  T* GetBufferPointer();
  const T* GetBufferPointer() const;
  void SetPixelContainer(PixelContainer*);
  PixelContainer* GetPixelContainer();
  const PixelContainer* GetPixelContainer() const;
  void SetSpacing(const SpacingType&);
  void SetSpacing(const double*);
  void SetSpacing(const float*);
  const SpacingType& GetSpacing() const;
  void SetOrigin(const PointType&);
  void SetOrigin(const double*);
  void SetOrigin(const float*);
  const PointType& GetOrigin() const;
  void SetDirection(const DirectionType&);
  const DirectionType& GetDirection() const;
  T& GetPixel(const IndexType&);
  const T& GetPixel(const IndexType&) const;
  void SetPixel(const IndexType&, const T&);
  template <class C>
  void TransformContinuousIndexToPhysicalPoint(const ContinuousIndex<C, N>&, Point<C, N>&) const;
  template <class C>
  bool TransformPhysicalPointToContinuousIndex(const Point<C, N>&, ContinuousIndex<C, N>&) const;
  template <class C>
  void TransformIndexToPhysicalPoint(const IndexType&, Point<C, N>&) const;
  template <class C>
  bool TransformPhysicalPointToIndex(const Point<C, N>&, IndexType&) const;
  void CopyInformation(const Image*);
  void Update();
  void DisconnectPipeline();
  void Register() const;
  void UnRegister() const;
};
}  // namespace itk
#endif
