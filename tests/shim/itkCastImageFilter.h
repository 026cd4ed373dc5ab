// compile-check stand-in (see tests/shim/README.md)
#ifndef XRC_SHIM_ITK_CAST
#define XRC_SHIM_ITK_CAST
#include <itkProcessShim.h>
XRC_SHIM_ITK_FILTER(CastImageFilter)
#endif
