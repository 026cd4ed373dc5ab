// compile-check stand-in for <boost/compute/container/vector.hpp> (see tests/shim/README.md)
#ifndef XRC_SHIM_BOOST_COMPUTE_VECTOR
#define XRC_SHIM_BOOST_COMPUTE_VECTOR
#include <cstddef>
namespace boost
{
namespace compute
{
class context;
class device;
class command_queue
{
public:
  command_queue();
};
template <class T>
class vector
{
public:
  vector();
  std::size_t size() const;
};
}  // namespace compute
}  // namespace boost
#endif
