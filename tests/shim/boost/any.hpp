// compile-check stand-in for <boost/any.hpp> (see tests/shim/README.md)
#ifndef XRC_SHIM_BOOST_ANY
#define XRC_SHIM_BOOST_ANY
namespace boost
{
class any
{
public:
  any();
  template <class T>
  any(const T&);
  template <class T>
  any& operator=(const T&);
  bool empty() const;
};
template <class T>
T any_cast(const any&);
}  // namespace boost
#endif
