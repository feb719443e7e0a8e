// TEST INFRASTRUCTURE ONLY (oracle). Minimal stand-in for boost::shared_array (Boost is absent
// from this image; /root/reference/include/mcarray/mcadefs.h:51 needs it).
#ifndef ORACLE_STANDIN_BOOST_SHARED_ARRAY_HPP
#define ORACLE_STANDIN_BOOST_SHARED_ARRAY_HPP
#include <cstddef>
#include <memory>
namespace boost {
template <class T> class shared_array {
 public:
  shared_array() {}
  explicit shared_array(T *p) : _p(p, std::default_delete<T[]>()) {}
  template <class D> shared_array(T *p, D d) : _p(p, d) {}
  void reset() { _p.reset(); }
  void reset(T *p) { _p.reset(p, std::default_delete<T[]>()); }
  template <class D> void reset(T *p, D d) { _p.reset(p, d); }
  T *get() const { return _p.get(); }
  T &operator[](std::ptrdiff_t i) const { return _p.get()[i]; }
  explicit operator bool() const { return bool(_p); }
  bool operator!() const { return !_p; }
 private:
  std::shared_ptr<T> _p;
};
}  // namespace boost
#endif
