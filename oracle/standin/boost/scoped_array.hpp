// TEST INFRASTRUCTURE ONLY (oracle). Minimal stand-in for boost::scoped_array.
#ifndef ORACLE_STANDIN_BOOST_SCOPED_ARRAY_HPP
#define ORACLE_STANDIN_BOOST_SCOPED_ARRAY_HPP
#include <cstddef>
#include <memory>
namespace boost {
template <class T> class scoped_array {
 public:
  scoped_array() {}
  explicit scoped_array(T *p) : _p(p) {}
  void reset(T *p = nullptr) { _p.reset(p); }
  T *get() const { return _p.get(); }
  T &operator[](std::ptrdiff_t i) const { return _p[size_t(i)]; }
 private:
  scoped_array(const scoped_array &);
  scoped_array &operator=(const scoped_array &);
  std::unique_ptr<T[]> _p;
};
}  // namespace boost
#endif
