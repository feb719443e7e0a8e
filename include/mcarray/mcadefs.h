/* mcarray (B200 build) — base types.  Same names as the reference's include/mcarray/mcadefs.h:82-89 so user code compiles
 * unchanged: BaseType is double at the API (the GPU path computes in fp32 behind the C ABI), SignalPtr is a shared array
 * of doubles, SignalVector one buffer per channel.  The reference uses boost::shared_array; Boost is picked up when it is
 * installed, otherwise an equivalent minimal shared array is used. */
#ifndef MCARRAY_B200_MCADEFS_H
#define MCARRAY_B200_MCADEFS_H

#include <stdint.h>

#include <cstddef>
#include <memory>
#include <vector>

#if defined(__has_include)
#if __has_include(<boost/shared_array.hpp>) && !defined(MCARRAY_NO_BOOST)
#include <boost/shared_array.hpp>
#define MCARRAY_HAVE_BOOST_SHARED_ARRAY 1
#endif
#endif

namespace mca {

#ifdef MCARRAY_HAVE_BOOST_SHARED_ARRAY
template <class T> using shared_array = boost::shared_array<T>;
#else
/** shared ownership of a new[]-allocated array (the subset of boost::shared_array the mcarray API uses) */
template <class T> class shared_array {
 public:
  shared_array() {}
  explicit shared_array(T *p) : _p(p, std::default_delete<T[]>()) {}
  template <class D> shared_array(T *p, D d) : _p(p, d) {}
  void reset() { _p.reset(); }
  void reset(T *p) { _p.reset(p, std::default_delete<T[]>()); }
  template <class D> void reset(T *p, D d) { _p.reset(p, d); }
  T &operator[](std::ptrdiff_t i) const { return _p.get()[i]; }
  T *get() const { return _p.get(); }
  explicit operator bool() const { return static_cast<bool>(_p); }
  long use_count() const { return _p.use_count(); }
 private:
  std::shared_ptr<T> _p;
};
#endif

typedef float BaseType32;
typedef double BaseType64;
typedef signed short BaseType16s;
typedef shared_array<BaseType32> SignalPtr32;
typedef shared_array<BaseType64> SignalPtr64;
typedef shared_array<BaseType16s> SignalPtr16s;
typedef std::vector<SignalPtr32> SignalVector32;
typedef std::vector<SignalPtr64> SignalVector64;
typedef std::vector<SignalPtr16s> SignalVector16s;

typedef BaseType64 BaseType;
typedef shared_array<BaseType> BaseTypePtr;
typedef shared_array<BaseType> SignalPtr;
typedef std::vector<SignalPtr> SignalVector;

}  // namespace mca

#endif
