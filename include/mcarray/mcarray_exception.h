/* MCArrayException : std::runtime_error — thrown for configuration errors only, like the reference
 * (include/mcarray/mcarray_exception.h:51-56; FastBinauralMasking.cpp:88-91, ArrayDescription.cpp:118-119).  Status codes
 * of the C ABI are turned into this exception by the C++ wrappers; nothing throws across the C boundary. */
#ifndef MCARRAY_B200_EXCEPTION_H
#define MCARRAY_B200_EXCEPTION_H

#include <stdexcept>
#include <string>

namespace mca {

class MCArrayException : public std::runtime_error {
 public:
  explicit MCArrayException(const std::string &msg) : std::runtime_error(msg) {}
};

}  // namespace mca

#endif
