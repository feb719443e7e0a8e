// TEST INFRASTRUCTURE ONLY (oracle). Minimal stand-in for boost::scoped_ptr.
#ifndef ORACLE_STANDIN_BOOST_SCOPED_PTR_HPP
#define ORACLE_STANDIN_BOOST_SCOPED_PTR_HPP
#include <memory>
namespace boost { template <class T> using scoped_ptr = std::unique_ptr<T>; }
#endif
