// TEST INFRASTRUCTURE ONLY (oracle). Minimal stand-in for boost::tuple<double,double,double>
// with the member get<N>() that /root/reference/src/mcarray/ArrayDescription.cpp:69-71,198-205 uses.
#ifndef ORACLE_STANDIN_BOOST_TUPLE_HPP
#define ORACLE_STANDIN_BOOST_TUPLE_HPP
#include <cstddef>
#include <string>
#include <tuple>
#include <vector>
namespace boost {
template <class... Ts> class tuple {
 public:
  tuple() : _t() {}
  tuple(const Ts &... v) : _t(v...) {}
  template <int I> typename std::tuple_element<static_cast<std::size_t>(I), std::tuple<Ts...>>::type &get() { return std::get<static_cast<std::size_t>(I)>(_t); }
  template <int I> const typename std::tuple_element<static_cast<std::size_t>(I), std::tuple<Ts...>>::type &get() const { return std::get<static_cast<std::size_t>(I)>(_t); }
 private:
  std::tuple<Ts...> _t;
};
}  // namespace boost
#endif
