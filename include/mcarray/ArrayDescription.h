/* mca::ArrayDescription — microphone ids -> (x, y, z), names, distances.  Host-only geometry with the public surface of
 * the reference class (include/mcarray/ArrayDescription.h:31-72, src/mcarray/ArrayDescription.cpp:41-301): pushPosition
 * returns consecutive ids, duplicate names throw, distance() is the Euclidean distance (ArrayDescription.cpp:57-64),
 * getBandwidth() = c / (2 maxDistance).  ArrayPosition is a plain struct instead of boost::tuple. */
#ifndef MCARRAY_B200_ARRAYDESCRIPTION_H
#define MCARRAY_B200_ARRAYDESCRIPTION_H

#include <mcarray/mcarray_exception.h>

#include <cmath>
#include <map>
#include <ostream>
#include <string>
#include <vector>

namespace mca {

class ArrayDescription {
 public:
  struct ArrayPosition {
    double x, y, z;
    ArrayPosition() : x(0), y(0), z(0) {}
    ArrayPosition(double x_, double y_, double z_) : x(x_), y(y_), z(z_) {}
  };
  typedef int ElementId;

  ArrayDescription() {}
  virtual ~ArrayDescription() {}

  ElementId pushPosition(double x, double y, double z, const std::string &name = "") { return pushPosition(ArrayPosition(x, y, z), name); }
  ElementId pushPosition(const ArrayPosition &position, const std::string &name = "") {
    const ElementId id = static_cast<ElementId>(_list.size());
    const std::string name_ = name.empty() ? std::to_string(id) : name;   // unnamed elements are called by their id (ArrayDescription.cpp:110-112)
    if (_names.count(name_)) throw MCArrayException("Array element name aready used");   // message as ArrayDescription.cpp:115
    _names[name_] = id;
    _list.push_back(position);
    return id;
  }

  size_t size() const { return _list.size(); }
  bool empty() const { return _list.empty(); }

  double distance(ElementId i, ElementId j) const {
    if (!valid(i) || !valid(j)) return 0;
    const ArrayPosition &a = _list[i], &b = _list[j];
    return std::sqrt(std::pow(b.x - a.x, 2) + std::pow(b.y - a.y, 2) + std::pow(b.z - a.z, 2));
  }
  double distance(const std::string &i, const std::string &j) const { return distance(getId(i), getId(j)); }
  double maxDistance() const {
    double m = 0;
    for (size_t i = 0; i < _list.size(); ++i)
      for (size_t j = i + 1; j < _list.size(); ++j) m = std::max(m, distance(ElementId(i), ElementId(j)));
    return m;
  }
  /** smallest distance between two different elements (the reference's version always returns 0: SURVEY.md §8c, not reproduced) */
  double minDistance() const {
    double m = 0;
    bool first = true;
    for (size_t i = 0; i < _list.size(); ++i)
      for (size_t j = i + 1; j < _list.size(); ++j) {
        const double d = distance(ElementId(i), ElementId(j));
        if (first || d < m) { m = d; first = false; }
      }
    return m;
  }

  void getPosition(ElementId id, ArrayPosition &position) const { if (valid(id)) position = _list[id]; }
  void getPosition(const std::string &name, ArrayPosition &position) const { getPosition(getId(name), position); }

  double getX(const ElementId &id) const { return valid(id) ? _list[id].x : 0; }
  double getY(const ElementId &id) const { return valid(id) ? _list[id].y : 0; }
  double getZ(const ElementId &id) const { return valid(id) ? _list[id].z : 0; }
  void getX(std::vector<double> &x) const { x.clear(); for (size_t i = 0; i < _list.size(); ++i) x.push_back(_list[i].x); }
  void getY(std::vector<double> &y) const { y.clear(); for (size_t i = 0; i < _list.size(); ++i) y.push_back(_list[i].y); }
  void getZ(std::vector<double> &z) const { z.clear(); for (size_t i = 0; i < _list.size(); ++i) z.push_back(_list[i].z); }
  double getX(const std::string &name) const { return getX(getId(name)); }
  double getY(const std::string &name) const { return getY(getId(name)); }
  double getZ(const std::string &name) const { return getZ(getId(name)); }

  std::string getName(ElementId id) const {
    for (std::map<std::string, int>::const_iterator it = _names.begin(); it != _names.end(); ++it)
      if (it->second == id) return it->first;
    return "";
  }
  ElementId getId(const std::string &name) const {
    std::map<std::string, int>::const_iterator it = _names.find(name);
    return it == _names.end() ? -1 : it->second;
  }

  /** spatial-aliasing limit c / (2 d_max), c = 346.1 m/s (microhponeArrayHelpers.cpp:38-43) */
  double getBandwidth() const { const double d = maxDistance(); return d > 0 ? 346.1 / (2 * d) : 0; }

  static ArrayDescription make_linear_array_description(const std::vector<double> &x) {
    ArrayDescription a;
    for (size_t i = 0; i < x.size(); ++i) a.pushPosition(x[i], 0, 0);
    return a;
  }

  /** [M][3] row-major coordinates in id order: the layout the C ABI takes */
  std::vector<double> xyz() const {
    std::vector<double> v;
    for (size_t i = 0; i < _list.size(); ++i) { v.push_back(_list[i].x); v.push_back(_list[i].y); v.push_back(_list[i].z); }
    return v;
  }

 private:
  bool valid(ElementId id) const { return id >= 0 && static_cast<size_t>(id) < _list.size(); }
  std::map<std::string, int> _names;
  std::vector<ArrayPosition> _list;
};

inline std::ostream &operator<<(std::ostream &os, const ArrayDescription &d) {
  for (size_t i = 0; i < d.size(); ++i)
    os << "[" << i << " " << d.getName(int(i)) << ": " << d.getX(int(i)) << ", " << d.getY(int(i)) << ", " << d.getZ(int(i)) << "] ";
  return os;
}

}  // namespace mca

#endif
