// Drop-in counterpart of common/Point.h:18-128 of zivy/LSQRRecipes (re-authored, VNL-free).
// Layout contract: exactly n values of T (sizeof(Point<double,3>) == 24), see SURVEY.md 8a-11.
#ifndef LSQR_B200_POINT_H
#define LSQR_B200_POINT_H
#include <cstring>
#include <ostream>

#include "Vector.h"

namespace lsqrRecipes {

template <class T, unsigned int n>
class Point {
 public:
  enum { dimension = n };
  Point() { std::memset(data, 0, sizeof(data)); }
  Point(T* fillData) { std::memcpy(data, fillData, sizeof(data)); }
  Point(const Point& other) { std::memcpy(data, other.data, sizeof(data)); }
  Point& operator=(const Point& other) { std::memcpy(data, other.data, sizeof(data)); return *this; }

  T& operator[](int index) { return data[index]; }
  const T& operator[](int index) const { return data[index]; }
  void set(T* fillData) { std::memcpy(data, fillData, sizeof(data)); }
  unsigned int size() { return n; }

  double distanceSquared(const Point& other) {
    double s = 0;
    for (unsigned int i = 0; i < n; i++) { const double d = data[i] - other.data[i]; s += d * d; }
    return s;
  }
  Point operator+(const Vector<T, n>& v) { Point r(*this); for (unsigned int i = 0; i < n; i++) r.data[i] += v[i]; return r; }
  Point operator-(const Vector<T, n>& v) { Point r(*this); for (unsigned int i = 0; i < n; i++) r.data[i] -= v[i]; return r; }
  Vector<T, n> operator-(const Point& p) { Vector<T, n> r(data); for (unsigned int i = 0; i < n; i++) r[i] -= p.data[i]; return r; }

  friend std::ostream& operator<<(std::ostream& out, const Point& p) {
    out << "[ " << p.data[0];
    for (unsigned int i = 1; i < n; i++) out << ", " << p.data[i];
    return out << " ]";
  }

 private:
  T data[n];
};

}  // namespace lsqrRecipes
#endif
