// Drop-in counterpart of common/Vector.h:18-158 of zivy/LSQRRecipes (re-authored, VNL-free).
// Layout contract: exactly n values of T, nothing else -- this is what crosses
// the C ABI as a record of the host AoS layout.
#ifndef LSQR_B200_VECTOR_H
#define LSQR_B200_VECTOR_H
#include <cmath>
#include <cstring>
#include <ostream>

namespace lsqrRecipes {

template <class T, unsigned int n>
class Vector {
 public:
  enum { dimension = n };
  Vector() { std::memset(data, 0, sizeof(data)); }
  Vector(T* fillData) { std::memcpy(data, fillData, sizeof(data)); }
  Vector(const Vector& other) { std::memcpy(data, other.data, sizeof(data)); }
  Vector& operator=(const Vector& other) { std::memcpy(data, other.data, sizeof(data)); return *this; }

  T& operator[](int index) { return data[index]; }
  const T& operator[](int index) const { return data[index]; }
  void set(T* fillData) { std::memcpy(data, fillData, sizeof(data)); }
  unsigned int size() { return n; }

  Vector operator*(const T& scalar) const { Vector r(*this); for (unsigned int i = 0; i < n; i++) r.data[i] *= scalar; return r; }
  T operator*(const Vector& right) const { T s = 0; for (unsigned int i = 0; i < n; i++) s += data[i] * right.data[i]; return s; }
  Vector operator+(const Vector& right) const { Vector r(*this); for (unsigned int i = 0; i < n; i++) r.data[i] += right.data[i]; return r; }
  Vector operator-(const Vector& right) const { Vector r(*this); for (unsigned int i = 0; i < n; i++) r.data[i] -= right.data[i]; return r; }

  T l2Norm() { T s = 0; for (unsigned int i = 0; i < n; i++) s += data[i] * data[i]; return std::sqrt(s); }
  void normalize() { const T norm = l2Norm(); for (unsigned int i = 0; i < n; i++) data[i] /= norm; }

  friend std::ostream& operator<<(std::ostream& out, const Vector& v) {
    out << "[ " << v.data[0];
    for (unsigned int i = 1; i < n; i++) out << ", " << v.data[i];
    return out << " ]";
  }

 private:
  T data[n];
};

template <class T, unsigned int n>
inline Vector<T, n> operator*(const T& s, const Vector<T, n>& v) { return v * s; }

}  // namespace lsqrRecipes
#endif
