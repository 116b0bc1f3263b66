// sycl_ext_complex.hpp — host stand-in for the third-party SyclCPLX complex type
// (userspace/3rdparty/SyclCPLX/include/sycl_ext_complex.hpp). Arithmetic follows that header for
// finite operands: operator* is (ac - bd, ad + bc) from four separately rounded products
// (:569-623), complex / scalar divides both parts, norm = re*re + im*im, conj flips the sign.
#pragma once
#include <cmath>

#ifndef _SYCL_CPLX_NAMESPACE
#define _SYCL_CPLX_NAMESPACE sycl::ext::cplx
#endif

namespace _SYCL_CPLX_NAMESPACE {

template <class T>
class complex {
  T re_, im_;

 public:
  using value_type = T;
  constexpr complex(T re = T{}, T im = T{}) : re_{re}, im_{im} {}
  template <class U>
  constexpr complex(const complex<U>& o) : re_{static_cast<T>(o.real())}, im_{static_cast<T>(o.imag())} {}
  constexpr T real() const { return re_; }
  constexpr T imag() const { return im_; }
  constexpr void real(T v) { re_ = v; }
  constexpr void imag(T v) { im_ = v; }
  constexpr complex& operator+=(const complex& o) { re_ += o.re_; im_ += o.im_; return *this; }
  constexpr complex& operator-=(const complex& o) { re_ -= o.re_; im_ -= o.im_; return *this; }
  constexpr complex& operator*=(const complex& o) { *this = *this * o; return *this; }
  constexpr complex& operator*=(T s) { re_ *= s; im_ *= s; return *this; }
  constexpr complex& operator/=(T s) { re_ /= s; im_ /= s; return *this; }
  friend constexpr complex operator+(const complex& a, const complex& b) { return {a.re_ + b.re_, a.im_ + b.im_}; }
  friend constexpr complex operator-(const complex& a, const complex& b) { return {a.re_ - b.re_, a.im_ - b.im_}; }
  friend constexpr complex operator-(const complex& a) { return {-a.re_, -a.im_}; }
  friend constexpr complex operator*(const complex& z, const complex& w) {
    const T a = z.re_, b = z.im_, c = w.re_, d = w.im_;
    const T ac = a * c, bd = b * d, ad = a * d, bc = b * c;
    return {ac - bd, ad + bc};
  }
  friend constexpr complex operator*(const complex& a, T s) { return {a.re_ * s, a.im_ * s}; }
  friend constexpr complex operator*(T s, const complex& a) { return {a.re_ * s, a.im_ * s}; }
  friend constexpr complex operator/(const complex& a, T s) { return {a.re_ / s, a.im_ / s}; }
  friend constexpr bool operator==(const complex& a, const complex& b) { return a.re_ == b.re_ && a.im_ == b.im_; }
  friend constexpr bool operator!=(const complex& a, const complex& b) { return !(a == b); }
};

template <class T>
constexpr T norm(const complex<T>& c) { return c.real() * c.real() + c.imag() * c.imag(); }
template <class T>
constexpr complex<T> conj(const complex<T>& c) { return {c.real(), -c.imag()}; }
template <class T>
inline T abs(const complex<T>& c) { return std::hypot(c.real(), c.imag()); }

}  // namespace _SYCL_CPLX_NAMESPACE
