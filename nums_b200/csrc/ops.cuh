// Scalar semantics of the elementwise ufuncs (device functors).
//
// Semantics follow NumPy's inner loops (the reference calls np.<ufunc> at
// nums/core/systems/numpy_compute.py:184-186 and :233-238): IEEE arithmetic for
// floating point, wrap-around for integers, 0 for integer division by zero, Python-sign
// floor_divide/remainder, NaN-propagating maximum/minimum vs NaN-ignoring fmax/fmin.
#pragma once
#include <cmath>
#include "common.cuh"

namespace nums {
namespace op {

template <typename T> constexpr bool is_fp = std::is_floating_point<T>::value;
template <typename T> constexpr bool is_int = std::is_same<T, int32_t>::value || std::is_same<T, int64_t>::value;
template <typename T> constexpr bool is_bool = std::is_same<T, bool>::value;

template <typename T> __device__ __forceinline__ bool nan_(T v) {
  if constexpr (is_fp<T>) return v != v;
  else return false;
}

// npy_divmod: floor division and Python-sign modulus for floating point.
template <typename T>
__device__ __forceinline__ void fp_divmod(T a, T b, T& quo, T& rem) {
  T mod = fmod(a, b);
  if (b == T(0)) {  // NumPy: a / b for the quotient, fmod's NaN for the remainder
    quo = a / b;
    rem = mod;
    return;
  }
  T div = (a - mod) / b;
  if (mod != T(0)) {
    if ((b < T(0)) != (mod < T(0))) {
      mod += b;
      div -= T(1);
    }
  } else {
    mod = copysign(T(0), b);
  }
  T fl;
  if (div != T(0)) {
    fl = floor(div);
    if (div - fl > T(0.5)) fl += T(1);
  } else {
    fl = copysign(T(0), a / b);
  }
  quo = fl;
  rem = mod;
}

template <typename T> __device__ __forceinline__ T int_floor_div(T a, T b) {
  if (b == 0) return 0;
  if (b == T(-1)) return T(0) - a;  // also avoids the INT_MIN / -1 trap; wraps like NumPy
  T q = a / b;
  if ((a % b != 0) && ((a < 0) != (b < 0))) --q;
  return q;
}
template <typename T> __device__ __forceinline__ T int_py_mod(T a, T b) {
  if (b == 0) return 0;
  if (b == T(-1)) return 0;
  T r = a % b;
  if (r != 0 && ((r < 0) != (b < 0))) r += b;
  return r;
}
template <typename T> __device__ __forceinline__ T int_pow(T base, T e) {
  if (e < 0) return 0;  // NumPy raises ValueError for negative integer exponents
  using U = typename std::make_unsigned<T>::type;
  U result = 1, b = (U)base;
  while (e) {
    if (e & 1) result *= b;
    b *= b;
    e >>= 1;
  }
  return (T)result;
}
template <typename T> __device__ __forceinline__ T int_gcd(T a, T b) {
  using U = typename std::make_unsigned<T>::type;
  U x = a < 0 ? U(0) - (U)a : (U)a, y = b < 0 ? U(0) - (U)b : (U)b;
  while (y) {
    U t = x % y;
    x = y;
    y = t;
  }
  return (T)x;
}

#define NUMS_BINARY(NAME, OUT_T, ...)                                      \
  template <typename T> struct NAME {                                      \
    using Out = OUT_T;                                                     \
    static __device__ __forceinline__ Out apply(T a, T b) { __VA_ARGS__ }  \
  };

// ---- arithmetic -----------------------------------------------------------------------------------
NUMS_BINARY(Add, T, if constexpr (is_bool<T>) return a || b; else return a + b;)
NUMS_BINARY(Subtract, T, if constexpr (is_bool<T>) return a != b; else return a - b;)
NUMS_BINARY(Multiply, T, if constexpr (is_bool<T>) return a && b; else return a * b;)
NUMS_BINARY(TrueDivide, T, return a / b;)
NUMS_BINARY(FloorDivide, T,
            if constexpr (is_fp<T>) { T q; T r; fp_divmod(a, b, q, r); return q; }
            else return int_floor_div(a, b);)
NUMS_BINARY(Remainder, T,
            if constexpr (is_fp<T>) { T q; T r; fp_divmod(a, b, q, r); return r; }
            else return int_py_mod(a, b);)
NUMS_BINARY(Fmod, T,
            if constexpr (is_fp<T>) return fmod(a, b);
            else { if (b == 0) return 0; if (b == T(-1)) return 0; return a % b; })
NUMS_BINARY(Power, T, if constexpr (is_fp<T>) return pow(a, b); else return int_pow(a, b);)
NUMS_BINARY(Maximum, T,
            if constexpr (is_bool<T>) return a || b;
            else { if (nan_(a)) return a; if (nan_(b)) return b; return a >= b ? a : b; })
NUMS_BINARY(Minimum, T,
            if constexpr (is_bool<T>) return a && b;
            else { if (nan_(a)) return a; if (nan_(b)) return b; return a <= b ? a : b; })
NUMS_BINARY(Fmax, T,
            if constexpr (is_bool<T>) return a || b;
            else { if (nan_(b)) return a; if (nan_(a)) return b; return a >= b ? a : b; })
NUMS_BINARY(Fmin, T,
            if constexpr (is_bool<T>) return a && b;
            else { if (nan_(b)) return a; if (nan_(a)) return b; return a <= b ? a : b; })
// ---- floating-point only ---------------------------------------------------------------------------
NUMS_BINARY(Arctan2, T, return atan2(a, b);)
NUMS_BINARY(Hypot, T, return hypot(a, b);)
NUMS_BINARY(Copysign, T, return copysign(a, b);)
NUMS_BINARY(Nextafter, T, return nextafter(a, b);)
NUMS_BINARY(Heaviside, T,
            if (nan_(a)) return a; if (a == T(0)) return b; return a < T(0) ? T(0) : T(1);)
NUMS_BINARY(Logaddexp, T,
            if (a == b) return a + T(0.693147180559945309417232121458176568);
            T d = a - b;
            if (d > T(0)) return a + log1p(exp(-d));
            if (d <= T(0)) return b + log1p(exp(d));
            return d;)
NUMS_BINARY(Logaddexp2, T,
            if (a == b) return a + T(1);
            T d = a - b;
            const T log2e = T(1.442695040888963407359924681001892137);
            if (d > T(0)) return a + log2e * log1p(exp2(-d));
            if (d <= T(0)) return b + log2e * log1p(exp2(d));
            return d;)
NUMS_BINARY(Ldexp, T,
            { T e = b; if (e > T(65536)) e = T(65536); if (e < T(-65536)) e = T(-65536);
              return ldexp(a, (int)e); })
NUMS_BINARY(Xlogy, T, if (a == T(0) && !nan_(b)) return T(0); return a * log(b);)
// ---- comparisons / logic -----------------------------------------------------------------------------
NUMS_BINARY(Less, bool, return a < b;)
NUMS_BINARY(LessEqual, bool, return a <= b;)
NUMS_BINARY(Greater, bool, return a > b;)
NUMS_BINARY(GreaterEqual, bool, return a >= b;)
NUMS_BINARY(Equal, bool, return a == b;)
NUMS_BINARY(NotEqual, bool, return a != b;)
NUMS_BINARY(LogicalAnd, bool, return (a != T(0)) && (b != T(0));)
NUMS_BINARY(LogicalOr, bool, return (a != T(0)) || (b != T(0));)
NUMS_BINARY(LogicalXor, bool, return (a != T(0)) != (b != T(0));)
// ---- integer / bool bit ops ------------------------------------------------------------------------
NUMS_BINARY(BitAnd, T, if constexpr (is_bool<T>) return a && b; else return a & b;)
NUMS_BINARY(BitOr, T, if constexpr (is_bool<T>) return a || b; else return a | b;)
NUMS_BINARY(BitXor, T, if constexpr (is_bool<T>) return a != b; else return a ^ b;)
NUMS_BINARY(LeftShift, T,
            { using U = typename std::make_unsigned<T>::type;
              if (b < 0 || b >= T(sizeof(T) * 8)) return T(0);
              return (T)((U)a << b); })
NUMS_BINARY(RightShift, T,
            if (b < 0 || b >= T(sizeof(T) * 8)) return a < 0 ? T(-1) : T(0); return a >> b;)
NUMS_BINARY(Gcd, T, return int_gcd(a, b);)
NUMS_BINARY(Lcm, T,
            { T g = int_gcd(a, b); if (g == 0) return T(0);
              T q = a / g; T r = q * b; return r < 0 ? T(0) - r : r; })

#undef NUMS_BINARY

// ---- unary -------------------------------------------------------------------------------------------
#define NUMS_UNARY(NAME, OUT_T, ...)                                 \
  template <typename T> struct NAME {                                \
    using Out = OUT_T;                                               \
    static __device__ __forceinline__ Out apply(T a) { __VA_ARGS__ } \
  };

NUMS_UNARY(Copy, T, return a;)
NUMS_UNARY(Abs, T,
           if constexpr (is_fp<T>) return fabs(a);
           else if constexpr (is_bool<T>) return a;
           else return a < 0 ? T(0) - a : a;)
NUMS_UNARY(Negative, T, if constexpr (is_bool<T>) return !a; else return T(0) - a;)
NUMS_UNARY(Positive, T, return a;)
NUMS_UNARY(Sign, T,
           if constexpr (is_bool<T>) return a;
           else { if (nan_(a)) return a; return a > T(0) ? T(1) : (a < T(0) ? T(-1) : T(0)); })
NUMS_UNARY(Sqrt, T, return sqrt(a);)
NUMS_UNARY(Cbrt, T, return cbrt(a);)
NUMS_UNARY(Square, T, if constexpr (is_bool<T>) return a; else return a * a;)
NUMS_UNARY(Reciprocal, T,
           if constexpr (is_fp<T>) return T(1) / a;
           else if constexpr (is_bool<T>) return a;
           else return a == 0 ? T(0) : T(1) / a;)
NUMS_UNARY(Exp, T, return exp(a);)
NUMS_UNARY(Exp2, T, return exp2(a);)
NUMS_UNARY(Expm1, T, return expm1(a);)
NUMS_UNARY(Log, T, return log(a);)
NUMS_UNARY(Log2, T, return log2(a);)
NUMS_UNARY(Log10, T, return log10(a);)
NUMS_UNARY(Log1p, T, return log1p(a);)
NUMS_UNARY(Sin, T, return sin(a);)
NUMS_UNARY(Cos, T, return cos(a);)
NUMS_UNARY(Tan, T, return tan(a);)
NUMS_UNARY(Arcsin, T, return asin(a);)
NUMS_UNARY(Arccos, T, return acos(a);)
NUMS_UNARY(Arctan, T, return atan(a);)
NUMS_UNARY(Sinh, T, return sinh(a);)
NUMS_UNARY(Cosh, T, return cosh(a);)
NUMS_UNARY(Tanh, T, return tanh(a);)
NUMS_UNARY(Arcsinh, T, return asinh(a);)
NUMS_UNARY(Arccosh, T, return acosh(a);)
NUMS_UNARY(Arctanh, T, return atanh(a);)
NUMS_UNARY(Floor, T, if constexpr (is_fp<T>) return floor(a); else return a;)
NUMS_UNARY(Ceil, T, if constexpr (is_fp<T>) return ceil(a); else return a;)
NUMS_UNARY(Trunc, T, if constexpr (is_fp<T>) return trunc(a); else return a;)
NUMS_UNARY(Rint, T, if constexpr (is_fp<T>) return rint(a); else return a;)
NUMS_UNARY(Deg2rad, T, return a * T(0.017453292519943295769236907684886127);)
NUMS_UNARY(Rad2deg, T, return a * T(57.295779513082320876798154814105170332);)
NUMS_UNARY(Spacing, T,
           if (isinf(a)) return T(NAN); if (nan_(a)) return a;
           return nextafter(a, copysign(T(INFINITY), a)) - a;)
NUMS_UNARY(Isnan, bool, return nan_(a);)
NUMS_UNARY(Isinf, bool, if constexpr (is_fp<T>) return isinf(a); else return false;)
NUMS_UNARY(Isfinite, bool, if constexpr (is_fp<T>) return isfinite(a); else return true;)
NUMS_UNARY(Signbit, bool, if constexpr (is_fp<T>) return signbit(a); else return a < T(0);)
NUMS_UNARY(LogicalNot, bool, return a == T(0);)
NUMS_UNARY(Invert, T, if constexpr (is_bool<T>) return !a; else return ~a;)

#undef NUMS_UNARY

}  // namespace op
}  // namespace nums
