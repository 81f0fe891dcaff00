// Text field -> number, with the exact results of the converters the reference applies in
// read_csv_block (nums/core/systems/filesystem.py:157-212):
//
//   float dtypes : floatconv(x) = float(x)      (:163-167)  correctly rounded decimal -> binary64
//   np.int64     : np.int64(x)                  (:173-174)  integer literal
//   other ints   : int(float(x))                (:175-176)  truncation toward zero of the double
//
// Python's float() grammar: optional surrounding whitespace, optional sign, then "inf" / "infinity" /
// "nan" (any case) or digits with at most one '.', optional exponent, single underscores between
// digits.  Hexadecimal literals (the reference routes '0x' to float.fromhex) are reported as
// unsupported.  The decimal -> binary64 conversion is Clinger's exact fast path for small inputs and
// the Eisel-Lemire algorithm otherwise (D. Lemire, "Number Parsing at a Gigabyte per Second", SPE
// 2021; that the 128-bit product never needs a fallback for 64-bit significands is proven in
// Mushtak & Lemire, "Fast Number Parsing Without Fallback", SPE 2023).  Inputs with more than 19
// significant digits are truncated and converted twice (w and w + 1); when both agree the result is
// exact, otherwise the field is reported as unsupported rather than guessed.
//
// The functions are __host__ __device__: tests/csv_host_check.cpp compiles them for the CPU and
// compares them with strtod on tens of millions of strings.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define NUMS_HD __host__ __device__ __forceinline__
#else
#define NUMS_HD inline
#endif

namespace nums {
namespace csv {

enum FieldStatus : int {
  FIELD_OK = 0,
  FIELD_INVALID = 1,      // Python raises ValueError
  FIELD_UNSUPPORTED = 2,  // valid for Python, not handled here (hex literal, undecidable > 19 digits)
};

struct Pow5 {
  uint64_t hi, lo;
};

#if defined(__CUDA_ARCH__)
#define NUMS_CSV_TABLE __device__
#else
#define NUMS_CSV_TABLE
#endif
constexpr int kPow5Low = -342, kPow5High = 308;
// one copy per compilation mode: device code reads the __device__ array, host code the host array
#if defined(__CUDACC__)
static __device__ const Pow5 kPow5Device[kPow5High - kPow5Low + 1] = {
#include "pow5_table.inc"
};
#endif
static const Pow5 kPow5Host[kPow5High - kPow5Low + 1] = {
#include "pow5_table.inc"
};

// exact powers of ten for Clinger's fast path (10^22 is the largest power of ten a double holds exactly)
#define NUMS_POW10_LIST 1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12, 1e13, 1e14, 1e15, 1e16, \
                        1e17, 1e18, 1e19, 1e20, 1e21, 1e22
#if defined(__CUDACC__)
static __device__ const double kPow10Device[23] = {NUMS_POW10_LIST};
#endif
static const double kPow10Host[23] = {NUMS_POW10_LIST};

NUMS_HD double pow10_exact(int e) {
#if defined(__CUDA_ARCH__)
  return kPow10Device[e];
#else
  return kPow10Host[e];
#endif
}

NUMS_HD Pow5 pow5(int q) {
#if defined(__CUDA_ARCH__)
  return kPow5Device[q - kPow5Low];
#else
  return kPow5Host[q - kPow5Low];
#endif
}

NUMS_HD void mul64(uint64_t a, uint64_t b, uint64_t* hi, uint64_t* lo) {
#if defined(__CUDA_ARCH__)
  *lo = a * b;
  *hi = __umul64hi(a, b);
#else
  const unsigned __int128 p = (unsigned __int128)a * b;
  *lo = (uint64_t)p;
  *hi = (uint64_t)(p >> 64);
#endif
}

NUMS_HD int clz64(uint64_t x) {
#if defined(__CUDA_ARCH__)
  return __clzll((long long)x);
#else
  return __builtin_clzll(x);
#endif
}

NUMS_HD double bits_to_double(uint64_t bits) {
#if defined(__CUDA_ARCH__)
  return __longlong_as_double((long long)bits);
#else
  double d;
  __builtin_memcpy(&d, &bits, sizeof(d));
  return d;
#endif
}

// w * 10^q (w != 0) -> binary64 bit pattern without the sign (Eisel-Lemire).
NUMS_HD uint64_t decimal_to_bits(uint64_t w, int q) {
  if (q < kPow5Low) return 0;                          // underflows to zero for every 64-bit w
  if (q > kPow5High) return 0x7FF0000000000000ULL;     // overflows to infinity
  const int lz = clz64(w);
  w <<= lz;
  const Pow5 t = pow5(q);
  uint64_t hi, lo;
  mul64(w, t.hi, &hi, &lo);
  if ((hi & 0x1FFu) == 0x1FFu) {   // the 55 bits we keep could still be changed by the low table word
    uint64_t hi2, lo2;
    mul64(w, t.lo, &hi2, &lo2);
    lo += hi2;
    if (hi2 > lo) ++hi;
  }
  const int upper = (int)(hi >> 63);
  const int shift = upper + 64 - 52 - 3;
  uint64_t m = hi >> shift;
  // binary exponent of 10^q's leading bit: floor(q * log2(10)) + 63, log2(10) ~ 217706 / 65536
  int power2 = (int)(((int64_t)217706 * q) >> 16) + 63 + upper - lz + 1023;
  if (power2 <= 0) {               // subnormal result
    if (-power2 + 1 >= 64) return 0;
    m >>= -power2 + 1;
    m += m & 1;
    m >>= 1;
    return m;                      // exponent field 0, or 1 when the rounding carried into bit 52
  }
  // exactly half-way between two doubles: only possible for small |q|; round to even
  if (lo <= 1 && q >= -4 && q <= 23 && (m & 3) == 1 && (m << shift) == hi) m &= ~(uint64_t)1;
  m += m & 1;
  m >>= 1;
  if (m >= ((uint64_t)2 << 52)) {
    m = (uint64_t)1 << 52;
    ++power2;
  }
  m &= ~((uint64_t)1 << 52);
  if (power2 >= 0x7FF) return 0x7FF0000000000000ULL;
  return ((uint64_t)power2 << 52) | m;
}

NUMS_HD bool is_space(uint8_t c) {   // str.isspace() for the ASCII range
  return c == ' ' || (c >= 9 && c <= 13) || (c >= 0x1c && c <= 0x1f);
}
NUMS_HD bool is_digit(uint8_t c) { return c >= '0' && c <= '9'; }
NUMS_HD uint8_t lower(uint8_t c) { return (c >= 'A' && c <= 'Z') ? (uint8_t)(c + 32) : c; }

NUMS_HD bool matches_word(const uint8_t* p, int n, const char* word, int len) {
  if (n != len) return false;
  for (int i = 0; i < len; ++i)
    if (lower(p[i]) != (uint8_t)word[i]) return false;
  return true;
}

// float(text[0:n]) -> *out.  Returns a FieldStatus.
NUMS_HD int parse_float(const uint8_t* p, int n, double* out) {
  while (n > 0 && is_space(p[0])) { ++p; --n; }
  while (n > 0 && is_space(p[n - 1])) --n;
  if (n == 0) return FIELD_INVALID;
  bool negative = false;
  if (p[0] == '+' || p[0] == '-') {
    negative = p[0] == '-';
    ++p; --n;
    if (n == 0) return FIELD_INVALID;
  }
  const uint64_t sign = negative ? 0x8000000000000000ULL : 0;
  if (!is_digit(p[0]) && p[0] != '.') {
    if (matches_word(p, n, "inf", 3) || matches_word(p, n, "infinity", 8)) {
      *out = bits_to_double(sign | 0x7FF0000000000000ULL);
      return FIELD_OK;
    }
    if (matches_word(p, n, "nan", 3)) {
      *out = bits_to_double(sign | 0x7FF8000000000000ULL);
      return FIELD_OK;
    }
    return FIELD_INVALID;
  }
  if (n >= 2 && p[0] == '0' && lower(p[1]) == 'x') return FIELD_UNSUPPORTED;   // float.fromhex route

  uint64_t w = 0;
  int digits = 0;            // significant digits accumulated into w (<= 19)
  int64_t exp10 = 0;         // decimal exponent to apply to w
  bool truncated = false;    // a non-zero digit did not fit into w
  bool any_digit = false, seen_point = false, prev_digit = false;
  int i = 0;
  for (; i < n; ++i) {
    const uint8_t c = p[i];
    if (is_digit(c)) {
      any_digit = true;
      prev_digit = true;
      const int d = c - '0';
      if (digits < 19) {
        if (w != 0 || d != 0) {       // leading zeros are not significant
          w = w * 10 + (uint64_t)d;
          ++digits;
        }
        if (seen_point) --exp10;
      } else {
        if (d != 0) truncated = true;
        if (!seen_point) ++exp10;
      }
    } else if (c == '.') {
      if (seen_point) return FIELD_INVALID;
      seen_point = true;
      prev_digit = false;
    } else if (c == '_') {             // only between two digits
      if (!prev_digit || i + 1 >= n || !is_digit(p[i + 1])) return FIELD_INVALID;
      prev_digit = false;
    } else {
      break;
    }
  }
  if (!any_digit) return FIELD_INVALID;
  if (i < n) {
    if (lower(p[i]) != 'e') return FIELD_INVALID;
    ++i;
    bool exp_negative = false;
    if (i < n && (p[i] == '+' || p[i] == '-')) {
      exp_negative = p[i] == '-';
      ++i;
    }
    if (i >= n || !is_digit(p[i])) return FIELD_INVALID;
    int64_t e = 0;
    prev_digit = false;
    for (; i < n; ++i) {
      const uint8_t c = p[i];
      if (is_digit(c)) {
        if (e < 100000000) e = e * 10 + (c - '0');
        prev_digit = true;
      } else if (c == '_') {
        if (!prev_digit || i + 1 >= n || !is_digit(p[i + 1])) return FIELD_INVALID;
        prev_digit = false;
      } else {
        return FIELD_INVALID;
      }
    }
    exp10 += exp_negative ? -e : e;
  }
  if (w == 0) {
    *out = bits_to_double(sign);
    return FIELD_OK;
  }
  if (exp10 < -100000) exp10 = -100000;
  if (exp10 > 100000) exp10 = 100000;
  const int q = (int)exp10;
  if (!truncated && w <= ((uint64_t)1 << 53) && q >= -22 && q <= 22) {
    // Clinger: both operands exact, one correctly rounded operation
    double d = (double)w;
    d = q < 0 ? d / pow10_exact(-q) : d * pow10_exact(q);
    *out = negative ? -d : d;
    return FIELD_OK;
  }
  const uint64_t bits = decimal_to_bits(w, q);
  if (truncated && decimal_to_bits(w + 1, q) != bits) return FIELD_UNSUPPORTED;
  *out = bits_to_double(sign | bits);
  return FIELD_OK;
}

// np.int64(text): optional whitespace, sign, decimal digits with single underscores.
NUMS_HD int parse_int64(const uint8_t* p, int n, int64_t* out) {
  while (n > 0 && is_space(p[0])) { ++p; --n; }
  while (n > 0 && is_space(p[n - 1])) --n;
  if (n == 0) return FIELD_INVALID;
  bool negative = false;
  if (p[0] == '+' || p[0] == '-') {
    negative = p[0] == '-';
    ++p; --n;
  }
  if (n == 0 || !is_digit(p[0])) return FIELD_INVALID;
  uint64_t v = 0;
  bool prev_digit = false;
  for (int i = 0; i < n; ++i) {
    const uint8_t c = p[i];
    if (is_digit(c)) {
      const uint64_t d = (uint64_t)(c - '0');
      if (v > (0x8000000000000000ULL - d) / 10) return FIELD_INVALID;   // OverflowError in NumPy
      v = v * 10 + d;
      prev_digit = true;
    } else if (c == '_') {
      if (!prev_digit || i + 1 >= n || !is_digit(p[i + 1])) return FIELD_INVALID;
      prev_digit = false;
    } else {
      return FIELD_INVALID;
    }
  }
  if (!negative && v > 0x7FFFFFFFFFFFFFFFULL) return FIELD_INVALID;
  *out = negative ? (int64_t)(0 - v) : (int64_t)v;
  return FIELD_OK;
}

}  // namespace csv
}  // namespace nums
