// CPU check of nums_b200/csrc/csv_parse.cuh against the C library's correctly rounded strtod
// (what Python's float() is specified to agree with).  Built and run by tests/test_csv_host.py.
//   usage: csv_host_check <cases per family> <seed>
#include <cinttypes>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>

#include "../../nums_b200/csrc/csv_parse.cuh"

using nums::csv::parse_float;

static uint64_t bits_of(double d) {
  uint64_t b;
  memcpy(&b, &d, 8);
  return b;
}

static long failures = 0, unsupported = 0, checked = 0;

static void check(const std::string& s, bool allow_unsupported) {
  double got = 0.0;
  const int st = parse_float(reinterpret_cast<const uint8_t*>(s.data()), (int)s.size(), &got);
  ++checked;
  if (st == nums::csv::FIELD_UNSUPPORTED && allow_unsupported) {
    ++unsupported;
    return;
  }
  char* end = nullptr;
  const double want = strtod(s.c_str(), &end);
  if (st != nums::csv::FIELD_OK || bits_of(got) != bits_of(want)) {
    if (++failures <= 20)
      fprintf(stderr, "MISMATCH '%s': status %d got %.17g (%016" PRIx64 ") want %.17g (%016" PRIx64 ")\n", s.c_str(), st,
              got, bits_of(got), want, bits_of(want));
  }
}

int main(int argc, char** argv) {
  const long n = argc > 1 ? atol(argv[1]) : 1000000;
  std::mt19937_64 rng(argc > 2 ? atoll(argv[2]) : 12345);
  char buf[128];
  // 1. random digit strings with a decimal point somewhere and a random exponent
  for (long i = 0; i < n; ++i) {
    const int nd = 1 + (int)(rng() % 19);
    std::string s;
    if (rng() & 1) s += (rng() & 1) ? '-' : '+';
    const int point = (int)(rng() % (nd + 1));
    for (int k = 0; k < nd; ++k) {
      if (k == point && (k > 0 || (rng() & 1))) s += '.';
      s += (char)('0' + rng() % 10);
    }
    if (rng() % 4) {
      snprintf(buf, sizeof buf, "%c%d", (rng() & 1) ? 'e' : 'E', (int)(rng() % 700) - 360);
      s += buf;
    }
    check(s, false);
  }
  // 2./3. printed doubles: np.savetxt's default %.18e, shortest round trip %.17g, short %.6f
  for (long i = 0; i < n; ++i) {
    uint64_t b = rng();
    double d;
    memcpy(&d, &b, 8);
    if (!std::isfinite(d)) continue;
    snprintf(buf, sizeof buf, "%.18e", d);
    check(buf, false);
    snprintf(buf, sizeof buf, "%.17g", d);
    check(buf, false);
    const double u = (double)(rng() >> 11) / 9007199254740992.0 * 200.0 - 100.0;
    snprintf(buf, sizeof buf, "%.6f", u);
    check(buf, false);
    snprintf(buf, sizeof buf, "%.18e", u);
    check(buf, false);
  }
  // 4. more than 19 significant digits: exact when decidable, otherwise reported
  for (long i = 0; i < n / 4; ++i) {
    uint64_t b = rng();
    double d;
    memcpy(&d, &b, 8);
    if (!std::isfinite(d)) continue;
    snprintf(buf, sizeof buf, "%.30e", d);
    check(buf, true);
  }
  // 5. subnormals, extremes, half-way cases
  const char* fixed[] = {"4.9406564584124654e-324", "2.4703282292062327e-324", "2.4703282292062328e-324", "1e-400",
                         "1.7976931348623157e308", "1.7976931348623159e308", "1e309", "2.2250738585072014e-308",
                         "2.2250738585072011e-308", "9007199254740993", "9007199254740992", "9007199254740995",
                         "0.1", "0.30000000000000004", "123456789012345678", "1e23", "8.5e22", "1e22", "1e-22",
                         "0", "-0", "0.0e10", ".5", "5.", "1_0.2_5e1_0", "00012.5000", "1E5", "  12.5  ", "\t7\n",
                         "6.0221407600000000e+23", "1.000000000000000000e+00", "8.692932128906250000e-01"};
  for (const char* s : fixed) {
    std::string t(s);
    std::string plain;
    for (char c : t) if (c != '_') plain += c;       // strtod has no underscores
    double got = 0.0;
    const int st = parse_float(reinterpret_cast<const uint8_t*>(t.data()), (int)t.size(), &got);
    const double want = strtod(plain.c_str(), nullptr);
    ++checked;
    if (st != nums::csv::FIELD_OK || bits_of(got) != bits_of(want)) {
      ++failures;
      fprintf(stderr, "MISMATCH fixed '%s': status %d got %.17g want %.17g\n", s, st, got, want);
    }
  }
  // 6. specials and rejects
  struct { const char* s; int status; } odd[] = {{"nan", 0}, {"-NaN", 0}, {"inf", 0}, {"-Infinity", 0}, {"+INF", 0},
      {"", 1}, {" ", 1}, {"abc", 1}, {"1.2.3", 1}, {"1e", 1}, {"e5", 1}, {"1_", 1}, {"_1", 1}, {"1__0", 1}, {"--1", 1},
      {"1e+", 1}, {".", 1}, {"+", 1}, {"infin", 1}, {"1 2", 1}, {"0x10", 2}, {"1._5", 1}, {"1_.5", 1}};
  for (auto& o : odd) {
    double got = 0.0;
    const int st = parse_float(reinterpret_cast<const uint8_t*>(o.s), (int)strlen(o.s), &got);
    ++checked;
    if (st != o.status) {
      ++failures;
      fprintf(stderr, "STATUS '%s': got %d want %d\n", o.s, st, o.status);
    }
  }
  printf("checked %ld strings, %ld mismatches, %ld reported undecidable (> 19 digits)\n", checked, failures, unsupported);
  return failures ? 1 : 0;
}
