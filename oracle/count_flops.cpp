// count_flops.cpp -- op-counting build of the CPU oracle (test infrastructure).
//
// Compiles oracle/bio_oracle.c unchanged as C++ with `double` replaced by a
// counting scalar, so that the ALGORITHMIC floating-point work of one env step
// (+ - x / sqrt and every transcendental counted as 1 each) can be
// measured instead of estimated.  Used by oracle/count_flops.py to write
// profiles/flop_count.json, which bench.py turns into roofline.achieved.
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static long long g_ops[6];  // 0 add/sub, 1 mul, 2 div, 3 sqrt, 4 transcendental (sin cos exp), 5 compare/abs/floor

struct cd {
    double v;
    cd() = default;
    cd(double x) : v(x) {}
    explicit operator double() const { return v; }
    explicit operator int() const { return (int)v; }
    explicit operator uint64_t() const { return (uint64_t)v; }
    cd& operator+=(const cd& o) { g_ops[0]++; v += o.v; return *this; }
    cd& operator-=(const cd& o) { g_ops[0]++; v -= o.v; return *this; }
    cd& operator*=(const cd& o) { g_ops[1]++; v *= o.v; return *this; }
    cd& operator/=(const cd& o) { g_ops[2]++; v /= o.v; return *this; }
};
static inline cd operator+(const cd& a, const cd& b) { g_ops[0]++; return cd(a.v + b.v); }
static inline cd operator-(const cd& a, const cd& b) { g_ops[0]++; return cd(a.v - b.v); }
static inline cd operator*(const cd& a, const cd& b) { g_ops[1]++; return cd(a.v * b.v); }
static inline cd operator/(const cd& a, const cd& b) { g_ops[2]++; return cd(a.v / b.v); }
static inline cd operator-(const cd& a) { return cd(-a.v); }
static inline bool operator<(const cd& a, const cd& b) { g_ops[5]++; return a.v < b.v; }
static inline bool operator>(const cd& a, const cd& b) { g_ops[5]++; return a.v > b.v; }
static inline bool operator<=(const cd& a, const cd& b) { g_ops[5]++; return a.v <= b.v; }
static inline bool operator>=(const cd& a, const cd& b) { g_ops[5]++; return a.v >= b.v; }
static inline bool operator==(const cd& a, const cd& b) { return a.v == b.v; }
static inline bool operator!=(const cd& a, const cd& b) { return a.v != b.v; }
static inline cd sqrt(const cd& a) { g_ops[3]++; return cd(::sqrt(a.v)); }
static inline cd sin(const cd& a) { g_ops[4]++; return cd(::sin(a.v)); }
static inline cd cos(const cd& a) { g_ops[4]++; return cd(::cos(a.v)); }
static inline cd exp(const cd& a) { g_ops[4]++; return cd(::exp(a.v)); }
static inline cd fabs(const cd& a) { g_ops[5]++; return cd(::fabs(a.v)); }
static inline cd floor(const cd& a) { g_ops[5]++; return cd(::floor(a.v)); }
static inline cd ceil(const cd& a) { g_ops[5]++; return cd(::ceil(a.v)); }
static inline cd pow(const cd& a, const cd& b) { g_ops[4]++; return cd(::pow(a.v, b.v)); }
static inline cd pow(const cd& a, double b) { g_ops[4]++; return cd(::pow(a.v, b)); }
static inline cd fmod(const cd& a, const cd& b) { g_ops[2]++; return cd(::fmod(a.v, b.v)); }
#undef isnan
#undef isfinite
static inline bool isnan(const cd& a) { return a.v != a.v; }
static inline bool isfinite(const cd& a) { return ::fabs(a.v) <= 1.79e308; }

extern "C" {
#define double cd
#include "bio_oracle.c"
#undef double

void orc_count_reset(void) { memset(g_ops, 0, sizeof g_ops); }
void orc_count_get(long long* out6) { memcpy(out6, g_ops, sizeof g_ops); }
}
