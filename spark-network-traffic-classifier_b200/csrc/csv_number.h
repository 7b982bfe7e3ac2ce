// CSV field grammar and exact decimal -> binary conversion for the device CSV reader (csv.cu; SURVEY.md 8f-3:
// `spark.read.csv(..., inferSchema=True)` at kdd99.py:25 and cicids17.py:19-20).
//
// Spark infers a column's type by trying, per field, Integer -> Long -> Double -> String  [recalled: Spark 2.4
// sql/execution/datasources/csv/CSVInferSchema.scala: tryParseInteger `field.toInt`, tryParseLong, tryParseDouble `field.toDouble`]
// and converts with Java's Integer.parseInt / Double.parseDouble, which is CORRECTLY ROUNDED.  This header restates the grammar
// and does the conversion exactly, without a big-number library:
//   * up to 19 significant digits go into a 64-bit integer w, the rest of the literal into a decimal exponent q;
//   * w <= 2^53 and |q| <= 22: one IEEE multiplication or division of two exactly representable doubles (Clinger's fast path);
//   * |q| <= 27: w * 5^q as a 128-bit product, or w / 5^-q by a 128 / 64-bit division with the remainder as sticky bit,
//     rounded to 53 bits half-to-even;
//   * more than 19 digits: both the truncated w and w + 1 are converted; if they agree the result is exact, otherwise — and for
//     |q| > 27 — the field is reported as UNSUPPORTED and the reader fails loudly (no silent approximation).
// Compiles for the host too (B2F_HD empty): tests/test_csv_number_host.py checks it against Python's float() without a GPU.
#pragma once
#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define B2F_HD __host__ __device__ __forceinline__
#else
#define B2F_HD static inline
#endif

namespace b200flow {

enum CsvClass : int { CSV_NULL = 0, CSV_INT = 1, CSV_LONG = 2, CSV_DOUBLE = 3, CSV_STRING = 4 };
enum CsvNum : int { CSVNUM_OK = 0, CSVNUM_NOT_A_NUMBER = 1, CSVNUM_UNSUPPORTED = 2 };

B2F_HD int csv_clz64(uint64_t v) {
#if defined(__CUDA_ARCH__)
    return __clzll((long long)v);
#else
    return __builtin_clzll(v);
#endif
}

B2F_HD bool csv_is_space(uint8_t c) { return c <= 0x20; }   // Java String.trim / univocity: every char <= ' '
B2F_HD bool csv_is_digit(uint8_t c) { return (uint8_t)(c - '0') < 10; }

B2F_HD uint64_t csv_pow5(int k) {   // 5^k, k in [0, 27] (5^27 < 2^63)
    uint64_t r = 1;
    for (int i = 0; i < k; ++i) r *= 5u;
    return r;
}

B2F_HD double csv_pow10_exact(int k) {   // 10^k, k in [0, 22]: exactly representable
    double r = 1.0;
    for (int i = 0; i < k; ++i) r *= 10.0;
    return r;
}

// nearest double (ties to even) of P * 2^e2 (+ a sticky remainder below P's last bit); P != 0; the result is a normal number
B2F_HD double csv_round_u128(unsigned __int128 P, int e2, bool sticky) {
    const uint64_t hi = (uint64_t)(P >> 64), lo = (uint64_t)P;
    const int nb = hi ? 128 - csv_clz64(hi) : 64 - csv_clz64(lo);
    if (nb <= 53) return ldexp((double)lo, e2);            // exact (sticky is never set in this case: see the callers)
    const int shift = nb - 53;
    uint64_t m = (uint64_t)(P >> shift);
    const unsigned __int128 rem = P & ((((unsigned __int128)1) << shift) - 1);
    const unsigned __int128 half = ((unsigned __int128)1) << (shift - 1);
    const bool up = rem > half || (rem == half && (sticky || (m & 1u)));
    m += up ? 1u : 0u;                                      // 2^53 stays exactly representable
    return ldexp((double)m, e2 + shift);
}

// w * 10^q, correctly rounded; w != 0, |q| <= 27
B2F_HD double csv_scale_exact(uint64_t w, int q) {
    if (w <= (1ull << 53) && q >= -22 && q <= 22) {
        const double d = (double)w;
        return q < 0 ? d / csv_pow10_exact(-q) : d * csv_pow10_exact(q);
    }
    if (q >= 0) return csv_round_u128((unsigned __int128)w * csv_pow5(q), q, false);
    const int k = -q;
    const uint64_t D = csv_pow5(k);
    const int a = csv_clz64(w), b = csv_clz64(D);
    const uint64_t wn = w << a, Dn = D << b;
    const unsigned __int128 N = ((unsigned __int128)wn) << 63;          // wn / 2 < 2^63 <= Dn: the quotient fits 64 bits
    const uint64_t Q = (uint64_t)(N / Dn);
    const bool sticky = (N % Dn) != 0;
    return csv_round_u128((unsigned __int128)Q, -63 + b - a - k, sticky);   // Q >= 2^62: never the nb <= 53 branch
}

struct CsvNumber {
    uint64_t w;        // first <= 19 significant digits
    int q;             // value = w * 10^q (+ dropped digits)
    bool neg, truncated, integer_syntax;   // integer_syntax: [+-]digits only, no surrounding blanks
    int n_digits;      // significant digits seen (after leading zeros)
};

B2F_HD bool csv_match(const uint8_t* p, int len, const char* lit) {
    int i = 0;
    for (; lit[i]; ++i) if (i >= len || p[i] != (uint8_t)lit[i]) return false;
    return i == len;
}

// Scans [ws] [sign] digits [. digits] [(e|E) [sign] digits] [ws]  (Java Double.parseDouble without hex floats and type suffixes).
// Returns false if the field is not of that form.
B2F_HD bool csv_scan_number(const uint8_t* p, int len, CsvNumber* out) {
    int i = 0, e = len;
    while (i < e && csv_is_space(p[i])) ++i;
    while (e > i && csv_is_space(p[e - 1])) --e;
    const bool blanks = i != 0 || e != len;
    if (i >= e) return false;
    bool neg = false;
    if (p[i] == '+' || p[i] == '-') { neg = p[i] == '-'; ++i; }
    uint64_t w = 0; int nd = 0, dropped = 0, frac = 0; bool trunc = false, any = false, isint = true;
    while (i < e && csv_is_digit(p[i])) {
        any = true;
        const int d = p[i] - '0';
        if (nd == 0 && d == 0) { ++i; continue; }              // leading zeros
        if (nd < 19) { w = w * 10u + (uint64_t)d; ++nd; } else { ++dropped; trunc |= d != 0; }
        ++i;
    }
    if (i < e && p[i] == '.') {
        isint = false; ++i;
        while (i < e && csv_is_digit(p[i])) {
            any = true;
            const int d = p[i] - '0';
            if (nd == 0 && d == 0) { ++frac; ++i; continue; }   // 0.000ddd: zeros only move the exponent
            if (nd < 19) { w = w * 10u + (uint64_t)d; ++nd; ++frac; } else { trunc |= d != 0; }
            ++i;
        }
    }
    if (!any) return false;
    int ex = 0;
    if (i < e && (p[i] == 'e' || p[i] == 'E')) {
        isint = false; ++i;
        bool eneg = false;
        if (i < e && (p[i] == '+' || p[i] == '-')) { eneg = p[i] == '-'; ++i; }
        if (i >= e || !csv_is_digit(p[i])) return false;
        while (i < e && csv_is_digit(p[i])) { if (ex < 100000) ex = ex * 10 + (p[i] - '0'); ++i; }
        if (eneg) ex = -ex;
    }
    if (i != e) return false;
    out->w = w; out->q = ex - frac + dropped; out->neg = neg; out->truncated = trunc; out->integer_syntax = isint && !blanks;
    out->n_digits = nd + dropped;
    return true;
}

// Type of one field as Spark's inference sees it (empty -> null).
B2F_HD int csv_classify(const uint8_t* p, int len) {
    if (len == 0) return CSV_NULL;
    CsvNumber n;
    if (csv_scan_number(p, len, &n)) {
        if (n.integer_syntax && n.n_digits <= 19 && !n.truncated) {
            // [+-]digits: int32, int64, or beyond (Spark: Decimal, then Double)
            const uint64_t lim32 = n.neg ? 2147483648ull : 2147483647ull, lim64 = n.neg ? 9223372036854775808ull : 9223372036854775807ull;
            if (n.w <= lim32) return CSV_INT;
            if (n.w <= lim64) return CSV_LONG;
        }
        return CSV_DOUBLE;
    }
    int i = 0, e = len;                                       // Double.parseDouble trims before it matches the named values
    while (i < e && csv_is_space(p[i])) ++i;
    while (e > i && csv_is_space(p[e - 1])) --e;
    const uint8_t* s = p + i; const int l = e - i;
    if (csv_match(s, l, "NaN") || csv_match(s, l, "Infinity") || csv_match(s, l, "+Infinity") || csv_match(s, l, "-Infinity") ||
        csv_match(s, l, "Inf") || csv_match(s, l, "-Inf") || csv_match(s, l, "+Inf"))      // CSVOptions nanValue / positiveInf / negativeInf
        return l == 0 ? CSV_NULL : CSV_DOUBLE;
    return CSV_STRING;
}

// Field of a DOUBLE column -> value.  Empty -> NaN (null).  CSVNUM_UNSUPPORTED: a literal this reader cannot round exactly.
B2F_HD int csv_parse_double(const uint8_t* p, int len, double* out) {
    if (len == 0) { *out = nan(""); return CSVNUM_OK; }
    CsvNumber n;
    if (!csv_scan_number(p, len, &n)) {
        int i = 0, e = len;
        while (i < e && csv_is_space(p[i])) ++i;
        while (e > i && csv_is_space(p[e - 1])) --e;
        const uint8_t* s = p + i; const int l = e - i;
        if (csv_match(s, l, "NaN")) { *out = nan(""); return CSVNUM_OK; }
        if (csv_match(s, l, "Infinity") || csv_match(s, l, "+Infinity") || csv_match(s, l, "Inf") || csv_match(s, l, "+Inf")) { *out = INFINITY; return CSVNUM_OK; }
        if (csv_match(s, l, "-Infinity") || csv_match(s, l, "-Inf")) { *out = -INFINITY; return CSVNUM_OK; }
        return CSVNUM_NOT_A_NUMBER;
    }
    if (n.w == 0) { *out = n.neg ? -0.0 : 0.0; return CSVNUM_OK; }
    if (n.q < -27 || n.q > 27) return CSVNUM_UNSUPPORTED;
    double v = csv_scale_exact(n.w, n.q);
    if (n.truncated) {                                        // the true value lies in (w, w + 1) * 10^q
        if (n.w == 0xFFFFFFFFFFFFFFFFull || csv_scale_exact(n.w + 1, n.q) != v) return CSVNUM_UNSUPPORTED;
    }
    *out = n.neg ? -v : v;
    return CSVNUM_OK;
}

// Field of an INT column -> value (the column's type was inferred from these very fields, so anything else is an error).
B2F_HD int csv_parse_int32(const uint8_t* p, int len, int32_t* out) {
    CsvNumber n;
    if (len == 0 || !csv_scan_number(p, len, &n) || !n.integer_syntax || n.truncated || n.n_digits > 19) return CSVNUM_NOT_A_NUMBER;
    if (n.w > (n.neg ? 2147483648ull : 2147483647ull)) return CSVNUM_NOT_A_NUMBER;
    *out = n.neg ? (int32_t)(0u - (uint32_t)n.w) : (int32_t)n.w;
    return CSVNUM_OK;
}

B2F_HD uint64_t csv_hash(const uint8_t* p, int len) {   // FNV-1a, never 0 (0 marks an empty slot)
    uint64_t h = 1469598103934665603ull;
    for (int i = 0; i < len; ++i) { h ^= p[i]; h *= 1099511628211ull; }
    return h ? h : 1ull;
}

}  // namespace b200flow
