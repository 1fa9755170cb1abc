// One BLAST-6 record -> (query id, hit id, bit score).  Host/device shared so that the grammar can be
// unit-tested on the CPU against Python's int()/float() (tests/test_parse_line_host.py).
//
// Mirrors blast_file_processor.py:66-82: fields split on '\t'; ids are
// int(field.split("_", 2)[1]) of fields 0 and 1 (checked first, :72-76), the score is float(field 11)
// (:78-82).  int()/float() tolerate surrounding ASCII whitespace (the BLAST fixtures right-align the
// score: " 694").  Decimal -> fp64 is exact (correctly rounded like float()) for <= 19 significant
// digits with the mantissa <= 2^53 and |decimal exponent| <= 22 (one IEEE division or
// multiplication by an exactly representable power of ten); anything Python would accept beyond that
// (inf, nan, digit underscores, longer mantissas, huge exponents) is reported as
// OFG_LINE_SCORE_FORMAT instead of being approximated.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define OFG_HD_INLINE __host__ __device__ __forceinline__
#else
#define OFG_HD_INLINE static inline
#endif

#define OFG_PL_OK 0
#define OFG_PL_BAD_ID 1
#define OFG_PL_BAD_SCORE 2
#define OFG_PL_SCORE_FORMAT 6

OFG_HD_INLINE bool ofg_is_term(int c) { return c == '\n' || c == '\r'; }
// Python str.strip()/int()/float() whitespace that can occur inside a tab-separated field
OFG_HD_INLINE bool ofg_is_space(int c) { return c == ' ' || c == '\v' || c == '\f' || c == 0x1c || c == 0x1d || c == 0x1e || c == 0x1f; }

// 10^k for 0 <= k <= 22, all exactly representable
#if defined(__CUDACC__)
__device__ __constant__ double OFG_POW10_DEV[23] = {1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11,
                                                    1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
#endif
#if defined(__CUDACC__)
// correctly rounded 1 / 10^k, 0 <= k <= 8
__device__ __constant__ double OFG_RCP10_DEV[9] = {1.0, 1.0 / 1e1, 1.0 / 1e2, 1.0 / 1e3, 1.0 / 1e4, 1.0 / 1e5, 1.0 / 1e6, 1.0 / 1e7, 1.0 / 1e8};
#endif
OFG_HD_INLINE double ofg_pow10_exact(int k)
{
#if defined(__CUDA_ARCH__)
    return OFG_POW10_DEV[k];
#else
    static const double t[23] = {1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11,
                                 1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
    return t[k];
#endif
}

// Parses the id field starting at pos (ends at '\t' or a line terminator).  On success *id holds the
// integer and pos is left ON the delimiter that ended the field.
template <class Src>
OFG_HD_INLINE int ofg_parse_id_field(Src& src, int64_t& pos, uint32_t* id)
{
    int c = src.get(pos);
    // text before the first '_' (the species id) is ignored (:72)
    while (c != '_') {
        if (c == '\t' || ofg_is_term(c)) return OFG_PL_BAD_ID;  // IndexError: no "_" in the field
        c = src.get(++pos);
    }
    c = src.get(++pos);
    while (ofg_is_space(c)) c = src.get(++pos);
    bool neg = false;
    if (c == '+' || c == '-') { neg = (c == '-'); c = src.get(++pos); }
    if (c < '0' || c > '9') return OFG_PL_BAD_ID;
    uint64_t v = 0;
    while (c >= '0' && c <= '9') {
        v = v * 10 + (uint64_t)(c - '0');
        if (v > 0x7fffffffULL) v = 0x7fffffffULL;  // saturate: reported later as out of range
        c = src.get(++pos);
    }
    while (ofg_is_space(c)) c = src.get(++pos);
    if (c == '_') {  // split("_", 2): the rest of the field is ignored
        while (c != '\t' && !ofg_is_term(c)) c = src.get(++pos);
    } else if (c != '\t' && !ofg_is_term(c)) {
        return OFG_PL_BAD_ID;
    }
    if (neg && v != 0) return OFG_PL_BAD_ID;  // negative ids index from the end in the reference; rejected here
    *id = (uint32_t)v;
    return OFG_PL_OK;
}

// Parses the score field starting at pos; the field ends at '\t' or a terminator.
template <class Src>
OFG_HD_INLINE int ofg_parse_score_field(Src& src, int64_t pos, double* out)
{
    int c = src.get(pos);
    while (ofg_is_space(c)) c = src.get(++pos);
    bool neg = false;
    if (c == '+' || c == '-') { neg = (c == '-'); c = src.get(++pos); }
    uint64_t mant = 0;   // significant digits read so far, trailing zeros held back in pz
    int pz = 0, pzf = 0; // pending zeros after the last non-zero digit (total / of which fractional)
    int frac_in = 0;     // fractional digits already folded into mant
    bool any = false, inexact = false, odd = false, after_point = false;
    for (;;) {
        if (c == '.' && !after_point) { after_point = true; c = src.get(++pos); continue; }
        if (c < '0' || c > '9') break;
        any = true;
        if (c == '0') {
            if (mant != 0) { pz++; if (after_point) pzf++; }
            else if (after_point) frac_in++;  // leading fractional zeros only move the exponent
        } else {
            for (int t = 0; t <= pz; t++) {
                if (mant >= 1000000000000000000ULL) { inexact = true; break; }
                mant *= 10;
            }
            mant += (uint64_t)(c - '0');
            frac_in += pzf + (after_point ? 1 : 0);
            pz = pzf = 0;
        }
        c = src.get(++pos);
    }
    int dec_exp = (pz - pzf) - frac_in;
    if (!any) {
        // "inf", "nan", "infinity" (any case) are valid for float(): unsupported here, everything else is a ValueError
        int l = c | 0x20;
        if (l == 'i' || l == 'n') odd = true;
        return odd ? OFG_PL_SCORE_FORMAT : OFG_PL_BAD_SCORE;
    }
    if (c == '_') return OFG_PL_SCORE_FORMAT;  // "1_000" is valid for float()
    if (c == 'e' || c == 'E') {
        c = src.get(++pos);
        bool eneg = false;
        if (c == '+' || c == '-') { eneg = (c == '-'); c = src.get(++pos); }
        if (c < '0' || c > '9') return OFG_PL_BAD_SCORE;
        int e = 0;
        while (c >= '0' && c <= '9') {
            if (e < 100000) e = e * 10 + (c - '0');
            c = src.get(++pos);
        }
        if (c == '_') return OFG_PL_SCORE_FORMAT;
        dec_exp += eneg ? -e : e;
    }
    while (ofg_is_space(c)) c = src.get(++pos);
    if (c != '\t' && !ofg_is_term(c)) return OFG_PL_BAD_SCORE;
    double v;
    if (mant == 0) v = 0.0;
    else if (inexact || mant > 9007199254740992ULL) return OFG_PL_SCORE_FORMAT;
    else if (dec_exp == 0) v = (double)mant;
    else if (dec_exp < 0 && dec_exp >= -22) v = (double)mant / ofg_pow10_exact(-dec_exp);
    else if (dec_exp > 0 && dec_exp <= 22) v = (double)mant * ofg_pow10_exact(dec_exp);
    else return OFG_PL_SCORE_FORMAT;
    *out = neg ? -v : v;
    return OFG_PL_OK;
}

// Full record starting at pos (first byte of a non-empty line).
template <class Src>
OFG_HD_INLINE int ofg_parse_record(Src& src, int64_t pos, uint32_t* f0, uint32_t* f1, double* score)
{
    int rc = ofg_parse_id_field(src, pos, f0);
    if (rc) return rc;
    if (src.get(pos) != '\t') return OFG_PL_BAD_ID;  // row[1] does not exist
    ++pos;
    rc = ofg_parse_id_field(src, pos, f1);
    if (rc) return rc;
    // skip to the start of field 11: ten more tabs (we are on the delimiter after field 1)
    int c = src.get(pos);
    for (int k = 0; k < 10; k++) {
        if (c != '\t') return OFG_PL_BAD_SCORE;  // IndexError on row[11]
        c = src.get(++pos);
        if (k < 9) while (c != '\t' && !ofg_is_term(c)) c = src.get(++pos);
    }
    return ofg_parse_score_field(src, pos, score);
}
