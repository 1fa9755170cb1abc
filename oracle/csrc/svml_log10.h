/* Bit-exact emulation of the log10 that numpy applies to float64 arrays on AVX512_SKX hosts.
 *
 * TEST INFRASTRUCTURE (oracle).  The reference calls np.log10 at gathering.py:82 and :120
 * (scnorm.loglinear / CalculateFittingParameters).  numpy >= 1.22 dispatches float64 log10 on
 * AVX512_SKX machines to Intel SVML's __svml_log108_ha (numpy/_core/src/umath/svml, linked into
 * _multiarray_umath).  Its result is within 0.5x ULP but NOT always correctly rounded (about 1
 * input in 10^4 differs from the correctly rounded value), and the Levenberg-Marquardt fit that
 * consumes it is sensitive to single ULPs (SURVEY.md finding 5), so the operation sequence is
 * restated here: mantissa/exponent split, r = RNDSCALE5(VRCP14PD(m)), R = r*m - 1 (fma),
 * 16-entry hi/lo table of -log10(r), degree-9 polynomial evaluated with the same FMA tree and
 * the same final double-double style summation.
 *
 * VRCP14PD followed by rounding to 5 fractional bits was measured exhaustively on the build
 * host to depend only on the top 16 mantissa bits and to be a monotone step function with the
 * 16 thresholds in RCP_THR (tools/probe_rcp14.c regenerates them).
 * Constants are the published SVML table __svml_dlog10_ha_data_internal_avx512.
 * Domain: positive, finite, normal doubles (all the path ever feeds it).  Pinned against
 * np.log10 by tests/test_log10_emul.py.
 */
#ifndef OFG_ORACLE_SVML_LOG10_H
#define OFG_ORACLE_SVML_LOG10_H
#include <stdint.h>
#include <string.h>
#include <math.h>

static const uint64_t SVML_TH_BITS[16] = {
    0x0000000000000000ULL, 0xbf9af5f92b000000ULL, 0xbfaa30a9d60a0000ULL, 0xbfb31b3055c40000ULL,
    0xbfb8cf1838860000ULL, 0xbfbe3bc1ab0e0000ULL, 0xbfc1b3e71ec98000ULL, 0xbfc42c7e7fe40000ULL,
    0x3fbffbfc2bbc0000ULL, 0x3fbb721cd1710000ULL, 0x3fb715d0ce360000ULL, 0x3fb2e3a740b80000ULL,
    0x3fadb11ed7660000ULL, 0x3fa5e3966b7e0000ULL, 0x3f9cb38fccd80000ULL, 0x3f8c3d0837780000ULL};
static const uint64_t SVML_TL_BITS[16] = {
    0x0000000000000000ULL, 0xbd4cc1f4dbc14db5ULL, 0x3d20163d7e67d282ULL, 0xbd5c46006d082e6dULL,
    0xbd520326ca2c7e5fULL, 0xbd39fe3d562a53f0ULL, 0x3d5842a222666db0ULL, 0x3d1ff1d22abd8abeULL,
    0x3d5e00dd60df12c3ULL, 0x3d55f8a57d982e77ULL, 0x3d5ebeeb67bf155aULL, 0xbd5ffc5aef53f2edULL,
    0x3d557e865b878377ULL, 0x3d5252a37739a5f7ULL, 0x3d47fb6d2d6528c7ULL, 0x3d2310272fe17537ULL};
/* top-16-mantissa-bit thresholds at which RNDSCALE5(RCP14(m)) steps down by 1/32 */
static const uint16_t RCP_THR[16] = {0x040f, 0x0c97, 0x15b4, 0x1f70, 0x29e6, 0x3523, 0x4143, 0x4e5f,
                                     0x5c99, 0x6c15, 0x7d07, 0x8f9c, 0xa41a, 0xbad0, 0xd41c, 0xf080};

static inline double svml_bits2d(uint64_t b) { double d; memcpy(&d, &b, 8); return d; }

static inline double svml_log10_emul(double x)
{
    uint64_t xb; memcpy(&xb, &x, 8);
    int ebits = (int)((xb >> 52) & 0x7ff);
    if ((xb >> 63) || ebits == 0 || ebits == 0x7ff) return log10(x); /* outside the emulated domain */
    const double C180 = svml_bits2d(0x3fa8c2d828480370ULL), C1c0 = svml_bits2d(0xbfabd80d96029814ULL);
    const double C200 = svml_bits2d(0x3fafc3f6f38b58a2ULL), C240 = svml_bits2d(0xbfb287a63464dc80ULL);
    const double C280 = svml_bits2d(0x3fb63c62777f27d9ULL), C2c0 = svml_bits2d(0xbfbbcb7b153c06a3ULL);
    const double C300 = svml_bits2d(0x3fc287a7636f428cULL), C340 = svml_bits2d(0xbfcbcb7b1526e4dbULL);
    const double C380 = svml_bits2d(0x3c69a5fdaaf11825ULL), C3c0 = svml_bits2d(0x3fdbcb7b1526e50eULL);
    const double L2H = svml_bits2d(0x3fd34413509f0000ULL), L2L = svml_bits2d(0x3d7e7fbcc47c4acdULL);
    double e = (double)(ebits - 1023);                                   /* VGETEXPPD */
    double m = svml_bits2d((xb & 0x000fffffffffffffULL) | 0x3ff0000000000000ULL); /* VGETMANTPD [1,2) */
    unsigned top16 = (unsigned)((xb >> 36) & 0xffff);
    int steps = 0;
    for (int i = 0; i < 16; i++) steps += (top16 >= RCP_THR[i]);
    int r32 = 32 - steps;                                                /* r = r32/32 in {16..32}/32 */
    double r = (double)r32 / 32.0;
    int idx = r32 & 15;                                                  /* top 4 mantissa bits of r */
    double R = fma(r, m, -1.0);                                          /* vfmsub213pd */
    if (r < 0.75) e = e + 1.0;
    double pa = fma(C200, R, C240);
    double pb = fma(C180, R, C1c0);
    double R2 = R * R;
    double R4 = R2 * R2;
    double p1 = fma(R2, pb, pa);
    double Th = svml_bits2d(SVML_TH_BITS[idx]);
    double Tl = svml_bits2d(SVML_TL_BITS[idx]);
    double Kh = fma(L2H, e, Th);
    double Kl = fma(L2L, e, Tl);
    double pc = fma(C300, R, C340);
    double pd = fma(C280, R, C2c0);
    double Rh = fma(R, C3c0, Kh);
    double p2 = fma(R2, pd, pc);
    double d = Rh - Kh;
    double p = fma(R4, p1, p2);
    p = fma(R, p, C380);
    double Rl = fma(R, C3c0, -d);
    Kl = Kl + Rl;
    p = fma(R, p, Kl);
    return Rh + p;
}
#endif
