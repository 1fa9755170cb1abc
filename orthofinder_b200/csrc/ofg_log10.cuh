// Device log10 that reproduces, bit for bit, what numpy.log10 returns for float64 arrays on an
// AVX512_SKX host (Intel SVML __svml_log108_ha, linked into numpy's _multiarray_umath), which is what
// the reference feeds to curve_fit at gathering.py:82,120.  A correctly rounded log10 would differ
// from it on about 1 input in 10^4 and that is enough to flip the Levenberg-Marquardt accept test
// (SURVEY.md finding 5), so the operation sequence is reproduced: r = RNDSCALE5(VRCP14PD(m)) as a
// 16-threshold step function of the top 16 mantissa bits (measured exhaustively), R = fma(r, m, -1),
// hi/lo table of -log10(r), degree-9 polynomial with SVML's FMA tree, double-double style final sum.
// Domain: positive finite normal doubles.  The oracle's independent C restatement is
// oracle/csrc/svml_log10.h; both are pinned against np.log10 in tests.
#pragma once
#include <stdint.h>

__device__ __constant__ unsigned long long OFG_SVML_TH[16] = {
    0x0000000000000000ULL, 0xbf9af5f92b000000ULL, 0xbfaa30a9d60a0000ULL, 0xbfb31b3055c40000ULL,
    0xbfb8cf1838860000ULL, 0xbfbe3bc1ab0e0000ULL, 0xbfc1b3e71ec98000ULL, 0xbfc42c7e7fe40000ULL,
    0x3fbffbfc2bbc0000ULL, 0x3fbb721cd1710000ULL, 0x3fb715d0ce360000ULL, 0x3fb2e3a740b80000ULL,
    0x3fadb11ed7660000ULL, 0x3fa5e3966b7e0000ULL, 0x3f9cb38fccd80000ULL, 0x3f8c3d0837780000ULL};
__device__ __constant__ unsigned long long OFG_SVML_TL[16] = {
    0x0000000000000000ULL, 0xbd4cc1f4dbc14db5ULL, 0x3d20163d7e67d282ULL, 0xbd5c46006d082e6dULL,
    0xbd520326ca2c7e5fULL, 0xbd39fe3d562a53f0ULL, 0x3d5842a222666db0ULL, 0x3d1ff1d22abd8abeULL,
    0x3d5e00dd60df12c3ULL, 0x3d55f8a57d982e77ULL, 0x3d5ebeeb67bf155aULL, 0xbd5ffc5aef53f2edULL,
    0x3d557e865b878377ULL, 0x3d5252a37739a5f7ULL, 0x3d47fb6d2d6528c7ULL, 0x3d2310272fe17537ULL};
__device__ __constant__ unsigned short OFG_RCP_THR[16] = {0x040f, 0x0c97, 0x15b4, 0x1f70, 0x29e6, 0x3523, 0x4143, 0x4e5f,
                                                          0x5c99, 0x6c15, 0x7d07, 0x8f9c, 0xa41a, 0xbad0, 0xd41c, 0xf080};

__device__ __forceinline__ double ofg_b2d(unsigned long long b) { return __longlong_as_double((long long)b); }

__device__ __forceinline__ double ofg_log10_svml(double x)
{
    const unsigned long long xb = (unsigned long long)__double_as_longlong(x);
    const int eb = (int)((xb >> 52) & 0x7ff);
    if ((xb >> 63) || eb == 0 || eb == 0x7ff) return log10(x);  // outside the emulated domain
    double e = (double)(eb - 1023);
    const double m = ofg_b2d((xb & 0x000fffffffffffffULL) | 0x3ff0000000000000ULL);
    const unsigned top16 = (unsigned)((xb >> 36) & 0xffffu);
    int steps = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) steps += (top16 >= (unsigned)OFG_RCP_THR[i]);
    const int r32 = 32 - steps;
    const double r = (double)r32 * 0.03125;
    const int idx = r32 & 15;
    const double R = __fma_rn(r, m, -1.0);
    if (r < 0.75) e = __dadd_rn(e, 1.0);
    const double pa = __fma_rn(ofg_b2d(0x3fafc3f6f38b58a2ULL), R, ofg_b2d(0xbfb287a63464dc80ULL));
    const double pb = __fma_rn(ofg_b2d(0x3fa8c2d828480370ULL), R, ofg_b2d(0xbfabd80d96029814ULL));
    const double R2 = __dmul_rn(R, R);
    const double R4 = __dmul_rn(R2, R2);
    const double p1 = __fma_rn(R2, pb, pa);
    const double Th = ofg_b2d(OFG_SVML_TH[idx]);
    const double Tl = ofg_b2d(OFG_SVML_TL[idx]);
    const double Kh = __fma_rn(ofg_b2d(0x3fd34413509f0000ULL), e, Th);
    double Kl = __fma_rn(ofg_b2d(0x3d7e7fbcc47c4acdULL), e, Tl);
    const double pc = __fma_rn(ofg_b2d(0x3fc287a7636f428cULL), R, ofg_b2d(0xbfcbcb7b1526e4dbULL));
    const double pd = __fma_rn(ofg_b2d(0x3fb63c62777f27d9ULL), R, ofg_b2d(0xbfbbcb7b153c06a3ULL));
    const double C3c0 = ofg_b2d(0x3fdbcb7b1526e50eULL);
    const double Rh = __fma_rn(R, C3c0, Kh);
    const double p2 = __fma_rn(R2, pd, pc);
    const double d = __dsub_rn(Rh, Kh);
    double p = __fma_rn(R4, p1, p2);
    p = __fma_rn(R, p, ofg_b2d(0x3c69a5fdaaf11825ULL));
    const double Rl = __fma_rn(R, C3c0, -d);
    Kl = __dadd_rn(Kl, Rl);
    p = __fma_rn(R, p, Kl);
    return __dadd_rn(Rh, p);
}
