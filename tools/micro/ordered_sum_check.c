// Exactness check of an integer emulation of a sequentially rounded fp64 sum (an experiment for fit_kernel):
// while every partial sum stays inside the binade of s, RN(s_i + a_i) = s_i + g * rint(a_i / g) with g = ulp(s), so the additions are
// integer additions that a warp could do in any order; ties (a_i / g = k + 1/2) need the parity of the running sum and are rejected here.
// Compares the emulation with plain sequential sums chunk by chunk and reports how often the fast path applies.
//   gcc -O2 -ffp-contract=off -o ordered_sum_check tools/micro/ordered_sum_check.c -lm && ./ordered_sum_check
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include <string.h>
#include <math.h>
#define CH 512
static uint64_t rs = 88172645463325252ull;
static double urand(void){ rs ^= rs<<13; rs ^= rs>>7; rs ^= rs<<17; return (rs>>11) * (1.0/9007199254740992.0); }
static double nrand(void){ double u=urand()+1e-300, v=urand(); return sqrt(-2*log(u))*cos(6.283185307179586*v); }
// returns 1 if the fast path applied (and *out set), 0 otherwise
static int fast_chunk(double s, const double* q, double* out)
{
    uint64_t sb; memcpy(&sb, &s, 8);
    int e = (int)((sb >> 52) & 0x7ff) - 1023;
    if (!(e > -900 && e < 900)) return 0;
    uint64_t scb = ((uint64_t)(1023 + 52 - e) << 52) | (sb & 0x8000000000000000ull);
    uint64_t gb = ((uint64_t)(1023 + e - 52) << 52) | (sb & 0x8000000000000000ull);
    double sc, g; memcpy(&sc, &scb, 8); memcpy(&g, &gb, 8);
    long long S = (long long)((sb & 0xfffffffffffffull) | 0x10000000000000ull);
    const long long LO = (1ll << 52) + 1, HI = (1ll << 53) - 1;
    long long P = S;
    for (int i = 0; i < CH; i++) {
        double x = q[i] * sc;
        if (!(fabs(x) < 9007199254740992.0)) return 0;
        double kf = rint(x);
        if (fabs(x - kf) == 0.5) return 0;
        P += (long long)kf;
        if (P < LO || P > HI) return 0;
    }
    *out = (double)P * g;
    return 1;
}
int main(void)
{
    long chunks = 0, fast = 0, bad = 0;
    for (int trial = 0; trial < 3000; trial++) {
        int kind = trial % 6;
        int m = 512 * (20 + trial % 90);
        double s = 0.0;
        double a0 = 0.5 + urand(), b0 = urand() * 3 - 1.5;
        for (int c = 0; c < m / CH; c++) {
            double q[CH];
            for (int i = 0; i < CH; i++) {
                double lx = 3.0 + 4.0 * urand(), r = 0.3 * nrand();
                double f = a0 * lx + b0 + r - (a0 * lx + b0);
                switch (kind) {
                case 0: q[i] = f * f; break;                       // enorm of residuals
                case 1: q[i] = lx * lx; break;                     // column norm
                case 2: q[i] = lx * f; break;                      // J^T f (mixed signs)
                case 3: q[i] = lx; break;                          // dot with ones
                case 4: q[i] = f; break;                           // residual sum (mixed)
                default: q[i] = (i % 7 == 0) ? ldexp(floor(urand() * 1024), -3) : lx * lx * 1e-3; break;  // many exact ties
                }
            }
            double seq = s;
            for (int i = 0; i < CH; i++) seq = seq + q[i];
            double fo;
            chunks++;
            if (fast_chunk(s, q, &fo)) { fast++; if (memcmp(&fo, &seq, 8)) { bad++; if (bad < 10) printf("MISMATCH kind %d chunk %d: %a vs %a\n", kind, c, fo, seq); } }
            s = seq;
        }
    }
    printf("chunks %ld fast %ld (%.1f%%) mismatches %ld\n", chunks, fast, 100.0 * fast / chunks, bad);
    return bad != 0;
}
