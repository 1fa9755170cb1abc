// Build: gcc -O2 -fopenmp -mfma -ffp-contract=off div_fast_check.c -lm   (fit_div of csrc/ofg_fit_core.h)
// statistical + structured check: q0 = x*r; q1 = fma(fma(-q0, d, x), r, q0) with r = RN(1/d) == x / d ?
#include <stdio.h>
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <omp.h>
static inline uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
static inline uint64_t next(uint64_t* s) { uint64_t s0 = s[0], s1 = s[1], r = s0 + s1; s1 ^= s0; s[0] = rotl(s0, 55) ^ s1 ^ (s1 << 14); s[1] = rotl(s1, 36); return r; }
static inline double mk(uint64_t bits, int emin, int emax) {  // random significand, exponent in [emin, emax]
    uint64_t m = bits & 0xFFFFFFFFFFFFFULL; int e = emin + (int)((bits >> 52) % (uint64_t)(emax - emin + 1));
    uint64_t u = ((uint64_t)(e + 1023) << 52) | m; if (bits >> 63) u |= 1ULL << 63; double d; memcpy(&d, &u, 8); return d; }
int main(void) {
    long bad = 0, total = 0;
    #pragma omp parallel reduction(+:bad,total)
    {
        uint64_t s[2] = {0x9E3779B97F4A7C15ULL * (omp_get_thread_num() + 1), 0xD1B54A32D192ED03ULL + omp_get_thread_num()};
        for (long it = 0; it < 40000000L; it++) {
            const double d = mk(next(s), -40, 40);
            const double r = 1.0 / d;
            for (int k = 0; k < 16; k++) {
                uint64_t b = next(s);
                if ((k & 3) == 3) b &= ~0xFFFFFFFFULL >> (b & 31);  // few significant bits: exact / near-tie quotients
                const double x = mk(b, -60, 60);
                const double q0 = x * r, q1 = fma(fma(-q0, d, x), r, q0);
                if (q1 != x / d) { bad++; if (bad < 5) printf("BAD x=%a d=%a got %a want %a\n", x, d, q1, x / d); }
                total++;
            }
        }
    }
    printf("checked %ld pairs, bad %ld\n", total, bad);
    // all-ones significand divisor (Markstein's exceptional case)
    long bad1 = 0; uint64_t s[2] = {1, 2};
    for (long it = 0; it < 100000000L; it++) { uint64_t u = 0x3FFFFFFFFFFFFFFFULL; double d; memcpy(&d, &u, 8); const double r = 1.0 / d; const double x = mk(next(s), -10, 10);
        const double q0 = x * r, q1 = fma(fma(-q0, d, x), r, q0); if (q1 != x / d) bad1++; }
    printf("all-ones divisor: bad %ld of 1e8\n", bad1);
    return 0;
}
