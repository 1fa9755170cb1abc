// Levenberg-Marquardt fit of log10(score) = a*log10(Lf) + b exactly as scipy.optimize.curve_fit runs it
// for scnorm.CalculateFittingParameters (gathering.py:119-121): MINPACK lmdif from (1,1), forward-
// difference Jacobian, ftol = xtol = 1.49012e-8, gtol = 0, factor = 100, maxfev = 600.
//
// The result depends on the ORDER of every O(m) floating-point sum (SURVEY.md finding 5, Appendix A),
// so those sums run as strictly sequential chains (one thread, index order, no FMA); the O(m)
// element-wise work (residuals, finite-difference columns, Householder updates) is spread over the
// CTA, and the O(1) bookkeeping (qrfac pivoting, lmpar, qrsolv, trust region) is executed redundantly
// by every thread on identical broadcast inputs.
//
// One source, two builds: nvcc compiles it as the body of fit_kernel (one CTA per species pair);
// g++ compiles it with a one-thread "CTA" so the exact same control flow can be checked on the CPU
// against scipy and the oracle (tests/test_host_shared_code.py).
#pragma once
#include <math.h>
#include <float.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define FIT_FN __device__
#define FIT_MFN __device__
#define FIT_TID ((int)threadIdx.x)
#define FIT_NT ((int)blockDim.x)
#define FIT_SYNC() __syncthreads()
#define FADD(a, b) __dadd_rn((a), (b))
#define FSUB(a, b) __dsub_rn((a), (b))
#define FMUL(a, b) __dmul_rn((a), (b))
#define FDIV(a, b) __ddiv_rn((a), (b))
#define FSQRT(a) __dsqrt_rn((a))
#define FFMA(a, b, c) __fma_rn((a), (b), (c))
#else
#define FIT_FN static
#define FIT_MFN
#define FIT_TID 0
#define FIT_NT 1
#define FIT_SYNC() do {} while (0)
#define FADD(a, b) ((a) + (b))
#define FSUB(a, b) ((a) - (b))
#define FMUL(a, b) ((a) * (b))
#define FDIV(a, b) ((a) / (b))
#define FSQRT(a) sqrt((a))
#define FFMA(a, b, c) fma((a), (b), (c))
#endif

#define FIT_CH 512   // doubles staged per chunk and buffer for the sequential chains
#define FIT_MAXV 3   // chains that can run side by side (one per warp)

#define FIT_RING 3    // staging ring depth: chunk c+2 is being fetched while chunk c is accumulated

// Exact division by a divisor that is reused for a whole chain.  With r = RN(1/d), q0 = RN(x r) is within one ulp
// of x / d and the correction q0 + (x - q0 d) r, the residual taken exactly by an fma, rounds to RN(x / d)
// (Markstein's theorem; 5e9 random pairs and the all-ones significand checked in tools/micro/div_fast_check.c).
// Three fp64 instructions instead of the ~45 of the division routine - the staging warps are bound by the fp64
// pipe.  Only used when divisor, dividend and quotient are comfortably normal; anything else divides for real.
FIT_FN double fit_rcp(double d)
{
    const double ad = fabs(d);
    return (ad > 1e-150 && ad < 1e150) ? FDIV(1.0, d) : 0.0;  // 0: no fast path for this divisor
}
FIT_FN double fit_div(double x, double d, double r)
{
    const double ax = fabs(x);
    if (r != 0.0 && ax > 1e-150 && ax < 1e150) {
        const double q0 = FMUL(x, r);
        return FFMA(FFMA(-q0, d, x), r, q0);
    }
    return FDIV(x, d);
}

struct FitState {
    double x0, x1;               // current iterate
    double xp00, xp01, h0;       // x + h0*e0 and h0
    double xp10, xp11, h1;       // x + h1*e1 and h1
    double aj0, t, aj1, temp0;   // signed column norms, reflection coefficients
    double xn0, xn1;             // trial point
    double xnp0, hn0, xnp1, hn1; // its forward-difference points (xn0 + hn0, xn1 + hn1) and steps
    double rh0, rh1, raj0, raj1, rhn0, rhn1;  // fit_rcp of the six divisors (0 = divide for real)
    int piv, hh1, hh2;           // columns swapped; first / second reflection applied
};

struct FitShared {  // lives in shared memory on the device (37 KB)
    double buf[FIT_MAXV][FIT_RING][FIT_CH];  // staged squares (enorm) / products (dot)
    double out[FIT_MAXV];
    int tag[FIT_MAXV][FIT_RING];      // chunk tag set when a staged chunk holds a value outside MINPACK's mid range
    FitState st;                      // the scalars the staging warps read (published before every chain)
    double rx[FIT_RING][FIT_CH];      // raw (lx, ly) of the chunks being staged, filled one chunk ahead by cp.async
    double ry[FIT_RING][FIT_CH];
};

// ordered sum of the staged terms q[0..len): strictly left to right.  On the device the next eight terms
// are fetched (four 16-byte shared-memory loads) while the current eight are being added, so that the loop
// runs at the latency of the dependent fp64 additions and nothing else.
FIT_FN double fit_ordered_sum_var(double s, const double* q, int len);

// Full chunks take the constant-trip-count instance (the compiler keeps the loads one batch ahead of the
// additions there: 8.1 cycles per term measured on B200, the latency of a dependent DADD).
FIT_FN double fit_ordered_sum(double s, const double* q, int len)
{
    if (len == FIT_CH) return fit_ordered_sum_var(s, q, FIT_CH);
    return fit_ordered_sum_var(s, q, len);
}

FIT_FN double fit_ordered_sum_var(double s, const double* q, int len)
{
    int i = 0;
#if defined(__CUDACC__)
    if (len >= 8) {
        const double2* q2 = reinterpret_cast<const double2*>(q);
        double2 a = q2[0], b = q2[1], c = q2[2], d = q2[3];
        for (; i + 16 <= len; i += 8) {
            const double2 na = q2[(i >> 1) + 4], nb = q2[(i >> 1) + 5], nc = q2[(i >> 1) + 6], nd = q2[(i >> 1) + 7];
            s = FADD(s, a.x); s = FADD(s, a.y); s = FADD(s, b.x); s = FADD(s, b.y);
            s = FADD(s, c.x); s = FADD(s, c.y); s = FADD(s, d.x); s = FADD(s, d.y);
            a = na; b = nb; c = nc; d = nd;
        }
        s = FADD(s, a.x); s = FADD(s, a.y); s = FADD(s, b.x); s = FADD(s, b.y);
        s = FADD(s, c.x); s = FADD(s, c.y); s = FADD(s, d.x); s = FADD(s, d.y);
        i += 8;
    }
#endif
    for (; i < len; i++) s = FADD(s, q[i]);
    return s;
}

struct FitEnormAcc {
    double s1, s2, s3, x1max, x3max, agiant;
    FIT_MFN void init(double n) { s1 = s2 = s3 = x1max = x3max = 0.0; agiant = FDIV(1.304e19, n); }
    FIT_MFN void add(double v)
    {
        const double rdwarf = 3.834e-20;
        const double xabs = fabs(v);
        if (xabs > rdwarf && xabs < agiant) { s2 = FADD(s2, FMUL(xabs, xabs)); return; }
        if (xabs <= rdwarf) {
            if (xabs > x3max) { double t = FDIV(x3max, xabs); s3 = FADD(1.0, FMUL(s3, FMUL(t, t))); x3max = xabs; }
            else if (xabs != 0.0) { double t = FDIV(xabs, x3max); s3 = FADD(s3, FMUL(t, t)); }
        } else {
            if (xabs > x1max) { double t = FDIV(x1max, xabs); s1 = FADD(1.0, FMUL(s1, FMUL(t, t))); x1max = xabs; }
            else { double t = FDIV(xabs, x1max); s1 = FADD(s1, FMUL(t, t)); }
        }
    }
    // same sequence of operations as add() applied to v[0..len), with the squares of a batch of 8
    // computed ahead of the (strictly ordered) additions when all 8 are in MINPACK's mid range
    FIT_MFN void add_run(const double* v, int len)
    {
        const double rdwarf = 3.834e-20;
        int i = 0;
        for (; i + 8 <= len; i += 8) {
            double x[8];
            bool mid = true;
#pragma unroll
            for (int k = 0; k < 8; k++) { x[k] = fabs(v[i + k]); mid = mid && (x[k] > rdwarf && x[k] < agiant); }
            if (mid) {
                double q[8];
#pragma unroll
                for (int k = 0; k < 8; k++) q[k] = FMUL(x[k], x[k]);
#pragma unroll
                for (int k = 0; k < 8; k++) s2 = FADD(s2, q[k]);
            } else {
                for (int k = 0; k < 8; k++) add(v[i + k]);
            }
        }
        for (; i < len; i++) add(v[i]);
    }
    FIT_MFN double result() const
    {
        if (s1 != 0.0) return FMUL(x1max, FSQRT(FADD(s1, FDIV(FDIV(s2, x1max), x1max))));
        if (s2 != 0.0) {
            if (s2 >= x3max) return FSQRT(FMUL(s2, FADD(1.0, FMUL(FDIV(x3max, s2), FMUL(x3max, s3)))));
            return FSQRT(FMUL(x3max, FADD(FDIV(s2, x3max), FMUL(x3max, s3))));
        }
        return FMUL(x3max, FSQRT(s3));
    }
};

FIT_FN double fit_enorm_small(int n, const double* x)
{
    FitEnormAcc acc;
    acc.init((double)n);
    for (int i = 0; i < n; i++) acc.add(x[i]);
    return acc.result();
}

// ---- vectors as pure functions -------------------------------------------------------------------------
// Every long vector of the algorithm (residual, the two Jacobian columns, the Householder vectors, Q^T f) is
// an element-wise function of (log10 Lf_i, log10 S_i) and a handful of scalars.  They are therefore never
// stored: the warps that stage a chain recompute the needed term from lx[i], ly[i] with exactly the
// operations MINPACK would have applied to the stored arrays, in the same order.  This removes all O(m)
// scratch traffic; the fit costs the latency of its ordered sums and nothing else.
enum { FIT_K_F = 0,      // f_i = (x0*lx + x1) - ly                       (enorm term)
       FIT_K_J0,         // forward-difference column of parameter 0      (enorm term)
       FIT_K_J1,         // forward-difference column of parameter 1      (enorm term)
       FIT_K_AB,         // A'_i * B_i   first Householder vector times second (pivoted) column   (dot term)
       FIT_K_AF,         // A'_i * f_i                                                           (dot term)
       FIT_K_B1,         // B'_i = B_i - t*A'_i   second column after the first reflection       (enorm term)
       FIT_K_B2W,        // B''_i * w'_i  second Householder vector times once-reflected residual (dot term)
       FIT_K_FNEW,       // residual at the trial point                   (enorm term)
       FIT_K_J0N,        // forward-difference columns AT the trial point: the next iteration's column norms,
       FIT_K_J1N };      // summed speculatively next to FIT_K_FNEW (they are needed iff the step is accepted)


FIT_FN bool fit_kind_is_enorm(int kind) { return kind == FIT_K_F || kind == FIT_K_J0 || kind == FIT_K_J1 || kind == FIT_K_B1 || kind >= FIT_K_FNEW; }

FIT_FN double fit_elem_f(const FitState& S, double lx, double ly) { return FSUB(FADD(FMUL(S.x0, lx), S.x1), ly); }
FIT_FN double fit_elem_j0(const FitState& S, double lx, double ly, double f) { return fit_div(FSUB(FSUB(FADD(FMUL(S.xp00, lx), S.xp01), ly), f), S.h0, S.rh0); }
FIT_FN double fit_elem_j1(const FitState& S, double lx, double ly, double f) { return fit_div(FSUB(FSUB(FADD(FMUL(S.xp10, lx), S.xp11), ly), f), S.h1, S.rh1); }
FIT_FN double fit_elem_ap(const FitState& S, long i, double A)
{
    if (!S.hh1) return A;
    const double ap = fit_div(A, S.aj0, S.raj0);
    return i == 0 ? FADD(ap, 1.0) : ap;
}
FIT_FN double fit_elem_bp(const FitState& S, double B, double Ap) { return S.hh1 ? FSUB(B, FMUL(S.t, Ap)) : B; }
FIT_FN double fit_elem_bpp(const FitState& S, long i, double Bp)
{
    if (!S.hh2 || i < 1) return Bp;
    const double b = fit_div(Bp, S.aj1, S.raj1);
    return i == 1 ? FADD(b, 1.0) : b;
}
FIT_FN double fit_elem_wp(const FitState& S, double f, double Ap) { return S.hh1 ? FADD(f, FMUL(Ap, S.temp0)) : f; }

// value of the chain term `kind` at absolute index i
FIT_FN double fit_term(int kind, long i, double lx, double ly, const FitState& S)
{
    if (kind >= FIT_K_FNEW) {
        const double fn = FSUB(FADD(FMUL(S.xn0, lx), S.xn1), ly);
        if (kind == FIT_K_FNEW) return fn;
        if (kind == FIT_K_J0N) return fit_div(FSUB(FSUB(FADD(FMUL(S.xnp0, lx), S.xn1), ly), fn), S.hn0, S.rhn0);
        return fit_div(FSUB(FSUB(FADD(FMUL(S.xn0, lx), S.xnp1), ly), fn), S.hn1, S.rhn1);
    }
    const double f = fit_elem_f(S, lx, ly);
    if (kind == FIT_K_F) return f;
    if (kind == FIT_K_J0) return fit_elem_j0(S, lx, ly, f);
    if (kind == FIT_K_J1) return fit_elem_j1(S, lx, ly, f);
    const double J0 = fit_elem_j0(S, lx, ly, f), J1 = fit_elem_j1(S, lx, ly, f);
    const double A = S.piv ? J1 : J0, B = S.piv ? J0 : J1;
    const double Ap = fit_elem_ap(S, i, A);
    if (kind == FIT_K_AB) return FMUL(Ap, B);
    if (kind == FIT_K_AF) return FMUL(Ap, f);
    const double Bp = fit_elem_bp(S, B, Ap);
    if (kind == FIT_K_B1) return Bp;
    return FMUL(fit_elem_bpp(S, i, Bp), fit_elem_wp(S, f, Ap));  // FIT_K_B2W
}

#if defined(__CUDA_ARCH__)
__device__ __forceinline__ void fit_cp_async8(void* smem, const void* gmem)
{
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(gmem) : "memory");
}
#endif

// All terms of one element for the (up to three) chains of a call, sharing the sub-expressions they have in
// common (the residual, the two difference quotients, the Householder vector element).  Same operations, same
// order as fit_term.
FIT_FN void fit_terms(int nv, int k0, int k1, int k2, long i, double lx, double ly, const FitState& S, double* tv)
{
    const int ks[3] = {k0, k1, k2};
    const int kmax = nv == 1 ? k0 : (nv == 2 ? (k0 > k1 ? k0 : k1) : (k0 > k1 ? (k0 > k2 ? k0 : k2) : (k1 > k2 ? k1 : k2)));
    const int kmin = nv == 1 ? k0 : (nv == 2 ? (k0 < k1 ? k0 : k1) : (k0 < k1 ? (k0 < k2 ? k0 : k2) : (k1 < k2 ? k1 : k2)));
    double fn = 0.0, j0n = 0.0, j1n = 0.0;
    if (kmax >= FIT_K_FNEW) {
        fn = FSUB(FADD(FMUL(S.xn0, lx), S.xn1), ly);
        if (kmax >= FIT_K_J0N) {
            j0n = fit_div(FSUB(FSUB(FADD(FMUL(S.xnp0, lx), S.xn1), ly), fn), S.hn0, S.rhn0);
            j1n = fit_div(FSUB(FSUB(FADD(FMUL(S.xn0, lx), S.xnp1), ly), fn), S.hn1, S.rhn1);
        }
    }
    double f = 0.0, J0 = 0.0, J1 = 0.0, Ap = 0.0, B = 0.0, Bp = 0.0, b2w = 0.0;
    if (kmin < FIT_K_FNEW) {
        f = fit_elem_f(S, lx, ly);
        const int kc = kmax < FIT_K_FNEW ? kmax : FIT_K_B2W;  // highest current-point kind that may be present
        if (kc >= FIT_K_J0) {
            J0 = fit_elem_j0(S, lx, ly, f);
            J1 = fit_elem_j1(S, lx, ly, f);
        }
        if (kc >= FIT_K_AB) {
            const double A = S.piv ? J1 : J0;
            B = S.piv ? J0 : J1;
            Ap = fit_elem_ap(S, i, A);
            if (kc >= FIT_K_B1) Bp = fit_elem_bp(S, B, Ap);
            if (kc >= FIT_K_B2W) b2w = FMUL(fit_elem_bpp(S, i, Bp), fit_elem_wp(S, f, Ap));
        }
    }
#pragma unroll
    for (int v = 0; v < 3; v++) {
        if (v < nv) {
            const int k = ks[v];
            tv[v] = k == FIT_K_F ? f : k == FIT_K_J0 ? J0 : k == FIT_K_J1 ? J1 : k == FIT_K_AB ? FMUL(Ap, B) : k == FIT_K_AF ? FMUL(Ap, f)
                  : k == FIT_K_B1 ? Bp : k == FIT_K_B2W ? b2w : k == FIT_K_FNEW ? fn : k == FIT_K_J0N ? j0n : j1n;
        }
    }
}

#ifndef FIT_SKIP_SUMS
#define FIT_SKIP_SUMS 0   // tools/micro/fit_timing.cu sets 1 to time the staging warps alone
#endif

// Up to FIT_MAXV chains side by side over the absolute index range [i0, i0 + n).  Chain v is an enorm (MINPACK's
// scaled 2-norm) or a plain ordered sum depending on its kind; it is accumulated in index order by lane 0 of
// warp v from terms staged through a ring of shared-memory chunks by the remaining warps.  With fewer than 128
// threads (the host build) staging and accumulation simply alternate.
FIT_FN void fit_chain_multi(int nv, int k0, int k1, int k2, long i0, long n, const double* lx, const double* ly,
                            const FitState& S_local, FitShared* sh, double* out)
{
    // every thread carries the same scalars; the staging loops read them from shared memory, which keeps them out
    // of the registers of the hot loops
    if (FIT_TID == 0) sh->st = S_local;
    FIT_SYNC();
    const FitState& S = sh->st;
    const double rdwarf = 3.834e-20;
    const double agiant = FDIV(1.304e19, (double)n);
    const bool en0 = fit_kind_is_enorm(k0), en1 = fit_kind_is_enorm(k1), en2 = fit_kind_is_enorm(k2);
    const bool overlap = FIT_NT >= 128;
    if (!overlap) {
        for (int v = 0; v < nv; v++) {
            const int kind = v == 0 ? k0 : (v == 1 ? k1 : k2);
            const bool en = fit_kind_is_enorm(kind);
            FitEnormAcc acc;
            acc.init((double)n);
            double sum = 0.0;
            for (long c0 = 0; c0 < n; c0 += FIT_CH) {
                const int len = (int)((n - c0) < FIT_CH ? (n - c0) : FIT_CH);
                for (int i = FIT_TID; i < len; i += FIT_NT) {
                    const long gi = i0 + c0 + i;
                    double tv[3];
                    fit_terms(nv, k0, k1, k2, gi, lx[gi], ly[gi], S, tv);
                    sh->buf[0][0][i] = tv[v];
                }
                FIT_SYNC();
                if (FIT_TID == 0) {
                    if (en) acc.add_run(sh->buf[0][0], len);
                    else sum = fit_ordered_sum(sum, sh->buf[0][0], len);
                }
                FIT_SYNC();
            }
            if (FIT_TID == 0) sh->out[v] = en ? acc.result() : sum;
        }
    } else {
        const int my_v = (FIT_TID & 31) == 0 ? (FIT_TID >> 5) : -1;
        const bool summer = my_v >= 0 && my_v < nv;
        const int my_kind = my_v == 0 ? k0 : (my_v == 1 ? k1 : k2);
        const bool my_en = summer && fit_kind_is_enorm(my_kind);
        FitEnormAcc acc;
        acc.init((double)n);
        double sum = 0.0;
        const int first_loader = nv * 32, n_loaders = FIT_NT - nv * 32;  // warps without a chain to sum stage terms
        const int lt = FIT_TID - first_loader;
        const long n_chunks = (n + FIT_CH - 1) / FIT_CH;
        // Staging warps: element k of a chunk belongs to loader (k mod n_loaders).  Its (lx, ly) are copied into shared
        // memory by cp.async one chunk ahead of their use, so the global-memory latency hides behind the summing warps
        // and costs no registers.
        auto fetch_chunk = [&](long cl) {
            const long c1 = cl * FIT_CH;
            if (c1 < n) {
                const int len = (int)((n - c1) < FIT_CH ? (n - c1) : FIT_CH);
                const int rs = (int)(cl % FIT_RING);
                for (int i = lt; i < len; i += n_loaders) {
                    const long gi = i0 + c1 + i;
#if defined(__CUDA_ARCH__)
                    fit_cp_async8(&sh->rx[rs][i], lx + gi);
                    fit_cp_async8(&sh->ry[rs][i], ly + gi);
#else
                    sh->rx[rs][i] = lx[gi];
                    sh->ry[rs][i] = ly[gi];
#endif
                }
            }
#if defined(__CUDA_ARCH__)
            asm volatile("cp.async.commit_group;" ::: "memory");
#endif
        };
        if (FIT_TID >= first_loader) fetch_chunk(0);
        for (long c = -(FIT_RING - 1); c < n_chunks; c++) {
            if (FIT_TID >= first_loader) {
                const long cl = c + (FIT_RING - 1);  // chunk to stage now
                const long c1 = cl * FIT_CH;
                fetch_chunk(cl + 1);
#if defined(__CUDA_ARCH__)
                asm volatile("cp.async.wait_group 1;" ::: "memory");  // everything but the newest group has landed
#endif
                if (c1 < n) {
                    const int slot = (int)(cl % FIT_RING);
                    const int len = (int)((n - c1) < FIT_CH ? (n - c1) : FIT_CH);
                    bool odd0 = false, odd1 = false, odd2 = false;
                    for (int i = lt; i < len; i += n_loaders) {
                        double tv[3];
                        fit_terms(nv, k0, k1, k2, i0 + c1 + i, sh->rx[slot][i], sh->ry[slot][i], S, tv);
                        {
                            const double av = fabs(tv[0]);
                            if (en0) odd0 = odd0 || !(av > rdwarf && av < agiant);
                            sh->buf[0][slot][i] = en0 ? FMUL(av, av) : tv[0];
                        }
                        if (nv > 1) {
                            const double av = fabs(tv[1]);
                            if (en1) odd1 = odd1 || !(av > rdwarf && av < agiant);
                            sh->buf[1][slot][i] = en1 ? FMUL(av, av) : tv[1];
                        }
                        if (nv > 2) {
                            const double av = fabs(tv[2]);
                            if (en2) odd2 = odd2 || !(av > rdwarf && av < agiant);
                            sh->buf[2][slot][i] = en2 ? FMUL(av, av) : tv[2];
                        }
                    }
                    if (odd0) sh->tag[0][slot] = (int)(cl + 1);
                    if (odd1) sh->tag[1][slot] = (int)(cl + 1);
                    if (odd2) sh->tag[2][slot] = (int)(cl + 1);
                }
            } else if (c >= 0 && summer && !FIT_SKIP_SUMS) {
                const long c0 = c * FIT_CH;
                const int len = (int)((n - c0) < FIT_CH ? (n - c0) : FIT_CH);
                const int slot = (int)(c % FIT_RING);
                if (!my_en) {
                    sum = fit_ordered_sum(sum, sh->buf[my_v][slot], len);
                } else if (sh->tag[my_v][slot] == (int)(c + 1)) {  // a value outside MINPACK's mid range: exact slow path
                    for (int i = 0; i < len; i++) {
                        const long gi = i0 + c0 + i;
                        acc.add(fit_term(my_kind, gi, lx[gi], ly[gi], S));
                    }
                } else {
                    acc.s2 = fit_ordered_sum(acc.s2, sh->buf[my_v][slot], len);
                }
            }
            FIT_SYNC();
        }
        if (summer) sh->out[my_v] = my_en ? acc.result() : sum;
    }
    FIT_SYNC();
    for (int v = 0; v < nv; v++) out[v] = sh->out[v];
    FIT_SYNC();
}

FIT_FN double fit_chain_one(int kind, long i0, long n, const double* lx, const double* ly, const FitState& S, FitShared* sh)
{
    double o[3];
    fit_chain_multi(1, kind, kind, kind, i0, n, lx, ly, S, sh, o);
    return o[0];
}

FIT_FN void fit_qrsolv(double r[2][2], const int* ipvt, const double* diag, const double* qtb, double* x, double* sdiag)
{
    double wa[2];
    for (int j = 0; j < 2; j++) {
        for (int i = j; i < 2; i++) r[i][j] = r[j][i];
        x[j] = r[j][j];
        wa[j] = qtb[j];
    }
    for (int j = 0; j < 2; j++) {
        const int l = ipvt[j];
        if (diag[l] != 0.0) {
            for (int k = j; k < 2; k++) sdiag[k] = 0.0;
            sdiag[j] = diag[l];
            double qtbpj = 0.0;
            for (int k = j; k < 2; k++) {
                if (sdiag[k] == 0.0) continue;
                double c, s;
                if (fabs(r[k][k]) < fabs(sdiag[k])) {
                    const double cotan = FDIV(r[k][k], sdiag[k]);
                    s = FDIV(0.5, FSQRT(FADD(0.25, FMUL(0.25, FMUL(cotan, cotan)))));
                    c = FMUL(s, cotan);
                } else {
                    const double t = FDIV(sdiag[k], r[k][k]);
                    c = FDIV(0.5, FSQRT(FADD(0.25, FMUL(0.25, FMUL(t, t)))));
                    s = FMUL(c, t);
                }
                r[k][k] = FADD(FMUL(c, r[k][k]), FMUL(s, sdiag[k]));
                double temp = FADD(FMUL(c, wa[k]), FMUL(s, qtbpj));
                qtbpj = FADD(FMUL(-s, wa[k]), FMUL(c, qtbpj));
                wa[k] = temp;
                for (int i = k + 1; i < 2; i++) {
                    temp = FADD(FMUL(c, r[i][k]), FMUL(s, sdiag[i]));
                    sdiag[i] = FADD(FMUL(-s, r[i][k]), FMUL(c, sdiag[i]));
                    r[i][k] = temp;
                }
            }
        }
        sdiag[j] = r[j][j];
        r[j][j] = x[j];
    }
    int nsing = 2;
    for (int j = 0; j < 2; j++) {
        if (sdiag[j] == 0.0 && nsing == 2) nsing = j;
        if (nsing < 2) wa[j] = 0.0;
    }
    for (int k = 1; k <= nsing; k++) {
        const int j = nsing - k;
        double sum = 0.0;
        for (int i = j + 1; i < nsing; i++) sum = FADD(sum, FMUL(r[i][j], wa[i]));
        wa[j] = FDIV(FSUB(wa[j], sum), sdiag[j]);
    }
    for (int j = 0; j < 2; j++) x[ipvt[j]] = wa[j];
}

FIT_FN void fit_lmpar(double r[2][2], const int* ipvt, const double* diag, const double* qtb, double delta, double* par,
                      double* x, double* sdiag, int* iterative)
{
    const double dwarf = DBL_MIN;
    double wa1[2], wa2[2];
    int nsing = 2;
    for (int j = 0; j < 2; j++) {
        wa1[j] = qtb[j];
        if (r[j][j] == 0.0 && nsing == 2) nsing = j;
        if (nsing < 2) wa1[j] = 0.0;
    }
    for (int k = 1; k <= nsing; k++) {
        const int j = nsing - k;
        wa1[j] = FDIV(wa1[j], r[j][j]);
        const double temp = wa1[j];
        for (int i = 0; i < j; i++) wa1[i] = FSUB(wa1[i], FMUL(r[i][j], temp));
    }
    for (int j = 0; j < 2; j++) x[ipvt[j]] = wa1[j];
    int iter = 0;
    for (int j = 0; j < 2; j++) wa2[j] = FMUL(diag[j], x[j]);
    double dxnorm = fit_enorm_small(2, wa2);
    double fp = FSUB(dxnorm, delta);
    if (fp <= FMUL(0.1, delta)) { *par = 0.0; return; }
    *iterative = 1;
    double parl = 0.0;
    if (nsing >= 2) {
        for (int j = 0; j < 2; j++) { const int l = ipvt[j]; wa1[j] = FMUL(diag[l], FDIV(wa2[l], dxnorm)); }
        for (int j = 0; j < 2; j++) {
            double sum = 0.0;
            for (int i = 0; i < j; i++) sum = FADD(sum, FMUL(r[i][j], wa1[i]));
            wa1[j] = FDIV(FSUB(wa1[j], sum), r[j][j]);
        }
        const double temp = fit_enorm_small(2, wa1);
        parl = FDIV(FDIV(FDIV(fp, delta), temp), temp);
    }
    for (int j = 0; j < 2; j++) {
        double sum = 0.0;
        for (int i = 0; i <= j; i++) sum = FADD(sum, FMUL(r[i][j], qtb[i]));
        wa1[j] = FDIV(sum, diag[ipvt[j]]);
    }
    const double gnorm = fit_enorm_small(2, wa1);
    double paru = FDIV(gnorm, delta);
    if (paru == 0.0) paru = FDIV(dwarf, fmin(delta, 0.1));
    *par = fmax(*par, parl);
    *par = fmin(*par, paru);
    if (*par == 0.0) *par = FDIV(gnorm, dxnorm);
    for (;;) {
        iter++;
        if (*par == 0.0) *par = fmax(dwarf, FMUL(0.001, paru));
        double temp = FSQRT(*par);
        for (int j = 0; j < 2; j++) wa1[j] = FMUL(temp, diag[j]);
        fit_qrsolv(r, ipvt, wa1, qtb, x, sdiag);
        for (int j = 0; j < 2; j++) wa2[j] = FMUL(diag[j], x[j]);
        dxnorm = fit_enorm_small(2, wa2);
        temp = fp;
        fp = FSUB(dxnorm, delta);
        if (fabs(fp) <= FMUL(0.1, delta) || (parl == 0.0 && fp <= temp && temp < 0.0) || iter == 10) break;
        for (int j = 0; j < 2; j++) { const int l = ipvt[j]; wa1[j] = FMUL(diag[l], FDIV(wa2[l], dxnorm)); }
        for (int j = 0; j < 2; j++) {
            wa1[j] = FDIV(wa1[j], sdiag[j]);
            temp = wa1[j];
            for (int i = j + 1; i < 2; i++) wa1[i] = FSUB(wa1[i], FMUL(r[i][j], temp));
        }
        temp = fit_enorm_small(2, wa1);
        const double parc = FDIV(FDIV(FDIV(fp, delta), temp), temp);
        if (fp > 0.0) parl = fmax(parl, *par);
        if (fp < 0.0) paru = fmin(paru, *par);
        *par = fmax(parl, FADD(*par, parc));
    }
    if (iter == 0) *par = 0.0;
}

// lx, ly: m selected points in the reference's order.  out[0..5] = a, b, m, info, nfev, iterative-lmpar flag.
// Returns MINPACK info.
FIT_FN int fit_loglinear(const double* lx, const double* ly, long m, FitShared* sh, double* out)
{
    const double ftol = 1.49012e-08, xtol = 1.49012e-08, gtol = 0.0, factor = 100.0, epsmch = DBL_EPSILON;
    const int maxfev = 600;
    double x[2] = {1.0, 1.0}, diag[2], qtf[2], wa1[2], wa2[2], wa3[2], acnorm[2], rdiag[2], wa[2];
    int ipvt[2];
    int info = 0, nfev = 0, iterative = 0, iter = 1;
    double par = 0.0, delta = 0.0, xnorm = 0.0, gnorm = 0.0;
    const double eps = FSQRT(epsmch);
    const double lx0 = lx[0], ly0 = ly[0], lx1 = m > 1 ? lx[1] : 0.0, ly1 = m > 1 ? ly[1] : 0.0;
    FitState S;
    S.xn0 = S.xn1 = 0.0;
    S.xnp0 = S.xnp1 = 0.0; S.hn0 = S.hn1 = 1.0; S.rhn0 = S.rhn1 = 0.0;

    nfev = 1;
    double fnorm = 0.0;
    bool have_fnorm = false;
    bool spec_valid = false;      // spec_acnorm holds the column norms at the current x
    double spec_acnorm[2] = {0.0, 0.0};
    for (;;) {
        // fdjac2: forward differences, h_j = eps*|x_j|
        S.x0 = x[0]; S.x1 = x[1];
        S.piv = 0; S.hh1 = 0; S.hh2 = 0; S.aj0 = S.aj1 = 1.0; S.raj0 = S.raj1 = 0.0; S.t = S.temp0 = 0.0;
        {
            double h[2];
            for (int j = 0; j < 2; j++) {
                h[j] = FMUL(eps, fabs(x[j]));
                if (h[j] == 0.0) h[j] = eps;
            }
            S.h0 = h[0]; S.xp00 = FADD(x[0], h[0]); S.xp01 = x[1]; S.rh0 = fit_rcp(h[0]); S.rh1 = fit_rcp(h[1]);
            S.h1 = h[1]; S.xp10 = x[0]; S.xp11 = FADD(x[1], h[1]);
        }
        nfev += 2;
        // qrfac with column pivoting; the initial |fvec| rides along with the two column norms
        if (spec_valid) {  // summed next to the accepted trial residual
            acnorm[0] = spec_acnorm[0];
            acnorm[1] = spec_acnorm[1];
            spec_valid = false;
        } else {
            double o[3];
            fit_chain_multi(have_fnorm ? 2 : 3, FIT_K_J0, FIT_K_J1, FIT_K_F, 0, m, lx, ly, S, sh, o);
            acnorm[0] = o[0];
            acnorm[1] = o[1];
            if (!have_fnorm) { fnorm = o[2]; have_fnorm = true; }
        }
        for (int j = 0; j < 2; j++) {
            rdiag[j] = acnorm[j];
            wa[j] = rdiag[j];
            ipvt[j] = j;
        }
        // j = 0: pivot, first Householder reflection
        double qtf_dot0 = 0.0;
        double r01 = 0.0;  // element (0,1) of R: second column's first entry after the first reflection
        {
            if (rdiag[1] > rdiag[0]) {
                S.piv = 1;
                rdiag[1] = rdiag[0];
                wa[1] = wa[0];
                ipvt[0] = 1; ipvt[1] = 0;
            }
            // elements 0 and 1 of the pivoted raw columns
            const double f0 = fit_elem_f(S, lx0, ly0), f1 = fit_elem_f(S, lx1, ly1);
            const double J00 = fit_elem_j0(S, lx0, ly0, f0), J10 = fit_elem_j1(S, lx0, ly0, f0);
            const double A0 = S.piv ? J10 : J00, B0 = S.piv ? J00 : J10;
            double ajnorm = acnorm[ipvt[0]];  // enorm(a[0:, 0]) of the pivoted column
            if (ajnorm != 0.0) {
                if (A0 < 0.0) ajnorm = -ajnorm;
                S.aj0 = ajnorm; S.raj0 = fit_rcp(ajnorm);
                S.hh1 = 1;
                const double Ap0 = fit_elem_ap(S, 0, A0);
                // A'.B and A'.f (the latter is needed for qtf) side by side
                double o[3];
                fit_chain_multi(2, FIT_K_AB, FIT_K_AF, FIT_K_AF, 0, m, lx, ly, S, sh, o);
                qtf_dot0 = o[1];
                S.t = FDIV(o[0], Ap0);
                r01 = fit_elem_bp(S, B0, Ap0);
                if (rdiag[1] != 0.0) {
                    const double t2 = FDIV(r01, rdiag[1]);
                    rdiag[1] = FMUL(rdiag[1], FSQRT(fmax(0.0, FSUB(1.0, FMUL(t2, t2)))));
                    const double q = FDIV(rdiag[1], wa[1]);
                    if (FMUL(0.05, FMUL(q, q)) <= epsmch) {
                        rdiag[1] = fit_chain_one(FIT_K_B1, 1, m - 1, lx, ly, S, sh);
                        wa[1] = rdiag[1];
                    }
                }
            } else {
                r01 = B0;
            }
            rdiag[0] = -ajnorm;
            (void)f1;
        }
        // j = 1: second Householder reflection (no pivot choice left)
        if (m >= 2) {
            double ajnorm = fit_chain_one(FIT_K_B1, 1, m - 1, lx, ly, S, sh);
            if (ajnorm != 0.0) {
                const double f1 = fit_elem_f(S, lx1, ly1);
                const double J01 = fit_elem_j0(S, lx1, ly1, f1), J11 = fit_elem_j1(S, lx1, ly1, f1);
                const double A1 = S.piv ? J11 : J01, B1 = S.piv ? J01 : J11;
                const double Bp1 = fit_elem_bp(S, B1, fit_elem_ap(S, 1, A1));
                if (Bp1 < 0.0) ajnorm = -ajnorm;
                S.aj1 = ajnorm; S.raj1 = fit_rcp(ajnorm);
                S.hh2 = 1;
            }
            rdiag[1] = -ajnorm;
        }
        if (iter == 1) {
            for (int j = 0; j < 2; j++) { diag[j] = acnorm[j]; if (acnorm[j] == 0.0) diag[j] = 1.0; }
            for (int j = 0; j < 2; j++) wa3[j] = FMUL(diag[j], x[j]);
            xnorm = fit_enorm_small(2, wa3);
            delta = FMUL(factor, xnorm);
            if (delta == 0.0) delta = factor;
        }
        // qtf: Q^T f through the two reflections
        {
            const double f0 = fit_elem_f(S, lx0, ly0), f1 = fit_elem_f(S, lx1, ly1);
            const double J00 = fit_elem_j0(S, lx0, ly0, f0), J10 = fit_elem_j1(S, lx0, ly0, f0);
            const double J01 = fit_elem_j0(S, lx1, ly1, f1), J11 = fit_elem_j1(S, lx1, ly1, f1);
            const double A0 = S.piv ? J10 : J00, A1 = S.piv ? J11 : J01, B1 = S.piv ? J01 : J11;
            const double Ap0 = fit_elem_ap(S, 0, A0), Ap1 = fit_elem_ap(S, 1, A1);
            if (S.hh1) S.temp0 = FDIV(-qtf_dot0, Ap0);
            qtf[0] = fit_elem_wp(S, f0, Ap0);
            double w1 = fit_elem_wp(S, f1, Ap1);
            if (S.hh2) {
                const double Bpp1 = fit_elem_bpp(S, 1, fit_elem_bp(S, B1, Ap1));
                const double sum = fit_chain_one(FIT_K_B2W, 1, m - 1, lx, ly, S, sh);
                const double temp = FDIV(-sum, Bpp1);
                w1 = FADD(w1, FMUL(Bpp1, temp));
            }
            qtf[1] = w1;
        }
        double r[2][2] = {{rdiag[0], r01}, {0.0, rdiag[1]}};
        gnorm = 0.0;
        if (fnorm != 0.0) {
            for (int j = 0; j < 2; j++) {
                const int l = ipvt[j];
                if (acnorm[l] == 0.0) continue;
                double sum = 0.0;
                for (int i = 0; i <= j; i++) sum = FADD(sum, FMUL(r[i][j], FDIV(qtf[i], fnorm)));
                gnorm = fmax(gnorm, fabs(FDIV(sum, acnorm[l])));
            }
        }
        if (gnorm <= gtol) { info = 4; break; }
        for (int j = 0; j < 2; j++) diag[j] = fmax(diag[j], acnorm[j]);
        double ratio;
        do {
            double sdiag[2], rr[2][2];
            for (int i = 0; i < 2; i++) for (int j = 0; j < 2; j++) rr[i][j] = r[i][j];
            fit_lmpar(rr, ipvt, diag, qtf, delta, &par, wa1, sdiag, &iterative);
            for (int j = 0; j < 2; j++) {
                wa1[j] = -wa1[j];
                wa2[j] = FADD(x[j], wa1[j]);
                wa3[j] = FMUL(diag[j], wa1[j]);
            }
            const double pnorm = fit_enorm_small(2, wa3);
            if (iter == 1) delta = fmin(delta, pnorm);
            S.xn0 = wa2[0]; S.xn1 = wa2[1];
            {   // fdjac2's steps at the trial point
                double hn0 = FMUL(eps, fabs(wa2[0])), hn1 = FMUL(eps, fabs(wa2[1]));
                if (hn0 == 0.0) hn0 = eps;
                if (hn1 == 0.0) hn1 = eps;
                S.hn0 = hn0; S.xnp0 = FADD(wa2[0], hn0); S.rhn0 = fit_rcp(hn0);
                S.hn1 = hn1; S.xnp1 = FADD(wa2[1], hn1); S.rhn1 = fit_rcp(hn1);
            }
            nfev++;
            double fnorm1;
            {
                double o[3];
                fit_chain_multi(3, FIT_K_FNEW, FIT_K_J0N, FIT_K_J1N, 0, m, lx, ly, S, sh, o);
                fnorm1 = o[0];
                spec_acnorm[0] = o[1];
                spec_acnorm[1] = o[2];
            }
            double actred = -1.0;
            if (FMUL(0.1, fnorm1) < fnorm) { const double q = FDIV(fnorm1, fnorm); actred = FSUB(1.0, FMUL(q, q)); }
            for (int j = 0; j < 2; j++) {
                wa3[j] = 0.0;
                const double temp = wa1[ipvt[j]];
                for (int i = 0; i <= j; i++) wa3[i] = FADD(wa3[i], FMUL(r[i][j], temp));
            }
            const double temp1 = FDIV(fit_enorm_small(2, wa3), fnorm);
            const double temp2 = FDIV(FMUL(FSQRT(par), pnorm), fnorm);
            const double prered = FADD(FMUL(temp1, temp1), FDIV(FMUL(temp2, temp2), 0.5));
            const double dirder = -FADD(FMUL(temp1, temp1), FMUL(temp2, temp2));
            ratio = 0.0;
            if (prered != 0.0) ratio = FDIV(actred, prered);
            if (ratio <= 0.25) {
                double temp = 0.5;
                if (actred < 0.0) temp = FDIV(FMUL(0.5, dirder), FADD(dirder, FMUL(0.5, actred)));
                if (FMUL(0.1, fnorm1) >= fnorm || temp < 0.1) temp = 0.1;
                delta = FMUL(temp, fmin(delta, FDIV(pnorm, 0.1)));
                par = FDIV(par, temp);
            } else if (par == 0.0 || ratio >= 0.75) {
                delta = FDIV(pnorm, 0.5);
                par = FMUL(0.5, par);
            }
            if (ratio >= 1e-4) {
                for (int j = 0; j < 2; j++) { x[j] = wa2[j]; wa2[j] = FMUL(diag[j], x[j]); }  // f(x) follows implicitly
                xnorm = fit_enorm_small(2, wa2);
                fnorm = fnorm1;
                iter++;
                spec_valid = true;
            }
            if (fabs(actred) <= ftol && prered <= ftol && FMUL(0.5, ratio) <= 1.0) info = 1;
            if (delta <= FMUL(xtol, xnorm)) info = 2;
            if (fabs(actred) <= ftol && prered <= ftol && FMUL(0.5, ratio) <= 1.0 && info == 2) info = 3;
            if (info != 0) break;
            if (nfev >= maxfev) info = 5;
            if (fabs(actred) <= epsmch && prered <= epsmch && FMUL(0.5, ratio) <= 1.0) info = 6;
            if (delta <= FMUL(epsmch, xnorm)) info = 7;
            if (gnorm <= epsmch) info = 8;
            if (info != 0) break;
        } while (ratio < 1e-4);
        if (info != 0) break;
    }
    if (FIT_TID == 0) {
        out[0] = x[0];
        out[1] = x[1];
        out[2] = (double)m;
        out[3] = (double)info;
        out[4] = (double)nfev;
        out[5] = (double)iterative;
    }
    return info;
}
