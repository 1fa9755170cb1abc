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
// against scipy and the oracle (tests/test_fit_core_host.py).
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
#endif

#define FIT_CH 512   // doubles staged per chunk and buffer for the sequential chains
#define FIT_MAXV 3   // chains that can run side by side (one per warp)

#define FIT_RING 3    // staging ring depth: chunk c+2 is being fetched while chunk c is accumulated

struct FitShared {  // lives in shared memory on the device (36.9 KB)
    double buf[FIT_MAXV][FIT_RING][FIT_CH];  // staged squares (enorm) / products (dot)
    double out[FIT_MAXV];
    int tag[FIT_MAXV][FIT_RING];      // chunk tag set when a staged chunk holds a value outside MINPACK's mid range
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

// enorm of up to FIT_MAXV long vectors side by side.  Vector v is accumulated in index order by lane 0
// of warp v while the warps >= FIT_MAXV stage the next chunk of every vector (double buffering).  With
// fewer than 128 threads (the host build) staging and accumulation simply alternate.
FIT_FN void fit_enorm_multi(int nv, const double* const* x, const long* n, FitShared* sh, double* out)
{
    FitEnormAcc acc;
    acc.init(1.0);
    const bool overlap = FIT_NT >= 128;
    const int my_v = (FIT_TID & 31) == 0 ? (FIT_TID >> 5) : -1;  // chain this thread accumulates (device)
    long nmax = 0;
    for (int v = 0; v < nv; v++) nmax = n[v] > nmax ? n[v] : nmax;
    if (!overlap) {
        for (int v = 0; v < nv; v++) {
            acc.init((double)n[v]);
            for (long c0 = 0; c0 < n[v]; c0 += FIT_CH) {
                const int len = (int)((n[v] - c0) < FIT_CH ? (n[v] - c0) : FIT_CH);
                for (int i = FIT_TID; i < len; i += FIT_NT) sh->buf[0][0][i] = x[v][c0 + i];
                FIT_SYNC();
                if (FIT_TID == 0) acc.add_run(sh->buf[0][0], len);
                FIT_SYNC();
            }
            if (FIT_TID == 0) sh->out[v] = acc.result();
        }
    } else {
        if (my_v >= 0 && my_v < nv) acc.init((double)n[my_v]);
        const int first_loader = FIT_MAXV * 32, n_loaders = FIT_NT - FIT_MAXV * 32;
        const long n_chunks = (nmax + FIT_CH - 1) / FIT_CH;
        // The loaders stage the SQUARES of chunk c+1 (independent, any order) while lane 0 of warp v adds the
        // squares of chunk c in index order: same operations as FitEnormAcc::add for mid-range values.  A chunk
        // holding a value outside the mid range is tagged and accumulated from the raw values instead.
        const double rdwarf = 3.834e-20;
        const int lt = FIT_TID - first_loader;  // loader index
        for (long c = -(FIT_RING - 1); c < n_chunks; c++) {
            if (FIT_TID >= first_loader) {
                const long cl = c + (FIT_RING - 1);          // chunk to stage now
                const long c1 = cl * FIT_CH;
                const int slot = (int)(cl % FIT_RING);
                for (int v = 0; v < nv; v++) {
                    if (c1 >= n[v]) continue;
                    const int len = (int)((n[v] - c1) < FIT_CH ? (n[v] - c1) : FIT_CH);
                    const double agiant = FDIV(1.304e19, (double)n[v]);
                    bool odd = false;
                    for (int i0 = 0; i0 < len; i0 += 4 * n_loaders) {
                        double xv[4];
#pragma unroll
                        for (int u = 0; u < 4; u++) {  // four independent loads in flight per loader
                            const int i = i0 + u * n_loaders + lt;
                            xv[u] = i < len ? fabs(x[v][c1 + i]) : 1.0;
                        }
#pragma unroll
                        for (int u = 0; u < 4; u++) {
                            const int i = i0 + u * n_loaders + lt;
                            if (i < len) {
                                odd = odd || !(xv[u] > rdwarf && xv[u] < agiant);
                                sh->buf[v][slot][i] = FMUL(xv[u], xv[u]);
                            }
                        }
                    }
                    if (odd) sh->tag[v][slot] = (int)(cl + 1);
                }
            } else if (c >= 0 && my_v >= 0 && my_v < nv) {
                const long c0 = c * FIT_CH;
                if (c0 < n[my_v]) {
                    const int len = (int)((n[my_v] - c0) < FIT_CH ? (n[my_v] - c0) : FIT_CH);
                    const int slot = (int)(c % FIT_RING);
                    if (sh->tag[my_v][slot] == (int)(c + 1)) {
                        for (int i = 0; i < len; i++) acc.add(x[my_v][c0 + i]);
                    } else {
#ifdef FIT_PROFILE
                        const long long tp0 = clock64();
#endif
                        acc.s2 = fit_ordered_sum(acc.s2, sh->buf[my_v][slot], len);
#ifdef FIT_PROFILE
                        if (my_v == 0) { g_t_sum += clock64() - tp0; g_n_chunks++; }
#endif
                    }
                }
            }
#ifdef FIT_PROFILE
            const long long tp1 = clock64();
#endif
            FIT_SYNC();
#ifdef FIT_PROFILE
            if (FIT_TID == 0) g_t_wait += clock64() - tp1;
#endif
        }
        if (my_v >= 0 && my_v < nv) sh->out[my_v] = acc.result();
    }
    FIT_SYNC();
    for (int v = 0; v < nv; v++) out[v] = sh->out[v];
    FIT_SYNC();
}

FIT_FN double fit_enorm_big(const double* x, long n, FitShared* sh)
{
    const double* xs[1] = {x};
    long ns[1] = {n};
    double o[1];
    fit_enorm_multi(1, xs, ns, sh, o);
    return o[0];
}

// Up to FIT_MAXV dot products side by side: out[v] = sum_i a[v][i]*b[v][i], index order, product rounded
// before the add.  Chain v is accumulated by lane 0 of warp v from products staged (double buffered) by the
// warps >= FIT_MAXV; the host build alternates staging and accumulation.
FIT_FN void fit_dot_multi(int nv, const double* const* a, const double* const* b, const long* n, FitShared* sh, double* out)
{
    const bool overlap = FIT_NT >= 128;
    double sum = 0.0;
    if (!overlap) {
        for (int v = 0; v < nv; v++) {
            sum = 0.0;
            for (long c0 = 0; c0 < n[v]; c0 += FIT_CH) {
                const int len = (int)((n[v] - c0) < FIT_CH ? (n[v] - c0) : FIT_CH);
                for (int i = FIT_TID; i < len; i += FIT_NT) sh->buf[0][0][i] = FMUL(a[v][c0 + i], b[v][c0 + i]);
                FIT_SYNC();
                if (FIT_TID == 0) sum = fit_ordered_sum(sum, sh->buf[0][0], len);
                FIT_SYNC();
            }
            if (FIT_TID == 0) sh->out[v] = sum;
        }
    } else {
        const int my_v = (FIT_TID & 31) == 0 ? (FIT_TID >> 5) : -1;
        const int first_loader = FIT_MAXV * 32, n_loaders = FIT_NT - FIT_MAXV * 32;
        long nmax = 0;
        for (int v = 0; v < nv; v++) nmax = n[v] > nmax ? n[v] : nmax;
        const long n_chunks = (nmax + FIT_CH - 1) / FIT_CH;
        const int lt = FIT_TID - first_loader;
        for (long c = -(FIT_RING - 1); c < n_chunks; c++) {
            if (FIT_TID >= first_loader) {
                const long cl = c + (FIT_RING - 1);
                const long c1 = cl * FIT_CH;
                const int slot = (int)(cl % FIT_RING);
                for (int v = 0; v < nv; v++) {
                    if (c1 >= n[v]) continue;
                    const int len = (int)((n[v] - c1) < FIT_CH ? (n[v] - c1) : FIT_CH);
                    for (int i0 = 0; i0 < len; i0 += 4 * n_loaders) {
                        double av[4], bv[4];
#pragma unroll
                        for (int u = 0; u < 4; u++) {
                            const int i = i0 + u * n_loaders + lt;
                            av[u] = i < len ? a[v][c1 + i] : 0.0;
                            bv[u] = i < len ? b[v][c1 + i] : 0.0;
                        }
#pragma unroll
                        for (int u = 0; u < 4; u++) {
                            const int i = i0 + u * n_loaders + lt;
                            if (i < len) sh->buf[v][slot][i] = FMUL(av[u], bv[u]);
                        }
                    }
                }
            } else if (c >= 0 && my_v >= 0 && my_v < nv) {
                const long c0 = c * FIT_CH;
                if (c0 < n[my_v]) {
                    const int len = (int)((n[my_v] - c0) < FIT_CH ? (n[my_v] - c0) : FIT_CH);
                    sum = fit_ordered_sum(sum, sh->buf[my_v][c % FIT_RING], len);
                }
            }
            FIT_SYNC();
        }
        if (my_v >= 0 && my_v < nv) sh->out[my_v] = sum;
    }
    FIT_SYNC();
    for (int v = 0; v < nv; v++) out[v] = sh->out[v];
    FIT_SYNC();
}

FIT_FN double fit_dot_big(const double* a, const double* b, long n, FitShared* sh)
{
    const double* as[1] = {a};
    const double* bs[1] = {b};
    long ns[1] = {n};
    double o[1];
    fit_dot_multi(1, as, bs, ns, sh, o);
    return o[0];
}

// f(x,a,b) - y with f = a*log10(x) + b : three separately rounded operations (gathering.py:82)
FIT_FN void fit_resid(const double* lx, const double* ly, long m, double pa, double pb, double* out)
{
    for (long k = FIT_TID; k < m; k += FIT_NT) out[k] = FSUB(FADD(FMUL(pa, lx[k]), pb), ly[k]);
    FIT_SYNC();
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

// lx, ly: m selected points in the reference's order.  fvec, wa4, c0, c1: m doubles of scratch each.
// out[0..5] = a, b, m, info, nfev, iterative-lmpar flag.  Returns MINPACK info.
FIT_FN int fit_loglinear(const double* lx, const double* ly, long m, double* fvec, double* wa4, double* c0, double* c1,
                         FitShared* sh, double* out)
{
    const double ftol = 1.49012e-08, xtol = 1.49012e-08, gtol = 0.0, factor = 100.0, epsmch = DBL_EPSILON;
    const int maxfev = 600;
    double* col[2] = {c0, c1};
    double x[2] = {1.0, 1.0}, diag[2], qtf[2], wa1[2], wa2[2], wa3[2], acnorm[2], rdiag[2], wa[2];
    int ipvt[2];
    int info = 0, nfev = 0, iterative = 0, iter = 1;
    double par = 0.0, delta = 0.0, xnorm = 0.0, gnorm = 0.0;
    const double eps = FSQRT(epsmch);

    fit_resid(lx, ly, m, x[0], x[1], fvec);
    nfev = 1;
    double fnorm = 0.0;
    bool have_fnorm = false;
    for (;;) {
        // fdjac2: both forward-difference columns in one pass (same per-element operations as evaluating
        // f at x + h_j e_j into a scratch vector first)
        {
            double h[2], xp[2][2];
            for (int j = 0; j < 2; j++) {
                const double temp = x[j];
                h[j] = FMUL(eps, fabs(temp));
                if (h[j] == 0.0) h[j] = eps;
                xp[j][0] = x[0];
                xp[j][1] = x[1];
                xp[j][j] = FADD(temp, h[j]);
            }
            double* cA = col[0];
            double* cB = col[1];
            for (long i = FIT_TID; i < m; i += FIT_NT) {
                const double lxi = lx[i], lyi = ly[i], f = fvec[i];
                const double w0 = FSUB(FADD(FMUL(xp[0][0], lxi), xp[0][1]), lyi);
                const double w1 = FSUB(FADD(FMUL(xp[1][0], lxi), xp[1][1]), lyi);
                cA[i] = FDIV(FSUB(w0, f), h[0]);
                cB[i] = FDIV(FSUB(w1, f), h[1]);
            }
            FIT_SYNC();
        }
        nfev += 2;
        // qrfac with column pivoting; the initial |fvec| rides along with the two column norms
        {
            const double* xs[3] = {col[0], col[1], fvec};
            long ns[3] = {m, m, m};
            double o[3];
            fit_enorm_multi(have_fnorm ? 2 : 3, xs, ns, sh, o);
            acnorm[0] = o[0];
            acnorm[1] = o[1];
            if (!have_fnorm) { fnorm = o[2]; have_fnorm = true; }
        }
        for (int j = 0; j < 2; j++) {
            rdiag[j] = acnorm[j];
            wa[j] = rdiag[j];
            ipvt[j] = j;
        }
        const int minmn = m < 2 ? (int)m : 2;
        double qtf_dot0 = 0.0;
        bool have_qtf_dot0 = false;
        for (int j = 0; j < minmn; j++) {
            int kmax = j;
            for (int k = j; k < 2; k++) if (rdiag[k] > rdiag[kmax]) kmax = k;
            if (kmax != j) {
                double* t = col[j]; col[j] = col[kmax]; col[kmax] = t;
                rdiag[kmax] = rdiag[j];
                wa[kmax] = wa[j];
                const int k = ipvt[j]; ipvt[j] = ipvt[kmax]; ipvt[kmax] = k;
            }
            double* cj = col[j];
            // enorm(a[j:, j]); for j = 0 this is the full (pivoted) column, already computed as acnorm
            double ajnorm = (j == 0) ? acnorm[ipvt[0]] : fit_enorm_big(cj + j, m - j, sh);
            if (ajnorm != 0.0) {
                if (cj[j] < 0.0) ajnorm = -ajnorm;
                FIT_SYNC();
                for (long i = j + FIT_TID; i < m; i += FIT_NT) cj[i] = FDIV(cj[i], ajnorm);
                FIT_SYNC();
                if (FIT_TID == 0) cj[j] = FADD(cj[j], 1.0);
                FIT_SYNC();
                for (int k = j + 1; k < 2; k++) {
                    double* ck = col[k];
                    // the first Householder vector is final here: its product with the residual (needed for
                    // qtf below) rides along with its product with the second column
                    const double* das[2] = {cj + j, cj + j};
                    const double* dbs[2] = {ck + j, fvec + j};
                    long dns[2] = {m - j, m - j};
                    double dout[2];
                    fit_dot_multi(2, das, dbs, dns, sh, dout);
                    const double sum = dout[0];
                    qtf_dot0 = dout[1];
                    have_qtf_dot0 = true;
                    const double temp = FDIV(sum, cj[j]);
                    for (long i = j + FIT_TID; i < m; i += FIT_NT) ck[i] = FSUB(ck[i], FMUL(temp, cj[i]));
                    FIT_SYNC();
                    if (rdiag[k] != 0.0) {
                        const double t2 = FDIV(ck[j], rdiag[k]);
                        rdiag[k] = FMUL(rdiag[k], FSQRT(fmax(0.0, FSUB(1.0, FMUL(t2, t2)))));
                        const double q = FDIV(rdiag[k], wa[k]);
                        if (FMUL(0.05, FMUL(q, q)) <= epsmch) {
                            rdiag[k] = fit_enorm_big(ck + j + 1, m - j - 1, sh);
                            wa[k] = rdiag[k];
                        }
                    }
                }
            }
            rdiag[j] = -ajnorm;
        }
        if (iter == 1) {
            for (int j = 0; j < 2; j++) { diag[j] = acnorm[j]; if (acnorm[j] == 0.0) diag[j] = 1.0; }
            for (int j = 0; j < 2; j++) wa3[j] = FMUL(diag[j], x[j]);
            xnorm = fit_enorm_small(2, wa3);
            delta = FMUL(factor, xnorm);
            if (delta == 0.0) delta = factor;
        }
        // qtf (wa4 = fvec, then the two Householder reflections; the copy is folded into the first one)
        bool wa4_ready = false;
        for (int j = 0; j < 2; j++) {
            double* cj = col[j];
            if (j < m && cj[j] != 0.0) {
                if (j == 0 && have_qtf_dot0) {
                    const double temp = FDIV(-qtf_dot0, cj[0]);
                    for (long i = FIT_TID; i < m; i += FIT_NT) wa4[i] = FADD(fvec[i], FMUL(cj[i], temp));
                    FIT_SYNC();
                    wa4_ready = true;
                } else {
                    if (!wa4_ready) {
                        for (long i = FIT_TID; i < m; i += FIT_NT) wa4[i] = fvec[i];
                        FIT_SYNC();
                        wa4_ready = true;
                    }
                    const double sum = fit_dot_big(cj + j, wa4 + j, m - j, sh);
                    const double temp = FDIV(-sum, cj[j]);
                    for (long i = j + FIT_TID; i < m; i += FIT_NT) wa4[i] = FADD(wa4[i], FMUL(cj[i], temp));
                    FIT_SYNC();
                }
            } else if (!wa4_ready) {
                for (long i = FIT_TID; i < m; i += FIT_NT) wa4[i] = fvec[i];
                FIT_SYNC();
                wa4_ready = true;
            }
            FIT_SYNC();
            if (FIT_TID == 0 && j < m) cj[j] = rdiag[j];
            FIT_SYNC();
            qtf[j] = j < m ? wa4[j] : 0.0;
        }
        double r[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
        for (int j = 0; j < 2; j++)
            for (int i = 0; i <= j; i++) r[i][j] = (i < m) ? col[j][i] : 0.0;
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
            fit_resid(lx, ly, m, wa2[0], wa2[1], wa4);
            nfev++;
            const double fnorm1 = fit_enorm_big(wa4, m, sh);
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
                for (int j = 0; j < 2; j++) { x[j] = wa2[j]; wa2[j] = FMUL(diag[j], x[j]); }
                { double* t = fvec; fvec = wa4; wa4 = t; }  // fvec <- f(x_new): swap the two work vectors
                xnorm = fit_enorm_small(2, wa2);
                fnorm = fnorm1;
                iter++;
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
