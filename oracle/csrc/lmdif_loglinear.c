/* Sequential restatement of what scipy.optimize.curve_fit does at gathering.py:120
 * (scnorm.CalculateFittingParameters): curve_fit(f, x, y) with f(x,a,b) = a*log10(x) + b, no p0,
 * no bounds -> method 'lm' -> scipy.optimize.leastsq(Dfun=None) -> MINPACK lmdif with
 * ftol = xtol = 1.49012e-8, gtol = 0, epsfcn = DBL_EPSILON, factor = 100, maxfev = 200*(n+1),
 * start (1, 1)  (scipy/optimize/_minpack_py.py: curve_fit, leastsq).
 *
 * TEST INFRASTRUCTURE (oracle).  MINPACK is a third-party dependency of the reference (scipy,
 * unpinned in requirements.txt; 1.18.1 in the build image); its published algorithm (lmdif,
 * fdjac2, qrfac, lmpar, qrsolv, enorm — More, Garbow, Hillstrom 1980) is restated here for
 * n = 2 and pinned bit-for-bit against scipy by tests/test_oracle_lmfit.py.  Every sum is a
 * plain left-to-right loop, no FMA contraction (compile with -ffp-contract=off): the result of
 * the fit depends on that order (SURVEY.md finding 5 and Appendix A).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

#define N 2
static const double EPSMCH = DBL_EPSILON;       /* dpmpar(1) */
static const double DWARF = DBL_MIN;            /* dpmpar(2) */

static double enorm(long n, const double *x)
{
    const double rdwarf = 3.834e-20, rgiant = 1.304e19;
    double s1 = 0, s2 = 0, s3 = 0, x1max = 0, x3max = 0;
    double agiant = rgiant / (double)n;
    for (long i = 0; i < n; i++) {
        double xabs = fabs(x[i]);
        if (xabs > rdwarf && xabs < agiant) { s2 += xabs * xabs; continue; }
        if (xabs <= rdwarf) {
            if (xabs > x3max) { double t = x3max / xabs; s3 = 1.0 + s3 * (t * t); x3max = xabs; }
            else if (xabs != 0) { double t = xabs / x3max; s3 += t * t; }
        } else {
            if (xabs > x1max) { double t = x1max / xabs; s1 = 1.0 + s1 * (t * t); x1max = xabs; }
            else { double t = xabs / x1max; s1 += t * t; }
        }
    }
    if (s1 != 0) return x1max * sqrt(s1 + (s2 / x1max) / x1max);
    if (s2 != 0) {
        if (s2 >= x3max) return sqrt(s2 * (1.0 + (x3max / s2) * (x3max * s3)));
        return sqrt(x3max * ((s2 / x3max) + (x3max * s3)));
    }
    return x3max * sqrt(s3);
}

/* residual of f(x,a,b) - y with f = a*log10(x) + b : three separately rounded operations */
static void resid(const double *lx, const double *ly, long m, const double *p, double *out)
{
    double a = p[0], b = p[1];
    for (long k = 0; k < m; k++) out[k] = ((a * lx[k]) + b) - ly[k];
}

/* r is n x n upper triangular stored r[i][j]; lower part used as scratch like MINPACK */
static void qrsolv(double r[N][N], const int *ipvt, const double *diag, const double *qtb,
                   double *x, double *sdiag)
{
    double wa[N];
    for (int j = 0; j < N; j++) {
        for (int i = j; i < N; i++) r[i][j] = r[j][i];
        x[j] = r[j][j];
        wa[j] = qtb[j];
    }
    for (int j = 0; j < N; j++) {
        int l = ipvt[j];
        if (diag[l] != 0) {
            for (int k = j; k < N; k++) sdiag[k] = 0;
            sdiag[j] = diag[l];
            double qtbpj = 0;
            for (int k = j; k < N; k++) {
                if (sdiag[k] == 0) continue;
                double c, s;
                if (fabs(r[k][k]) < fabs(sdiag[k])) {
                    double cotan = r[k][k] / sdiag[k];
                    s = 0.5 / sqrt(0.25 + 0.25 * (cotan * cotan));
                    c = s * cotan;
                } else {
                    double t = sdiag[k] / r[k][k];
                    c = 0.5 / sqrt(0.25 + 0.25 * (t * t));
                    s = c * t;
                }
                r[k][k] = c * r[k][k] + s * sdiag[k];
                double temp = c * wa[k] + s * qtbpj;
                qtbpj = -s * wa[k] + c * qtbpj;
                wa[k] = temp;
                for (int i = k + 1; i < N; i++) {
                    temp = c * r[i][k] + s * sdiag[i];
                    sdiag[i] = -s * r[i][k] + c * sdiag[i];
                    r[i][k] = temp;
                }
            }
        }
        sdiag[j] = r[j][j];
        r[j][j] = x[j];
    }
    int nsing = N;
    for (int j = 0; j < N; j++) {
        if (sdiag[j] == 0 && nsing == N) nsing = j;
        if (nsing < N) wa[j] = 0;
    }
    for (int k = 1; k <= nsing; k++) {
        int j = nsing - k;
        double sum = 0;
        for (int i = j + 1; i < nsing; i++) sum += r[i][j] * wa[i];
        wa[j] = (wa[j] - sum) / sdiag[j];
    }
    for (int j = 0; j < N; j++) x[ipvt[j]] = wa[j];
}

static void lmpar(double r[N][N], const int *ipvt, const double *diag, const double *qtb,
                  double delta, double *par, double *x, double *sdiag, int *iterative)
{
    double wa1[N], wa2[N];
    int nsing = N;
    for (int j = 0; j < N; j++) {
        wa1[j] = qtb[j];
        if (r[j][j] == 0 && nsing == N) nsing = j;
        if (nsing < N) wa1[j] = 0;
    }
    for (int k = 1; k <= nsing; k++) {
        int j = nsing - k;
        wa1[j] = wa1[j] / r[j][j];
        double temp = wa1[j];
        for (int i = 0; i < j; i++) wa1[i] -= r[i][j] * temp;
    }
    for (int j = 0; j < N; j++) x[ipvt[j]] = wa1[j];
    int iter = 0;
    for (int j = 0; j < N; j++) wa2[j] = diag[j] * x[j];
    double dxnorm = enorm(N, wa2);
    double fp = dxnorm - delta;
    if (fp <= 0.1 * delta) { *par = 0; return; }
    *iterative = 1;
    double parl = 0;
    if (nsing >= N) {
        for (int j = 0; j < N; j++) { int l = ipvt[j]; wa1[j] = diag[l] * (wa2[l] / dxnorm); }
        for (int j = 0; j < N; j++) {
            double sum = 0;
            for (int i = 0; i < j; i++) sum += r[i][j] * wa1[i];
            wa1[j] = (wa1[j] - sum) / r[j][j];
        }
        double temp = enorm(N, wa1);
        parl = ((fp / delta) / temp) / temp;
    }
    for (int j = 0; j < N; j++) {
        double sum = 0;
        for (int i = 0; i <= j; i++) sum += r[i][j] * qtb[i];
        wa1[j] = sum / diag[ipvt[j]];
    }
    double gnorm = enorm(N, wa1);
    double paru = gnorm / delta;
    if (paru == 0) paru = DWARF / fmin(delta, 0.1);
    *par = fmax(*par, parl);
    *par = fmin(*par, paru);
    if (*par == 0) *par = gnorm / dxnorm;
    for (;;) {
        iter++;
        if (*par == 0) *par = fmax(DWARF, 0.001 * paru);
        double temp = sqrt(*par);
        for (int j = 0; j < N; j++) wa1[j] = temp * diag[j];
        qrsolv(r, ipvt, wa1, qtb, x, sdiag);
        for (int j = 0; j < N; j++) wa2[j] = diag[j] * x[j];
        dxnorm = enorm(N, wa2);
        temp = fp;
        fp = dxnorm - delta;
        if (fabs(fp) <= 0.1 * delta || (parl == 0 && fp <= temp && temp < 0) || iter == 10) break;
        for (int j = 0; j < N; j++) { int l = ipvt[j]; wa1[j] = diag[l] * (wa2[l] / dxnorm); }
        for (int j = 0; j < N; j++) {
            wa1[j] = wa1[j] / sdiag[j];
            temp = wa1[j];
            for (int i = j + 1; i < N; i++) wa1[i] -= r[i][j] * temp;
        }
        temp = enorm(N, wa1);
        double parc = ((fp / delta) / temp) / temp;
        if (fp > 0) parl = fmax(parl, *par);
        if (fp < 0) paru = fmin(paru, *par);
        *par = fmax(parl, *par + parc);
    }
    if (iter == 0) *par = 0;
}

/* Returns MINPACK info (1..8), or -1 on allocation failure.  out[0]=a, out[1]=b.
 * stats[0]=nfev, stats[1]=1 if lmpar's iterative branch ran, stats[2]=outer iterations. */
int ofg_oracle_lmfit(const double *lx, const double *ly, long m, double *out, int *stats)
{
    const double ftol = 1.49012e-08, xtol = 1.49012e-08, gtol = 0.0, factor = 100.0;
    const int maxfev = 200 * (N + 1);
    double *fvec = malloc(sizeof(double) * m), *wa4 = malloc(sizeof(double) * m);
    double *col[N];
    col[0] = malloc(sizeof(double) * m);
    col[1] = malloc(sizeof(double) * m);
    if (!fvec || !wa4 || !col[0] || !col[1]) return -1;
    double x[N] = {1.0, 1.0}, diag[N], qtf[N], wa1[N], wa2[N], wa3[N];
    double acnorm[N], rdiag[N], wa[N];
    int ipvt[N];
    int info = 0, nfev = 0, iterative = 0, iter = 1;
    double par = 0, delta = 0, xnorm = 0, gnorm = 0;
    double eps = sqrt(fmax(EPSMCH, EPSMCH)); /* epsfcn = eps */

    resid(lx, ly, m, x, fvec);
    nfev = 1;
    double fnorm = enorm(m, fvec);
    for (;;) {
        /* fdjac2 */
        for (int j = 0; j < N; j++) {
            double temp = x[j];
            double h = eps * fabs(temp);
            if (h == 0) h = eps;
            x[j] = temp + h;
            resid(lx, ly, m, x, wa4);
            x[j] = temp;
            for (long i = 0; i < m; i++) col[j][i] = (wa4[i] - fvec[i]) / h;
        }
        nfev += N;
        /* qrfac with pivoting */
        for (int j = 0; j < N; j++) {
            acnorm[j] = enorm(m, col[j]);
            rdiag[j] = acnorm[j];
            wa[j] = rdiag[j];
            ipvt[j] = j;
        }
        long minmn = m < N ? m : N;
        for (int j = 0; j < minmn; j++) {
            int kmax = j;
            for (int k = j; k < N; k++) if (rdiag[k] > rdiag[kmax]) kmax = k;
            if (kmax != j) {
                double *t = col[j]; col[j] = col[kmax]; col[kmax] = t;
                rdiag[kmax] = rdiag[j];
                wa[kmax] = wa[j];
                int k = ipvt[j]; ipvt[j] = ipvt[kmax]; ipvt[kmax] = k;
            }
            double ajnorm = enorm(m - j, col[j] + j);
            if (ajnorm != 0) {
                if (col[j][j] < 0) ajnorm = -ajnorm;
                for (long i = j; i < m; i++) col[j][i] /= ajnorm;
                col[j][j] += 1.0;
                for (int k = j + 1; k < N; k++) {
                    double sum = 0;
                    for (long i = j; i < m; i++) sum += col[j][i] * col[k][i];
                    double temp = sum / col[j][j];
                    for (long i = j; i < m; i++) col[k][i] -= temp * col[j][i];
                    if (rdiag[k] != 0) {
                        temp = col[k][j] / rdiag[k];
                        rdiag[k] *= sqrt(fmax(0.0, 1.0 - temp * temp));
                        double q = rdiag[k] / wa[k];
                        if (0.05 * (q * q) <= EPSMCH) {
                            rdiag[k] = enorm(m - j - 1, col[k] + j + 1);
                            wa[k] = rdiag[k];
                        }
                    }
                }
            }
            rdiag[j] = -ajnorm;
        }
        if (iter == 1) {
            for (int j = 0; j < N; j++) { diag[j] = acnorm[j]; if (acnorm[j] == 0) diag[j] = 1.0; }
            for (int j = 0; j < N; j++) wa3[j] = diag[j] * x[j];
            xnorm = enorm(N, wa3);
            delta = factor * xnorm;
            if (delta == 0) delta = factor;
        }
        /* qtf */
        memcpy(wa4, fvec, sizeof(double) * m);
        for (int j = 0; j < N; j++) {
            if (j < m && col[j][j] != 0) {
                double sum = 0;
                for (long i = j; i < m; i++) sum += col[j][i] * wa4[i];
                double temp = -sum / col[j][j];
                for (long i = j; i < m; i++) wa4[i] += col[j][i] * temp;
            }
            col[j][j] = rdiag[j];
            qtf[j] = wa4[j];
        }
        double r[N][N];
        memset(r, 0, sizeof r);
        for (int j = 0; j < N; j++) for (int i = 0; i <= j; i++) r[i][j] = col[j][i];
        gnorm = 0;
        if (fnorm != 0) {
            for (int j = 0; j < N; j++) {
                int l = ipvt[j];
                if (acnorm[l] == 0) continue;
                double sum = 0;
                for (int i = 0; i <= j; i++) sum += r[i][j] * (qtf[i] / fnorm);
                gnorm = fmax(gnorm, fabs(sum / acnorm[l]));
            }
        }
        if (gnorm <= gtol) { info = 4; break; }
        for (int j = 0; j < N; j++) diag[j] = fmax(diag[j], acnorm[j]);
        double ratio;
        do {
            double sdiag[N], rr[N][N];
            memcpy(rr, r, sizeof r);
            lmpar(rr, ipvt, diag, qtf, delta, &par, wa1, sdiag, &iterative);
            for (int j = 0; j < N; j++) {
                wa1[j] = -wa1[j];
                wa2[j] = x[j] + wa1[j];
                wa3[j] = diag[j] * wa1[j];
            }
            double pnorm = enorm(N, wa3);
            if (iter == 1) delta = fmin(delta, pnorm);
            resid(lx, ly, m, wa2, wa4);
            nfev++;
            double fnorm1 = enorm(m, wa4);
            double actred = -1.0;
            if (0.1 * fnorm1 < fnorm) { double q = fnorm1 / fnorm; actred = 1.0 - q * q; }
            for (int j = 0; j < N; j++) {
                wa3[j] = 0;
                double temp = wa1[ipvt[j]];
                for (int i = 0; i <= j; i++) wa3[i] += r[i][j] * temp;
            }
            double temp1 = enorm(N, wa3) / fnorm;
            double temp2 = (sqrt(par) * pnorm) / fnorm;
            double prered = temp1 * temp1 + (temp2 * temp2) / 0.5;
            double dirder = -(temp1 * temp1 + temp2 * temp2);
            ratio = 0;
            if (prered != 0) ratio = actred / prered;
            if (ratio <= 0.25) {
                double temp = 0.5;
                if (actred < 0) temp = 0.5 * dirder / (dirder + 0.5 * actred);
                if (0.1 * fnorm1 >= fnorm || temp < 0.1) temp = 0.1;
                delta = temp * fmin(delta, pnorm / 0.1);
                par = par / temp;
            } else if (par == 0 || ratio >= 0.75) {
                delta = pnorm / 0.5;
                par = 0.5 * par;
            }
            if (ratio >= 1e-4) {
                for (int j = 0; j < N; j++) { x[j] = wa2[j]; wa2[j] = diag[j] * x[j]; }
                memcpy(fvec, wa4, sizeof(double) * m);
                xnorm = enorm(N, wa2);
                fnorm = fnorm1;
                iter++;
            }
            if (fabs(actred) <= ftol && prered <= ftol && 0.5 * ratio <= 1.0) info = 1;
            if (delta <= xtol * xnorm) info = 2;
            if (fabs(actred) <= ftol && prered <= ftol && 0.5 * ratio <= 1.0 && info == 2) info = 3;
            if (info != 0) break;
            if (nfev >= maxfev) info = 5;
            if (fabs(actred) <= EPSMCH && prered <= EPSMCH && 0.5 * ratio <= 1.0) info = 6;
            if (delta <= EPSMCH * xnorm) info = 7;
            if (gnorm <= EPSMCH) info = 8;
            if (info != 0) break;
        } while (ratio < 1e-4);
        if (info != 0) break;
    }
    out[0] = x[0];
    out[1] = x[1];
    if (stats) { stats[0] = nfev; stats[1] = iterative; stats[2] = iter; }
    free(fvec); free(wa4); free(col[0]); free(col[1]);
    return info;
}
