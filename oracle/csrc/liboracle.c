/* Entry points of the oracle's C helpers (TEST INFRASTRUCTURE; see oracle/README.md). */
#include "svml_log10.h"
void ofg_oracle_log10(const double *x, double *y, long n)
{
    for (long i = 0; i < n; i++) y[i] = svml_log10_emul(x[i]);
}
