// CPU build of the host/device-shared pieces of libofg (fit core, record parser, synthetic
// generator) so their control flow can be checked without a GPU.  Test helper only: the product
// never loads this library.
#include <stdlib.h>
#include <string.h>
#include "ofg_fit_core.h"
#include "ofg_parse_line.h"

struct HostSrc {
    const unsigned char* s;
    int64_t limit;
    int get(int64_t pos) { return pos >= limit ? '\n' : s[pos]; }
};

extern "C" {

int ofg_hostcheck_fit(const double* lx, const double* ly, long m, double* out)
{
    static FitShared sh;
    int info = fit_loglinear(lx, ly, m, &sh, out);
    return info;
}

// parses every non-empty record of text; returns count or -(offset+1) of the first bad line
long ofg_hostcheck_parse(const char* text, long n, unsigned* f0, unsigned* f1, double* sc, int* kind)
{
    HostSrc src{(const unsigned char*)text, n};
    long cnt = 0, pos = 0;
    while (pos < n) {
        long ls = pos;
        while (pos < n && !ofg_is_term(text[pos])) pos++;
        long le = pos;
        if (pos < n) pos++;
        if (le == ls) continue;
        int rc = ofg_parse_record(src, ls, &f0[cnt], &f1[cnt], &sc[cnt]);
        if (rc) { *kind = rc; return -(ls + 1); }
        cnt++;
    }
    return cnt;
}

int ofg_hostcheck_score(const char* text, long n, double* out)
{
    HostSrc src{(const unsigned char*)text, n};
    return ofg_parse_score_field(src, 0, out);
}
}
