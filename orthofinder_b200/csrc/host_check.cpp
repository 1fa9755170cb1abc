// CPU build of the host/device-shared pieces of libofg (fit core, record parser, synthetic
// generator) so their control flow can be checked without a GPU.  Test helper only: the product
// never loads this library.
#include <stdio.h>
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

// ---------------------------------------------------------------- MCL on the device, walked on the host
// The engine of ofg_mcl_driver.h over a backend that runs every "kernel" as a loop: per-column functions phase by phase
// (ofg_mcl_core.h), scans sequentially.  Same batching, offsets and buffer hand-over as the CUDA backend of ofg_mcl.cu.
#include "ofg_mcl_driver.h"

namespace {
struct MclHostBackend {
    std::string e;
    const char* error() { return e.c_str(); }
    bool alloc(void** p, size_t bytes) { *p = malloc(bytes ? bytes : 16); if (!*p) e = "malloc"; return *p != nullptr; }
    void free(void* p) { ::free(p); }
    bool memset(void* p, int v, size_t bytes) { ::memset(p, v, bytes); return true; }
    bool h2d(void* d, const void* h, size_t bytes) { memcpy(d, h, bytes); return true; }
    bool d2h(void* h, const void* d, size_t bytes) { memcpy(h, d, bytes); return true; }
    bool d2d(void* d, const void* s, size_t bytes) { memcpy(d, s, bytes); return true; }
    bool start_count(const MclStartArgs& a) { for (int64_t c = 0; c < a.n; c++) mcl_start_count(a, c); return true; }
    bool start_fill(const MclStartArgs& a) { for (int64_t c = 0; c < a.n; c++) mcl_start_fill(a, c); return true; }
    bool plan(const MclPlanArgs& a) { for (int64_t c = 0; c < a.n; c++) mcl_plan_column(a, c); return true; }
    template <typename T> bool scan(const T* in, int64_t* out, int64_t count, int64_t base)
    {
        int64_t r = base;
        for (int64_t i = 0; i < count; i++) { const int64_t x = (int64_t)in[i]; out[i] = r; r += x; }
        out[count] = r;
        return true;
    }
    bool scan_i64(const int64_t* in, int64_t* out, int64_t count, int64_t base) { return scan(in, out, count, base); }
    bool scan_i32(const int32_t* in, int64_t* out, int64_t count, int64_t base) { return scan(in, out, count, base); }
    bool expand(const MclExpandArgs& a, int64_t ncols)
    {
        static MclColShared sh;
        static MclColSharedLite lite;
        for (int64_t b = 0; b < ncols; b++) mcl_expand_column_hashed(sh, a, a.c0 + b);
        for (int64_t b = 0; b < ncols; b++) mcl_expand_column_direct(lite, a, a.c0 + b);
        return true;
    }
    bool inflate(const MclInflateArgs& a, int64_t ncols)
    {
        static MclRedShared sh;
        for (int64_t b = 0; b < ncols; b++) mcl_inflate_column(sh, a, a.c0 + b);
        return true;
    }
    bool dag(const MclDagArgs& a) { for (int64_t c = 0; c < a.n; c++) mcl_dag_column(a, c); return true; }
    bool attractor_round(const MclDagArgs& a) { for (int64_t c = 0; c < a.n; c++) mcl_attractor_round(a, c); return true; }
    bool reach_round(const MclDagArgs& a) { for (int64_t c = 0; c < a.n; c++) mcl_reach_round(a, c); return true; }
};
typedef MclEngine<MclHostBackend> HostMcl;
}  // namespace

extern "C" {

void* ofg_hostcheck_mcl_new() { return new HostMcl(); }
void ofg_hostcheck_mcl_free(void* h) { delete (HostMcl*)h; }
int ofg_hostcheck_mcl_params(void* h, int P, int S, int R, double pct, int overhead, long table_budget, long stage_budget, int direct_factor)
{
    HostMcl* m = (HostMcl*)h;
    m->par.P = P; m->par.S = S; m->par.R = R; m->par.pct = pct; m->par.overhead = overhead;
    if (table_budget > 0) m->par.table_budget = table_budget;
    if (stage_budget > 0) m->par.stage_budget = stage_budget;
    if (direct_factor >= 0) m->par.direct_factor = direct_factor;
    return 0;
}
int ofg_hostcheck_mcl_start(void* h, long n, const int64_t* cp, const int32_t* ri, const float* v) { return ((HostMcl*)h)->start(n, cp, ri, v) ? 0 : -1; }
int ofg_hostcheck_mcl_step(void* h, double inflation, double* chaos) { return ((HostMcl*)h)->step(inflation, chaos) ? 0 : -1; }
long ofg_hostcheck_mcl_nnz(void* h) { return (long)((HostMcl*)h)->cur.nnz; }
int ofg_hostcheck_mcl_get(void* h, int64_t* cp, int32_t* ri, float* v) { return ((HostMcl*)h)->copy_iterand(cp, ri, v) ? 0 : -1; }
int ofg_hostcheck_mcl_interpret(void* h, int32_t* labels, long* n_clusters, long* n_overlap)
{
    HostMcl* m = (HostMcl*)h;
    if (!m->interpret()) return -1;
    memcpy(labels, m->labels.data(), (size_t)m->n * 4);
    *n_clusters = (long)m->n_clusters; *n_overlap = (long)m->n_overlap;
    return 0;
}
long ofg_hostcheck_mcl_text(void* h, char* buf, long cap)
{
    HostMcl* m = (HostMcl*)h;
    const std::string t = mcl_clusters_text(m->labels, m->n_clusters);
    if (buf && (long)t.size() <= cap) memcpy(buf, t.data(), t.size());
    return (long)t.size();
}
}

extern "C" int ofg_hostcheck_mcl_weight(double w, float* out) { return mcl_weight_as_read(w, out) ? 1 : 0; }
// parses a graph file's text like ofg_mcl_set_graph_text; arrays are returned through the caller's buffers when they are large enough
extern "C" long ofg_hostcheck_mcl_parse(const char* text, long nbytes, long* n, int64_t* cp, long cp_cap, int32_t* ri, float* v, long e_cap, char* err, long err_cap)
{
    std::vector<int64_t> c; std::vector<int32_t> r; std::vector<float> w;
    int64_t nn = 0;
    const std::string msg = mcl_parse_graph_text(text, (size_t)nbytes, &nn, c, r, w);
    if (!msg.empty()) { snprintf(err, (size_t)err_cap, "%s", msg.c_str()); return -1; }
    *n = (long)nn;
    if ((long)c.size() <= cp_cap && (long)r.size() <= e_cap) {
        memcpy(cp, c.data(), c.size() * 8);
        if (!r.empty()) { memcpy(ri, r.data(), r.size() * 4); memcpy(v, w.data(), w.size() * 4); }
    }
    return (long)r.size();
}
