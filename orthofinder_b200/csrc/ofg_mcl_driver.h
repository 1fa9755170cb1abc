// MCL on the device: the iteration loop, memory plan and interpretation, written once over a small backend interface.
// ofg_mcl.cu instantiates it with the CUDA backend (kernels of ofg_mcl_core.h); host_check.cpp instantiates it with a
// host backend that runs the same per-column functions phase by phase, so the batching, the offsets and every buffer
// hand-over are exercised on the CPU as well (tests/test_mcl_host_core.py).
//
// One iteration (reference: the expansion / inflation loop inside the `mcl` binary that RunMCL starts, mcl.py:214-218):
//   plan      per column: summands -> hash capacity, dense flag, stage size            (thread per column)
//   scans     table / stage offsets; read back for the batching
//   batches   of consecutive columns whose global tables and stage fit the budgets:
//             expand (CTA per column) -> scan of the kept counts -> inflate + pack into the next iterand (CTA per column)
//   chaos     max over columns, read back; the loop ends when it is below 1e-4
// Interpretation: dag + attractor flags (thread per column), label rounds until nothing changes, cluster order on the host.
#pragma once
#include <stdint.h>
#include <stdlib.h>
#include <algorithm>
#include <string>
#include <vector>

#include "ofg_mcl_core.h"

struct MclParams {
    int P = 10000, S = 1100, R = 1400;     // resource scheme 6 of mcl 14-137 (the binary's default)
    double pct = 0.9;
    int overhead = 10;                     // -sparse
    int max_iter = 10000;                  // -L
    double chaos_limit = 1e-4;
    int64_t table_budget = (int64_t)1 << 28;   // slots of global hash tables / row-indexed accumulators per batch (12 bytes each)
    int64_t stage_budget = (int64_t)1 << 27;   // stage entries per batch (8 bytes each)
    int direct_factor = MCL_DIRECT_FACTOR;     // see MclPlanArgs
};

template <class B> class MclEngine {
public:
    B be;
    MclParams par;
    std::string err;
    int64_t n = 0;
    int64_t iterations = 0;
    double last_chaos = 0.0;
    int64_t n_clusters = 0, n_overlap = 0;
    int64_t launches = 0;
    std::vector<int32_t> labels;           // node -> cluster index in output order (after interpret)
    std::vector<double> chaos_log;

    struct Mx {
        int64_t* cp = nullptr;
        int32_t* ri = nullptr;
        float* v = nullptr;
        int64_t cap = 0, nnz = 0;
    };
    Mx cur, nxt;

    ~MclEngine() { release(); }

    void release()
    {
        free_mx(cur); free_mx(nxt);
        void** all[] = {(void**)&tcap, (void**)&dense, (void**)&need, (void**)&stg, (void**)&table_off, (void**)&stage_off, (void**)&cnt,
                        (void**)&gkeys, (void**)&gacc, (void**)&stage, (void**)&chaos_bits, (void**)&keep, (void**)&attr, (void**)&sys,
                        (void**)&second, (void**)&changed};
        for (void** p : all) { if (*p) be.free(*p); *p = nullptr; }
        table_cap = stage_cap = keep_cap = 0;
        n = 0;
    }

    // graph as the binary reads it from the file: column c = line c, rows ascending (device arrays of the backend)
    bool start(int64_t n_, const int64_t* in_cp, const int32_t* in_ri, const float* in_v)
    {
        if (!alloc_columns(n_)) return false;
        MclStartArgs a{};
        a.n = n; a.in_cp = in_cp; a.in_ri = in_ri; a.in_v = in_v; a.cnt = cnt;
        if (!be.start_count(a)) return fail();
        launches++;
        if (!be.scan_i32(cnt, cur.cp, n, 0)) return fail();
        launches++;
        int64_t tot = 0;
        if (!be.d2h(&tot, cur.cp + n, 8)) return fail();
        if (!reserve(cur, tot, false)) return false;
        a.out_cp = cur.cp; a.out_ri = cur.ri; a.out_v = cur.v;
        if (!be.start_fill(a)) return fail();
        launches++;
        cur.nnz = tot;
        iterations = 0;
        chaos_log.clear();
        labels.clear();
        return true;
    }

    bool step(double inflation, double* chaos_out)
    {
        MclPlanArgs pl{};
        pl.n = n; pl.cp = cur.cp; pl.ri = cur.ri; pl.overhead = par.overhead; pl.direct_factor = par.direct_factor;
        pl.tcap = tcap; pl.dense = dense; pl.need = need; pl.stg = stg;
        if (!be.plan(pl)) return fail();
        if (!be.scan_i64(need, table_off, n, 0) || !be.scan_i64(stg, stage_off, n, 0)) return fail();
        launches += 3;
        h_table.resize((size_t)n + 1); h_stage.resize((size_t)n + 1);
        if (!be.d2h(h_table.data(), table_off, (size_t)(n + 1) * 8) || !be.d2h(h_stage.data(), stage_off, (size_t)(n + 1) * 8)) return fail();
        if (!be.memset(chaos_bits, 0, 16)) return fail();
        int64_t base = 0;
        for (int64_t c0 = 0; c0 < n;) {
            int64_t c1 = c0 + 1;
            while (c1 < n && h_table[(size_t)c1 + 1] - h_table[(size_t)c0] <= par.table_budget &&
                   h_stage[(size_t)c1 + 1] - h_stage[(size_t)c0] <= par.stage_budget && c1 - c0 < ((int64_t)1 << 30))
                c1++;
            const int64_t T = h_table[(size_t)c1] - h_table[(size_t)c0], SG = h_stage[(size_t)c1] - h_stage[(size_t)c0];
            if (T > table_cap) {
                if (gkeys) be.free(gkeys);
                if (gacc) be.free(gacc);
                gkeys = nullptr; gacc = nullptr; table_cap = 0;
                if (!be.alloc((void**)&gkeys, (size_t)T * 4) || !be.alloc((void**)&gacc, (size_t)T * 8)) return fail();
                table_cap = T;
            }
            if (SG > stage_cap) {
                if (stage) be.free(stage);
                stage = nullptr; stage_cap = 0;
                if (!be.alloc((void**)&stage, (size_t)SG * 8)) return fail();
                stage_cap = SG;
            }
            if (T && !be.memset(gkeys, 0xFF, (size_t)T * 4)) return fail();
            MclExpandArgs e{};
            e.n = n; e.cp = cur.cp; e.ri = cur.ri; e.v = cur.v; e.c0 = c0;
            e.table_off = table_off; e.stage_off = stage_off; e.tcap = tcap; e.dense = dense;
            e.gkeys = gkeys; e.gacc = gacc; e.stage = stage; e.cnt = cnt; e.chaos_bits = chaos_bits; e.err = (int32_t*)(chaos_bits + 1);
            e.cut = par.P ? 0.99999 / (double)par.P : 0.0; e.S = par.S; e.R = par.R; e.pct = par.pct;
            if (!be.expand(e, c1 - c0)) return fail();
            if (!be.scan_i32(cnt + c0, nxt.cp + c0, c1 - c0, base)) return fail();
            launches += 3;
            int64_t nb = 0;
            if (!be.d2h(&nb, nxt.cp + c1, 8)) return fail();
            if (!reserve(nxt, nb, true, base)) return false;
            MclInflateArgs f{};
            f.c0 = c0; f.stage_off = stage_off; f.stage = stage; f.cnt = cnt; f.new_cp = nxt.cp; f.out_ri = nxt.ri; f.out_v = nxt.v;
            f.power = (double)(float)inflation;
            if (!be.inflate(f, c1 - c0)) return fail();
            launches++;
            base = nb;
            c0 = c1;
        }
        nxt.nnz = base;
        std::swap(cur, nxt);
        mcl_u64 back[2] = {0, 0};
        if (!be.d2h(back, chaos_bits, 16)) return fail();
        if (back[1]) { err = "hash table of a column overflowed (inconsistent plan)"; return false; }
        double ch;
        memcpy(&ch, &back[0], 8);
        last_chaos = ch;
        chaos_log.push_back(ch);
        iterations++;
        if (chaos_out) *chaos_out = ch;
        return true;
    }

    bool run(double inflation)
    {
        while (iterations < par.max_iter) {
            double ch = 0.0;
            if (!step(inflation, &ch)) return false;
            if (ch < par.chaos_limit) break;
        }
        return interpret();
    }

    // clusters of the current iterand (see ofg_mcl_core.h); fills labels, n_clusters, n_overlap
    bool interpret()
    {
        if (cur.nnz > keep_cap) {
            if (keep) be.free(keep);
            keep = nullptr; keep_cap = 0;
            if (!be.alloc((void**)&keep, (size_t)cur.nnz + 16)) return fail();
            keep_cap = cur.nnz;
        }
        MclDagArgs d{};
        d.n = n; d.cp = cur.cp; d.ri = cur.ri; d.v = cur.v; d.keep = keep; d.attr = attr; d.sys = sys; d.second = second; d.changed = changed;
        if (!be.dag(d)) return fail();
        launches++;
        for (int phase = 0; phase < 2; phase++) {
            for (int64_t round = 0; round <= n + 1; round++) {
                if (!be.memset(changed, 0, 4)) return fail();
                if (!(phase == 0 ? be.attractor_round(d) : be.reach_round(d))) return fail();
                launches++;
                int32_t ch = 0;
                if (!be.d2h(&ch, changed, 4)) return fail();
                if (!ch) break;
            }
        }
        std::vector<int32_t> hs((size_t)n), h2((size_t)n);
        if (n && (!be.d2h(hs.data(), sys, (size_t)n * 4) || !be.d2h(h2.data(), second, (size_t)n * 4))) return fail();
        n_overlap = 0;
        int32_t garbage = -1;
        for (int64_t i = 0; i < n; i++) {
            if (h2[(size_t)i] != MCL_NONE) n_overlap++;
            if (hs[(size_t)i] == MCL_NONE) {     // reaches no attractor system: the binary puts all such nodes into ONE cluster
                if (garbage < 0) garbage = (int32_t)i;   // ("added <k> garbage entries"); id = its smallest member, never an attractor
                hs[(size_t)i] = garbage;
            }
        }
        // output order of the binary: size descending, then smallest member
        std::vector<int32_t> size((size_t)n, 0), first((size_t)n, -1);
        for (int64_t i = 0; i < n; i++) {
            const int32_t s = hs[(size_t)i];
            size[(size_t)s]++;
            if (first[(size_t)s] < 0) first[(size_t)s] = (int32_t)i;
        }
        std::vector<int32_t> order;
        for (int64_t s = 0; s < n; s++) if (size[(size_t)s]) order.push_back((int32_t)s);
        std::sort(order.begin(), order.end(), [&](int32_t x, int32_t y) {
            if (size[(size_t)x] != size[(size_t)y]) return size[(size_t)x] > size[(size_t)y];
            return first[(size_t)x] < first[(size_t)y];
        });
        std::vector<int32_t> rank((size_t)n, -1);
        for (size_t i = 0; i < order.size(); i++) rank[(size_t)order[i]] = (int32_t)i;
        labels.resize((size_t)n);
        for (int64_t i = 0; i < n; i++) labels[(size_t)i] = rank[(size_t)hs[(size_t)i]];
        n_clusters = (int64_t)order.size();
        return true;
    }

    // current iterand to host arrays (parity tests)
    bool copy_iterand(int64_t* cp, int32_t* ri, float* v)
    {
        if (cp && !be.d2h(cp, cur.cp, (size_t)(n + 1) * 8)) return fail();
        if (ri && cur.nnz && !be.d2h(ri, cur.ri, (size_t)cur.nnz * 4)) return fail();
        if (v && cur.nnz && !be.d2h(v, cur.v, (size_t)cur.nnz * 4)) return fail();
        return true;
    }

private:
    int32_t* tcap = nullptr;
    uint8_t* dense = nullptr;
    int64_t *need = nullptr, *stg = nullptr, *table_off = nullptr, *stage_off = nullptr;
    int32_t* cnt = nullptr;
    int32_t* gkeys = nullptr;
    double* gacc = nullptr;
    mcl_u64* stage = nullptr;
    mcl_u64* chaos_bits = nullptr;
    uint8_t *keep = nullptr, *attr = nullptr;
    int32_t *sys = nullptr, *second = nullptr, *changed = nullptr;
    int64_t table_cap = 0, stage_cap = 0, keep_cap = 0;
    std::vector<int64_t> h_table, h_stage;

    bool fail()
    {
        err = be.error();
        if (err.empty()) err = "backend failure";
        return false;
    }
    void free_mx(Mx& m)
    {
        if (m.cp) be.free(m.cp);
        if (m.ri) be.free(m.ri);
        if (m.v) be.free(m.v);
        m = Mx();
    }
    bool alloc_columns(int64_t n_)
    {
        release();
        n = n_;
        const size_t c = (size_t)n + 1;
        bool ok = be.alloc((void**)&cur.cp, c * 8) && be.alloc((void**)&nxt.cp, c * 8) && be.alloc((void**)&tcap, c * 4) &&
                  be.alloc((void**)&dense, c) && be.alloc((void**)&need, c * 8) && be.alloc((void**)&stg, c * 8) &&
                  be.alloc((void**)&table_off, c * 8) && be.alloc((void**)&stage_off, c * 8) && be.alloc((void**)&cnt, c * 4) &&
                  be.alloc((void**)&chaos_bits, 16) && be.alloc((void**)&attr, c) && be.alloc((void**)&sys, c * 4) &&
                  be.alloc((void**)&second, c * 4) && be.alloc((void**)&changed, 4);
        return ok ? true : fail();
    }
    // capacity for `need_nnz` entries; keep_upto > 0 preserves the first entries already written
    bool reserve(Mx& m, int64_t need_nnz, bool preserve, int64_t keep_upto = 0)
    {
        if (need_nnz <= m.cap) return true;
        int64_t want = need_nnz + need_nnz / 4 + 1024;
        int32_t* nri = nullptr;
        float* nv = nullptr;
        if (!be.alloc((void**)&nri, (size_t)want * 4) || !be.alloc((void**)&nv, (size_t)want * 4)) return fail();
        if (preserve && keep_upto > 0 && m.ri) {
            if (!be.d2d(nri, m.ri, (size_t)keep_upto * 4) || !be.d2d(nv, m.v, (size_t)keep_upto * 4)) return fail();
        }
        if (m.ri) be.free(m.ri);
        if (m.v) be.free(m.v);
        m.ri = nri; m.v = nv; m.cap = want;
        return true;
    }
};

// the clusters file from `(mclheader` on, exactly as mcl 14-137 writes it (ids left-justified to the width w of n, one
// space before every member, a new line indented by w + 2 after a member that brings the line to 70 - w characters or
// more, the closing " $" never on a line of its own) - what GetPredictedOGs reads (scripts_of/mcl.py:36-63)
inline std::string mcl_clusters_text(const std::vector<int32_t>& labels, int64_t n_clusters)
{
    const int64_t n = (int64_t)labels.size();
    std::vector<int64_t> off((size_t)n_clusters + 1, 0);
    for (int64_t i = 0; i < n; i++) off[(size_t)labels[(size_t)i] + 1]++;
    for (int64_t k = 0; k < n_clusters; k++) off[(size_t)k + 1] += off[(size_t)k];
    std::vector<int32_t> members((size_t)n);
    {
        std::vector<int64_t> cur(off.begin(), off.end() - 1);
        for (int64_t i = 0; i < n; i++) members[(size_t)cur[(size_t)labels[(size_t)i]]++] = (int32_t)i;   // ascending within a cluster
    }
    const int w = (int)std::to_string((long long)n).size();
    std::string out = "(mclheader\nmcltype matrix\ndimensions " + std::to_string((long long)n) + "x" + std::to_string((long long)n_clusters) +
                      "\n)\n(mclmatrix\nbegin\n";
    out.reserve(out.size() + (size_t)n * 8 + (size_t)n_clusters * 8);
    for (int64_t k = 0; k < n_clusters; k++) {
        std::string line = std::to_string((long long)k);
        if ((int)line.size() < w) line.append((size_t)(w - (int)line.size()), ' ');
        for (int64_t i = off[(size_t)k]; i < off[(size_t)k + 1]; i++) {
            line += ' ';
            line += std::to_string((int)members[(size_t)i]);
            if ((int)line.size() + w >= 70 && i + 1 < off[(size_t)k + 1]) {
                out += line;
                out += '\n';
                line.assign((size_t)(w + 2), ' ');
            }
        }
        out += line;
        out += " $\n";
    }
    out += ")\n";
    return out;
}

// The graph file that WriteGraphParallel writes and the binary reads (MCL "matrix" interchange format: header up to
// `begin`, then `<column> <row>:<value> ... $` records, `)` at the end; gathering.py:39-48, :279-280) -> column-compressed
// arrays with the values the binary stores (strtod, then float).  Returns an empty string or the reason for refusing.
inline std::string mcl_parse_graph_text(const char* text, size_t nbytes, int64_t* n_out, std::vector<int64_t>& cp, std::vector<int32_t>& ri,
                                        std::vector<float>& v)
{
    const char* p = text;
    const char* end = text + nbytes;
    auto find = [&](const char* from, const char* word) -> const char* {
        const size_t L = strlen(word);
        for (const char* q = from; q + L <= end; q++) if (memcmp(q, word, L) == 0) return q;
        return nullptr;
    };
    const char* dim = find(p, "dimensions");
    if (!dim) return "no `dimensions` line";
    dim += 10;
    char* e = nullptr;
    std::string head(dim, (size_t)std::min<ptrdiff_t>(end - dim, 64));
    const long long nr = strtoll(head.c_str(), &e, 10);
    if (!e || *e != 'x') return "malformed `dimensions` line";
    const long long nc = strtoll(e + 1, &e, 10);
    if (nr < 1 || nr != nc || nr > 0x7ffffff0LL) return "the matrix must be square with 1 .. 2^31 - 16 nodes";
    const char* b = find(dim, "begin");
    if (!b) return "no `begin`";
    p = b + 5;
    const int64_t n = (int64_t)nr;
    std::vector<int64_t> count((size_t)n, 0);
    ri.clear(); v.clear();
    std::vector<int64_t> col_of;      // column of every entry, in file order (columns ascend, so entries are already grouped)
    int64_t cur = -1, last_col = -1;
    std::string tok;
    while (p < end) {
        while (p < end && (*p == ' ' || *p == '\n' || *p == '\r' || *p == '\t')) p++;
        if (p >= end) break;
        if (*p == ')') break;
        if (*p == '$') { cur = -1; p++; continue; }
        const char* q = p;
        while (q < end && *q != ' ' && *q != '\n' && *q != '\r' && *q != '\t' && *q != '$') q++;
        tok.assign(p, (size_t)(q - p));
        p = q;
        if (cur < 0) {                                   // column id
            const long long c = strtoll(tok.c_str(), &e, 10);
            if (*e != 0 || c <= last_col || c >= n) return "column ids must ascend and stay below the dimension";
            cur = last_col = (int64_t)c;
            continue;
        }
        const size_t colon = tok.find(':');
        const long long r = strtoll(tok.c_str(), &e, 10);
        if ((colon == std::string::npos ? *e != 0 : e != tok.c_str() + colon) || r < 0 || r >= n) return "malformed entry `" + tok + "`";
        double w = 1.0;
        if (colon != std::string::npos) {
            w = strtod(tok.c_str() + colon + 1, &e);
            if (*e != 0 || e == tok.c_str() + colon + 1) return "malformed value in `" + tok + "`";
        }
        if (!(w >= 0.0) || w > 3.0e38) return "weights must be finite and non-negative";
        if (count[(size_t)cur] > 0 && ri.back() >= (int32_t)r) return "rows of a column must ascend";
        ri.push_back((int32_t)r);
        v.push_back((float)w);
        count[(size_t)cur]++;
    }
    cp.assign((size_t)n + 1, 0);
    for (int64_t c = 0; c < n; c++) cp[(size_t)c + 1] = cp[(size_t)c] + count[(size_t)c];
    *n_out = n;
    return std::string();
}
