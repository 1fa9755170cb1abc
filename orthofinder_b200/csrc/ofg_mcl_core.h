// MCL on the device: the per-column work of one iteration of the process that the reference
// delegates to its bundled `mcl` binary (scripts_of/mcl.py:214-218 RunMCL; mcl 14-137, default resource scheme).
//
// Host/device-shared: every function here is written in PHASES - `MCL_PHASE(t) { ... }` is the body of thread t of a
// 256-thread CTA followed by a barrier.  On the device a phase is `for (t = threadIdx.x; once; __syncthreads())`, on the
// host it is a loop over all 256 t, so the CPU build (libofg_hostcheck.so, tests/test_mcl_host_core.py) walks exactly
// the control flow, the hash-table probing, the sorting network and the selection that the CTA executes.  Code between
// phases is CTA-uniform (every thread computes the same scalars from shared memory).
//
// Arithmetic (what makes the iterands bit-identical to the binary's; pinned on the CPU by the test oracle):
//   expansion   entry r of column c of M*M = sum over the entries (k, f) of column c, IN THEIR ORDER, of f * M[r,k],
//               accumulated in double; the product is float*float rounded to float for columns with
//               summands * overhead >= n (the binary's dense walk), the exact double product otherwise.
//               One CTA per column: a hash table (shared memory up to 2 048 slots, global memory beyond) keyed by
//               row - or, for a column that may touch a quarter of all rows, one accumulator per row (no probing, and
//               the result is collected in row order without a sort); source entries are taken one at a time (barrier in between), the 256 threads spread over the
//               rows of M[:,k] - distinct rows, so every accumulator sees its addends in source order.
//   pruning     val >= 0.99999/P; recovery / selection with "the K largest, ties kept" by a radix select on the float
//               bits; chaos on what is left; val / (double sum) - the sum of floats >= 1e-8 that add up to < 2 is exact in
//               double in any order, smaller entries (recovery only) are summed sequentially.
//   inflation   (float) pow((double) val, (double)(float) I), / the double sum (same exactness argument).
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define MCL_HD __host__ __device__ __forceinline__
#else
#define MCL_HD inline
#endif

typedef unsigned long long mcl_u64;

constexpr int MCL_NT = 256;            // threads per CTA (the radix-select histogram relies on 256)
constexpr int MCL_SMEM_SLOTS = 2048;   // hash slots kept in shared memory
constexpr int MCL_SMEM_LIST = 1024;    // list entries kept in shared memory (= MCL_SMEM_SLOTS / 2)
constexpr int MCL_DIRECT_FACTOR = 4;   // columns that may touch n / 4 rows or more accumulate in a row-indexed array
constexpr int MCL_RUN = 16;            // rows per thread and step when such an array is collected

#if defined(__CUDA_ARCH__)
#define MCL_PHASE(t) for (int t = (int)threadIdx.x, mcl_once_ = 1; mcl_once_; mcl_once_ = 0, __syncthreads())
#define MCL_ATOMIC_CAS_I32(addr, cmp, val) atomicCAS((int*)(addr), (int)(cmp), (int)(val))
#define MCL_ATOMIC_ADD_U32(addr, val) atomicAdd((unsigned*)(addr), (unsigned)(val))
#define MCL_ATOMIC_MIN_I32(addr, val) atomicMin((int*)(addr), (int)(val))
#define MCL_ATOMIC_MAX_U64(addr, val) atomicMax((unsigned long long*)(addr), (unsigned long long)(val))
#define MCL_FMUL(a, b) __fmul_rn((a), (b))
#define MCL_DMUL(a, b) __dmul_rn((a), (b))
#define MCL_DADD(a, b) __dadd_rn((a), (b))
#define MCL_DDIV(a, b) __ddiv_rn((a), (b))
#define MCL_D2F(a) __double2float_rn(a)
#else
#define MCL_PHASE(t) for (int t = 0; t < MCL_NT; ++t)
static inline int32_t mcl_host_cas(int32_t* a, int32_t cmp, int32_t val) { int32_t o = *a; if (o == cmp) *a = val; return o; }
static inline int32_t mcl_host_min(int32_t* a, int32_t val) { int32_t o = *a; if (val < o) *a = val; return o; }
static inline mcl_u64 mcl_host_max64(mcl_u64* a, mcl_u64 val) { mcl_u64 o = *a; if (val > o) *a = val; return o; }
#define MCL_ATOMIC_CAS_I32(addr, cmp, val) mcl_host_cas((int32_t*)(addr), (cmp), (val))
#define MCL_ATOMIC_ADD_U32(addr, val) (*(addr) += (val))
#define MCL_ATOMIC_MIN_I32(addr, val) mcl_host_min((int32_t*)(addr), (val))
#define MCL_ATOMIC_MAX_U64(addr, val) mcl_host_max64((mcl_u64*)(addr), (val))
#define MCL_FMUL(a, b) ((float)((float)(a) * (float)(b)))   // host build: -ffp-contract=off
#define MCL_DMUL(a, b) ((double)(a) * (double)(b))
#define MCL_DADD(a, b) ((double)(a) + (double)(b))
#define MCL_DDIV(a, b) ((double)(a) / (double)(b))
#define MCL_D2F(a) ((float)(a))
#endif

MCL_HD uint32_t mcl_float_bits(float f) { uint32_t b; memcpy(&b, &f, 4); return b; }
MCL_HD float mcl_bits_float(uint32_t b) { float f; memcpy(&f, &b, 4); return f; }
MCL_HD mcl_u64 mcl_double_bits(double d) { mcl_u64 b; memcpy(&b, &d, 8); return b; }
MCL_HD uint32_t mcl_hash(int32_t r) { uint32_t x = (uint32_t)r * 2654435761u; return x ^ (x >> 15); }
MCL_HD mcl_u64 mcl_pack(int32_t row, float val) { return ((mcl_u64)(uint32_t)row << 32) | (mcl_u64)mcl_float_bits(val); }
MCL_HD int32_t mcl_row_of(mcl_u64 e) { return (int32_t)(uint32_t)(e >> 32); }
MCL_HD float mcl_val_of(mcl_u64 e) { return mcl_bits_float((uint32_t)e); }

// sums of floats this large (and adding up to < 2) are exact in double whatever the order
constexpr float MCL_EXACT_SUM_MIN = 1e-8f;

// ---------------------------------------------------------------------------------------------- start matrix
// As the binary reads the graph (one thread per column): entries that read as 0 and existing loops are dropped, the loop
// is the largest other entry (1 for an empty column), the column is divided by its double sum.
struct MclStartArgs {
    int64_t n;
    const int64_t* in_cp;   // n + 1
    const int32_t* in_ri;
    const float* in_v;
    int32_t* cnt;           // entries of the start column (count pass)
    const int64_t* out_cp;  // n + 1 (fill pass)
    int32_t* out_ri;
    float* out_v;
};

MCL_HD void mcl_start_count(const MclStartArgs& a, int64_t c)
{
    int k = 1;
    for (int64_t p = a.in_cp[c]; p < a.in_cp[c + 1]; p++)
        if (a.in_ri[p] != (int32_t)c && a.in_v[p] != 0.0f) k++;
    a.cnt[c] = k;
}

MCL_HD void mcl_start_fill(const MclStartArgs& a, int64_t c)
{
    float mx = 0.0f;
    bool any = false;
    for (int64_t p = a.in_cp[c]; p < a.in_cp[c + 1]; p++)
        if (a.in_ri[p] != (int32_t)c && a.in_v[p] != 0.0f) { if (!any || a.in_v[p] > mx) mx = a.in_v[p]; any = true; }
    const float loop = any ? mx : 1.0f;
    const int64_t b = a.out_cp[c];
    int64_t o = b;
    bool placed = false;
    for (int64_t p = a.in_cp[c]; p < a.in_cp[c + 1]; p++) {
        if (a.in_ri[p] == (int32_t)c || a.in_v[p] == 0.0f) continue;
        if (!placed && a.in_ri[p] > (int32_t)c) { a.out_ri[o] = (int32_t)c; a.out_v[o++] = loop; placed = true; }
        a.out_ri[o] = a.in_ri[p]; a.out_v[o++] = a.in_v[p];
    }
    if (!placed) { a.out_ri[o] = (int32_t)c; a.out_v[o++] = loop; }
    double s = 0.0;
    for (int64_t p = b; p < o; p++) s = MCL_DADD(s, (double)a.out_v[p]);
    for (int64_t p = b; p < o; p++) a.out_v[p] = MCL_D2F(MCL_DDIV((double)a.out_v[p], s));
}

// The value mcl stores for a weight that WriteGraph_perSpecies printed with "%.3f" (gathering.py:44): strtod of the decimal
// string, then float.  X = round-half-even(w * 1000) on the exact binary value of w (what glibc's printf prints); X / 1000 in
// double is the correctly rounded value of that decimal.  false: negative, not finite or too large.
MCL_HD bool mcl_weight_as_read(double w, float* out)
{
    const mcl_u64 bits = mcl_double_bits(w);
    *out = 0.0f;
    if (bits >> 63) return (bits << 1) == 0;          // -0.0 prints as "-0.000" and reads as 0
    const int eb = (int)((bits >> 52) & 0x7ff);
    if (eb == 0x7ff) return false;
    mcl_u64 mant = bits & 0x000fffffffffffffULL;
    int e;
    if (eb == 0) e = -1074; else { mant |= 0x0010000000000000ULL; e = eb - 1075; }
    const mcl_u64 P = mant * 1000ULL;                 // < 2^63
    mcl_u64 X;
    if (e >= 0) {
        if (e >= 11 || (e > 0 && (P >> (64 - e)) != 0)) return false;
        X = P << e;
    } else {
        const int sh = -e;
        if (sh >= 64) X = 0;
        else {
            X = P >> sh;
            const mcl_u64 rem = P & ((1ULL << sh) - 1ULL), half = 1ULL << (sh - 1);
            if (rem > half || (rem == half && (X & 1ULL))) X++;
        }
    }
    if (X >= (1ULL << 53)) return false;
    *out = MCL_D2F(MCL_DDIV((double)X, 1000.0));
    return true;
}

// ---------------------------------------------------------------------------------------------- plan of one iteration
struct MclPlanArgs {
    int64_t n;
    const int64_t* cp;
    const int32_t* ri;
    int overhead;
    int direct_factor; // columns that may touch n / direct_factor rows or more use row-indexed accumulators (0: never)
    int32_t* tcap;     // hash capacity of the column (power of two); 0: row-indexed accumulators (n of them)
    uint8_t* dense;    // 1: products rounded to float (summands * overhead >= n)
    int64_t* need;     // slots the column needs in the GLOBAL table area (0: its table fits shared memory)
    int64_t* stg;      // entries the column needs in the stage area
};

MCL_HD void mcl_plan_column(const MclPlanArgs& a, int64_t c)
{
    int64_t ns = 0;
    for (int64_t p = a.cp[c]; p < a.cp[c + 1]; p++) ns += a.cp[a.ri[p] + 1] - a.cp[a.ri[p]];
    a.dense[c] = (ns * (int64_t)a.overhead < a.n) ? 0 : 1;
    int64_t bound = ns < a.n ? ns : a.n;
    if (bound < 1) bound = 1;
    int64_t cap = 64;
    while (cap < 2 * bound) cap <<= 1;
    const bool in_smem = cap <= MCL_SMEM_SLOTS;
    if (!in_smem && a.direct_factor > 0 && bound * (int64_t)a.direct_factor >= a.n) {      // a large share of all rows: one accumulator per row, no hashing, no sort
        a.tcap[c] = 0;
        a.need[c] = a.n;
        a.stg[c] = bound;
        return;
    }
    a.tcap[c] = (int32_t)cap;
    a.need[c] = in_smem ? 0 : cap;
    int64_t lst = 1;
    while (lst < bound) lst <<= 1;          // the sorting network pads the list to a power of two
    a.stg[c] = in_smem ? bound : lst;
}

// ---------------------------------------------------------------------------------------------- expansion + pruning
struct MclExpandArgs {
    int64_t n;
    const int64_t* cp;
    const int32_t* ri;
    const float* v;
    int64_t c0;                 // first column of the batch
    const int64_t* table_off;   // exclusive scans over ALL columns (n + 1); the batch's areas start at [c0]
    const int64_t* stage_off;
    const int32_t* tcap;
    const uint8_t* dense;
    int32_t* gkeys;             // global tables of the batch (keys preset to -1)
    double* gacc;
    mcl_u64* stage;             // per column: list scratch, then the kept entries (row << 32 | float bits)
    int32_t* cnt;               // kept entries per column
    mcl_u64* chaos_bits;        // max over columns (bits of a non-negative double)
    int32_t* err;               // set when a column's table overflowed (inconsistent plan)
    double cut;                 // 0.99999 / P
    int S, R;
    double pct;
};

struct MclColShared {
    int32_t keys[MCL_SMEM_SLOTS];
    double acc[MCL_SMEM_SLOTS];
    mcl_u64 list[MCL_SMEM_LIST];
    mcl_u64 tmp64[MCL_NT];
    int32_t scan[MCL_NT];
    int32_t scan2[16];
    double red_a[MCL_NT], red_b[MCL_NT], red_c[MCL_NT];
    double red2_a[16], red2_b[16], red2_c[16];
    int32_t red_i[MCL_NT], red2_i[16];
    float red_f[MCL_NT], red2_f[16];
    uint32_t hist[256];
    int32_t scan_total;
    double tot_a, tot_b, tot_c;   // sums a, b; max c
    int32_t tot_i;                // sum
    float tot_f;                  // min
    int32_t sel_bin, sel_above;
    double seq_sum;
};

// the reduction scratch alone (inflation needs nothing else)
struct MclRedShared {
    double red_a[MCL_NT], red_b[MCL_NT], red_c[MCL_NT];
    double red2_a[16], red2_b[16], red2_c[16];
    int32_t red_i[MCL_NT], red2_i[16];
    float red_f[MCL_NT], red2_f[16];
    double tot_a, tot_b, tot_c;
    int32_t tot_i;
    float tot_f;
    double seq_sum;
};

// the same without the hash table and the list (columns with row-indexed accumulators)
struct MclColSharedLite {
    mcl_u64 tmp64[MCL_NT];
    int32_t scan[MCL_NT];
    int32_t scan2[16];
    double red_a[MCL_NT], red_b[MCL_NT], red_c[MCL_NT];
    double red2_a[16], red2_b[16], red2_c[16];
    int32_t red_i[MCL_NT], red2_i[16];
    float red_f[MCL_NT], red2_f[16];
    uint32_t hist[256];
    int32_t scan_total;
    double tot_a, tot_b, tot_c;
    int32_t tot_i;
    float tot_f;
    int32_t sel_bin, sel_above;
    double seq_sum;
};

// exclusive scan of sh.scan[0..256), total in sh.scan_total
template <class SH> MCL_HD void mcl_block_scan(SH& sh)
{
    MCL_PHASE(t) {
        if (t < 16) {
            int s = 0;
            for (int i = 0; i < 16; i++) { const int x = sh.scan[t * 16 + i]; sh.scan[t * 16 + i] = s; s += x; }
            sh.scan2[t] = s;
        }
    }
    MCL_PHASE(t) {
        if (t == 0) {
            int s = 0;
            for (int i = 0; i < 16; i++) { const int x = sh.scan2[i]; sh.scan2[i] = s; s += x; }
            sh.scan_total = s;
        }
    }
    MCL_PHASE(t) { sh.scan[t] += sh.scan2[t >> 4]; }
}

// reduces the per-thread partials red_a/red_b (sums), red_c (max), red_i (sum), red_f (min) into tot_*
template <class SH> MCL_HD void mcl_block_reduce(SH& sh)
{
    MCL_PHASE(t) {
        if (t < 16) {
            double a = 0.0, b = 0.0, c = sh.red_c[t * 16];
            int k = 0;
            float f = sh.red_f[t * 16];
            for (int i = 0; i < 16; i++) {
                a = MCL_DADD(a, sh.red_a[t * 16 + i]);
                b = MCL_DADD(b, sh.red_b[t * 16 + i]);
                if (sh.red_c[t * 16 + i] > c) c = sh.red_c[t * 16 + i];
                k += sh.red_i[t * 16 + i];
                if (sh.red_f[t * 16 + i] < f) f = sh.red_f[t * 16 + i];
            }
            sh.red2_a[t] = a; sh.red2_b[t] = b; sh.red2_c[t] = c; sh.red2_i[t] = k; sh.red2_f[t] = f;
        }
    }
    MCL_PHASE(t) {
        if (t == 0) {
            double a = 0.0, b = 0.0, c = sh.red2_c[0];
            int k = 0;
            float f = sh.red2_f[0];
            for (int i = 0; i < 16; i++) {
                a = MCL_DADD(a, sh.red2_a[i]);
                b = MCL_DADD(b, sh.red2_b[i]);
                if (sh.red2_c[i] > c) c = sh.red2_c[i];
                k += sh.red2_i[i];
                if (sh.red2_f[i] < f) f = sh.red2_f[i];
            }
            sh.tot_a = a; sh.tot_b = b; sh.tot_c = c; sh.tot_i = k; sh.tot_f = f;
        }
    }
}

// the K-th largest value among the list entries with (double)val >= cutd (there are at least K of them): radix select on
// the float bits (all values are >= 0, so the bit patterns order like the values)
template <class SH> MCL_HD float mcl_kth_largest(SH& sh, const mcl_u64* list, int nt, double cutd, int K)
{
    uint32_t prefix = 0, pmask = 0;
    int kk = K;
    for (int shift = 24; shift >= 0; shift -= 8) {
        MCL_PHASE(t) { sh.hist[t] = 0; }
        MCL_PHASE(t) {
            for (int i = t; i < nt; i += MCL_NT) {
                const uint32_t b = (uint32_t)list[i];
                if ((double)mcl_bits_float(b) >= cutd && (b & pmask) == prefix) MCL_ATOMIC_ADD_U32(&sh.hist[(b >> shift) & 255u], 1u);
            }
        }
        MCL_PHASE(t) {
            if (t == 0) {
                int s = 0, bin = 255;
                for (; bin > 0; bin--) {
                    if (s + (int)sh.hist[bin] >= kk) break;
                    s += (int)sh.hist[bin];
                }
                sh.sel_bin = bin; sh.sel_above = s;
            }
        }
        prefix |= (uint32_t)sh.sel_bin << shift;
        pmask |= 255u << shift;
        kk -= sh.sel_above;
    }
    return mcl_bits_float(prefix);
}

// steps 6-8 on the expanded column: `list` holds nt entries (row << 32 | float bits) in row order
template <class SH> MCL_HD void mcl_prune_column(SH& sh, const MclExpandArgs& a, int64_t c, mcl_u64* list, int nt, mcl_u64* stage)
{
    // 6. pruning rule
    MCL_PHASE(t) {
        double ma = 0.0, mp = 0.0;
        int np = 0;
        for (int i = t; i < nt; i += MCL_NT) {
            const float val = mcl_val_of(list[i]);
            ma = MCL_DADD(ma, (double)val);
            if ((double)val >= a.cut) { mp = MCL_DADD(mp, (double)val); np++; }
        }
        sh.red_a[t] = ma; sh.red_b[t] = mp; sh.red_c[t] = 0.0; sh.red_i[t] = np; sh.red_f[t] = 0.0f;
    }
    mcl_block_reduce(sh);
    const double mass_all = sh.tot_a, mass_p = sh.tot_b;
    const int n_p = sh.tot_i;
    double thr_cut = a.cut;
    float bar = 0.0f;
    if (a.R && n_p < a.R && mass_p < a.pct * mass_all) {        // recovery from the whole vector
        thr_cut = 0.0;
        if (nt > a.R) bar = mcl_kth_largest(sh, list, nt, 0.0, a.R);
    } else if (a.S && n_p > a.S) {                              // selection among the pruned entries
        bar = mcl_kth_largest(sh, list, nt, a.cut, a.S);
        MCL_PHASE(t) {
            double ms = 0.0;
            int k = 0;
            for (int i = t; i < nt; i += MCL_NT) {
                const float val = mcl_val_of(list[i]);
                if ((double)val >= a.cut && val >= bar) { ms = MCL_DADD(ms, (double)val); k++; }
            }
            sh.red_a[t] = ms; sh.red_b[t] = 0.0; sh.red_c[t] = 0.0; sh.red_i[t] = k; sh.red_f[t] = 0.0f;
        }
        mcl_block_reduce(sh);
        const double mass_s = sh.tot_a;
        const int n_s = sh.tot_i;
        if (a.R && n_s < a.R && mass_s < a.pct * mass_all)      // recovery from the pruned vector
            bar = (n_p <= a.R) ? 0.0f : mcl_kth_largest(sh, list, nt, a.cut, a.R);
    }
    // 7. what is left: count, max, sum of squares (chaos), sum, smallest entry
    MCL_PHASE(t) {
        double s = 0.0, ssq = 0.0, mx = 0.0;
        int k = 0;
        float mn = 3.0e38f;
        for (int i = t; i < nt; i += MCL_NT) {
            const float val = mcl_val_of(list[i]);
            if ((double)val >= thr_cut && val >= bar) {
                s = MCL_DADD(s, (double)val);
                ssq = MCL_DADD(ssq, MCL_DMUL((double)val, (double)val));
                if ((double)val > mx) mx = (double)val;
                if (val < mn) mn = val;
                k++;
            }
        }
        sh.red_a[t] = s; sh.red_b[t] = ssq; sh.red_c[t] = mx; sh.red_i[t] = k; sh.red_f[t] = mn;
    }
    mcl_block_reduce(sh);
    const int kept = sh.tot_i;
    const bool exact = sh.tot_f >= MCL_EXACT_SUM_MIN;
    const double chaos = MCL_DMUL(MCL_DADD(sh.tot_c, -sh.tot_b), (double)kept);
    double sum = sh.tot_a;
    MCL_PHASE(t) {
        if (t == 0) {
            if (chaos > 0.0) MCL_ATOMIC_MAX_U64(a.chaos_bits, mcl_double_bits(chaos));
            if (!exact) {                       // tiny entries (recovery): the reference order, one addend at a time
                double s = 0.0;
                for (int i = 0; i < nt; i++) {
                    const float val = mcl_val_of(list[i]);
                    if ((double)val >= thr_cut && val >= bar) s = MCL_DADD(s, (double)val);
                }
                sh.seq_sum = s;
            }
        }
    }
    if (!exact) sum = sh.seq_sum;
    // 8. kept entries, rescaled, in row order -> stage
    int outn = 0;
    for (int base = 0; base < nt; base += MCL_NT) {
        MCL_PHASE(t) {
            mcl_u64 e = 0;
            int flag = 0;
            if (base + t < nt) {
                e = list[base + t];
                const float val = mcl_val_of(e);
                flag = ((double)val >= thr_cut && val >= bar) ? 1 : 0;
            }
            sh.tmp64[t] = e;
            sh.scan[t] = flag;
            sh.red_i[t] = flag;
        }
        mcl_block_scan(sh);
        MCL_PHASE(t) {
            if (sh.red_i[t]) {
                const mcl_u64 e = sh.tmp64[t];
                const float w = MCL_D2F(MCL_DDIV((double)mcl_val_of(e), sum));
                stage[outn + sh.scan[t]] = (e & 0xffffffff00000000ull) | (mcl_u64)mcl_float_bits(w);
            }
        }
        outn += sh.scan_total;
    }
    MCL_PHASE(t) { if (t == 0) a.cnt[c] = outn; }
}

// expansion of a column with hashed accumulators (shared memory up to MCL_SMEM_SLOTS slots, else the global table area)
template <class SH> MCL_HD void mcl_expand_column_hashed(SH& sh, const MclExpandArgs& a, int64_t c)
{
    const int tcap = a.tcap[c];
    if (tcap == 0) return;                         // a column of the row-indexed kernel (CTA-uniform: before any barrier)
    const int64_t p0 = a.cp[c], p1 = a.cp[c + 1];
    const bool in_smem = tcap <= MCL_SMEM_SLOTS;
    int32_t* keys = in_smem ? sh.keys : a.gkeys + (a.table_off[c] - a.table_off[a.c0]);
    double* acc = in_smem ? sh.acc : a.gacc + (a.table_off[c] - a.table_off[a.c0]);
    mcl_u64* stage = a.stage + (a.stage_off[c] - a.stage_off[a.c0]);
    mcl_u64* list = in_smem ? sh.list : stage;
    const uint32_t mask = (uint32_t)tcap - 1u;
    const bool dense = a.dense[c] != 0;

    // 1. empty table (the global tables were preset by the host)
    if (in_smem) {
        MCL_PHASE(t) { for (int i = t; i < tcap; i += MCL_NT) keys[i] = -1; }
    }
    // 2. accumulate: one source entry at a time, its column spread over the threads
    for (int64_t p = p0; p < p1; p++) {
        const int32_t k = a.ri[p];
        const float f = a.v[p];
        const int64_t q0 = a.cp[k], q1 = a.cp[k + 1];
        MCL_PHASE(t) {
            for (int64_t q = q0 + t; q < q1; q += MCL_NT) {
                const int32_t r = a.ri[q];
                const float w = a.v[q];
                uint32_t h = mcl_hash(r) & mask;
                int probes = 0;
                for (; probes <= tcap; probes++) {
                    const int32_t old = MCL_ATOMIC_CAS_I32(&keys[h], -1, r);
                    if (old == -1) { acc[h] = 0.0; break; }
                    if (old == r) break;
                    h = (h + 1u) & mask;
                }
                if (probes > tcap) { *a.err = 1; continue; }     // cannot happen with a consistent plan: never spin on the GPU
                const double pr = dense ? (double)MCL_FMUL(f, w) : MCL_DMUL((double)f, (double)w);
                acc[h] = MCL_DADD(acc[h], pr);
            }
        }
    }
    // 3. list of the occupied slots: (row << 32) | slot
    int nt = 0;
    for (int base = 0; base < tcap; base += MCL_NT) {
        MCL_PHASE(t) { sh.scan[t] = (base + t < tcap && keys[base + t] != -1) ? 1 : 0; }
        mcl_block_scan(sh);
        MCL_PHASE(t) {
            if (base + t < tcap && keys[base + t] != -1)
                list[nt + sh.scan[t]] = ((mcl_u64)(uint32_t)keys[base + t] << 32) | (mcl_u64)(uint32_t)(base + t);
        }
        nt += sh.scan_total;
    }
    // 4. rows ascending: bitonic network over the list padded to a power of two
    int m = 1;
    while (m < nt) m <<= 1;
    MCL_PHASE(t) { for (int i = nt + t; i < m; i += MCL_NT) list[i] = ~0ull; }
    for (int k2 = 2; k2 <= m; k2 <<= 1) {
        for (int j = k2 >> 1; j > 0; j >>= 1) {
            MCL_PHASE(t) {
                for (int i = t; i < m; i += MCL_NT) {
                    const int l = i ^ j;
                    if (l > i) {
                        const mcl_u64 x = list[i], y = list[l];
                        const bool up = (i & k2) == 0;
                        if ((x > y) == up) { list[i] = y; list[l] = x; }
                    }
                }
            }
        }
    }
    // 5. values: (row << 32) | bits of (float) sum
    MCL_PHASE(t) {
        for (int i = t; i < nt; i += MCL_NT) {
            const mcl_u64 e = list[i];
            list[i] = (e & 0xffffffff00000000ull) | (mcl_u64)mcl_float_bits(MCL_D2F(acc[(uint32_t)e]));
        }
    }
    mcl_prune_column(sh, a, c, list, nt, stage);
}

// expansion of a column that may touch a large share of the rows: one accumulator per row in the global table area (no
// probing), four source rows in flight per thread, the result collected in row order (no sort)
template <class SH> MCL_HD void mcl_expand_column_direct(SH& sh, const MclExpandArgs& a, int64_t c)
{
    if (a.tcap[c] != 0) return;                    // a column of the hashed kernel (CTA-uniform: before any barrier)
    const int64_t p0 = a.cp[c], p1 = a.cp[c + 1];
    int32_t* keys = a.gkeys + (a.table_off[c] - a.table_off[a.c0]);      // -1: row not touched yet
    double* acc = a.gacc + (a.table_off[c] - a.table_off[a.c0]);
    mcl_u64* stage = a.stage + (a.stage_off[c] - a.stage_off[a.c0]);
    mcl_u64* list = stage;
    const bool dense = a.dense[c] != 0;
    for (int64_t p = p0; p < p1; p++) {
        const int32_t k = a.ri[p];
        const float f = a.v[p];
        const int64_t q0 = a.cp[k], q1 = a.cp[k + 1];
        MCL_PHASE(t) {
            for (int64_t q = q0 + t; q < q1; q += 4 * (int64_t)MCL_NT) {     // rows of one source column are distinct: no two meet
                int32_t r[4], kk[4];
                double pr[4], av[4];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
                for (int j = 0; j < 4; j++) {
                    const int64_t qq = q + (int64_t)j * MCL_NT;
                    r[j] = -1;
                    if (qq < q1) {
                        r[j] = a.ri[qq];
                        const float w = a.v[qq];
                        pr[j] = dense ? (double)MCL_FMUL(f, w) : MCL_DMUL((double)f, (double)w);
                    }
                }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
                for (int j = 0; j < 4; j++)
                    if (r[j] >= 0) { kk[j] = keys[r[j]]; av[j] = kk[j] == -1 ? 0.0 : acc[r[j]]; }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
                for (int j = 0; j < 4; j++)
                    if (r[j] >= 0) {
                        acc[r[j]] = MCL_DADD(av[j], pr[j]);
                        if (kk[j] == -1) keys[r[j]] = r[j];
                    }
            }
        }
    }
    // the touched rows in row order, with their values: the list is sorted as it is written
    int nt = 0;
    for (int64_t base = 0; base < a.n; base += (int64_t)MCL_RUN * MCL_NT) {
        MCL_PHASE(t) {
            const int64_t lo = base + (int64_t)t * MCL_RUN;
            int k = 0;
            for (int64_t i = lo; i < lo + MCL_RUN && i < a.n; i++) k += keys[i] != -1;
            sh.scan[t] = k;
        }
        mcl_block_scan(sh);
        MCL_PHASE(t) {
            const int64_t lo = base + (int64_t)t * MCL_RUN;
            int o = nt + sh.scan[t];
            for (int64_t i = lo; i < lo + MCL_RUN && i < a.n; i++)
                if (keys[i] != -1) list[o++] = mcl_pack((int32_t)i, MCL_D2F(acc[i]));
        }
        nt += sh.scan_total;
    }
    mcl_prune_column(sh, a, c, list, nt, stage);
}

// ---------------------------------------------------------------------------------------------- inflation + packing
struct MclInflateArgs {
    int64_t c0;
    const int64_t* stage_off;
    const mcl_u64* stage;
    const int32_t* cnt;
    const int64_t* new_cp;   // n + 1, absolute
    int32_t* out_ri;
    float* out_v;
    double power;            // (double)(float) inflation
};

template <class SH> MCL_HD void mcl_inflate_column(SH& sh, const MclInflateArgs& a, int64_t c)
{
    const mcl_u64* src = a.stage + (a.stage_off[c] - a.stage_off[a.c0]);
    const int cnt = a.cnt[c];
    const int64_t d0 = a.new_cp[c];
    MCL_PHASE(t) {
        double s = 0.0;
        float mn = 3.0e38f;
        for (int i = t; i < cnt; i += MCL_NT) {
            const mcl_u64 e = src[i];
            const float p = MCL_D2F(pow((double)mcl_val_of(e), a.power));
            a.out_ri[d0 + i] = mcl_row_of(e);
            a.out_v[d0 + i] = p;
            s = MCL_DADD(s, (double)p);
            if (p < mn) mn = p;
        }
        sh.red_a[t] = s; sh.red_b[t] = 0.0; sh.red_c[t] = 0.0; sh.red_i[t] = 0; sh.red_f[t] = mn;
    }
    mcl_block_reduce(sh);
    const bool exact = sh.tot_f >= MCL_EXACT_SUM_MIN;
    double sum = sh.tot_a;
    MCL_PHASE(t) {
        if (t == 0 && !exact) {
            double s = 0.0;
            for (int i = 0; i < cnt; i++) s = MCL_DADD(s, (double)a.out_v[d0 + i]);
            sh.seq_sum = s;
        }
    }
    if (!exact) sum = sh.seq_sum;
    MCL_PHASE(t) {
        for (int i = t; i < cnt; i += MCL_NT) a.out_v[d0 + i] = MCL_D2F(MCL_DDIV((double)a.out_v[d0 + i], sum));
    }
}

// ---------------------------------------------------------------------------------------------- interpretation
// dag of the limit: per column keep val >= bar, bar = 0.999 self + 0.001 max if self < max, else self / 1.01; a node with a
// loop in the dag is an attractor.  One thread per column.
constexpr int32_t MCL_NONE = 0x7fffffff;

struct MclDagArgs {
    int64_t n;
    const int64_t* cp;
    const int32_t* ri;
    const float* v;
    uint8_t* keep;      // per entry
    uint8_t* attr;      // per node
    int32_t* sys;       // smallest attractor of the first system the node reaches (MCL_NONE: none yet)
    int32_t* second;    // of another system it reaches (overlap), or MCL_NONE
    int32_t* changed;
};

MCL_HD void mcl_dag_column(const MclDagArgs& a, int64_t c)
{
    double self = 0.0, mx = 0.0;
    for (int64_t p = a.cp[c]; p < a.cp[c + 1]; p++) {
        if (a.ri[p] == (int32_t)c) self = (double)a.v[p];
        if ((double)a.v[p] > mx) mx = (double)a.v[p];
    }
    const double bar = self < mx ? MCL_DADD(MCL_DMUL(0.999, self), MCL_DMUL(0.001, mx)) : MCL_DDIV(self, 1.01);
    bool at = false;
    for (int64_t p = a.cp[c]; p < a.cp[c + 1]; p++) {
        const bool k = (double)a.v[p] >= bar;
        a.keep[p] = k ? 1 : 0;
        if (k && a.ri[p] == (int32_t)c) at = true;
    }
    a.attr[c] = at ? 1 : 0;
    a.sys[c] = at ? (int32_t)c : MCL_NONE;
    a.second[c] = MCL_NONE;
}

// attractor systems: labels flow along the dag arcs between attractors in both directions until nothing changes
MCL_HD void mcl_attractor_round(const MclDagArgs& a, int64_t c)
{
    if (!a.attr[c]) return;
    for (int64_t p = a.cp[c]; p < a.cp[c + 1]; p++) {
        const int32_t r = a.ri[p];
        if (!a.keep[p] || !a.attr[r] || r == (int32_t)c) continue;
        const int32_t x = a.sys[c], y = a.sys[r];
        if (x == y) continue;
        const int32_t m = x < y ? x : y;
        if (MCL_ATOMIC_MIN_I32(&a.sys[c], m) > m) *a.changed = 1;
        if (MCL_ATOMIC_MIN_I32(&a.sys[r], m) > m) *a.changed = 1;
    }
}

// the other nodes: the two smallest systems reachable through their dag arcs
MCL_HD void mcl_reach_round(const MclDagArgs& a, int64_t c)
{
    if (a.attr[c]) return;
    int32_t s1 = a.sys[c], s2 = a.second[c];
    bool ch = false;
    for (int64_t p = a.cp[c]; p < a.cp[c + 1]; p++) {
        if (!a.keep[p]) continue;
        const int32_t r = a.ri[p];
        if (r == (int32_t)c) continue;
        const int32_t cand[2] = {a.sys[r], a.second[r]};
        for (int t = 0; t < 2; t++) {
            const int32_t s = cand[t];
            if (s == MCL_NONE || s == s1 || s == s2) continue;
            if (s < s1) { s2 = s1; s1 = s; ch = true; }
            else if (s < s2) { s2 = s; ch = true; }
        }
    }
    if (ch) { a.sys[c] = s1; a.second[c] = s2; *a.changed = 1; }
}
