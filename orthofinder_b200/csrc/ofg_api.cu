// libofg: context, memory plan and stage orchestration behind the C-ABI of include/ofg.h.
// One translation unit; kernels live in the .cuh files next to it.  Built for sm_100a only.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <math.h>
#include <stdarg.h>
#include <algorithm>
#include <string>
#include <vector>

#include "../../include/ofg.h"
#include "ofg_common.cuh"
#include "ofg_parse.cuh"
#include "ofg_rows.cuh"
#include "ofg_norm.cuh"
#include "ofg_graph.cuh"
#include "ofg_synth.cuh"

static_assert(sizeof(ofg_synth_params) == sizeof(OfgSynthParams), "synth parameter layouts must match");

namespace {

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 16 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct FileInfo {
    int64_t off = -1, raw = 0, padded = 0;
    int swap = 0;
};

}  // namespace

struct ofg_ctx {
    int device = 0;
    int n_sm = 148;
    cudaStream_t stream = nullptr;
    std::string err;
    // species
    int S = 0;
    std::vector<int64_t> n_seqs, seq_start, len_off;
    std::vector<double> lengths;
    int64_t N = 0;
    // partition
    std::vector<int32_t> owned;
    int with_columns = 0;
    std::vector<PairDesc> pairs;
    std::vector<int32_t> pair_of;
    std::vector<int64_t> pair_row_base, pair_col_base;
    int64_t n_rows = 0, n_cols = 0;
    std::vector<FileInfo> files;
    // options
    int allow_empty = 0;
    bool have_length_factors = false;  // ofg_set_length_factors was called for the current species
    bool layout_dirty = true;   // species / partition changed: per-gene constant vectors must be rebuilt on the device
    bool scores_v2 = false;     // option scores_v2: NormalisedBitScore + percentile-ratio thresholds (gathering.py:157-174,314-388)
    int64_t bytes_per_line_inv = 24;
    int64_t sel_frac_inv = 4;
    // text arena
    DevBuf text;
    int64_t text_used = 0;
    // state
    int stage_done = 0;
    bool thresholds_complete = false;
    bool pairs_dirty = true;
    // device buffers
    DevBuf d_pairs, d_pair_row_base, d_pair_of, d_lengths, d_tile_prefix, d_small;
    DevBuf coo_row, coo_col, coo_val, rowcount, row_start, rowlen, cols, vals, key, keyval, key2, keyval2;
    DevBuf long_rows, counters, scan_partial, sort_tile_prefix, pair_slot, bin_tasks, bin_cut, bin_cnt, sel_off, task_prefix;
    DevBuf sort_desc, zrow, pair_factor, d_owned, gene_cntg, gene_bytesg, most_init, d_length_factors, lx, ly, fit, rowfac, colfac, rowmax, best, most, owned_gene;
    DevBuf gene_of_local, species_of_local, d_seq_start, unit_cnt, unit_bytes, cnt_off, byte_off, out_indices, out_data, out_text;
    DevBuf species_ids, synth_bytes, synth_off, file_off, pair_sel_off, row_flagcnt, row_flagoff;
    // results
    std::vector<u64> pair_lines, pair_valid, pair_nnz;
    std::vector<double> fit_host;
    int64_t tot_lines = 0, tot_records = 0, tot_nnz = 0, tot_edges = 0, tot_text = 0, n_owned_genes = 0;
    u64 sel_cap = 0, sel_total = 0;
    int key_bits = 32;
    float ms[10] = {0};
    int64_t launches = 0;
    cudaEvent_t ev[12] = {nullptr};
    cudaEvent_t tl_ev[192] = {nullptr};
    const char* tl_name[192] = {nullptr};
    int tl_n = 0;
    std::string kernel_report;
};

namespace {

int fail(ofg_ctx* c, int code, const char* fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    c->err = buf;
    return code;
}

#define CK(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess) return fail(c, OFG_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

#define CKL(name)                                                                                        \
    do {                                                                                                 \
        c->launches++;                                                                                   \
        note_launch(c, name);                                                                            \
        cudaError_t e_ = cudaGetLastError();                                                             \
        if (e_ != cudaSuccess) return fail(c, OFG_ERR_CUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

// per-kernel timeline: one event after every launch (cheap), resolved after the build
void note_launch(ofg_ctx* c, const char* name)
{
    if (c->tl_n >= (int)(sizeof(c->tl_ev) / sizeof(c->tl_ev[0]))) return;
    if (!c->tl_ev[c->tl_n]) cudaEventCreate(&c->tl_ev[c->tl_n]);
    cudaEventRecord(c->tl_ev[c->tl_n], c->stream);
    c->tl_name[c->tl_n] = name;
    c->tl_n++;
}
void mark_timeline(ofg_ctx* c)  // start-of-segment marker (after a host sync the previous event is stale)
{
    note_launch(c, nullptr);
}

int grid_for(ofg_ctx* c, int64_t work_items, int per_cta, int ctas_per_sm)
{
    int64_t need = (work_items + per_cta - 1) / per_cta;
    int64_t maxg = (int64_t)c->n_sm * ctas_per_sm;
    if (need < 1) need = 1;
    return (int)std::min(need, maxg);
}

__global__ void gather_u64_kernel(const u64* src, const u64* idx, int n, u64* out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = src[idx[i]];
}

// unit-test kernels (ofg_test_*)
__global__ void test_log10_kernel(const double* x, double* y, int64_t n)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) y[i] = ofg_log10_svml(x[i]);
}
__global__ void test_format_kernel(const double* x, char* out, int64_t n)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        bool ok = true;
        char* o = out + i * 32;
        const int len = ofg_format_entry(7ULL, x[i], o, &ok);
        o[ok ? len : 0] = 0;
    }
}
__global__ void test_score_kernel(const uint8_t* text, int64_t nbytes, const int64_t* starts, int64_t n, double* y, int32_t* ok)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        GlobalSrc gs{text, nbytes};
        double v = 0.0;
        ok[i] = ofg_parse_score_field(gs, starts[i], &v);
        y[i] = v;
    }
}
__global__ void __launch_bounds__(FIT_THREADS) test_fit_kernel(const double* lx, const double* ly, long m, double* out)
{
    extern __shared__ __align__(16) unsigned char fit_smem_raw[];  // sizeof(FitShared) > 48 KB: dynamic, opt-in
    FitShared& sh = *reinterpret_cast<FitShared*>(fit_smem_raw);
    fit_loglinear(lx, ly, m, &sh, out);
}

template <typename TO>
int device_scan(ofg_ctx* c, const u32* in, int64_t n, TO* out, int write_total)
{
    if (n <= 0) {
        if (write_total) CK(cudaMemsetAsync(out, 0, sizeof(TO), c->stream));
        return 0;
    }
    const int64_t n_chunks = (n + SCAN_CHUNK - 1) / SCAN_CHUNK;
    CK(c->scan_partial.ensure((size_t)(n_chunks + 1) * sizeof(u64)));
    TO* part = c->scan_partial.as<TO>();
    const int g = grid_for(c, n_chunks, 1, 8);
    scan_reduce_kernel<TO><<<g, SCAN_THREADS, 0, c->stream>>>(in, n, part);
    CKL("scan_reduce_kernel<TO>");
    scan_partials_kernel<TO><<<1, 1024, 0, c->stream>>>(part, n_chunks, (TO*)nullptr);
    CKL("scan_partials_kernel<TO>");
    scan_apply_kernel<TO><<<g, SCAN_THREADS, 0, c->stream>>>(in, n, part, out, write_total);
    CKL("scan_apply_kernel<TO>");
    return 0;
}

int rebuild_pairs(ofg_ctx* c)
{
    if (c->S <= 0) return fail(c, OFG_ERR_ARG, "ofg_set_species has not been called");
    std::vector<int32_t> owned = c->owned;
    if (owned.empty()) for (int p = 0; p < c->S; p++) owned.push_back(p);
    std::sort(owned.begin(), owned.end());
    std::vector<char> is_owned(c->S, 0);
    for (int p : owned) is_owned[p] = 1;
    std::vector<FileInfo> old_files = c->files;
    std::vector<PairDesc> old_pairs = c->pairs;
    c->pairs.clear();
    c->pair_of.assign((size_t)c->S * c->S, -1);
    auto add = [&](int p, int j, int owned_row) {
        PairDesc d;
        memset(&d, 0, sizeof d);
        d.p = p; d.j = j; d.same = (p == j); d.Gi = (int32_t)c->n_seqs[p]; d.Gj = (int32_t)c->n_seqs[j];
        d.owned_row = owned_row;
        d.li_off = c->len_off[p]; d.lj_off = c->len_off[j];
        d.gi_start = c->seq_start[p]; d.gj_start = c->seq_start[j];
        c->pair_of[(size_t)p * c->S + j] = (int32_t)c->pairs.size();
        c->pairs.push_back(d);
    };
    for (int p : owned) for (int j = 0; j < c->S; j++) add(p, j, 1);
    if (c->with_columns)
        for (int p : owned) for (int j = 0; j < c->S; j++) if (!is_owned[j]) add(j, p, 0);
    const int P = (int)c->pairs.size();
    c->pair_row_base.assign(P + 1, 0);
    c->pair_col_base.assign(P + 1, 0);
    for (int i = 0; i < P; i++) {
        c->pairs[i].row_base = c->pair_row_base[i];
        c->pairs[i].col_base = c->pair_col_base[i];
        c->pair_row_base[i + 1] = c->pair_row_base[i] + c->pairs[i].Gi;
        c->pair_col_base[i + 1] = c->pair_col_base[i] + c->pairs[i].Gj;
    }
    c->n_rows = c->pair_row_base[P];
    c->n_cols = c->pair_col_base[P];
    c->files.assign(P, FileInfo());
    // keep text already added for pairs that still exist
    for (size_t i = 0; i < old_pairs.size() && i < old_files.size(); i++) {
        int np = c->pair_of[(size_t)old_pairs[i].p * c->S + old_pairs[i].j];
        if (np >= 0) c->files[np] = old_files[i];
    }
    c->owned = owned;
    c->n_owned_genes = 0;
    for (int p : owned) c->n_owned_genes += c->n_seqs[p];
    c->pairs_dirty = false;
    c->stage_done = 0;
    return 0;
}

int ensure_pairs(ofg_ctx* c)
{
    if (c->pairs_dirty) return rebuild_pairs(c);
    return 0;
}

int text_reserve(ofg_ctx* c, int64_t need)
{
    // tail pad so that staged overlap reads never leave the arena
    const int64_t tail = PS_RING + 64;
    if ((size_t)(need + tail) <= c->text.cap) return 0;
    DevBuf nb;
    int64_t want = std::max<int64_t>(need + tail, (int64_t)(c->text.cap + c->text.cap / 2));
    CK(nb.ensure((size_t)want));
    CK(cudaMemsetAsync(nb.p, '\n', nb.cap, c->stream));
    if (c->text.p && c->text_used > 0) CK(cudaMemcpyAsync(nb.p, c->text.p, (size_t)c->text_used, cudaMemcpyDeviceToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    c->text.release();
    c->text = nb;
    return 0;
}

int upload(ofg_ctx* c, DevBuf& b, const void* src, size_t bytes)
{
    CK(b.ensure(std::max<size_t>(bytes, 16)));
    if (bytes) CK(cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, c->stream));
    return 0;
}

void describe_line(ofg_ctx* c, int pair, u64 errword, std::string* out)
{
    const PairDesc& pd = c->pairs[pair];
    const FileInfo& f = c->files[pair];
    const int64_t off = (int64_t)(errword >> 4);
    const int kind = (int)(errword & 15);
    // find the line around the offset (the quote kind reports the quote's own offset)
    int64_t lo = std::max<int64_t>(0, off - 4096), hi = std::min<int64_t>(f.raw, off + 4096);
    std::vector<char> buf((size_t)(hi - lo) + 1, 0);
    cudaMemcpy(buf.data(), c->text.as<uint8_t>() + f.off + lo, (size_t)(hi - lo), cudaMemcpyDeviceToHost);
    int64_t s = off - lo, e = off - lo;
    while (s > 0 && buf[s - 1] != '\n' && buf[s - 1] != '\r') s--;
    while (e < hi - lo && buf[e] != '\n' && buf[e] != '\r') e++;
    char head[256];
    snprintf(head, sizeof head, "pair=%d,%d kind=%d offset=%lld line=", pd.p, pd.j, kind, (long long)(lo + s));
    *out = head;
    out->append(buf.data() + s, (size_t)(e - s));
}

// ------------------------------------------------------------------------------------------ stages
int stage_raw(ofg_ctx* c)
{
    mark_timeline(c);
    const int P = (int)c->pairs.size();
    // text offsets into the descriptors
    int64_t total_padded = 0;
    for (int i = 0; i < P; i++) {
        FileInfo& f = c->files[i];
        PairDesc& d = c->pairs[i];
        if (f.off < 0) {
            if (!c->allow_empty) return fail(c, OFG_ERR_MISSING, "hit text for pair (%d,%d) was not provided", d.p, d.j);
            d.present = 0; d.text_off = 0; d.text_len = 0; d.text_raw = 0; d.swap = 0;
        } else {
            d.present = 1; d.text_off = f.off; d.text_len = f.padded; d.text_raw = f.raw; d.swap = f.swap;
        }
        total_padded += d.text_len;
    }
    std::vector<int64_t> tile_prefix(P + 1, 0);  // parse segments before each pair
    for (int i = 0; i < P; i++) tile_prefix[i + 1] = tile_prefix[i] + (c->pairs[i].text_len + PS_SEG - 1) / PS_SEG;
    const int64_t n_tiles = tile_prefix[P];
    const int parse_grid = grid_for(c, (n_tiles + PS_WARPS - 1) / PS_WARPS, 1, 4);
    if (upload(c, c->d_pairs, c->pairs.data(), sizeof(PairDesc) * P)) return OFG_ERR_CUDA;
    if (upload(c, c->d_pair_row_base, c->pair_row_base.data(), sizeof(int64_t) * (P + 1))) return OFG_ERR_CUDA;
    if (upload(c, c->d_pair_of, c->pair_of.data(), sizeof(int32_t) * c->pair_of.size())) return OFG_ERR_CUDA;
    if (upload(c, c->d_lengths, c->lengths.data(), sizeof(double) * c->lengths.size())) return OFG_ERR_CUDA;
    if (upload(c, c->d_tile_prefix, tile_prefix.data(), sizeof(int64_t) * (P + 1))) return OFG_ERR_CUDA;
    if (upload(c, c->d_seq_start, c->seq_start.data(), sizeof(int64_t) * c->S)) return OFG_ERR_CUDA;

    // every parse warp may leave up to PS_RESERVE - 1 reserved slots unused
    const u64 cap = (u64)(total_padded / std::max<int64_t>(c->bytes_per_line_inv, 2)) + 1024 + (u64)parse_grid * PS_WARPS * PS_RESERVE;
    if (cap >= 0xfffffff0ULL) return fail(c, OFG_ERR_CAPACITY, "more than 2^32 records in one context (%llu); shard by query species", (unsigned long long)cap);
    CK(c->coo_row.ensure(cap * 4));
    CK(c->coo_col.ensure(cap * 4));
    CK(c->coo_val.ensure(cap * 8));
    CK(c->rowcount.ensure((size_t)(c->n_rows + 1) * 4));
    CK(c->rowlen.ensure((size_t)(c->n_rows + 1) * 4));
    CK(c->row_start.ensure((size_t)(c->n_rows + 2) * 4));
    // small counters block: [0] coo_count, [1] n_long, [2] err_flags, then pair_lines[P], pair_valid[P], pair_err[P], pair_nnz[P]
    const size_t small_words = 8 + 4 * (size_t)P;
    CK(c->d_small.ensure(small_words * 8));
    CK(cudaMemsetAsync(c->d_small.p, 0, small_words * 8, c->stream));
    u64* sm = c->d_small.as<u64>();
    u64* d_pair_lines = sm + 8;
    u64* d_pair_valid = sm + 8 + P;
    u64* d_pair_err = sm + 8 + 2 * (size_t)P;
    u64* d_pair_nnz = sm + 8 + 3 * (size_t)P;
    CK(cudaMemsetAsync(d_pair_err, 0xff, sizeof(u64) * P, c->stream));
    CK(cudaMemsetAsync(c->rowcount.p, 0, (size_t)(c->n_rows + 1) * 4, c->stream));
    CK(cudaMemsetAsync(c->rowlen.p, 0, (size_t)(c->n_rows + 1) * 4, c->stream));

    CK(cudaEventRecord(c->ev[0], c->stream));
    if (n_tiles > 0) {
        ParseArgs a;
        a.text = c->text.as<uint8_t>(); a.pairs = c->d_pairs.as<PairDesc>(); a.P = P;
        a.seg_prefix = c->d_tile_prefix.as<int64_t>(); a.n_segs = n_tiles;
        a.coo_row = c->coo_row.as<u32>(); a.coo_col = c->coo_col.as<u32>(); a.coo_val = c->coo_val.as<double>();
        a.cap = cap; a.coo_count = sm; a.rowcount = c->rowcount.as<u32>();
        a.pair_lines = d_pair_lines; a.pair_valid = d_pair_valid; a.pair_err = d_pair_err; a.one = 1u;
        const int parse_smem = (int)(sizeof(PsWarp) * PS_WARPS);
        CK(cudaFuncSetAttribute(parse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, parse_smem));
        parse_kernel<<<parse_grid, PARSE_THREADS, parse_smem, c->stream>>>(a);
        CKL("parse_kernel");
    }
    CK(cudaEventRecord(c->ev[1], c->stream));
    // host sync 1: line counts, errors
    std::vector<u64> small(small_words);
    CK(cudaMemcpyAsync(small.data(), sm, small_words * 8, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    mark_timeline(c);
    c->pair_lines.assign(small.begin() + 8, small.begin() + 8 + P);
    c->pair_valid.assign(small.begin() + 8 + P, small.begin() + 8 + 2 * P);
    const u64 coo_count = small[0];
    for (int i = 0; i < P; i++) {
        const u64 e = small[8 + 2 * (size_t)P + i];
        if (e != ~0ULL) {
            std::string d;
            describe_line(c, i, e, &d);
            const int kind = (int)(e & 15);
            const int code = (kind == OFG_LINE_RANGE_Q || kind == OFG_LINE_RANGE_S) ? OFG_ERR_INCONSISTENT
                             : (kind == OFG_LINE_QUOTE || kind == OFG_LINE_SCORE_FORMAT) ? OFG_ERR_UNSUPPORTED : OFG_ERR_PARSE;
            return fail(c, code, "%s", d.c_str());
        }
    }
    if (coo_count > cap)
        return fail(c, OFG_ERR_CAPACITY, "record capacity exceeded: %llu lines but room for %llu (lines shorter than %lld bytes on average); lower option max_bytes_per_line_inv",
                    (unsigned long long)coo_count, (unsigned long long)cap, (long long)c->bytes_per_line_inv);
    c->tot_lines = 0; c->tot_records = 0;
    for (int i = 0; i < P; i++) { c->tot_lines += (int64_t)c->pair_lines[i]; c->tot_records += (int64_t)c->pair_valid[i]; }
    const u64 n_slots = (u64)c->tot_records;
    CK(c->cols.ensure((n_slots + 64) * 4));
    CK(c->vals.ensure((n_slots + 64) * 8));
    CK(c->key.ensure((n_slots + 64) * 4));
    CK(c->keyval.ensure((n_slots + 64) * 8));
    CK(c->long_rows.ensure((size_t)(c->n_rows + 1) * 4));

    if (device_scan<u32>(c, c->rowcount.as<u32>(), c->n_rows, c->row_start.as<u32>(), 1)) return OFG_ERR_CUDA;
    if (coo_count > 0) {
        scatter_rows_kernel<<<grid_for(c, (int64_t)coo_count, 256 * 8, 8), 256, 0, c->stream>>>(
            c->coo_row.as<u32>(), c->coo_col.as<u32>(), c->coo_val.as<double>(), coo_count, c->row_start.as<u32>(),
            c->rowlen.as<u32>(), c->cols.as<u32>(), c->vals.as<double>());
        CKL("scatter_rows_kernel");
    }
    // sentinel key: one bit above the largest possible length product
    double maxL = 1.0;
    for (double v : c->lengths) maxL = std::max(maxL, v);
    const double max_lf = maxL * maxL;
    if (max_lf >= 2147483648.0) return fail(c, OFG_ERR_UNSUPPORTED, "sequence length product %.0f does not fit 31 bits", max_lf);
    int bits = 1;
    while ((double)(1ULL << bits) <= max_lf) bits++;
    const u32 key_sentinel = 1u << bits;
    RowSortArgs r;
    r.n_rows = c->n_rows; r.row_start = c->row_start.as<u32>(); r.rowlen = c->rowlen.as<u32>();
    r.cols = c->cols.as<u32>(); r.vals = c->vals.as<double>(); r.pairs = c->d_pairs.as<PairDesc>();
    r.pair_row_base = c->d_pair_row_base.as<int64_t>(); r.P = P; r.lengths = c->d_lengths.as<double>();
    r.key = c->key.as<u32>(); r.keyval = c->keyval.as<double>(); r.key_sentinel = key_sentinel;
    r.long_rows = c->long_rows.as<u32>(); r.n_long = (u32*)(sm + 1); r.pair_nnz = d_pair_nnz;
    r.scratch_cols = c->coo_col.as<u32>(); r.scratch_vals = c->coo_val.as<double>();
    if (c->n_rows > 0) {
        rowsort_kernel<<<grid_for(c, c->n_rows, ROWSORT_WARPS, 16), ROWSORT_WARPS * 32, 0, c->stream>>>(r);
        CKL("rowsort_kernel");
        rowsort_long_kernel<<<c->n_sm * 2, 256, 0, c->stream>>>(r);
        CKL("rowsort_long_kernel");
    }
    CK(cudaEventRecord(c->ev[2], c->stream));
    // host sync 2: stored entries per pair
    std::vector<u64> nnz(P);
    CK(cudaMemcpyAsync(nnz.data(), d_pair_nnz, sizeof(u64) * P, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    mark_timeline(c);
    c->pair_nnz = nnz;
    c->tot_nnz = 0;
    for (int i = 0; i < P; i++) c->tot_nnz += (int64_t)nnz[i];
    c->stage_done = OFG_STAGE_RAW;
    // remember sort parameters
    c->fit_host.assign((size_t)P * 8, 0.0);
    c->key_bits = bits + 1;  // including the sentinel bit
    return 0;
}

int stage_norm(ofg_ctx* c)
{
    mark_timeline(c);
    const int P = (int)c->pairs.size();
    u64* sm = c->d_small.as<u64>();
    u32* err_flags = (u32*)(sm + 2);
    const int key_bits = c->key_bits;
    const u64 n_slots = (u64)c->tot_records;
    if (c->scores_v2) {
        // --scores-v2 (gathering.py:157-174 NormalisedBitScore): B' = diag(Lq^-0.5) B diag(Lh^-0.5), no fit and no pair is
        // zeroed.  Expressed through the same factor kernels as a fit with a = 0.5, b = 0 (10^-0 = 1 exactly).
        std::vector<double> fv((size_t)P * 8, 0.0);
        for (int i = 0; i < P; i++) { fv[(size_t)i * 8 + 0] = 0.5; fv[(size_t)i * 8 + 3] = 1.0; fv[(size_t)i * 8 + 6] = 1.0; }
        CK(c->fit.ensure((size_t)P * 8 * 8 + 64));
        if (P > 0 && upload(c, c->fit, fv.data(), fv.size() * 8)) return OFG_ERR_CUDA;
        CK(cudaStreamSynchronize(c->stream));  // fv goes out of scope
        CK(cudaEventRecord(c->ev[3], c->stream));
        CK(cudaEventRecord(c->ev[4], c->stream));
        CK(cudaEventRecord(c->ev[5], c->stream));
    } else {
        // ---- segmented stable radix sort by Lf
        std::vector<int64_t> pair_slot(P + 1, 0), tile_prefix(P + 1, 0);
        for (int i = 0; i < P; i++) {
            pair_slot[i + 1] = pair_slot[i] + (int64_t)c->pair_valid[i];
            tile_prefix[i + 1] = tile_prefix[i] + ((int64_t)c->pair_valid[i] + RS_TILE - 1) / RS_TILE;
        }
        const int64_t n_tiles = tile_prefix[P];
        if (upload(c, c->pair_slot, pair_slot.data(), sizeof(int64_t) * (P + 1))) return OFG_ERR_CUDA;
        if (upload(c, c->sort_tile_prefix, tile_prefix.data(), sizeof(int64_t) * (P + 1))) return OFG_ERR_CUDA;
        const int n_pass = (key_bits + RS_MAX_BITS - 1) / RS_MAX_BITS;
        const int pass_bits = (key_bits + n_pass - 1) / n_pass;
        CK(c->key2.ensure((n_slots + 64) * 4));
        CK(c->keyval2.ensure((n_slots + 64) * 8));
        CK(c->counters.ensure(((size_t)n_tiles << pass_bits) * 4 + 64));
        // the first pass reads the scores straight from the CSR values (same slots as the keys), so the row sort does
        // not have to write a second copy; from then on the two sort buffers ping-pong
        u32* kin = c->key.as<u32>(); double* vin = c->vals.as<double>();
        u32* kout = c->key2.as<u32>(); double* vout = c->keyval2.as<double>();
        if (n_tiles > 0) {
            CK(c->sort_desc.ensure((size_t)n_tiles * sizeof(RsTileDesc)));
            for (int pass = 0; pass < n_pass; pass++) {
                RadixArgs a;
                a.key_in = kin; a.val_in = vin; a.key_out = kout; a.val_out = vout;
                a.tile_prefix = c->sort_tile_prefix.as<int64_t>(); a.pair_slot = c->pair_slot.as<int64_t>();
                a.P = P; a.n_tiles = n_tiles; a.shift = pass * pass_bits; a.bits = pass_bits; a.counters = c->counters.as<u32>();
                a.desc = c->sort_desc.as<RsTileDesc>();
                if (pass == 0) {
                    rs_desc_kernel<<<(unsigned)((n_tiles + 255) / 256), 256, 0, c->stream>>>(a, c->sort_desc.as<RsTileDesc>());
                    CKL("rs_desc_kernel");
                }
                const int g = grid_for(c, n_tiles, 1, 2);  // scatter: two resident CTAs per SM (registers, shared memory)
                radix_hist_kernel<<<grid_for(c, n_tiles, 1, 4), RS_THREADS, 0, c->stream>>>(a);
                CKL("radix_hist_kernel");
                if (device_scan<u32>(c, a.counters, n_tiles << pass_bits, a.counters, 0)) return OFG_ERR_CUDA;
                CK(cudaFuncSetAttribute(radix_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RadixSmem)));
                radix_scatter_kernel<<<g, RS_THREADS, sizeof(RadixSmem), c->stream>>>(a);
                CKL("radix_scatter_kernel");
                std::swap(kin, kout);
                if (pass == 0) { vin = vout; vout = c->keyval.as<double>(); } else std::swap(vin, vout);
            }
        }
        CK(cudaEventRecord(c->ev[3], c->stream));
        // ---- bins
        std::vector<BinTask> tasks;
        std::vector<int64_t> task_prefix(P + 1, 0);
        for (int i = 0; i < P; i++) {
            const u64 H = c->pair_nnz[i];
            const u32 first = (u32)pair_slot[i];
            if (H > 0 && H < 100) {
                tasks.push_back(BinTask{first, (u32)H, i, 1});
            } else if (H >= 100) {
                const u32 n_in = H > 5000 ? 1000u : (H > 1000 ? 200u : 20u);
                const u64 n_bins = H / n_in;
                for (u64 b = 0; b < n_bins; b++) tasks.push_back(BinTask{(u32)(first + b * n_in), n_in, i, 0});
            }
            task_prefix[i + 1] = (int64_t)tasks.size();
        }
        const int64_t n_tasks = (int64_t)tasks.size();
        if (upload(c, c->bin_tasks, tasks.data(), sizeof(BinTask) * tasks.size())) return OFG_ERR_CUDA;
        std::vector<u64> tp64(task_prefix.begin(), task_prefix.end());
        if (upload(c, c->task_prefix, tp64.data(), sizeof(u64) * (P + 1))) return OFG_ERR_CUDA;
        CK(c->bin_cut.ensure((size_t)(n_tasks + 1) * 8));
        CK(c->bin_cnt.ensure((size_t)(n_tasks + 1) * 4));
        CK(c->sel_off.ensure((size_t)(n_tasks + 2) * 8));
        c->sel_cap = std::max<u64>(n_slots / (u64)std::max<int64_t>(c->sel_frac_inv, 1), 1u << 16);
        if (c->sel_cap > n_slots + 1) c->sel_cap = n_slots + 1;
        CK(c->lx.ensure(c->sel_cap * 8)); CK(c->ly.ensure(c->sel_cap * 8));
        CK(c->fit.ensure((size_t)P * 8 * 8 + 64));
        CK(cudaMemsetAsync(c->fit.p, 0, (size_t)P * 8 * 8, c->stream));
        BinArgs b;
        b.tasks = c->bin_tasks.as<BinTask>(); b.n_tasks = n_tasks; b.key = kin; b.val = vin;
        {
            // numpy: virtual index (n-1)*q with q = 95/100, gamma = vi - floor(vi)
            const double q = 95.0 / 100.0;
            const double v20 = 19.0 * q, v200 = 199.0 * q, v1000 = 999.0 * q;
            b.k20 = (int)floor(v20); b.g20 = v20 - floor(v20);
            b.k200 = (int)floor(v200); b.g200 = v200 - floor(v200);
            b.k1000 = (int)floor(v1000); b.g1000 = v1000 - floor(v1000);
        }
        b.cut = c->bin_cut.as<double>(); b.cnt = c->bin_cnt.as<u32>(); b.sel_off = c->sel_off.as<u64>();
        b.lx = c->lx.as<double>(); b.ly = c->ly.as<double>(); b.sel_cap = c->sel_cap;
        // staging: two free fp64 arrays of at least n_slots entries - the parse output scores and the idle sort buffer
        b.stage_lx = c->coo_val.as<double>(); b.stage_ly = vout;
        if (n_tasks > 0) {
            const int g = grid_for(c, n_tasks, BIN_WARPS, 6);
            bin_cut_kernel<<<g, BIN_THREADS, 0, c->stream>>>(b);
            CKL("bin_cut_kernel");
            if (device_scan<u64>(c, b.cnt, n_tasks, c->sel_off.as<u64>(), 1)) return OFG_ERR_CUDA;
            bin_select_kernel<<<grid_for(c, n_tasks, BIN_WARPS, 16), BIN_THREADS, 0, c->stream>>>(b);
            CKL("bin_select_kernel");
        } else {
            CK(cudaMemsetAsync(c->sel_off.p, 0, 16, c->stream));
        }
        CK(cudaEventRecord(c->ev[4], c->stream));
        // ---- fit
        CK(c->pair_sel_off.ensure((size_t)(P + 2) * 8));
        gather_u64_kernel<<<(P + 1 + 255) / 256, 256, 0, c->stream>>>(c->sel_off.as<u64>(), c->task_prefix.as<u64>(), P + 1,
                                                                       c->pair_sel_off.as<u64>());
        CKL("gather_u64_kernel");
        FitArgs f;
        f.P = P; f.pair_sel_off = c->pair_sel_off.as<u64>(); f.sel_cap = c->sel_cap; f.lx = b.lx; f.ly = b.ly;
        f.fit = c->fit.as<double>(); f.err_flags = err_flags;
        if (P > 0) {
            CK(cudaFuncSetAttribute(fit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FitShared)));
            fit_kernel<<<P, FIT_THREADS, sizeof(FitShared), c->stream>>>(f);
            CKL("fit_kernel");
        }
        CK(cudaEventRecord(c->ev[5], c->stream));
    }
    // ---- factors, scale, best hits
    CK(c->rowfac.ensure((size_t)(c->n_rows + 1) * 8));
    CK(c->colfac.ensure((size_t)(c->n_cols + 1) * 8));
    CK(c->rowmax.ensure((size_t)(c->n_rows + 1) * 8));
    CK(c->best.ensure((size_t)(c->N + 1) * 8));
    CK(c->most.ensure((size_t)(c->N + 1) * 8));
    CK(cudaMemsetAsync(c->best.p, 0, (size_t)(c->N + 1) * 8, c->stream));
    if (P > 0 && c->n_rows > 0) {
        FactorArgs fa;
        fa.pairs = c->d_pairs.as<PairDesc>(); fa.P = P; fa.fit = c->fit.as<double>(); fa.lengths = c->d_lengths.as<double>();
        fa.rowfac = c->rowfac.as<double>(); fa.colfac = c->colfac.as<double>();
        fa.given = (c->scores_v2 && c->have_length_factors) ? c->d_length_factors.as<double>() : nullptr;
        int64_t maxg = 1;
        for (int i = 0; i < P; i++) maxg = std::max<int64_t>(maxg, (int64_t)c->pairs[i].Gi + c->pairs[i].Gj);
        dim3 g((unsigned)P, (unsigned)std::min<int64_t>((maxg + 255) / 256, 64));
        factors_kernel<<<g, 256, 0, c->stream>>>(fa);
        CKL("factors_kernel");
        ScaleArgs sa;
        sa.n_rows = c->n_rows; sa.row_start = c->row_start.as<u32>(); sa.rowlen = c->rowlen.as<u32>();
        sa.cols = c->cols.as<u32>(); sa.vals = c->vals.as<double>(); sa.pairs = c->d_pairs.as<PairDesc>();
        sa.pair_row_base = c->d_pair_row_base.as<int64_t>(); sa.P = P; sa.fit = c->fit.as<double>();
        sa.rowfac = c->rowfac.as<double>(); sa.colfac = c->colfac.as<double>(); sa.rowmax = c->rowmax.as<double>();
        sa.best = c->best.as<double>();
        CK(c->row_flagcnt.ensure((size_t)(c->n_rows + 1) * 4));
        CK(cudaMemsetAsync(c->row_flagcnt.p, 0, (size_t)(c->n_rows + 1) * 4, c->stream));
        sa.row_flagcnt = c->row_flagcnt.as<u32>();
        const int g2 = grid_for(c, c->n_rows, 16, 8);
        scale_bh_kernel<16><<<g2, 256, 0, c->stream>>>(sa);
        CKL("scale_bh_kernel");
        std::vector<int32_t> diag;
        int64_t max_gi = 1;
        for (int i = 0; i < P; i++) if (c->pairs[i].same) { diag.push_back(i); max_gi = std::max<int64_t>(max_gi, c->pairs[i].Gi); }
        if (!diag.empty()) {
            if (upload(c, c->file_off, diag.data(), diag.size() * 4)) return OFG_ERR_CUDA;  // small scratch buffer
            dim3 gd((unsigned)std::min<int64_t>((max_gi + 7) / 8, 256), (unsigned)diag.size());
            bh_self_kernel<<<gd, 256, 0, c->stream>>>(sa, c->file_off.as<int32_t>());
            CKL("bh_self_kernel");
        }
    }
    CK(cudaEventRecord(c->ev[6], c->stream));
    // host sync 3: fit table, error flags
    c->fit_host.assign((size_t)P * 8, 0.0);
    u32 flags = 0;
    if (P > 0) CK(cudaMemcpyAsync(c->fit_host.data(), c->fit.p, (size_t)P * 8 * 8, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(&flags, err_flags, 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    mark_timeline(c);
    if (flags & 2u)
        return fail(c, OFG_ERR_CAPACITY, "selection capacity exceeded (%llu points); lower option sel_capacity_frac_inv", (unsigned long long)c->sel_cap);
    if (flags & 1u) {
        for (int i = 0; i < P; i++)
            if (c->fit_host[(size_t)i * 8 + 2] > 1 && c->fit_host[(size_t)i * 8 + 6] == 0.0)
                return fail(c, OFG_ERR_FIT, "Optimal parameters not found for pair (%d,%d): MINPACK info %d", c->pairs[i].p, c->pairs[i].j,
                            (int)c->fit_host[(size_t)i * 8 + 3]);
    }
    c->stage_done = OFG_STAGE_NORM;
    return 0;
}

int stage_thresh(ofg_ctx* c)
{
    mark_timeline(c);
    const int P = (int)c->pairs.size();
    // owned gene mask and initial thresholds: 1e9 for owned genes (gathering.py:296), +inf elsewhere.  Both depend
    // only on the species / partition layout: built once and kept on the device.
    if (c->layout_dirty) {
        std::vector<uint8_t> og((size_t)c->N + 1, 0);
        std::vector<double> init((size_t)c->N + 1, INFINITY);
        for (int p : c->owned)
            for (int64_t g = c->seq_start[p]; g < c->seq_start[p] + c->n_seqs[p]; g++) { og[g] = 1; init[g] = 1e9; }
        if (upload(c, c->owned_gene, og.data(), og.size())) return OFG_ERR_CUDA;
        if (upload(c, c->most_init, init.data(), init.size() * 8)) return OFG_ERR_CUDA;
        std::vector<int64_t> gene_of;
        std::vector<int32_t> sp_of;
        for (int p : c->owned)
            for (int64_t q = 0; q < c->n_seqs[p]; q++) { gene_of.push_back(c->seq_start[p] + q); sp_of.push_back(p); }
        if (upload(c, c->gene_of_local, gene_of.data(), gene_of.size() * 8)) return OFG_ERR_CUDA;
        if (upload(c, c->species_of_local, sp_of.data(), sp_of.size() * 4)) return OFG_ERR_CUDA;
        CK(cudaStreamSynchronize(c->stream));  // host vectors go out of scope
        c->layout_dirty = false;
    }
    CK(c->most.ensure((size_t)(c->N + 1) * 8));
    CK(cudaMemcpyAsync(c->most.p, c->most_init.p, (size_t)(c->N + 1) * 8, cudaMemcpyDeviceToDevice, c->stream));
    if (P > 0 && c->n_rows > 0) {
        GraphArgs g;
        g.n_rows = c->n_rows; g.row_start = c->row_start.as<u32>(); g.rowlen = c->rowlen.as<u32>(); g.cols = c->cols.as<u32>();
        g.vals = c->vals.as<double>(); g.pairs = c->d_pairs.as<PairDesc>(); g.pair_row_base = c->d_pair_row_base.as<int64_t>();
        g.P = P; g.pair_of = c->d_pair_of.as<int32_t>(); g.S = c->S; g.most = c->most.as<double>(); g.best = c->best.as<double>(); g.row_flagcnt = nullptr;
        g.zrow = nullptr; g.pair_factor = nullptr; g.gene_cnt = nullptr; g.gene_bytes = nullptr; g.err_flags = nullptr;
        if (c->scores_v2) {
            // gathering.py:314-388 GetMostDistant_s_estimate_from_relative_rbhs
            CK(c->zrow.ensure((size_t)(c->n_rows + 1) * 8));
            CK(cudaMemsetAsync(c->zrow.p, 0, (size_t)(c->n_rows + 1) * 8, c->stream));
            std::vector<double> ones((size_t)P, 1.0);
            std::vector<int32_t> owned(c->owned.begin(), c->owned.end());
            if (upload(c, c->pair_factor, ones.data(), ones.size() * 8)) return OFG_ERR_CUDA;
            if (upload(c, c->d_owned, owned.data(), owned.size() * 4)) return OFG_ERR_CUDA;
            CK(cudaStreamSynchronize(c->stream));  // host vectors go out of scope
            g.zrow = c->zrow.as<double>(); g.pair_factor = c->pair_factor.as<double>();
            rbh_kernel<8, 1><<<grid_for(c, c->n_rows, 32, 8), 256, 0, c->stream>>>(g);
            CKL("rbh_kernel");
            RatioArgs ra;
            ra.pairs = c->d_pairs.as<PairDesc>(); ra.pair_of = c->d_pair_of.as<int32_t>(); ra.S = c->S;
            ra.owned = c->d_owned.as<int32_t>(); ra.n_owned = (int)owned.size(); ra.zrow = c->zrow.as<double>();
            ra.pair_factor = c->pair_factor.as<double>();
            const int64_t n_tasks = (int64_t)owned.size() * c->S * c->S;
            if (n_tasks > 0x7fffffffLL) return fail(c, OFG_ERR_CAPACITY, "too many species for the --scores-v2 ratio table (%lld tasks)", (long long)n_tasks);
            if (n_tasks > 0) {
                v2_ratio_kernel<<<(unsigned)n_tasks, 256, 0, c->stream>>>(ra);
                CKL("v2_ratio_kernel");
            }
            rbh_kernel<8, 2><<<grid_for(c, c->n_rows, 32, 8), 256, 0, c->stream>>>(g);
            CKL("rbh_kernel");
        } else {
            rbh_kernel<2, 0><<<grid_for(c, c->n_rows, 128, 8), 256, 0, c->stream>>>(g);
            CKL("rbh_kernel");
        }
        thresh_final_kernel<<<grid_for(c, c->N, 256, 4), 256, 0, c->stream>>>(c->most.as<double>(), c->best.as<double>(),
                                                                               c->owned_gene.as<uint8_t>(), c->N);
        CKL("thresh_final_kernel");
    }
    CK(cudaEventRecord(c->ev[7], c->stream));
    c->thresholds_complete = ((int)c->owned.size() == c->S);
    c->stage_done = OFG_STAGE_THRESH;
    return 0;
}

int stage_connect(ofg_ctx* c)
{
    mark_timeline(c);
    const int P = (int)c->pairs.size();
    if (!c->thresholds_complete && c->with_columns)
        return fail(c, OFG_ERR_ARG, "thresholds of species owned by other contexts are missing: exchange them and call ofg_thresholds_mark_complete");
    GraphArgs g;
    g.n_rows = c->n_rows; g.row_start = c->row_start.as<u32>(); g.rowlen = c->rowlen.as<u32>(); g.cols = c->cols.as<u32>();
    g.vals = c->vals.as<double>(); g.pairs = c->d_pairs.as<PairDesc>(); g.pair_row_base = c->d_pair_row_base.as<int64_t>();
    g.P = P; g.pair_of = c->d_pair_of.as<int32_t>(); g.S = c->S; g.most = c->most.as<double>(); g.best = c->best.as<double>(); g.row_flagcnt = nullptr; g.zrow = nullptr; g.pair_factor = nullptr; g.gene_cnt = nullptr; g.gene_bytes = nullptr; g.err_flags = nullptr;
    // per-gene edge / text-byte counters, maintained by connect_kernel and by REVERSE imports (edge_account)
    CK(c->gene_cntg.ensure((size_t)(c->N + 1) * 4));
    CK(c->gene_bytesg.ensure((size_t)(c->N + 1) * 4));
    CK(cudaMemsetAsync(c->gene_cntg.p, 0, (size_t)(c->N + 1) * 4, c->stream));
    CK(cudaMemsetAsync(c->gene_bytesg.p, 0, (size_t)(c->N + 1) * 4, c->stream));
    if (P > 0 && c->n_rows > 0) {
        CK(c->row_flagcnt.ensure((size_t)(c->n_rows + 1) * 4));
        CK(cudaMemsetAsync(c->row_flagcnt.p, 0, (size_t)(c->n_rows + 1) * 4, c->stream));
        g.row_flagcnt = c->row_flagcnt.as<u32>();
        g.gene_cnt = c->gene_cntg.as<u32>(); g.gene_bytes = c->gene_bytesg.as<u32>(); g.err_flags = (u32*)(c->d_small.as<u64>() + 2);
        connect_kernel<2><<<grid_for(c, c->n_rows, 128, 8), 256, 0, c->stream>>>(g);
        CKL("connect_kernel");
    }
    CK(cudaEventRecord(c->ev[8], c->stream));
    c->stage_done = OFG_STAGE_CONNECT;
    return 0;
}

int stage_emit(ofg_ctx* c)
{
    mark_timeline(c);
    const int P = (int)c->pairs.size();
    u64* sm = c->d_small.as<u64>();
    u32* err_flags = (u32*)(sm + 2);
    GraphArgs g;
    g.n_rows = c->n_rows; g.row_start = c->row_start.as<u32>(); g.rowlen = c->rowlen.as<u32>(); g.cols = c->cols.as<u32>();
    g.vals = c->vals.as<double>(); g.pairs = c->d_pairs.as<PairDesc>(); g.pair_row_base = c->d_pair_row_base.as<int64_t>();
    g.P = P; g.pair_of = c->d_pair_of.as<int32_t>(); g.S = c->S; g.most = c->most.as<double>(); g.best = c->best.as<double>(); g.row_flagcnt = nullptr; g.zrow = nullptr; g.pair_factor = nullptr; g.gene_cnt = nullptr; g.gene_bytes = nullptr; g.err_flags = nullptr;
    // emission units: owned genes in list order, one per graph row (gene_of_local / species_of_local are layout
    // constants uploaded by stage_thresh)
    const int64_t n_units = c->n_owned_genes;
    CK(c->unit_cnt.ensure((size_t)(n_units + 1) * 4));
    CK(c->unit_bytes.ensure((size_t)(n_units + 1) * 4));
    CK(c->cnt_off.ensure((size_t)(n_units + 2) * 8));
    CK(c->byte_off.ensure((size_t)(n_units + 2) * 8));
    EmitArgs e;
    e.g = g; e.n_genes = n_units; e.gene_of_local = c->gene_of_local.as<int64_t>();
    e.species_of_local = c->species_of_local.as<int32_t>(); e.seq_start = c->d_seq_start.as<int64_t>();
    e.gene_cnt = c->unit_cnt.as<u32>(); e.gene_bytes = c->unit_bytes.as<u32>();
    e.cnt_off = c->cnt_off.as<u64>(); e.byte_off = c->byte_off.as<u64>();
    e.out_indices = nullptr; e.out_data = nullptr; e.out_text = nullptr; e.err_flags = err_flags;
    c->tot_edges = 0; c->tot_text = 0;
    if (n_units > 0) {
        const int ge = grid_for(c, n_units, EMIT_WARPS, 6);
        emit_sizes_kernel<<<grid_for(c, n_units, 256, 4), 256, 0, c->stream>>>(e, c->gene_cntg.as<u32>(), c->gene_bytesg.as<u32>());
        CKL("emit_sizes_kernel");
        if (device_scan<u64>(c, e.gene_cnt, n_units, c->cnt_off.as<u64>(), 1)) return OFG_ERR_CUDA;
        if (device_scan<u64>(c, e.gene_bytes, n_units, c->byte_off.as<u64>(), 1)) return OFG_ERR_CUDA;
        u64 tots[2] = {0, 0};
        CK(cudaMemcpyAsync(&tots[0], c->cnt_off.as<u64>() + n_units, 8, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaMemcpyAsync(&tots[1], c->byte_off.as<u64>() + n_units, 8, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
    mark_timeline(c);  // host sync 4: output sizes
        c->tot_edges = (int64_t)tots[0]; c->tot_text = (int64_t)tots[1];
        CK(c->out_indices.ensure((size_t)(tots[0] + 16) * 4));
        CK(c->out_data.ensure((size_t)(tots[0] + 16) * 8));
        CK(c->out_text.ensure((size_t)tots[1] + 64));
        e.out_indices = c->out_indices.as<int32_t>(); e.out_data = c->out_data.as<double>(); e.out_text = c->out_text.as<char>();
        emit_kernel<true><<<ge, EMIT_WARPS * 32, 0, c->stream>>>(e);
        CKL("emit_kernel<true>");
    }
    CK(cudaEventRecord(c->ev[9], c->stream));
    u32 flags = 0;
    CK(cudaMemcpyAsync(&flags, err_flags, 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    mark_timeline(c);
    if (flags & 4u) return fail(c, OFG_ERR_UNSUPPORTED, "a graph weight is too large for \"%%.3f\" formatting on the device");
    c->stage_done = OFG_STAGE_GRAPH;
    return 0;
}

}  // namespace

// ============================================================================================ C-ABI
extern "C" {

int ofg_create(ofg_ctx** out, int device)
{
    if (!out) return OFG_ERR_ARG;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n) return OFG_ERR_CUDA;
    if (cudaSetDevice(device) != cudaSuccess) return OFG_ERR_CUDA;
    ofg_ctx* c = new ofg_ctx();
    c->device = device;
    cudaDeviceGetAttribute(&c->n_sm, cudaDevAttrMultiProcessorCount, device);
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) { delete c; return OFG_ERR_CUDA; }
    for (auto& e : c->ev) cudaEventCreate(&e);
    *out = c;
    return OFG_OK;
}

void ofg_destroy(ofg_ctx* c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    DevBuf* bufs[] = {&c->text, &c->d_pairs, &c->d_pair_row_base, &c->d_pair_of, &c->d_lengths, &c->d_tile_prefix, &c->d_small,
                      &c->coo_row, &c->coo_col, &c->coo_val, &c->rowcount, &c->row_start, &c->rowlen, &c->cols, &c->vals, &c->key,
                      &c->keyval, &c->key2, &c->keyval2, &c->long_rows, &c->counters, &c->scan_partial, &c->sort_tile_prefix,
                      &c->pair_slot, &c->bin_tasks, &c->bin_cut, &c->bin_cnt, &c->sel_off, &c->task_prefix, &c->lx, &c->ly, &c->sort_desc, &c->zrow, &c->pair_factor, &c->d_owned, &c->gene_cntg, &c->gene_bytesg, &c->most_init, &c->d_length_factors,
                      &c->fit, &c->rowfac, &c->colfac, &c->rowmax, &c->best, &c->most, &c->owned_gene,
                      &c->gene_of_local, &c->species_of_local, &c->d_seq_start, &c->unit_cnt, &c->unit_bytes, &c->cnt_off,
                      &c->byte_off, &c->out_indices, &c->out_data, &c->out_text, &c->species_ids, &c->synth_bytes, &c->synth_off,
                      &c->file_off, &c->pair_sel_off, &c->row_flagcnt, &c->row_flagoff};
    for (DevBuf* b : bufs) b->release();
    for (auto& e : c->ev) if (e) cudaEventDestroy(e);
    cudaStreamDestroy(c->stream);
    delete c;
}

const char* ofg_last_error(const ofg_ctx* c) { return c ? c->err.c_str() : "null context"; }

void* ofg_stream(ofg_ctx* c) { return c ? (void*)c->stream : nullptr; }

int ofg_set_species(ofg_ctx* c, int n_species, const int64_t* n_seqs, const double* lengths_concat)
{
    if (!c || n_species <= 0 || !n_seqs || !lengths_concat) return c ? fail(c, OFG_ERR_ARG, "bad arguments to ofg_set_species") : OFG_ERR_ARG;
    cudaSetDevice(c->device);
    c->S = n_species;
    c->layout_dirty = true;
    c->have_length_factors = false;
    c->n_seqs.assign(n_seqs, n_seqs + n_species);
    c->seq_start.assign(n_species, 0);
    c->len_off.assign(n_species, 0);
    int64_t acc = 0;
    for (int p = 0; p < n_species; p++) {
        if (n_seqs[p] < 0 || n_seqs[p] > (int64_t)OFG_COL_MASK) return fail(c, OFG_ERR_UNSUPPORTED, "species %d has %lld sequences (limit 2^29-1)", p, (long long)n_seqs[p]);
        c->seq_start[p] = acc;
        c->len_off[p] = acc;
        acc += n_seqs[p];
    }
    if (acc >= 2147483647LL) return fail(c, OFG_ERR_UNSUPPORTED, "more than 2^31-1 sequences in total");
    c->N = acc;
    c->lengths.assign(lengths_concat, lengths_concat + acc);
    c->owned.clear();
    c->pairs_dirty = true;
    c->files.clear();
    c->pairs.clear();
    c->text_used = 0;
    c->stage_done = 0;
    return OFG_OK;
}

int ofg_set_partition(ofg_ctx* c, int n_owned, const int32_t* owned_rows, int with_columns)
{
    if (!c) return OFG_ERR_ARG;
    if (c->S <= 0) return fail(c, OFG_ERR_ARG, "call ofg_set_species first");
    c->layout_dirty = true;
    std::vector<int32_t> o;
    for (int i = 0; i < n_owned; i++) {
        if (owned_rows[i] < 0 || owned_rows[i] >= c->S) return fail(c, OFG_ERR_ARG, "owned row %d out of range", owned_rows[i]);
        o.push_back(owned_rows[i]);
    }
    c->owned = o;
    c->with_columns = with_columns;
    c->pairs_dirty = true;
    return ensure_pairs(c);
}

int ofg_set_option(ofg_ctx* c, const char* name, int64_t value)
{
    if (!c || !name) return OFG_ERR_ARG;
    if (!strcmp(name, "allow_empty")) c->allow_empty = (int)value;
    else if (!strcmp(name, "scores_v2")) c->scores_v2 = value != 0;
    else if (!strcmp(name, "max_bytes_per_line_inv")) c->bytes_per_line_inv = value;
    else if (!strcmp(name, "sel_capacity_frac_inv")) c->sel_frac_inv = value;
    else if (!strcmp(name, "reserve_text_bytes")) { cudaSetDevice(c->device); return text_reserve(c, value); }
    else return fail(c, OFG_ERR_ARG, "unknown option %s", name);
    return OFG_OK;
}

int ofg_set_length_factors(ofg_ctx* c, const double* factors_concat)
{
    if (!c || !factors_concat) return OFG_ERR_ARG;
    if (c->S <= 0) return fail(c, OFG_ERR_ARG, "call ofg_set_species first");
    cudaSetDevice(c->device);
    if (upload(c, c->d_length_factors, factors_concat, sizeof(double) * c->lengths.size())) return OFG_ERR_CUDA;
    CK(cudaStreamSynchronize(c->stream));  // the caller's buffer may go away
    c->have_length_factors = true;
    return OFG_OK;
}

int ofg_clear_hits(ofg_ctx* c)
{
    if (!c) return OFG_ERR_ARG;
    for (auto& f : c->files) f = FileInfo();
    c->text_used = 0;
    c->stage_done = 0;
    return OFG_OK;
}

int ofg_add_hits_text(ofg_ctx* c, int p, int j, const void* ptr, size_t nbytes, int swap_cols, int is_device)
{
    if (!c) return OFG_ERR_ARG;
    cudaSetDevice(c->device);
    if (int rc = ensure_pairs(c)) return rc;
    if (p < 0 || j < 0 || p >= c->S || j >= c->S) return fail(c, OFG_ERR_ARG, "pair (%d,%d) out of range", p, j);
    const int pi = c->pair_of[(size_t)p * c->S + j];
    if (pi < 0) return fail(c, OFG_ERR_ARG, "pair (%d,%d) is not part of this context's partition", p, j);
    const int64_t off = (c->text_used + 15) & ~15LL;
    const int64_t padded = ((int64_t)nbytes + 16) & ~15LL;  // always at least one '\n' of padding
    if (int rc = text_reserve(c, off + padded)) return rc;
    uint8_t* dst = c->text.as<uint8_t>() + off;
    if (nbytes) CK(cudaMemcpyAsync(dst, ptr, nbytes, is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemsetAsync(dst + nbytes, '\n', (size_t)(padded - (int64_t)nbytes), c->stream));
    FileInfo& f = c->files[pi];
    f.off = off; f.raw = (int64_t)nbytes; f.padded = padded; f.swap = swap_cols;
    c->text_used = off + padded;
    c->stage_done = 0;
    return OFG_OK;
}

int ofg_synth_hits(ofg_ctx* c, const ofg_synth_params* params, const int32_t* species_ids)
{
    if (!c || !params || !species_ids) return OFG_ERR_ARG;
    cudaSetDevice(c->device);
    if (int rc = ensure_pairs(c)) return rc;
    const int P = (int)c->pairs.size();
    for (auto& f : c->files) f = FileInfo();
    c->text_used = 0;
    if (upload(c, c->d_pairs, c->pairs.data(), sizeof(PairDesc) * P)) return OFG_ERR_CUDA;
    if (upload(c, c->d_pair_row_base, c->pair_row_base.data(), sizeof(int64_t) * (P + 1))) return OFG_ERR_CUDA;
    if (upload(c, c->d_lengths, c->lengths.data(), sizeof(double) * c->lengths.size())) return OFG_ERR_CUDA;
    if (upload(c, c->species_ids, species_ids, sizeof(int32_t) * c->S)) return OFG_ERR_CUDA;
    CK(c->synth_bytes.ensure((size_t)(c->n_rows + 1) * 4));
    CK(c->synth_off.ensure((size_t)(c->n_rows + 2) * 8));
    CK(c->file_off.ensure((size_t)(P + 1) * 8));
    SynthArgs a;
    memcpy(&a.P, params, sizeof(OfgSynthParams));
    a.pairs = c->d_pairs.as<PairDesc>(); a.pair_row_base = c->d_pair_row_base.as<int64_t>(); a.n_pairs = P;
    a.n_rows = c->n_rows; a.species_ids = c->species_ids.as<int32_t>(); a.lengths = c->d_lengths.as<double>();
    a.row_bytes = c->synth_bytes.as<u32>(); a.row_off = c->synth_off.as<u64>(); a.file_off = c->file_off.as<int64_t>();
    a.text = nullptr;
    const int g = grid_for(c, c->n_rows, 128, 16);
    synth_kernel<false><<<g, 128, 0, c->stream>>>(a);
    CKL("synth_kernel<false>");
    if (device_scan<u64>(c, a.row_bytes, c->n_rows, c->synth_off.as<u64>(), 1)) return OFG_ERR_CUDA;
    std::vector<u64> offs((size_t)c->n_rows + 1);
    CK(cudaMemcpyAsync(offs.data(), c->synth_off.p, offs.size() * 8, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    std::vector<int64_t> file_off(P);
    int64_t used = 0;
    for (int i = 0; i < P; i++) {
        const int64_t raw = (int64_t)(offs[c->pair_row_base[i + 1]] - offs[c->pair_row_base[i]]);
        const int64_t off = (used + 15) & ~15LL;
        const int64_t padded = (raw + 16) & ~15LL;
        FileInfo& f = c->files[i];
        f.off = off; f.raw = raw; f.padded = padded; f.swap = 0;
        file_off[i] = off;
        used = off + padded;
    }
    if (int rc = text_reserve(c, used)) return rc;
    CK(cudaMemsetAsync(c->text.p, '\n', (size_t)used + PS_RING + 64, c->stream));
    if (upload(c, c->file_off, file_off.data(), sizeof(int64_t) * P)) return OFG_ERR_CUDA;
    a.text = c->text.as<uint8_t>();
    synth_kernel<true><<<g, 128, 0, c->stream>>>(a);
    CKL("synth_kernel<true>");
    CK(cudaStreamSynchronize(c->stream));
    c->text_used = used;
    c->stage_done = 0;
    return OFG_OK;
}

int ofg_hits_text_size(ofg_ctx* c, int p, int j, size_t* nbytes)
{
    if (!c || !nbytes) return OFG_ERR_ARG;
    if (int rc = ensure_pairs(c)) return rc;
    if (p < 0 || j < 0 || p >= c->S || j >= c->S) return fail(c, OFG_ERR_ARG, "pair (%d,%d) out of range", p, j);
    const int pi = c->pair_of[(size_t)p * c->S + j];
    if (pi < 0 || c->files[pi].off < 0) return fail(c, OFG_ERR_MISSING, "no text for pair (%d,%d)", p, j);
    *nbytes = (size_t)c->files[pi].raw;
    return OFG_OK;
}

int ofg_hits_text_copy(ofg_ctx* c, int p, int j, void* host_dst, size_t cap)
{
    size_t n = 0;
    if (int rc = ofg_hits_text_size(c, p, j, &n)) return rc;
    if (cap < n) return fail(c, OFG_ERR_ARG, "destination too small");
    cudaSetDevice(c->device);
    const int pi = c->pair_of[(size_t)p * c->S + j];
    CK(cudaMemcpyAsync(host_dst, c->text.as<uint8_t>() + c->files[pi].off, n, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return OFG_OK;
}

int ofg_build(ofg_ctx* c, int stop_after)
{
    if (!c) return OFG_ERR_ARG;
    cudaSetDevice(c->device);
    if (int rc = ensure_pairs(c)) return rc;
    if (stop_after < OFG_STAGE_RAW || stop_after > OFG_STAGE_GRAPH) return fail(c, OFG_ERR_ARG, "bad stage %d", stop_after);
    if (c->stage_done >= stop_after) c->stage_done = 0;  // asked for a stage that is already built: rebuild from the text
    const bool fresh = c->stage_done == 0;
    if (fresh) {
        c->tl_n = 0;
        c->launches = 0;
        for (int i = 0; i < 10; i++) c->ms[i] = 0.f;
    }
    int first_ev = -1, last_ev = -1;
    if (c->stage_done < OFG_STAGE_RAW) {
        if (int rc = stage_raw(c)) { c->stage_done = 0; return rc; }
        first_ev = 0; last_ev = 2;
    }
    if (stop_after >= OFG_STAGE_NORM && c->stage_done < OFG_STAGE_NORM) {
        if (first_ev < 0) { CK(cudaEventRecord(c->ev[2], c->stream)); first_ev = 2; }
        if (int rc = stage_norm(c)) { c->stage_done = 0; return rc; }
        last_ev = 6;
    }
    if (stop_after >= OFG_STAGE_THRESH && c->stage_done < OFG_STAGE_THRESH) {
        if (first_ev < 0) { CK(cudaEventRecord(c->ev[6], c->stream)); first_ev = 6; }
        if (int rc = stage_thresh(c)) { c->stage_done = 0; return rc; }
        last_ev = 7;
    }
    if (stop_after >= OFG_STAGE_CONNECT && c->stage_done < OFG_STAGE_CONNECT) {
        if (first_ev < 0) { CK(cudaEventRecord(c->ev[7], c->stream)); first_ev = 7; }
        if (int rc = stage_connect(c)) return rc;
        last_ev = 8;
    }
    if (stop_after >= OFG_STAGE_GRAPH && c->stage_done < OFG_STAGE_GRAPH) {
        if (first_ev < 0) { CK(cudaEventRecord(c->ev[8], c->stream)); first_ev = 8; }
        if (int rc = stage_emit(c)) return rc;
        last_ev = 9;
    }
    CK(cudaStreamSynchronize(c->stream));
    // group timings between consecutive events that were recorded in this call
    if (first_ev >= 0) {
        for (int k = first_ev; k < last_ev; k++) {
            float t = 0.f;
            if (cudaEventElapsedTime(&t, c->ev[k], c->ev[k + 1]) == cudaSuccess) c->ms[k] = t;
        }
        float tot = 0.f;
        for (int k = 0; k < 9; k++) tot += c->ms[k];
        c->ms[9] = tot;
    }
    return OFG_OK;
}

int ofg_kernel_report(ofg_ctx* c, char* buf, size_t cap)
{
    if (!c || !buf || cap == 0) return OFG_ERR_ARG;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    struct Agg { const char* name; int n; double ms; };
    std::vector<Agg> agg;
    for (int i = 1; i < c->tl_n; i++) {
        if (!c->tl_name[i]) continue;
        float t = 0.f;
        if (cudaEventElapsedTime(&t, c->tl_ev[i - 1], c->tl_ev[i]) != cudaSuccess) continue;
        bool found = false;
        for (auto& a : agg) if (!strcmp(a.name, c->tl_name[i])) { a.n++; a.ms += t; found = true; break; }
        if (!found) agg.push_back(Agg{c->tl_name[i], 1, (double)t});
    }
    std::string r;
    char line[256];
    for (auto& a : agg) { snprintf(line, sizeof line, "%s %d %.6f\n", a.name, a.n, a.ms); r += line; }
    if (r.size() + 1 > cap) return fail(c, OFG_ERR_ARG, "report buffer too small");
    memcpy(buf, r.c_str(), r.size() + 1);
    return OFG_OK;
}

int ofg_last_timing(ofg_ctx* c, float ms[10], int64_t* n_launches)
{
    if (!c) return OFG_ERR_ARG;
    if (ms) memcpy(ms, c->ms, sizeof(float) * 10);
    if (n_launches) *n_launches = c->launches;
    return OFG_OK;
}

int ofg_thresholds_dev(ofg_ctx* c, double** dev_ptr, int64_t* n)
{
    if (!c) return OFG_ERR_ARG;
    if (c->stage_done < OFG_STAGE_THRESH) return fail(c, OFG_ERR_ARG, "thresholds are not built yet");
    if (dev_ptr) *dev_ptr = c->most.as<double>();
    if (n) *n = c->N;
    return OFG_OK;
}

int ofg_thresholds_mark_complete(ofg_ctx* c)
{
    if (!c) return OFG_ERR_ARG;
    if (c->stage_done < OFG_STAGE_THRESH) return fail(c, OFG_ERR_ARG, "thresholds are not built yet");
    c->thresholds_complete = true;
    return OFG_OK;
}

int ofg_thresholds_copy(ofg_ctx* c, double* host_dst, int64_t n)
{
    if (!c || !host_dst) return OFG_ERR_ARG;
    if (c->stage_done < OFG_STAGE_THRESH) return fail(c, OFG_ERR_ARG, "thresholds are not built yet");
    cudaSetDevice(c->device);
    CK(cudaMemcpyAsync(host_dst, c->most.p, (size_t)std::min<int64_t>(n, c->N) * 8, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return OFG_OK;
}

int ofg_export_flags(ofg_ctx* c, int which, int n_ranks, const int32_t* species_rank, int my_rank, int64_t* counts, int32_t** dev_list)
{
    if (!c || !species_rank || !counts || !dev_list || n_ranks <= 0) return OFG_ERR_ARG;
    cudaSetDevice(c->device);
    const int need = which == 0 ? OFG_STAGE_NORM : OFG_STAGE_CONNECT;
    if (c->stage_done < need) return fail(c, OFG_ERR_ARG, "export of %s lists needs stage %d", which == 0 ? "best-hit" : "connect", need);
    const int P = (int)c->pairs.size();
    std::vector<uint8_t> ex(P, 0);
    for (int i = 0; i < P; i++) {
        const PairDesc& d = c->pairs[i];
        ex[i] = d.owned_row && species_rank[d.j] != my_rank;  // (k, p): column species p lives elsewhere
    }
    DevBuf& flagbuf = c->synth_bytes;  // scratch buffers of the generator are free at this point
    DevBuf& offbuf = c->synth_off;
    for (int r = 0; r < n_ranks; r++) counts[r] = 0;
    *dev_list = nullptr;
    if (P == 0 || c->n_rows == 0) return OFG_OK;
    if (upload(c, flagbuf, ex.data(), (size_t)P)) return OFG_ERR_CUDA;
    // per-row counts were written by scale_bh_kernel (best hits) / connect_kernel (connect): scan them
    CK(c->row_flagoff.ensure((size_t)(c->n_rows + 2) * 8));
    if (device_scan<u64>(c, c->row_flagcnt.as<u32>(), c->n_rows, c->row_flagoff.as<u64>(), 1)) return OFG_ERR_CUDA;
    CK(offbuf.ensure((size_t)(2 * P + 4) * 8));
    std::vector<u64> bidx(P + 1);
    for (int i = 0; i <= P; i++) bidx[i] = (u64)c->pair_row_base[i];
    CK(cudaMemcpyAsync(offbuf.as<u64>() + P + 2, bidx.data(), sizeof(u64) * (P + 1), cudaMemcpyHostToDevice, c->stream));
    gather_u64_kernel<<<(P + 1 + 255) / 256, 256, 0, c->stream>>>(c->row_flagoff.as<u64>(), offbuf.as<u64>() + P + 2, P + 1, offbuf.as<u64>());
    CKL("gather_u64_kernel");
    std::vector<u64> pb(P + 1);
    CK(cudaMemcpyAsync(pb.data(), offbuf.p, sizeof(u64) * (P + 1), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    // lay the exported pairs out grouped by destination rank
    std::vector<u64> off(P, 0);
    u64 total = 0;
    for (int r = 0; r < n_ranks; r++)
        for (int i = 0; i < P; i++)
            if (ex[i] && species_rank[c->pairs[i].j] == r) { off[i] = total; total += pb[i + 1] - pb[i]; counts[r] += (int64_t)(pb[i + 1] - pb[i]); }
    CK(c->unit_bytes.ensure((size_t)(total + 1) * 8));  // list buffer (free until the emit stage)
    CK(cudaMemcpyAsync(offbuf.p, off.data(), sizeof(u64) * P, cudaMemcpyHostToDevice, c->stream));
    ExportArgs a;
    a.n_rows = c->n_rows; a.row_start = c->row_start.as<u32>(); a.rowlen = c->rowlen.as<u32>(); a.cols = c->cols.as<u32>();
    a.pairs = c->d_pairs.as<PairDesc>(); a.pair_row_base = c->d_pair_row_base.as<int64_t>(); a.P = P;
    a.pair_export = flagbuf.as<uint8_t>(); a.flag = which == 0 ? OFG_FLAG_BH : OFG_FLAG_CONNECT;
    a.row_off = c->row_flagoff.as<u64>(); a.pair_off = offbuf.as<u64>(); a.out = c->unit_bytes.as<int2>();
    if (total > 0) {
        export_flags_kernel<<<grid_for(c, c->n_rows, 8, 8), 256, 0, c->stream>>>(a);
        CKL("export_flags_kernel");
    }
    CK(cudaStreamSynchronize(c->stream));
    *dev_list = c->unit_bytes.as<int32_t>();
    return OFG_OK;
}

int ofg_import_flags(ofg_ctx* c, int which, const int32_t* dev_list, int64_t n_entries)
{
    if (!c || n_entries < 0 || (n_entries > 0 && !dev_list)) return OFG_ERR_ARG;
    cudaSetDevice(c->device);
    const int need = which == 0 ? OFG_STAGE_NORM : OFG_STAGE_CONNECT;
    if (c->stage_done < need) return fail(c, OFG_ERR_ARG, "import of %s lists needs stage %d", which == 0 ? "best-hit" : "connect", need);
    if (n_entries == 0) return OFG_OK;
    u32* err_flags = (u32*)(c->d_small.as<u64>() + 2);
    ImportArgs a;
    a.list = reinterpret_cast<const int2*>(dev_list); a.n = n_entries; a.seq_start = c->d_seq_start.as<int64_t>(); a.S = c->S;
    a.pair_of = c->d_pair_of.as<int32_t>(); a.pairs = c->d_pairs.as<PairDesc>(); a.row_start = c->row_start.as<u32>();
    a.rowlen = c->rowlen.as<u32>(); a.cols = c->cols.as<u32>(); a.set_flag = which == 0 ? OFG_FLAG_TBH : OFG_FLAG_REVERSE;
    a.err_flags = err_flags; a.vals = c->vals.as<double>(); a.gene_cnt = c->gene_cntg.as<u32>(); a.gene_bytes = c->gene_bytesg.as<u32>();
    import_flags_kernel<<<grid_for(c, n_entries, 256, 8), 256, 0, c->stream>>>(a);
    CKL("import_flags_kernel");
    u32 flags = 0;
    CK(cudaMemcpyAsync(&flags, err_flags, 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    if (flags & 8u) return fail(c, OFG_ERR_ARG, "imported list addresses a species pair this context does not hold");
    return OFG_OK;
}

int ofg_counts(ofg_ctx* c, int64_t* lines, int64_t* records, int64_t* nnz, int64_t* edges)
{
    if (!c) return OFG_ERR_ARG;
    if (lines) *lines = c->tot_lines;
    if (records) *records = c->tot_records;
    if (nnz) *nnz = c->tot_nnz;
    if (edges) *edges = c->tot_edges;
    return OFG_OK;
}

int ofg_pair_csr(ofg_ctx* c, int p, int j, int64_t* n_rows, int64_t* nnz, int32_t* row_len, uint32_t* cols, double* vals)
{
    if (!c) return OFG_ERR_ARG;
    if (c->stage_done < OFG_STAGE_RAW) return fail(c, OFG_ERR_ARG, "nothing built yet");
    if (p < 0 || j < 0 || p >= c->S || j >= c->S) return fail(c, OFG_ERR_ARG, "pair (%d,%d) out of range", p, j);
    const int pi = c->pair_of[(size_t)p * c->S + j];
    if (pi < 0) return fail(c, OFG_ERR_ARG, "pair (%d,%d) is not part of this context", p, j);
    cudaSetDevice(c->device);
    const PairDesc& pd = c->pairs[pi];
    const int64_t G = pd.Gi;
    std::vector<u32> rl((size_t)G), rs((size_t)G + 1);
    CK(cudaMemcpyAsync(rl.data(), c->rowlen.as<u32>() + pd.row_base, (size_t)G * 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(rs.data(), c->row_start.as<u32>() + pd.row_base, (size_t)(G + 1) * 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    int64_t tot = 0;
    for (int64_t r = 0; r < G; r++) tot += rl[r];
    if (n_rows) *n_rows = G;
    if (nnz) *nnz = tot;
    if (row_len) for (int64_t r = 0; r < G; r++) row_len[r] = (int32_t)rl[r];
    if (cols || vals) {
        const u32 s0 = rs[0], s1 = rs[G];
        std::vector<u32> hc((size_t)(s1 - s0) + 1);
        std::vector<double> hv((size_t)(s1 - s0) + 1);
        if (s1 > s0) {
            CK(cudaMemcpyAsync(hc.data(), c->cols.as<u32>() + s0, (size_t)(s1 - s0) * 4, cudaMemcpyDeviceToHost, c->stream));
            CK(cudaMemcpyAsync(hv.data(), c->vals.as<double>() + s0, (size_t)(s1 - s0) * 8, cudaMemcpyDeviceToHost, c->stream));
            CK(cudaStreamSynchronize(c->stream));
        }
        int64_t o = 0;
        for (int64_t r = 0; r < G; r++)
            for (u32 k = 0; k < rl[r]; k++, o++) {
                if (cols) cols[o] = hc[rs[r] - s0 + k];
                if (vals) vals[o] = hv[rs[r] - s0 + k];
            }
    }
    return OFG_OK;
}

int ofg_pair_fit(ofg_ctx* c, int p, int j, double out[6])
{
    if (!c || !out) return OFG_ERR_ARG;
    if (c->stage_done < OFG_STAGE_NORM) return fail(c, OFG_ERR_ARG, "normalisation not built yet");
    if (p < 0 || j < 0 || p >= c->S || j >= c->S) return fail(c, OFG_ERR_ARG, "pair (%d,%d) out of range", p, j);
    const int pi = c->pair_of[(size_t)p * c->S + j];
    if (pi < 0) return fail(c, OFG_ERR_ARG, "pair (%d,%d) is not part of this context", p, j);
    const double* f = c->fit_host.data() + (size_t)pi * 8;
    out[0] = f[0]; out[1] = f[1]; out[2] = f[2]; out[3] = f[3]; out[4] = f[4]; out[5] = f[6];
    return OFG_OK;
}

int ofg_graph_csr_size(ofg_ctx* c, int64_t* n_rows, int64_t* nnz)
{
    if (!c) return OFG_ERR_ARG;
    if (c->stage_done < OFG_STAGE_GRAPH) return fail(c, OFG_ERR_ARG, "graph not built yet");
    if (n_rows) *n_rows = c->n_owned_genes;
    if (nnz) *nnz = c->tot_edges;
    return OFG_OK;
}

int ofg_graph_csr_copy(ofg_ctx* c, int64_t* row_ids, int64_t* indptr, int32_t* indices, double* data)
{
    if (!c) return OFG_ERR_ARG;
    if (c->stage_done < OFG_STAGE_GRAPH) return fail(c, OFG_ERR_ARG, "graph not built yet");
    cudaSetDevice(c->device);
    const int64_t n = c->n_owned_genes;
    if (row_ids) {
        int64_t o = 0;
        for (int p : c->owned) for (int64_t q = 0; q < c->n_seqs[p]; q++) row_ids[o++] = c->seq_start[p] + q;
    }
    if (indptr) {
        std::vector<u64> off((size_t)n + 1);
        if (n > 0) CK(cudaMemcpyAsync(off.data(), c->cnt_off.p, off.size() * 8, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        for (int64_t g = 0; g < n; g++) indptr[g] = (int64_t)off[(size_t)g];
        indptr[n] = c->tot_edges;
    }
    if (indices && c->tot_edges) CK(cudaMemcpyAsync(indices, c->out_indices.p, (size_t)c->tot_edges * 4, cudaMemcpyDeviceToHost, c->stream));
    if (data && c->tot_edges) CK(cudaMemcpyAsync(data, c->out_data.p, (size_t)c->tot_edges * 8, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return OFG_OK;
}

int ofg_graph_text_size(ofg_ctx* c, size_t* nbytes)
{
    if (!c || !nbytes) return OFG_ERR_ARG;
    if (c->stage_done < OFG_STAGE_GRAPH) return fail(c, OFG_ERR_ARG, "graph not built yet");
    *nbytes = (size_t)c->tot_text;
    return OFG_OK;
}

int ofg_graph_text_copy(ofg_ctx* c, void* host_dst, size_t cap)
{
    if (!c || !host_dst) return OFG_ERR_ARG;
    if (c->stage_done < OFG_STAGE_GRAPH) return fail(c, OFG_ERR_ARG, "graph not built yet");
    if (cap < (size_t)c->tot_text) return fail(c, OFG_ERR_ARG, "destination too small");
    cudaSetDevice(c->device);
    if (c->tot_text) CK(cudaMemcpyAsync(host_dst, c->out_text.p, (size_t)c->tot_text, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return OFG_OK;
}

// ---------------------------------------------------------------- unit-test entry points
int ofg_test_log10(ofg_ctx* c, const double* host_x, double* host_y, int64_t n)
{
    if (!c || n <= 0) return OFG_ERR_ARG;
    cudaSetDevice(c->device);
    DevBuf x, y;
    CK(x.ensure((size_t)n * 8)); CK(y.ensure((size_t)n * 8));
    CK(cudaMemcpyAsync(x.p, host_x, (size_t)n * 8, cudaMemcpyHostToDevice, c->stream));
    test_log10_kernel<<<grid_for(c, n, 256, 8), 256, 0, c->stream>>>(x.as<double>(), y.as<double>(), n);
    CKL("test_log10_kernel");
    CK(cudaMemcpyAsync(host_y, y.p, (size_t)n * 8, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    x.release(); y.release();
    return OFG_OK;
}

int ofg_test_format(ofg_ctx* c, const double* host_x, char* host_out, int64_t n)
{
    if (!c || n <= 0) return OFG_ERR_ARG;
    cudaSetDevice(c->device);
    DevBuf x, o;
    CK(x.ensure((size_t)n * 8)); CK(o.ensure((size_t)n * 32));
    CK(cudaMemcpyAsync(x.p, host_x, (size_t)n * 8, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemsetAsync(o.p, 0, (size_t)n * 32, c->stream));
    test_format_kernel<<<grid_for(c, n, 256, 8), 256, 0, c->stream>>>(x.as<double>(), o.as<char>(), n);
    CKL("test_format_kernel");
    CK(cudaMemcpyAsync(host_out, o.p, (size_t)n * 32, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    x.release(); o.release();
    return OFG_OK;
}

int ofg_test_parse_score(ofg_ctx* c, const char* host_text, int64_t nbytes, double* host_y, int32_t* host_ok, int64_t n_fields)
{
    if (!c || n_fields <= 0) return OFG_ERR_ARG;
    cudaSetDevice(c->device);
    std::vector<int64_t> starts;
    starts.push_back(0);
    for (int64_t i = 0; i < nbytes && (int64_t)starts.size() < n_fields; i++) if (host_text[i] == '\n') starts.push_back(i + 1);
    if ((int64_t)starts.size() != n_fields) return fail(c, OFG_ERR_ARG, "expected %lld fields", (long long)n_fields);
    DevBuf t, s, y, k;
    CK(t.ensure((size_t)nbytes + 16)); CK(s.ensure((size_t)n_fields * 8)); CK(y.ensure((size_t)n_fields * 8)); CK(k.ensure((size_t)n_fields * 4));
    CK(cudaMemcpyAsync(t.p, host_text, (size_t)nbytes, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(s.p, starts.data(), (size_t)n_fields * 8, cudaMemcpyHostToDevice, c->stream));
    test_score_kernel<<<grid_for(c, n_fields, 256, 8), 256, 0, c->stream>>>(t.as<uint8_t>(), nbytes, s.as<int64_t>(), n_fields, y.as<double>(), k.as<int32_t>());
    CKL("test_score_kernel");
    CK(cudaMemcpyAsync(host_y, y.p, (size_t)n_fields * 8, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(host_ok, k.p, (size_t)n_fields * 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    t.release(); s.release(); y.release(); k.release();
    return OFG_OK;
}

int ofg_test_lmfit(ofg_ctx* c, const double* host_lx, const double* host_ly, int64_t m, double out[4])
{
    if (!c || m < 2) return OFG_ERR_ARG;
    cudaSetDevice(c->device);
    DevBuf x, y, o;
    CK(x.ensure((size_t)m * 8)); CK(y.ensure((size_t)m * 8)); CK(o.ensure(64));
    CK(cudaMemcpyAsync(x.p, host_lx, (size_t)m * 8, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(y.p, host_ly, (size_t)m * 8, cudaMemcpyHostToDevice, c->stream));
    CK(cudaFuncSetAttribute(test_fit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FitShared)));
    test_fit_kernel<<<1, FIT_THREADS, sizeof(FitShared), c->stream>>>(x.as<double>(), y.as<double>(), (long)m, o.as<double>());
    CKL("test_fit_kernel");
    double h[8];
    CK(cudaMemcpyAsync(h, o.p, 48, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    out[0] = h[0]; out[1] = h[1]; out[2] = h[3]; out[3] = h[4];
    x.release(); y.release(); o.release();
    return OFG_OK;
}

}  // extern "C"
