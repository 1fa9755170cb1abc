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
#include "ofg_bins.cuh"
#include "ofg_graph.cuh"
#include "ofg_synth.cuh"
#include "ofg_inflate.cuh"
#include "ofg_dendro.cuh"

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

// One wave = a contiguous range of species pairs (whole query species) whose RAW + NORM stages run together out of
// one set of scratch buffers; only the slotted CSR and the per-row / per-gene vectors persist across waves (the
// reference processes one query species at a time for the same reason: gathering.py:484-485).
struct Wave {
    int P0 = 0, P1 = 0;       // pairs [P0, P1)
    int64_t R0 = 0, R1 = 0;   // rows  [R0, R1)
    int64_t text_bytes = 0;   // padded text of the wave
    int64_t n_segs = 0;       // parse segments
    u64 cap = 0;              // record-slot capacity (upper bound from the text size)
};

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
    bool has_partition = false;  // ofg_set_partition was called (an empty owned list then means "no rows", not "all rows")
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
    bool homology = false;      // option homology_graph: `--c-homologs` (gathering.py:51-73, 520-521) instead of the connect / RBH graph
    bool keep_self = false;     // option keep_self_hits: qExcludeSelfHits=False of GetBLAST6Scores (DendroBLAST, orthologues.py:272-274); RAW stage only
    int rowsort_short = -1;     // option rowsort_short: -1 = by the wave's text bytes per row, 0 = never, 1 = always use the half-warp sort for rows <= 16 records
    bool radix_only = false;    // option length_sort_radix: always take the LSD radix path (the fallback of the bucket path)
    int64_t fallback_waves = 0; // waves of the last build that took the radix path
    int wave_species = 0;       // option wave_species: query species per wave (0: the whole partition in one wave)
    bool async = false;         // option async: ofg_build returns after enqueueing; results are read back by ofg_sync
    double bpl_learned = 0.0;   // bytes of text per record slot seen by the last complete build (sizes the next one)
    // text arena
    DevBuf text;
    int64_t text_used = 0;
    // state
    int stage_done = 0;
    bool thresholds_complete = false;
    bool pairs_dirty = true;
    bool pending = false;       // device results of the last build(s) have not been read back yet
    bool graph_pending = false; // ... including the graph sizes
    std::vector<Wave> waves;
    std::vector<int64_t> seg_prefix_host;   // parse segments before each pair, per wave (kept alive for the async uploads)
    std::vector<int32_t> diag_host;
    bool raw_scratch_valid = false;         // the wave scratch still holds the RAW output of the (only) wave
    // device buffers
    DevBuf d_pairs, d_pair_row_base, d_pair_of, d_lengths, d_tile_prefix, d_small;
    DevBuf coo_row, coo_col, coo_val, rowcount, row_start, rowlen, cols, vals, key, keyval, key2, keyval2;
    DevBuf long_rows, counters, scan_partial, wave_tile_prefix, wave_pair_slot, bin_tasks, bin_cut, bin_cnt, sel_off, wave_task_prefix;
    DevBuf sort_desc, zrow, pair_factor, d_owned, gene_cntg, gene_bytesg, most_init, d_length_factors, lx, ly, fit, rowfac, colfac, rowmax, best, most, owned_gene;
    DevBuf gene_of_local, species_of_local, d_seq_start, unit_cnt, unit_bytes, cnt_off, byte_off, out_indices, out_data, out_text;
    DevBuf species_ids, synth_bytes, synth_off, file_off, pair_sel_off, row_flagcnt, row_flagoff, diag_pairs;
    DevBuf hom_cnt, hom_off, hom_fill, hom_col;
    // gzip input: compressed bytes of the hit files and their member table; inflated on the device at the start of a build
    DevBuf gz_comp, gz_members_dev, gz_err;
    int64_t gz_used = 0;
    std::vector<GzMember> gz_members;
    bool gz_pending = false;      // members were added since the last inflate
    unsigned long long* h_gz_err = nullptr;
    DevBuf bk_hist, bk_rec, bk_bbidx, bk_shist, bk_thr, bk_below, bk_bin16, bk_bb, bk_bbfill, bk_split, bk_gcursor, bin_val;
    DevBuf x_flag, x_off, x_list, x_pair_dst, x_order, x_dst_first, x_inbox_ptrs;
    // pinned read-back area: [small block][fit table]
    u64* h_small = nullptr;
    size_t h_small_cap = 0;
    double* h_fit = nullptr;
    size_t h_fit_cap = 0;
    size_t small_words = 0, small_pair_base = 0;
    // results
    std::vector<u64> pair_lines, pair_valid, pair_nnz;
    std::vector<double> fit_host;
    int64_t tot_lines = 0, tot_records = 0, tot_nnz = 0, tot_edges = 0, tot_text = 0, n_owned_genes = 0;
    u64 sel_cap = 0;
    int key_bits = 32;
    float ms[10] = {0};
    int64_t launches = 0;
    std::vector<cudaEvent_t> tl_ev;
    std::vector<const char*> tl_name;
    int tl_n = 0;
    EmitArgs last_emit;
    int last_emit_grid = 0;
    HomEmitArgs last_hom;
    bool last_is_hom = false;
    std::vector<int64_t> x_sig;             // what the exchange tables on the device were built from
    std::vector<unsigned char> plan_sig;    // what the descriptor tables on the device were built from
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

// per-kernel timeline: one event after every launch (cheap), resolved when the results are read back
void note_launch(ofg_ctx* c, const char* name)
{
    if (c->tl_n >= (int)c->tl_ev.size()) {
        if (c->tl_ev.size() >= 8192) return;
        cudaEvent_t e = nullptr;
        cudaEventCreate(&e);
        c->tl_ev.push_back(e);
        c->tl_name.push_back(nullptr);
    }
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

// offsets of every pair's selected points: pair_sel_off[i] = sel_off[task_prefix[i]] (both on the device)
__global__ void pair_sel_off_kernel(const u64* sel_off, const u64* task_prefix, int n, u64* out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = sel_off[task_prefix[i]];
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
// Exclusive scan of n u32 values.  n is a host-side bound; the real count may live on the device (n_dev / n_shift,
// see scan_count).  base_dev (nullable): device value added to every output.  write_total: out[n] = total.
template <typename TO>
int device_scan(ofg_ctx* c, const u32* in, int64_t n, TO* out, int write_total, const int64_t* n_dev = nullptr, int n_shift = 0,
                const TO* base_dev = nullptr)
{
    if (n <= 0) {
        if (write_total) {
            if (base_dev) CK(cudaMemcpyAsync(out, base_dev, sizeof(TO), cudaMemcpyDeviceToDevice, c->stream));
            else CK(cudaMemsetAsync(out, 0, sizeof(TO), c->stream));
        }
        return 0;
    }
    const int64_t n_chunks = (n + SCAN_CHUNK - 1) / SCAN_CHUNK;
    CK(c->scan_partial.ensure((size_t)(n_chunks + 1) * sizeof(u64)));
    TO* part = c->scan_partial.as<TO>();
    const int g = grid_for(c, n_chunks, 1, 8);
    scan_reduce_kernel<TO><<<g, SCAN_THREADS, 0, c->stream>>>(in, n, part, n_dev, n_shift);
    CKL("scan_reduce_kernel<TO>");
    scan_partials_kernel<TO><<<1, 1024, 0, c->stream>>>(part, n, (TO*)nullptr, n_dev, n_shift, base_dev);
    CKL("scan_partials_kernel<TO>");
    scan_apply_kernel<TO><<<g, SCAN_THREADS, 0, c->stream>>>(in, n, part, out, write_total, n_dev, n_shift, base_dev);
    CKL("scan_apply_kernel<TO>");
    return 0;
}

int rebuild_pairs(ofg_ctx* c)
{
    if (c->S <= 0) return fail(c, OFG_ERR_ARG, "ofg_set_species has not been called");
    std::vector<int32_t> owned = c->owned;
    if (owned.empty() && !c->has_partition) for (int p = 0; p < c->S; p++) owned.push_back(p);
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

// ------------------------------------------------------------------------------------------ plan
u64* small_ptr(ofg_ctx* c) { return c->d_small.as<u64>(); }
u32* err_flags_ptr(ofg_ctx* c) { return (u32*)(small_ptr(c) + 2); }
int64_t* n_tiles_ptr(ofg_ctx* c) { return (int64_t*)(small_ptr(c) + 3); }
int64_t* n_tasks_ptr(ofg_ctx* c) { return (int64_t*)(small_ptr(c) + 4); }
u32* bk_flags_ptr(ofg_ctx* c) { return (u32*)(small_ptr(c) + 5); }
int64_t* n_tiles_fb_ptr(ofg_ctx* c) { return (int64_t*)(small_ptr(c) + 6); }
int64_t* n_tasks_fb_ptr(ofg_ctx* c) { return (int64_t*)(small_ptr(c) + 7); }
int64_t* n_tiles_nw_ptr(ofg_ctx* c) { return (int64_t*)(small_ptr(c) + 8); }
int64_t* n_tasks_nw_ptr(ofg_ctx* c) { return (int64_t*)(small_ptr(c) + 9); }
u32* n_bb_ptr(ofg_ctx* c) { return (u32*)(small_ptr(c) + 10); }
unsigned long long* list_used_ptr(ofg_ctx* c) { return (unsigned long long*)(small_ptr(c) + 11); }
u64* fallback_waves_ptr(ofg_ctx* c) { return small_ptr(c) + 12; }   // waves that took the radix path (statistics)
u64* wave_count_ptr(ofg_ctx* c, int w) { return small_ptr(c) + 16 + w; }
u64* pair_lines_ptr(ofg_ctx* c) { return small_ptr(c) + c->small_pair_base; }
u64* pair_valid_ptr(ofg_ctx* c) { return small_ptr(c) + c->small_pair_base + c->pairs.size(); }
u64* pair_err_ptr(ofg_ctx* c) { return small_ptr(c) + c->small_pair_base + 2 * c->pairs.size(); }
u64* pair_nnz_ptr(ofg_ctx* c) { return small_ptr(c) + c->small_pair_base + 3 * c->pairs.size(); }

// bytes of text that one record slot is assumed to need at least (capacity planning)
double bytes_per_slot(ofg_ctx* c)
{
    return std::max<double>((double)std::max<int64_t>(c->bytes_per_line_inv, 2), c->bpl_learned);
}

// Waves, text offsets and capacities; uploads the descriptors.  Everything here depends only on the host-side
// description of the job (file sizes), never on device results.
int plan_build(ofg_ctx* c)
{
    const int P = (int)c->pairs.size();
    for (int i = 0; i < P; i++) {
        FileInfo& f = c->files[i];
        PairDesc& d = c->pairs[i];
        if (f.off < 0) {
            if (!c->allow_empty) return fail(c, OFG_ERR_MISSING, "hit text for pair (%d,%d) was not provided", d.p, d.j);
            d.present = 0; d.text_off = 0; d.text_len = 0; d.text_raw = 0; d.swap = 0;
        } else {
            d.present = 1; d.text_off = f.off; d.text_len = f.padded; d.text_raw = f.raw; d.swap = f.swap;
        }
    }
    // waves: whole query species (S consecutive pairs each), wave_species of them at a time
    c->waves.clear();
    const int S = c->S;
    const int n_groups = S > 0 ? (P + S - 1) / S : 0;   // row groups: owned species, then (design A) the column pairs in chunks of S
    const int per = c->wave_species > 0 ? c->wave_species : std::max(n_groups, 1);
    c->seg_prefix_host.clear();
    const double bps = bytes_per_slot(c);
    for (int g0 = 0; g0 < n_groups; g0 += per) {
        Wave w;
        w.P0 = g0 * S;
        w.P1 = std::min(P, (g0 + per) * S);
        w.R0 = c->pair_row_base[w.P0];
        w.R1 = c->pair_row_base[w.P1];
        for (int i = w.P0; i < w.P1; i++) {
            w.text_bytes += c->pairs[i].text_len;
            w.n_segs += (c->pairs[i].text_len + PS_SEG - 1) / PS_SEG;
        }
        const int parse_grid = grid_for(c, (w.n_segs + PS_WARPS - 1) / PS_WARPS, 1, 4);
        // every parse warp may leave up to PS_RESERVE - 1 reserved slots unused
        w.cap = (u64)((double)w.text_bytes / bps) + 1024 + (u64)parse_grid * PS_WARPS * PS_RESERVE;
        c->waves.push_back(w);
    }
    if (c->waves.empty()) c->waves.push_back(Wave());
    u64 tot_cap = 0;
    for (const Wave& w : c->waves) tot_cap += w.cap;
    if (tot_cap >= 0xfffffff0ULL)
        return fail(c, OFG_ERR_CAPACITY, "more than 2^32 record slots in one context (%llu); shard by query species or raise max_bytes_per_line_inv", (unsigned long long)tot_cap);
    // parse segments before each pair, per wave, back to back: [wave 0: Pw + 1 entries][wave 1: ...]
    for (const Wave& w : c->waves) {
        int64_t acc = 0;
        for (int i = w.P0; i < w.P1; i++) {
            c->seg_prefix_host.push_back(acc);
            acc += (c->pairs[i].text_len + PS_SEG - 1) / PS_SEG;
        }
        c->seg_prefix_host.push_back(acc);
    }
    {
        // descriptor tables: uploaded only when they changed (a job that is built repeatedly enqueues without host stalls)
        std::vector<unsigned char> sig;
        auto put = [&](const void* p, size_t n) { const unsigned char* b = (const unsigned char*)p; sig.insert(sig.end(), b, b + n); };
        put(c->pairs.data(), sizeof(PairDesc) * P);
        put(c->pair_of.data(), sizeof(int32_t) * c->pair_of.size());
        put(c->seg_prefix_host.data(), sizeof(int64_t) * c->seg_prefix_host.size());
        put(c->seq_start.data(), sizeof(int64_t) * c->S);
        if (sig != c->plan_sig || c->layout_dirty) {
            if (upload(c, c->d_pairs, c->pairs.data(), sizeof(PairDesc) * P)) return OFG_ERR_CUDA;
            if (upload(c, c->d_pair_row_base, c->pair_row_base.data(), sizeof(int64_t) * (P + 1))) return OFG_ERR_CUDA;
            if (upload(c, c->d_pair_of, c->pair_of.data(), sizeof(int32_t) * c->pair_of.size())) return OFG_ERR_CUDA;
            if (upload(c, c->d_lengths, c->lengths.data(), sizeof(double) * c->lengths.size())) return OFG_ERR_CUDA;
            if (upload(c, c->d_tile_prefix, c->seg_prefix_host.data(), sizeof(int64_t) * c->seg_prefix_host.size())) return OFG_ERR_CUDA;
            if (upload(c, c->d_seq_start, c->seq_start.data(), sizeof(int64_t) * c->S)) return OFG_ERR_CUDA;
            c->plan_sig.swap(sig);
        }
    }
    // small counters block: [1] n_long, [2] err_flags, [3] n_tiles, [4] n_tasks, [5..12] bucket-path routing, [16 + w] record
    // slots reserved by wave w, then pair_lines[P], pair_valid[P], pair_err[P], pair_nnz[P]
    c->small_pair_base = 16 + ((c->waves.size() + 7) & ~(size_t)7);
    c->small_words = c->small_pair_base + 4 * (size_t)P;
    CK(c->d_small.ensure(c->small_words * 8));
    CK(cudaMemsetAsync(c->d_small.p, 0, c->small_words * 8, c->stream));
    if (P > 0) CK(cudaMemsetAsync(pair_err_ptr(c), 0xff, sizeof(u64) * P, c->stream));
    // persistent per-row / per-slot arrays
    CK(c->rowcount.ensure((size_t)(c->n_rows + 1) * 4));
    CK(c->rowlen.ensure((size_t)(c->n_rows + 1) * 4));
    CK(c->row_start.ensure((size_t)(c->n_rows + 2) * 4));
    CK(c->long_rows.ensure((size_t)(c->n_rows + 1) * 4));
    CK(cudaMemsetAsync(c->rowcount.p, 0, (size_t)(c->n_rows + 1) * 4, c->stream));
    CK(cudaMemsetAsync(c->rowlen.p, 0, (size_t)(c->n_rows + 1) * 4, c->stream));
    CK(cudaMemsetAsync(c->row_start.p, 0, 4, c->stream));   // the first wave starts at slot 0
    CK(c->cols.ensure((tot_cap + 64) * 4));
    CK(c->vals.ensure((tot_cap + 64) * 8));
    CK(c->fit.ensure((size_t)P * 8 * 8 + 64));
    CK(cudaMemsetAsync(c->fit.p, 0, (size_t)P * 8 * 8 + 64, c->stream));
    // sentinel key: one bit above the largest possible length product
    double maxL = 1.0;
    for (double v : c->lengths) maxL = std::max(maxL, v);
    const double max_lf = maxL * maxL;
    if (max_lf >= 2147483648.0) return fail(c, OFG_ERR_UNSUPPORTED, "sequence length product %.0f does not fit 31 bits", max_lf);
    int bits = 1;
    while ((double)(1ULL << bits) <= max_lf) bits++;
    c->key_bits = bits + 1;  // including the sentinel bit
    // pinned read-back area
    if (c->h_small_cap < c->small_words) {
        if (c->h_small) cudaFreeHost(c->h_small);
        c->h_small = nullptr;
        CK(cudaHostAlloc((void**)&c->h_small, (c->small_words + 64) * 8, cudaHostAllocDefault));
        c->h_small_cap = c->small_words + 64;
    }
    if (c->h_fit_cap < (size_t)P * 8 + 16) {
        if (c->h_fit) cudaFreeHost(c->h_fit);
    if (c->h_gz_err) cudaFreeHost(c->h_gz_err);
        c->h_fit = nullptr;
        CK(cudaHostAlloc((void**)&c->h_fit, ((size_t)P * 8 + 16 + 64) * 8, cudaHostAllocDefault));
        c->h_fit_cap = (size_t)P * 8 + 16 + 64;
    }
    c->tot_edges = 0; c->tot_text = 0;
    return 0;
}

// ------------------------------------------------------------------------------------------ stages, per wave
int wave_raw(ofg_ctx* c, int wi)
{
    const Wave& w = c->waves[wi];
    const int Pw = w.P1 - w.P0;
    const int64_t nr = w.R1 - w.R0;
    if (Pw <= 0) return 0;
    size_t seg_off = 0;  // position of this wave's segment prefix inside d_tile_prefix
    for (int k = 0; k < wi; k++) seg_off += (size_t)(c->waves[k].P1 - c->waves[k].P0) + 1;
    const u64 cap = w.cap;
    CK(c->coo_row.ensure(cap * 4));
    CK(c->coo_col.ensure(cap * 4));
    CK(c->coo_val.ensure(cap * 8));
    CK(c->key.ensure((cap + 64) * 4));
    u64* d_count = wave_count_ptr(c, wi);
    if (w.n_segs > 0) {
        const int parse_grid = grid_for(c, (w.n_segs + PS_WARPS - 1) / PS_WARPS, 1, 4);
        ParseArgs a;
        a.text = c->text.as<uint8_t>(); a.pairs = c->d_pairs.as<PairDesc>() + w.P0; a.P = Pw;
        a.seg_prefix = c->d_tile_prefix.as<int64_t>() + seg_off; a.n_segs = w.n_segs;
        a.coo_row = c->coo_row.as<u32>(); a.coo_col = c->coo_col.as<u32>(); a.coo_val = c->coo_val.as<double>();
        a.cap = cap; a.coo_count = d_count; a.rowcount = c->rowcount.as<u32>();
        a.pair_lines = pair_lines_ptr(c) + w.P0; a.pair_valid = pair_valid_ptr(c) + w.P0; a.pair_err = pair_err_ptr(c) + w.P0; a.one = 1u; a.keep_self = c->keep_self ? 1 : 0;
        const int parse_smem = (int)(sizeof(PsWarp) * PS_WARPS);
        CK(cudaFuncSetAttribute(parse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, parse_smem));
        parse_kernel<<<parse_grid, PARSE_THREADS, parse_smem, c->stream>>>(a);
        CKL("parse_kernel");
    }
    // row_start of the wave's rows: exclusive scan of the row histogram, starting at the slot where the previous wave ended
    u32* rs = c->row_start.as<u32>() + w.R0;
    if (device_scan<u32>(c, c->rowcount.as<u32>() + w.R0, nr, rs, 1, nullptr, 0, rs)) return OFG_ERR_CUDA;
    scatter_rows_kernel<<<grid_for(c, (int64_t)cap, 256 * 8, 8), 256, 0, c->stream>>>(
        c->coo_row.as<u32>(), c->coo_col.as<u32>(), c->coo_val.as<double>(), d_count, cap, c->row_start.as<u32>(),
        c->rowlen.as<u32>(), c->cols.as<u32>(), c->vals.as<double>());
    CKL("scatter_rows_kernel");
    CK(cudaMemsetAsync(small_ptr(c) + 1, 0, 8, c->stream));  // n_long of this wave
    RowSortArgs r;
    r.row_begin = w.R0; r.n_rows = w.R1; r.row_start = c->row_start.as<u32>(); r.rowlen = c->rowlen.as<u32>();
    r.cols = c->cols.as<u32>(); r.vals = c->vals.as<double>(); r.pairs = c->d_pairs.as<PairDesc>();
    r.pair_row_base = c->d_pair_row_base.as<int64_t>(); r.P = (int)c->pairs.size(); r.lengths = c->d_lengths.as<double>();
    r.key = c->key.as<u32>(); r.slot0 = rs; r.keyval = nullptr; r.key_sentinel = 1u << (c->key_bits - 1);
    r.long_rows = c->long_rows.as<u32>(); r.n_long = (u32*)(small_ptr(c) + 1); r.pair_nnz = pair_nnz_ptr(c);
    r.scratch_cols = c->coo_col.as<u32>(); r.scratch_vals = c->coo_val.as<double>();
    // rows of a few records (jobs with a handful of hits per query and target species) go through the half-warp kernel first
    const bool short_rows = c->rowsort_short >= 0 ? c->rowsort_short != 0 : (nr > 0 && (double)w.text_bytes / 60.0 / (double)nr < 32.0);
    r.short_max = short_rows ? ROWSORT_SHORT : 0;
    if (nr > 0) {
        if (short_rows) {
            rowsort_short_kernel<<<grid_for(c, nr, 64, 8), 256, 0, c->stream>>>(r);
            CKL("rowsort_short_kernel");
        }
        rowsort_kernel<<<grid_for(c, nr, ROWSORT_WARPS, 16), ROWSORT_WARPS * 32, 0, c->stream>>>(r);
        CKL("rowsort_kernel");
        rowsort_long_kernel<<<c->n_sm * 2, 256, 0, c->stream>>>(r);
        CKL("rowsort_long_kernel");
    }
    return 0;
}

int wave_norm(ofg_ctx* c, int wi)
{
    const Wave& w = c->waves[wi];
    const int Pw = w.P1 - w.P0;
    const int64_t nr = w.R1 - w.R0;
    if (Pw <= 0) return 0;
    const int P = (int)c->pairs.size();
    u32* err_flags = err_flags_ptr(c);
    const int key_bits = c->key_bits;
    const u64 cap = w.cap;
    const u32* slot0 = c->row_start.as<u32>() + w.R0;
    double* fit_w = c->fit.as<double>() + (size_t)w.P0 * 8;
    if (c->scores_v2) {
        // --scores-v2 (gathering.py:157-174 NormalisedBitScore): B' = diag(Lq^-0.5) B diag(Lh^-0.5), no fit and no pair is
        // zeroed.  Expressed through the same factor kernels as a fit with a = 0.5, b = 0 (10^-0 = 1 exactly).
        double* fv = c->h_fit;  // pinned staging (also the read-back area; rewritten by the read-back later)
        for (int i = 0; i < Pw; i++) {
            double* f = fv + (size_t)(w.P0 + i) * 8;
            for (int k = 0; k < 8; k++) f[k] = 0.0;
            f[0] = 0.5; f[3] = 1.0; f[6] = 1.0;
        }
        CK(cudaMemcpyAsync(fit_w, fv + (size_t)w.P0 * 8, (size_t)Pw * 8 * 8, cudaMemcpyHostToDevice, c->stream));
    } else {
        // ---- layout of the wave on the device: slots, sort tiles and length bins of every pair
        const int64_t max_tiles = (int64_t)(cap / RS_TILE) + Pw + 1;
        // bins: H / 1000 for H > 5000; at most 25 (H <= 5000, bins of 200) or 50 (H <= 1000, bins of 20) or 1 per smaller pair
        const int64_t max_tasks = (int64_t)(cap / 1000) + 76 * (int64_t)Pw + 16;
        CK(c->wave_pair_slot.ensure((size_t)(Pw + 2) * 8));
        CK(c->wave_tile_prefix.ensure((size_t)(Pw + 2) * 8));
        CK(c->wave_task_prefix.ensure((size_t)(Pw + 2) * 8));
        WaveLayoutArgs la;
        la.row_start = c->row_start.as<u32>(); la.pair_row_base = c->d_pair_row_base.as<int64_t>(); la.P0 = w.P0; la.P1 = w.P1;
        la.pair_nnz = pair_nnz_ptr(c); la.pair_slot = c->wave_pair_slot.as<int64_t>(); la.tile_prefix = c->wave_tile_prefix.as<int64_t>();
        la.task_prefix = c->wave_task_prefix.as<u64>(); la.n_tiles = n_tiles_ptr(c); la.n_tasks = n_tasks_ptr(c);
        la.max_tiles = max_tiles; la.max_tasks = max_tasks; la.err_flags = err_flags;
        wave_layout_kernel<<<1, 1024, 0, c->stream>>>(la, 1);
        CKL("wave_layout_kernel");
        const int n_pass = (key_bits + RS_MAX_BITS - 1) / RS_MAX_BITS;     // of the radix path
        const int pass_bits = (key_bits + n_pass - 1) / n_pass;
        CK(c->sort_desc.ensure((size_t)max_tiles * sizeof(RsTileDesc)));
        rs_desc_kernel<<<grid_for(c, max_tiles, 256, 4), 256, 0, c->stream>>>(c->wave_tile_prefix.as<int64_t>(), c->wave_pair_slot.as<int64_t>(), Pw,
                                                                              n_tiles_ptr(c), pass_bits, c->sort_desc.as<RsTileDesc>());
        CKL("rs_desc_kernel");
        CK(c->keyval.ensure((cap + 64) * 8));
        CK(c->key2.ensure((cap + 64) * 4));
        CK(c->keyval2.ensure((cap + 64) * 8));
        CK(c->bin_val.ensure((cap + 64) * 8));
        CK(c->bin_tasks.ensure((size_t)(max_tasks + 1) * sizeof(BinTask)));
        CK(c->bin_cut.ensure((size_t)(max_tasks + 1) * 8));
        CK(c->bin_cnt.ensure((size_t)(max_tasks + 1) * 4));
        CK(c->sel_off.ensure((size_t)(max_tasks + 2) * 8));
        c->sel_cap = std::max<u64>(cap / (u64)std::max<int64_t>(c->sel_frac_inv, 1), 1u << 16);
        CK(c->lx.ensure(c->sel_cap * 8)); CK(c->ly.ensure(c->sel_cap * 8));
        bin_tasks_kernel<<<grid_for(c, max_tasks, 256, 4), 256, 0, c->stream>>>(c->wave_task_prefix.as<u64>(), c->wave_pair_slot.as<int64_t>(), pair_nnz_ptr(c),
                                                                                w.P0, Pw, n_tasks_ptr(c), c->bin_tasks.as<BinTask>());
        CKL("bin_tasks_kernel");
        BinArgs b;
        b.tasks = c->bin_tasks.as<BinTask>(); b.n_tasks_dev = n_tasks_ptr(c); b.key = nullptr; b.val = nullptr; b.rec = nullptr; b.cand_cnt = nullptr; b.below = nullptr;
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
        // staging of the selected hits of every bin (bin_cut -> bin_select): two fp64 arrays of at least `cap` entries
        b.stage_lx = c->coo_val.as<double>(); b.stage_ly = c->bin_val.as<double>();
        // ---- bucket path (ofg_bins.cuh): bins without a full sort; routes itself to the radix path on the device
        CK(cudaMemsetAsync(small_ptr(c) + 5, 0, 7 * 8, c->stream));   // routing words of this wave
        if (c->radix_only) {
            const u32 one = BK_FALLBACK;
            CK(cudaMemcpyAsync(bk_flags_ptr(c), &one, 4, cudaMemcpyHostToDevice, c->stream));
        }
        const size_t tab = (size_t)Pw * BK_NB * 4;
        CK(c->bk_hist.ensure(tab)); CK(c->bk_bbidx.ensure(tab));
        CK(c->bk_rec.ensure((cap + 64) * 16));
        CK(c->bk_bb.ensure((size_t)(max_tasks + 1) * sizeof(BbDesc)));
        CK(c->bk_bbfill.ensure((size_t)(max_tasks + 1) * 4));
        CK(c->bk_split.ensure((size_t)(max_tasks + 1) * 8));
        CK(c->bk_gcursor.ensure((size_t)(max_tasks + 1) * 4));
        CK(c->bk_shist.ensure((size_t)(max_tasks + 1) * (BK_SB / 2) * 4));
        CK(c->bk_thr.ensure((size_t)(max_tasks + 1) * 2));
        CK(c->bk_below.ensure((size_t)(max_tasks + 1) * 2));
        CK(c->bk_bin16.ensure((cap + 64) * 2));
        BkArgs k;
        k.key = c->key.as<u32>(); k.key_sentinel = 1u << (key_bits - 1); k.vals = c->vals.as<double>(); k.slot0 = slot0;
        k.desc = c->sort_desc.as<RsTileDesc>(); k.n_tiles_all = n_tiles_ptr(c); k.n_tiles_nw = n_tiles_nw_ptr(c);
        k.pair_slot = c->wave_pair_slot.as<int64_t>(); k.task_prefix = c->wave_task_prefix.as<u64>(); k.pair_nnz = pair_nnz_ptr(c) + w.P0; k.Pw = Pw;
        k.hist = c->bk_hist.as<u32>(); k.binof = c->bk_bbidx.as<u32>();
        k.bb = c->bk_bb.as<BbDesc>(); k.bb_cap = (u32)std::min<int64_t>(max_tasks, 0x7fffffff); k.n_bb = n_bb_ptr(c); k.list_used = list_used_ptr(c);
        k.list_cap = cap; k.list = c->keyval.as<u64>(); k.bbfill = c->bk_bbfill.as<u32>(); k.split = c->bk_split.as<u64>();
        k.gcursor = c->bk_gcursor.as<u32>(); k.rec_b = c->bk_rec.as<uint4>(); k.flags = bk_flags_ptr(c);
        if (!c->radix_only) {
            CK(cudaMemsetAsync(c->bk_hist.p, 0, tab, c->stream));
            CK(cudaMemsetAsync(c->bk_bbidx.p, 0, tab, c->stream));
            CK(cudaMemsetAsync(c->bk_bbfill.p, 0, (size_t)(max_tasks + 1) * 4, c->stream));
            CK(cudaMemsetAsync(c->bk_gcursor.p, 0, (size_t)(max_tasks + 1) * 4, c->stream));
            CK(cudaMemsetAsync(c->bk_shist.p, 0, (size_t)(max_tasks + 1) * (BK_SB / 2) * 4, c->stream));
            CK(cudaFuncSetAttribute(bk_hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BK_NB * 2));
            bk_hist_kernel<<<grid_for(c, max_tiles, BKH_TILES, 1), BKH_THREADS, BK_NB * 2, c->stream>>>(k);
            CKL("bk_hist_kernel");
            bk_scan_kernel<<<Pw, 1024, 0, c->stream>>>(k);
            CKL("bk_scan_kernel");
        }
        bk_route_kernel<<<1, 1, 0, c->stream>>>(bk_flags_ptr(c), n_tiles_ptr(c), n_tasks_ptr(c), n_bb_ptr(c), k.bb_cap, n_tiles_fb_ptr(c), n_tasks_fb_ptr(c),
                                                n_tiles_nw_ptr(c), n_tasks_nw_ptr(c), fallback_waves_ptr(c));
        CKL("bk_route_kernel");
        if (!c->radix_only) {
            bk_gather_kernel<<<grid_for(c, max_tiles, 1, 4), 512, 0, c->stream>>>(k);
            CKL("bk_gather_kernel");
            bk_split_kernel<<<grid_for(c, max_tasks, BKS_WARPS, 8), BKS_WARPS * 32, 0, c->stream>>>(k);
            CKL("bk_split_kernel");
            BkCandArgs kc;
            kc.bin16 = c->bk_bin16.as<unsigned short>(); kc.shist = c->bk_shist.as<u32>(); kc.thr = c->bk_thr.as<unsigned short>();
            kc.below = c->bk_below.as<unsigned short>(); kc.n_tasks_nw = n_tasks_nw_ptr(c); kc.tasks = c->bin_tasks.as<BinTask>();
            kc.k20 = b.k20; kc.k200 = b.k200; kc.k1000 = b.k1000;
            bk_binhist_kernel<<<grid_for(c, max_tiles, 1, 4), 512, 0, c->stream>>>(k, kc);
            CKL("bk_binhist_kernel");
            bk_thr_kernel<<<grid_for(c, max_tasks, 8, 8), 256, 0, c->stream>>>(kc);
            CKL("bk_thr_kernel");
            bk_cand_kernel<<<grid_for(c, max_tiles, 1, 3), 512, 0, c->stream>>>(k, kc);
            CKL("bk_cand_kernel");
            BinArgs bu = b;
            bu.n_tasks_dev = n_tasks_nw_ptr(c); bu.key = nullptr; bu.val = nullptr; bu.rec = c->bk_rec.as<uint4>();
            bu.cand_cnt = c->bk_gcursor.as<u32>(); bu.below = c->bk_below.as<unsigned short>();
            bin_cut_kernel<true><<<grid_for(c, max_tasks, BIN_WARPS, 6), BIN_THREADS, 0, c->stream>>>(bu);
            CKL("bin_cut_kernel");
        }
        // ---- radix path (fallback): segmented stable LSD radix sort by Lf, then the bins of the sorted order.  Its tile and
        // task counts are zero unless the bucket path routed the wave here.
        {
            CK(c->counters.ensure(((size_t)max_tiles << pass_bits) * 4 + 64));
            // the first pass reads the scores straight from the CSR values (same slots as the keys), so the row sort does
            // not have to write a second copy; from then on the two sort buffers ping-pong
            u32* kin = c->key.as<u32>(); double* vin = c->vals.as<double>();
            u32* kout = c->key2.as<u32>(); double* vout = c->keyval2.as<double>();
            CK(cudaFuncSetAttribute(radix_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RadixSmem)));
            // a wave that stays on the bucket path only pays for the (empty) launches; keep their grids small unless forced
            const int gh = c->radix_only ? grid_for(c, max_tiles, 1, 4) : std::min(grid_for(c, max_tiles, 1, 4), c->n_sm * 4);
            for (int pass = 0; pass < n_pass; pass++) {
                RadixArgs a;
                a.key_in = kin; a.val_in = vin; a.key_out = kout; a.val_out = vout;
                a.n_tiles_dev = n_tiles_fb_ptr(c); a.val_in_off_dev_valid = pass == 0 ? 1 : 0; a.slot0 = slot0;
                a.shift = pass * pass_bits; a.bits = pass_bits; a.counters = c->counters.as<u32>();
                a.desc = c->sort_desc.as<RsTileDesc>();
                radix_hist_kernel<<<gh, RS_THREADS, 0, c->stream>>>(a);
                CKL("radix_hist_kernel");
                if (device_scan<u32>(c, a.counters, max_tiles << pass_bits, a.counters, 0, n_tiles_fb_ptr(c), pass_bits)) return OFG_ERR_CUDA;
                radix_scatter_kernel<<<grid_for(c, max_tiles, 1, 2), RS_THREADS, sizeof(RadixSmem), c->stream>>>(a);  // two resident CTAs per SM
                CKL("radix_scatter_kernel");
                std::swap(kin, kout);
                if (pass == 0) { vin = vout; vout = c->keyval.as<double>(); } else std::swap(vin, vout);
            }
            BinArgs bs = b;
            bs.n_tasks_dev = n_tasks_fb_ptr(c); bs.key = kin; bs.val = vin; bs.rec = nullptr;
            bin_cut_kernel<false><<<grid_for(c, max_tasks, BIN_WARPS, 6), BIN_THREADS, 0, c->stream>>>(bs);
            CKL("bin_cut_kernel");
        }
        if (device_scan<u64>(c, b.cnt, max_tasks, c->sel_off.as<u64>(), 1, n_tasks_ptr(c), 0)) return OFG_ERR_CUDA;
        bin_select_kernel<<<grid_for(c, max_tasks, BIN_WARPS, 16), BIN_THREADS, 0, c->stream>>>(b);
        CKL("bin_select_kernel");
        // ---- fit
        CK(c->pair_sel_off.ensure((size_t)(Pw + 2) * 8));
        pair_sel_off_kernel<<<(Pw + 1 + 255) / 256, 256, 0, c->stream>>>(c->sel_off.as<u64>(), c->wave_task_prefix.as<u64>(), Pw + 1,
                                                                         c->pair_sel_off.as<u64>());
        CKL("pair_sel_off_kernel");
        FitArgs f;
        f.P = Pw; f.pair_sel_off = c->pair_sel_off.as<u64>(); f.sel_cap = c->sel_cap; f.lx = b.lx; f.ly = b.ly;
        f.fit = fit_w; f.err_flags = err_flags;
        CK(ofg_launch_fit(f, Pw, c->stream));
        CKL("fit_kernel");
    }
    // ---- factors, scale, best hits
    if (nr > 0) {
        FactorArgs fa;
        fa.pairs = c->d_pairs.as<PairDesc>() + w.P0; fa.P = Pw; fa.fit = fit_w; fa.lengths = c->d_lengths.as<double>();
        fa.rowfac = c->rowfac.as<double>(); fa.colfac = c->colfac.as<double>();
        fa.given = (c->scores_v2 && c->have_length_factors) ? c->d_length_factors.as<double>() : nullptr;
        int64_t maxg = 1;
        for (int i = w.P0; i < w.P1; i++) maxg = std::max<int64_t>(maxg, (int64_t)c->pairs[i].Gi + c->pairs[i].Gj);
        dim3 g((unsigned)Pw, (unsigned)std::min<int64_t>((maxg + 255) / 256, 64));
        factors_kernel<<<g, 256, 0, c->stream>>>(fa);
        CKL("factors_kernel");
        ScaleArgs sa;
        sa.row_begin = w.R0; sa.n_rows = w.R1; sa.row_start = c->row_start.as<u32>(); sa.rowlen = c->rowlen.as<u32>();
        sa.cols = c->cols.as<u32>(); sa.vals = c->vals.as<double>(); sa.pairs = c->d_pairs.as<PairDesc>();
        sa.pair_row_base = c->d_pair_row_base.as<int64_t>(); sa.P = P; sa.fit = c->fit.as<double>();
        sa.rowfac = c->rowfac.as<double>(); sa.colfac = c->colfac.as<double>(); sa.rowmax = c->rowmax.as<double>();
        sa.best = c->best.as<double>();
        sa.row_flagcnt = c->row_flagcnt.as<u32>();
        // lanes per row by the wave's records per row (estimated from its text size: ~60 bytes per hit line)
        if ((double)w.text_bytes / 60.0 / (double)nr < 14.0) scale_bh_kernel<8><<<grid_for(c, nr, 32, 8), 256, 0, c->stream>>>(sa);
        else scale_bh_kernel<16><<<grid_for(c, nr, 16, 8), 256, 0, c->stream>>>(sa);
        CKL("scale_bh_kernel");
    }
    return 0;
}

// buffers of the NORM stage that span all waves
int norm_begin(ofg_ctx* c)
{
    CK(c->rowfac.ensure((size_t)(c->n_rows + 1) * 8));
    CK(c->colfac.ensure((size_t)(c->n_cols + 1) * 8));
    CK(c->rowmax.ensure((size_t)(c->n_rows + 1) * 8));
    CK(c->best.ensure((size_t)(c->N + 1) * 8));
    CK(c->most.ensure((size_t)(c->N + 1) * 8));
    CK(cudaMemsetAsync(c->best.p, 0, (size_t)(c->N + 1) * 8, c->stream));
    CK(c->row_flagcnt.ensure((size_t)(c->n_rows + 1) * 4));
    CK(cudaMemsetAsync(c->row_flagcnt.p, 0, (size_t)(c->n_rows + 1) * 4, c->stream));
    return 0;
}

// best hits of the within-species matrices: needs the per-gene best of every other species, i.e. all waves
int norm_finish(ofg_ctx* c)
{
    const int P = (int)c->pairs.size();
    c->diag_host.clear();
    int64_t max_gi = 1;
    for (int i = 0; i < P; i++) if (c->pairs[i].same) { c->diag_host.push_back(i); max_gi = std::max<int64_t>(max_gi, c->pairs[i].Gi); }
    if (!c->diag_host.empty() && c->n_rows > 0) {
        if (upload(c, c->diag_pairs, c->diag_host.data(), c->diag_host.size() * 4)) return OFG_ERR_CUDA;
        ScaleArgs sa;
        sa.row_begin = 0; sa.n_rows = c->n_rows; sa.row_start = c->row_start.as<u32>(); sa.rowlen = c->rowlen.as<u32>();
        sa.cols = c->cols.as<u32>(); sa.vals = c->vals.as<double>(); sa.pairs = c->d_pairs.as<PairDesc>();
        sa.pair_row_base = c->d_pair_row_base.as<int64_t>(); sa.P = P; sa.fit = c->fit.as<double>();
        sa.rowfac = c->rowfac.as<double>(); sa.colfac = c->colfac.as<double>(); sa.rowmax = c->rowmax.as<double>();
        sa.best = c->best.as<double>(); sa.row_flagcnt = c->row_flagcnt.as<u32>();
        dim3 gd((unsigned)std::min<int64_t>((max_gi + 7) / 8, 256), (unsigned)c->diag_host.size());
        bh_self_kernel<<<gd, 256, 0, c->stream>>>(sa, c->diag_pairs.as<int32_t>());
        CKL("bh_self_kernel");
    }
    return 0;
}

// gzip members added since the last build: inflated into the text arena, one warp per member
int inflate_pending(ofg_ctx* c)
{
    if (!c->gz_pending) return 0;
    const int64_t n = (int64_t)c->gz_members.size();
    CK(c->gz_err.ensure(16));
    CK(cudaMemsetAsync(c->gz_err.p, 0xff, 8, c->stream));
    if (n > 0) {
        if (upload(c, c->gz_members_dev, c->gz_members.data(), sizeof(GzMember) * (size_t)n)) return OFG_ERR_CUDA;
        inflate_kernel<<<grid_for(c, n, GZ_WARPS, 8), GZ_WARPS * 32, 0, c->stream>>>(c->gz_comp.as<uint8_t>(), c->gz_members_dev.as<GzMember>(), n,
                                                                                     c->text.as<uint8_t>(), c->gz_err.as<unsigned long long>());
        CKL("inflate_kernel");
    }
    c->gz_pending = false;
    return 0;
}

int stage_raw_norm(ofg_ctx* c, int stop_after)
{
    mark_timeline(c);
    if (int rc = inflate_pending(c)) return rc;
    if (int rc = plan_build(c)) return rc;
    const bool norm = stop_after >= OFG_STAGE_NORM;
    if (norm) if (int rc = norm_begin(c)) return rc;
    for (int wi = 0; wi < (int)c->waves.size(); wi++) {
        if (int rc = wave_raw(c, wi)) return rc;
        if (norm) if (int rc = wave_norm(c, wi)) return rc;
    }
    if (norm) if (int rc = norm_finish(c)) return rc;
    c->raw_scratch_valid = !norm && c->waves.size() == 1;
    c->stage_done = norm ? OFG_STAGE_NORM : OFG_STAGE_RAW;
    c->pending = true;
    return 0;
}

// NORM on top of a finished RAW stage (single wave only: the wave scratch still holds the sort keys)
int stage_norm_only(ofg_ctx* c)
{
    mark_timeline(c);
    if (int rc = norm_begin(c)) return rc;
    if (int rc = wave_norm(c, 0)) return rc;
    if (int rc = norm_finish(c)) return rc;
    c->raw_scratch_valid = false;
    c->stage_done = OFG_STAGE_NORM;
    c->pending = true;
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
        CK(cudaStreamSynchronize(c->stream));  // host vectors go out of scope (once per species / partition layout)
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
    c->thresholds_complete = ((int)c->owned.size() == c->S);
    c->stage_done = OFG_STAGE_THRESH;
    c->pending = true;
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
        g.gene_cnt = c->gene_cntg.as<u32>(); g.gene_bytes = c->gene_bytesg.as<u32>(); g.err_flags = err_flags_ptr(c);
        connect_kernel<2><<<grid_for(c, c->n_rows, 128, 8), 256, 0, c->stream>>>(g);
        CKL("connect_kernel");
    }
    c->stage_done = OFG_STAGE_CONNECT;
    c->pending = true;
    return 0;
}

int launch_emit(ofg_ctx* c)
{
    if (c->last_is_hom) {
        HomEmitArgs& h = c->last_hom;
        h.out_indices = c->out_indices.as<int32_t>(); h.out_data = c->out_data.as<double>(); h.out_text = c->out_text.as<char>();
        h.cap_edges = c->out_indices.cap / 4 > 16 ? c->out_indices.cap / 4 - 16 : 0;
        h.cap_edges = std::min<u64>(h.cap_edges, c->out_data.cap / 8 > 16 ? c->out_data.cap / 8 - 16 : 0);
        h.cap_text = c->out_text.cap > 64 ? c->out_text.cap - 64 : 0;
        hom_emit_kernel<true><<<c->last_emit_grid, HOM_WARPS * 32, 0, c->stream>>>(h);
        CKL("hom_emit_kernel<true>");
        return 0;
    }
    EmitArgs& e = c->last_emit;
    e.out_indices = c->out_indices.as<int32_t>(); e.out_data = c->out_data.as<double>(); e.out_text = c->out_text.as<char>();
    e.cap_edges = c->out_indices.cap / 4 > 16 ? c->out_indices.cap / 4 - 16 : 0;
    e.cap_edges = std::min<u64>(e.cap_edges, c->out_data.cap / 8 > 16 ? c->out_data.cap / 8 - 16 : 0);
    e.cap_text = c->out_text.cap > 64 ? c->out_text.cap - 64 : 0;
    emit_kernel<true><<<c->last_emit_grid, EMIT_WARPS * 32, 0, c->stream>>>(e);
    CKL("emit_kernel<true>");
    return 0;
}

int stage_emit(ofg_ctx* c)
{
    mark_timeline(c);
    const int P = (int)c->pairs.size();
    u32* err_flags = err_flags_ptr(c);
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
    e.cap_edges = 0; e.cap_text = 0;
    c->last_emit = e;
    c->last_is_hom = false;
    c->last_emit_grid = 0;
    c->tot_edges = 0; c->tot_text = 0;
    if (n_units > 0) {
        c->last_emit_grid = grid_for(c, n_units, EMIT_WARPS, 6);
        emit_sizes_kernel<<<grid_for(c, n_units, 256, 4), 256, 0, c->stream>>>(e, c->gene_cntg.as<u32>(), c->gene_bytesg.as<u32>());
        CKL("emit_sizes_kernel");
        if (device_scan<u64>(c, e.gene_cnt, n_units, c->cnt_off.as<u64>(), 1)) return OFG_ERR_CUDA;
        if (device_scan<u64>(c, e.gene_bytes, n_units, c->byte_off.as<u64>(), 1)) return OFG_ERR_CUDA;
        // the output sizes are known only on the device: emit into the buffers of the previous build and let the kernel
        // check the capacity; the read-back grows them and emits again when they were too small (first build)
        if (int rc = launch_emit(c)) return rc;
    }
    c->stage_done = OFG_STAGE_GRAPH;
    c->pending = true;
    c->graph_pending = true;
    return 0;
}

// `--c-homologs` (gathering.py:51-73, :520-524): the graph is the symmetrised structure of the normalised matrices; replaces
// the THRESH / CONNECT / GRAPH stages.  Needs the whole pair grid in this context.
int stage_homology(ofg_ctx* c)
{
    mark_timeline(c);
    const int P = (int)c->pairs.size();
    if ((int)c->owned.size() != c->S) return fail(c, OFG_ERR_UNSUPPORTED, "the homology graph needs every query species in one context");
    if (c->layout_dirty) {
        std::vector<int64_t> gene_of;
        std::vector<int32_t> sp_of;
        for (int p : c->owned)
            for (int64_t q = 0; q < c->n_seqs[p]; q++) { gene_of.push_back(c->seq_start[p] + q); sp_of.push_back(p); }
        if (upload(c, c->gene_of_local, gene_of.data(), gene_of.size() * 8)) return OFG_ERR_CUDA;
        if (upload(c, c->species_of_local, sp_of.data(), sp_of.size() * 4)) return OFG_ERR_CUDA;
        CK(cudaStreamSynchronize(c->stream));
        // (layout_dirty stays set: stage_thresh builds the rest of the layout vectors when the default graph is asked for)
    }
    GraphArgs g;
    g.n_rows = c->n_rows; g.row_start = c->row_start.as<u32>(); g.rowlen = c->rowlen.as<u32>(); g.cols = c->cols.as<u32>();
    g.vals = c->vals.as<double>(); g.pairs = c->d_pairs.as<PairDesc>(); g.pair_row_base = c->d_pair_row_base.as<int64_t>();
    g.P = P; g.pair_of = c->d_pair_of.as<int32_t>(); g.S = c->S; g.most = nullptr; g.best = nullptr; g.row_flagcnt = nullptr; g.zrow = nullptr; g.pair_factor = nullptr; g.gene_cnt = nullptr; g.gene_bytes = nullptr; g.err_flags = nullptr;
    CK(c->hom_cnt.ensure((size_t)(c->n_rows + 1) * 4));
    CK(c->hom_off.ensure((size_t)(c->n_rows + 2) * 4));
    CK(c->hom_fill.ensure((size_t)(c->n_rows + 1) * 4));
    CK(cudaMemsetAsync(c->hom_cnt.p, 0, (size_t)(c->n_rows + 1) * 4, c->stream));
    CK(cudaMemsetAsync(c->hom_fill.p, 0, (size_t)(c->n_rows + 1) * 4, c->stream));
    // extras are entries of the transposed direction: never more than the stored entries
    u64 tot_cap = 0;
    for (const Wave& w : c->waves) tot_cap += w.cap;
    CK(c->hom_col.ensure((size_t)(tot_cap + 64) * 4));
    HomArgs h;
    h.g = g; h.extra_cnt = c->hom_cnt.as<u32>(); h.extra_off = c->hom_off.as<u32>(); h.extra_fill = c->hom_fill.as<u32>(); h.extra_col = c->hom_col.as<u32>();
    const int64_t n_units = c->n_owned_genes;
    CK(c->unit_cnt.ensure((size_t)(n_units + 1) * 4));
    CK(c->unit_bytes.ensure((size_t)(n_units + 1) * 4));
    CK(c->cnt_off.ensure((size_t)(n_units + 2) * 8));
    CK(c->byte_off.ensure((size_t)(n_units + 2) * 8));
    c->tot_edges = 0; c->tot_text = 0;
    c->last_emit_grid = 0;
    c->last_is_hom = true;
    if (P > 0 && c->n_rows > 0 && n_units > 0) {
        hom_extras_kernel<false><<<grid_for(c, c->n_rows, 64, 8), 256, 0, c->stream>>>(h);
        CKL("hom_extras_kernel");
        if (device_scan<u32>(c, h.extra_cnt, c->n_rows, c->hom_off.as<u32>(), 1)) return OFG_ERR_CUDA;
        hom_extras_kernel<true><<<grid_for(c, c->n_rows, 64, 8), 256, 0, c->stream>>>(h);
        CKL("hom_extras_kernel");
        HomEmitArgs e;
        e.g = g; e.extra_off = c->hom_off.as<u32>(); e.extra_col = c->hom_col.as<u32>(); e.n_genes = n_units;
        e.gene_of_local = c->gene_of_local.as<int64_t>(); e.species_of_local = c->species_of_local.as<int32_t>(); e.seq_start = c->d_seq_start.as<int64_t>();
        e.gene_cnt = c->unit_cnt.as<u32>(); e.gene_bytes = c->unit_bytes.as<u32>(); e.cnt_off = c->cnt_off.as<u64>(); e.byte_off = c->byte_off.as<u64>();
        e.out_indices = nullptr; e.out_data = nullptr; e.out_text = nullptr; e.err_flags = err_flags_ptr(c); e.cap_edges = 0; e.cap_text = 0;
        c->last_emit_grid = grid_for(c, n_units, HOM_WARPS, 8);
        hom_emit_kernel<false><<<c->last_emit_grid, HOM_WARPS * 32, 0, c->stream>>>(e);
        CKL("hom_emit_kernel<false>");
        if (device_scan<u64>(c, e.gene_cnt, n_units, c->cnt_off.as<u64>(), 1)) return OFG_ERR_CUDA;
        if (device_scan<u64>(c, e.gene_bytes, n_units, c->byte_off.as<u64>(), 1)) return OFG_ERR_CUDA;
        c->last_hom = e;
        if (int rc = launch_emit(c)) return rc;
    }
    c->stage_done = OFG_STAGE_GRAPH;
    c->pending = true;
    c->graph_pending = true;
    return 0;
}

const char* const STAGE_OF_KERNEL[][2] = {
    {"inflate_kernel", "0"}, {"parse_kernel", "0"}, {"scatter_rows_kernel", "1"}, {"rowsort_kernel", "1"}, {"rowsort_short_kernel", "1"}, {"rowsort_long_kernel", "1"},
    {"wave_layout_kernel", "2"}, {"rs_desc_kernel", "2"}, {"radix_hist_kernel", "2"}, {"radix_scatter_kernel", "2"},
    {"bk_hist_kernel", "2"}, {"bk_scan_kernel", "2"}, {"bk_route_kernel", "2"}, {"bk_gather_kernel", "2"}, {"bk_split_kernel", "2"}, {"bk_binhist_kernel", "2"}, {"bk_thr_kernel", "3"}, {"bk_cand_kernel", "3"},
    {"bin_tasks_kernel", "3"}, {"bin_cut_kernel", "3"}, {"bin_select_kernel", "3"}, {"pair_sel_off_kernel", "4"}, {"fit_kernel", "4"},
    {"factors_kernel", "5"}, {"scale_bh_kernel", "5"}, {"bh_self_kernel", "5"}, {"rbh_kernel", "6"}, {"v2_ratio_kernel", "6"},
    {"thresh_final_kernel", "6"}, {"connect_kernel", "7"}, {"emit_sizes_kernel", "8"}, {"emit_kernel<true>", "8"}, {"hom_extras_kernel", "7"}, {"hom_emit_kernel<false>", "8"}, {"hom_emit_kernel<true>", "8"}};

// One host synchronisation for everything a build left on the device: line counts, parse errors, capacity flags,
// fit table, graph sizes, kernel timeline.  Called at the end of ofg_build (unless option async) and by every accessor.
int sync_results(ofg_ctx* c)
{
    if (!c->pending) return 0;
    const int P = (int)c->pairs.size();
    u64 tots[2] = {0, 0};
    for (int attempt = 0; attempt < 3; attempt++) {
        CK(cudaMemcpyAsync(c->h_small, c->d_small.p, c->small_words * 8, cudaMemcpyDeviceToHost, c->stream));
        if (c->gz_err.p) {
            if (!c->h_gz_err) CK(cudaHostAlloc((void**)&c->h_gz_err, 16, cudaHostAllocDefault));
            CK(cudaMemcpyAsync(c->h_gz_err, c->gz_err.p, 8, cudaMemcpyDeviceToHost, c->stream));
        }
        if (P > 0 && c->stage_done >= OFG_STAGE_NORM) CK(cudaMemcpyAsync(c->h_fit, c->fit.p, (size_t)P * 8 * 8, cudaMemcpyDeviceToHost, c->stream));
        const int64_t n_units = c->n_owned_genes;
        if (c->graph_pending && n_units > 0) {
            CK(cudaMemcpyAsync(c->h_fit + (size_t)P * 8, c->cnt_off.as<u64>() + n_units, 8, cudaMemcpyDeviceToHost, c->stream));
            CK(cudaMemcpyAsync(c->h_fit + (size_t)P * 8 + 1, c->byte_off.as<u64>() + n_units, 8, cudaMemcpyDeviceToHost, c->stream));
        }
        CK(cudaStreamSynchronize(c->stream));
        const u32 flags = (u32)c->h_small[2];
        if (c->graph_pending && n_units > 0) {
            memcpy(tots, c->h_fit + (size_t)P * 8, 16);
            if (flags & 64u) {
                // output buffers too small: grow and emit again
                CK(c->out_indices.ensure((size_t)(tots[0] + 32) * 4));
                CK(c->out_data.ensure((size_t)(tots[0] + 32) * 8));
                CK(c->out_text.ensure((size_t)tots[1] + 128));
                u32 cleared = flags & ~64u;
                CK(cudaMemcpyAsync(err_flags_ptr(c), &cleared, 4, cudaMemcpyHostToDevice, c->stream));
                CK(cudaStreamSynchronize(c->stream));
                if (int rc = launch_emit(c)) return rc;
                continue;
            }
        }
        break;
    }
    c->pending = false;
    const u64* small = c->h_small;
    const u32 flags = (u32)small[2];
    c->fallback_waves = (int64_t)small[12];
    const size_t pb = c->small_pair_base;
    c->pair_lines.assign(small + pb, small + pb + P);
    c->pair_valid.assign(small + pb + P, small + pb + 2 * (size_t)P);
    c->pair_nnz.assign(small + pb + 3 * (size_t)P, small + pb + 4 * (size_t)P);
    c->tot_lines = 0; c->tot_records = 0; c->tot_nnz = 0;
    for (int i = 0; i < P; i++) { c->tot_lines += (int64_t)c->pair_lines[i]; c->tot_records += (int64_t)c->pair_valid[i]; c->tot_nnz += (int64_t)c->pair_nnz[i]; }
    // resolve the timeline into the per-stage totals
    for (int i = 0; i < 10; i++) c->ms[i] = 0.f;
    for (int i = 1; i < c->tl_n; i++) {
        if (!c->tl_name[i]) continue;
        float t = 0.f;
        if (cudaEventElapsedTime(&t, c->tl_ev[i - 1], c->tl_ev[i]) != cudaSuccess) continue;
        int st = -1;
        for (auto& m : STAGE_OF_KERNEL) if (!strcmp(m[0], c->tl_name[i])) { st = m[1][0] - '0'; break; }
        if (st < 0 && strstr(c->tl_name[i], "scan_")) {
            // scans belong to the stage of the next non-scan kernel
            for (int k = i + 1; k < c->tl_n && st < 0; k++)
                if (c->tl_name[k] && !strstr(c->tl_name[k], "scan_"))
                    for (auto& m : STAGE_OF_KERNEL) if (!strcmp(m[0], c->tl_name[k])) { st = m[1][0] - '0'; break; }
        }
        if (st >= 0) c->ms[st] += t;
        c->ms[9] += t;
    }
    // a corrupt .gz: gzip.open raises inside the reference's reader (blast_file_processor.py:63-66)
    if (c->gz_err.p && c->h_gz_err && *c->h_gz_err != ~0ULL && !c->gz_members.empty()) {
        const unsigned long long e = *c->h_gz_err;
        const size_t mi = (size_t)(e >> 8);
        const int kind = (int)(e & 255);
        c->stage_done = 0;
        if (mi < c->gz_members.size()) {
            const PairDesc& pd = c->pairs[c->gz_members[mi].file];
            return fail(c, OFG_ERR_PARSE, "pair=%d,%d gzip: compressed hit file is corrupt (member %zu, error kind %d)", pd.p, pd.j, mi, kind);
        }
        return fail(c, OFG_ERR_PARSE, "gzip: compressed hit file is corrupt");
    }
    // parse errors first (the reference fails on the first offending line of a file)
    for (int i = 0; i < P; i++) {
        const u64 e = small[pb + 2 * (size_t)P + i];
        if (e != ~0ULL) {
            std::string d;
            describe_line(c, i, e, &d);
            const int kind = (int)(e & 15);
            const int code = (kind == OFG_LINE_RANGE_Q || kind == OFG_LINE_RANGE_S) ? OFG_ERR_INCONSISTENT
                             : (kind == OFG_LINE_QUOTE || kind == OFG_LINE_SCORE_FORMAT) ? OFG_ERR_UNSUPPORTED : OFG_ERR_PARSE;
            c->stage_done = 0;
            return fail(c, code, "%s", d.c_str());
        }
    }
    for (size_t w = 0; w < c->waves.size(); w++) {
        const u64 used = small[16 + w];
        if (used > c->waves[w].cap) {
            const bool learned = c->bpl_learned > (double)c->bytes_per_line_inv;
            c->bpl_learned = 0.0;
            c->stage_done = 0;
            return fail(c, OFG_ERR_CAPACITY, "record capacity exceeded: %llu lines but room for %llu (%s); %s",
                        (unsigned long long)used, (unsigned long long)c->waves[w].cap,
                        learned ? "sized from the previous build of this context" : "lines shorter than max_bytes_per_line_inv bytes on average",
                        learned ? "build again" : "lower option max_bytes_per_line_inv");
        }
    }
    if (flags & 16u) { c->stage_done = 0; return fail(c, OFG_ERR_CAPACITY, "sort tile / length bin capacity exceeded"); }
    if (c->stage_done >= OFG_STAGE_NORM) {
        c->fit_host.assign(c->h_fit, c->h_fit + (size_t)P * 8);
        if (flags & 2u) {
            c->stage_done = 0;
            return fail(c, OFG_ERR_CAPACITY, "selection capacity exceeded (%llu points); lower option sel_capacity_frac_inv", (unsigned long long)c->sel_cap);
        }
        if (flags & 1u) {
            for (int i = 0; i < P; i++)
                if (c->fit_host[(size_t)i * 8 + 2] > 1 && c->fit_host[(size_t)i * 8 + 6] == 0.0) {
                    c->stage_done = 0;
                    return fail(c, OFG_ERR_FIT, "Optimal parameters not found for pair (%d,%d): MINPACK info %d", c->pairs[i].p, c->pairs[i].j,
                                (int)c->fit_host[(size_t)i * 8 + 3]);
                }
        }
    } else {
        c->fit_host.assign((size_t)P * 8, 0.0);
    }
    if (flags & 8u) return fail(c, OFG_ERR_ARG, "imported list addresses a species pair this context does not hold");
    if (flags & 32u) return fail(c, OFG_ERR_CAPACITY, "exchange inbox too small for the exported lists; enlarge it and repeat the step");
    if (c->graph_pending) {
        c->graph_pending = false;
        c->tot_edges = (int64_t)tots[0]; c->tot_text = (int64_t)tots[1];
        if (flags & 64u) return fail(c, OFG_ERR_CAPACITY, "graph output buffers could not be sized");
        if (flags & 4u) return fail(c, OFG_ERR_UNSUPPORTED, "a graph weight is too large for \"%%.3f\" formatting on the device");
    }
    // size the next build of this context from what this one saw (with a margin)
    if (c->tot_records > 0 && c->stage_done >= OFG_STAGE_RAW) {
        int64_t text = 0;
        u64 slots = 0;
        for (const Wave& w : c->waves) text += w.text_bytes;
        for (size_t w = 0; w < c->waves.size(); w++) slots += small[16 + w];
        if (slots > 0) c->bpl_learned = 0.85 * (double)text / (double)slots;
    }
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
                      &c->keyval, &c->key2, &c->keyval2, &c->long_rows, &c->counters, &c->scan_partial, &c->wave_tile_prefix,
                      &c->wave_pair_slot, &c->bin_tasks, &c->bin_cut, &c->bin_cnt, &c->sel_off, &c->wave_task_prefix, &c->lx, &c->ly, &c->sort_desc, &c->zrow, &c->pair_factor, &c->d_owned, &c->gene_cntg, &c->gene_bytesg, &c->most_init, &c->d_length_factors,
                      &c->fit, &c->rowfac, &c->colfac, &c->rowmax, &c->best, &c->most, &c->owned_gene,
                      &c->gene_of_local, &c->species_of_local, &c->d_seq_start, &c->unit_cnt, &c->unit_bytes, &c->cnt_off,
                      &c->byte_off, &c->out_indices, &c->out_data, &c->out_text, &c->species_ids, &c->synth_bytes, &c->synth_off,
                      &c->file_off, &c->pair_sel_off, &c->row_flagcnt, &c->row_flagoff, &c->diag_pairs,
                      &c->hom_cnt, &c->hom_off, &c->hom_fill, &c->hom_col, &c->gz_comp, &c->gz_members_dev, &c->gz_err,
                      &c->bk_hist, &c->bk_rec, &c->bk_bbidx, &c->bk_shist, &c->bk_thr, &c->bk_below, &c->bk_bin16, &c->bk_bb, &c->bk_bbfill, &c->bk_split, &c->bk_gcursor, &c->bin_val,
                      &c->x_flag, &c->x_off, &c->x_list, &c->x_pair_dst, &c->x_order, &c->x_dst_first, &c->x_inbox_ptrs};
    for (DevBuf* b : bufs) b->release();
    for (auto& e : c->tl_ev) if (e) cudaEventDestroy(e);
    if (c->h_small) cudaFreeHost(c->h_small);
    if (c->h_fit) cudaFreeHost(c->h_fit);
    if (c->h_gz_err) cudaFreeHost(c->h_gz_err);
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
        if (n_seqs[p] < 0 || n_seqs[p] > (int64_t)OFG_COL_MASK) return fail(c, OFG_ERR_UNSUPPORTED, "species %d has %lld sequences (limit 2^28-1)", p, (long long)n_seqs[p]);
        c->seq_start[p] = acc;
        c->len_off[p] = acc;
        acc += n_seqs[p];
    }
    if (acc >= 2147483647LL) return fail(c, OFG_ERR_UNSUPPORTED, "more than 2^31-1 sequences in total");
    c->N = acc;
    c->lengths.assign(lengths_concat, lengths_concat + acc);
    c->owned.clear();
    c->has_partition = false;
    c->plan_sig.clear();
    c->x_sig.clear();
    c->pending = false;
    c->graph_pending = false;
    c->bpl_learned = 0.0;
    c->pairs_dirty = true;
    c->files.clear();
    c->pairs.clear();
    c->text_used = 0;
    c->gz_used = 0;
    c->gz_members.clear();
    c->gz_pending = false;
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
    c->has_partition = true;
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
    else if (!strcmp(name, "wave_species")) { c->wave_species = (int)std::max<int64_t>(value, 0); c->stage_done = 0; }
    else if (!strcmp(name, "async")) c->async = value != 0;
    else if (!strcmp(name, "length_sort_radix")) c->radix_only = value != 0;
    else if (!strcmp(name, "keep_self_hits")) { c->keep_self = value != 0; c->stage_done = 0; }
    else if (!strcmp(name, "rowsort_short")) c->rowsort_short = (int)value;
    else if (!strcmp(name, "homology_graph")) { c->homology = value != 0; c->stage_done = std::min(c->stage_done, (int)OFG_STAGE_NORM); }
    else if (!strcmp(name, "reserve_text_bytes")) { cudaSetDevice(c->device); return text_reserve(c, value); }
    else if (!strcmp(name, "reserve_gz_bytes")) { cudaSetDevice(c->device); CK(c->gz_comp.ensure((size_t)value + 64)); }
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
    c->gz_used = 0;
    c->gz_members.clear();
    c->gz_pending = false;
    c->stage_done = 0;
    return OFG_OK;
}

// One gzip member starting at data[o]: header fields (RFC 1952), the "BC" extra subfield of BGZF-style writers gives the
// member's total size.  Returns false if it is not a gzip member.
static bool gz_parse_member(const uint8_t* d, size_t n, size_t o, size_t* deflate_start, size_t* member_size)
{
    if (o + 18 > n || d[o] != 0x1f || d[o + 1] != 0x8b || d[o + 2] != 8) return false;
    const int flg = d[o + 3];
    size_t q = o + 10;
    *member_size = 0;
    if (flg & 4) {
        if (q + 2 > n) return false;
        const size_t xlen = d[q] | ((size_t)d[q + 1] << 8);
        q += 2;
        if (q + xlen > n) return false;
        size_t x = q;
        while (x + 4 <= q + xlen) {
            const size_t slen = d[x + 2] | ((size_t)d[x + 3] << 8);
            if (d[x] == 'B' && d[x + 1] == 'C' && slen == 2 && x + 6 <= q + xlen) *member_size = (size_t)(d[x + 4] | ((size_t)d[x + 5] << 8)) + 1;
            else if (d[x] == 'O' && d[x + 1] == 'F' && slen == 4 && x + 8 <= q + xlen)   // this repo's writer: 32-bit member size
                *member_size = (size_t)d[x + 4] | ((size_t)d[x + 5] << 8) | ((size_t)d[x + 6] << 16) | ((size_t)d[x + 7] << 24);
            x += 4 + slen;
        }
        q += xlen;
    }
    if (flg & 8) { while (q < n && d[q]) q++; q++; }
    if (flg & 16) { while (q < n && d[q]) q++; q++; }
    if (flg & 2) q += 2;
    if (q + 8 > n) return false;
    *deflate_start = q;
    return true;
}

int ofg_add_hits_gz(ofg_ctx* c, int p, int j, const void* host_ptr, size_t nbytes, int swap_cols)
{
    if (!c || (!host_ptr && nbytes)) return OFG_ERR_ARG;
    cudaSetDevice(c->device);
    if (int rc = ensure_pairs(c)) return rc;
    if (p < 0 || j < 0 || p >= c->S || j >= c->S) return fail(c, OFG_ERR_ARG, "pair (%d,%d) out of range", p, j);
    const int pi = c->pair_of[(size_t)p * c->S + j];
    if (pi < 0) return fail(c, OFG_ERR_ARG, "pair (%d,%d) is not part of this context's partition", p, j);
    const uint8_t* d = (const uint8_t*)host_ptr;
    // walk the members: with a size hint in every header the walk never inflates; a file without hints is taken as ONE member
    std::vector<GzMember> ms;
    const int64_t cbase = (c->gz_used + 15) & ~15LL;
    size_t o = 0;
    u64 raw_total = 0;
    while (o < nbytes) {
        if (d[o] == 0) { o++; continue; }   // zero padding between members is legal (gzip.py skips it)
        size_t ds = 0, msz = 0;
        if (!gz_parse_member(d, nbytes, o, &ds, &msz)) return fail(c, OFG_ERR_UNSUPPORTED, "Blast file of pair (%d,%d): not a gzip stream at byte %zu", p, j, o);
        size_t end;
        if (msz) {
            end = o + msz;
            if (end > nbytes || end < ds + 8) return fail(c, OFG_ERR_PARSE, "Blast file of pair (%d,%d): gzip member at byte %zu is truncated", p, j, o);
        } else {
            if (!ms.empty()) return fail(c, OFG_ERR_UNSUPPORTED, "Blast file of pair (%d,%d): several gzip members without size hints", p, j);
            end = nbytes;   // a classic single-member file
        }
        GzMember m;
        m.in_off = (u64)cbase + ds; m.in_len = end - ds - 8; m.out_off = raw_total; m.file = pi; m.pad = 0;
        m.out_len = (u64)d[end - 4] | ((u64)d[end - 3] << 8) | ((u64)d[end - 2] << 16) | ((u64)d[end - 1] << 24);
        if (m.in_len >= 0xffffff00ull) return fail(c, OFG_ERR_UNSUPPORTED, "Blast file of pair (%d,%d): a gzip member of 4 GiB or more", p, j);
        raw_total += m.out_len;
        ms.push_back(m);
        o = end;
    }
    // text arena range of the inflated file
    const int64_t off = (c->text_used + 15) & ~15LL;
    const int64_t padded = ((int64_t)raw_total + 16) & ~15LL;
    if (int rc = text_reserve(c, off + padded)) return rc;
    // compressed arena (kept until the build: the members are inflated together)
    const int64_t cneed = cbase + (int64_t)nbytes + 64;
    if ((size_t)cneed > c->gz_comp.cap) {
        DevBuf nb;
        CK(nb.ensure((size_t)std::max<int64_t>(cneed, (int64_t)(c->gz_comp.cap + c->gz_comp.cap / 2))));
        if (c->gz_comp.p && c->gz_used > 0) CK(cudaMemcpyAsync(nb.p, c->gz_comp.p, (size_t)c->gz_used, cudaMemcpyDeviceToDevice, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        c->gz_comp.release();
        c->gz_comp = nb;
    }
    if (nbytes) CK(cudaMemcpyAsync(c->gz_comp.as<uint8_t>() + cbase, host_ptr, nbytes, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemsetAsync(c->text.as<uint8_t>() + off + raw_total, '\n', (size_t)(padded - (int64_t)raw_total), c->stream));
    for (auto& m : ms) { m.out_off += (u64)off; c->gz_members.push_back(m); }
    FileInfo& f = c->files[pi];
    f.off = off; f.raw = (int64_t)raw_total; f.padded = padded; f.swap = swap_cols;
    c->text_used = off + padded;
    c->gz_used = cbase + (int64_t)nbytes;
    c->gz_pending = true;
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
    c->plan_sig.clear();   // the descriptor table on the device is rewritten here
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
    if (c->gz_pending) {   // compressed files are inflated on demand
        if (int rc = inflate_pending(c)) return rc;
        unsigned long long e = ~0ULL;
        CK(cudaMemcpyAsync(&e, c->gz_err.p, 8, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        if (e != ~0ULL) {
            const size_t mi = (size_t)(e >> 8);
            const PairDesc& pd = c->pairs[c->gz_members[std::min(mi, c->gz_members.size() - 1)].file];
            return fail(c, OFG_ERR_PARSE, "pair=%d,%d gzip: compressed hit file is corrupt (member %zu, error kind %d)", pd.p, pd.j, mi, (int)(e & 255));
        }
    }
    const int pi = c->pair_of[(size_t)p * c->S + j];
    CK(cudaMemcpyAsync(host_dst, c->text.as<uint8_t>() + c->files[pi].off, n, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return OFG_OK;
}

static int build_once(ofg_ctx* c, int stop_after)
{
    if (c->stage_done >= stop_after) c->stage_done = 0;  // asked for a stage that is already built: rebuild from the text
    if (c->stage_done == OFG_STAGE_RAW && !(c->raw_scratch_valid && c->waves.size() == 1)) c->stage_done = 0;  // sort keys of the waves are gone
    if (c->stage_done == 0) {
        if (c->pending) { c->pending = false; c->graph_pending = false; CK(cudaStreamSynchronize(c->stream)); }
        c->tl_n = 0;
        c->launches = 0;
    }
    if (c->stage_done < OFG_STAGE_RAW) {
        if (int rc = stage_raw_norm(c, stop_after)) { c->stage_done = 0; return rc; }
    } else if (stop_after >= OFG_STAGE_NORM && c->stage_done < OFG_STAGE_NORM) {
        if (int rc = stage_norm_only(c)) { c->stage_done = 0; return rc; }
    }
    if (c->homology && stop_after > OFG_STAGE_NORM) {
        if (c->stage_done < OFG_STAGE_GRAPH) if (int rc = stage_homology(c)) return rc;
        if (c->async) return OFG_OK;
        return sync_results(c);
    }
    if (stop_after >= OFG_STAGE_THRESH && c->stage_done < OFG_STAGE_THRESH)
        if (int rc = stage_thresh(c)) { c->stage_done = 0; return rc; }
    if (stop_after >= OFG_STAGE_CONNECT && c->stage_done < OFG_STAGE_CONNECT)
        if (int rc = stage_connect(c)) return rc;
    if (stop_after >= OFG_STAGE_GRAPH && c->stage_done < OFG_STAGE_GRAPH)
        if (int rc = stage_emit(c)) return rc;
    if (c->async) return OFG_OK;
    return sync_results(c);
}

int ofg_build(ofg_ctx* c, int stop_after)
{
    if (!c) return OFG_ERR_ARG;
    cudaSetDevice(c->device);
    if (int rc = ensure_pairs(c)) return rc;
    if (stop_after < OFG_STAGE_RAW || stop_after > OFG_STAGE_GRAPH) return fail(c, OFG_ERR_ARG, "bad stage %d", stop_after);
    if (c->keep_self && stop_after > OFG_STAGE_RAW)
        return fail(c, OFG_ERR_ARG, "option keep_self_hits is for the raw matrices only (OFG_STAGE_RAW, DendroBLAST): the graph path drops self hits");
    const bool sized_from_last = c->bpl_learned > (double)c->bytes_per_line_inv;
    int rc = build_once(c, stop_after);
    if (rc == OFG_ERR_CAPACITY && sized_from_last && c->bpl_learned == 0.0 && !c->async)
        rc = build_once(c, stop_after);  // the capacities learned from the previous input were too small for this one: conservative sizes
    return rc;
}

int ofg_sync(ofg_ctx* c)
{
    if (!c) return OFG_ERR_ARG;
    cudaSetDevice(c->device);
    if (!c->pending) { CK(cudaStreamSynchronize(c->stream)); return OFG_OK; }
    return sync_results(c);
}

int ofg_get_stat(ofg_ctx* c, const char* name, int64_t* value)
{
    if (!c || !name || !value) return OFG_ERR_ARG;
    cudaSetDevice(c->device);
    if (int rc = sync_results(c)) return rc;
    if (!strcmp(name, "radix_fallback_waves")) *value = c->fallback_waves;
    else if (!strcmp(name, "waves")) *value = (int64_t)c->waves.size();
    else if (!strcmp(name, "record_slot_capacity")) { int64_t t = 0; for (const Wave& w : c->waves) t += (int64_t)w.cap; *value = t; }
    else return fail(c, OFG_ERR_ARG, "unknown statistic %s", name);
    return OFG_OK;
}

int ofg_wait_for(ofg_ctx* c, ofg_ctx* other)
{
    if (!c || !other) return OFG_ERR_ARG;
    if (c == other) return OFG_OK;
    cudaSetDevice(other->device);
    cudaEvent_t e = nullptr;
    CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    CK(cudaEventRecord(e, other->stream));
    cudaSetDevice(c->device);
    CK(cudaStreamWaitEvent(c->stream, e, 0));
    CK(cudaEventDestroy(e));  // released once the recorded work has completed
    return OFG_OK;
}

int ofg_kernel_report(ofg_ctx* c, char* buf, size_t cap)
{
    if (!c || !buf || cap == 0) return OFG_ERR_ARG;
    cudaSetDevice(c->device);
    if (int rc = sync_results(c)) return rc;
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
    cudaSetDevice(c->device);
    if (int rc = sync_results(c)) return rc;
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
    cudaSetDevice(c->device);
    if (int rc_ = sync_results(c)) return rc_;
    if (c->stage_done < OFG_STAGE_THRESH) return fail(c, OFG_ERR_ARG, "thresholds are not built yet");
    cudaSetDevice(c->device);
    CK(cudaMemcpyAsync(host_dst, c->most.p, (size_t)std::min<int64_t>(n, c->N) * 8, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return OFG_OK;
}

// row_flagcnt is one buffer: scale_bh_kernel fills it with best-hit counts, connect_kernel overwrites it with connect
// counts.  The best-hit lists can therefore only be exported between NORM and CONNECT, the connect lists from CONNECT on.
static int check_exchange_stage(ofg_ctx* c, int which, const char* what)
{
    if (which != 0 && which != 1) return fail(c, OFG_ERR_ARG, "which must be 0 (best hits) or 1 (connect)");
    if (which == 0 && !(c->stage_done == OFG_STAGE_NORM || c->stage_done == OFG_STAGE_THRESH))
        return fail(c, OFG_ERR_ARG, "%s of best-hit lists is only possible right after stage NORM / THRESH (stage is %d)", what, c->stage_done);
    if (which == 1 && c->stage_done != OFG_STAGE_CONNECT)
        return fail(c, OFG_ERR_ARG, "%s of connect lists is only possible right after stage CONNECT (stage is %d)", what, c->stage_done);
    return 0;
}

int ofg_export_flags(ofg_ctx* c, int which, int n_ranks, const int32_t* species_rank, int my_rank, int64_t* counts, int32_t** dev_list)
{
    if (!c || !species_rank || !counts || !dev_list || n_ranks <= 0) return OFG_ERR_ARG;
    cudaSetDevice(c->device);
    if (int rc = sync_results(c)) return rc;
    if (int rc = check_exchange_stage(c, which, "export")) return rc;
    const int P = (int)c->pairs.size();
    std::vector<uint8_t> ex(P, 0);
    for (int i = 0; i < P; i++) {
        const PairDesc& d = c->pairs[i];
        ex[i] = d.owned_row && species_rank[d.j] != my_rank;  // (k, p): column species p lives elsewhere
    }
    DevBuf& flagbuf = c->x_flag;
    DevBuf& offbuf = c->x_off;
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
    CK(c->x_list.ensure((size_t)(total + 1) * 8));
    CK(cudaMemcpyAsync(offbuf.p, off.data(), sizeof(u64) * P, cudaMemcpyHostToDevice, c->stream));
    ExportArgs a;
    a.n_rows = c->n_rows; a.row_start = c->row_start.as<u32>(); a.rowlen = c->rowlen.as<u32>(); a.cols = c->cols.as<u32>();
    a.pairs = c->d_pairs.as<PairDesc>(); a.pair_row_base = c->d_pair_row_base.as<int64_t>(); a.P = P;
    a.pair_export = flagbuf.as<uint8_t>(); a.flag = which == 0 ? OFG_FLAG_BH : OFG_FLAG_CONNECT;
    a.row_off = c->row_flagoff.as<u64>(); a.pair_off = offbuf.as<u64>(); a.out = c->x_list.as<int2>();
    if (total > 0) {
        export_flags_kernel<<<grid_for(c, c->n_rows, 8, 8), 256, 0, c->stream>>>(a);
        CKL("export_flags_kernel");
    }
    CK(cudaStreamSynchronize(c->stream));
    *dev_list = c->x_list.as<int32_t>();
    return OFG_OK;
}

static void fill_import_args(ofg_ctx* c, int which, ImportArgs& a)
{
    a.list = nullptr; a.n = 0; a.seq_start = c->d_seq_start.as<int64_t>(); a.S = c->S;
    a.pair_of = c->d_pair_of.as<int32_t>(); a.pairs = c->d_pairs.as<PairDesc>(); a.row_start = c->row_start.as<u32>();
    a.rowlen = c->rowlen.as<u32>(); a.cols = c->cols.as<u32>(); a.set_flag = which == 0 ? OFG_FLAG_TBH : OFG_FLAG_REVERSE;
    a.err_flags = err_flags_ptr(c); a.vals = c->vals.as<double>(); a.gene_cnt = c->gene_cntg.as<u32>(); a.gene_bytes = c->gene_bytesg.as<u32>();
    a.inbox = nullptr; a.seg_bytes = 0; a.my_rank = -1;
}

int ofg_import_flags(ofg_ctx* c, int which, const int32_t* dev_list, int64_t n_entries)
{
    if (!c || n_entries < 0 || (n_entries > 0 && !dev_list)) return OFG_ERR_ARG;
    cudaSetDevice(c->device);
    if (int rc = check_exchange_stage(c, which, "import")) return rc;
    if (n_entries == 0) return OFG_OK;
    ImportArgs a;
    fill_import_args(c, which, a);
    a.list = reinterpret_cast<const int2*>(dev_list); a.n = n_entries;
    import_flags_kernel<<<grid_for(c, n_entries, 256, 8), 256, 0, c->stream>>>(a);
    CKL("import_flags_kernel");
    c->pending = true;
    if (c->async) return OFG_OK;
    return sync_results(c);
}

// ---- peer exchange: lists written straight into the destinations' inboxes, everything sized and placed on the device
int ofg_export_flags_peer(ofg_ctx* c, int which, int n_ranks, const int32_t* species_rank, int my_rank, void* const* inbox_ptrs,
                          size_t seg_bytes)
{
    if (!c || !species_rank || !inbox_ptrs || n_ranks <= 0 || n_ranks > 32 || my_rank < 0 || my_rank >= n_ranks || seg_bytes < 16 || (seg_bytes & 7))
        return c ? fail(c, OFG_ERR_ARG, "bad arguments to ofg_export_flags_peer") : OFG_ERR_ARG;
    cudaSetDevice(c->device);
    if (int rc = check_exchange_stage(c, which, "export")) return rc;
    const int P = (int)c->pairs.size();
    // static part (depends on the partition only): destination of every pair, exported pairs grouped by destination
    std::vector<int32_t> pair_dst(std::max(P, 1), -1), order, dst_first(n_ranks + 1, 0);
    for (int i = 0; i < P; i++) {
        const PairDesc& d = c->pairs[i];
        const int r = species_rank[d.j];
        if (r < 0 || r >= n_ranks) return fail(c, OFG_ERR_ARG, "species %d has no owner rank", d.j);
        if (d.owned_row && r != my_rank) pair_dst[i] = r;
    }
    for (int r = 0; r < n_ranks; r++) {
        dst_first[r] = (int32_t)order.size();
        for (int i = 0; i < P; i++) if (pair_dst[i] == r) order.push_back(i);
    }
    dst_first[n_ranks] = (int32_t)order.size();
    if (order.empty()) order.push_back(0);
    // uploaded once per (partition, owner map, inbox set): the steady state of a job enqueues the exchange without a host stall
    std::vector<int64_t> sig;
    sig.push_back(n_ranks); sig.push_back(my_rank); sig.push_back(P);
    for (int i = 0; i < P; i++) sig.push_back(pair_dst[i]);
    for (int r = 0; r < n_ranks; r++) sig.push_back((int64_t)(uintptr_t)inbox_ptrs[r]);
    if (sig != c->x_sig) {
        if (upload(c, c->x_pair_dst, pair_dst.data(), pair_dst.size() * 4)) return OFG_ERR_CUDA;
        if (upload(c, c->x_order, order.data(), order.size() * 4)) return OFG_ERR_CUDA;
        if (upload(c, c->x_dst_first, dst_first.data(), dst_first.size() * 4)) return OFG_ERR_CUDA;
        if (upload(c, c->x_inbox_ptrs, inbox_ptrs, sizeof(void*) * n_ranks)) return OFG_ERR_CUDA;
        CK(cudaStreamSynchronize(c->stream));  // host vectors go out of scope (a few KB; the lists themselves never touch the host)
        c->x_sig = sig;
    }
    CK(c->row_flagoff.ensure((size_t)(c->n_rows + 2) * 8));
    CK(c->x_off.ensure((size_t)(P + 4) * 8));
    if (device_scan<u64>(c, c->row_flagcnt.as<u32>(), c->n_rows, c->row_flagoff.as<u64>(), 1)) return OFG_ERR_CUDA;
    PeerExportArgs a;
    a.n_rows = c->n_rows; a.row_start = c->row_start.as<u32>(); a.rowlen = c->rowlen.as<u32>(); a.cols = c->cols.as<u32>();
    a.pairs = c->d_pairs.as<PairDesc>(); a.pair_row_base = c->d_pair_row_base.as<int64_t>(); a.P = P;
    a.pair_dst = c->x_pair_dst.as<int32_t>(); a.flag = which == 0 ? OFG_FLAG_BH : OFG_FLAG_CONNECT;
    a.row_off = c->row_flagoff.as<u64>(); a.pair_off = c->x_off.as<u64>(); a.order = c->x_order.as<int32_t>();
    a.dst_first = c->x_dst_first.as<int32_t>(); a.inbox = reinterpret_cast<unsigned char* const*>(c->x_inbox_ptrs.p);
    a.seg_bytes = (u64)seg_bytes; a.my_rank = my_rank; a.n_ranks = n_ranks; a.err_flags = err_flags_ptr(c);
    export_layout_kernel<<<1, 32 * n_ranks, 0, c->stream>>>(a);
    CKL("export_layout_kernel");
    if (P > 0 && c->n_rows > 0 && dst_first[n_ranks] > 0) {
        export_peer_kernel<4><<<grid_for(c, c->n_rows, 64, 8), 256, 0, c->stream>>>(a);
        CKL("export_peer_kernel");
    }
    c->pending = true;
    return OFG_OK;
}

int ofg_import_flags_peer(ofg_ctx* c, int which, int n_ranks, int my_rank, const void* inbox, size_t seg_bytes)
{
    if (!c || !inbox || n_ranks <= 0 || my_rank < 0 || my_rank >= n_ranks) return c ? fail(c, OFG_ERR_ARG, "bad arguments to ofg_import_flags_peer") : OFG_ERR_ARG;
    cudaSetDevice(c->device);
    if (int rc = check_exchange_stage(c, which, "import")) return rc;
    ImportArgs a;
    fill_import_args(c, which, a);
    a.inbox = reinterpret_cast<const unsigned char*>(inbox); a.seg_bytes = (u64)seg_bytes; a.my_rank = my_rank;
    dim3 g((unsigned)std::max(1, c->n_sm * 2 / std::max(n_ranks - 1, 1)), (unsigned)n_ranks);
    import_flags_kernel<<<g, 256, 0, c->stream>>>(a);
    CKL("import_flags_kernel");
    c->pending = true;
    return OFG_OK;
}

int ofg_ipc_get_handle(ofg_ctx* c, void* dev_ptr, void* handle_out64)
{
    if (!c || !dev_ptr || !handle_out64) return OFG_ERR_ARG;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaSetDevice(c->device);
    CK(cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(handle_out64), dev_ptr));
    return OFG_OK;
}

int ofg_ipc_open_handle(ofg_ctx* c, const void* handle64, void** dev_ptr_out)
{
    if (!c || !handle64 || !dev_ptr_out) return OFG_ERR_ARG;
    cudaSetDevice(c->device);
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof h);
    CK(cudaIpcOpenMemHandle(dev_ptr_out, h, cudaIpcMemLazyEnablePeerAccess));
    return OFG_OK;
}

int ofg_ipc_close_handle(ofg_ctx* c, void* dev_ptr)
{
    if (!c || !dev_ptr) return OFG_ERR_ARG;
    cudaSetDevice(c->device);
    CK(cudaIpcCloseMemHandle(dev_ptr));
    return OFG_OK;
}

int ofg_device_alloc(ofg_ctx* c, size_t bytes, void** dev_ptr_out)
{
    if (!c || !dev_ptr_out) return OFG_ERR_ARG;
    cudaSetDevice(c->device);
    CK(cudaMalloc(dev_ptr_out, bytes));
    CK(cudaMemset(*dev_ptr_out, 0, bytes));
    return OFG_OK;
}

int ofg_device_free(ofg_ctx* c, void* dev_ptr)
{
    if (!c) return OFG_ERR_ARG;
    cudaSetDevice(c->device);
    if (dev_ptr) CK(cudaFree(dev_ptr));
    return OFG_OK;
}

int ofg_counts(ofg_ctx* c, int64_t* lines, int64_t* records, int64_t* nnz, int64_t* edges)
{
    if (!c) return OFG_ERR_ARG;
    cudaSetDevice(c->device);
    if (int rc_ = sync_results(c)) return rc_;
    if (lines) *lines = c->tot_lines;
    if (records) *records = c->tot_records;
    if (nnz) *nnz = c->tot_nnz;
    if (edges) *edges = c->tot_edges;
    return OFG_OK;
}

int ofg_pair_csr(ofg_ctx* c, int p, int j, int64_t* n_rows, int64_t* nnz, int32_t* row_len, uint32_t* cols, double* vals)
{
    if (!c) return OFG_ERR_ARG;
    cudaSetDevice(c->device);
    if (int rc_ = sync_results(c)) return rc_;
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

// DendroBLAST distance matrices of orthogroups on the raw matrices (orthologues.py:264-291, 396-410).
int ofg_dendro_matrices(ofg_ctx* c, int64_t n_ogs, const int64_t* og_off, const int32_t* species, const int32_t* iseq, int complete,
                        double* host_out, int64_t out_len, double* host_max_og)
{
    if (!c || !og_off || n_ogs < 0 || (n_ogs > 0 && (!species || !iseq || !host_out))) return OFG_ERR_ARG;
    cudaSetDevice(c->device);
    if (int rc_ = sync_results(c)) return rc_;
    if (c->stage_done != OFG_STAGE_RAW)
        return fail(c, OFG_ERR_ARG, "the distance matrices need the raw scores: build to OFG_STAGE_RAW and no further (stage is %d)", c->stage_done);
    if (c->has_partition) return fail(c, OFG_ERR_ARG, "the distance matrices need every species pair in one context");
    if (og_off[0] != 0) return fail(c, OFG_ERR_ARG, "og_off[0] must be 0");
    const int64_t K = og_off[n_ogs];
    std::vector<int32_t> mem_og((size_t)K);
    std::vector<int64_t> out_off((size_t)n_ogs + 1, 0);
    for (int64_t g = 0; g < n_ogs; g++) {
        const int64_t n = og_off[g + 1] - og_off[g];
        if (n < 0) return fail(c, OFG_ERR_ARG, "og_off must not decrease (orthogroup %lld)", (long long)g);
        out_off[(size_t)g + 1] = out_off[(size_t)g] + n * n;
        for (int64_t k = og_off[g]; k < og_off[g + 1]; k++) {
            if (species[k] < 0 || species[k] >= c->S || iseq[k] < 0 || iseq[k] >= c->n_seqs[species[k]])
                return fail(c, OFG_ERR_ARG, "orthogroup %lld member %lld: gene %d_%d does not exist", (long long)g, (long long)(k - og_off[g]), species[k], iseq[k]);
            mem_og[(size_t)k] = (int32_t)g;
        }
    }
    if (out_len < out_off[(size_t)n_ogs]) return fail(c, OFG_ERR_ARG, "output holds %lld values, %lld needed", (long long)out_len, (long long)out_off[(size_t)n_ogs]);
    if (n_ogs == 0 || K == 0) return OFG_OK;
    DevBuf gmin, gmaxinv, d_mem_og, d_og_off, d_out_off, d_sp, d_seq, dM, dD, d_max;
    DevBuf* tmp[] = {&gmin, &gmaxinv, &d_mem_og, &d_og_off, &d_out_off, &d_sp, &d_seq, &dM, &dD, &d_max};
    auto done = [&](int rc) { for (DevBuf* b : tmp) b->release(); return rc; };
    const int64_t E = out_off[(size_t)n_ogs];
    std::vector<unsigned long long> max_init((size_t)n_ogs, 0ull);
    {
        const double lo = -9e99;   // max_og starts at -9e99 (:401)
        unsigned long long b; memcpy(&b, &lo, 8);
        b = (b >> 63) ? ~b : (b | 0x8000000000000000ull);
        std::fill(max_init.begin(), max_init.end(), b);
    }
    if (gmin.ensure((size_t)c->N * 8) != cudaSuccess || gmaxinv.ensure((size_t)c->N * 8) != cudaSuccess || dM.ensure((size_t)E * 8) != cudaSuccess ||
        (complete && dD.ensure((size_t)E * 8) != cudaSuccess))
        return done(fail(c, OFG_ERR_CUDA, "out of device memory for %lld matrix elements", (long long)E));
    if (upload(c, d_mem_og, mem_og.data(), (size_t)K * 4) || upload(c, d_og_off, og_off, (size_t)(n_ogs + 1) * 8) ||
        upload(c, d_out_off, out_off.data(), (size_t)(n_ogs + 1) * 8) || upload(c, d_sp, species, (size_t)K * 4) ||
        upload(c, d_seq, iseq, (size_t)K * 4) || upload(c, d_max, max_init.data(), (size_t)n_ogs * 8))
        return done(OFG_ERR_CUDA);
    DendroArgs a;
    a.pairs = c->d_pairs.as<PairDesc>(); a.pair_of = c->d_pair_of.as<int32_t>(); a.S = c->S; a.seq_start = c->d_seq_start.as<int64_t>();
    a.row_start = c->row_start.as<u32>(); a.rowlen = c->rowlen.as<u32>(); a.cols = c->cols.as<u32>(); a.vals = c->vals.as<double>();
    a.gmin = gmin.as<double>(); a.gmaxinv = gmaxinv.as<double>(); a.N = c->N;
    a.n_members = K; a.mem_og = d_mem_og.as<int32_t>(); a.og_off = d_og_off.as<int64_t>(); a.out_off = d_out_off.as<int64_t>();
    a.mem_sp = d_sp.as<int32_t>(); a.mem_seq = d_seq.as<int32_t>(); a.M = dM.as<double>(); a.D = dD.as<double>();
    a.og_max = d_max.as<unsigned long long>();
    dendro_minmax_kernel<<<(unsigned)((c->N + 7) / 8), 256, 0, c->stream>>>(a);
    dendro_rows_kernel<<<(unsigned)((K + 7) / 8), 256, 0, c->stream>>>(a);
    if (complete) dendro_complete_kernel<<<(unsigned)((K + 7) / 8), 256, 0, c->stream>>>(a);
    if (cudaGetLastError() != cudaSuccess) return done(fail(c, OFG_ERR_CUDA, "distance matrix kernels failed to launch"));
    cudaError_t e = cudaMemcpyAsync(host_out, complete ? dD.p : dM.p, (size_t)E * 8, cudaMemcpyDeviceToHost, c->stream);
    std::vector<unsigned long long> mx((size_t)n_ogs);
    if (e == cudaSuccess && host_max_og && complete) e = cudaMemcpyAsync(mx.data(), d_max.p, (size_t)n_ogs * 8, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) return done(fail(c, OFG_ERR_CUDA, "distance matrices: %s", cudaGetErrorString(e)));
    if (host_max_og)
        for (int64_t g = 0; g < n_ogs; g++) {
            unsigned long long b = complete ? mx[(size_t)g] : max_init[0];
            b = (b >> 63) ? (b & 0x7fffffffffffffffull) : ~b;
            memcpy(&host_max_og[g], &b, 8);
        }
    return done(OFG_OK);
}

int ofg_pair_fit(ofg_ctx* c, int p, int j, double out[6])
{
    if (!c || !out) return OFG_ERR_ARG;
    cudaSetDevice(c->device);
    if (int rc_ = sync_results(c)) return rc_;
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
    cudaSetDevice(c->device);
    if (int rc_ = sync_results(c)) return rc_;
    if (c->stage_done < OFG_STAGE_GRAPH) return fail(c, OFG_ERR_ARG, "graph not built yet");
    if (n_rows) *n_rows = c->n_owned_genes;
    if (nnz) *nnz = c->tot_edges;
    return OFG_OK;
}

int ofg_graph_csr_copy(ofg_ctx* c, int64_t* row_ids, int64_t* indptr, int32_t* indices, double* data)
{
    if (!c) return OFG_ERR_ARG;
    cudaSetDevice(c->device);
    if (int rc_ = sync_results(c)) return rc_;
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

int ofg_graph_dev(ofg_ctx* c, int64_t* n_rows, int64_t* nnz, const uint64_t** row_off_dev, const int32_t** indices_dev,
                  const double** weights_dev)
{
    if (!c) return OFG_ERR_ARG;
    cudaSetDevice(c->device);
    if (int rc_ = sync_results(c)) return rc_;
    if (c->stage_done < OFG_STAGE_GRAPH) return fail(c, OFG_ERR_ARG, "graph not built yet");
    if (c->n_owned_genes != c->N) return fail(c, OFG_ERR_ARG, "the device graph of a partitioned context holds only its own rows");
    CK(cudaStreamSynchronize(c->stream));
    if (n_rows) *n_rows = c->n_owned_genes;
    if (nnz) *nnz = c->tot_edges;
    if (row_off_dev) *row_off_dev = (const uint64_t*)c->cnt_off.p;
    if (indices_dev) *indices_dev = c->out_indices.as<int32_t>();
    if (weights_dev) *weights_dev = c->out_data.as<double>();
    return OFG_OK;
}

int ofg_graph_text_size(ofg_ctx* c, size_t* nbytes)
{
    if (!c || !nbytes) return OFG_ERR_ARG;
    cudaSetDevice(c->device);
    if (int rc_ = sync_results(c)) return rc_;
    if (c->stage_done < OFG_STAGE_GRAPH) return fail(c, OFG_ERR_ARG, "graph not built yet");
    *nbytes = (size_t)c->tot_text;
    return OFG_OK;
}

int ofg_graph_text_copy(ofg_ctx* c, void* host_dst, size_t cap)
{
    if (!c || !host_dst) return OFG_ERR_ARG;
    cudaSetDevice(c->device);
    if (int rc_ = sync_results(c)) return rc_;
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
    CK(ofg_launch_test_fit(x.as<double>(), y.as<double>(), (long)m, o.as<double>(), c->stream));
    CKL("test_fit_kernel");
    double h[8];
    CK(cudaMemcpyAsync(h, o.p, 48, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    out[0] = h[0]; out[1] = h[1]; out[2] = h[3]; out[3] = h[4];
    x.release(); y.release(); o.release();
    return OFG_OK;
}

}  // extern "C"
