// libofg, MCL on the device: C-ABI `ofg_mcl_*` of include/ofg.h.  Replaces the external
// process that the reference starts in RunMCL (scripts_of/mcl.py:214-218: `mcl graph -I inflation -o clusters -te n -V all`)
// for the graph that is already in device memory.  Kernels = the per-column functions of ofg_mcl_core.h; the iteration
// loop is ofg_mcl_driver.h over the CUDA backend below.  Built for sm_100a only; no CPU fallback: every entry point fails
// when there is no device.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <vector>

#include "../../include/ofg.h"
#include "ofg_mcl_driver.h"

namespace {

__global__ void __launch_bounds__(256) mcl_start_count_kernel(MclStartArgs a)
{
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < a.n; c += (int64_t)gridDim.x * blockDim.x) mcl_start_count(a, c);
}
__global__ void __launch_bounds__(256) mcl_start_fill_kernel(MclStartArgs a)
{
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < a.n; c += (int64_t)gridDim.x * blockDim.x) mcl_start_fill(a, c);
}
__global__ void __launch_bounds__(256) mcl_plan_kernel(MclPlanArgs a)
{
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < a.n; c += (int64_t)gridDim.x * blockDim.x) mcl_plan_column(a, c);
}
__global__ void __launch_bounds__(MCL_NT) mcl_expand_kernel(MclExpandArgs a)
{
    __shared__ MclColShared sh;
    mcl_expand_column_hashed(sh, a, a.c0 + (int64_t)blockIdx.x);
}
__global__ void __launch_bounds__(MCL_NT) mcl_expand_direct_kernel(MclExpandArgs a)
{
    __shared__ MclColSharedLite sh;
    mcl_expand_column_direct(sh, a, a.c0 + (int64_t)blockIdx.x);
}
__global__ void __launch_bounds__(MCL_NT) mcl_inflate_kernel(MclInflateArgs a)
{
    __shared__ MclRedShared sh;
    mcl_inflate_column(sh, a, a.c0 + (int64_t)blockIdx.x);
}
__global__ void __launch_bounds__(256) mcl_dag_kernel(MclDagArgs a)
{
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < a.n; c += (int64_t)gridDim.x * blockDim.x) mcl_dag_column(a, c);
}
__global__ void __launch_bounds__(256) mcl_attractor_kernel(MclDagArgs a)
{
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < a.n; c += (int64_t)gridDim.x * blockDim.x) mcl_attractor_round(a, c);
}
__global__ void __launch_bounds__(256) mcl_reach_kernel(MclDagArgs a)
{
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < a.n; c += (int64_t)gridDim.x * blockDim.x) mcl_reach_round(a, c);
}

// exclusive scan by one CTA: out[i] = base + in[0] + ... + in[i-1] for i in [0, count]
template <typename T> __global__ void __launch_bounds__(1024) mcl_scan_kernel(const T* in, int64_t* out, int64_t count, int64_t base)
{
    __shared__ int64_t part[1024];
    const int t = (int)threadIdx.x;
    const int64_t per = (count + 1023) / 1024;
    const int64_t b = (int64_t)t * per < count ? (int64_t)t * per : count;
    const int64_t e = b + per < count ? b + per : count;
    int64_t s = 0;
    for (int64_t i = b; i < e; i++) s += (int64_t)in[i];
    part[t] = s;
    __syncthreads();
    if (t == 0) {
        int64_t r = base;
        for (int i = 0; i < 1024; i++) { const int64_t x = part[i]; part[i] = r; r += x; }
        out[count] = r;
    }
    __syncthreads();
    s = part[t];
    for (int64_t i = b; i < e; i++) { const int64_t x = (int64_t)in[i]; out[i] = s; s += x; }
}

// graph of a built context -> the values the binary would read from OrthoFinder_graph.txt (mcl_weight_as_read)
__global__ void __launch_bounds__(256) mcl_weights_kernel(const double* w, float* out, int64_t nnz, int* bad)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nnz; i += (int64_t)gridDim.x * blockDim.x) {
        float f = 0.0f;
        if (!mcl_weight_as_read(w[i], &f)) *bad = 1;
        out[i] = f;
    }
}

__global__ void __launch_bounds__(256) mcl_rowptr_kernel(const unsigned long long* row_off, int64_t n, int64_t nnz, int64_t* cp)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= n; i += (int64_t)gridDim.x * blockDim.x)
        cp[i] = i < n ? (int64_t)row_off[i] : nnz;
}

struct CudaBackend {
    cudaStream_t stream = nullptr;
    std::string e;
    int n_sm = 148;
    const char* error() { return e.c_str(); }
    bool ck(cudaError_t r, const char* what)
    {
        if (r == cudaSuccess) return true;
        e = std::string(what) + ": " + cudaGetErrorString(r);
        return false;
    }
    bool launched(const char* what) { return ck(cudaGetLastError(), what); }
    int grid_for(int64_t n) const
    {
        int64_t g = (n + 255) / 256;
        const int64_t cap = (int64_t)n_sm * 16;
        return (int)(g < 1 ? 1 : (g > cap ? cap : g));
    }
    bool alloc(void** p, size_t bytes) { return ck(cudaMalloc(p, bytes ? bytes : 16), "cudaMalloc"); }
    void free(void* p) { cudaFree(p); }
    bool memset(void* p, int v, size_t bytes) { return ck(cudaMemsetAsync(p, v, bytes, stream), "cudaMemsetAsync"); }
    bool h2d(void* d, const void* h, size_t bytes)
    {
        return ck(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, stream), "h2d") && ck(cudaStreamSynchronize(stream), "sync");
    }
    bool d2h(void* h, const void* d, size_t bytes)
    {
        return ck(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, stream), "d2h") && ck(cudaStreamSynchronize(stream), "sync");
    }
    bool d2d(void* d, const void* s, size_t bytes) { return ck(cudaMemcpyAsync(d, s, bytes, cudaMemcpyDeviceToDevice, stream), "d2d"); }
    bool start_count(const MclStartArgs& a) { mcl_start_count_kernel<<<grid_for(a.n), 256, 0, stream>>>(a); return launched("mcl_start_count_kernel"); }
    bool start_fill(const MclStartArgs& a) { mcl_start_fill_kernel<<<grid_for(a.n), 256, 0, stream>>>(a); return launched("mcl_start_fill_kernel"); }
    bool plan(const MclPlanArgs& a) { mcl_plan_kernel<<<grid_for(a.n), 256, 0, stream>>>(a); return launched("mcl_plan_kernel"); }
    bool scan_i64(const int64_t* in, int64_t* out, int64_t count, int64_t base)
    {
        mcl_scan_kernel<int64_t><<<1, 1024, 0, stream>>>(in, out, count, base);
        return launched("mcl_scan_kernel");
    }
    bool scan_i32(const int32_t* in, int64_t* out, int64_t count, int64_t base)
    {
        mcl_scan_kernel<int32_t><<<1, 1024, 0, stream>>>(in, out, count, base);
        return launched("mcl_scan_kernel");
    }
    bool expand(const MclExpandArgs& a, int64_t ncols)
    {
        mcl_expand_kernel<<<(unsigned)ncols, MCL_NT, 0, stream>>>(a);               // columns with hashed accumulators
        if (!launched("mcl_expand_kernel")) return false;
        mcl_expand_direct_kernel<<<(unsigned)ncols, MCL_NT, 0, stream>>>(a);        // columns with row-indexed accumulators
        return launched("mcl_expand_direct_kernel");
    }
    bool inflate(const MclInflateArgs& a, int64_t ncols)
    {
        mcl_inflate_kernel<<<(unsigned)ncols, MCL_NT, 0, stream>>>(a);
        return launched("mcl_inflate_kernel");
    }
    bool dag(const MclDagArgs& a) { mcl_dag_kernel<<<grid_for(a.n), 256, 0, stream>>>(a); return launched("mcl_dag_kernel"); }
    bool attractor_round(const MclDagArgs& a) { mcl_attractor_kernel<<<grid_for(a.n), 256, 0, stream>>>(a); return launched("mcl_attractor_kernel"); }
    bool reach_round(const MclDagArgs& a) { mcl_reach_kernel<<<grid_for(a.n), 256, 0, stream>>>(a); return launched("mcl_reach_kernel"); }
};

}  // namespace

struct ofg_mcl {
    int device = 0;
    MclEngine<CudaBackend> eng;
    std::string err;
    std::string text;          // clusters file body of the last run
    bool have_graph = false, have_clusters = false;
    float ms_last = 0.f;       // device time of the last ofg_mcl_run (CUDA events on the library's stream)
    // staging of the input graph
    int64_t* in_cp = nullptr;
    int32_t* in_ri = nullptr;
    float* in_v = nullptr;
    void drop_input()
    {
        if (in_cp) cudaFree(in_cp);
        if (in_ri) cudaFree(in_ri);
        if (in_v) cudaFree(in_v);
        in_cp = nullptr; in_ri = nullptr; in_v = nullptr;
    }
};

static int mcl_fail(ofg_mcl* m, int code, const std::string& msg)
{
    m->err = msg;
    return code;
}

extern "C" {

int ofg_mcl_create(ofg_mcl** out, int device)
{
    if (!out) return OFG_ERR_ARG;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) return OFG_ERR_CUDA;
    if (cudaSetDevice(device) != cudaSuccess) return OFG_ERR_CUDA;
    ofg_mcl* m = new ofg_mcl();
    m->device = device;
    if (cudaStreamCreateWithFlags(&m->eng.be.stream, cudaStreamNonBlocking) != cudaSuccess) { delete m; return OFG_ERR_CUDA; }
    cudaDeviceGetAttribute(&m->eng.be.n_sm, cudaDevAttrMultiProcessorCount, device);
    *out = m;
    return OFG_OK;
}

void ofg_mcl_destroy(ofg_mcl* m)
{
    if (!m) return;
    cudaSetDevice(m->device);
    m->drop_input();
    m->eng.release();
    if (m->eng.be.stream) cudaStreamDestroy(m->eng.be.stream);
    delete m;
}

const char* ofg_mcl_last_error(ofg_mcl* m) { return m ? m->err.c_str() : "null handle"; }

int ofg_mcl_set_params(ofg_mcl* m, int prune_number, int select_number, int recover_number, double recover_pct, int sparse_overhead,
                       int max_iterations)
{
    if (!m) return OFG_ERR_ARG;
    if (prune_number < 0 || select_number < 0 || recover_number < 0 || !(recover_pct >= 0.0 && recover_pct <= 100.0) || sparse_overhead < 1 ||
        max_iterations < 1)
        return mcl_fail(m, OFG_ERR_ARG, "bad resource parameters");
    m->eng.par.P = prune_number; m->eng.par.S = select_number; m->eng.par.R = recover_number;
    m->eng.par.pct = recover_pct / 100.0; m->eng.par.overhead = sparse_overhead; m->eng.par.max_iter = max_iterations;
    return OFG_OK;
}

int ofg_mcl_set_budgets(ofg_mcl* m, int64_t table_slots, int64_t stage_entries, int direct_factor)
{
    if (!m || table_slots < 1 || stage_entries < 1 || direct_factor < 0) return OFG_ERR_ARG;
    m->eng.par.table_budget = table_slots;
    m->eng.par.stage_budget = stage_entries;
    m->eng.par.direct_factor = direct_factor;
    return OFG_OK;
}

int ofg_mcl_set_graph_host(ofg_mcl* m, int64_t n, const int64_t* colptr, const int32_t* rows, const float* vals)
{
    if (!m || n < 1 || n > 0x7ffffff0 || !colptr) return OFG_ERR_ARG;
    cudaSetDevice(m->device);
    const int64_t nnz = colptr[n];
    if (nnz < 0 || colptr[0] != 0 || (nnz && (!rows || !vals))) return mcl_fail(m, OFG_ERR_ARG, "bad column pointers");
    for (int64_t c = 0; c < n; c++) {
        if (colptr[c + 1] < colptr[c]) return mcl_fail(m, OFG_ERR_ARG, "column pointers must not decrease");
        for (int64_t p = colptr[c]; p < colptr[c + 1]; p++) {
            if (rows[p] < 0 || rows[p] >= n) return mcl_fail(m, OFG_ERR_ARG, "row index out of range");
            if (p > colptr[c] && rows[p] <= rows[p - 1]) return mcl_fail(m, OFG_ERR_ARG, "rows of a column must ascend");
            if (!(vals[p] >= 0.0f) || vals[p] > 3.0e38f) return mcl_fail(m, OFG_ERR_ARG, "weights must be finite and non-negative");
        }
    }
    m->drop_input();
    m->have_graph = m->have_clusters = false;
    CudaBackend& be = m->eng.be;
    if (!be.alloc((void**)&m->in_cp, (size_t)(n + 1) * 8) || !be.alloc((void**)&m->in_ri, (size_t)(nnz + 1) * 4) ||
        !be.alloc((void**)&m->in_v, (size_t)(nnz + 1) * 4))
        return mcl_fail(m, OFG_ERR_CUDA, be.e);
    if (!be.h2d(m->in_cp, colptr, (size_t)(n + 1) * 8) || (nnz && (!be.h2d(m->in_ri, rows, (size_t)nnz * 4) || !be.h2d(m->in_v, vals, (size_t)nnz * 4))))
        return mcl_fail(m, OFG_ERR_CUDA, be.e);
    if (!m->eng.start(n, m->in_cp, m->in_ri, m->in_v)) return mcl_fail(m, OFG_ERR_CUDA, m->eng.err);
    if (!be.ck(cudaStreamSynchronize(be.stream), "sync")) return mcl_fail(m, OFG_ERR_CUDA, be.e);
    m->drop_input();
    m->have_graph = true;
    return OFG_OK;
}

int ofg_mcl_set_graph_text(ofg_mcl* m, const char* text, size_t nbytes)
{
    if (!m || !text) return OFG_ERR_ARG;
    std::vector<int64_t> cp;
    std::vector<int32_t> ri;
    std::vector<float> v;
    int64_t n = 0;
    const std::string msg = mcl_parse_graph_text(text, nbytes, &n, cp, ri, v);
    if (!msg.empty()) return mcl_fail(m, OFG_ERR_PARSE, "graph file: " + msg);
    return ofg_mcl_set_graph_host(m, n, cp.data(), ri.data(), v.data());
}

int ofg_mcl_set_graph_dev(ofg_mcl* m, int64_t n, const uint64_t* row_off_dev, int64_t nnz, const int32_t* indices_dev, const double* weights_dev)
{
    if (!m || n < 1 || n > 0x7ffffff0 || nnz < 0 || !row_off_dev || (nnz && (!indices_dev || !weights_dev))) return OFG_ERR_ARG;
    cudaSetDevice(m->device);
    m->drop_input();
    m->have_graph = m->have_clusters = false;
    CudaBackend& be = m->eng.be;
    int* bad = nullptr;
    if (!be.alloc((void**)&m->in_cp, (size_t)(n + 1) * 8) || !be.alloc((void**)&m->in_v, (size_t)(nnz + 1) * 4) || !be.alloc((void**)&bad, 4))
        return mcl_fail(m, OFG_ERR_CUDA, be.e);
    be.memset(bad, 0, 4);
    mcl_rowptr_kernel<<<be.grid_for(n + 1), 256, 0, be.stream>>>((const unsigned long long*)row_off_dev, n, nnz, m->in_cp);
    if (nnz) mcl_weights_kernel<<<be.grid_for(nnz), 256, 0, be.stream>>>(weights_dev, m->in_v, nnz, bad);
    if (!be.launched("mcl_weights_kernel")) { cudaFree(bad); return mcl_fail(m, OFG_ERR_CUDA, be.e); }
    int hbad = 0;
    const bool okc = be.d2h(&hbad, bad, 4);
    cudaFree(bad);
    if (!okc) return mcl_fail(m, OFG_ERR_CUDA, be.e);
    if (hbad) return mcl_fail(m, OFG_ERR_ARG, "a graph weight is negative, not finite or too large to print");
    if (!m->eng.start(n, m->in_cp, indices_dev, m->in_v)) return mcl_fail(m, OFG_ERR_CUDA, m->eng.err);
    if (!be.ck(cudaStreamSynchronize(be.stream), "sync")) return mcl_fail(m, OFG_ERR_CUDA, be.e);
    m->drop_input();
    m->have_graph = true;
    return OFG_OK;
}

int ofg_mcl_step(ofg_mcl* m, double inflation, double* chaos)
{
    if (!m) return OFG_ERR_ARG;
    if (!m->have_graph) return mcl_fail(m, OFG_ERR_ARG, "no graph set");
    if (!(inflation > 0.0)) return mcl_fail(m, OFG_ERR_ARG, "inflation must be positive");
    cudaSetDevice(m->device);
    if (!m->eng.step(inflation, chaos)) return mcl_fail(m, OFG_ERR_CUDA, m->eng.err);
    return OFG_OK;
}

int ofg_mcl_interpret(ofg_mcl* m)
{
    if (!m) return OFG_ERR_ARG;
    if (!m->have_graph) return mcl_fail(m, OFG_ERR_ARG, "no graph set");
    cudaSetDevice(m->device);
    if (!m->eng.interpret()) return mcl_fail(m, OFG_ERR_CUDA, m->eng.err);
    m->text = mcl_clusters_text(m->eng.labels, m->eng.n_clusters);
    m->have_clusters = true;
    return OFG_OK;
}

int ofg_mcl_run(ofg_mcl* m, double inflation)
{
    if (!m) return OFG_ERR_ARG;
    if (!m->have_graph) return mcl_fail(m, OFG_ERR_ARG, "no graph set");
    if (!(inflation > 0.0)) return mcl_fail(m, OFG_ERR_ARG, "inflation must be positive");
    cudaSetDevice(m->device);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0, m->eng.be.stream);
    const bool ok = m->eng.run(inflation);
    cudaEventRecord(e1, m->eng.be.stream);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&m->ms_last, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (!ok) return mcl_fail(m, OFG_ERR_CUDA, m->eng.err);
    m->text = mcl_clusters_text(m->eng.labels, m->eng.n_clusters);
    m->have_clusters = true;
    return OFG_OK;
}

int ofg_mcl_result(ofg_mcl* m, int64_t* iterations, int64_t* n_clusters, int64_t* n_overlap, double* last_chaos, int64_t* nnz, float* ms,
                   int64_t* n_launches)
{
    if (!m) return OFG_ERR_ARG;
    if (iterations) *iterations = m->eng.iterations;
    if (n_clusters) *n_clusters = m->have_clusters ? m->eng.n_clusters : -1;
    if (n_overlap) *n_overlap = m->have_clusters ? m->eng.n_overlap : -1;
    if (last_chaos) *last_chaos = m->eng.last_chaos;
    if (nnz) *nnz = m->eng.cur.nnz;
    if (ms) *ms = m->ms_last;
    if (n_launches) *n_launches = m->eng.launches;
    return OFG_OK;
}

int ofg_mcl_labels(ofg_mcl* m, int32_t* labels, int64_t n)
{
    if (!m || !labels) return OFG_ERR_ARG;
    if (!m->have_clusters) return mcl_fail(m, OFG_ERR_ARG, "no clustering yet");
    if (n != m->eng.n) return mcl_fail(m, OFG_ERR_ARG, "label buffer size differs from the number of nodes");
    memcpy(labels, m->eng.labels.data(), (size_t)n * 4);
    return OFG_OK;
}

int ofg_mcl_clusters_text(ofg_mcl* m, char* buf, size_t cap, size_t* need)
{
    if (!m) return OFG_ERR_ARG;
    if (!m->have_clusters) return mcl_fail(m, OFG_ERR_ARG, "no clustering yet");
    if (need) *need = m->text.size();
    if (buf) {
        if (cap < m->text.size()) return mcl_fail(m, OFG_ERR_ARG, "text buffer too small");
        memcpy(buf, m->text.data(), m->text.size());
    }
    return OFG_OK;
}

int ofg_mcl_iterand_copy(ofg_mcl* m, int64_t* colptr, int32_t* rows, float* vals)
{
    if (!m) return OFG_ERR_ARG;
    if (!m->have_graph) return mcl_fail(m, OFG_ERR_ARG, "no graph set");
    cudaSetDevice(m->device);
    if (!m->eng.copy_iterand(colptr, rows, vals)) return mcl_fail(m, OFG_ERR_CUDA, m->eng.err);
    return OFG_OK;
}

}  // extern "C"
