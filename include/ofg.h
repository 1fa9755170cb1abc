/* libofg — B200-native orthogroup-graph builder: C-ABI of the WaterfallMethod hot path.
 *
 * The reference (davidemms/OrthoFinder, scripts_of/) has no FFI for this path; the seam is the Python
 * call site gathering.py:476-512 inside DoOrthogroups.  Every entry point below cites
 * the reference code whose work it replaces.  Conventions: plain pointers and sizes only, return 0 on
 * success or a negative OFG_ERR_* code (message via ofg_last_error); the library never exits the
 * process; buffers handed out stay owned by the context; one context per GPU, not thread-safe.
 * Species are addressed by their position in seqsInfo.speciesToUse ("list position"), exactly as
 * iSpecies/jSpecies in gathering.py:177-196.
 */
#ifndef OFG_H
#define OFG_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ofg_ctx ofg_ctx;

enum {
    OFG_OK = 0,
    OFG_ERR_CUDA = -1,         /* CUDA runtime failure */
    OFG_ERR_ARG = -2,          /* bad argument / call order */
    OFG_ERR_PARSE = -3,        /* malformed hit line: blast_file_processor.py:74-82,99-103 */
    OFG_ERR_INCONSISTENT = -4, /* sequence id >= nSeqs: blast_file_processor.py:89-98 */
    OFG_ERR_FIT = -5,          /* curve_fit would raise "Optimal parameters not found" (gathering.py:120) */
    OFG_ERR_CAPACITY = -6,     /* an internal capacity was exceeded; see message */
    OFG_ERR_UNSUPPORTED = -7,  /* input outside the supported grammar / ranges; see message */
    OFG_ERR_MISSING = -8       /* a hit file required by the partition was not provided */
};

/* kinds of offending line reported with OFG_ERR_PARSE / OFG_ERR_INCONSISTENT */
enum {
    OFG_LINE_OK = 0,
    OFG_LINE_BAD_ID = 1,     /* "Query or hit sequence ID ... missing or incorrectly formatted" (:75) */
    OFG_LINE_BAD_SCORE = 2,  /* "12th field ... should be the bit-score" (:81) */
    OFG_LINE_QUOTE = 3,      /* csv quoting: not produced by BLAST/DIAMOND/MMseqs2, rejected */
    OFG_LINE_RANGE_Q = 4,    /* query id >= nSeqs_i (:89-98) */
    OFG_LINE_RANGE_S = 5,    /* hit id >= nSeqs_j (:89-98) */
    OFG_LINE_SCORE_FORMAT = 6 /* float() would accept it but the device grammar does not (inf/nan/>19 digits) */
};

/* stages of ofg_build(stop_after) — each one ends where the reference writes a pickle or the graph */
enum {
    OFG_STAGE_RAW = 1,      /* GetBLAST6Scores for every pair (blast_file_processor.py:38-104) */
    OFG_STAGE_NORM = 2,     /* NormaliseScores + GetBH_s -> "B", "BH" (gathering.py:140-155,213-248) */
    OFG_STAGE_THRESH = 3,   /* RBH + GetMostDistant_s for owned species (gathering.py:251-259,294-310) */
    OFG_STAGE_CONNECT = 4,  /* ConnectAllBetterThanCutoff_s -> "connect" (gathering.py:391-408) */
    OFG_STAGE_GRAPH = 5     /* WriteGraph_perSpecies: (connect + connect^T) .* B, CSR + text (gathering.py:23-49) */
};

/* flag bits stored in the top of the column word returned by ofg_pair_csr */
#define OFG_FLAG_BH 0x80000000u      /* entry of BH_pj   (gathering.py:213-248) */
#define OFG_FLAG_CONNECT 0x40000000u /* entry of connect_pj (gathering.py:391-408) */
#define OFG_FLAG_REVERSE 0x20000000u /* connect_jp[c,r] is set (gathering.py:30, m2tr) */
#define OFG_FLAG_TBH 0x10000000u     /* BH_jp[c,r] is set, imported from the context that owns species j */
#define OFG_COL_MASK 0x0fffffffu

/* synthetic hit-file parameters (orthofinder_b200/csrc/synth_hits.h) */
typedef struct {
    uint64_t seed;
    int32_t hq, integer_scores, ident_levels, p_dup_permille, p_missing_permille, p_nohit_permille,
        p_selfonly_permille, p_twin_permille, single_i, single_j;
} ofg_synth_params;

/* Create / destroy a context bound to CUDA device `device`. */
int ofg_create(ofg_ctx** out, int device);
void ofg_destroy(ofg_ctx* ctx);
const char* ofg_last_error(const ofg_ctx* ctx);

/* Index space and sequence lengths: util.SequencesInfo (util.py:54-84) and
 * gathering.GetSequenceLengths (gathering.py:444-464).  n_seqs[p] genes for list position p,
 * lengths_concat holds the S length vectors back to back (fp64, integer valued). */
int ofg_set_species(ofg_ctx* ctx, int n_species, const int64_t* n_seqs, const double* lengths_concat);

/* Query species (rows of the S x S pair grid) owned by this context: the unit the reference
 * parallelises over (gathering.py:484-485).  Default: all.  With with_columns != 0 the context
 * also ingests Blast{j}_{p} for every owned p (the transposed direction the reference reads back
 * from pickles at gathering.py:254 and :30), so that only per-gene thresholds need exchanging. */
int ofg_set_partition(ofg_ctx* ctx, int n_owned, const int32_t* owned_rows, int with_columns);

/* Hit text of Blast{p}_{j}.txt (already decompressed): replaces the open()/csv.reader loop of
 * blast_file_processor.py:59-88.  `ptr` is a host pointer (copied H2D here) or, with is_device != 0,
 * a device pointer that is copied device-to-device into the context's text arena.  swap_cols != 0
 * reads query/hit from columns 1/0 (the -1 one-way mode, blast_file_processor.py:41-54).
 * A pair that is never added is an empty file (q_allow_empty, :62-63) only if allow_empty was set
 * with ofg_set_option; otherwise ofg_build fails with OFG_ERR_MISSING. */
int ofg_add_hits_text(ofg_ctx* ctx, int p, int j, const void* ptr, size_t nbytes, int swap_cols, int is_device);

/* The same for a compressed file: `host_ptr` holds the bytes of Blast{p}_{j}.txt.gz as they are on disk (what the reference
 * opens with gzip.open, blast_file_processor.py:63; DIAMOND writes its hit files compressed, config.json:48-52).  Only the
 * compressed bytes are copied to the device; ofg_build inflates them there, one warp per gzip member.  Files whose members
 * carry their compressed size in an extra field (bgzip / BGZF "BC", or the 32-bit "OF" field of this repo's writer) are
 * inflated by as many warps as they have members; a plain single-member file is one warp's work.  Several members without
 * size hints are OFG_ERR_UNSUPPORTED (inflate on the host and use ofg_add_hits_text).  A corrupt stream or a wrong ISIZE
 * fails the build with OFG_ERR_PARSE; the CRC-32 is not verified. */
int ofg_add_hits_gz(ofg_ctx* ctx, int p, int j, const void* host_ptr, size_t nbytes, int swap_cols);

/* Drop all hit text (keeps species / partition). */
int ofg_clear_hits(ofg_ctx* ctx);

/* Generate Blast{species_ids[p]}_{species_ids[j]}.txt for every pair this context needs directly in
 * device memory (bench / tests); byte-identical to the host generator for the same parameters. */
int ofg_synth_hits(ofg_ctx* ctx, const ofg_synth_params* params, const int32_t* species_ids);
/* Size of / copy out one generated (or added) file, for tests and for the host-buffer e2e bench. */
int ofg_hits_text_size(ofg_ctx* ctx, int p, int j, size_t* nbytes);
int ofg_hits_text_copy(ofg_ctx* ctx, int p, int j, void* host_dst, size_t cap);

/* Options: "allow_empty" (0/1), "max_bytes_per_line_inv" (capacity: records <= bytes/this, default 24; after a
 * complete build the context sizes the next one from the line density it saw),
 * "sel_capacity_frac_inv" (selection capacity = records/this, default 4), "reserve_text_bytes" (size the text arena
 * once), "reserve_gz_bytes" (the same for the compressed arena of ofg_add_hits_gz), "wave_species" (RAW + NORM run over this many query species at a time out of one set of scratch buffers -
 * the reference's "one query species per task", gathering.py:484-485; 0 = the whole partition at once),
 * "rowsort_short" (-1 = decide per wave from the text bytes per row (default), 0 = never, 1 = always: rows of at most 16 records are
 * sorted by half a warp each),
 * "length_sort_radix" (0/1: build the length bins with the segmented LSD radix sort only - the fallback of the bucket path),
 * "async" (0/1: ofg_build only enqueues; counts, errors and sizes are read back by ofg_sync or by the first accessor),
 * "scores_v2" (0/1): the `--scores-v2` variant of the path - NormalisedBitScore (gathering.py:157-174)
 * instead of the length-bin fit and GetMostDistant_s_estimate_from_relative_rbhs (gathering.py:314-388) instead of
 * GetMostDistant_s; what the reference passes as v2_scores to ProcessBlastHits / ConnectCognates (:177, :251),
 * "homology_graph" (0/1): the `--c-homologs` graph of gathering.py:51-73 instead of the connect graph,
 * "keep_self_hits" (0/1): GetBLAST6Scores(..., qExcludeSelfHits=False) - the raw matrices of DendroBLAST
 * (orthologues.py:272-274); only OFG_STAGE_RAW can be built with it (see ofg_dendro_matrices). */
int ofg_set_option(ofg_ctx* ctx, const char* name, int64_t value);

/* Optional, only read with "scores_v2": the per-gene factors Lengths[p] ** -0.5 of NormalisedBitScore
 * (gathering.py:171-172) as the caller's numpy computes them, laid out like lengths_concat.  numpy's pow is
 * not correctly rounded (SVML), so handing its values over makes the v2 graph byte-identical to the
 * reference's on the same host; without this call the device evaluates pow(L, -0.5) itself (within 2 ulp).
 * Reset by ofg_set_species. */
int ofg_set_length_factors(ofg_ctx* ctx, const double* factors_concat);

/* Run the path up to and including `stop_after` (OFG_STAGE_*).  OFG_STAGE_THRESH stops before the
 * thresholds of other ranks are needed; call ofg_thresholds_get / _set around the exchange and then
 * ofg_build(ctx, OFG_STAGE_GRAPH) to finish.  Single GPU: ofg_build(ctx, OFG_STAGE_GRAPH).  A later stage
 * includes every earlier one that has not run yet. */
int ofg_build(ofg_ctx* ctx, int stop_after);

/* Waits for everything enqueued on the context's stream and reads the results of the build(s) back (one host
 * synchronisation: line counts, parse errors, capacity flags, fit table, graph sizes).  Returns what a synchronous
 * ofg_build would have returned.  Every accessor below calls it implicitly. */
int ofg_sync(ofg_ctx* ctx);
/* Statistics of the last build: "radix_fallback_waves" (waves whose length bins were built by the LSD radix sort because
 * the bucket path could not take them), "waves", "record_slot_capacity". */
int ofg_get_stat(ofg_ctx* ctx, const char* name, int64_t* value);

/* DendroBLAST distance matrices of orthogroups (replaces the worker
 * orthologues.py:264-291 `Worker_OGMatrices_ReadBLASTAndUpdateDistances` and, with complete != 0, the completion loop of
 * `DendroBLASTTrees.CompleteAndWriteOGMatrices` :396-410).  The context must be built to OFG_STAGE_RAW and no further, with
 * option "keep_self_hits" = 1 (GetBLAST6Scores(..., qExcludeSelfHits=False), orthologues.py:272-274), every species pair in
 * one context.
 *   og_off[n_ogs + 1]  first member of every orthogroup in species[] / iseq[] (og_off[0] = 0)
 *   species[k], iseq[k] member k: species list position and sequence index (util.py:54-61)
 *   complete == 0:     M_ij = 0.5 * max(B_ij, Bmin_i) * (1 / Bmax_i), bit-identical to the reference
 *   complete != 0:     D_ij = D_ji = -log(M_ij + M_ji) for i != j (CUDA log, <= 1 ulp from numpy's), D_ii = M_ii;
 *                      host_max_og[g] (may be NULL) = max over j < i of D_ij, -9e99 if there is none
 *   host_out           sum over orthogroups of n^2 doubles, orthogroup after orthogroup, row-major; out_len = its capacity
 * Errors: OFG_ERR_ARG for a gene that does not exist, a wrong stage, a partitioned context or a short output. */
int ofg_dendro_matrices(ofg_ctx* ctx, int64_t n_ogs, const int64_t* og_off, const int32_t* species, const int32_t* iseq,
                        int complete, double* host_out, int64_t out_len, double* host_max_og);
/* Orders the context's stream after everything enqueued so far on `other`'s stream (event wait, no host block). */
int ofg_wait_for(ofg_ctx* ctx, ofg_ctx* other);

/* Per-gene cut-off "mostDistant" (gathering.py:294-310), global gene indexing (seqStartingIndices).
 * _dev returns the device array of nSeqs doubles (entries of species not owned hold +inf until set);
 * _set_dev tells the library that every entry is now valid (after the all-gather). */
int ofg_thresholds_dev(ofg_ctx* ctx, double** dev_ptr, int64_t* n);
int ofg_thresholds_mark_complete(ofg_ctx* ctx);
int ofg_thresholds_copy(ofg_ctx* ctx, double* host_dst, int64_t n);

/* Sparse exchange between contexts that own disjoint query species and do NOT ingest column pairs
 * (design B of SURVEY.md 8e): instead of re-reading Blast{j}_{p}, the owner of species j ships the index
 * lists the reference reads back from the transposed pickles — BH_jp (matrices.LoadMatrixArray(...,
 * row=False), gathering.py:254) after OFG_STAGE_NORM, connect_jp (gathering.py:30) after OFG_STAGE_CONNECT.
 * which = 0: best hits, 1: connect.  species_rank[S] gives the owner rank of every species.  The list holds
 * one (gene of the remote species, gene of the local species) pair of global gene ids (2 x int32) per flagged
 * entry, grouped by destination rank; counts[r] entries go to rank r.  The device buffer stays owned by the
 * context until the next export.  ofg_import_flags marks the received entries (TBH / REVERSE flags). */
int ofg_export_flags(ofg_ctx* ctx, int which, int n_ranks, const int32_t* species_rank, int my_rank, int64_t* counts,
                     int32_t** dev_list);
int ofg_import_flags(ofg_ctx* ctx, int which, const int32_t* dev_list, int64_t n_entries);

/* Peer variant of the same exchange, without host round trips: every context owns an INBOX of n_ranks segments of
 * seg_bytes ([u64 count][count x (int32, int32)] per source rank).  ofg_export_flags_peer computes all offsets on
 * the device and writes this context's lists straight into segment `my_rank` of each destination's inbox
 * (inbox_ptrs[r]: device pointer valid on this GPU - another context of the same process, or a peer GPU's buffer
 * opened with ofg_ipc_open_handle: the stores travel over NVLink while the kernel runs).  After all exporters are
 * done (ofg_wait_for between contexts of a process, a barrier between ranks) ofg_import_flags_peer marks every
 * received entry.  A segment that is too small sets OFG_ERR_CAPACITY at the next ofg_sync; enlarge and repeat. */
int ofg_export_flags_peer(ofg_ctx* ctx, int which, int n_ranks, const int32_t* species_rank, int my_rank,
                          void* const* inbox_ptrs, size_t seg_bytes);
int ofg_import_flags_peer(ofg_ctx* ctx, int which, int n_ranks, int my_rank, const void* inbox, size_t seg_bytes);
/* Plumbing for the inboxes: zeroed device allocations on the context's GPU and CUDA IPC handles (64 bytes) so that
 * the ranks of one node can map each other's inbox. */
int ofg_device_alloc(ofg_ctx* ctx, size_t bytes, void** dev_ptr_out);
int ofg_device_free(ofg_ctx* ctx, void* dev_ptr);
int ofg_ipc_get_handle(ofg_ctx* ctx, void* dev_ptr, void* handle_out64);
int ofg_ipc_open_handle(ofg_ctx* ctx, const void* handle64, void** dev_ptr_out);
int ofg_ipc_close_handle(ofg_ctx* ctx, void* dev_ptr);

/* Totals after a build: lines = non-empty records seen, records = lines kept after the self-hit and
 * score > 0 filters, nnz = stored (q,s) after max-dedup, edges = entries of the final graph. */
int ofg_counts(ofg_ctx* ctx, int64_t* lines, int64_t* records, int64_t* nnz, int64_t* edges);

/* Introspection for the parity gates: matrix of pair (p,j) as row lengths + packed
 * (col | flags, value) entries in row-major / column-ascending order.  Values are raw bit scores
 * after OFG_STAGE_RAW and normalised scores afterwards ("B" pickles).  Pass NULL to query sizes. */
int ofg_pair_csr(ofg_ctx* ctx, int p, int j, int64_t* n_rows, int64_t* nnz, int32_t* row_len, uint32_t* cols,
                 double* vals);
/* out[0]=a, out[1]=b (curve_fit parameters, gathering.py:119-121), out[2]=number of selected hits,
 * out[3]=MINPACK info, out[4]=nfev, out[5]=1 if normalised / 0 if the pair was zeroed (:152-155). */
int ofg_pair_fit(ofg_ctx* ctx, int p, int j, double out[6]);

/* Final graph W = (connect_pj + connect_jp^T) .* B_pj for the owned rows (gathering.py:23-49):
 * CSR with global gene indices, rows [row0, row0 + n_rows) contiguous per owned species in list order. */
int ofg_graph_csr_size(ofg_ctx* ctx, int64_t* n_rows, int64_t* nnz);
int ofg_graph_csr_copy(ofg_ctx* ctx, int64_t* row_ids, int64_t* indptr, int32_t* indices, double* data);
/* MCL text of the owned rows ("%d    " + "%d:%.3f " ... + "$\n", gathering.py:39-48), in device memory;
 * header (:279-280) and the closing ")\n" (:48) are added by the host wrapper. */
int ofg_graph_text_size(ofg_ctx* ctx, size_t* nbytes);
int ofg_graph_text_copy(ofg_ctx* ctx, void* host_dst, size_t cap);

/* Device pointers of the finished graph of a context that owns every species (single GPU): row_off[g] = first edge of
 * gene g (n_rows entries; the end of the last row is nnz), global column indices, fp64 weights.  Owned by ctx, valid
 * until the next build or ofg_destroy; everything enqueued on the context's stream has completed on return.  This is
 * what ofg_mcl_set_graph_dev takes, so the graph never leaves the device between WriteGraphParallel's replacement and
 * RunMCL's (gathering.py:512 -> :517). */
int ofg_graph_dev(ofg_ctx* ctx, int64_t* n_rows, int64_t* nnz, const uint64_t** row_off_dev, const int32_t** indices_dev,
                  const double** weights_dev);

/* ---- MCL on the device ------------------------------------------------------------------
 * Replaces the external process of MCL.RunMCL (scripts_of/mcl.py:214-218: `mcl graphFilename -I inflation -o
 * clustersFilename -te nProcesses -V all`, bundled binary scripts_of/bin/mcl = mcl 14-137 with its default resource
 * scheme: prune 10000, select 1100, recover 1400, recover percentage 90).  The result is the body of the clusters file
 * the binary writes - what GetPredictedOGs (mcl.py:36-63) and ConvertSingleIDsToIDPair (mcl.py:93-125) read - and the
 * per-gene cluster index.  One handle per GPU; calls on one handle are not thread-safe.  Every entry point returns
 * OFG_OK or a negative OFG_ERR_* code, message via ofg_mcl_last_error. */
typedef struct ofg_mcl ofg_mcl;
int ofg_mcl_create(ofg_mcl** out, int device);
void ofg_mcl_destroy(ofg_mcl* m);
const char* ofg_mcl_last_error(ofg_mcl* m);
/* the binary's -P, -S, -R, -pct (percent), -sparse and -L options; defaults = the binary's */
int ofg_mcl_set_params(ofg_mcl* m, int prune_number, int select_number, int recover_number, double recover_pct, int sparse_overhead,
                       int max_iterations);
/* scratch budgets per batch of columns: slots of global-memory hash tables / row-indexed accumulators (12 bytes each), stage
 * entries (8 bytes each); direct_factor: a column that may touch n / direct_factor rows or more accumulates in a row-indexed
 * array instead of a hash table (default 4; 0 = never).  Results do not depend on any of the three. */
int ofg_mcl_set_budgets(ofg_mcl* m, int64_t table_slots, int64_t stage_entries, int direct_factor);
/* the graph as the binary reads it from a file: column c = line c of the graph file, rows ascending, the float values
 * of the printed weights (host arrays; entries that read as 0 and loops are dropped like the binary does) */
int ofg_mcl_set_graph_host(ofg_mcl* m, int64_t n, const int64_t* colptr, const int32_t* rows, const float* vals);
/* the same from the bytes of the graph FILE (what RunMCL's graphFilename holds): parsed on the host like the binary reads it
 * (strtod, then float); OFG_ERR_PARSE with the reason for anything that is not a square MCL matrix with ascending ids */
int ofg_mcl_set_graph_text(ofg_mcl* m, const char* text, size_t nbytes);
/* the graph of ofg_graph_dev, on the same device: weights go through "%.3f" -> strtod -> float on the device, exactly the
 * values the binary would have read from OrthoFinder_graph.txt */
int ofg_mcl_set_graph_dev(ofg_mcl* m, int64_t n, const uint64_t* row_off_dev, int64_t nnz, const int32_t* indices_dev,
                          const double* weights_dev);
/* the whole process: expansion / inflation until the chaos falls below 1e-4, then the clusters */
int ofg_mcl_run(ofg_mcl* m, double inflation);
/* one iteration / the clusters of the current iterand (parity tests against the binary's -dump ite files) */
int ofg_mcl_step(ofg_mcl* m, double inflation, double* chaos);
int ofg_mcl_interpret(ofg_mcl* m);
int ofg_mcl_iterand_copy(ofg_mcl* m, int64_t* colptr, int32_t* rows, float* vals);   /* nnz from ofg_mcl_result */
/* any pointer may be NULL; n_clusters / n_overlap are -1 before the first clustering; ms = device time of the last run */
int ofg_mcl_result(ofg_mcl* m, int64_t* iterations, int64_t* n_clusters, int64_t* n_overlap, double* last_chaos, int64_t* nnz, float* ms,
                   int64_t* n_launches);
/* labels[g] = index of gene g's cluster in the order of the clusters file (size descending, then smallest member) */
int ofg_mcl_labels(ofg_mcl* m, int32_t* labels, int64_t n);
/* the clusters file from "(mclheader" on, byte for byte what the binary writes after its "# cline:" comment line */
int ofg_mcl_clusters_text(ofg_mcl* m, char* buf, size_t cap, size_t* need);

/* Kernel timing of the last ofg_build (CUDA events on the library's stream), milliseconds per group:
 * 0 parse, 1 rows (bucket + row sort), 2 length sort, 3 select, 4 fit, 5 scale+BH, 6 thresholds,
 * 7 connect, 8 emit, 9 total.  n_launches = kernels launched by the last build. */
int ofg_last_timing(ofg_ctx* ctx, float ms[10], int64_t* n_launches);
/* Per-kernel device time of the last build: one text line "name launches total_ms" per kernel, measured
 * with CUDA events recorded after every launch on the library's stream (host-sync gaps excluded). */
int ofg_kernel_report(ofg_ctx* ctx, char* buf, size_t cap);
/* Stream the library launches on (cudaStream_t), so callers can record their own events on it. */
void* ofg_stream(ofg_ctx* ctx);

/* Standalone device functions exposed for unit parity tests (each runs one small kernel). */
int ofg_test_log10(ofg_ctx* ctx, const double* host_x, double* host_y, int64_t n);      /* np.log10 emulation */
int ofg_test_format(ofg_ctx* ctx, const double* host_x, char* host_out, int64_t n);     /* "%.3f", 32 B slots */
int ofg_test_parse_score(ofg_ctx* ctx, const char* host_text, int64_t nbytes, double* host_y, int32_t* host_ok,
                         int64_t n_fields);                                             /* '\n'-separated */
int ofg_test_lmfit(ofg_ctx* ctx, const double* host_lx, const double* host_ly, int64_t m, double out[4]);

#ifdef __cplusplus
}
#endif
#endif
