// Device build of the synthetic hit-file generator (synth_hits.h): writes Blast{i}_{j}.txt for every
// pair of the context straight into the text arena.  Two passes (size, write), one thread per
// (pair, query); byte-identical to libofg_synth.so (host build of the same header).
#pragma once
#include "ofg_common.cuh"
#include "synth_hits.h"

struct SynthArgs {
    OfgSynthParams P;
    const PairDesc* pairs;
    const int64_t* pair_row_base;
    int n_pairs;
    int64_t n_rows;
    const int32_t* species_ids;
    const double* lengths;
    u32* row_bytes;            // [n_rows] size pass output
    const u64* row_off;        // [n_rows + 1] exclusive scan of row_bytes
    const int64_t* file_off;   // [n_pairs] arena offset of each file
    uint8_t* text;
};

template <bool WRITE>
__global__ void __launch_bounds__(128) synth_kernel(SynthArgs a)
{
    char buf[2 * OFG_SYNTH_MAX_LINE];
    for (int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; row < a.n_rows; row += (int64_t)gridDim.x * blockDim.x) {
        const int pair = upper_bound_dev(a.pair_row_base, a.n_pairs + 1, row) - 1;
        const PairDesc& pd = a.pairs[pair];
        const int64_t q = row - pd.row_base;
        const double* Li = a.lengths + pd.li_off;
        const double* Lj = a.lengths + pd.lj_off;
        const double* Llo = pd.p < pd.j ? Li : Lj;
        const int64_t Glo = pd.p < pd.j ? pd.Gi : pd.Gj;
        const int ns = ofg_synth_num_slots(a.P, q);
        u32 total = 0;
        uint8_t* dst = nullptr;
        if (WRITE) dst = a.text + a.file_off[pair] + (a.row_off[row] - a.row_off[pd.row_base]);
        for (int k = 0; k < ns; k++) {
            const int n = ofg_synth_slot(a.P, pd.p, pd.j, a.species_ids[pd.p], a.species_ids[pd.j], pd.Gi, pd.Gj, Li, Lj,
                                         Llo, Glo, q, k, buf);
            if (WRITE)
                for (int t = 0; t < n; t++) dst[total + t] = (uint8_t)buf[t];
            total += (u32)n;
        }
        if (!WRITE) a.row_bytes[row] = total;
    }
}
