// Host build of the synthetic hit-file generator (same bytes as the device kernels in ofg_synth.cu).
#include "synth_hits.h"
#include <string.h>

extern "C" {

// Bytes and lines of Blast{si}_{sj}.txt; if out != nullptr the text is written there (cap bytes).
int64_t ofg_synth_pair_host(const OfgSynthParams* P, int pi, int pj, int si, int sj, int64_t Gi, int64_t Gj,
                            const double* Li, const double* Lj, char* out, int64_t cap, int64_t* n_lines)
{
    const double* Llo = pi < pj ? Li : Lj;
    int64_t Glo = pi < pj ? Gi : Gj;
    char buf[2 * OFG_SYNTH_MAX_LINE];
    int64_t total = 0, lines = 0;
    for (int64_t q = 0; q < Gi; q++) {
        int ns = ofg_synth_num_slots(*P, q);
        for (int k = 0; k < ns; k++) {
            int n = ofg_synth_slot(*P, pi, pj, si, sj, Gi, Gj, Li, Lj, Llo, Glo, q, k, buf);
            if (out) {
                if (total + n > cap) return -1;
                memcpy(out + total, buf, (size_t)n);
            }
            for (int t = 0; t < n; t++) lines += (buf[t] == '\n');
            total += n;
        }
    }
    if (n_lines) *n_lines = lines;
    return total;
}

}  // extern "C"
