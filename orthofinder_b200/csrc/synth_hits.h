// Synthetic all-vs-all hit files (BLAST outfmt-6 / DIAMOND tabular) for parity tests and bench.
//
// One source for host and device: the functions below are compiled by g++ into
// libofg_synth.so (small configs written to disk, fed to the live reference and the oracle) and by
// nvcc into libofg.so (the bench configs are generated straight into HBM), and both produce the
// same bytes for the same parameters.  Only integer arithmetic is used so that this holds.
//
// Shape follows SURVEY.md §8(d): gene g of every species belongs to family g; for pair (i,j) each
// query q gets its family partner plus hq-1 neighbours at symmetric offsets (so that the reverse
// file holds the reciprocal line), scores are a shared per-(gene pair) identity times
// min(Lq, Ls) plus a small per-direction jitter, ~p_dup of (q,s) get a second lower-scoring HSP
// line.  Stress knobs: integer scores with quantised identities (exact ties), missing reciprocal
// lines, genes without hits, genes with hits only inside their species, "twin" genes whose two
// targets land inside the 1e-3 best-hit band with different normalised values, one pair with a
// single line (the "too few hits" branch of gathering.py:152-155).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define OFG_HD __host__ __device__ __forceinline__
#else
#define OFG_HD static inline
#endif

struct OfgSynthParams {
    uint64_t seed;
    int32_t hq;                  // target lines per query per pair (before duplicates)
    int32_t integer_scores;      // 0: one decimal (DIAMOND style); 1: integer bit scores (MMseqs2 style)
    int32_t ident_levels;        // 0: continuous identities; >0: identity in {0.20 + 0.01*k, k < levels}
    int32_t p_dup_permille;      // second, lower-scoring HSP line for a (q,s)
    int32_t p_missing_permille;  // line dropped when species_i > species_j (missing reciprocal)
    int32_t p_nohit_permille;    // genes that get no line at all as a query
    int32_t p_selfonly_permille; // genes with lines only against their own species
    int32_t p_twin_permille;     // families (2t, 2t+1) with near-tied cross hits
    int32_t single_i, single_j;  // species pair that holds exactly one line (-1: none)
};

#define OFG_SYNTH_MAX_LINE 160

OFG_HD uint64_t ofg_mix64(uint64_t z)
{
    z += 0x9e3779b97f4a7c15ULL;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}
OFG_HD uint64_t ofg_hash4(uint64_t seed, uint64_t a, uint64_t b, uint64_t c, uint64_t d)
{
    return ofg_mix64(ofg_mix64(ofg_mix64(ofg_mix64(seed ^ a) ^ b) ^ c) ^ d);
}

OFG_HD int ofg_put_uint(char* out, uint64_t v)
{
    char tmp[24];
    int n = 0;
    do { tmp[n++] = (char)('0' + (v % 10)); v /= 10; } while (v);
    for (int k = 0; k < n; k++) out[k] = tmp[n - 1 - k];
    return n;
}

OFG_HD int ofg_synth_is_twin(const OfgSynthParams& P, int64_t q)
{
    return P.p_twin_permille > 0 &&
           (int)(ofg_hash4(P.seed, 0x7717, (uint64_t)(q >> 1), 3, 5) % 1000) < P.p_twin_permille;
}

// number of candidate lines of query q in pair (i,j) before drop/dup decisions
OFG_HD int ofg_synth_num_slots(const OfgSynthParams& P, int64_t q) { return P.hq + (ofg_synth_is_twin(P, q) ? 1 : 0); }

// Writes the lines of (species pair (si,sj) with list positions (pi,pj), query q, slot k) into
// out (>= 2*OFG_SYNTH_MAX_LINE bytes); returns the number of bytes (0 if the slot emits nothing).
OFG_HD int ofg_synth_slot(const OfgSynthParams& P, int pi, int pj, int si, int sj, int64_t Gi, int64_t Gj,
                          const double* Li, const double* Lj, const double* Llo, int64_t Glo, int64_t q, int k,
                          char* out)
{
    if (P.single_i == pi && P.single_j == pj) {
        if (!(q == (7 % Gi) && k == 0)) return 0;
    }
    // query-level drops
    uint64_t hq_ = ofg_hash4(P.seed, 0x51, (uint64_t)si, (uint64_t)q, 1);
    if ((int)(hq_ % 1000) < P.p_nohit_permille) return 0;
    if (pi != pj && (int)((hq_ >> 20) % 1000) < P.p_selfonly_permille) return 0;
    // reverse-direction query may be "self only"/"no hit" as well: nothing to do, asymmetry is intended

    int twin = ofg_synth_is_twin(P, q);
    int is_twin_slot = twin && (k == P.hq);
    // unordered pair stride: odd, shared by both directions
    int lo = pi < pj ? pi : pj, hi = pi < pj ? pj : pi;
    int64_t d = 1 + 2 * (int64_t)(ofg_hash4(P.seed, 0xd1, (uint64_t)lo, (uint64_t)hi, 2) % 48);
    int64_t off = (k == 0) ? 0 : ((k & 1) ? (int64_t)((k + 1) >> 1) : -(int64_t)(k >> 1));
    int64_t qq = (Gi == Gj) ? q : (q * Gj) / Gi;
    int64_t s;
    if (is_twin_slot) s = (qq ^ 1) % Gj;
    else s = ((qq + off * d) % Gj + Gj) % Gj;

    // symmetric per-(gene pair) hash
    uint64_t ka = ((uint64_t)si << 40) | (uint64_t)q, kb = ((uint64_t)sj << 40) | (uint64_t)s;
    uint64_t kl = ka < kb ? ka : kb, kh = ka < kb ? kb : ka;
    uint64_t hs = ofg_hash4(P.seed, 0xa7, kl, kh, 3);
    uint64_t hd = ofg_hash4(P.seed, 0xb9, ka, kb, 4);  // directional
    if (si > sj && (int)(hd % 1000) < P.p_missing_permille) return 0;

    int64_t aoff = off < 0 ? -off : off;
    int ident;  // permille
    double lq = Li[q], ls = Lj[s];
    int64_t minL = (int64_t)(lq < ls ? lq : ls);
    int same_twin_family = twin && ((q >> 1) == (s >> 1));
    if (same_twin_family) {
        ident = 700 + (int)(ofg_hash4(P.seed, 0x7718, (uint64_t)(q >> 1), 0, 6) % 200);
        minL = (int64_t)Llo[((q >> 1) << 1) % Glo];  // same value for all four lines, both directions
    } else if (P.ident_levels > 0) {
        ident = 200 + 10 * (int)(hs % (uint64_t)P.ident_levels);
        if (aoff == 0) ident += 400;
    } else if (aoff == 0) {
        ident = 600 + (int)(hs % 351);
    } else {
        int base = 520 - 12 * (int)aoff;
        if (base < 240) base = 240;
        ident = base + (int)(hs % 81) - 40;
    }
    int jit = (same_twin_family || P.integer_scores) ? 0 : (int)((hd >> 12) % 41) - 20;
    int64_t score10;  // tenths of a bit
    if (P.integer_scores) {
        score10 = 10 * ((6 * (int64_t)ident * minL + 5000) / 10000);
        if (score10 < 10) score10 = 10;
    } else {
        score10 = (19 * (int64_t)ident * minL * (1000 + jit)) / 1000000 + 200;
    }

    int nlines = 1;
    if ((int)((hs >> 24) % 1000) < P.p_dup_permille) nlines = 2;
    int n = 0;
    for (int rep = 0; rep < nlines; rep++) {
        int64_t sc = score10;
        if (rep == 1) {
            sc = (score10 * (int64_t)(300 + (hd >> 30) % 500)) / 1000;
            if (P.integer_scores) sc = (sc / 10) * 10;
            if (sc < 10) sc = 10;
        }
        uint64_t hf = ofg_mix64(hd + (uint64_t)rep);
        int64_t alen = (minL * (int64_t)(800 + hf % 200)) / 1000;
        if (alen < 10) alen = 10;
        int64_t mism = (alen * (1000 - ident)) / 1000;
        int64_t qs = 1 + (int64_t)((hf >> 10) % 20), ss = 1 + (int64_t)((hf >> 16) % 20);
        n += ofg_put_uint(out + n, (uint64_t)si); out[n++] = '_'; n += ofg_put_uint(out + n, (uint64_t)q); out[n++] = '\t';
        n += ofg_put_uint(out + n, (uint64_t)sj); out[n++] = '_'; n += ofg_put_uint(out + n, (uint64_t)s); out[n++] = '\t';
        n += ofg_put_uint(out + n, (uint64_t)(ident / 10)); out[n++] = '.'; out[n++] = (char)('0' + ident % 10); out[n++] = '\t';
        n += ofg_put_uint(out + n, (uint64_t)alen); out[n++] = '\t';
        n += ofg_put_uint(out + n, (uint64_t)mism); out[n++] = '\t';
        n += ofg_put_uint(out + n, (hf >> 22) % 10); out[n++] = '\t';
        n += ofg_put_uint(out + n, (uint64_t)qs); out[n++] = '\t';
        n += ofg_put_uint(out + n, (uint64_t)(qs + alen - 1)); out[n++] = '\t';
        n += ofg_put_uint(out + n, (uint64_t)ss); out[n++] = '\t';
        n += ofg_put_uint(out + n, (uint64_t)(ss + alen - 1)); out[n++] = '\t';
        // e-value: d.de-XX scaled with the score
        int64_t ex = sc / 30;
        if (ex > 180) ex = 180;
        out[n++] = (char)('1' + (hf >> 28) % 9); out[n++] = '.'; out[n++] = (char)('0' + (hf >> 34) % 10);
        out[n++] = 'e'; out[n++] = '-';
        if (ex < 10) out[n++] = '0';
        n += ofg_put_uint(out + n, (uint64_t)ex); out[n++] = '\t';
        n += ofg_put_uint(out + n, (uint64_t)(sc / 10));
        if (!P.integer_scores) { out[n++] = '.'; out[n++] = (char)('0' + sc % 10); }
        out[n++] = '\n';
    }
    return n;
}
