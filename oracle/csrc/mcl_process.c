/* CPU restatement of the MCL process that the reference hands the graph to (TEST INFRASTRUCTURE; see the header of
 * oracle/mcl_oracle.py).
 *
 * The reference runs a bundled binary without source: `mcl <graph> -I <inflation> -o <clusters> -te <n> -V all`
 * (scripts_of/mcl.py:214-218, binary scripts_of/bin/mcl = mcl 14-137, default resource scheme 6: P = 10000, S = 1100,
 * R = 1400, pct = 90).  Everything below restates van Dongen's published algorithm (expansion, inflation, pruning,
 * interpretation of the limit as attractor systems); the arithmetic details that the publication leaves open were pinned
 * against the binary's own dumps (`-write-graphx`, `-write-expanded`, `-dump ite`, `-dump dag`, `--write-binary`):
 *   start matrix   entries that read as 0 and existing loops dropped, loop := max of the column (1 for an empty column), column / its double sum;
 *   expansion      column c of M*M; per result entry the products are accumulated in double in the order of the source
 *                  column; the product itself is float*float rounded to float when the column has
 *                  summands * overhead >= n ("dense" walk, overhead 10 = the binary's -sparse default), double otherwise;
 *   pruning        keep val >= 0.99999/P; if fewer than R entries and less than pct of the mass are left: the R largest of the
 *                  whole vector; else if more than S entries are left: the S largest of them, and if that leaves fewer than
 *                  R entries and less than pct of the mass: the R largest of the pruned vector.  "The K largest" keeps
 *                  every entry equal to the K-th largest.  Then val / (double sum of what is left);
 *   chaos          max over columns of (max - sum of squares) * entries, on what pruning left of the column and before
 *                  that is rescaled (fits the binary's log and the iteration counts of all 21 golden runs); stop when < 1e-4;
 *   inflation      val := (float) pow((double) val, (double)(float) I), then / the double sum;
 *   dag            per column keep val >= bar, bar = 0.999 self + 0.001 max if self < max, else self / 1.01;
 *   clusters       nodes with a loop in the dag are attractors, connected attractors form a system, every node joins
 *                  the systems it reaches; a node reaching several stays in the first (systems in order of their smallest
 *                  attractor); nodes that reach no system form one cluster together; output order: size descending, then
 *                  smallest member.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int n;
    int64_t *cp;   /* n + 1 */
    int32_t *ri;
    float *v;
    int64_t cap;
} mclo_mx;

typedef struct {
    int n;
    mclo_mx cur, nxt, exp;   /* current iterand, scratch, pruned expansion of the last step (before inflation) */
    double *acc;
    int32_t *touched;
    unsigned char *seen;
    float *tmp;
} mclo;

static void mx_init(mclo_mx *m, int n, int64_t cap)
{
    m->n = n;
    m->cp = (int64_t *)calloc((size_t)n + 1, sizeof(int64_t));
    m->cap = cap > 16 ? cap : 16;
    m->ri = (int32_t *)malloc(sizeof(int32_t) * (size_t)m->cap);
    m->v = (float *)malloc(sizeof(float) * (size_t)m->cap);
}
static void mx_reserve(mclo_mx *m, int64_t need)
{
    if (need <= m->cap) return;
    while (m->cap < need) m->cap *= 2;
    m->ri = (int32_t *)realloc(m->ri, sizeof(int32_t) * (size_t)m->cap);
    m->v = (float *)realloc(m->v, sizeof(float) * (size_t)m->cap);
}
static void mx_free(mclo_mx *m) { free(m->cp); free(m->ri); free(m->v); }

static int cmp_i32(const void *a, const void *b)
{
    int32_t x = *(const int32_t *)a, y = *(const int32_t *)b;
    return (x > y) - (x < y);
}
static int cmp_f32_desc(const void *a, const void *b)
{
    float x = *(const float *)a, y = *(const float *)b;
    return (x < y) - (x > y);
}

/* start matrix from the graph as mcl reads it: column c = line c of the graph file */
void *mclo_new(int n, const int64_t *cp, const int32_t *ri, const float *v)
{
    mclo *h = (mclo *)calloc(1, sizeof(mclo));
    h->n = n;
    int64_t nnz = cp[n];
    mx_init(&h->cur, n, nnz + n);
    mx_init(&h->nxt, n, nnz + n);
    mx_init(&h->exp, n, nnz + n);
    h->acc = (double *)malloc(sizeof(double) * (size_t)n);
    h->touched = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
    h->seen = (unsigned char *)calloc((size_t)n, 1);
    h->tmp = (float *)malloc(sizeof(float) * (size_t)(n > 16 ? n : 16));
    int64_t o = 0;
    for (int c = 0; c < n; c++) {
        float mx = 0.0f;
        int any = 0;
        for (int64_t p = cp[c]; p < cp[c + 1]; p++)
            if (ri[p] != c && v[p] != 0.0f) { if (!any || v[p] > mx) mx = v[p]; any = 1; }
        float loop = any ? mx : 1.0f;
        int64_t b = o;
        int placed = 0;
        for (int64_t p = cp[c]; p < cp[c + 1]; p++) {
            if (ri[p] == c || v[p] == 0.0f) continue;   /* loops are replaced, zero entries are dropped on read */
            if (!placed && ri[p] > c) { h->cur.ri[o] = c; h->cur.v[o++] = loop; placed = 1; }
            h->cur.ri[o] = ri[p]; h->cur.v[o++] = v[p];
        }
        if (!placed) { h->cur.ri[o] = c; h->cur.v[o++] = loop; }
        double s = 0.0;
        for (int64_t p = b; p < o; p++) s += (double)h->cur.v[p];
        for (int64_t p = b; p < o; p++) h->cur.v[p] = (float)((double)h->cur.v[p] / s);
        h->cur.cp[c + 1] = o;
    }
    return h;
}

void mclo_free(void *hh)
{
    mclo *h = (mclo *)hh;
    mx_free(&h->cur); mx_free(&h->nxt); mx_free(&h->exp);
    free(h->acc); free(h->touched); free(h->seen); free(h->tmp); free(h);
}

static float kth_largest(const float *v, int n, int k, float *tmp)
{
    memcpy(tmp, v, sizeof(float) * (size_t)n);
    qsort(tmp, (size_t)n, sizeof(float), cmp_f32_desc);
    return tmp[k - 1];
}

/* the pruning rule; vals[0..n) in row order, keep[] out; returns the number kept */
static int prune_vector(const float *v, int n, double cut, int S, int R, double pct, unsigned char *keep, float *tmp, float *tmp2)
{
    double mass_all = 0.0, mass_p = 0.0;
    int n_p = 0;
    for (int i = 0; i < n; i++) mass_all += (double)v[i];
    for (int i = 0; i < n; i++) {
        keep[i] = (double)v[i] >= cut;
        if (keep[i]) { n_p++; mass_p += (double)v[i]; }
    }
    if (R && n_p < R && mass_p < pct * mass_all) {          /* recovery from the whole vector */
        if (n <= R) { memset(keep, 1, (size_t)n); return n; }
        float bar = kth_largest(v, n, R, tmp);
        int k = 0;
        for (int i = 0; i < n; i++) { keep[i] = v[i] >= bar; k += keep[i]; }
        return k;
    }
    if (S && n_p > S) {                                     /* selection among the pruned entries */
        int m = 0;
        for (int i = 0; i < n; i++) if (keep[i]) tmp2[m++] = v[i];
        float bar = kth_largest(tmp2, m, S, tmp);
        int k = 0;
        double mass_s = 0.0;
        for (int i = 0; i < n; i++) if (keep[i] && v[i] >= bar) { k++; mass_s += (double)v[i]; }
        if (R && k < R && mass_s < pct * mass_all) {        /* recovery from the pruned vector */
            if (m <= R) return n_p;
            bar = kth_largest(tmp2, m, R, tmp);
        }
        k = 0;
        for (int i = 0; i < n; i++) { keep[i] = keep[i] && v[i] >= bar; k += keep[i]; }
        return k;
    }
    return n_p;
}

/* one expansion + inflation; returns the chaos of the expansion */
double mclo_step(void *hh, double inflation, int P, int S, int R, double pct, int overhead)
{
    mclo *h = (mclo *)hh;
    const mclo_mx *M = &h->cur;
    mclo_mx *E = &h->exp, *N = &h->nxt;
    int n = h->n;
    double cut = P ? 0.99999 / (double)P : 0.0, chaos = 0.0;
    double power = (double)(float)inflation;
    float *vals = (float *)malloc(sizeof(float) * (size_t)n * 3);
    float *tmp = vals + n, *tmp2 = vals + 2 * (size_t)n;
    unsigned char *keep = (unsigned char *)malloc((size_t)n);
    int64_t o = 0;
    for (int c = 0; c < n; c++) {
        int64_t ns = 0;
        for (int64_t p = M->cp[c]; p < M->cp[c + 1]; p++) ns += M->cp[M->ri[p] + 1] - M->cp[M->ri[p]];
        int dense = !(ns * (int64_t)overhead < (int64_t)n);
        int nt = 0;
        for (int64_t p = M->cp[c]; p < M->cp[c + 1]; p++) {
            int k = M->ri[p];
            float f = M->v[p];
            for (int64_t q = M->cp[k]; q < M->cp[k + 1]; q++) {
                int r = M->ri[q];
                if (!h->seen[r]) { h->seen[r] = 1; h->touched[nt++] = r; h->acc[r] = 0.0; }
                if (dense) { float pr = f * M->v[q]; h->acc[r] += (double)pr; }
                else h->acc[r] += (double)f * (double)M->v[q];
            }
        }
        qsort(h->touched, (size_t)nt, sizeof(int32_t), cmp_i32);
        for (int i = 0; i < nt; i++) { vals[i] = (float)h->acc[h->touched[i]]; h->seen[h->touched[i]] = 0; }
        int kept = prune_vector(vals, nt, cut, S, R, pct, keep, tmp, tmp2);
        mx_reserve(E, o + kept); mx_reserve(N, o + kept);
        double s = 0.0;
        for (int i = 0; i < nt; i++) if (keep[i]) s += (double)vals[i];
        int64_t b = o;
        {   /* chaos: on what pruning left, before it is rescaled */
            double mx0 = 0.0, ssq0 = 0.0;
            for (int i = 0; i < nt; i++)
                if (keep[i]) { if ((double)vals[i] > mx0) mx0 = (double)vals[i]; ssq0 += (double)vals[i] * (double)vals[i]; }
            double ch0 = (mx0 - ssq0) * (double)kept;
            if (ch0 > chaos) chaos = ch0;
        }
        for (int i = 0; i < nt; i++) {
            if (!keep[i]) continue;
            float w = (float)((double)vals[i] / s);
            E->ri[o] = h->touched[i]; E->v[o] = w; N->ri[o] = h->touched[i];
            o++;
        }
        double ps = 0.0;
        for (int64_t p = b; p < o; p++) { N->v[p] = (float)pow((double)E->v[p], power); ps += (double)N->v[p]; }
        for (int64_t p = b; p < o; p++) N->v[p] = (float)((double)N->v[p] / ps);
        E->cp[c + 1] = o; N->cp[c + 1] = o;
    }
    free(vals); free(keep);
    mclo_mx t = h->cur; h->cur = h->nxt; h->nxt = t;
    return chaos;
}

/* which: 0 current iterand, 1 pruned expansion of the last step */
int64_t mclo_nnz(void *hh, int which) { mclo *h = (mclo *)hh; return (which ? h->exp : h->cur).cp[h->n]; }
void mclo_get(void *hh, int which, int64_t *cp, int32_t *ri, float *v)
{
    mclo *h = (mclo *)hh;
    const mclo_mx *m = which ? &h->exp : &h->cur;
    memcpy(cp, m->cp, sizeof(int64_t) * ((size_t)h->n + 1));
    memcpy(ri, m->ri, sizeof(int32_t) * (size_t)m->cp[h->n]);
    memcpy(v, m->v, sizeof(float) * (size_t)m->cp[h->n]);
}

static int cmp_key_pair(const void *a, const void *b)
{
    int64_t x = *(const int64_t *)a, y = *(const int64_t *)b;
    return (x > y) - (x < y);
}

static int uf_find(int32_t *par, int x)
{
    while (par[x] != x) { par[x] = par[par[x]]; x = par[x]; }
    return x;
}

/* dag + attractor systems of matrix `which`; label[node] = index of its cluster in OUTPUT order; returns the number of
 * clusters; *overlap = nodes that reached more than one system */
int mclo_interpret(void *hh, int which, int32_t *label, int32_t *overlap)
{
    mclo *h = (mclo *)hh;
    const mclo_mx *M = which ? &h->exp : &h->cur;
    int n = h->n;
    /* dag (as a keep mask over M's entries) */
    unsigned char *keep = (unsigned char *)malloc((size_t)M->cp[n] + 1);
    unsigned char *attr = (unsigned char *)calloc((size_t)n, 1);
    for (int c = 0; c < n; c++) {
        double self = 0.0, mx = 0.0;
        for (int64_t p = M->cp[c]; p < M->cp[c + 1]; p++) {
            if (M->ri[p] == c) self = (double)M->v[p];
            if ((double)M->v[p] > mx) mx = (double)M->v[p];
        }
        double bar = self < mx ? 0.999 * self + 0.001 * mx : self / 1.01;
        for (int64_t p = M->cp[c]; p < M->cp[c + 1]; p++) {
            keep[p] = (double)M->v[p] >= bar;
            if (keep[p] && M->ri[p] == c) attr[c] = 1;
        }
    }
    int32_t *par = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
    for (int i = 0; i < n; i++) par[i] = i;
    for (int c = 0; c < n; c++) {
        if (!attr[c]) continue;
        for (int64_t p = M->cp[c]; p < M->cp[c + 1]; p++)
            if (keep[p] && attr[M->ri[p]]) {
                int a = uf_find(par, c), b = uf_find(par, M->ri[p]);
                if (a != b) { if (a < b) par[b] = a; else par[a] = b; }   /* root = smallest attractor of the system */
            }
    }
    /* sys[node] = smallest-attractor root of the FIRST system the node reaches (systems ordered by that root) */
    int32_t *sys = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
    int32_t *second = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);   /* another system it reaches, or -1 */
    for (int i = 0; i < n; i++) { sys[i] = attr[i] ? uf_find(par, i) : -1; second[i] = -1; }
    /* nodes reach systems through dag arcs, possibly through other non-attractor nodes: propagate to a fixed point */
    int changed = 1, guard = 0;
    while (changed && guard++ < n + 2) {
        changed = 0;
        for (int c = 0; c < n; c++) {
            if (attr[c]) continue;
            for (int64_t p = M->cp[c]; p < M->cp[c + 1]; p++) {
                if (!keep[p]) continue;
                int r = M->ri[p];
                int cand[2] = { sys[r], second[r] };
                for (int t = 0; t < 2; t++) {
                    int s = cand[t];
                    if (s < 0 || s == sys[c] || s == second[c]) continue;
                    if (sys[c] < 0) sys[c] = s;
                    else if (s < sys[c]) { second[c] = sys[c]; sys[c] = s; }
                    else if (second[c] < 0 || s < second[c]) second[c] = s;
                    else continue;
                    changed = 1;
                }
            }
        }
    }
    int nov = 0, garbage = -1;
    for (int i = 0; i < n; i++) {
        if (second[i] >= 0) nov++;
        if (sys[i] < 0) {                    /* reaches no attractor system: the binary collects these nodes in ONE cluster */
            if (garbage < 0) garbage = i;    /* ("added <k> garbage entries"); its id = its smallest member, never an attractor */
            sys[i] = garbage;
        }
    }
    *overlap = nov;
    /* output order: size descending, then smallest member */
    int32_t *size = (int32_t *)calloc((size_t)n, sizeof(int32_t));
    int32_t *first = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
    for (int i = 0; i < n; i++) first[i] = -1;
    for (int i = 0; i < n; i++) { size[sys[i]]++; if (first[sys[i]] < 0) first[sys[i]] = i; }
    int nc = 0;
    int64_t *key = (int64_t *)malloc(sizeof(int64_t) * (size_t)n * 2);
    for (int s = 0; s < n; s++)
        if (size[s]) { key[2 * nc] = ((int64_t)(n - size[s]) << 32) | (int64_t)first[s]; key[2 * nc + 1] = s; nc++; }
    qsort(key, (size_t)nc, 2 * sizeof(int64_t), cmp_key_pair);
    int32_t *rank = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
    for (int i = 0; i < nc; i++) rank[key[2 * i + 1]] = i;
    for (int i = 0; i < n; i++) label[i] = rank[sys[i]];
    free(keep); free(attr); free(par); free(sys); free(second); free(size); free(first); free(key); free(rank);
    return nc;
}
