/* Sequential restatement of blast_file_processor.GetBLAST6Scores' line handling
 * (blast_file_processor.py:66-88) and of WriteGraph_perSpecies' text emission
 * (gathering.py:39-48).  TEST INFRASTRUCTURE (oracle).
 *
 * Text-mode csv.reader semantics that matter for BLAST-6 files: '\n', '\r\n' and '\r' end a
 * record, empty records are skipped (:68-69), fields split on '\t', fields 0/1 give
 * int(field.split("_", 2)[1]), field 11 gives float(field).  Python's int()/float() accept
 * surrounding ASCII whitespace; strtol/strtod are used with an explicit "whole field consumed"
 * check (strtod is correctly rounded in glibc, like Python's float()).  A '"' anywhere in a
 * line is reported as unsupported (csv quoting is never produced by BLAST/DIAMOND/MMseqs2).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <ctype.h>
#include <errno.h>

/* error kinds (shared numbering with include/ofg.h) */
enum { OFG_LINE_OK = 0, OFG_LINE_BAD_ID = 1, OFG_LINE_BAD_SCORE = 2, OFG_LINE_QUOTE = 3 };

static int parse_id_field(const char *f, const char *fe, long *out)
{
    /* field.split("_", 2)[1] : text after the first '_' up to the second '_' or the field end */
    const char *u = memchr(f, '_', (size_t)(fe - f));
    if (!u) return -1; /* IndexError */
    const char *b = u + 1;
    const char *e = memchr(b, '_', (size_t)(fe - b));
    if (!e) e = fe;
    while (b < e && isspace((unsigned char)*b)) b++;
    while (e > b && isspace((unsigned char)e[-1])) e--;
    if (b == e) return -1;
    const char *p = b;
    int neg = 0;
    if (*p == '+' || *p == '-') { neg = (*p == '-'); p++; }
    if (p == e) return -1;
    long v = 0;
    for (; p < e; p++) {
        if (*p < '0' || *p > '9') return -1;
        v = v * 10 + (*p - '0');
        if (v > 2000000000L) return -1;
    }
    *out = neg ? -v : v;
    return 0;
}

static int parse_score_field(const char *f, const char *fe, double *out)
{
    char buf[128];
    size_t n = (size_t)(fe - f);
    if (n == 0 || n >= sizeof buf) return -1;
    memcpy(buf, f, n);
    buf[n] = 0;
    char *b = buf, *e = buf + n;
    while (*b && isspace((unsigned char)*b)) b++;
    while (e > b && isspace((unsigned char)e[-1])) *--e = 0;
    if (b == e) return -1;
    for (char *p = b; p < e; p++) /* reject what strtod accepts but float() does not (hex), keep it simple */
        if (*p == 'x' || *p == 'X' || *p == '_') return -1;
    char *endp;
    errno = 0;
    double v = strtod(b, &endp);
    if (endp != e) return -1;
    *out = v;
    return 0;
}

/* Parses `text` (nbytes).  For every non-empty record appends (q, s, score) in file order.
 * Returns the number of records, or -(k) - 1 where the k-th (0-based) byte offset in
 * *err_off starts the first offending line and *err_kind says why. */
long ofg_oracle_parse_hits(const char *text, long nbytes, int swap_cols, int32_t *q_out,
                           int32_t *s_out, double *score_out, long cap, long *err_off, int *err_kind)
{
    long n = 0, pos = 0;
    while (pos < nbytes) {
        long ls = pos;
        while (pos < nbytes && text[pos] != '\n' && text[pos] != '\r') pos++;
        long le = pos;
        if (pos < nbytes) pos++;
        if (le == ls) continue;
        const char *tabs[13];
        int nt = 0;
        int quote = 0;
        for (long k = ls; k < le; k++) {
            if (text[k] == '\t' && nt < 12) tabs[nt++] = text + k;
            if (text[k] == '"') quote = 1;
        }
        if (quote) { *err_off = ls; *err_kind = OFG_LINE_QUOTE; return -1; }
        if (nt < 11) { /* fewer than 12 fields: ids may still parse, the score raises IndexError */
            *err_off = ls;
            *err_kind = nt < 1 ? OFG_LINE_BAD_ID : OFG_LINE_BAD_SCORE;
            long t;
            if (nt >= 1 && (parse_id_field(text + ls, tabs[0], &t) ||
                            parse_id_field(tabs[0] + 1, nt >= 2 ? tabs[1] : text + le, &t)))
                *err_kind = OFG_LINE_BAD_ID;
            return -1;
        }
        long a, b;
        if (parse_id_field(text + ls, tabs[0], &a) || parse_id_field(tabs[0] + 1, tabs[1], &b)) {
            *err_off = ls; *err_kind = OFG_LINE_BAD_ID; return -1;
        }
        double sc;
        const char *fe = nt >= 12 ? tabs[11] : text + le;
        if (parse_score_field(tabs[10] + 1, fe, &sc)) { *err_off = ls; *err_kind = OFG_LINE_BAD_SCORE; return -1; }
        if (n >= cap) { *err_off = ls; *err_kind = -1; return -1; }
        q_out[n] = (int32_t)(swap_cols ? b : a);
        s_out[n] = (int32_t)(swap_cols ? a : b);
        score_out[n] = sc;
        n++;
    }
    return n;
}

/* gathering.py:39-48 : "%d    " row id, then "%d:%.3f " per entry, then "$\n". */
long ofg_oracle_format_rows(const int64_t *indptr, const int32_t *indices, const double *data,
                            long row0, long row1, char *out, long cap)
{
    long w = 0;
    for (long r = row0; r < row1; r++) {
        if (cap - w < 64) return -1;
        w += sprintf(out + w, "%ld    ", r);
        for (int64_t k = indptr[r]; k < indptr[r + 1]; k++) {
            if (cap - w < 64) return -1;
            w += sprintf(out + w, "%d:%.3f ", indices[k], data[k]);
        }
        w += sprintf(out + w, "$\n");
    }
    return w;
}
