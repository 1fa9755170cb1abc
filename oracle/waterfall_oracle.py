"""CPU oracle: a numpy (+ small C helpers) restatement of OrthoFinder's WaterfallMethod hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under orthofinder_b200/ imports this module; only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it, and only
as the checker / the timed CPU baseline.

Parity status: PINNED.  tests/test_reference_pin.py (marker `reference`) runs the unmodified live reference
(oracle/ref_harness.py) in the build container on the ten usable fixtures of SURVEY.md section 8(c) and requires
this restatement to reproduce its OrthoFinder_graph.txt byte for byte, its curve_fit parameters bit for bit and its
per-gene thresholds exactly; the golden vectors committed under tests/golden/ (tools/make_golden.py, same check)
carry the live reference's outputs to where /root/reference is absent (tests/test_oracle_golden.py).

Each function cites the reference lines it follows (paths relative to /root/reference/scripts_of).
Flat CSR arrays per species pair are used instead of scipy.sparse; the arithmetic that decides
results (np.percentile's linear interpolation, np.log10, MINPACK lmdif via curve_fit, np.power,
the left-to-right diagonal products) is kept identical.
"""
import ctypes
import os
import subprocess
from collections import namedtuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

# util.py:54-61
SequencesInfo = namedtuple("SequencesInfo", "nSeqs nSpecies speciesToUse seqStartingIndices nSeqsPerSpecies")


def build_lib(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, "csrc", f) for f in os.listdir(os.path.join(_HERE, "csrc"))]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s", "liboracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build_lib())
        _LIB.ofg_oracle_lmfit.restype = ctypes.c_int
        _LIB.ofg_oracle_parse_hits.restype = ctypes.c_long
        _LIB.ofg_oracle_format_rows.restype = ctypes.c_long
    return _LIB


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


# ----------------------------------------------------------------------------------------------
# numpy.log10 as the reference sees it
# ----------------------------------------------------------------------------------------------
def numpy_uses_svml():
    try:
        from numpy._core._multiarray_umath import __cpu_features__ as f
    except Exception:  # pragma: no cover
        return False
    return bool(f.get("AVX512_SKX"))


def log10_ref(x):
    """np.log10 as numpy computes it on an AVX512_SKX host (SVML __svml_log108_ha), emulated in C
    (csrc/svml_log10.h) so that the oracle gives the same bits on every host."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.empty_like(x)
    lib().ofg_oracle_log10(_p(x), _p(y), ctypes.c_long(x.size))
    return y


# ----------------------------------------------------------------------------------------------
# A1/A2: index space and sequence lengths
# ----------------------------------------------------------------------------------------------
def get_species_to_use(species_ids_fn):
    """util.py:178-192 GetSpeciesToUse."""
    use, skipped = [], 0
    with open(species_ids_fn) as f:
        for line in f:
            line = line.rstrip()
            if not line:
                continue
            if line.startswith("#"):
                skipped += 1
            else:
                use.append(int(line.split(": ")[0]))
    return use, len(use) + skipped


def get_seqs_info(wd, species_to_use, n_sp_all):
    """util.py:64-84 GetSeqsInfo."""
    starts, n = [0], 0
    per = {}
    for k in range(n_sp_all):
        c = 0
        with open(os.path.join(wd, "Species%d.fa" % k)) as f:
            for line in f:
                if len(line) > 1 and line[0] == ">":
                    c += 1
        per[k] = c
        if k in species_to_use:
            n += c
            starts.append(n)
    return SequencesInfo(nSeqs=n, nSpecies=len(species_to_use), speciesToUse=list(species_to_use),
                         seqStartingIndices=starts[:-1], nSeqsPerSpecies=per)


def get_sequence_lengths(wd, si):
    """gathering.py:444-464 GetSequenceLengths."""
    out = []
    for k in si.speciesToUse:
        L = np.zeros(si.nSeqsPerSpecies[k])
        cur, idx, first = 0, -1, True
        with open(os.path.join(wd, "Species%d.fa" % k)) as f:
            for row in f:
                if len(row) > 1 and row[0] == ">":
                    if first:
                        first = False
                    else:
                        L[idx] = cur
                        cur = 0
                    idx = int(row[1:].split("_")[1])
                else:
                    cur += len(row.rstrip())
        L[idx] = cur
        out.append(L)
    return out


# ----------------------------------------------------------------------------------------------
# A3: one hit file -> CSR of raw bit scores
# ----------------------------------------------------------------------------------------------
class HitFileError(Exception):
    def __init__(self, kind, offset, line):
        super().__init__("malformed hit line (kind %d) at byte %d: %r" % (kind, offset, line))
        self.kind, self.offset, self.line = kind, offset, line


def parse_hits_text(text, swap_cols=False):
    """blast_file_processor.py:66-82: (q, s, score) per non-empty record, file order."""
    if isinstance(text, str):
        text = text.encode()
    cap = text.count(b"\n") + text.count(b"\r") + 1
    q = np.empty(cap, np.int32)
    s = np.empty(cap, np.int32)
    sc = np.empty(cap, np.float64)
    off = ctypes.c_long(0)
    kind = ctypes.c_int(0)
    n = lib().ofg_oracle_parse_hits(text, ctypes.c_long(len(text)), ctypes.c_int(int(swap_cols)), _p(q), _p(s),
                                    _p(sc), ctypes.c_long(cap), ctypes.byref(off), ctypes.byref(kind))
    if n < 0:
        end = off.value
        while end < len(text) and text[end:end + 1] not in (b"\n", b"\r"):
            end += 1
        raise HitFileError(kind.value, off.value, text[off.value:end])
    return q[:n], s[:n], sc[:n]


class InconsistentInput(Exception):
    pass


def get_blast6_scores(text, n_i, n_j, same_species, swap_cols=False):
    """blast_file_processor.py:38-104 GetBLAST6Scores -> (indptr, cols, vals) of the LIL matrix B.

    B[q,s] = max over lines, stored only if score > current (so scores <= 0 never stored, :87);
    self hits of a species against itself dropped (:83-84); ids past the matrix raise (:89-98).
    """
    q, s, sc = parse_hits_text(text, swap_cols)
    if same_species:
        keep = q != s
        q, s, sc = q[keep], s[keep], sc[keep]
    if q.size and (q.min() < 0 or s.min() < 0 or q.max() >= n_i or s.max() >= n_j):
        raise InconsistentInput("id out of range")
    keep = sc > 0
    q, s, sc = q[keep], s[keep], sc[keep]
    key = q.astype(np.int64) * n_j + s
    order = np.argsort(key, kind="stable")
    key, sc = key[order], sc[order]
    if key.size:
        first = np.concatenate(([True], key[1:] != key[:-1]))
        starts = np.flatnonzero(first)
        vals = np.maximum.reduceat(sc, starts)
        ukey = key[starts]
    else:
        vals, ukey = sc, key
    rows = (ukey // n_j).astype(np.int64)
    cols = (ukey % n_j).astype(np.int32)
    indptr = np.zeros(n_i + 1, np.int64)
    np.add.at(indptr, rows + 1, 1)
    return np.cumsum(indptr), cols, vals.astype(np.float64)


# ----------------------------------------------------------------------------------------------
# A5-A9: length normalisation
# ----------------------------------------------------------------------------------------------
def top_percentile_selection(Lf, S):
    """gathering.py:93-116 GetTopPercentileOfScores(L, S, 95): indices (into Lf/S) of the kept
    hits, in the order the reference emits them."""
    n = len(S)
    order = np.argsort(Lf, kind="stable")  # sorted(zip(L, range(n))) : ties by original index (:96)
    if n < 100:
        return order
    n_in = 1000 if n > 5000 else (200 if n > 1000 else 20)
    n_bins = n // n_in  # the remainder (largest Lf) is dropped (:104-107)
    idx = order[:n_bins * n_in].reshape(n_bins, n_in)
    sc = S[idx]
    # np.percentile(bin, 95), method 'linear' (numpy/lib/_function_base_impl.py: _quantile):
    # virtual index (n-1)*q, previous = floor, gamma = vi - previous, lerp a + (b-a)*gamma (gamma < .5)
    vi = (n_in - 1) * np.true_divide(95, 100)
    k = int(np.floor(vi))
    g = vi - k
    srt = np.sort(sc, axis=1)
    a, b = srt[:, k], srt[:, min(k + 1, n_in - 1)]
    cut = a + (b - a) * g
    keep = sc >= cut[:, None]
    return idx[keep]


def fit_loglinear(Lf_top, S_top):
    """gathering.py:119-121 CalculateFittingParameters == curve_fit(a*log10(x)+b, Lf, log10(S))."""
    lx = log10_ref(Lf_top)
    ly = log10_ref(S_top)
    out = np.zeros(2)
    stats = np.zeros(3, np.int32)
    info = lib().ofg_oracle_lmfit(_p(lx), _p(ly), ctypes.c_long(lx.size), _p(out), _p(stats))
    if info not in (1, 2, 3, 4):
        raise RuntimeError("Optimal parameters not found (MINPACK info %d)" % info)
    return out, info, stats


def normalise_scores(csr, Li, Lj):
    """gathering.py:140-155 NormaliseScores.  Returns (csr', fit) ; fit is None when the pair is zeroed."""
    indptr, cols, vals = csr
    n_i = len(indptr) - 1
    rows = np.repeat(np.arange(n_i), np.diff(indptr))
    Lf = Li[rows] * Lj[cols]  # :85-90,142
    sel = top_percentile_selection(Lf, vals)
    if len(sel) > 1:
        (a, b), info, stats = fit_loglinear(Lf[sel], vals[sel])
        a = np.float64(a)
        b = np.float64(b)
        # :124-131  10**(-b) * diag(Lq**-a) * B * diag(Lh**-a), evaluated left to right
        li = Li ** (-a)
        lj = Lj ** (-a)
        new = ((10 ** (-b) * li)[rows] * vals) * lj[cols]
        return (indptr, cols, new), (float(a), float(b), len(sel), info, int(stats[0]))
    empty = (np.zeros(n_i + 1, np.int64), np.zeros(0, np.int32), np.zeros(0))
    return empty, None


def normalised_bit_score(csr, Li, Lj):
    """gathering.py:157-174 NormalisedBitScore (A18, `--scores-v2`): diag(Lq**-0.5) * B * diag(Lh**-0.5), the two
    sparse products evaluated left to right; no fit, no pair is ever zeroed.  Returns csr'."""
    indptr, cols, vals = csr
    if vals.size == 0:  # :165-166
        return csr
    n_i = len(indptr) - 1
    rows = np.repeat(np.arange(n_i), np.diff(indptr))
    li = Li ** (-0.5)  # :171-172
    lj = Lj ** (-0.5)
    return indptr, cols, (li[rows] * vals) * lj[cols]


# ----------------------------------------------------------------------------------------------
# A11: best hits with tolerance
# ----------------------------------------------------------------------------------------------
def _row_max(indptr, vals, fill):
    n = len(indptr) - 1
    out = np.full(n, fill, np.float64)
    nz = np.diff(indptr) > 0
    if vals.size:
        out[nz] = np.maximum.reduceat(vals, indptr[:-1][nz])
    return out, nz


def get_bh(B_row, p, tol=1e-3):
    """gathering.py:213-248 GetBH_s: boolean flag per stored entry of B[p][j]."""
    n_i = len(B_row[0][0]) - 1
    best = -np.ones(n_i)
    flags = [None] * len(B_row)
    for j, (indptr, cols, vals) in enumerate(B_row):
        if j == p:
            continue
        m, nz = _row_max(indptr, vals, -1.0)
        best = np.where(nz & (m > best), m, best)
        rows = np.repeat(np.arange(n_i), np.diff(indptr))
        flags[j] = vals > (m[rows] - tol)
    indptr, cols, vals = B_row[p]
    rows = np.repeat(np.arange(n_i), np.diff(indptr))
    flags[p] = vals > (best[rows] - tol)
    return flags


# ----------------------------------------------------------------------------------------------
# A12-A14: reciprocal best hits, per-gene threshold, connect
# ----------------------------------------------------------------------------------------------
def _entry_keys(indptr, cols, n_cols):
    rows = np.repeat(np.arange(len(indptr) - 1, dtype=np.int64), np.diff(indptr))
    return rows, rows * n_cols + cols


def _rbh_flags(B, BH, p, k, n_seqs):
    """RBH_pk = BH_pk AND BH_kp^T (matrices.py:65-69) as a flag per stored entry of B[p][k]; also the entry rows."""
    indptr, cols, vals = B[p][k]
    rows, keys = _entry_keys(indptr, cols, n_seqs[k])
    f = BH[p][k]
    indptr2, cols2, _ = B[k][p]
    rows2, _ = _entry_keys(indptr2, cols2, n_seqs[p])
    f2 = BH[k][p]
    tkeys = cols2[f2].astype(np.int64) * n_seqs[k] + rows2[f2]  # (row of p, col of k)
    return f & np.isin(keys, tkeys), rows


def rbh_conversion_factors(B, BH, p, n_seqs, pct=90):
    """gathering.py:333-364, stage 1 of GetMostDistant_s_estimate_from_relative_rbhs: for query species p, the factor
    per other species k (in species order) that converts an RBH score in k into the expected score of the most
    distant RBH.  Z[k][i] is gene i's RBH score in species k if it has exactly ONE RBH there, else 0 (:340);
    rs[a][b] is the (100 - pct)th percentile of Z[a] * (1 / Z[b]) over the genes where both exist (:344-357);
    the factor of species b is the column minimum (:364)."""
    others = [k for k in range(len(B[p])) if k != p]
    n_i = n_seqs[p]
    Z = np.zeros((len(others), n_i))
    for a, k in enumerate(others):
        rbh, rows = _rbh_flags(B, BH, p, k, n_seqs)
        cnt = np.bincount(rows[rbh], minlength=n_i)
        vals = B[p][k][2]
        single = rbh & (cnt[rows] == 1)
        Z[a, rows[single]] = vals[single]
    with np.errstate(divide="ignore", invalid="ignore"):
        Zr = 1.0 / Z  # :342
        rs = np.ones((len(others), len(others)))
        for a in range(len(others)):
            for b in range(len(others)):
                if a == b:
                    continue
                ratios = Z[a] * Zr[b]  # :349 (reciprocal first, then the product)
                ok = np.isfinite(ratios) & (ratios > 0)
                if ok.any():
                    rs[a, b] = np.percentile(ratios[ok], 100 - pct)  # :355
    return (rs.min(axis=0) if len(others) else np.zeros(0)), others


def connect_cognates(B, BH, p, n_seqs, v2_scores=False):
    """gathering.py:251-259 ConnectCognates for query species p.
    B[p][j], BH[p][j] and BH[j][p] for all j are read.  Returns (connect flags per j, mostDistant).
    v2_scores: thresholds by GetMostDistant_s_estimate_from_relative_rbhs (:314-388) instead of GetMostDistant_s."""
    S = len(B[p])
    n_i = n_seqs[p]
    most = np.ones(n_i) * 1e9  # :296 / :366
    best = np.zeros(n_i)  # :298 / :368
    factor = {}
    if v2_scores:
        C, others = rbh_conversion_factors(B, BH, p, n_seqs)
        factor = {k: C[a] for a, k in enumerate(others)}
    for k in range(S):
        if k == p:
            continue
        indptr, cols, vals = B[p][k]
        m, _ = _row_max(indptr, vals, 0.0)  # matrices.py:75-78 sparse_max_row
        best = np.maximum(best, m)
        rbh, rows = _rbh_flags(B, BH, p, k, n_seqs)
        I = rows[rbh]
        if I.size:
            # :304-306 / :381-384  mostDistant[I] = np.minimum([factor *] B[I,J], mostDistant[I]) : fancy assignment
            # with repeated I is last-wins; entries are in (row, col) order so the last is the largest col
            v = np.minimum(factor[k] * vals[rbh] if v2_scores else vals[rbh], most[I])
            most[I] = v
    I = most > 1e8
    most[I] = best[I] + 1e-6  # :308-309
    conn = []
    for j in range(S):
        indptr, cols, vals = B[p][j]
        rows = np.repeat(np.arange(n_i), np.diff(indptr))
        c = vals >= most[rows]  # :391-408
        if j == p:
            c &= rows != cols
        conn.append(c)
    return conn, most


# ----------------------------------------------------------------------------------------------
# A15/A16: graph assembly and text
# ----------------------------------------------------------------------------------------------
def assemble_graph(B, conn, si_starts, n_seqs):
    """gathering.py:23-49: W_pj = (connect_pj + connect_jp^T) .* B_pj ; returns the global CSR."""
    S = len(B)
    N = int(sum(n_seqs))
    R, C, V = [], [], []
    for p in range(S):
        for j in range(S):
            indptr, cols, vals = B[p][j]
            rows, keys = _entry_keys(indptr, cols, n_seqs[j])
            c1 = conn[p][j].astype(np.float64)
            indptr2, cols2, _ = B[j][p]
            rows2, _ = _entry_keys(indptr2, cols2, n_seqs[p])
            f2 = conn[j][p]
            tkeys = cols2[f2].astype(np.int64) * n_seqs[j] + rows2[f2]
            c2 = np.isin(keys, tkeys).astype(np.float64)
            cnt = c1 + c2
            w = cnt * vals
            keep = cnt != 0  # x.multiply(y) keeps the structural intersection; zero-weight never written
            R.append(rows[keep] + si_starts[p])
            C.append(cols[keep].astype(np.int64) + si_starts[j])
            V.append(w[keep])
    R = np.concatenate(R) if R else np.zeros(0, np.int64)
    C = np.concatenate(C) if C else np.zeros(0, np.int64)
    V = np.concatenate(V) if V else np.zeros(0)
    order = np.lexsort((C, R))
    R, C, V = R[order], C[order], V[order]
    indptr = np.zeros(N + 1, np.int64)
    np.add.at(indptr, R + 1, 1)
    return np.cumsum(indptr), C.astype(np.int32), V


def assemble_homology_graph(B, si_starts, n_seqs):
    """gathering.py:51-73 WriteGraph_perSpecies_homology (`--c-homologs`, gathering_version (3, 2)): per query species the
    structural union W_pj = (B_pj + B_jp^T > 0) of the normalised matrices, every entry written with weight 1.0."""
    S = len(B)
    N = int(sum(n_seqs))
    R, C = [], []
    for p in range(S):
        for j in range(S):
            indptr, cols, vals = B[p][j]
            rows, keys = _entry_keys(indptr, cols, n_seqs[j])
            indptr2, cols2, vals2 = B[j][p]
            rows2, _ = _entry_keys(indptr2, cols2, n_seqs[p])
            tkeys = cols2.astype(np.int64) * n_seqs[j] + rows2      # entry (x, y) of B_jp is entry (y, x) of its transpose
            u = np.union1d(keys[vals > 0], tkeys[vals2 > 0])
            R.append(u // n_seqs[j] + si_starts[p])
            C.append(u % n_seqs[j] + si_starts[j])
    R = np.concatenate(R) if R else np.zeros(0, np.int64)
    C = np.concatenate(C) if C else np.zeros(0, np.int64)
    order = np.lexsort((C, R))
    R, C = R[order], C[order]
    indptr = np.zeros(N + 1, np.int64)
    np.add.at(indptr, R + 1, 1)
    return np.cumsum(indptr), C.astype(np.int32), np.ones(C.size)


def graph_text(indptr, indices, data, n_total):
    """gathering.py:273-291 header + :39-48 rows + ')' after the last species."""
    head = ("(mclheader\nmcltype matrix\ndimensions %dx%d\n)\n" % (n_total, n_total)).encode()
    head += b"\n(mclmatrix\nbegin\n\n"
    nnz = int(indptr[-1])
    cap = 64 * (n_total + 1) + 40 * nnz + 64
    buf = ctypes.create_string_buffer(cap)
    indptr = np.ascontiguousarray(indptr, np.int64)
    indices = np.ascontiguousarray(indices, np.int32)
    data = np.ascontiguousarray(data, np.float64)
    w = lib().ofg_oracle_format_rows(_p(indptr), _p(indices), _p(data), ctypes.c_long(0), ctypes.c_long(n_total),
                                     buf, ctypes.c_long(cap))
    assert w >= 0
    return head + buf.raw[:w] + b")\n"


# ----------------------------------------------------------------------------------------------
# whole path
# ----------------------------------------------------------------------------------------------
def waterfall(n_seqs, lengths, hit_text, q_double_blast=True, return_parts=False, v2_scores=False, homology=False):
    """WaterfallMethod end to end (gathering.py:476-512 of DoOrthogroups; v2_scores = `--scores-v2`, A18).

    n_seqs[p], lengths[p] per used species (list order); hit_text(p, j) -> bytes of Blast{p}_{j}
    (ids already translated to list positions by the caller) or None if absent.
    Returns dict(indptr, indices, data, fits, B, BH, connect, most_distant, warnings).
    """
    S = len(n_seqs)
    starts = np.concatenate(([0], np.cumsum(n_seqs)[:-1])).astype(np.int64)
    B = [[None] * S for _ in range(S)]
    fits = {}
    warns = []
    for p in range(S):
        for j in range(S):
            swap = (not q_double_blast) and p > j  # blast_file_processor.py:41-54
            txt = hit_text(j, p) if swap else hit_text(p, j)
            raw = get_blast6_scores(txt if txt is not None else b"", n_seqs[p], n_seqs[j], p == j, swap)
            if v2_scores:
                B[p][j], fits[(p, j)] = normalised_bit_score(raw, lengths[p], lengths[j]), None
                continue
            B[p][j], fits[(p, j)] = normalise_scores(raw, lengths[p], lengths[j])
            if fits[(p, j)] is None:
                warns.append((p, j, int(raw[0][-1])))
    if homology:   # gathering.py:520-521: no best hits, no thresholds - the symmetrised structure of B
        indptr, indices, data = assemble_homology_graph(B, starts, n_seqs)
        return dict(indptr=indptr, indices=indices, data=data, fits=fits, warnings=warns, most_distant=None)
    BH = [get_bh(B[p], p) for p in range(S)]
    conn, most = [], []
    for p in range(S):
        c, m = connect_cognates(B, BH, p, n_seqs, v2_scores)
        conn.append(c)
        most.append(m)
    indptr, indices, data = assemble_graph(B, conn, starts, n_seqs)
    out = dict(indptr=indptr, indices=indices, data=data, fits=fits, warnings=warns, most_distant=most)
    if return_parts:
        out.update(B=B, BH=BH, connect=conn)
    return out


def waterfall_from_directory(wd, q_double_blast=True, return_parts=False, v2_scores=False, homology=False):
    """Entry used by tests: reads SpeciesIDs.txt / Species*.fa / Blast*_*.txt like the reference."""
    import gzip
    use, n_all = get_species_to_use(os.path.join(wd, "SpeciesIDs.txt"))
    si = get_seqs_info(wd, use, n_all)
    L = get_sequence_lengths(wd, si)
    n_seqs = [si.nSeqsPerSpecies[k] for k in si.speciesToUse]

    def hit_text(p, j):
        fn = os.path.join(wd, "Blast%d_%d.txt" % (si.speciesToUse[p], si.speciesToUse[j]))
        if os.path.exists(fn + ".gz"):  # blast_file_processor.py:65
            with gzip.open(fn + ".gz", "rb") as f:
                return f.read()
        with open(fn, "rb") as f:
            return f.read()

    res = waterfall(n_seqs, L, hit_text, q_double_blast, return_parts, v2_scores, homology)
    res["seqsInfo"] = si
    res["lengths"] = L
    res["graph"] = graph_text(res["indptr"], res["indices"], res["data"], si.nSeqs)
    return res
