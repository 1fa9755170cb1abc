"""CPU restatement of the DendroBLAST distance matrices - TEST INFRASTRUCTURE ONLY.

Nothing under orthofinder_b200/ may import this module; it is the checker for tests/ (and pinned against the live
reference by tests/test_reference_pin.py, goldens written by tools/make_golden_dendro.py).

Follows /root/reference/scripts_of/orthologues.py:
  :226-236  lil_minmax                                       -> row_min_max
  :264-291  Worker_OGMatrices_ReadBLASTAndUpdateDistances    -> og_matrices   (M_ij = 0.5 * max(B_ij, Bmin_i) * (1 / Bmax_i))
  :396-411  DendroBLASTTrees.CompleteAndWriteOGMatrices      -> complete      (D_ij = D_ji = -log(M_ij + M_ji), j < i)
  :413-431  DendroBLASTTrees.WritePhylipMatrix               -> phylip_text
B is GetBLAST6Scores with qExcludeSelfHits=False (:272-274): the RAW scores, self hits kept.
"""
import numpy as np

from . import waterfall_oracle as wo


def raw_matrices(n_seqs, hit_text, p, q_double_blast=True):
    """[B_pj for j] as (indptr, cols, vals), blast_file_processor.py:38-104 with qExcludeSelfHits=False."""
    out = []
    for j in range(len(n_seqs)):
        swap = (not q_double_blast) and p > j
        txt = hit_text(j, p) if swap else hit_text(p, j)
        out.append(wo.get_blast6_scores(txt if txt is not None else b"", n_seqs[p], n_seqs[j], False, swap))
    return out


def row_min_max(Bs, n):
    """orthologues.py:275-280 + lil_minmax :226-236: min / max over the stored entries of a gene's rows in every B_pj
    (9e99 / 0 for a gene without hits)."""
    mins = np.full(n, 9e99)
    maxes = np.zeros(n)
    for indptr, _cols, vals in Bs:
        nz = np.flatnonzero(np.diff(indptr))
        if nz.size:
            mins[nz] = np.minimum(mins[nz], np.minimum.reduceat(vals, indptr[nz]))
            maxes[nz] = np.maximum(maxes[nz], np.maximum.reduceat(vals, indptr[nz]))
    return mins, maxes


def _lookup(B, n_j, rows, cols):
    """B[rows[:, None], cols[None, :]] of a CSR triple, 0 where absent."""
    indptr, c, v = B
    out = np.zeros((rows.size, cols.size))
    if c.size == 0 or rows.size == 0 or cols.size == 0:
        return out
    key = np.repeat(np.arange(indptr.size - 1, dtype=np.int64), np.diff(indptr)) * n_j + c
    want = rows[:, None].astype(np.int64) * n_j + cols[None, :]
    pos = np.searchsorted(key, want)
    pos_c = np.minimum(pos, key.size - 1)
    hit = key[pos_c] == want
    out[hit] = v[pos_c[hit]]
    return out


def og_matrices(n_seqs, hit_text, ogs, q_double_blast=True):
    """ogs: list of orthogroups, each a list of (species position, iSeq).  Returns the list of n x n matrices M
    (orthologues.py:264-291)."""
    S = len(n_seqs)
    mats = [np.zeros((len(og), len(og))) for og in ogs]
    members = [[(np.array([i for i, g in enumerate(og) if g[0] == p], np.int64),
                 np.array([g[1] for g in og if g[0] == p], np.int64)) for p in range(S)] for og in ogs]
    with np.errstate(divide="ignore", invalid="ignore"):
        for p in range(S):
            Bs = raw_matrices(n_seqs, hit_text, p, q_double_blast)
            mins, maxes = row_min_max(Bs, n_seqs[p])
            maxes_inv = 1. / maxes
            for j in range(S):
                for m, mem in zip(mats, members):
                    ii, qi = mem[p]
                    jj, qj = mem[j]
                    if ii.size == 0 or jj.size == 0:
                        continue
                    b = _lookup(Bs[j], n_seqs[j], qi, qj)
                    m[np.ix_(ii, jj)] = (0.5 * np.maximum(b, mins[qi][:, None])) * maxes_inv[qi][:, None]
    return mats


def complete(m):
    """orthologues.py:396-410: D (symmetric, the diagonal keeps M's) and max_og."""
    n = m.shape[0]
    d = m.copy()
    max_og = -9e99
    with np.errstate(divide="ignore"):
        for i in range(n):
            for j in range(i):
                d[i, j] = -np.log(m[i, j] + m[j, i])
                d[j, i] = d[i, j]
                max_og = max(max_og, d[i, j])
    return d, max_og


def phylip_text(d, names, max_og):
    """orthologues.py:413-431 WritePhylipMatrix."""
    max_og = 1.1 * max_og
    sliver = 1e-6
    n = d.shape[0]
    lines = ["%d\n" % n]
    for i in range(n):
        V = [0. + (0. if i == j else d[i][j] if d[i][j] > -9e99 else max_og) for j in range(n)]
        V = [sliver if 0 < v < sliver else v for v in V]
        lines.append(names[i] + " " + " ".join(["%.6f" % v for v in V]) + "\n")
    return "".join(lines).encode()
