"""Multi-process driver of the CPU oracle (TEST INFRASTRUCTURE: used only by bench.py's reference arm).

The reference parallelises over query species with multiprocessing (gathering.py:480-509); parse +
normalisation of one species pair depend only on that pair's file and the two length vectors, so the
same work is farmed out per pair here (finer grain, same arithmetic), the remaining per-species steps
run in the parent exactly as in oracle/waterfall_oracle.py.
"""
import multiprocessing as mp
import os

import numpy as np

from . import waterfall_oracle as wo

_G = {}


def _pair_task(args):
    p, j = args
    n_seqs, lengths, texts = _G["n_seqs"], _G["lengths"], _G["texts"]
    raw = wo.get_blast6_scores(texts[(p, j)], n_seqs[p], n_seqs[j], p == j)
    B, fit = wo.normalise_scores(raw, lengths[p], lengths[j])
    return p, j, B, fit


def waterfall_parallel(n_seqs, lengths, texts, n_procs=None, want_text=True):
    """Same result as wo.waterfall(); texts is a dict {(p, j): bytes} inherited by the forked workers."""
    S = len(n_seqs)
    n_procs = n_procs or min(os.cpu_count() or 1, S * S)
    _G.update(n_seqs=n_seqs, lengths=lengths, texts=texts)
    wo.lib()
    tasks = sorted(((p, j) for p in range(S) for j in range(S)), key=lambda t: -len(texts[t]))
    B = [[None] * S for _ in range(S)]
    fits = {}
    if n_procs > 1:
        ctx = mp.get_context("fork")
        with ctx.Pool(n_procs) as pool:
            for p, j, b, f in pool.imap_unordered(_pair_task, tasks, chunksize=1):
                B[p][j] = b
                fits[(p, j)] = f
    else:
        for t in tasks:
            p, j, b, f = _pair_task(t)
            B[p][j] = b
            fits[(p, j)] = f
    starts = np.concatenate(([0], np.cumsum(n_seqs)[:-1])).astype(np.int64)
    BH = [wo.get_bh(B[p], p) for p in range(S)]
    conn = []
    for p in range(S):
        c, _ = wo.connect_cognates(B, BH, p, n_seqs)
        conn.append(c)
    indptr, indices, data = wo.assemble_graph(B, conn, starts, n_seqs)
    out = dict(indptr=indptr, indices=indices, data=data, fits=fits)
    if want_text:
        out["graph"] = wo.graph_text(indptr, indices, data, int(sum(n_seqs)))
    return out
