#!/usr/bin/env python
"""Generate tests/golden/dendro_*.npz: DendroBLAST distance matrices of the UNMODIFIED live reference.

Runs only where /root/reference exists.  For every dataset of tools/make_golden.py it
  1. builds orthogroups from the golden graph (connected components with >= 4 genes, large ones cut into pieces of 80,
     sizes non-increasing as orthologues.py:331-333 asserts) plus one group of random genes and, where they exist, one
     group of genes without any hit (the 1/0 -> inf -> max_og branch of WritePhylipMatrix),
  2. runs the reference's own worker, completion and Phylip writer in-process (oracle/ref_harness.run_reference_dendroblast),
  3. requires oracle/dendroblast_oracle.py to reproduce M and D bit for bit and the Phylip text byte for byte,
  4. stores orthogroups and expected outputs, compressed.
"""
import os
import sys
import tempfile
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_harness as rh, dendroblast_oracle as do  # noqa: E402
from orthofinder_b200 import synth  # noqa: E402
from tests import helpers  # noqa: E402
from tools.make_golden import FIXTURES, SYNTH, provenance  # noqa: E402

MAX_OG = 80
MAX_OGS = 48
MAX_ENTRIES = 40000   # sum of n^2 over the orthogroups kept per dataset (file size)


def orthogroups_from_graph(g, rng):
    """list of orthogroups, each a list of (species position, iSeq)."""
    n_seqs = g.n_seqs
    starts = np.concatenate(([0], np.cumsum(n_seqs)))
    n = int(starts[-1])
    parent = np.arange(n)

    def find(x):
        while parent[x] != x:
            parent[x] = parent[parent[x]]
            x = parent[x]
        return x
    _n, rows = helpers.parse_graph_text(g.graph)
    has_edge = np.zeros(n, bool)
    for r, edges in rows:
        for c, _w in edges:
            c = int(c)
            has_edge[r] = has_edge[c] = True
            a, b = find(r), find(c)
            if a != b:
                parent[max(a, b)] = min(a, b)
    comp = {}
    for x in range(n):
        comp.setdefault(find(x), []).append(x)
    groups = []
    for members in comp.values():
        for k in range(0, len(members), MAX_OG):
            piece = members[k:k + MAX_OG]
            if len(piece) >= 4:
                groups.append(piece)
    groups.sort(key=lambda m: (-len(m), m[0]))
    groups = groups[:MAX_OGS]
    while len(groups) > 6 and sum(len(m) ** 2 for m in groups) > MAX_ENTRIES:
        groups.pop(len(groups) // 2)
    groups.append(sorted(rng.choice(n, size=min(12, n), replace=False).tolist()))
    lonely = np.flatnonzero(~has_edge)
    if lonely.size >= 4:
        groups.append(lonely[:6].tolist())
    groups.sort(key=lambda m: -len(m))

    def split(x):
        p = int(np.searchsorted(starts, x, side="right") - 1)
        return p, int(x - starts[p])
    return [[split(x) for x in m] for m in groups]


def run_one(name, wd):
    g = helpers.Golden(name)
    rng = np.random.default_rng(20240607)
    ogs = orthogroups_from_graph(g, rng)
    ogs_ids = [[(g.species[p], q) for p, q in og] for og in ogs]
    ref = rh.run_reference_dendroblast(wd, ogs_ids)
    M = do.og_matrices(g.n_seqs, g.hit_text, ogs)
    n_inf = 0
    phy = []
    for iog, (og, m) in enumerate(zip(ogs_ids, M)):
        assert np.array_equal(m, ref["M"][iog]), "%s og %d: oracle M differs from the reference" % (name, iog)
        d, max_og = do.complete(m)
        assert np.array_equal(d, ref["D"][iog]), "%s og %d: oracle D differs" % (name, iog)
        txt = do.phylip_text(d, ["%d_%d" % x for x in og], max_og)
        assert txt == ref["phylip"][iog], "%s og %d: Phylip text differs" % (name, iog)
        phy.append(txt)
        n_inf += int(np.isinf(d).sum())
    out = dict(og_off=np.cumsum([0] + [len(og) for og in ogs]).astype(np.int64),
               og_sp=np.array([p for og in ogs for p, _ in og], np.int32), og_seq=np.array([q for og in ogs for _, q in og], np.int32),
               M=np.concatenate([m.ravel() for m in ref["M"]]), D=np.concatenate([d.ravel() for d in ref["D"]]),
               phy_z=np.frombuffer(zlib.compress(b"".join(phy), 9), np.uint8), phy_sizes=np.array([len(t) for t in phy], np.int64),
               provenance=np.array(provenance()))
    fn = os.path.join(ROOT, "tests", "golden", "dendro_" + name + ".npz")
    np.savez_compressed(fn, **out)
    print("%-20s orthogroups=%d members=%d entries=%d inf=%d -> %s (%d KB)" % (
        name, len(ogs), out["og_sp"].size, out["M"].size, n_inf, os.path.relpath(fn, ROOT), os.path.getsize(fn) // 1024))


def main():
    for name, wd in FIXTURES.items():
        run_one(name, wd)
    for name in SYNTH:
        wd = tempfile.mkdtemp(prefix="ofg_golden_") + "/"
        synth.write_working_directory(wd, name)
        run_one("synth_" + name, wd)


if __name__ == "__main__":
    main()
