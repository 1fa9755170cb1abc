#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED live reference (oracle/ref_harness.py).

Runs only where /root/reference exists (the build container).  For every dataset it
  1. runs the reference in-process and keeps its OrthoFinder_graph.txt bytes, the curve_fit
     parameters of every species pair and the nnz of its B / BH / connect pickles,
  2. requires the CPU oracle (oracle/waterfall_oracle.py) to reproduce the graph byte for byte and the
     fit parameters bit for bit (this is what pins the oracle), for the default scores and for `--scores-v2`,
  3. stores inputs (only for the reference's own fixtures — synthetic inputs are regenerated from
     orthofinder_b200/synth.py) and expected outputs, compressed.
Provenance (numpy / scipy versions, CPU features) is recorded in each file.
"""
import os
import sys
import tempfile
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_harness as rh, waterfall_oracle as wo  # noqa: E402
from orthofinder_b200 import synth  # noqa: E402

REF_TESTS = "/root/reference/tests/"
FIXTURES = {
    # name: (directory, store inputs)
    "ref_AddTwoSpecies": REF_TESTS + "ExpectedOutput/AddTwoSpecies/",   # 5 species, 43 221 lines
    "ref_FromTrees": REF_TESTS + "Input/FromTrees/WorkingDirectory/",   # C1: the 4 ExampleData proteomes
    "ref_removeFirst": REF_TESTS + "Input/ExampleDataset_removeFirst/", # '#'-skipped species 0
}
SYNTH = ["C2_tiny", "C5_small"]


def provenance():
    import scipy
    return "numpy %s scipy %s svml_dispatch=%s reference=v3.0.1b1" % (np.__version__, scipy.__version__, wo.numpy_uses_svml())


def run_one(name, wd, store_inputs):
    ref = rh.run_reference(wd)
    mine = wo.waterfall_from_directory(wd)
    assert ref["graph"] == mine["graph"], "%s: oracle graph differs from the reference" % name
    assert np.array_equal(np.concatenate(ref["most_distant"]), np.concatenate(mine["most_distant"])), "%s: thresholds differ" % name
    S = ref["nSpecies"]
    fits = np.full((S, S, 3), np.nan)
    nnz = np.zeros((S, S, 3), np.int64)
    for (p, j), (pars, m, raw_nnz) in ref["fits"].items():
        f = mine["fits"][(p, j)]
        if pars is not None:
            assert f is not None and f[0] == pars[0] and f[1] == pars[1] and f[2] == m, (name, p, j)
            fits[p, j] = (pars[0], pars[1], m)
        else:
            assert f is None
            fits[p, j, 2] = m
        nnz[p, j] = [ref["mats"][(k, p, j)].nnz for k in ("B", "BH", "connect")]
    out = dict(graph_z=np.frombuffer(zlib.compress(ref["graph"], 9), np.uint8), fits=fits, nnz=nnz,
               n_seqs=np.array([ref["nSeqsPerSpecies"][k] for k in ref["speciesToUse"]], np.int64),
               species=np.array(ref["speciesToUse"], np.int64), most_distant=np.concatenate(ref["most_distant"]),
               provenance=np.array(provenance()))
    if store_inputs:
        out["lengths"] = np.concatenate(ref["lengths"])
        blobs, sizes = [], []
        for p in ref["speciesToUse"]:
            for j in ref["speciesToUse"]:
                with open(os.path.join(wd, "Blast%d_%d.txt" % (p, j)), "rb") as f:
                    b = f.read()
                blobs.append(b)
                sizes.append(len(b))
        out["hits_z"] = np.frombuffer(zlib.compress(b"".join(blobs), 9), np.uint8)
        out["hit_sizes"] = np.array(sizes, np.int64)
    # the `--scores-v2` variant of the same inputs: graph and thresholds of the live reference
    ref2 = rh.run_reference(wd, capture_fit=False, keep_matrices=False, v2_scores=True)
    mine2 = wo.waterfall_from_directory(wd, v2_scores=True)
    assert ref2["graph"] == mine2["graph"], "%s: oracle v2 graph differs from the reference" % name
    out["graph_v2_z"] = np.frombuffer(zlib.compress(ref2["graph"], 9), np.uint8)
    assert np.array_equal(np.concatenate(ref2["most_distant"]), np.concatenate(mine2["most_distant"])), "%s: v2 thresholds differ" % name
    out["most_distant_v2"] = np.concatenate(ref2["most_distant"])
    # the `--c-homologs` graph (gathering_version (3, 2), gathering.py:51-73): symmetrised structure of the normalised matrices
    ref3 = rh.run_reference(wd, capture_fit=False, keep_matrices=False, homology=True)
    mine3 = wo.waterfall_from_directory(wd, homology=True)
    assert ref3["graph"] == mine3["graph"], "%s: oracle homology graph differs from the reference" % name
    out["graph_hom_z"] = np.frombuffer(zlib.compress(ref3["graph"], 9), np.uint8)
    fn = os.path.join(ROOT, "tests", "golden", name + ".npz")
    np.savez(fn, **out)
    print("%-20s S=%d genes=%d graph=%d B  entries=%d  -> %s (%d KB)" % (
        name, S, ref["nSeqs"], len(ref["graph"]), ref["graph"].count(b":"), os.path.relpath(fn, ROOT),
        os.path.getsize(fn) // 1024))


def main():
    for name, wd in FIXTURES.items():
        run_one(name, wd, True)
    for name in SYNTH:
        wd = tempfile.mkdtemp(prefix="ofg_golden_") + "/"
        synth.write_working_directory(wd, name)
        run_one("synth_" + name, wd, False)


if __name__ == "__main__":
    main()
