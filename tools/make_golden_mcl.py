"""Golden vectors of the MCL step from the reference's bundled binary, for the GPU box (where /root/reference
does not exist).  For every golden graph of tests/golden/ (the live reference's OrthoFinder_graph.txt) and two synthetic
graphs that reach the selection / recovery limits of the default resource scheme, runs
`mcl graph -I <inflation> -o clusters -te 1 -V all` (scripts_of/mcl.py:214-218) and stores the clusters file body, the
number of iterations and, for the start matrix and every iterand, the entry count and a SHA-256 over its column
pointers, row indices and float bit patterns (tests.helpers.matrix_digest) in tests/golden/mcl_<name>.npz.

    python tools/make_golden_mcl.py
"""
import os
import sys
import tempfile
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import mcl_oracle as mo, ref_harness as rh   # noqa: E402
from tests.helpers import GOLDEN_DIR, Golden, REF_GOLDENS, SYNTH_GOLDENS, matrix_digest, mcl_synthetic_graph   # noqa: E402

INFLATIONS = (1.2, 1.5, 3.0)   # the reference's default (__main__.py:223), the value its stored test results used, a coarse one


def pack(name, graph_text):
    out = {"provenance": "mcl 14-137 (scripts_of/bin/mcl), -te 1 -V all, default scheme", "inflations": np.array(INFLATIONS)}
    for I in INFLATIONS:
        with tempfile.TemporaryDirectory() as td:
            r = rh.run_mcl_binary(graph_text, I, td, dump_iterands=True)
        its = r["iterands"]
        tag = "I%s" % str(I).replace(".", "_")
        out[tag + "_clusters_z"] = np.frombuffer(zlib.compress(mo.clusters_body_of_file(r["clusters_text"]).encode(), 9), np.uint8)
        out[tag + "_iterations"] = np.int64(len(its) - 1)
        out[tag + "_nnz"] = np.array([len(ri) for _, ri, _ in its], np.int64)            # start matrix, then every iterand
        out[tag + "_sha256"] = np.array([matrix_digest(*m) for m in its])
        print(name, I, "iterations", len(its) - 1, "nnz after 1:", len(its[1][1]), "final:", len(its[-1][1]))
    np.savez_compressed(os.path.join(GOLDEN_DIR, "mcl_%s.npz" % name), **out)


if __name__ == "__main__":
    for g in REF_GOLDENS + SYNTH_GOLDENS:
        pack(g, Golden(g).graph)
    for kind in ("family", "ties"):
        pack("synth_" + kind, mcl_synthetic_graph(kind))
