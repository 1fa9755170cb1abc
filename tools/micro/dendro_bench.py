"""Timing of the DendroBLAST matrices on the C2 workload: raw stage with self hits + ofg_dendro_matrices for
20 000 orthogroups of 16 genes (gene q of every species) and for 50 orthogroups of 1 600 genes.  GPU box only."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from orthofinder_b200 import _lib, synth, waterfall  # noqa: E402

ids, n_seqs, L, P = synth.config_inputs("C2")
S, G = len(n_seqs), n_seqs[0]
gb = waterfall.GraphBuilder(0)
gb.set_species(n_seqs, L)
gb.set_option("keep_self_hits", 1)
gb.synth_hits(P, list(range(S)))
for _ in range(2):
    t0 = time.perf_counter()
    gb.build(_lib.STAGE_RAW)
    gb.sync()
    t_raw = time.perf_counter() - t0
print("raw stage (parse + rows, self hits kept): %.1f ms for %d lines" % (t_raw * 1e3, gb.counts()["lines"]))
for label, ogs in (("20000 orthogroups x 16 genes", [[(p, q) for p in range(S)] for q in range(G)]),
                   ("50 orthogroups x 1600 genes", [[(p, q) for p in range(S) for q in range(100 * k, 100 * k + 100)] for k in range(50)])):
    og_off = np.cumsum([0] + [len(o) for o in ogs])
    sp = np.array([p for o in ogs for p, _ in o], np.int32)
    sq = np.array([q for o in ogs for _, q in o], np.int32)
    for complete in (False, True):
        gb.dendro_matrices(og_off, sp, sq, complete)
        t0 = time.perf_counter()
        flat, mx = gb.dendro_matrices(og_off, sp, sq, complete)
        dt = time.perf_counter() - t0
        print("%s complete=%d: %.1f ms wall incl. uploads and the D2H of %.1f MB (%d matrix elements, %.0f M elements/s)" % (
            label, complete, dt * 1e3, flat.nbytes / 1e6, flat.size, flat.size / dt / 1e6))
gb.close()
