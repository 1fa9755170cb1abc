"""Device time of the MCL step on the graph of a synthetic job, taken from device memory (GPU box):
    python tools/mcl_timing.py C2_small 1.2 [C2 ...]   -> one JSON line per configuration"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from orthofinder_b200 import mcl as pm, synth, waterfall   # noqa: E402

inflation = 1.2
for arg in sys.argv[1:]:
    try:
        inflation = float(arg)
        continue
    except ValueError:
        pass
    ids, n_seqs, L, P = synth.config_inputs(arg)
    gb = waterfall.GraphBuilder(0)
    gb.set_species(n_seqs, L)
    gb.synth_hits(P, ids)
    gb.build()
    c = gb.counts()
    t0 = time.time()
    res = pm.MCL.RunMCLOnBuilder(gb, None, inflation)
    wall = time.time() - t0
    lab = res.pop("labels")
    print(json.dumps(dict(config=arg, inflation=inflation, genes=int(sum(n_seqs)), edges=c["edges"], wall_s=round(wall, 3), **res)), flush=True)
    gb.close()
