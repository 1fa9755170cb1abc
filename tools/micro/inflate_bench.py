"""inflate_kernel alone: S species x 20 000 genes of synthetic hit text, compressed on the host as BGZF-style members (zlib
level 1), inflated on the device; prints the kernel's time from the context's kernel report.  GPU box only.
  python tools/micro/inflate_bench.py [S] [reps]"""
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

sys.path.insert(0, ".")
from orthofinder_b200 import _lib, synth, waterfall  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 6
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
ids, n_seqs, L, P = synth.config_inputs("C2")
n_seqs, L = n_seqs[:S], L[:S]
gb = waterfall.GraphBuilder(0)
gb.set_species(n_seqs, L)
gb.synth_hits(P, list(range(S)))
texts = {(p, j): gb.hits_text(p, j) for p in range(S) for j in range(S)}
raw = sum(len(t) for t in texts.values())
t0 = time.perf_counter()
with ThreadPoolExecutor(16) as pool:
    blobs = dict(zip(texts, pool.map(lambda t: waterfall.gzip_members(np.frombuffer(t, np.uint8), level=1), texts.values())))
print("compressed %.2f GB -> %.2f GB in %.1f s" % (raw / 1e9, sum(len(b) for b in blobs.values()) / 1e9, time.perf_counter() - t0), flush=True)
for r in range(reps):
    gb.clear_hits()
    for (p, j), z in blobs.items():
        gb.add_hits_gz(p, j, z)
    gb.build(_lib.STAGE_RAW)
    gb.sync()
    rep = gb.kernel_report()
    n, ms = rep["inflate_kernel"]
    print("rep %d: inflate_kernel %d launch %.2f ms = %.1f GB/s of text (parse %.2f ms)" % (r, n, ms, raw / ms / 1e6, rep["parse_kernel"][1]), flush=True)
assert gb.hits_text(1, 2) == texts[(1, 2)] and gb.hits_text(S - 1, 0) == texts[(S - 1, 0)]
gb.close()
