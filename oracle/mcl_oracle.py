"""CPU oracle of the MCL step that consumes the graph - TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module; the
product (orthofinder_b200/) never does.

What it restates: the reference runs `mcl graph -I <inflation> -o clusters -te <n> -V all` (scripts_of/mcl.py:214-218)
with the bundled binary scripts_of/bin/mcl (mcl 14-137, no source in the reference tree), reads the clusters with
`GetPredictedOGs` (mcl.py:36-63) and rewrites the ids with `ConvertSingleIDsToIDPair` (mcl.py:93-125).  The process itself
is van Dongen's published algorithm; its arithmetic is restated in oracle/csrc/mcl_process.c.

Parity status: PINNED against the binary run in the build container (tests/test_reference_pin.py, marker `reference`):
every iterand of the five golden graphs bit for bit (`-dump ite --write-binary`), ten pruning parameter sets on the first
expansion (`-write-expanded`), the clusters file body byte for byte.  tests/golden/mcl_*.npz hold the binary's clusters
for the GPU box (tools/make_golden_mcl.py).
"""
import ctypes
import os
import re

import numpy as np

from . import waterfall_oracle as _wo

DEFAULTS = dict(P=10000, S=1100, R=1400, pct=0.9, overhead=10, max_iter=10000, chaos_limit=1e-4)

_LIB = None


def _lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(_wo.build_lib())
        L.mclo_new.restype = ctypes.c_void_p
        L.mclo_new.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        L.mclo_free.argtypes = [ctypes.c_void_p]
        L.mclo_step.restype = ctypes.c_double
        L.mclo_step.argtypes = [ctypes.c_void_p, ctypes.c_double, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double,
                                ctypes.c_int]
        L.mclo_nnz.restype = ctypes.c_int64
        L.mclo_nnz.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.mclo_get.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        L.mclo_interpret.restype = ctypes.c_int
        L.mclo_interpret.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        _LIB = L
    return _LIB


def parse_graph_text(text):
    """The MCL 'matrix' text that WriteGraphParallel produces (gathering.py:273-291) -> (n, colptr i64, rows i32,
    vals f32).  mcl reads a value with strtod and stores it as float."""
    if isinstance(text, (bytes, bytearray, memoryview)):
        text = bytes(text).decode()
    n = int(re.search(r"dimensions (\d+)x(\d+)", text).group(1))
    body = text[text.index("begin") + 5:text.rindex(")")]
    colptr = np.zeros(n + 1, np.int64)
    rows, vals = [], []
    cols = {}
    for rec in body.split("$"):
        t = rec.split()
        if not t:
            continue
        cols[int(t[0])] = t[1:]
    for c in range(n):
        for e in cols.get(c, ()):
            r, w = e.split(":")
            rows.append(int(r))
            vals.append(float(w))
        colptr[c + 1] = len(rows)
    return n, colptr, np.asarray(rows, np.int32), np.asarray(vals, np.float64).astype(np.float32)


class Process:
    """Step-wise access to the restated process (for the pin against the binary's dumps)."""

    def __init__(self, n, colptr, rows, vals):
        self.n = int(n)
        self._cp = np.ascontiguousarray(colptr, np.int64)
        self._ri = np.ascontiguousarray(rows, np.int32)
        self._v = np.ascontiguousarray(vals, np.float32)
        self._h = _lib().mclo_new(self.n, self._cp.ctypes.data, self._ri.ctypes.data, self._v.ctypes.data)

    def step(self, inflation, **kw):
        p = dict(DEFAULTS, **kw)
        return _lib().mclo_step(self._h, float(inflation), p["P"], p["S"], p["R"], p["pct"], p["overhead"])

    def matrix(self, which=0):
        nnz = _lib().mclo_nnz(self._h, which)
        cp = np.empty(self.n + 1, np.int64)
        ri = np.empty(nnz, np.int32)
        v = np.empty(nnz, np.float32)
        _lib().mclo_get(self._h, which, cp.ctypes.data, ri.ctypes.data, v.ctypes.data)
        return cp, ri, v

    def interpret(self, which=0):
        label = np.empty(self.n, np.int32)
        ov = ctypes.c_int32(0)
        nc = _lib().mclo_interpret(self._h, which, label.ctypes.data, ctypes.byref(ov))
        return nc, label, ov.value

    def close(self):
        if self._h:
            _lib().mclo_free(self._h)
            self._h = None

    def __del__(self):
        self.close()


def run(n, colptr, rows, vals, inflation, **kw):
    """The whole process: returns dict(labels, n_clusters, iterations, chaos, overlap)."""
    p = dict(DEFAULTS, **kw)
    pr = Process(n, colptr, rows, vals)
    it, chaos = 0, []
    while it < p["max_iter"]:
        ch = pr.step(inflation, **kw)
        it += 1
        chaos.append(ch)
        if ch < p["chaos_limit"]:
            break
    nc, label, ov = pr.interpret(0)
    pr.close()
    return dict(labels=label, n_clusters=nc, iterations=it, chaos=chaos, overlap=ov)


def clusters_from_labels(labels):
    """List of clusters (ascending members) in output order."""
    order = np.argsort(labels, kind="stable")
    counts = np.bincount(labels)
    return np.split(order.astype(np.int64), np.cumsum(counts)[:-1])


def clusters_text_body(labels, n):
    """The clusters file from `(mclheader` on, as mcl 14-137 writes it (fitted to the binary's files for id widths 1-6):
    ids left-justified to the width w of n, one space before every member, a new line indented by w + 2 after a member
    that brings the line to 70 - w characters or more, the closing ` $` never on a line of its own."""
    cl = clusters_from_labels(labels)
    out = ["(mclheader\nmcltype matrix\ndimensions %dx%d\n)\n(mclmatrix\nbegin\n" % (n, len(cl))]
    w = len(str(n))
    for i, members in enumerate(cl):
        line = str(i).ljust(w)
        toks = [str(m) for m in members.tolist()]
        for k, t in enumerate(toks):
            line += " " + t
            if len(line) + w >= 70 and k < len(toks) - 1:
                out.append(line + "\n")
                line = " " * (w + 2)
        out.append(line + " $\n")
    out.append(")\n")
    return "".join(out)


def clusters_body_of_file(text):
    """What the parity gate compares: the binary's file without its `# cline:` comment line."""
    return text[text.index("(mclheader"):]
