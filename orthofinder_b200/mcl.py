"""Host mirror of the reference's MCL step over the `ofg_mcl_*` C-ABI of libofg.so.

Reference interface: `MCL.RunMCL(graphFilename, clustersFilename, nProcesses, inflation)` (scripts_of/mcl.py:214-218), which
starts the bundled binary `mcl graph -I inflation -o clusters -te nProcesses -V all`, called from DoOrthogroups
(gathering.py:517, :523).  Here the same call clusters on the GPU and writes the same clusters file; `RunMCLOnBuilder`
takes the graph straight from the device memory of a finished `GraphBuilder` (no text round trip).  There is no CPU
fallback: libofg.so and a CUDA device are required.
"""
import ctypes
import os
import re

import numpy as np

from . import _lib

MCL_HEADER_COMMENT = "# cline: ofg_mcl (orthofinder_b200) -I %s\n"   # the binary writes its own command line here; readers skip to "begin"


def read_graph_text(text):
    """OrthoFinder_graph.txt (WriteGraphParallel, gathering.py:273-291) -> (n, colptr int64, rows int32, vals float32):
    line c of the file is column c of MCL's matrix; a weight becomes what the binary stores for it (strtod, then float)."""
    if isinstance(text, (bytes, bytearray, memoryview)):
        text = bytes(text).decode()
    n = int(re.search(r"dimensions\s+(\d+)x(\d+)", text).group(1))
    body = text[text.index("begin") + 5:text.rindex(")")]
    counts = np.zeros(n, np.int64)
    rows, vals = [], []
    last = -1
    for rec in body.split("$"):
        t = rec.split()
        if not t:
            continue
        c = int(t[0])
        if c <= last or c >= n:
            raise ValueError("graph rows must ascend and stay below the dimension (row %d)" % c)
        last = c
        ent = [e.split(":") for e in t[1:]]
        rows.append(np.array([int(r) for r, _ in ent], np.int32))
        vals.append(np.array([float(w) for _, w in ent], np.float64).astype(np.float32))
        counts[c] = len(ent)
    colptr = np.zeros(n + 1, np.int64)
    np.cumsum(counts, out=colptr[1:])
    rows = np.concatenate(rows) if rows else np.zeros(0, np.int32)
    vals = np.concatenate(vals) if vals else np.zeros(0, np.float32)
    return n, colptr, np.ascontiguousarray(rows), np.ascontiguousarray(vals)


class DeviceMCL:
    """One `ofg_mcl` handle (one GPU)."""

    def __init__(self, device=0):
        self._L = _lib.load()
        h = ctypes.c_void_p()
        rc = self._L.ofg_mcl_create(ctypes.byref(h), int(device))
        if rc != _lib.OFG_OK:
            raise RuntimeError("ofg_mcl_create failed (%s): a CUDA device is required, there is no CPU fallback" % _lib.ERR_NAMES.get(rc, rc))
        self._h = h
        self.n = 0

    def _ck(self, rc):
        if rc != _lib.OFG_OK:
            msg = self._L.ofg_mcl_last_error(self._h)
            raise RuntimeError("%s: %s" % (_lib.ERR_NAMES.get(rc, rc), msg.decode() if msg else ""))

    def set_params(self, prune_number=10000, select_number=1100, recover_number=1400, recover_pct=90.0, sparse_overhead=10, max_iterations=10000):
        self._ck(self._L.ofg_mcl_set_params(self._h, prune_number, select_number, recover_number, float(recover_pct), sparse_overhead, max_iterations))

    def set_budgets(self, table_slots=1 << 28, stage_entries=1 << 27, direct_factor=4):
        self._ck(self._L.ofg_mcl_set_budgets(self._h, int(table_slots), int(stage_entries), int(direct_factor)))

    def set_graph(self, n, colptr, rows, vals):
        colptr = np.ascontiguousarray(colptr, np.int64)
        rows = np.ascontiguousarray(rows, np.int32)
        vals = np.ascontiguousarray(vals, np.float32)
        self._ck(self._L.ofg_mcl_set_graph_host(self._h, int(n), colptr.ctypes.data, rows.ctypes.data, vals.ctypes.data))
        self.n = int(n)

    def set_graph_text(self, text):
        """The bytes of a graph file (parsed by the library on the host)."""
        text = bytes(text)
        self._ck(self._L.ofg_mcl_set_graph_text(self._h, text, len(text)))
        m = re.search(rb"dimensions\s+(\d+)x", text)
        self.n = int(m.group(1))

    def set_graph_from_builder(self, builder):
        """The finished graph of a waterfall.GraphBuilder, taken from device memory (same GPU)."""
        n, nnz = ctypes.c_int64(0), ctypes.c_int64(0)
        off, idx, w = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p()
        rc = self._L.ofg_graph_dev(builder.h, ctypes.byref(n), ctypes.byref(nnz), ctypes.byref(off), ctypes.byref(idx), ctypes.byref(w))
        if rc != _lib.OFG_OK:
            msg = self._L.ofg_last_error(builder.h)
            raise RuntimeError("%s: %s" % (_lib.ERR_NAMES.get(rc, rc), msg.decode() if msg else ""))
        self._ck(self._L.ofg_mcl_set_graph_dev(self._h, n.value, off, nnz.value, idx, w))
        self.n = int(n.value)

    def run(self, inflation):
        self._ck(self._L.ofg_mcl_run(self._h, float(inflation)))
        return self.result()

    def step(self, inflation):
        ch = ctypes.c_double(0.0)
        self._ck(self._L.ofg_mcl_step(self._h, float(inflation), ctypes.byref(ch)))
        return ch.value

    def interpret(self):
        self._ck(self._L.ofg_mcl_interpret(self._h))
        return self.result()

    def result(self):
        it, nc, ov, nnz, nl = (ctypes.c_int64(0) for _ in range(5))
        ch, ms = ctypes.c_double(0.0), ctypes.c_float(0.0)
        self._ck(self._L.ofg_mcl_result(self._h, ctypes.byref(it), ctypes.byref(nc), ctypes.byref(ov), ctypes.byref(ch), ctypes.byref(nnz),
                                        ctypes.byref(ms), ctypes.byref(nl)))
        return dict(iterations=it.value, n_clusters=nc.value, overlap=ov.value, chaos=ch.value, nnz=nnz.value, ms=ms.value, launches=nl.value)

    def iterand(self):
        nnz = self.result()["nnz"]
        cp, ri, v = np.empty(self.n + 1, np.int64), np.empty(nnz, np.int32), np.empty(nnz, np.float32)
        self._ck(self._L.ofg_mcl_iterand_copy(self._h, cp.ctypes.data, ri.ctypes.data, v.ctypes.data))
        return cp, ri, v

    def labels(self):
        lab = np.empty(self.n, np.int32)
        self._ck(self._L.ofg_mcl_labels(self._h, lab.ctypes.data, self.n))
        return lab

    def clusters_text(self):
        need = ctypes.c_size_t(0)
        self._ck(self._L.ofg_mcl_clusters_text(self._h, None, 0, ctypes.byref(need)))
        buf = ctypes.create_string_buffer(need.value)
        self._ck(self._L.ofg_mcl_clusters_text(self._h, buf, need.value, None))
        return buf.raw[:need.value].decode()

    def close(self):
        if self._h:
            self._L.ofg_mcl_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _write_clusters(clustersFilename, text, inflation):
    with open(clustersFilename, "w") as f:
        f.write(MCL_HEADER_COMMENT % inflation)
        f.write(text)


class MCL:
    """Same entry point as the reference's mcl.MCL for this step (scripts_of/mcl.py:214-218)."""

    @staticmethod
    def RunMCL(graphFilename, clustersFilename, nProcesses, inflation, device=0):
        """Clusters the graph file on the GPU and writes the clusters file the binary would have written (its body; the
        first line, a comment with the command line, differs).  `nProcesses` (the binary's -te) has no meaning here and
        does not change the result there either."""
        with open(graphFilename, "rb") as f:
            text = f.read()
        m = DeviceMCL(device)
        try:
            m.set_graph_text(text)
            res = m.run(inflation)
            _write_clusters(clustersFilename, m.clusters_text(), inflation)
        finally:
            m.close()
        return res

    @staticmethod
    def RunMCLOnBuilder(builder, clustersFilename, inflation, device=0):
        """As RunMCL, for the graph a waterfall.GraphBuilder holds in device memory (after build()): the weights go
        through the same "%.3f" rounding on the device, so the result is the one the graph FILE would give."""
        m = DeviceMCL(device)
        try:
            m.set_graph_from_builder(builder)
            res = m.run(inflation)
            if clustersFilename:
                _write_clusters(clustersFilename, m.clusters_text(), inflation)
            res["labels"] = m.labels()
        finally:
            m.close()
        return res


def GetPredictedOGs(clustersFilename):
    """Reader of the clusters file with the reference's semantics (scripts_of/mcl.py:36-63): list of sets of id strings."""
    ogs, og, header = [], None, True
    with open(clustersFilename) as f:
        for line in f:
            if header:
                if "begin" in line:
                    header = False
                continue
            if ")" in line:
                break
            if line[-2] == "$":
                line = line[:-3]
            if line[0] == " ":
                og.update(line.split())
            else:
                if og is not None:
                    ogs.append(og)
                og = set(line.split()[1:])
        if og is not None:
            ogs.append(og)
    return ogs
