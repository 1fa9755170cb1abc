"""Host side of the B200 orthogroup-graph builder: a drop-in for the hot path of
gathering.DoOrthogroups (reference scripts_of/gathering.py:476-512).

Two layers:
  * GraphBuilder      — thin object wrapper over the C-ABI context (include/ofg.h).
  * WaterfallMethod / DoOrthogroupsGraph — the reference-facing interface: same names, argument
    meaning and error behaviour as WaterfallMethod.ProcessBlastHits / ConnectCognates /
    WriteGraphParallel, taking a util.SequencesInfo-shaped object and the blast directory list and
    producing OrthoFinder_graph.txt.

All numerics run in hand-written CUDA kernels inside libofg.so; this module only moves bytes
(file -> pinned host -> device, device -> file) and translates error codes into the reference's
messages.  There is no CPU fallback.
"""
import ctypes
import gzip
import zlib
import os
import sys
from collections import namedtuple

import numpy as np

from . import _lib

# util.py:54-61
SequencesInfo = namedtuple("SequencesInfo", "nSeqs nSpecies speciesToUse seqStartingIndices nSeqsPerSpecies")


class OfgError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("%s: %s" % (_lib.ERR_NAMES.get(code, str(code)), message))
        self.code = code
        self.message = message


def _vp(a):
    return a.ctypes.data_as(ctypes.c_void_p)


class GraphBuilder:
    """One context per GPU (include/ofg.h).  Not thread-safe."""

    def __init__(self, device=0):
        self.lib = _lib.load()
        h = ctypes.c_void_p()
        rc = self.lib.ofg_create(ctypes.byref(h), int(device))
        if rc != 0:
            raise OfgError(rc, "ofg_create failed on device %d (is a CUDA device visible?)" % device)
        self.h = h
        self.device = device
        self.S = 0
        self.n_seqs = []

    def close(self):
        if getattr(self, "h", None):
            self.lib.ofg_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise OfgError(rc, self.lib.ofg_last_error(self.h).decode(errors="replace"))

    # ---- inputs
    def set_species(self, n_seqs, lengths):
        """n_seqs[p], lengths[p] (fp64 vectors) per used species in list order
        (util.SequencesInfo + gathering.GetSequenceLengths)."""
        n = np.ascontiguousarray(n_seqs, dtype=np.int64)
        L = np.ascontiguousarray(np.concatenate([np.asarray(x, dtype=np.float64) for x in lengths]) if len(lengths)
                                 else np.zeros(0))
        assert L.size == int(n.sum()), "lengths do not match n_seqs"
        self._ck(self.lib.ofg_set_species(self.h, len(n), _vp(n), _vp(L)))
        self.S = len(n)
        self.n_seqs = [int(x) for x in n]
        self._lengths = L
        self._factors_sent = False

    def set_partition(self, owned_rows, with_columns):
        o = np.ascontiguousarray(owned_rows, dtype=np.int32)
        self._ck(self.lib.ofg_set_partition(self.h, len(o), _vp(o), int(bool(with_columns))))

    def set_option(self, name, value):
        self._ck(self.lib.ofg_set_option(self.h, name.encode(), int(value)))
        if name == "scores_v2":
            self._v2 = bool(value)

    def add_hits(self, p, j, data, swap_cols=False):
        """data: bytes-like / numpy uint8 host buffer, or (device_ptr, nbytes) tuple."""
        if isinstance(data, tuple):
            ptr, n = data
            self._ck(self.lib.ofg_add_hits_text(self.h, p, j, ctypes.c_void_p(ptr), n, int(swap_cols), 1))
            return
        if isinstance(data, np.ndarray):
            buf = np.ascontiguousarray(data, dtype=np.uint8)
            self._ck(self.lib.ofg_add_hits_text(self.h, p, j, _vp(buf), buf.size, int(swap_cols), 0))
            return
        b = bytes(data)
        self._ck(self.lib.ofg_add_hits_text(self.h, p, j, ctypes.cast(ctypes.c_char_p(b), ctypes.c_void_p), len(b),
                                            int(swap_cols), 0))

    def add_hits_gz(self, p, j, data, swap_cols=False):
        """data: the bytes of Blast{p}_{j}.txt.gz as on disk (bytes / numpy uint8) or a (host pointer, nbytes) tuple; inflated
        on the device by ofg_build (one warp per gzip member)."""
        if isinstance(data, tuple):
            self._ck(self.lib.ofg_add_hits_gz(self.h, p, j, ctypes.c_void_p(data[0]), int(data[1]), int(swap_cols)))
            return
        if isinstance(data, np.ndarray):
            buf = np.ascontiguousarray(data, dtype=np.uint8)
            self._ck(self.lib.ofg_add_hits_gz(self.h, p, j, _vp(buf), buf.size, int(swap_cols)))
            return
        b = bytes(data)
        self._ck(self.lib.ofg_add_hits_gz(self.h, p, j, ctypes.cast(ctypes.c_char_p(b), ctypes.c_void_p), len(b), int(swap_cols)))

    def add_hits_ptr(self, p, j, host_ptr, nbytes, swap_cols=False):
        self._ck(self.lib.ofg_add_hits_text(self.h, p, j, ctypes.c_void_p(host_ptr), nbytes, int(swap_cols), 0))

    def clear_hits(self):
        self._ck(self.lib.ofg_clear_hits(self.h))

    def synth_hits(self, params, species_ids):
        ids = np.ascontiguousarray(species_ids, dtype=np.int32)
        self._ck(self.lib.ofg_synth_hits(self.h, ctypes.byref(params), _vp(ids)))

    def hits_text(self, p, j):
        n = ctypes.c_size_t(0)
        self._ck(self.lib.ofg_hits_text_size(self.h, p, j, ctypes.byref(n)))
        buf = np.empty(max(n.value, 1), np.uint8)
        self._ck(self.lib.ofg_hits_text_copy(self.h, p, j, _vp(buf), n.value))
        return buf[:n.value].tobytes()

    def hits_text_size(self, p, j):
        n = ctypes.c_size_t(0)
        self._ck(self.lib.ofg_hits_text_size(self.h, p, j, ctypes.byref(n)))
        return n.value

    def hits_text_into(self, p, j, host_ptr, cap):
        self._ck(self.lib.ofg_hits_text_copy(self.h, p, j, ctypes.c_void_p(host_ptr), cap))

    # ---- run
    def build(self, stop_after=_lib.STAGE_GRAPH):
        if getattr(self, "_v2", False) and getattr(self, "_lengths", None) is not None and not self._factors_sent:
            # gathering.py:171-172: li_vals = Lq**(-0.5) with the host's numpy, so the graph matches the reference's bytes
            fac = np.ascontiguousarray(self._lengths ** (-0.5))
            self._ck(self.lib.ofg_set_length_factors(self.h, _vp(fac)))
            self._factors_sent = True
        self._ck(self.lib.ofg_build(self.h, int(stop_after)))

    # ---- outputs
    def counts(self):
        a, b, c, d = (ctypes.c_int64(0) for _ in range(4))
        self._ck(self.lib.ofg_counts(self.h, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c), ctypes.byref(d)))
        return dict(lines=a.value, records=b.value, nnz=c.value, edges=d.value)

    def timing(self):
        ms = np.zeros(10, np.float32)
        n = ctypes.c_int64(0)
        self._ck(self.lib.ofg_last_timing(self.h, _vp(ms), ctypes.byref(n)))
        names = ["parse", "rows", "length_sort", "select", "fit", "scale_bh", "thresholds", "connect", "emit", "total"]
        d = {k: float(v) for k, v in zip(names, ms)}
        d["launches"] = n.value
        return d

    def kernel_report(self):
        """{kernel name: (launches, total ms)} of the last build."""
        buf = ctypes.create_string_buffer(1 << 16)
        self._ck(self.lib.ofg_kernel_report(self.h, buf, len(buf)))
        out = {}
        for line in buf.value.decode().splitlines():
            name, n, ms = line.rsplit(" ", 2)
            out[name] = (int(n), float(ms))
        return out

    def pair_csr(self, p, j):
        nr, nnz = ctypes.c_int64(0), ctypes.c_int64(0)
        self._ck(self.lib.ofg_pair_csr(self.h, p, j, ctypes.byref(nr), ctypes.byref(nnz), None, None, None))
        rl = np.zeros(max(nr.value, 1), np.int32)
        cols = np.zeros(max(nnz.value, 1), np.uint32)
        vals = np.zeros(max(nnz.value, 1), np.float64)
        self._ck(self.lib.ofg_pair_csr(self.h, p, j, ctypes.byref(nr), ctypes.byref(nnz), _vp(rl), _vp(cols), _vp(vals)))
        rl = rl[:nr.value]
        indptr = np.concatenate(([0], np.cumsum(rl, dtype=np.int64)))
        return indptr, cols[:nnz.value], vals[:nnz.value]

    def pair_fit(self, p, j):
        out = np.zeros(6)
        self._ck(self.lib.ofg_pair_fit(self.h, p, j, _vp(out)))
        return dict(a=out[0], b=out[1], m=int(out[2]), info=int(out[3]), nfev=int(out[4]), normalised=bool(out[5]))

    def thresholds(self):
        n = int(sum(self.n_seqs))
        out = np.zeros(max(n, 1))
        self._ck(self.lib.ofg_thresholds_copy(self.h, _vp(out), n))
        return out[:n]

    def thresholds_dev(self):
        p = ctypes.c_void_p()
        n = ctypes.c_int64(0)
        self._ck(self.lib.ofg_thresholds_dev(self.h, ctypes.byref(p), ctypes.byref(n)))
        return p.value, n.value

    def thresholds_mark_complete(self):
        self._ck(self.lib.ofg_thresholds_mark_complete(self.h))

    def sync(self):
        """Wait for the enqueued work and read its results back (option "async"); raises what the build would have."""
        self._ck(self.lib.ofg_sync(self.h))

    def stat(self, name):
        v = ctypes.c_int64(0)
        self._ck(self.lib.ofg_get_stat(self.h, name.encode(), ctypes.byref(v)))
        return v.value

    def dendro_matrices(self, og_off, species, iseq, complete=False):
        """DendroBLAST matrices of orthogroups on the raw scores (ofg_dendro_matrices; orthologues.py:264-291, 396-410).
        Returns (flat float64 array of sum n^2 values, max_og per orthogroup)."""
        og_off = np.ascontiguousarray(og_off, np.int64)
        species = np.ascontiguousarray(species, np.int32)
        iseq = np.ascontiguousarray(iseq, np.int32)
        n_ogs = og_off.size - 1
        sizes = np.diff(og_off)
        out = np.empty(int((sizes * sizes).sum()), np.float64)
        mx = np.empty(max(n_ogs, 1), np.float64)
        self._ck(self.lib.ofg_dendro_matrices(self.h, n_ogs, _vp(og_off), _vp(species), _vp(iseq), int(bool(complete)), _vp(out), out.size, _vp(mx)))
        return out, mx[:n_ogs]

    def wait_for(self, other):
        """Order this context's stream after everything enqueued on `other`'s (no host block)."""
        self._ck(self.lib.ofg_wait_for(self.h, other.h))

    def device_alloc(self, nbytes):
        p = ctypes.c_void_p()
        self._ck(self.lib.ofg_device_alloc(self.h, int(nbytes), ctypes.byref(p)))
        return p.value

    def device_free(self, ptr):
        self._ck(self.lib.ofg_device_free(self.h, ctypes.c_void_p(ptr)))

    def ipc_handle(self, ptr):
        buf = ctypes.create_string_buffer(64)
        self._ck(self.lib.ofg_ipc_get_handle(self.h, ctypes.c_void_p(ptr), buf))
        return buf.raw

    def ipc_open(self, handle):
        p = ctypes.c_void_p()
        self._ck(self.lib.ofg_ipc_open_handle(self.h, ctypes.c_char_p(handle), ctypes.byref(p)))
        return p.value

    def ipc_close(self, ptr):
        self._ck(self.lib.ofg_ipc_close_handle(self.h, ctypes.c_void_p(ptr)))

    def export_flags_peer(self, which, species_rank, my_rank, inbox_ptrs, seg_bytes):
        """Design B exchange without host round trips: this context's best-hit (which=0) / connect (which=1) lists go
        straight into segment `my_rank` of every destination's inbox (device pointers valid on this GPU)."""
        sr = np.ascontiguousarray(species_rank, dtype=np.int32)
        ptrs = (ctypes.c_void_p * len(inbox_ptrs))(*[ctypes.c_void_p(x) for x in inbox_ptrs])
        self._ck(self.lib.ofg_export_flags_peer(self.h, int(which), len(inbox_ptrs), _vp(sr), int(my_rank), ptrs, int(seg_bytes)))

    def import_flags_peer(self, which, n_ranks, my_rank, inbox_ptr, seg_bytes):
        self._ck(self.lib.ofg_import_flags_peer(self.h, int(which), int(n_ranks), int(my_rank), ctypes.c_void_p(inbox_ptr), int(seg_bytes)))

    def export_flags(self, which, species_rank, my_rank, n_ranks=None):
        """Design B exchange: (device pointer, counts per destination rank) of the best-hit (which=0) or
        connect (which=1) index lists for species pairs whose column species another rank owns."""
        sr = np.ascontiguousarray(species_rank, dtype=np.int32)
        if n_ranks is None:
            n_ranks = int(sr.max()) + 1 if sr.size else 1
        counts = np.zeros(n_ranks, np.int64)
        ptr = ctypes.c_void_p()
        self._ck(self.lib.ofg_export_flags(self.h, int(which), n_ranks, _vp(sr), int(my_rank), _vp(counts), ctypes.byref(ptr)))
        return ptr.value, counts

    def import_flags(self, which, dev_ptr, n_entries):
        self._ck(self.lib.ofg_import_flags(self.h, int(which), ctypes.c_void_p(dev_ptr), int(n_entries)))

    def graph_csr(self):
        nr, nnz = ctypes.c_int64(0), ctypes.c_int64(0)
        self._ck(self.lib.ofg_graph_csr_size(self.h, ctypes.byref(nr), ctypes.byref(nnz)))
        rows = np.zeros(max(nr.value, 1), np.int64)
        indptr = np.zeros(nr.value + 1, np.int64)
        idx = np.zeros(max(nnz.value, 1), np.int32)
        dat = np.zeros(max(nnz.value, 1), np.float64)
        self._ck(self.lib.ofg_graph_csr_copy(self.h, _vp(rows), _vp(indptr), _vp(idx), _vp(dat)))
        return rows[:nr.value], indptr, idx[:nnz.value], dat[:nnz.value]

    def graph_text_size(self):
        n = ctypes.c_size_t(0)
        self._ck(self.lib.ofg_graph_text_size(self.h, ctypes.byref(n)))
        return n.value

    def graph_text_body(self, out=None):
        n = self.graph_text_size()
        buf = out if out is not None else np.empty(max(n, 1), np.uint8)
        assert buf.size >= n
        self._ck(self.lib.ofg_graph_text_copy(self.h, _vp(buf), n))
        return buf[:n]


class PipelinedGraphBuilder:
    """The whole job on one GPU with the host->device copy hidden behind the compute.

    The query species are cut into `n_parts` contiguous row partitions, each in its own context (own stream,
    SURVEY.md §8e design B: rows only).  All copies are queued up front, partition after partition (a
    partition's copies wait for the previous partition's, so the PCIe link serves them in order at full
    rate), and partition g is parsed and normalised while partitions g+1.. are still in flight.  The
    transposed best-hit and connect lists are then handed between the contexts on the device
    (sharding.exchange_lists_local) and the stitched graph text is the concatenation of the parts - byte
    for byte the single-context result (tests/test_gpu_parity.py)."""

    def __init__(self, device=0, n_parts=4):
        self.device = device
        self.n_parts_req = n_parts
        self.ctxs = []
        self.parts = []
        self.xchg = None

    def close(self):
        if self.xchg is not None:
            self.xchg.close()
            self.xchg = None
        for c in self.ctxs:
            c.close()
        self.ctxs = []

    def set_species(self, n_seqs, lengths, weights=None):
        from . import sharding
        S = len(n_seqs)
        self.S = S
        parts = sharding.contiguous_partitions(S, self.n_parts_req, weights)
        if [len(x) for x in parts] != [len(x) for x in self.parts] or not self.ctxs:
            self.close()
            self.ctxs = [GraphBuilder(self.device) for _ in parts]
        self.parts = parts
        self.owner = sharding.species_owner(parts, S)
        for c, owned in zip(self.ctxs, parts):
            c.set_species(n_seqs, lengths)
            c.set_partition(owned, False)
            c.set_option("async", 1)   # the contexts only enqueue: their streams overlap with each other and with the copies
        if self.xchg is not None:
            self.xchg.close()
        self.xchg = sharding.LocalExchange(self.ctxs, self.owner, n_seqs) if len(self.ctxs) > 1 else None

    def set_option(self, name, value):
        for c in self.ctxs:
            c.set_option(name, value)

    def load(self, host_buffer_of, gz=False):
        """host_buffer_of(p, j) -> (host pointer, nbytes) (pinned memory for an asynchronous copy) or a
        bytes-like object.  Queues every copy; returns immediately.  gz: the buffers hold the .gz files as they are on
        disk; only the compressed bytes are copied and every context inflates its files on the device."""
        import torch
        prev = None
        for c, owned in zip(self.ctxs, self.parts):
            c.clear_hits()
            st = torch.cuda.ExternalStream(c.lib.ofg_stream(c.h), device=torch.device("cuda", self.device))
            if prev is not None:
                st.wait_event(prev)
            for p in owned:
                for j in range(self.S):
                    d = host_buffer_of(p, j)
                    if gz:
                        c.add_hits_gz(p, j, d)
                    elif isinstance(d, tuple):
                        c.add_hits_ptr(p, j, d[0], d[1])
                    else:
                        c.add_hits(p, j, d)
            prev = torch.cuda.Event()
            prev.record(st)

    def build(self):
        """Enqueues the whole job on the contexts' streams (no host synchronisation in between: the transposed lists go
        from context to context on the device) and then reads every context's results back."""
        for attempt in range(8):
            for c in self.ctxs:
                c.build(_lib.STAGE_NORM)
            if self.xchg is not None:
                self.xchg.exchange(0)
            for c in self.ctxs:
                c.build(_lib.STAGE_CONNECT)
            if self.xchg is not None:
                self.xchg.exchange(1)
            for c in self.ctxs:
                c.build(_lib.STAGE_GRAPH)
            try:
                for c in self.ctxs:
                    c.sync()
                return
            except OfgError as e:
                if e.code == _lib.ERR_CAPACITY and "inbox" in e.message and self.xchg is not None and attempt < 7:
                    for c in self.ctxs:     # drain the other contexts, then repeat the job with larger inboxes
                        try:
                            c.sync()
                        except OfgError:
                            pass
                    self.xchg.grow()
                    continue
                raise

    def counts(self):
        out = dict(lines=0, records=0, nnz=0, edges=0)
        for c in self.ctxs:
            for k, v in c.counts().items():
                out[k] += v
        return out

    def graph_text_size(self):
        return sum(c.graph_text_size() for c in self.ctxs)

    def graph_text_body(self, out=None):
        n = self.graph_text_size()
        buf = out if out is not None else np.empty(max(n, 1), np.uint8)
        assert buf.size >= n
        o = 0
        for c in self.ctxs:
            k = c.graph_text_size()
            if k:
                c.graph_text_body(buf[o:o + k])
            o += k
        return buf[:n]


def graph_header(n_seqs_total):
    """gathering.py:279-280."""
    return ("(mclheader\nmcltype matrix\ndimensions %dx%d\n)\n" % (n_seqs_total, n_seqs_total)).encode() + \
        b"\n(mclmatrix\nbegin\n\n"


GRAPH_FOOTER = b")\n"  # gathering.py:48, written after the last species


# --------------------------------------------------------------------------------------------------
# Reference-facing interface
# --------------------------------------------------------------------------------------------------
def GetSpeciesToUse(speciesIDsFN):
    """util.py:178-192 (host; not part of the device path): (speciesToUse, nSpAll, speciesToUse_names)."""
    use, names, skipped = [], [], 0
    with open(speciesIDsFN) as f:
        for line in f:
            line = line.rstrip()
            if not line:
                continue
            if line.startswith("#"):
                skipped += 1
            else:
                iSp, spName = line.split(": ")
                use.append(int(iSp))
                names.append(spName)
    return use, len(use) + skipped, names


def GetSeqsInfo(inputDirectory_list, speciesToUse, nSpAll):
    """util.py:64-84."""
    starts, n, per = [0], 0, {}
    for k in range(nSpAll):
        for d in inputDirectory_list:
            fn = os.path.join(d, "Species%d.fa" % k)
            if os.path.exists(fn):
                break
        c = 0
        with open(fn) as f:
            for line in f:
                if len(line) > 1 and line[0] == ">":
                    c += 1
        per[k] = c
        if k in speciesToUse:
            n += c
            starts.append(n)
    return SequencesInfo(nSeqs=n, nSpecies=len(speciesToUse), speciesToUse=list(speciesToUse),
                         seqStartingIndices=starts[:-1], nSeqsPerSpecies=per)


def GetSequenceLengths(seqsInfo, fasta_dir_list):
    """gathering.py:444-464: one fp64 vector per used species."""
    out = []
    for k in seqsInfo.speciesToUse:
        L = np.zeros(seqsInfo.nSeqsPerSpecies[k])
        for d in fasta_dir_list:
            fn = os.path.join(d, "Species%d.fa" % k)
            if os.path.exists(fn):
                break
        cur, idx, first = 0, -1, True
        with open(fn) as f:
            for row in f:
                if len(row) > 1 and row[0] == ">":
                    if first:
                        first = False
                    else:
                        L[idx] = cur
                        cur = 0
                    idx = int(row[1:].split("_")[1])
                else:
                    cur += len(row.rstrip())
        L[idx] = cur
        out.append(L)
    return out


BGZF_CHUNK = 65280   # text bytes per member, as bgzip: every member stays below 64 KB compressed


def gzip_members(data, level=1, chunk=BGZF_CHUNK):
    """Writes `data` as a gzip file of independent members of `chunk` text bytes each, BGZF style: every member's header
    carries its compressed size in a "BC" extra field (readable by gzip.open / zcat / bgzip like any multi-member gzip
    file).  This is the layout the device-side inflate parallelises over: one warp per member."""
    import struct
    out = []
    mv = memoryview(data)
    for o in range(0, max(len(mv), 1), chunk):
        part = bytes(mv[o:o + chunk])
        co = zlib.compressobj(level, zlib.DEFLATED, -15)
        body = co.compress(part) + co.flush()
        total = 18 + len(body) + 8
        assert total <= 65536
        out.append(b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + struct.pack("<H", total - 1))
        out.append(body)
        out.append(struct.pack("<II", zlib.crc32(part) & 0xffffffff, len(part) & 0xffffffff))
    return b"".join(out)


def _gunzip(data):
    """All members of a gzip stream in one go.  zlib releases the GIL while it inflates, so the reader threads of
    read_hit_files run in parallel (gzip.GzipFile.read loops in Python)."""
    out = []
    while data:
        d = zlib.decompressobj(wbits=31)
        out.append(d.decompress(data))
        out.append(d.flush())
        if not d.eof:  # what gzip.open (blast_file_processor.py:63) raises on a truncated member
            e = EOFError("Compressed file ended before the end-of-stream marker was reached")
            e.partial = b"".join(out)
            raise e
        data = d.unused_data.lstrip(b"\0")
    return b"".join(out) if len(out) != 2 else out[0] + out[1]


class GzBytes(bytes):
    """The raw bytes of a .gz hit file (to be inflated on the device)."""


def gz_has_size_hints(data):
    """True if the first gzip member carries its compressed size in a "BC" extra field (bgzip / BGZF, gzip_members): such
    files split into members without inflating, so the device can inflate them with one warp per member.  A classic
    single-member file (what DIAMOND's --compress 1 writes) is one warp's work on the device - host zlib is faster there."""
    if len(data) < 18 or data[0] != 0x1f or data[1] != 0x8b or not (data[3] & 4):
        return False
    xlen = data[10] | (data[11] << 8)
    o, end = 12, min(12 + xlen, len(data))
    while o + 4 <= end:
        ln = data[o + 2] | (data[o + 3] << 8)
        if data[o] == 0x42 and data[o + 1] == 0x43 and ln == 2:
            return True
        o += 4 + ln
    return False


def csv_unquote(data):
    """The reference reads hit files through csv.reader(delimiter='\\t') with the default dialect (blast_file_processor.py:65-66): a
    field that starts with '"' is quoted - '""' inside it is one quote, tabs and line breaks inside it belong to the field.  No search
    program writes such files and the device parser refuses a quote (OFG_LINE_QUOTE), so a file that contains one is rewritten HERE,
    by the csv module itself, into the plain rows it yields; the device then parses those.  Tabs and line breaks inside a field become
    spaces (int() / float() strip them like any other white space, blast_file_processor.py:72-79), a quote that is left inside a field
    becomes an apostrophe (either makes an id or a score unreadable, :74-82; the other columns are not read); a row without fields is a blank line
    (skipped, :68-69).  Text mode of open() / gzip.open('rt') means universal newlines, reproduced first."""
    import csv
    import io
    text = bytes(data).decode("utf-8", errors="surrogateescape").replace("\r\n", "\n").replace("\r", "\n")
    rows = []
    for row in csv.reader(io.StringIO(text, newline=""), delimiter="\t"):
        if row:
            rows.append("\t".join(f.replace("\t", " ").replace("\n", " ").replace('"', "'") for f in row))
    return ("\n".join(rows) + "\n").encode("utf-8", errors="surrogateescape") if rows else b""


def _plain_rows(data):
    """Host-side bytes of a hit file as the device takes them: files with csv quoting are rewritten (csv_unquote)."""
    return csv_unquote(data) if b'"' in data else data


def _read_hit_file(blastDir_list, iSpecies, jSpecies, q_allow_empty, raw_gz=False):
    """File lookup of blast_file_processor.py:59-65 (first directory holding the file wins, .gz accepted).
    raw_gz: hand .gz files over compressed (GzBytes) instead of inflating them with zlib on the host; "auto": only the files
    whose members carry size hints (gz_has_size_hints), the others are inflated here, in the reader thread."""
    fn = None
    for d in blastDir_list:
        fn = os.path.join(d, "Blast%d_%d.txt" % (iSpecies, jSpecies))
        if os.path.exists(fn) or os.path.exists(fn + ".gz"):
            break
    if not os.path.exists(fn) and not os.path.exists(fn + ".gz"):
        if q_allow_empty:
            return None, fn
        raise FileNotFoundError(fn)
    if os.path.exists(fn + ".gz"):
        with open(fn + ".gz", "rb") as f:
            raw = f.read()
            if raw_gz is True or (raw_gz == "auto" and gz_has_size_hints(raw)):
                return GzBytes(raw), fn
            try:
                return _plain_rows(_gunzip(raw)), fn
            except (EOFError, zlib.error) as e:
                e.hit_file = (iSpecies, jSpecies, d)
                raise
    with open(fn, "rb") as f:
        return _plain_rows(f.read()), fn


def read_hit_files(blastDir_list, requests, q_allow_empty=False, n_readers=None, window=None, raw_gz=False):
    """Yields (request, data, filename) for every request (iSpecies, jSpecies) IN ORDER while a pool of threads reads
    and inflates the files ahead of the consumer - DIAMOND writes its hit files gzip-compressed (config.json:51), and
    inflating 16 files at once is what keeps the device fed.  At most `window` files are held in memory."""
    from concurrent.futures import ThreadPoolExecutor
    requests = list(requests)
    n_readers = n_readers or max(1, min(16, os.cpu_count() or 1))
    window = window or 2 * n_readers
    if n_readers <= 1 or len(requests) <= 1:
        for r in requests:
            data, fn = _read_hit_file(blastDir_list, r[0], r[1], q_allow_empty, raw_gz)
            yield r, data, fn
        return
    with ThreadPoolExecutor(max_workers=n_readers) as pool:
        pending = []
        nxt = 0
        while nxt < len(requests) or pending:
            while nxt < len(requests) and len(pending) < window:
                r = requests[nxt]
                pending.append((r, pool.submit(_read_hit_file, blastDir_list, r[0], r[1], q_allow_empty, raw_gz)))
                nxt += 1
            r, fut = pending.pop(0)
            data, fn = fut.result()
            yield r, data, fn


class WaterfallMethod:
    """Mirror of gathering.WaterfallMethod for the graph hot path, backed by libofg."""

    @staticmethod
    def BuildGraph(seqsInfo, blastDir_list, Lengths, qDoubleBlast=True, v2_scores=False, q_allow_empty=False, device=0,
                   builder=None, n_readers=None, homology=False, device_gunzip="auto", stage=None, keep_self_hits=False):
        """ProcessBlastHits + ConnectCognates for every species (gathering.py:177-196,251-259), up to
        the device-resident graph.  Returns the GraphBuilder holding it.
        homology=True: the `--c-homologs` graph of gathering_version (3, 2) instead (gathering.py:51-73, 520-521).
        stage=_lib.STAGE_RAW with keep_self_hits=True: only the raw matrices of GetBLAST6Scores(qExcludeSelfHits=False), what
        DendroBLAST reads (orthologues.py:272-274; orthofinder_b200/dendroblast.py).
        device_gunzip: "auto" - .gz files whose members carry size hints (bgzip / BGZF layout) are inflated on the device, one
        warp per member, classic single-member files by zlib in the reader threads (a single member is one warp's work on the
        device); True - every .gz file goes to the device; False - every .gz file is inflated on the host."""
        stage = _lib.STAGE_GRAPH if stage is None else stage
        gb = builder or GraphBuilder(device)
        sp = list(seqsInfo.speciesToUse)
        n_seqs = [seqsInfo.nSeqsPerSpecies[k] for k in sp]
        gb.set_species(n_seqs, Lengths)
        gb.set_option("allow_empty", int(q_allow_empty))
        gb.set_option("scores_v2", int(bool(v2_scores)))  # --scores-v2: gathering.py:157-174, :314-388
        gb.set_option("homology_graph", int(bool(homology)))
        gb.set_option("keep_self_hits", int(bool(keep_self_hits)))
        cells = []
        for p, kp in enumerate(sp):
            for j, kj in enumerate(sp):
                swap = (not qDoubleBlast) and kp > kj  # blast_file_processor.py:41-54
                cells.append((kj if swap else kp, kp if swap else kj, p, j, swap))
        try:
            for (i_open, j_open, p, j, swap), data, fn in read_hit_files(blastDir_list, cells, q_allow_empty, n_readers, raw_gz=device_gunzip):
                if data is None:
                    continue
                if isinstance(data, GzBytes):
                    # the compressed bytes go to the device as they are (ofg_add_hits_gz); layouts the device cannot split
                    # into members are inflated here instead
                    try:
                        gb.add_hits_gz(p, j, data, swap_cols=swap)
                        continue
                    except OfgError as e:
                        if e.code != _lib.ERR_UNSUPPORTED:
                            raise
                        try:
                            data = _plain_rows(_gunzip(data))
                        except (EOFError, zlib.error) as e2:
                            e2.hit_file = (i_open, j_open, os.path.dirname(fn) + "/")
                            raise
                gb.add_hits(p, j, data, swap_cols=swap)
        except (EOFError, zlib.error) as e:
            # a truncated / damaged .gz ends the reference's csv loop with an exception: same report (blast_file_processor.py:99-103)
            if hasattr(e, "hit_file"):
                i_sp, j_sp, d = e.hit_file
                lines = getattr(e, "partial", b"").split(b"\n")
                last = lines[-2] if len(lines) > 1 else b""   # the last complete row the csv reader had returned
                print("ERROR: Blast%d_%d.txt is corrupted" % (i_sp, j_sp))
                sys.stderr.write("Malformatted line in %sBlast%d_%d.txt\nOffending line was:\n" % (d, i_sp, j_sp))
                sys.stderr.write(last.decode(errors="replace") + "\n")
                print("ERROR: Error processing files Blast%d_*" % i_sp)
            raise
        try:
            gb.build(stage)
        except OfgError as e:
            if device_gunzip and e.code == _lib.ERR_PARSE and " gzip: " in e.message:
                # a .gz the device could not inflate (e.g. concatenated members without size hints, which look like one member
                # with a wrong length): let zlib decide - it either inflates the file or raises what gzip.open raises
                return WaterfallMethod.BuildGraph(seqsInfo, blastDir_list, Lengths, qDoubleBlast, v2_scores, q_allow_empty, device,
                                                  gb, n_readers, homology, device_gunzip=False, stage=stage,
                                                  keep_self_hits=keep_self_hits)
            WaterfallMethod._report(e, sp, blastDir_list)
            raise
        if stage < _lib.STAGE_NORM:
            return gb
        for p in range(len(sp)):
            for j in range(len(sp)):
                f = gb.pair_fit(p, j)
                if not f["normalised"]:
                    # gathering.py:146-155
                    if p == j and f["m"] == 0:
                        print("WARNING: THIS IS UNCOMMON, there are zero hits when searching the genes in species %d "
                              "against itself. Check the input proteome contains all the genes from that species and "
                              "check the search program is working (default is diamond)." % p)
                    else:
                        print("WARNING: Too few hits between species %d and species %d to normalise the scores, "
                              "these hits will be ignored" % (p, j))
        return gb

    @staticmethod
    def _report(e, sp, blastDir_list):
        """Reproduce the reference's messages (blast_file_processor.py:74-103, gathering.py:209)."""
        msg = e.message
        if e.code == _lib.ERR_PARSE and msg.startswith("pair=") and " gzip: " in msg:
            # a damaged .gz: the reference's reader raises inside gzip.open and reports the file (blast_file_processor.py:99-103)
            p, j = [int(x) for x in msg.split()[0][5:].split(",")]
            print("ERROR: Blast%d_%d.txt is corrupted" % (sp[p], sp[j]))
            sys.stderr.write("Malformatted line in %sBlast%d_%d.txt\nOffending line was:\n\n" % (blastDir_list[0], sp[p], sp[j]))
            print("ERROR: Error processing files Blast%d_*" % sp[p])
            return
        if e.code in (_lib.ERR_PARSE, _lib.ERR_INCONSISTENT, _lib.ERR_UNSUPPORTED) and msg.startswith("pair="):
            head, line = msg.split(" line=", 1)
            p, j = [int(x) for x in head.split()[0][5:].split(",")]
            kind = int(head.split("kind=")[1].split()[0])
            i_sp, j_sp = sp[p], sp[j]
            if e.code == _lib.ERR_INCONSISTENT:
                sys.stderr.write("\nERROR: Inconsistent input files.\n")
            elif kind == 1:
                sys.stderr.write("\nERROR: Query or hit sequence ID in BLAST results file was missing or incorrectly formatted.\n")
            elif kind in (2, 6):
                sys.stderr.write("\nERROR: 12th field in BLAST results file line should be the bit-score for the hit\n")
            print("ERROR: Blast%d_%d.txt is corrupted" % (i_sp, j_sp))
            sys.stderr.write("Malformatted line in %sBlast%d_%d.txt\nOffending line was:\n" % (blastDir_list[0], i_sp, j_sp))
            sys.stderr.write(line + "\n")
            print("ERROR: Error processing files Blast%d_*" % i_sp)

    @staticmethod
    def WriteGraphParallel(gb, seqsInfo, graphFN):
        """gathering.py:273-291: header, rows of every species, ')' — one file."""
        with open(graphFN, "wb") as f:
            f.write(graph_header(seqsInfo.nSeqs))
            f.write(gb.graph_text_body().tobytes())
            f.write(GRAPH_FOOTER)
        return graphFN


def GetGraphFilename(wd_current, i_unassigned=None, fileIdentifierString="OrthoFinder"):
    """files.py:324-327 FileHandler.GetGraphFilename."""
    identifier = fileIdentifierString + ("" if i_unassigned is None else "_unassigned_clade_%d" % i_unassigned)
    return wd_current + "%s_graph.txt" % identifier


def DoOrthogroupsGraph(workingDir, graphFN=None, qDoubleBlast=True, q_allow_empty=None, device=0, blastDir_list=None,
                       v2_scores=False, i_unassigned=None, wd_current=None, n_readers=None, device_gunzip="auto"):
    """The part of gathering.DoOrthogroups between :476 and :512 for a `-b <dir> -og` run: reads SpeciesIDs.txt /
    Species*.fa from workingDir and Blast*_*.txt[.gz] from blastDir_list (default [workingDir]; the first directory
    holding a file wins, files.py:313-322) and writes the graph file.

    i_unassigned: the clade-specific run of `--assign` (gathering.py:469-483, __main__.py:1083-1093): only the latest
    blast directory is searched, missing hit files are empty (q_allow_empty), the caller passes v2_scores=True and the
    graph goes to OrthoFinder_unassigned_clade_<i>_graph.txt (files.py:324-327)."""
    if not workingDir.endswith(os.sep):
        workingDir += os.sep
    q_unassigned = i_unassigned is not None
    sp, n_all, _ = GetSpeciesToUse(workingDir + "SpeciesIDs.txt")
    si = GetSeqsInfo([workingDir], sp, n_all)
    L = GetSequenceLengths(si, [workingDir])
    dirs = [d if d.endswith(os.sep) else d + os.sep for d in (blastDir_list or [workingDir])]
    if q_unassigned:
        dirs = dirs[:1]   # gathering.py:482-483: only the latest directory holds the unassigned-gene searches
    if q_allow_empty is None:
        q_allow_empty = q_unassigned   # gathering.py:492 passes q_unassigned as q_allow_empty
    gb = WaterfallMethod.BuildGraph(si, dirs, L, qDoubleBlast=qDoubleBlast, v2_scores=v2_scores, q_allow_empty=q_allow_empty,
                                    device=device, n_readers=n_readers, device_gunzip=device_gunzip)
    if graphFN is None:
        graphFN = GetGraphFilename(wd_current or workingDir, i_unassigned)
    WaterfallMethod.WriteGraphParallel(gb, si, graphFN)
    return graphFN, gb, si


def BuildGraphDistributed(seqsInfo, blastDir_list, Lengths, graphFN, qDoubleBlast=True, v2_scores=False, q_allow_empty=False,
                          device=None, n_readers=None):
    """The graph stage of DoOrthogroups (gathering.py:480-512) over the GPUs of one node, one process per GPU
    (torch.distributed must be initialised, backend nccl).  What the reference does with a queue of query species and
    pickles on disk (gathering.py:484-509, matrices.py:37-57):

      * the query species are dealt to the ranks greedily by the size of their hit files (partition_rows);
      * every rank reads, inflates and uploads only the Blast{p}_{j} files of its own query species and runs
        ProcessBlastHits for them (design B of SURVEY.md 8e: rows only);
      * the two reads of the transposed direction - BH_jp at gathering.py:254, connect_jp at :30 - are index lists that
        the export kernel stores straight into the owning rank's inbox over NVLink (sharding.PeerExchange);
      * every rank writes the part files <graphFN>_<iSpec> of its species, exactly the files WriteGraph_perSpecies writes
        (gathering.py:26), and rank 0 assembles header + parts in species order + ")" like WriteGraphParallel (:279-290).

    Every rank returns graphFN once the file is complete."""
    import torch
    import torch.distributed as dist
    from . import sharding
    rank, world = dist.get_rank(), dist.get_world_size()
    if device is None:
        device = torch.cuda.current_device()
    sp = list(seqsInfo.speciesToUse)
    S = len(sp)
    n_seqs = [seqsInfo.nSeqsPerSpecies[k] for k in sp]

    def file_of(kp, kj):
        for d in blastDir_list:
            fn = os.path.join(d, "Blast%d_%d.txt" % (kp, kj))
            if os.path.exists(fn):
                return fn
            if os.path.exists(fn + ".gz"):
                return fn + ".gz"
        return None

    # weights: bytes on disk of every query species' files (a .gz counts a few times its size: it inflates)
    weights = np.zeros(S)
    for p, kp in enumerate(sp):
        for j, kj in enumerate(sp):
            swap = (not qDoubleBlast) and kp > kj
            fn = file_of(kj, kp) if swap else file_of(kp, kj)
            if fn is not None:
                weights[p] += os.path.getsize(fn) * (4 if fn.endswith(".gz") else 1)
    parts = sharding.partition_rows(S, world, weights)
    owner = sharding.species_owner(parts, S)
    owned = parts[rank]
    gb = GraphBuilder(device)
    try:
        gb.set_species(n_seqs, Lengths)
        gb.set_partition(owned, False)
        gb.set_option("allow_empty", int(q_allow_empty))
        gb.set_option("scores_v2", int(bool(v2_scores)))
        gb.set_option("async", 1)
        cells = []
        for p in owned:
            kp = sp[p]
            for j, kj in enumerate(sp):
                swap = (not qDoubleBlast) and kp > kj
                cells.append((kj if swap else kp, kp if swap else kj, p, j, swap))
        for (_, _, p, j, swap), data, fn in read_hit_files(blastDir_list, cells, q_allow_empty, n_readers):
            if data is not None:
                gb.add_hits(p, j, data, swap_cols=swap)
        px = sharding.PeerExchange(gb, owner, rank, world, device, n_seqs) if world > 1 else None
        err = None
        try:
            for attempt in range(6):
                gb.build(_lib.STAGE_NORM)
                if px is not None:
                    px.exchange(0)
                gb.build(_lib.STAGE_CONNECT)
                if px is not None:
                    px.exchange(1)
                gb.build(_lib.STAGE_GRAPH)
                status = 0
                try:
                    gb.sync()
                except OfgError as e:
                    err = e
                    status = 2 if (e.code == _lib.ERR_CAPACITY and "inbox" in e.message) else 1
                # every rank must take the same decision: an inbox that was too small anywhere repeats the job everywhere
                st = torch.tensor([status], dtype=torch.int32, device=torch.device("cuda", device))
                if world > 1:
                    dist.all_reduce(st, op=dist.ReduceOp.MAX)
                worst = int(st.item())
                if worst == 0:
                    err = None
                    break
                if worst == 2 and px is not None and attempt < 5:
                    px.grow()
                    err = None
                    continue
                break
            if err is not None:
                WaterfallMethod._report(err, sp, blastDir_list)
                raise err
            if worst != 0:
                raise OfgError(_lib.ERR_ARG, "another rank failed while building the graph")
        finally:
            if px is not None:
                px.close()
        for p in owned:
            for j in range(S):
                f = gb.pair_fit(p, j)
                if not f["normalised"]:
                    if p == j and f["m"] == 0:
                        print("WARNING: THIS IS UNCOMMON, there are zero hits when searching the genes in species %d "
                              "against itself. Check the input proteome contains all the genes from that species and "
                              "check the search program is working (default is diamond)." % p)
                    else:
                        print("WARNING: Too few hits between species %d and species %d to normalise the scores, "
                              "these hits will be ignored" % (p, j))
        # part files per owned species (rows of a species are one contiguous run of lines of the rank's block)
        body = gb.graph_text_body()
        ends = np.flatnonzero(body == 10)
        line0 = 0
        for p in owned:
            lo = 0 if line0 == 0 else int(ends[line0 - 1]) + 1
            hi = int(ends[line0 + n_seqs[p] - 1]) + 1 if n_seqs[p] else lo
            with open(graphFN + "_%d" % p, "wb") as f:
                f.write(body[lo:hi].tobytes())
            line0 += n_seqs[p]
    finally:
        gb.close()
    if world > 1:
        dist.barrier()
    if rank == 0:
        with open(graphFN, "wb") as out:
            out.write(graph_header(seqsInfo.nSeqs))
            for p in range(S):
                with open(graphFN + "_%d" % p, "rb") as part:
                    while True:
                        chunk = part.read(1 << 24)
                        if not chunk:
                            break
                        out.write(chunk)
                os.remove(graphFN + "_%d" % p)
            out.write(GRAPH_FOOTER)
    if world > 1:
        dist.barrier()
    return graphFN
