"""Sharding of the pair grid by query species.

Rank r owns a set of query species p ("rows" of the S x S pair grid), exactly the unit the reference
parallelises over (gathering.py:484-485).  Everything that is reduced per gene (row maxima, best hit in
another species, the "most distant" threshold) stays local to the owner.  The reference reads the
transposed direction from pickles in two places: BH_jp (gathering.py:254) and connect_jp (:30).

Design B (default): a rank ingests only its rows; the two transposed matrices travel as sparse index lists
(8 bytes per flagged entry).  `PeerExchange` (one process per GPU) and `LocalExchange` (row partitions in
several contexts of one GPU) let the export kernel write the lists straight into the destinations' inboxes -
peer memory over NVLink, mapped with CUDA IPC - with all offsets computed on the device; the only collective
is a 4-byte all-reduce that orders "every exporter is done" before the imports.  `exchange_lists` is the
older variant that ships the same lists with two NCCL all-to-all calls (host-sized).

Design A (north-star wording): a rank also ingests the column pairs Blast{j}_{p}; their fit, normalisation
and best hits depend only on that file and the two length vectors, so they are recomputed locally and
bit-identically, and the only data exchanged is the per-gene threshold vector (8 bytes per gene): one
all-reduce(MIN) over a vector that holds +inf for genes a rank does not own (`exchange_thresholds`).
"""
import numpy as np


def partition_rows(n_species, world, weights=None):
    """Greedy longest-first assignment of query species to ranks, balanced by `weights` (e.g. hit-file
    bytes per query species; default: equal).  Returns a list of sorted species lists, one per rank
    (ranks beyond the number of species own nothing)."""
    w = np.ones(n_species) if weights is None else np.asarray(weights, dtype=np.float64)
    load = [0.0] * world
    owned = [[] for _ in range(world)]
    for p in sorted(range(n_species), key=lambda k: (-w[k], k)):
        r = min(range(world), key=lambda k: (load[k], k))
        owned[r].append(p)
        load[r] += w[p]
    return [sorted(o) for o in owned]


def contiguous_partitions(n_species, n_parts, weights=None):
    """Split species 0..S-1 into n_parts contiguous ranges of roughly equal weight (used by the single-GPU
    copy/compute pipeline, where the stitched graph is the plain concatenation of the parts)."""
    n_parts = max(1, min(n_parts, n_species))
    w = np.ones(n_species) if weights is None else np.asarray(weights, dtype=np.float64)
    cum = np.concatenate(([0.0], np.cumsum(w)))
    cuts = [0]
    for g in range(1, n_parts):
        k = int(np.searchsorted(cum, cum[-1] * g / n_parts, side="left"))
        k = min(max(k, cuts[-1] + 1), n_species - (n_parts - g))
        cuts.append(k)
    cuts.append(n_species)
    return [list(range(cuts[g], cuts[g + 1])) for g in range(n_parts)]


def exchange_lists_local(ctxs, which, owner):
    """Design B exchange between row partitions that live in separate contexts of ONE process / GPU: every
    context exports its best-hit (which=0) or connect (which=1) lists grouped by destination and the
    destinations import them straight from the exporter's device buffer (no collective, no host copy)."""
    outbox = []
    for r, c in enumerate(ctxs):
        ptr, counts = c.export_flags(which, owner, r, len(ctxs))   # synchronises c's stream: the lists are complete
        outbox.append((ptr, counts))
    total = 0
    for src, (ptr, counts) in enumerate(outbox):
        off = 0
        for dst in range(len(counts)):
            n = int(counts[dst])
            if n and dst != src and dst < len(ctxs):
                ctxs[dst].import_flags(which, ptr + off * 8, n)
            off += n
            total += n
    return total


def _inbox_segment_bytes(owner, n_seqs, n_ranks, factor=4):
    """Initial size of one inbox segment: `factor` entries per matrix row of the largest (source, destination) block
    of species pairs; best-hit lists hold ~1 entry per row, connect lists a few (grown on demand)."""
    n_seqs = np.asarray(n_seqs, dtype=np.int64)
    owner = np.asarray(owner)
    rows = 0
    for src in range(n_ranks):
        g_src = int(n_seqs[owner == src].sum())
        for dst in range(n_ranks):
            if dst != src:
                rows = max(rows, g_src * int((owner == dst).sum()))
    return int(8 + 8 * (factor * rows + 4096))


class LocalExchange:
    """Design B exchange between row partitions held by several contexts of ONE process / GPU: every context exports
    straight into the other contexts' inboxes and the imports are ordered by stream events - nothing touches the host."""

    def __init__(self, ctxs, owner, n_seqs):
        self.ctxs, self.owner, self.n = ctxs, np.asarray(owner, dtype=np.int32), len(ctxs)
        self.seg = _inbox_segment_bytes(owner, n_seqs, self.n)
        self.inbox = None
        self._alloc()

    def _alloc(self):
        self.close()
        # two inboxes per context (best-hit lists, connect lists): a context may still be importing one while the next
        # export already arrives
        self.inbox = [[c.device_alloc(self.seg * self.n) for c in self.ctxs] for _ in range(2)]

    def grow(self):
        self.seg = 8 + 4 * (self.seg - 8)
        self._alloc()

    def close(self):
        if self.inbox:
            for boxes in self.inbox:
                for c, p in zip(self.ctxs, boxes):
                    if getattr(c, "h", None):
                        c.device_free(p)
        self.inbox = None

    def exchange(self, which):
        boxes = self.inbox[which]
        for r, c in enumerate(self.ctxs):
            c.export_flags_peer(which, self.owner, r, boxes, self.seg)
        for r, c in enumerate(self.ctxs):
            for src, other in enumerate(self.ctxs):
                if src != r:
                    c.wait_for(other)
            c.import_flags_peer(which, self.n, r, boxes[r], self.seg)


class PeerExchange:
    """Design B exchange between the ranks of one node (one process per GPU): every rank maps the other ranks' inboxes
    with CUDA IPC, the export kernel stores the lists into them over NVLink, and a 4-byte NCCL all-reduce on the
    builder's stream is the only collective (it orders "all exporters are done" before the imports)."""

    def __init__(self, gb, species_rank, rank, world, device_index, n_seqs):
        import torch
        import torch.distributed as dist
        self.gb, self.rank, self.world = gb, rank, world
        self.owner = np.asarray(species_rank, dtype=np.int32)
        self.dev = torch.device("cuda", device_index)
        self.stream = torch.cuda.ExternalStream(gb.lib.ofg_stream(gb.h), device=self.dev)
        self.token = torch.zeros(1, dtype=torch.int32, device=self.dev)
        self.seg = _inbox_segment_bytes(species_rank, n_seqs, world)
        self.mine, self.ptrs = None, None
        self._alloc()

    def _alloc(self):
        import torch.distributed as dist
        self._release()
        self.mine = [self.gb.device_alloc(self.seg * self.world) for _ in range(2)]
        handles = [None] * self.world
        dist.all_gather_object(handles, [self.gb.ipc_handle(p) for p in self.mine])
        self.ptrs = []
        for which in range(2):
            self.ptrs.append([self.mine[which] if r == self.rank else self.gb.ipc_open(handles[r][which])
                              for r in range(self.world)])

    def _release(self):
        import torch
        import torch.distributed as dist
        if self.ptrs:
            torch.cuda.synchronize()
            dist.barrier()
            for which in range(2):
                for r, p in enumerate(self.ptrs[which]):
                    if r != self.rank:
                        self.gb.ipc_close(p)
            dist.barrier()
            for p in self.mine:
                self.gb.device_free(p)
        self.mine, self.ptrs = None, None

    def grow(self):
        self.seg = 8 + 4 * (self.seg - 8)
        self._alloc()

    def close(self):
        self._release()

    def exchange(self, which):
        import torch
        import torch.distributed as dist
        self.gb.export_flags_peer(which, self.owner, self.rank, self.ptrs[which], self.seg)
        with torch.cuda.stream(self.stream):
            dist.all_reduce(self.token)   # stream-ordered: starts after this rank's export, ends after every rank's
        self.gb.import_flags_peer(which, self.world, self.rank, self.mine[which], self.seg)


def rank_pairs(owned, n_species):
    """(row pairs, column pairs) a rank ingests: (p, j) for owned p and every j, plus (j, p) for j it does
    not own."""
    own = set(owned)
    rows = [(p, j) for p in owned for j in range(n_species)]
    cols = [(j, p) for p in owned for j in range(n_species) if j not in own]
    return rows, cols


def merge_thresholds_host(parts):
    """Element-wise minimum of per-rank threshold vectors (+inf where a rank owns nothing): what the
    all-reduce computes."""
    out = np.array(parts[0], dtype=np.float64, copy=True)
    for p in parts[1:]:
        np.minimum(out, p, out=out)
    return out


class _DeviceVector:
    """Zero-copy view of a device vector for torch.as_tensor (CUDA array interface)."""

    def __init__(self, ptr, n, typestr="<f8"):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 3}


def species_owner(owned_per_rank, n_species):
    """owner rank of every species from partition_rows' output."""
    out = np.full(n_species, -1, np.int32)
    for r, o in enumerate(owned_per_rank):
        out[o] = r
    return out


def exchange_lists(gb, which, species_rank, rank, world, device_index):
    """Design B (no column pairs): all-to-all of the index lists the reference reads from the transposed
    pickles — best hits (which=0, gathering.py:254) or connect (which=1, gathering.py:30).  Two NCCL calls:
    the per-rank counts, then the variable-size lists (one int64 = two packed int32 gene ids per entry)."""
    import torch
    import torch.distributed as dist
    dev = torch.device("cuda", device_index)
    ptr, counts = gb.export_flags(which, species_rank, rank, world)
    send_counts = torch.from_numpy(np.ascontiguousarray(counts[:world], dtype=np.int64)).to(dev)
    recv_counts = torch.empty(world, dtype=torch.int64, device=dev)
    dist.all_to_all_single(recv_counts, send_counts)
    rc = [int(x) for x in recv_counts.cpu()]
    sc = [int(x) for x in counts[:world]]
    n_send = sum(sc)
    send = (torch.as_tensor(_DeviceVector(ptr, n_send, "<i8"), device=dev) if n_send
            else torch.empty(0, dtype=torch.int64, device=dev))
    recv = torch.empty(sum(rc), dtype=torch.int64, device=dev)
    dist.all_to_all_single(recv, send, output_split_sizes=rc, input_split_sizes=sc)
    torch.cuda.synchronize()
    gb.import_flags(which, recv.data_ptr(), recv.numel())
    return n_send, recv.numel()


def exchange_thresholds(gb, device_index):
    """In-place NCCL all-reduce(MIN) of the library's threshold vector (one collective over NVLink)."""
    import torch
    import torch.distributed as dist
    ptr, n = gb.thresholds_dev()
    t = torch.as_tensor(_DeviceVector(ptr, n), device=torch.device("cuda", device_index))
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    torch.cuda.synchronize()
    gb.thresholds_mark_complete()
