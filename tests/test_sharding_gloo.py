"""CPU, world_size 2 over gloo: the sharding protocol of orthofinder_b200/sharding.py (design A).  The
per-rank compute is done by the oracle here (no GPU in this container); what is under test is the
partition, the set of files each rank ingests, and that ONE all-reduce(MIN) of per-gene thresholds is
all the communication needed for the stitched graph to equal the single-process graph."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import waterfall_oracle as wo
from orthofinder_b200 import sharding
from tests.helpers import Golden


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _rank_graph(g, owned, S, most_all):
    """Rows of the final graph for the owned species, using only row pairs + column pairs + thresholds."""
    rows_p, cols_p = sharding.rank_pairs(owned, S)
    have = set(rows_p) | set(cols_p)
    B = [[None] * S for _ in range(S)]
    for (p, j) in have:
        raw = wo.get_blast6_scores(g.hit_text(p, j), g.n_seqs[p], g.n_seqs[j], p == j)
        B[p][j], _ = wo.normalise_scores(raw, g.lengths[p], g.lengths[j])
    return B, have


def _worker(rank, world, port, name, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = Golden(name)
    S = len(g.n_seqs)
    sizes = [sum(len(g.hit_text(p, j)) for j in range(S)) for p in range(S)]
    owned = sharding.partition_rows(S, world, sizes)[rank]
    B, have = _rank_graph(g, owned, S, None)
    starts = np.concatenate(([0], np.cumsum(g.n_seqs)))
    # best hits: rows of owned species need every j; column pairs (j, p), j != p, need only their own row maxima
    BH = [[None] * S for _ in range(S)]
    for p in owned:
        flags = wo.get_bh(B[p], p)
        for j in range(S):
            BH[p][j] = flags[j]
    for (j, p) in have:
        if BH[j][p] is None:
            indptr, cols, vals = B[j][p]
            m, _ = wo._row_max(indptr, vals, -1.0)
            rows = np.repeat(np.arange(g.n_seqs[j]), np.diff(indptr))
            BH[j][p] = vals > (m[rows] - 1e-3)
    # thresholds of owned species; +inf elsewhere; ONE collective
    most = np.full(int(starts[-1]), np.inf)
    for p in owned:
        # connect_cognates only reads B[p][k], BH[p][k], B[k][p], BH[k][p]
        _, md = wo.connect_cognates(B, BH, p, g.n_seqs)
        most[starts[p]:starts[p + 1]] = md
    t = torch.from_numpy(most)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    most = t.numpy()
    assert np.all(np.isfinite(most))
    # connect flags for every pair this rank holds, then its rows of W
    conn = [[None] * S for _ in range(S)]
    for (p, j) in have:
        indptr, cols, vals = B[p][j]
        rows = np.repeat(np.arange(g.n_seqs[p]), np.diff(indptr))
        c = vals >= most[starts[p]:starts[p + 1]][rows]
        if p == j:
            c &= rows != cols
        conn[p][j] = c
    R, C, V = [], [], []
    for p in owned:
        for j in range(S):
            indptr, cols, vals = B[p][j]
            rows, keys = wo._entry_keys(indptr, cols, g.n_seqs[j])
            i2, c2, _ = B[j][p]
            rows2, _ = wo._entry_keys(i2, c2, g.n_seqs[p])
            tk = c2[conn[j][p]].astype(np.int64) * g.n_seqs[j] + rows2[conn[j][p]]
            cnt = conn[p][j].astype(np.float64) + np.isin(keys, tk)
            keep = cnt != 0
            R.append(rows[keep] + starts[p]); C.append(cols[keep].astype(np.int64) + starts[j]); V.append((cnt * vals)[keep])
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), R=np.concatenate(R), C=np.concatenate(C), V=np.concatenate(V),
             owned=np.array(owned), most=most)
    dist.destroy_process_group()


@pytest.mark.parametrize("name", ["synth_C5_small", "ref_AddTwoSpecies"])
def test_two_ranks_one_allreduce_reproduce_the_graph(tmp_path, name):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, name, str(tmp_path)), nprocs=world, join=True)
    g = Golden(name)
    ref = wo.waterfall(g.n_seqs, g.lengths, g.hit_text)
    parts = [np.load(os.path.join(str(tmp_path), "rank%d.npz" % r)) for r in range(world)]
    owned = sorted(int(x) for p in parts for x in p["owned"])
    assert owned == list(range(len(g.n_seqs)))
    np.testing.assert_array_equal(parts[0]["most"], np.concatenate(ref["most_distant"]))
    R = np.concatenate([p["R"] for p in parts]); C = np.concatenate([p["C"] for p in parts]); V = np.concatenate([p["V"] for p in parts])
    order = np.lexsort((C, R))
    rows_ref = np.repeat(np.arange(len(ref["indptr"]) - 1), np.diff(ref["indptr"]))
    np.testing.assert_array_equal(R[order], rows_ref)
    np.testing.assert_array_equal(C[order], ref["indices"])
    np.testing.assert_array_equal(V[order], ref["data"])


def test_partition_is_balanced_and_complete():
    w = [5, 1, 1, 1, 4, 4, 2, 2]
    parts = sharding.partition_rows(8, 3, w)
    assert sorted(p for o in parts for p in o) == list(range(8))
    loads = [sum(w[p] for p in o) for o in parts]
    assert max(loads) - min(loads) <= 2
    rows, cols = sharding.rank_pairs(parts[0], 8)
    assert len(rows) == len(parts[0]) * 8 and len(cols) == len(parts[0]) * (8 - len(parts[0]))


def test_contiguous_partitions_cover_all_species_in_order():
    """sharding.contiguous_partitions (row partitions of the single-GPU copy/compute pipeline): contiguous, ordered,
    non-empty, balanced by weight."""
    from orthofinder_b200 import sharding
    for S in (1, 2, 3, 5, 16, 23, 64):
        for G in (1, 2, 4, 7, 100):
            parts = sharding.contiguous_partitions(S, G)
            assert len(parts) == min(S, G) and all(parts)
            assert [p for part in parts for p in part] == list(range(S))
            assert max(len(x) for x in parts) - min(len(x) for x in parts) <= 1
    parts = sharding.contiguous_partitions(6, 3, [100, 1, 1, 1, 1, 100])
    assert parts[0] == [0] and parts[-1][-1] == 5 and [p for part in parts for p in part] == list(range(6))
    owner = sharding.species_owner(parts, 6)
    assert list(owner) == [g for g, part in enumerate(parts) for _ in part]
