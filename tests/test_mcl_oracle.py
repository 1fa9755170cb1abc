"""MCL step: the CPU oracle (oracle/mcl_oracle.py) against what the reference's bundled mcl binary produced
(tests/golden/mcl_*.npz, tools/make_golden_mcl.py) - every iterand bit for bit, the clusters file body byte for byte - and,
where the reference is present, against the binary itself (marker `reference`)."""
import numpy as np
import pytest

from oracle import mcl_oracle as mo, ref_harness as rh
from tests.helpers import MCL_GOLDENS, MclGolden, matrix_digest, mcl_golden_graph, mcl_synthetic_graph


def _cases():
    for name in MCL_GOLDENS:
        for I in ((1.5,) if name == "synth_family" else (1.2, 1.5, 3.0)):   # the 1 250-gene family costs ~0.5 s per iteration
            yield name, I


@pytest.mark.parametrize("name,inflation", list(_cases()))
def test_oracle_reproduces_the_binary(name, inflation):
    g = MclGolden(name)
    n, cp, ri, v = mo.parse_graph_text(mcl_golden_graph(name))
    sha, nnz = g.digests(inflation)
    pr = mo.Process(n, cp, ri, v)
    assert matrix_digest(*pr.matrix(0)) == sha[0]                       # start matrix: loops, column sums
    k = 0
    while True:
        chaos = pr.step(inflation)
        k += 1
        m = pr.matrix(0)
        assert len(m[1]) == nnz[k], (k, len(m[1]), nnz[k])
        assert matrix_digest(*m) == sha[k], "iterand %d differs" % k
        if chaos < mo.DEFAULTS["chaos_limit"]:
            break
    assert k == g.iterations(inflation)
    nc, label, ov = pr.interpret(0)
    assert mo.clusters_text_body(label, n) == g.clusters_body(inflation)


needs_reference = [pytest.mark.reference, pytest.mark.skipif(not rh.reference_available(), reason="live reference not present (GPU box)")]


@pytest.mark.parametrize("P,S,R,pct", [(200, 10, 15, 90), (100, 8, 20, 95), (50, 20, 5, 90), (20, 3, 8, 99), (10000, 4, 4, 50),
                                       (40, 6, 12, 97), (500, 12, 9, 93)])
@pytest.mark.reference
@pytest.mark.skipif(not rh.reference_available(), reason="live reference not present (GPU box)")
def test_pruning_rule_against_the_binary(tmp_path, P, S, R, pct):
    """Small prune / select / recover numbers make every branch of the rule fire on a fixture-sized graph; compared on the
    binary's own first expansion (-write-expanded) and on its clusters."""
    graph = mcl_golden_graph("ref_AddTwoSpecies")
    ref = rh.run_mcl_binary(graph, 1.2, str(tmp_path), dump_iterands=True,
                            extra=("-P", str(P), "-S", str(S), "-R", str(R), "-pct", str(pct), "-write-expanded", "expanded.bin"))
    n, cp, ri, v = mo.parse_graph_text(graph)
    kw = dict(P=P, S=S, R=R, pct=pct / 100.0)
    pr = mo.Process(n, cp, ri, v)
    pr.step(1.2, **kw)
    assert matrix_digest(*pr.matrix(1)) == matrix_digest(*ref["expanded"])
    for k in range(1, len(ref["iterands"])):
        if k > 1:
            pr.step(1.2, **kw)
        assert matrix_digest(*pr.matrix(0)) == matrix_digest(*ref["iterands"][k]), k
    nc, label, ov = pr.interpret(0)
    assert mo.clusters_text_body(label, n) == mo.clusters_body_of_file(ref["clusters_text"])


@pytest.mark.parametrize("name", ["ref_FromTrees", "synth_ties"])
@pytest.mark.reference
@pytest.mark.skipif(not rh.reference_available(), reason="live reference not present (GPU box)")
def test_goldens_are_what_the_binary_writes_today(tmp_path, name):
    g = MclGolden(name)
    ref = rh.run_mcl_binary(mcl_golden_graph(name), 1.2, str(tmp_path), dump_iterands=True, threads=4)   # -te does not change results
    assert mo.clusters_body_of_file(ref["clusters_text"]) == g.clusters_body(1.2)
    assert [matrix_digest(*m) for m in ref["iterands"]] == g.digests(1.2)[0]


def test_graph_text_reader_drops_what_mcl_drops():
    n, cp, ri, v = mo.parse_graph_text(mcl_synthetic_graph("ties"))
    pr = mo.Process(n, cp, ri, v)
    scp, sri, sv = pr.matrix(0)
    assert np.all(sv > 0) and np.all(np.diff(scp) >= 1)                 # 0.000 entries gone, every column has its loop
    sums = np.add.reduceat(sv.astype(np.float64), scp[:-1])
    assert np.allclose(sums, 1.0, atol=1e-6)


@pytest.mark.parametrize("name,inflation", [("ref_AddTwoSpecies", 1.2), ("synth_ties", 1.5)])
@pytest.mark.reference
@pytest.mark.skipif(not rh.reference_available(), reason="live reference not present (GPU box)")
def test_interim_clusterings_against_the_binarys_progress_table(tmp_path, name, inflation):
    """Every iteration the binary interprets the dag of the pruned expansion and prints the number of clusters and of nodes in
    overlap (-v all); the oracle's dag, attractor systems and reach sets give the same pair on every line, and its chaos the
    printed one (two decimals).  (Checked by hand in the container as well: the raw overlapping clusters of `-dump cls` are
    exactly the oracle's reach sets.  WHICH cluster keeps an overlapping node of an unconverged iterand is not reproduced:
    `-L k` runs differ there; at the limit - no golden graph but synth_ties has overlap left - the rule is pinned by the clusters files.)"""
    graph = mcl_golden_graph(name)
    ref = rh.run_mcl_binary(graph, inflation, str(tmp_path), log=True)
    pr = mo.Process(*mo.parse_graph_text(graph))
    assert len(ref["log"]) == MclGolden(name).iterations(inflation)
    for chaos_printed, n_cls, n_olap, n_missing in ref["log"]:
        chaos = pr.step(inflation)
        nc, label, ov = pr.interpret(1)
        assert (nc, ov) == (n_cls + (1 if n_missing else 0), n_olap)      # nodes that reach no attractor: one more cluster
        assert abs(chaos - chaos_printed) <= 0.00501


@pytest.mark.reference
@pytest.mark.skipif(not rh.reference_available(), reason="live reference not present (GPU box)")
def test_nodes_without_an_attractor_share_one_cluster(tmp_path):
    """Stopped after 10 iterations (-L 10) the equal-weight 9-cycle of synth_ties has no attractor yet: the binary reports 9 'garbage
    entries' and writes them as ONE cluster; so does the oracle (which cluster keeps the overlapping nodes of such an unconverged
    iterand is not reproduced, so only that cluster and the cluster count are compared)."""
    import numpy as np
    graph = mcl_golden_graph("synth_ties")
    ref = rh.run_mcl_binary(graph, 1.5, str(tmp_path), extra=("-L", "10"))
    n_ref = int(ref["clusters_text"].split("dimensions ")[1].split("x")[1].split()[0])
    pr = mo.Process(*mo.parse_graph_text(graph))
    for _ in range(10):
        pr.step(1.5)
    nc, label, ov = pr.interpret(0)
    assert nc == n_ref
    cl = [c.tolist() for c in mo.clusters_from_labels(label)]
    garbage = [c for c in cl if len(c) == 9]
    assert len(garbage) == 1 and " ".join(map(str, garbage[0])) + " $" in ref["clusters_text"]
