"""GPU: MCL on the device through the C-ABI (ofg_mcl_*) against what the reference's bundled mcl binary
produced (tests/golden/mcl_*.npz) and against the CPU oracle.

Bar: the clusters file body byte-identical to the binary's, the same number of iterations, every iterand with the binary's
structure; values bit-equal to the binary's iterands unless CUDA's pow (<= 2 ulp in double, where glibc's is < 1) rounds
an inflated entry to the neighbouring float - then the run is compared with the oracle entry by entry within 1e-5
relative (a last-place difference of one entry spreads to its column through the rescaling)."""
import os

import numpy as np
import pytest

from oracle import mcl_oracle as mo
from orthofinder_b200 import mcl as pm
from tests.helpers import MCL_GOLDENS, MclGolden, matrix_digest, mcl_golden_graph

pytestmark = pytest.mark.gpu


def _stepwise(name, inflation, **budgets):
    g = MclGolden(name)
    sha, nnz = g.digests(inflation)
    graph = mcl_golden_graph(name)
    n, cp, ri, v = pm.read_graph_text(graph)
    m = pm.DeviceMCL(0)
    if budgets:
        m.set_budgets(**budgets)
    m.set_graph(n, cp, ri, v)
    assert matrix_digest(*m.iterand()) == sha[0], "start matrix differs"
    oracle = None
    k, bit_equal = 0, 0
    while True:
        chaos = m.step(inflation)
        k += 1
        it = m.iterand()
        assert len(it[1]) == nnz[k], (name, inflation, k, len(it[1]), nnz[k])
        if oracle is None and matrix_digest(*it) == sha[k]:
            bit_equal += 1
        else:
            if oracle is None:                      # first last-place difference: bring the oracle to this iteration
                oracle = mo.Process(*mo.parse_graph_text(graph))
                for _ in range(k - 1):
                    oracle.step(inflation)
            oracle.step(inflation)
            ocp, ori, ov = oracle.matrix(0)
            assert np.array_equal(it[0], ocp) and np.array_equal(it[1], ori), (name, inflation, k, "structure differs")
            assert np.allclose(it[2], ov, rtol=1e-5, atol=0), (name, inflation, k)
        if chaos < 1e-4:
            break
        assert k <= g.iterations(inflation)
    assert k == g.iterations(inflation)
    res = m.interpret()
    assert res["iterations"] == k
    assert m.clusters_text() == g.clusters_body(inflation)
    lab = m.labels()
    assert lab.min() == 0 and lab.max() == res["n_clusters"] - 1
    m.close()
    return bit_equal, k


@pytest.mark.parametrize("name,inflation", [("ref_AddTwoSpecies", 1.2), ("ref_FromTrees", 1.5), ("ref_removeFirst", 3.0),
                                            ("synth_ties", 1.2), ("synth_ties", 3.0), ("synth_C5_small", 1.5), ("synth_C2_tiny", 3.0)])
def test_iterands_and_clusters_equal_the_binary(name, inflation):
    bit_equal, k = _stepwise(name, inflation)
    print("%s I=%s: %d of %d iterands bit-equal to the binary's" % (name, inflation, bit_equal, k))


def test_family_graph_reaches_selection_and_recovery():
    _stepwise("synth_family", 3.0)


def test_small_batches_hash_tables_in_global_memory_and_row_indexed_accumulators_agree():
    _stepwise("synth_C2_tiny", 3.0, table_slots=20000, stage_entries=30000, direct_factor=0)      # hash tables only
    _stepwise("synth_family", 3.0, table_slots=9000, stage_entries=5000, direct_factor=0)
    _stepwise("synth_family", 3.0, table_slots=9000, stage_entries=5000, direct_factor=1000000)   # row-indexed wherever shared memory is too small
    _stepwise("synth_C5_small", 1.5, direct_factor=1000000)


@pytest.mark.parametrize("name", MCL_GOLDENS)
def test_whole_run_writes_the_binarys_clusters(name):
    g = MclGolden(name)
    n, cp, ri, v = pm.read_graph_text(mcl_golden_graph(name))
    m = pm.DeviceMCL(0)
    m.set_graph(n, cp, ri, v)
    res = m.run(1.5)
    assert res["iterations"] == g.iterations(1.5)
    assert m.clusters_text() == g.clusters_body(1.5)
    assert res["launches"] > 0 and res["ms"] > 0
    m.close()


def test_graph_from_device_memory_gives_the_clusters_of_the_graph_file(tmp_path):
    """GraphBuilder -> ofg_graph_dev -> ofg_mcl_set_graph_dev: the "%.3f" rounding happens on the device, so the clusters are
    those the binary finds in the graph FILE (golden of synth_C2_tiny); RunMCL on the written file gives the same."""
    from orthofinder_b200 import synth, waterfall
    ids, n_seqs, L, P = synth.config_inputs("C2_tiny")
    gb = waterfall.GraphBuilder(0)
    gb.set_species(n_seqs, L)
    gb.synth_hits(P, ids)
    gb.build()
    g = MclGolden("synth_C2_tiny")
    out = str(tmp_path / "clusters_dev.txt")
    res = pm.MCL.RunMCLOnBuilder(gb, out, 1.2)
    body = open(out).read()
    assert body[body.index("(mclheader"):] == g.clusters_body(1.2)
    assert res["iterations"] == g.iterations(1.2)
    graph_fn = str(tmp_path / "OrthoFinder_graph.txt")
    with open(graph_fn, "wb") as f:
        f.write(waterfall.graph_header(sum(n_seqs)) + gb.graph_text_body().tobytes() + waterfall.GRAPH_FOOTER)
    out2 = str(tmp_path / "clusters_file.txt")
    pm.MCL.RunMCL(graph_fn, out2, 4, 1.2)
    assert open(out2).read() == body
    ogs = pm.GetPredictedOGs(out2)
    assert len(ogs) == res["n_clusters"] and sum(len(o) for o in ogs) == sum(n_seqs)
    gb.close()


def test_errors():
    m = pm.DeviceMCL(0)
    with pytest.raises(RuntimeError):
        m.run(1.2)                                  # no graph
    cp = np.array([0, 1, 2], np.int64)
    with pytest.raises(RuntimeError):
        m.set_graph(2, cp, np.array([1, 5], np.int32), np.array([1, 1], np.float32))      # row out of range
    with pytest.raises(RuntimeError):
        m.set_graph(2, cp, np.array([1, 0], np.int32), np.array([1, -1], np.float32))     # negative weight
    m.set_graph(2, cp, np.array([1, 0], np.int32), np.array([1, 1], np.float32))
    with pytest.raises(RuntimeError):
        m.run(0.0)
    res = m.run(2.0)
    assert res["n_clusters"] == 1 and list(m.labels()) == [0, 0]
    m.close()
