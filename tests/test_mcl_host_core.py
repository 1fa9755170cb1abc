"""CPU: the MCL engine of libofg (ofg_mcl_driver.h + the per-column functions of ofg_mcl_core.h that the CUDA kernels
run) compiled for the host and walked phase by phase (libofg_hostcheck.so) - against what the reference's bundled mcl
binary produced (tests/golden/mcl_*.npz): every iterand bit for bit, the clusters file body byte for byte.  The host build
uses glibc's pow like the binary; on the device CUDA's pow may differ in the last place (tests/test_zz_mcl_gpu.py)."""
import ctypes
import os

import numpy as np
import pytest

from tests.helpers import MCL_GOLDENS, MclGolden, matrix_digest, mcl_golden_graph

PKG = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "orthofinder_b200")


def _lib():
    L = ctypes.CDLL(os.path.join(PKG, "libofg_hostcheck.so"))
    L.ofg_hostcheck_mcl_new.restype = ctypes.c_void_p
    L.ofg_hostcheck_mcl_free.argtypes = [ctypes.c_void_p]
    L.ofg_hostcheck_mcl_params.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_int,
                                           ctypes.c_long, ctypes.c_long, ctypes.c_int]
    L.ofg_hostcheck_mcl_start.argtypes = [ctypes.c_void_p, ctypes.c_long, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    L.ofg_hostcheck_mcl_step.argtypes = [ctypes.c_void_p, ctypes.c_double, ctypes.c_void_p]
    L.ofg_hostcheck_mcl_nnz.restype = ctypes.c_long
    L.ofg_hostcheck_mcl_nnz.argtypes = [ctypes.c_void_p]
    L.ofg_hostcheck_mcl_get.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    L.ofg_hostcheck_mcl_interpret.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    L.ofg_hostcheck_mcl_text.restype = ctypes.c_long
    L.ofg_hostcheck_mcl_text.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_long]
    return L


class HostEngine:
    def __init__(self, graph_text, P=10000, S=1100, R=1400, pct=0.9, overhead=10, table_budget=0, stage_budget=0, direct_factor=-1):
        from orthofinder_b200 import mcl as pm
        self.L = _lib()
        self.n, cp, ri, v = pm.read_graph_text(graph_text)
        self.h = self.L.ofg_hostcheck_mcl_new()
        self.L.ofg_hostcheck_mcl_params(self.h, P, S, R, pct, overhead, table_budget, stage_budget, direct_factor)
        assert self.L.ofg_hostcheck_mcl_start(self.h, self.n, cp.ctypes.data, ri.ctypes.data, v.ctypes.data) == 0

    def step(self, inflation):
        ch = ctypes.c_double(0)
        assert self.L.ofg_hostcheck_mcl_step(self.h, float(inflation), ctypes.byref(ch)) == 0
        return ch.value

    def matrix(self):
        nnz = self.L.ofg_hostcheck_mcl_nnz(self.h)
        cp, ri, v = np.empty(self.n + 1, np.int64), np.empty(nnz, np.int32), np.empty(nnz, np.float32)
        assert self.L.ofg_hostcheck_mcl_get(self.h, cp.ctypes.data, ri.ctypes.data, v.ctypes.data) == 0
        return cp, ri, v

    def clusters_text(self):
        lab = np.empty(self.n, np.int32)
        nc, ov = ctypes.c_long(0), ctypes.c_long(0)
        assert self.L.ofg_hostcheck_mcl_interpret(self.h, lab.ctypes.data, ctypes.byref(nc), ctypes.byref(ov)) == 0
        size = self.L.ofg_hostcheck_mcl_text(self.h, None, 0)
        buf = ctypes.create_string_buffer(size)
        self.L.ofg_hostcheck_mcl_text(self.h, buf, size)
        return buf.raw.decode(), lab, nc.value, ov.value

    def close(self):
        self.L.ofg_hostcheck_mcl_free(self.h)


def _run_against_golden(name, inflation, **kw):
    g = MclGolden(name)
    sha, nnz = g.digests(inflation)
    e = HostEngine(mcl_golden_graph(name), **kw)
    assert matrix_digest(*e.matrix()) == sha[0]
    k = 0
    while True:
        chaos = e.step(inflation)
        k += 1
        m = e.matrix()
        assert len(m[1]) == nnz[k], (k, len(m[1]), nnz[k])
        assert matrix_digest(*m) == sha[k], "iterand %d differs" % k
        if chaos < 1e-4:
            break
    assert k == g.iterations(inflation)
    text, lab, nc, ov = e.clusters_text()
    assert text == g.clusters_body(inflation)
    e.close()


@pytest.mark.parametrize("name,inflation", [("ref_AddTwoSpecies", 1.2), ("ref_FromTrees", 1.5), ("ref_removeFirst", 3.0),
                                            ("synth_ties", 1.2), ("synth_ties", 1.5), ("synth_ties", 3.0), ("synth_C5_small", 1.5)])
def test_engine_on_the_host_reproduces_the_binary(name, inflation):
    _run_against_golden(name, inflation)


def test_engine_with_many_small_batches_and_global_tables():
    """Budgets of a few thousand entries cut every iteration into dozens of batches; C2_tiny's columns of up to 2 400 entries
    and the 1 250-gene family need hash tables in 'global' memory and lists longer than the shared-memory list."""
    _run_against_golden("synth_C2_tiny", 3.0, table_budget=20000, stage_budget=30000, direct_factor=0)
    _run_against_golden("synth_family", 3.0, table_budget=9000, stage_budget=5000, direct_factor=0)


def test_engine_with_row_indexed_accumulators():
    """Columns that may touch a large share of the rows skip the hash table and the sort (default: a quarter; here also
    every column whose table would not fit shared memory)."""
    _run_against_golden("synth_C2_tiny", 3.0)
    _run_against_golden("synth_family", 3.0, table_budget=9000, stage_budget=5000, direct_factor=1000000)
    _run_against_golden("synth_C5_small", 3.0, direct_factor=1000000)


def test_weight_as_the_binary_reads_it_from_the_graph_file():
    """"%.3f" -> strtod -> float, for the device path that never writes the file."""
    L = _lib()
    L.ofg_hostcheck_mcl_weight.argtypes = [ctypes.c_double, ctypes.c_void_p]
    rng = np.random.default_rng(5)
    xs = np.concatenate([rng.uniform(0, 4, 20000), rng.uniform(0, 3000, 5000), 10.0 ** rng.uniform(-9, 12.9, 5000),
                         (rng.integers(0, 4000, 5000) + 0.5) / 1000.0, [0.0, 0.0004, 0.0005, 0.0015, 0.0025, 9.0e12]])
    out = ctypes.c_float(0)
    for x in xs:
        assert L.ofg_hostcheck_mcl_weight(float(x), ctypes.byref(out)) == 1, x
        assert out.value == float(np.float32(float("%.3f" % x))), x
    for bad in (-1.0, float("inf"), float("nan"), 1e13, 1e19):      # beyond 2^53 thousandths: rejected, never approximated
        assert L.ofg_hostcheck_mcl_weight(bad, ctypes.byref(out)) == 0


@pytest.mark.reference
def test_clusters_file_is_read_by_the_live_reference_like_the_binarys(tmp_path):
    """The file MCL.RunMCL writes (own `# cline:` comment + the body) through the reference's own readers: GetPredictedOGs
    (mcl.py:36-63) and ConvertSingleIDsToIDPair (mcl.py:93-125) give what they give for the binary's file."""
    from oracle import ref_harness as rh
    if not rh.reference_available():
        pytest.skip("live reference not present (GPU box)")
    from orthofinder_b200 import mcl as pm
    from tests.helpers import Golden
    gathering, util, files = rh._import_reference_real()[:3]
    import scripts_of.mcl as ref_mcl
    name = "ref_AddTwoSpecies"
    graph = mcl_golden_graph(name)
    ref = rh.run_mcl_binary(graph, 1.2, str(tmp_path / "bin"))
    e = HostEngine(graph)
    while e.step(1.2) >= 1e-4:
        pass
    text, lab, nc, ov = e.clusters_text()
    e.close()
    mine = str(tmp_path / "clusters_mine.txt")
    pm._write_clusters(mine, text, 1.2)
    theirs = str(tmp_path / "bin" / "clusters.txt")
    assert ref_mcl.GetPredictedOGs(mine) == ref_mcl.GetPredictedOGs(theirs)
    assert pm.GetPredictedOGs(mine) == ref_mcl.GetPredictedOGs(theirs)
    g = Golden(name)
    starts = np.concatenate(([0], np.cumsum(g.n_seqs)[:-1]))
    seqsInfo = util.SequencesInfo(nSeqs=sum(g.n_seqs), nSpecies=len(g.n_seqs), speciesToUse=list(g.species), seqStartingIndices=[int(x) for x in starts],
                                  nSeqsPerSpecies={sp: n for sp, n in zip(g.species, g.n_seqs)})
    outs = []
    for fn in (mine, theirs):
        out = fn + ".pairs"
        ref_mcl.ConvertSingleIDsToIDPair(seqsInfo, fn, out)
        outs.append(open(out).read().split("\n", 1)[1])        # without the first line (the command-line comment)
    assert outs[0] == outs[1]


def test_graph_file_reader_of_the_library():
    """mcl_parse_graph_text (ofg_mcl_set_graph_text, what RunMCL uses) == the Python reader, on every golden graph; refusals."""
    from orthofinder_b200 import mcl as pm
    L = _lib()
    L.ofg_hostcheck_mcl_parse.restype = ctypes.c_long
    L.ofg_hostcheck_mcl_parse.argtypes = [ctypes.c_char_p, ctypes.c_long, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_long, ctypes.c_void_p,
                                          ctypes.c_void_p, ctypes.c_long, ctypes.c_char_p, ctypes.c_long]
    err = ctypes.create_string_buffer(256)
    for name in MCL_GOLDENS:
        text = bytes(mcl_golden_graph(name))
        n, cp, ri, v = pm.read_graph_text(text)
        n2 = ctypes.c_long(0)
        cp2, ri2, v2 = np.empty(n + 1, np.int64), np.empty(len(ri) + 1, np.int32), np.empty(len(ri) + 1, np.float32)
        k = L.ofg_hostcheck_mcl_parse(text, len(text), ctypes.byref(n2), cp2.ctypes.data, n + 1, ri2.ctypes.data, v2.ctypes.data, len(ri) + 1, err, 256)
        assert k == len(ri) and n2.value == n, (name, err.value)
        assert np.array_equal(cp2, cp) and np.array_equal(ri2[:k], ri) and np.array_equal(v2[:k].view(np.uint32), v.view(np.uint32)), name
    head = b"(mclheader\nmcltype matrix\ndimensions 3x3\n)\n(mclmatrix\nbegin\n"
    good = head + b"0 1:0.5 2:1e-3 $\n1 0:0.5\n   2 $\n)\n"          # continuation line, entry without a value (reads as 1)
    n2 = ctypes.c_long(0)
    cp2, ri2, v2 = np.empty(4, np.int64), np.empty(8, np.int32), np.empty(8, np.float32)
    assert L.ofg_hostcheck_mcl_parse(good, len(good), ctypes.byref(n2), cp2.ctypes.data, 4, ri2.ctypes.data, v2.ctypes.data, 8, err, 256) == 4
    assert list(cp2) == [0, 2, 4, 4] and list(ri2[:4]) == [1, 2, 0, 2] and list(v2[:4]) == [np.float32(0.5), np.float32(1e-3), np.float32(0.5), 1.0]
    for bad in (b"nothing", head.replace(b"3x3", b"3x4") + b")\n", head + b"1 0:1 $\n0 1:1 $\n)\n", head + b"0 2:1 1:1 $\n)\n",
                head + b"0 7:1 $\n)\n", head + b"0 1:abc $\n)\n", head + b"0 1:-2 $\n)\n"):
        assert L.ofg_hostcheck_mcl_parse(bad, len(bad), ctypes.byref(n2), cp2.ctypes.data, 4, ri2.ctypes.data, v2.ctypes.data, 8, err, 256) == -1, bad


def _clique_graph(n, n_cliques, seed):
    """Disjoint cliques on random members of 0..n-1 (all other nodes without edges): cluster lines with members of 1 to
    len(str(n)) digits, to exercise the line wrapping of the clusters file."""
    from tests.helpers import mcl_graph_text
    rng = np.random.default_rng(seed)
    perm = rng.permutation(n)
    cols = [dict() for _ in range(n)]
    pos = 0
    for _ in range(n_cliques):
        size = int(rng.integers(15, 60))
        members = [int(x) for x in perm[pos:pos + size]]
        pos += size
        for a in members:
            cols[a] = {b: 1.0 for b in members if b != a}
    return mcl_graph_text(n, cols)


@pytest.mark.parametrize("n,n_cliques", [(9, 1), (700, 8), (12000, 60), (120000, 150)])
def test_clusters_text_wrapping_for_every_id_width(n, n_cliques):
    """The C++ writer of the library (mcl_clusters_text) == the oracle's writer (fitted to the binary's files for id widths 1-6)."""
    from oracle import mcl_oracle as mo
    graph = _clique_graph(n, n_cliques if n > 9 else 0, 11) if n > 9 else mcl_golden_graph("synth_ties")
    e = HostEngine(graph)
    while e.step(2.0) >= 1e-4:
        pass
    text, lab, nc, ov = e.clusters_text()
    e.close()
    assert text == mo.clusters_text_body(lab, e.n)


@pytest.mark.reference
@pytest.mark.parametrize("n,n_cliques", [(700, 8), (12000, 60), (120000, 150)])
def test_clusters_text_wrapping_against_the_binary(tmp_path, n, n_cliques):
    from oracle import mcl_oracle as mo, ref_harness as rh
    if not rh.reference_available():
        pytest.skip("live reference not present (GPU box)")
    graph = _clique_graph(n, n_cliques, 11)
    ref = rh.run_mcl_binary(graph, 2.0, str(tmp_path), threads=4)
    e = HostEngine(graph)
    while e.step(2.0) >= 1e-4:
        pass
    text, lab, nc, ov = e.clusters_text()
    e.close()
    assert text == mo.clusters_body_of_file(ref["clusters_text"])


@pytest.mark.reference
@pytest.mark.parametrize("name,variant,inflation", [("ref_AddTwoSpecies", "graph_v2", 1.5), ("ref_FromTrees", "graph_hom", 1.2),
                                                    ("synth_C5_small", "graph_v2", 2.0), ("ref_removeFirst", "graph_hom", 4.0)])
def test_other_graphs_of_the_path_against_the_binary(tmp_path, name, variant, inflation):
    """The --scores-v2 and --c-homologs graphs (all weights 1.000 / 2.000 there: ties everywhere) through the binary, the oracle and
    the host-walked engine: iterands bit for bit, clusters file byte for byte."""
    from oracle import mcl_oracle as mo, ref_harness as rh
    if not rh.reference_available():
        pytest.skip("live reference not present (GPU box)")
    from tests.helpers import Golden
    graph = getattr(Golden(name), variant)
    ref = rh.run_mcl_binary(graph, inflation, str(tmp_path), dump_iterands=True)
    e = HostEngine(graph)
    pr = mo.Process(*mo.parse_graph_text(graph))
    assert matrix_digest(*e.matrix()) == matrix_digest(*ref["iterands"][0]) == matrix_digest(*pr.matrix(0))
    for k in range(1, len(ref["iterands"])):
        ch = e.step(inflation)
        ch2 = pr.step(inflation)
        want = matrix_digest(*ref["iterands"][k])
        assert matrix_digest(*e.matrix()) == want and matrix_digest(*pr.matrix(0)) == want, k
        assert (ch < 1e-4) == (k == len(ref["iterands"]) - 1) == (ch2 < 1e-4), (k, ch, ch2)
    text = e.clusters_text()[0]
    e.close()
    assert text == mo.clusters_body_of_file(ref["clusters_text"])
    assert mo.clusters_text_body(pr.interpret(0)[1], pr.n) == text


def test_unconverged_iterand_engine_and_oracle_agree():
    """After 10 of 17 iterations synth_ties still has nodes in overlap and nine nodes without an attractor (one shared cluster):
    the engine's dag / label rounds / cluster order give the oracle's clusters."""
    from oracle import mcl_oracle as mo
    graph = mcl_golden_graph("synth_ties")
    e = HostEngine(graph)
    pr = mo.Process(*mo.parse_graph_text(graph))
    for _ in range(10):
        e.step(1.5)
        pr.step(1.5)
    text, lab, nc, ov = e.clusters_text()
    nc2, lab2, ov2 = pr.interpret(0)
    e.close()
    assert (nc, ov) == (nc2, ov2) and ov > 0
    assert np.array_equal(lab, lab2) and text == mo.clusters_text_body(lab2, pr.n)
    assert 9 in np.bincount(lab).tolist()


def _random_graph(rng):
    """2-70 nodes, four weight styles (continuous, a few equal values, values that print as 0.000 or with large integer parts,
    heavy-tailed), loops in the input, one-way and unequal two-way edges, isolated nodes."""
    from tests.helpers import mcl_graph_text
    n = int(rng.integers(2, 70))
    cols = [dict() for _ in range(n)]
    kind = int(rng.integers(0, 4))
    dens = float(rng.choice([0.03, 0.1, 0.3, 0.8]))
    for i in range(n):
        for j in range(n):
            if rng.random() < dens and (i != j or rng.random() < 0.2):
                if kind == 0:
                    w = float(rng.uniform(0.001, 3))
                elif kind == 1:
                    w = float(rng.choice([0.5, 1.0, 1.0, 2.0]))
                elif kind == 2:
                    w = float(rng.choice([0.0004, 0.001, 0.25, 1.5, 12.345, 250.0]))
                else:
                    w = round(float(rng.lognormal(0, 1.5)), 3)
                cols[i][j] = w
                if rng.random() < 0.7:
                    cols[j][i] = w if rng.random() < 0.6 else w * float(rng.uniform(0.5, 1.5))
    return n, mcl_graph_text(n, cols)


@pytest.mark.reference
def test_random_graphs_against_the_binary(tmp_path):
    """120 random graphs at random inflations through the binary, the oracle and the host-walked engine: every iterand bit for
    bit, the same last iteration, the clusters file byte for byte (550 more were run by hand in the build container)."""
    from oracle import mcl_oracle as mo, ref_harness as rh
    if not rh.reference_available():
        pytest.skip("live reference not present (GPU box)")
    rng = np.random.default_rng(77)
    for trial in range(120):
        n, graph = _random_graph(rng)
        inflation = float(rng.choice([1.2, 1.5, 2.0, 3.3, 5.0]))
        ref = rh.run_mcl_binary(graph, inflation, str(tmp_path / "w"), dump_iterands=True)
        its = ref["iterands"]
        pr = mo.Process(*mo.parse_graph_text(graph))
        e = HostEngine(graph)
        assert matrix_digest(*pr.matrix(0)) == matrix_digest(*e.matrix()) == matrix_digest(*its[0]), trial
        for k in range(1, len(its)):
            ch, ch2 = pr.step(inflation), e.step(inflation)
            want = matrix_digest(*its[k])
            assert matrix_digest(*pr.matrix(0)) == want and matrix_digest(*e.matrix()) == want, (trial, k)
            assert (ch < 1e-4) == (ch2 < 1e-4) == (k == len(its) - 1), (trial, k, ch, ch2)
        body = mo.clusters_body_of_file(ref["clusters_text"])
        assert mo.clusters_text_body(pr.interpret(0)[1], n) == body, trial
        assert e.clusters_text()[0] == body, trial
        e.close()


@pytest.mark.parametrize("P,S,R,pct", [(200, 10, 15, 90), (100, 8, 20, 95), (50, 20, 5, 90), (20, 3, 8, 99), (10000, 4, 4, 50), (40, 6, 12, 97)])
def test_engine_pruning_branches_equal_the_oracle(P, S, R, pct):
    """Small prune / select / recover numbers (the binary's -P -S -R -pct; ofg_mcl_set_params) make every branch of the pruning rule
    and the radix select fire on a fixture-sized graph: the engine's iterands and clusters == the oracle's, which
    tests/test_mcl_oracle.py pins against the binary for the same settings."""
    from oracle import mcl_oracle as mo
    graph = mcl_golden_graph("ref_AddTwoSpecies")
    kw = dict(P=P, S=S, R=R, pct=pct / 100.0)
    e = HostEngine(graph, **kw)
    pr = mo.Process(*mo.parse_graph_text(graph))
    for k in range(1, 200):
        ch, ch2 = pr.step(1.2, **kw), e.step(1.2)
        assert matrix_digest(*e.matrix()) == matrix_digest(*pr.matrix(0)), k
        assert (ch < 1e-4) == (ch2 < 1e-4)
        if ch < 1e-4:
            break
    text = e.clusters_text()[0]
    e.close()
    assert text == mo.clusters_text_body(pr.interpret(0)[1], pr.n)
