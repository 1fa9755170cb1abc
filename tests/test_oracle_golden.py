"""CPU: the oracle restatement against the golden vectors produced by the live reference."""
import numpy as np
import pytest

from oracle import waterfall_oracle as wo
from tests.helpers import Golden, REF_GOLDENS, SYNTH_GOLDENS


@pytest.mark.parametrize("name", REF_GOLDENS + SYNTH_GOLDENS)
def test_oracle_reproduces_reference_graph(name):
    g = Golden(name)
    res = wo.waterfall(g.n_seqs, g.lengths, g.hit_text)
    text = wo.graph_text(res["indptr"], res["indices"], res["data"], sum(g.n_seqs))
    assert text == g.graph, "oracle graph differs from the live reference's (%s)" % g.provenance
    S = len(g.n_seqs)
    for p in range(S):
        for j in range(S):
            f = res["fits"][(p, j)]
            if np.isnan(g.fits[p, j, 0]):
                assert f is None
            else:
                # curve_fit parameters bit for bit
                assert f[0] == g.fits[p, j, 0] and f[1] == g.fits[p, j, 1] and f[2] == int(g.fits[p, j, 2])
    np.testing.assert_array_equal(np.concatenate(res["most_distant"]), g.most_distant)


@pytest.mark.parametrize("name", REF_GOLDENS + SYNTH_GOLDENS)
def test_oracle_reproduces_reference_graph_with_v2_scores(name):
    """`--scores-v2` (gathering.py:157-174 NormalisedBitScore, :314-388 percentile-ratio thresholds):
    the golden graph comes from the live reference run with v2_scores=True (tools/make_golden.py)."""
    g = Golden(name)
    res = wo.waterfall(g.n_seqs, g.lengths, g.hit_text, v2_scores=True)
    text = wo.graph_text(res["indptr"], res["indices"], res["data"], sum(g.n_seqs))
    assert text == g.graph_v2, "oracle v2 graph differs from the live reference's (%s)" % g.provenance
    assert np.array_equal(np.concatenate(res["most_distant"]), g.most_distant_v2)
    assert all(f is None for f in res["fits"].values())   # no fit in this mode


def test_oracle_matrix_counts_match_reference_pickles():
    g = Golden("ref_AddTwoSpecies")
    res = wo.waterfall(g.n_seqs, g.lengths, g.hit_text, return_parts=True)
    S = len(g.n_seqs)
    for p in range(S):
        for j in range(S):
            assert len(res["B"][p][j][1]) == g.nnz[p, j, 0]
            assert int(res["BH"][p][j].sum()) == g.nnz[p, j, 1]
            assert int(res["connect"][p][j].sum()) == g.nnz[p, j, 2]
    # known-answer vectors of SURVEY.md §8c
    assert tuple(g.nnz[0, 1]) == (863, 377, 319)
    assert abs(g.fits[0, 0, 0] - 0.463130427080) < 1e-12 and abs(g.fits[0, 0, 1] - 0.447571928288) < 1e-12
    assert int(g.fits[0, 0, 2]) == 208


def test_stress_config_exercises_the_quirks():
    """C5 stand-in: 'too few hits' pair present; the last-column-wins rule (gathering.py:304-306) is observable."""
    g = Golden("synth_C5_small")
    res = wo.waterfall(g.n_seqs, g.lengths, g.hit_text, return_parts=True)
    assert res["warnings"] == [(4, 3, 1)]
    B, BH = res["B"], res["BH"]
    S = len(g.n_seqs)
    differs = 0
    for p in range(S):
        for k in range(S):
            if k == p:
                continue
            indptr, cols, vals = B[p][k]
            rows, keys = wo._entry_keys(indptr, cols, g.n_seqs[k])
            i2, c2, _ = B[k][p]
            rows2, _ = wo._entry_keys(i2, c2, g.n_seqs[p])
            f2 = BH[k][p]
            tk = c2[f2].astype(np.int64) * g.n_seqs[k] + rows2[f2]
            rbh = BH[p][k] & np.isin(keys, tk)
            r, v = rows[rbh], vals[rbh]
            u, cnt = np.unique(r, return_counts=True)
            for gene in u[cnt > 1]:
                vv = v[r == gene]
                differs += vv.min() != vv[-1]
    assert differs >= 3, "stress data no longer distinguishes last-wins from min"


def test_empty_and_ragged_inputs():
    n_seqs = [3, 2]
    L = [np.array([100., 200., 300.]), np.array([150., 250.])]
    texts = {(0, 0): b"", (0, 1): b"0_0\t1_1\t1\t1\t1\t1\t1\t1\t1\t1\t1e-5\t50\n\n\r\n0_0\t1_1\t1\t1\t1\t1\t1\t1\t1\t1\t1e-5\t70.5",
             (1, 0): b"1_1\t0_0\t1\t1\t1\t1\t1\t1\t1\t1\t1e-5\t 60\r\n", (1, 1): b"1_0\t1_0\t1\t1\t1\t1\t1\t1\t1\t1\t1e-5\t99\n"}
    res = wo.waterfall(n_seqs, L, lambda p, j: texts[(p, j)])
    # every pair has < 2 hits -> all zeroed, graph has no edges but all rows
    assert res["indptr"][-1] == 0 and len(res["indptr"]) == 6
    q, s, sc = wo.parse_hits_text(texts[(0, 1)])
    assert list(q) == [0, 0] and list(s) == [1, 1] and list(sc) == [50.0, 70.5]
    raw = wo.get_blast6_scores(texts[(0, 1)], 3, 2, False)
    assert list(raw[2]) == [70.5]  # max over duplicate HSPs
    raw = wo.get_blast6_scores(texts[(1, 1)], 2, 2, True)
    assert raw[0][-1] == 0  # self hit dropped


@pytest.mark.parametrize("name", REF_GOLDENS + SYNTH_GOLDENS)
def test_oracle_reproduces_reference_homology_graph(name):
    """`--c-homologs` (gathering_version (3, 2), gathering.py:51-73, dispatched at :520-521; the rest of SURVEY N1): the golden
    comes from the live reference's WriteGraph_perSpecies_homology (tools/make_golden.py)."""
    g = Golden(name)
    res = wo.waterfall(g.n_seqs, g.lengths, g.hit_text, homology=True)
    text = wo.graph_text(res["indptr"], res["indices"], res["data"], sum(g.n_seqs))
    assert text == g.graph_hom, "oracle homology graph differs from the live reference's (%s)" % g.provenance
    assert np.all(res["data"] == 1.0) and res["indices"].size > g.graph.count(b":")


@pytest.mark.parametrize("name", REF_GOLDENS + SYNTH_GOLDENS)
def test_oracle_reproduces_reference_dendroblast_matrices(name):
    """DendroBLAST distance matrices (orthologues.py:264-291 worker on raw scores with self hits, :396-410 completion,
    :413-431 Phylip text); goldens from the live reference's own functions (tools/make_golden_dendro.py)."""
    from oracle import dendroblast_oracle as do
    from tests.helpers import DendroGolden
    g, dg = Golden(name), DendroGolden(name)
    M = do.og_matrices(g.n_seqs, g.hit_text, dg.ogs)
    assert len(M) == len(dg.M) > 0
    for iog, (og, m) in enumerate(zip(dg.ogs, M)):
        assert np.array_equal(m, dg.M[iog]), (name, iog)
        d, max_og = do.complete(m)
        assert np.array_equal(d, dg.D[iog]), (name, iog)
        assert do.phylip_text(d, ["%d_%d" % (g.species[p], q) for p, q in og], max_og) == dg.phylip[iog], (name, iog)
