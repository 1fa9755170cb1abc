"""The pin of the oracle: the UNMODIFIED live reference (/root/reference, driven in-process by oracle/ref_harness.py) ==
the committed goldens == the CPU oracle, on every usable fixture of SURVEY.md section 8(c).  Runs only where the
reference is present (the build container); on the GPU box these tests are skipped and the goldens stand in.

Also checks the drop-in activation of INTEGRATION.md section 1 (orthofinder_b200/integration.py) against the live
gathering module with a stub builder: the replacement of DoOrthogroups must hand seqsInfo, the blast directory list
(only the first directory for unassigned-clade runs, gathering.py:482-483), qDoubleBlast, v2_scores and q_allow_empty
to the builder and write to FileHandler.GetGraphFilename(i_unassigned) (files.py:324-327).
"""
import os

import numpy as np
import pytest

from oracle import ref_harness as rh, waterfall_oracle as wo
from tests.helpers import Golden

pytestmark = [pytest.mark.reference,
              pytest.mark.skipif(not rh.reference_available(), reason="live reference not present (GPU box)")]

REF_TESTS = os.path.join(rh.REFERENCE_ROOT, "tests")
# name -> (directory, golden committed under tests/golden/ or None)
FIXTURES = {
    "AddTwoSpecies": ("ExpectedOutput/AddTwoSpecies/", "ref_AddTwoSpecies"),
    "FromTrees": ("Input/FromTrees/WorkingDirectory/", "ref_FromTrees"),
    "removeFirst": ("Input/ExampleDataset_removeFirst/", "ref_removeFirst"),
    "FromOrthogroups": ("Input/FromOrthogroups/WorkingDirectory/", None),
    "AddOneSpecies": ("ExpectedOutput/AddOneSpecies/", None),
    "removeMiddle": ("Input/ExampleDataset_removeMiddle/", None),
    "removeLast": ("Input/ExampleDataset_removeLast/", None),
    "ExampleBlastDir": ("Input/SmallExampleDataset_ExampleBlastDir/", None),
    "skipSpecies": ("Input/SmallExampleDataset_ExampleBlastDir_skipSpecies/", None),
    "Issue83": ("Input/ISSUES/Issue83/", None),
}


@pytest.mark.parametrize("name", sorted(FIXTURES))
def test_live_reference_equals_oracle_and_goldens(name):
    sub, golden = FIXTURES[name]
    wd = os.path.join(REF_TESTS, sub)
    ref = rh.run_reference(wd)
    mine = wo.waterfall_from_directory(wd)
    assert ref["graph"] == mine["graph"]                        # OrthoFinder_graph.txt, byte for byte
    n_fit = 0
    for (p, j), (pars, m, raw_nnz) in ref["fits"].items():       # curve_fit parameters, bit for bit
        f = mine["fits"][(p, j)]
        if pars is None:
            assert f is None
        else:
            assert f is not None and f[0] == pars[0] and f[1] == pars[1] and f[2] == m, (p, j)
            n_fit += 1
    assert n_fit > 0
    assert np.array_equal(np.concatenate(ref["most_distant"]), np.concatenate(mine["most_distant"]))   # thresholds (gathering.py:294-310)
    if golden:
        g = Golden(golden)
        assert g.graph == ref["graph"]
        assert np.array_equal(g.most_distant, np.concatenate(ref["most_distant"]))
        S = len(g.n_seqs)
        for p in range(S):
            for j in range(S):
                pars = ref["fits"][(p, j)][0]
                if pars is None:
                    assert np.isnan(g.fits[p, j, 0])
                else:
                    assert g.fits[p, j, 0] == pars[0] and g.fits[p, j, 1] == pars[1]


@pytest.mark.parametrize("name,double", [("AddTwoSpecies", True), ("FromTrees", True), ("removeFirst", True), ("FromTrees", False)])
def test_live_reference_dendroblast_equals_oracle_and_goldens(name, double):
    """the reference's own worker / completion / Phylip writer (orthologues.py:264-291, 396-431) on the orthogroups
    of the committed goldens == oracle/dendroblast_oracle.py (bit for bit) == the goldens."""
    from oracle import dendroblast_oracle as do
    from tests.helpers import DendroGolden
    sub, golden = FIXTURES[name]
    wd = os.path.join(REF_TESTS, sub)
    g, dg = Golden(golden), DendroGolden(golden)
    ids = [[(g.species[p], q) for p, q in og] for og in dg.ogs]
    ref = rh.run_reference_dendroblast(wd, ids, q_double_blast=double)
    M = do.og_matrices(g.n_seqs, g.hit_text, dg.ogs, q_double_blast=double)
    for iog, og in enumerate(ids):
        assert np.array_equal(M[iog], ref["M"][iog]), iog
        d, max_og = do.complete(M[iog])
        assert np.array_equal(d, ref["D"][iog]), iog
        assert do.phylip_text(d, ["%d_%d" % x for x in og], max_og) == ref["phylip"][iog], iog
        if double:
            assert np.array_equal(dg.M[iog], ref["M"][iog]) and np.array_equal(dg.D[iog], ref["D"][iog]) and dg.phylip[iog] == ref["phylip"][iog]


@pytest.mark.parametrize("name", ["AddTwoSpecies", "removeMiddle", "Issue83"])
def test_live_reference_equals_oracle_with_v2_scores(name):
    wd = os.path.join(REF_TESTS, FIXTURES[name][0])
    ref = rh.run_reference(wd, capture_fit=False, keep_matrices=False, v2_scores=True)
    mine = wo.waterfall_from_directory(wd, v2_scores=True)
    assert ref["graph"] == mine["graph"]
    assert np.array_equal(np.concatenate(ref["most_distant"]), np.concatenate(mine["most_distant"]))
    golden = FIXTURES[name][1]
    if golden:
        assert Golden(golden).graph_v2 == ref["graph"]


class _Options:
    def __init__(self, **kw):
        self.gathering_version = (1, 0)
        self.qDoubleBlast = True
        self.v2_scores = False
        self.nProcessAlg = 1
        self.mclInflation = 1.2
        self.__dict__.update(kw)


@pytest.mark.parametrize("i_unassigned,double,v2,version", [(None, True, False, (1, 0)), (3, False, True, (1, 0)), (None, True, False, (3, 2))])
def test_patched_DoOrthogroups_plumbing(tmp_path, i_unassigned, double, v2, version):
    """INTEGRATION.md section 1 executed against the live gathering module, with a stub in place of the CUDA builder."""
    gathering, util, files, matrices = rh._import_reference_real()
    from orthofinder_b200 import integration, waterfall
    wd = os.path.join(REF_TESTS, FIXTURES["FromTrees"][0])
    sp, n_all, _ = util.GetSpeciesToUse(wd + "SpeciesIDs.txt")
    assert (sp, n_all) == tuple(waterfall.GetSpeciesToUse(wd + "SpeciesIDs.txt")[:2])
    si = util.GetSeqsInfo([wd], sp, n_all)
    mine_si = waterfall.GetSeqsInfo([wd], sp, n_all)
    assert tuple(si) == tuple(mine_si)
    out = str(tmp_path) + "/"
    fh = files.FileHandler
    saved = (fh.wd_base, fh.wd_current)
    fh.wd_base = [wd, out]          # two directories: the unassigned run must only search the first
    fh.wd_current = out
    calls = {}

    def build_graph(seqsInfo, blastDir_list, Lengths, qDoubleBlast, v2_scores, q_allow_empty, device, homology=False):
        calls["build"] = dict(seqsInfo=seqsInfo, blastDir_list=list(blastDir_list), n_len=[len(x) for x in Lengths],
                              qDoubleBlast=qDoubleBlast, v2_scores=v2_scores, q_allow_empty=q_allow_empty, homology=homology)
        return object()

    def write_graph(gb, seqsInfo, graphFN):
        calls["graphFN"] = graphFN
        open(graphFN, "w").write("stub")
        return graphFN

    mcl_calls = []
    saved_mcl = (gathering.mcl.MCL.RunMCL, gathering.mcl.ConvertSingleIDsToIDPair, gathering.post_clustering_orthogroups)
    gathering.mcl.MCL.RunMCL = staticmethod(lambda g, c, n, i: mcl_calls.append((g, c, n, i)))
    gathering.mcl.ConvertSingleIDsToIDPair = lambda *a, **k: mcl_calls.append(("convert",) + a)
    gathering.post_clustering_orthogroups = lambda *a, **k: mcl_calls.append(("post",))
    try:
        integration.patch_gathering(gathering, build_graph=build_graph, write_graph=write_graph)
        opts = _Options(qDoubleBlast=double, v2_scores=v2, gathering_version=version)
        gathering.DoOrthogroups(opts, None, si, {}, None, i_unassigned)
    finally:
        integration.unpatch_gathering(gathering)
        gathering.mcl.MCL.RunMCL, gathering.mcl.ConvertSingleIDsToIDPair, gathering.post_clustering_orthogroups = saved_mcl
        fh.wd_base, fh.wd_current = saved
    b = calls["build"]
    assert b["seqsInfo"] is si and b["qDoubleBlast"] == double and b["v2_scores"] == v2
    assert b["homology"] == (version == (3, 2))      # --c-homologs: gathering.py:520-521
    assert b["n_len"] == [si.nSeqsPerSpecies[k] for k in si.speciesToUse]
    if i_unassigned is None:
        assert b["blastDir_list"] == [wd, out] and b["q_allow_empty"] is False
        assert calls["graphFN"] == out + "OrthoFinder_graph.txt" == waterfall.GetGraphFilename(out)
        assert mcl_calls[-1] == ("post",)
    else:
        assert b["blastDir_list"] == [wd] and b["q_allow_empty"] is True          # gathering.py:482-483, :492
        assert calls["graphFN"] == out + "OrthoFinder_unassigned_clade_3_graph.txt" == waterfall.GetGraphFilename(out, 3)
        assert all(c[0] != "post" for c in mcl_calls)                              # gathering.py:525
    assert mcl_calls[0][0] == calls["graphFN"] and mcl_calls[0][3] == 1.2           # mcl.MCL.RunMCL(graphFilename, ..., inflation)


def test_patched_DoOrthogroups_can_cluster_on_the_device(tmp_path):
    """mcl_on_device=True: the builder goes to RunMCLOnBuilder with the clusters filename of FileHandler.CreateUnusedClustersFN and
    the inflation; the reference's external mcl process is not started; ConvertSingleIDsToIDPair gets the same clusters file."""
    gathering, util, files, matrices = rh._import_reference_real()
    from orthofinder_b200 import integration
    wd = os.path.join(REF_TESTS, FIXTURES["FromTrees"][0])
    sp, n_all, _ = util.GetSpeciesToUse(wd + "SpeciesIDs.txt")
    si = util.GetSeqsInfo([wd], sp, n_all)
    out = str(tmp_path) + "/"
    fh = files.FileHandler
    saved = (fh.wd_base, fh.wd_current)
    fh.wd_base, fh.wd_current = [wd], out
    builder, calls = object(), []
    saved_mcl = (gathering.mcl.MCL.RunMCL, gathering.mcl.ConvertSingleIDsToIDPair, gathering.post_clustering_orthogroups)
    gathering.mcl.MCL.RunMCL = staticmethod(lambda *a: calls.append(("binary",) + a))
    gathering.mcl.ConvertSingleIDsToIDPair = lambda *a, **k: calls.append(("convert",) + a)
    gathering.post_clustering_orthogroups = lambda *a, **k: calls.append(("post",))
    try:
        integration.patch_gathering(gathering, build_graph=lambda *a, **k: builder, write_graph=lambda gb, s_, fn: fn, mcl_on_device=True,
                                    run_mcl_on_builder=lambda gb, fn, infl: calls.append(("device", gb, fn, infl)))
        gathering.DoOrthogroups(_Options(qDoubleBlast=True, v2_scores=False, gathering_version=(1, 0)), None, si, {}, None, None)
    finally:
        integration.unpatch_gathering(gathering)
        gathering.mcl.MCL.RunMCL, gathering.mcl.ConvertSingleIDsToIDPair, gathering.post_clustering_orthogroups = saved_mcl
        fh.wd_base, fh.wd_current = saved
    assert [c[0] for c in calls] == ["device", "convert", "post"]
    assert calls[0][1] is builder and calls[0][3] == 1.2 and calls[0][2].startswith(out + "clusters_OrthoFinder_I1.2")
    assert calls[1][2] == calls[0][2]                # ConvertSingleIDsToIDPair(seqsInfo, clustersFilename, ...)
