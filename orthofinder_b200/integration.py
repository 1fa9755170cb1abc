"""Drop-in activation inside the reference: `patch_gathering(gathering)` replaces gathering.DoOrthogroups by a function
that keeps the reference's own orchestration (lengths, blast directory list, graph filename, MCL, cluster conversion,
post-clustering tables) and hands the graph stage - gathering.py:478-512, i.e. ProcessBlastHits, ConnectCognates and
WriteGraphParallel for every species - to libofg.  This is INTEGRATION.md section 1 as code:

    from scripts_of import gathering
    from orthofinder_b200 import integration
    integration.patch_gathering(gathering)          # before __main__.main() reaches DoOrthogroups

Everything the replacement touches outside this repo is looked up on the `gathering` module object that is passed in
(files.FileHandler, util, mcl, GetSequenceLengths, post_clustering_orthogroups), so the reference stays unmodified.
CUDA is first touched inside BuildGraph, i.e. after the ParallelTaskManager singleton has forked (__main__.py:42-43).
"""
from . import waterfall


def patch_gathering(gathering, device=0, build_graph=None, write_graph=None, mcl_on_device=False, run_mcl_on_builder=None):
    """build_graph / write_graph default to waterfall.WaterfallMethod.BuildGraph / WriteGraphParallel (tests pass stubs).
    mcl_on_device: cluster the graph that is still in the builder's device memory (orthofinder_b200.mcl.MCL.RunMCLOnBuilder,
    INTEGRATION.md section 7) instead of starting the reference's external `mcl` process on the graph file (mcl.py:214-218);
    the graph file is written all the same and the clusters file has the binary's body."""
    files, util, mcl = gathering.files, gathering.util, gathering.mcl
    build_graph = build_graph or waterfall.WaterfallMethod.BuildGraph
    write_graph = write_graph or waterfall.WaterfallMethod.WriteGraphParallel
    reference_DoOrthogroups = gathering.DoOrthogroups

    def DoOrthogroups(options, speciesInfoObj, seqsInfo, speciesNamesDict, speciesXML=None, i_unassigned=None):
        homology = options.gathering_version == (3, 2)       # --c-homologs (__main__.py:520-521, gathering.py:520-524)
        if not (options.gathering_version < (3, 0) or homology):
            return reference_DoOrthogroups(options, speciesInfoObj, seqsInfo, speciesNamesDict, speciesXML, i_unassigned)
        q_unassigned = i_unassigned is not None
        util.PrintUnderline("Running OrthoFinder algorithm" + (" for clade-specific genes" if q_unassigned else ""))
        Lengths = gathering.GetSequenceLengths(seqsInfo)                     # gathering.py:476
        util.PrintTime("Initial processing of each species")
        blastDir_list = files.FileHandler.GetBlastResultsDir()               # :481
        if q_unassigned:
            blastDir_list = blastDir_list[:1]                                # :482-483
        try:
            kw = dict(homology=True) if homology else {}
            gb = build_graph(seqsInfo, blastDir_list, Lengths, qDoubleBlast=options.qDoubleBlast, v2_scores=options.v2_scores,
                             q_allow_empty=q_unassigned, device=device, **kw)   # :484-509 (q_allow_empty: :492)
        except waterfall.OfgError:
            util.Fail()                                                      # what ManageQueue does when a worker dies
            raise
        if not homology:
            util.PrintTime("Connected putative homologues")                  # :511
        graphFilename = write_graph(gb, seqsInfo, files.FileHandler.GetGraphFilename(i_unassigned))   # :512, files.py:324-327
        clustersFilename, clustersFilename_pairs = files.FileHandler.CreateUnusedClustersFN(
            "_I%0.1f" % options.mclInflation, i_unassigned)                  # :515-516
        if mcl_on_device:
            if run_mcl_on_builder is None:
                from . import mcl as device_mcl
                device_mcl.MCL.RunMCLOnBuilder(gb, clustersFilename, options.mclInflation, device=device)
            else:
                run_mcl_on_builder(gb, clustersFilename, options.mclInflation)
            util.PrintTime("Ran MCL")                                        # mcl.py:218
        if hasattr(gb, "close"):
            gb.close()
        # unchanged from here on (gathering.py:517-527)
        if not mcl_on_device:
            mcl.MCL.RunMCL(graphFilename, clustersFilename, options.nProcessAlg, options.mclInflation)
        mcl.ConvertSingleIDsToIDPair(seqsInfo, clustersFilename, clustersFilename_pairs, q_unassigned)
        if not q_unassigned:
            gathering.post_clustering_orthogroups(clustersFilename_pairs, speciesInfoObj, seqsInfo, speciesNamesDict, options, speciesXML)
        return clustersFilename_pairs

    DoOrthogroups.reference = reference_DoOrthogroups
    gathering.DoOrthogroups = DoOrthogroups
    return DoOrthogroups


def unpatch_gathering(gathering):
    ref = getattr(gathering.DoOrthogroups, "reference", None)
    if ref is not None:
        gathering.DoOrthogroups = ref
