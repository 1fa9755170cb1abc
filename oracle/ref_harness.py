"""Drive the UNMODIFIED live reference (/root/reference) in-process to produce golden vectors.

TEST INFRASTRUCTURE ONLY.  Runs only in the build container (the reference is not present on
the GPU box); its outputs are committed under tests/golden/ by tools/make_golden.py.

Follows the recipe in SURVEY.md §8(c): import scripts_of with a stub ``Bio`` package, then call
  gathering.GetSequenceLengths            (gathering.py:444-464)
  WaterfallMethod.ProcessBlastHits        (gathering.py:177-196)
  WaterfallMethod.ConnectCognates         (gathering.py:251-259)
  gathering.WriteGraph_perSpecies         (gathering.py:23-49)
and assemble header + parts exactly like WriteGraphParallel (gathering.py:273-291).
The fp64 B / BH / connect pickles are read back before deletion so that weights can be compared
at 1e-9 instead of the file's %.3f.
"""
import os
import re
import sys
import io
import pickle
import shutil
import tempfile
import contextlib
import warnings

import numpy as np

REFERENCE_ROOT = os.environ.get("OFG_REFERENCE_ROOT", "/root/reference")
_STUBS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "stubs")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "scripts_of"))




def _import_reference_real():
    if not reference_available():
        raise RuntimeError("live reference not present at %s" % REFERENCE_ROOT)
    for p in (_STUBS, REFERENCE_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from scripts_of import gathering, util, files, matrices
    return gathering, util, files, matrices


def run_reference(wd, q_double_blast=True, capture_fit=True, keep_matrices=True, v2_scores=False, homology=False):
    """Run the reference hot path on WorkingDirectory ``wd`` (must end with '/').

    Returns dict with: graph (bytes), seqsInfo fields, lengths, per-pair CSR of B (normalised),
    BH, connect, fit params per pair (None when the pair was zeroed), stdout text.
    """
    gathering, util, files, matrices = _import_reference_real()
    if not wd.endswith("/"):
        wd += "/"
    out = tempfile.mkdtemp(prefix="ofg_ref_") + "/"
    pic = out + "pickle/"
    os.mkdir(pic)
    fh = files.FileHandler
    saved = (fh.wd_base, fh.wd_current)
    fh.wd_base = [wd]
    fh.wd_current = out
    res = {}
    buf = io.StringIO()
    fits = {}
    try:
        sp, n_all, _ = util.GetSpeciesToUse(wd + "SpeciesIDs.txt")
        si = util.GetSeqsInfo([wd], sp, n_all)
        L = gathering.GetSequenceLengths(si)
        orig_fit = gathering.scnorm.CalculateFittingParameters
        cur = {}

        def fit_probe(Lf, S):
            pars = orig_fit(Lf, S)
            cur["pars"] = np.array(pars, dtype=np.float64)
            cur["m"] = len(S)
            return pars

        orig_norm = gathering.WaterfallMethod.NormaliseScores

        def norm_probe(B, Lengths, i, j):
            cur.clear()
            r = orig_norm(B, Lengths, i, j)
            fits[(norm_probe.p, j)] = (cur.get("pars"), cur.get("m", 0), B.nnz)
            return r

        most = {}
        orig_md = gathering.WaterfallMethod.GetMostDistant_s
        orig_md2 = gathering.WaterfallMethod.GetMostDistant_s_estimate_from_relative_rbhs

        def md_probe(RBH, B, seqsInfo, iSpec):
            m = orig_md(RBH, B, seqsInfo, iSpec)
            most[iSpec] = np.array(m, dtype=np.float64, copy=True)
            return m

        def md2_probe(RBH, B, seqsInfo, iSpec):
            m = orig_md2(RBH, B, seqsInfo, iSpec)
            most[iSpec] = np.array(m, dtype=np.float64, copy=True)
            return m

        gathering.WaterfallMethod.GetMostDistant_s = staticmethod(md_probe)
        gathering.WaterfallMethod.GetMostDistant_s_estimate_from_relative_rbhs = staticmethod(md2_probe)
        if capture_fit:
            gathering.scnorm.CalculateFittingParameters = staticmethod(fit_probe)
            gathering.WaterfallMethod.NormaliseScores = staticmethod(norm_probe)
        try:
            with contextlib.redirect_stdout(buf):
                for p in range(si.nSpecies):
                    norm_probe.p = p
                    gathering.WaterfallMethod.ProcessBlastHits(si, [wd], L, p, pic, q_double_blast, v2_scores)
                if not homology:   # gathering_version (3, 2) (--c-homologs) skips ConnectCognates (gathering.py:498,520)
                    for p in range(si.nSpecies):
                        gathering.WaterfallMethod.ConnectCognates(si, p, pic, v2_scores)
                mats = {}
                if keep_matrices and not homology:
                    for p in range(si.nSpecies):
                        for j in range(si.nSpecies):
                            for name in ("B", "BH", "connect"):
                                with open(pic + "%s%d_%d.pic" % (name, p, j), "rb") as f:
                                    mats[(name, p, j)] = pickle.load(f).tocsr()
                graph_fn = out + "OrthoFinder_graph.txt"
                with open(graph_fn, "w") as g:
                    g.write("(mclheader\nmcltype matrix\ndimensions %dx%d\n)\n" % (si.nSeqs, si.nSeqs))
                    g.write("\n(mclmatrix\nbegin\n\n")
                for p in range(si.nSpecies):
                    if homology:
                        gathering.WriteGraph_perSpecies_homology((si, graph_fn, p, pic))   # gathering.py:51-73
                    else:
                        gathering.WriteGraph_perSpecies((si, graph_fn, p, pic))
                with open(graph_fn, "ab") as g:
                    for p in range(si.nSpecies):
                        with open(graph_fn + "_%d" % p, "rb") as part:
                            g.write(part.read())
        finally:
            gathering.scnorm.CalculateFittingParameters = staticmethod(orig_fit)
            gathering.WaterfallMethod.NormaliseScores = staticmethod(orig_norm)
            gathering.WaterfallMethod.GetMostDistant_s = staticmethod(orig_md)
            gathering.WaterfallMethod.GetMostDistant_s_estimate_from_relative_rbhs = staticmethod(orig_md2)
        with open(graph_fn, "rb") as g:
            res["graph"] = g.read()
        res.update(nSeqs=si.nSeqs, nSpecies=si.nSpecies, speciesToUse=list(si.speciesToUse),
                   seqStartingIndices=list(si.seqStartingIndices),
                   nSeqsPerSpecies=dict(si.nSeqsPerSpecies), lengths=[np.asarray(x) for x in L],
                   mats=mats, fits=fits, stdout=buf.getvalue(),
                   most_distant=[most[p] for p in range(si.nSpecies)] if not homology else None)
    finally:
        fh.wd_base, fh.wd_current = saved
        shutil.rmtree(out, ignore_errors=True)
    return res


def run_reference_dendroblast(wd, ogs, q_double_blast=True):
    """The live reference's DendroBLAST distance matrices (orthologues.py:264-291 worker, :396-411 completion, :413-431
    Phylip text) for WorkingDirectory ``wd``.  ogs: list of orthogroups, each a list of (species ID, iSeq), sizes
    non-increasing and >= 4 like GetOGMatrices_FullParallel :327-337 guarantees.
    Returns dict(M=[n x n arrays before completion], D=[after], phylip=[bytes per orthogroup])."""
    import queue
    _import_reference_real()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from scripts_of import orthologues, util, files
    if not wd.endswith("/"):
        wd += "/"
    out = tempfile.mkdtemp(prefix="ofg_refd_") + "/"
    sp, n_all, _ = util.GetSpeciesToUse(wd + "SpeciesIDs.txt")
    si = util.GetSeqsInfo([wd], sp, n_all)
    og_seqs = [[orthologues.Seq((int(a), int(b))) for a, b in og] for og in ogs]
    # orthologues.py:338-341
    ogsPerSpecies = [[[(g, i) for i, g in enumerate(og) if g.iSp == iSp] for iSp in si.speciesToUse] for og in og_seqs]
    nGenes = [len(og) for og in og_seqs]
    ogMatrices = [[np.zeros(n) for _ in range(n)] for n in nGenes]

    class _Cmds:
        def __init__(self, items):
            self.items = list(items)

        def get(self, *_a):
            if not self.items:
                raise queue.Empty
            return self.items.pop(0)

    class _Status:
        def put(self, _x):
            pass

    cmds = _Cmds((iiSp, sp1, si.nSeqsPerSpecies[sp1]) for iiSp, sp1 in enumerate(si.speciesToUse))   # :344-345
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        orthologues.Worker_OGMatrices_ReadBLASTAndUpdateDistances(cmds, _Status(), 0, ogMatrices, nGenes, si, [wd], ogsPerSpecies,
                                                                  q_double_blast)
        M = [np.array([np.array(r) for r in m]) for m in ogMatrices]
        fh = files.FileHandler
        saved = getattr(fh, "GetOGsDistMatFN", None)
        fh.GetOGsDistMatFN = staticmethod(lambda iog: out + "OG%07d.phy" % iog)
        try:
            inst = orthologues.DendroBLASTTrees.__new__(orthologues.DendroBLASTTrees)
            orthologues.DendroBLASTTrees.CompleteAndWriteOGMatrices(inst, og_seqs, ogMatrices)
        finally:
            if saved is not None:
                fh.GetOGsDistMatFN = saved
    D = [np.array([np.array(r) for r in m]) for m in ogMatrices]
    phy = []
    for iog in range(len(ogs)):
        with open(out + "OG%07d.phy" % iog, "rb") as f:
            phy.append(f.read())
    shutil.rmtree(out, ignore_errors=True)
    return dict(M=M, D=D, phylip=phy, seqsInfo=si)


# ---------------------------------------------------------------- the consumer: the bundled mcl binary
MCL_BINARY = os.path.join(REFERENCE_ROOT, "scripts_of", "bin", "mcl")


def read_mcl_binary_matrix(path):
    """A matrix written by the binary with --write-binary (canonical domains): (colptr i64, rows i32, vals f32)."""
    import numpy as np
    b = open(path, "rb").read()
    nc, nr, fl = np.frombuffer(b[4:28], dtype="<i8")
    assert fl == 3, "non-canonical domains"
    offs = np.frombuffer(b[28:28 + 8 * (nc + 1)], dtype="<i8")
    base = 28 + 8 * (nc + 1)
    rec = np.dtype([("i", "<i4"), ("v", "<f4")])
    colptr = np.zeros(nc + 1, np.int64)
    parts = []
    for c in range(nc):
        o = base + int(offs[c])
        n = int(np.frombuffer(b[o:o + 8], dtype="<i8")[0])
        parts.append(np.frombuffer(b[o + 24:o + 24 + 8 * n], dtype=rec))
        colptr[c + 1] = colptr[c] + n
    a = np.concatenate(parts) if parts else np.zeros(0, rec)
    return colptr, a["i"].astype(np.int32), a["v"].astype(np.float32)


def run_mcl_binary(graph_text, inflation, workdir, dump_iterands=False, extra=(), threads=1, log=False):
    """Runs the reference's bundled mcl exactly as RunMCL does (scripts_of/mcl.py:214-218: `mcl graph -I <inflation>
    -o clusters -te <n> -V all`) on `graph_text`; returns dict(clusters_text, iterands=[(colptr, rows, vals), ...] incl.
    the start matrix when dump_iterands, expanded=(...) when extra holds -write-expanded)."""
    import glob
    import subprocess
    os.makedirs(workdir, exist_ok=True)
    for f in glob.glob(os.path.join(workdir, "ite-*.D")):
        os.remove(f)
    g = os.path.join(workdir, "graph.txt")
    with open(g, "wb") as fh:
        fh.write(graph_text if isinstance(graph_text, bytes) else graph_text.encode())
    cmd = [MCL_BINARY, "graph.txt", "-I", str(inflation), "-o", "clusters.txt", "-te", str(threads), "-v" if log else "-V", "all"]
    if dump_iterands:
        cmd += ["--write-binary", "-dump", "ite", "-dump-stem", "D", "-dump-interval", "0:100000"]
    cmd += list(extra)
    done = subprocess.run(cmd, cwd=workdir, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE if log else subprocess.DEVNULL)
    out = {"clusters_text": open(os.path.join(workdir, "clusters.txt")).read()}
    if log:
        # per iteration: (chaos, clusters of the interim dag, nodes in overlap) from the binary's progress table (-v all)
        rows = [ln.split() for ln in done.stderr.decode(errors="replace").splitlines() if re.match(r"^\s+\d+\s+\d+\.\d+\s", ln)]
        out["log"] = [(float(r[1]), int(r[10]), int(r[11]), int(r[13][4:]) if len(r) > 13 and r[13].startswith("[!m=") else 0) for r in rows]
        # (chaos, clusters, nodes in overlap, nodes in no cluster - printed as a trailing "[!m=.. e=..]")
    if dump_iterands:
        k, its = 0, []
        while os.path.exists(os.path.join(workdir, "ite-%d.D" % k)):
            its.append(read_mcl_binary_matrix(os.path.join(workdir, "ite-%d.D" % k)))
            k += 1
        out["iterands"] = its
    if "-write-expanded" in extra:
        out["expanded"] = read_mcl_binary_matrix(os.path.join(workdir, extra[list(extra).index("-write-expanded") + 1]))
    return out
