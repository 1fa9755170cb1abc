"""Host mirror of the reference's DendroBLAST distance-matrix stage (scripts_of/orthologues.py:264-431), backed by libofg.

SURVEY.md §8f row N4: the second consumer of GetBLAST6Scores.  The hit files are parsed on the device exactly like for the
graph (RAW stage, self hits kept: orthologues.py:272-274), the per-gene min / max (:275-282, lil_minmax :226-236), the
orthogroup matrices M_ij = 0.5 * max(B_ij, Bmin_i) / Bmax_i (:283-287) and their completion D_ij = -log(M_ij + M_ji)
(:396-410) are kernels (csrc/ofg_dendro.cuh); the Phylip files (:413-431) are written by the host.

Orthogroup members are anything with `.iSp` / `.iSeq` (the reference's Seq, orthologues.py:73-90) or (iSp, iSeq) tuples,
iSp being the species ID of SpeciesIDs.txt.  There is no CPU fallback: without libofg.so / a GPU this module raises.
"""
import numpy as np

from . import _lib
from .waterfall import WaterfallMethod


def _members(ogs, seqsInfo):
    pos = {k: i for i, k in enumerate(seqsInfo.speciesToUse)}
    og_off = np.zeros(len(ogs) + 1, np.int64)
    sp, sq = [], []
    for g, og in enumerate(ogs):
        for m in og:
            i_sp, i_seq = (m.iSp, m.iSeq) if hasattr(m, "iSp") else m
            sp.append(pos[int(i_sp)])
            sq.append(int(i_seq))
        og_off[g + 1] = len(sp)
    return og_off, np.array(sp, np.int32), np.array(sq, np.int32)


def _name(m):
    return m.ToString() if hasattr(m, "ToString") else "%d_%d" % (m[0], m[1])


class DendroBLASTTrees:
    """The distance-matrix half of orthologues.DendroBLASTTrees (:299-431)."""

    def __init__(self, seqsInfo, blastDir_list, qDoubleBlast=True, device=0, n_readers=None):
        self.seqsInfo = seqsInfo
        self.blastDir_list = blastDir_list
        self.qDoubleBlast = qDoubleBlast
        self.device = device
        self.n_readers = n_readers
        self.gb = None

    def _raw(self):
        """GetBLAST6Scores(..., qExcludeSelfHits=False) for every species pair (:272-274), resident on the device."""
        if self.gb is None:
            sp = list(self.seqsInfo.speciesToUse)
            ones = [np.ones(self.seqsInfo.nSeqsPerSpecies[k]) for k in sp]   # sequence lengths play no part here
            self.gb = WaterfallMethod.BuildGraph(self.seqsInfo, self.blastDir_list, ones, qDoubleBlast=self.qDoubleBlast,
                                                 device=self.device, n_readers=self.n_readers, stage=_lib.STAGE_RAW,
                                                 keep_self_hits=True)
        return self.gb

    def close(self):
        if self.gb is not None:
            self.gb.close()
            self.gb = None

    def GetOGMatrices_FullParallel(self, ogs_all):
        """:310-379: (ogs with >= 4 genes, their matrices M as n x n arrays)."""
        lengths = [len(og) for og in ogs_all]
        for iog in range(1, len(ogs_all)):
            assert lengths[iog] <= lengths[iog - 1], iog          # :331-333
        iog4 = next((i for i, n in enumerate(lengths) if n < 4), len(ogs_all))   # :334-337
        ogs = ogs_all[:iog4]
        og_off, sp, sq = _members(ogs, self.seqsInfo)
        flat, _ = self._raw().dendro_matrices(og_off, sp, sq, complete=False)
        return ogs, self._split(flat, og_off)

    @staticmethod
    def _split(flat, og_off):
        out, o = [], 0
        for n in np.diff(og_off):
            out.append(flat[o:o + n * n].reshape(n, n))
            o += n * n
        return out

    def CompleteOGMatrices(self, ogs):
        """:381-394 / :396-410 on the device: (D per orthogroup, max_og per orthogroup)."""
        og_off, sp, sq = _members(ogs, self.seqsInfo)
        flat, mx = self._raw().dendro_matrices(og_off, sp, sq, complete=True)
        return self._split(flat, og_off), mx

    def CompleteAndWriteOGMatrices(self, ogs, dist_fn):
        """:396-411: writes dist_fn(iog) (files.FileHandler.GetOGsDistMatFN) for every orthogroup, returns the matrices."""
        D, mx = self.CompleteOGMatrices(ogs)
        for iog, (og, d) in enumerate(zip(ogs, D)):
            self.WritePhylipMatrix(d, [_name(g) for g in og], dist_fn(iog), mx[iog])
        return D

    @staticmethod
    def WritePhylipMatrix(m, names, outFN, max_og):
        """:413-431."""
        with open(outFN, "wb") as f:
            f.write(phylip_text(m, names, max_og))


def phylip_text(m, names, max_og):
    """The bytes WritePhylipMatrix (:413-431) writes: -inf -> 1.1 * max_og, values in (0, 1e-6) -> 1e-6, "%.6f", no "-0"."""
    m = np.asarray(m, np.float64)
    n = m.shape[0]
    v = np.where(m > -9e99, m, 1.1 * max_og)
    v[np.arange(n), np.arange(n)] = 0.0
    v = 0.0 + v
    v = np.where((v > 0) & (v < 1e-6), 1e-6, v)
    lines = ["%d\n" % n]
    for i in range(n):
        lines.append(names[i] + " " + " ".join(["%.6f" % x for x in v[i]]) + "\n")
    return "".join(lines).encode()
