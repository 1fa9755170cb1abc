"""Golden-vector loading and comparison helpers shared by the CPU and GPU tests."""
import os
import zlib

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
REF_GOLDENS = ["ref_AddTwoSpecies", "ref_FromTrees", "ref_removeFirst"]
SYNTH_GOLDENS = ["synth_C2_tiny", "synth_C5_small"]


class Golden:
    """Inputs (n_seqs, lengths, hit text per pair) and the live reference's outputs for one dataset."""

    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        self.name = name
        self.graph = zlib.decompress(z["graph_z"].tobytes())
        self.graph_v2 = zlib.decompress(z["graph_v2_z"].tobytes()) if "graph_v2_z" in z.files else None
        self.most_distant_v2 = z["most_distant_v2"] if "most_distant_v2" in z.files else None
        self.graph_hom = zlib.decompress(z["graph_hom_z"].tobytes()) if "graph_hom_z" in z.files else None
        self.fits = z["fits"]
        self.nnz = z["nnz"]
        self.n_seqs = [int(x) for x in z["n_seqs"]]
        self.species = [int(x) for x in z["species"]]
        self.most_distant = z["most_distant"]
        self.provenance = str(z["provenance"])
        S = len(self.n_seqs)
        if "hits_z" in z.files:
            blob = zlib.decompress(z["hits_z"].tobytes())
            off = np.concatenate(([0], np.cumsum(z["hit_sizes"])))
            self.texts = {(p, j): blob[off[p * S + j]:off[p * S + j + 1]] for p in range(S) for j in range(S)}
            L = z["lengths"]
            o = np.concatenate(([0], np.cumsum(self.n_seqs)))
            self.lengths = [L[o[p]:o[p + 1]].copy() for p in range(S)]
            self.synth = None
        else:
            from orthofinder_b200 import synth
            cfg = name[len("synth_"):]
            ids, n_seqs, L, P = synth.config_inputs(cfg)
            assert n_seqs == self.n_seqs
            self.lengths = L
            self.synth = (P, ids)
            lib = synth.host_lib()
            self.texts = {}
            for p in range(S):
                for j in range(S):
                    self.texts[(p, j)] = synth.pair_text(lib, P, p, j, ids[p], ids[j], L[p], L[j])[0]

    def hit_text(self, p, j):
        return self.texts[(p, j)]


class DendroGolden:
    """Orthogroups and the live reference's DendroBLAST matrices / Phylip files for one dataset
    (tools/make_golden_dendro.py; orthologues.py:264-291, 396-431)."""

    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN_DIR, "dendro_" + name + ".npz"))
        self.og_off = z["og_off"]
        self.og_sp = z["og_sp"]          # species list positions
        self.og_seq = z["og_seq"]
        self.ogs = [[(int(self.og_sp[k]), int(self.og_seq[k])) for k in range(self.og_off[g], self.og_off[g + 1])]
                    for g in range(self.og_off.size - 1)]
        n = np.diff(self.og_off)
        eo = np.concatenate(([0], np.cumsum(n * n)))
        self.M = [z["M"][eo[g]:eo[g + 1]].reshape(n[g], n[g]) for g in range(n.size)]
        self.D = [z["D"][eo[g]:eo[g + 1]].reshape(n[g], n[g]) for g in range(n.size)]
        blob = zlib.decompress(z["phy_z"].tobytes())
        po = np.concatenate(([0], np.cumsum(z["phy_sizes"])))
        self.phylip = [blob[po[g]:po[g + 1]] for g in range(n.size)]
        self.provenance = str(z["provenance"])


def parse_graph_text(graph):
    """OrthoFinder_graph.txt -> (n, list of (row, [(col, 'w.www')...]))."""
    lines = graph.split(b"\n")
    assert lines[0] == b"(mclheader" and lines[1] == b"mcltype matrix"
    n = int(lines[2].split()[1].split(b"x")[0])
    i = lines.index(b"begin") + 2
    rows = []
    while lines[i] != b")":
        parts = lines[i].split()
        assert parts[-1] == b"$"
        rows.append((int(parts[0]), [tuple(x.split(b":")) for x in parts[1:-1]]))
        i += 1
    return n, rows


def graph_text_diff(a, b):
    """(rows with different edge sets, entries whose third decimal differs) between two graph texts."""
    na, ra = parse_graph_text(a)
    nb, rb = parse_graph_text(b)
    assert na == nb and len(ra) == len(rb)
    edge_diff = dec_diff = 0
    for (r1, e1), (r2, e2) in zip(ra, rb):
        assert r1 == r2
        if [c for c, _ in e1] != [c for c, _ in e2]:
            edge_diff += 1
        else:
            dec_diff += sum(w1 != w2 for (_, w1), (_, w2) in zip(e1, e2))
    return edge_diff, dec_diff
