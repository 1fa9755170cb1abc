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


def mcl_graph_text(n, cols):
    """MCL 'matrix' text exactly as WriteGraphParallel / WriteGraph_perSpecies lay it out (gathering.py:39-48, :279-280)."""
    out = ["(mclheader\nmcltype matrix\ndimensions %dx%d\n)\n" % (n, n), "\n(mclmatrix\nbegin\n\n"]
    for i in range(n):
        out.append("%d    " % i + "".join("%d:%.3f " % (j, w) for j, w in sorted(cols[i].items())) + "$\n")
    out.append(")\n")
    return "".join(out).encode()


def mcl_synthetic_graph(kind):
    """Seeded graphs for the MCL step beyond what the golden datasets exercise.
    'family': one gene family of 1 250 genes with ~28 hits each (columns of the expansion pass the selection number
    1 100 and the recovery number 1 400 of mcl's default scheme; the dense product walk), 60 four-gene cliques, genes
    without edges.  'ties': exactly symmetric structures (a gene between two equal triangles / two equal stars, cycles and
    paths with equal weights, weights that print as 0.000) - overlap cut, attractor systems of several genes."""
    rng = np.random.default_rng(20261020)
    if kind == "family":
        n, fam = 1600, 1250
        cols = [dict() for _ in range(n)]
        for i in range(fam):
            for j in rng.choice(fam, 28, replace=False):
                if i != j:
                    w = float(rng.uniform(0.05, 2.5))
                    cols[i][int(j)] = w
                    if rng.random() < 0.9:
                        cols[int(j)][i] = w * float(rng.uniform(0.8, 1.2))
        for i in range(fam, fam + 240, 4):
            for a in range(i, i + 4):
                for b in range(i, i + 4):
                    if a != b:
                        cols[a][b] = float(rng.uniform(0.3, 2.0))
        return mcl_graph_text(n, cols)
    if kind == "ties":
        cols, base = [], 0

        def block(k):
            nonlocal base
            cols.extend(dict() for _ in range(k))
            base += k
            return base - k

        def add(a, b, w):
            cols[a][b] = w
            cols[b][a] = w
        for rep in range(6):
            o = block(7)                       # gene o+3 between two equal triangles
            for tri in ((0, 1, 2), (4, 5, 6)):
                for a in tri:
                    for b in tri:
                        if a < b:
                            add(o + a, o + b, 1.0)
            for k in (0, 1, 2, 4, 5, 6):
                add(o + 3, o + k, 0.2 + 0.1 * rep)
            o = block(10)                      # gene o+3 between two equal stars (hubs o+5 and o+1), o+4 isolated
            for leaf in (0, 6, 7):
                add(o + 5, o + leaf, 1.0)
            for leaf in (2, 8, 9):
                add(o + 1, o + leaf, 1.0)
            add(o + 3, o + 5, 0.3)
            add(o + 3, o + 1, 0.3)
            o = block(5 + rep)                 # cycle with equal weights
            for a in range(5 + rep):
                add(o + a, o + (a + 1) % (5 + rep), 0.75)
            o = block(4 + rep)                 # path with equal weights
            for a in range(3 + rep):
                add(o + a, o + a + 1, 1.5)
            o = block(3)                       # an edge that prints as 0.000 (dropped when mcl reads it) and a one-way edge
            cols[o][o + 1] = 0.0004
            cols[o + 1][o + 2] = 0.8
        block(3)
        return mcl_graph_text(base, cols)
    raise KeyError(kind)


def matrix_digest(colptr, rows, vals):
    """SHA-256 over a column-compressed float matrix (int64 pointers, int32 rows, float32 bit patterns)."""
    import hashlib
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(colptr, np.int64).tobytes())
    h.update(np.ascontiguousarray(rows, np.int32).tobytes())
    h.update(np.ascontiguousarray(vals, np.float32).tobytes())
    return h.hexdigest()


class MclGolden:
    """What the reference's bundled mcl binary produced for one golden graph (tools/make_golden_mcl.py)."""

    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN_DIR, "mcl_%s.npz" % name))
        self.inflations = [float(x) for x in z["inflations"]]
        self._z = z

    def _tag(self, inflation):
        return "I%s" % str(float(inflation)).replace(".", "_")

    def clusters_body(self, inflation):
        return zlib.decompress(self._z[self._tag(inflation) + "_clusters_z"].tobytes()).decode()

    def iterations(self, inflation):
        return int(self._z[self._tag(inflation) + "_iterations"])

    def digests(self, inflation):
        return [str(x) for x in self._z[self._tag(inflation) + "_sha256"]], [int(x) for x in self._z[self._tag(inflation) + "_nnz"]]


MCL_GOLDENS = REF_GOLDENS + SYNTH_GOLDENS + ["synth_family", "synth_ties"]


def mcl_golden_graph(name):
    return mcl_synthetic_graph(name[len("synth_"):]) if name in ("synth_family", "synth_ties") else Golden(name).graph


def csv_quote_hit_text(text, seed=3):
    """Rewrites plain hit rows with csv quoting that csv.reader(delimiter='\\t') undoes (blast_file_processor.py:65-66): quoted ids and
    scores, an ignored column holding a tab, a line break or a doubled quote inside quotes, \\r\\n line ends.  Same rows after parsing."""
    rng = np.random.default_rng(seed)
    out = []
    for k, line in enumerate(bytes(text).split(b"\n")):
        if not line:
            continue
        f = line.split(b"\t")
        kind = int(rng.integers(0, 6))
        if kind == 0:
            f[0] = b'"' + f[0] + b'"'
        elif kind == 1:
            f[1] = b'"' + f[1] + b'"'
            f[11] = b'"' + f[11] + b'"'
        elif kind == 2:
            f[4] = b'"a\tb"'                     # a tab inside an ignored column
        elif kind == 3:
            f[5] = b'"x""y"'                     # doubled quote
        elif kind == 4:
            f[6] = b'"two\nlines"'               # a line break inside an ignored column
        out.append(b"\t".join(f) + (b"\r\n" if k % 3 == 0 else b"\n"))
    return b"".join(out)
