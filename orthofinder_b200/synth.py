"""Synthetic WorkingDirectories (SpeciesIDs.txt, Species{k}.fa, Blast{i}_{j}.txt) — SURVEY.md §8(d).

The hit lines come from csrc/synth_hits.h (one source, compiled for host here and for the device in
libofg.so, byte-identical).  Gene lengths are drawn on the host only (numpy) and handed to both.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


class SynthParams(ctypes.Structure):
    _fields_ = [("seed", ctypes.c_uint64), ("hq", ctypes.c_int32), ("integer_scores", ctypes.c_int32),
                ("ident_levels", ctypes.c_int32), ("p_dup_permille", ctypes.c_int32),
                ("p_missing_permille", ctypes.c_int32), ("p_nohit_permille", ctypes.c_int32),
                ("p_selfonly_permille", ctypes.c_int32), ("p_twin_permille", ctypes.c_int32),
                ("single_i", ctypes.c_int32), ("single_j", ctypes.c_int32)]


def params(seed=1, hq=50, integer_scores=0, ident_levels=0, p_dup=100, p_missing=0, p_nohit=0, p_selfonly=0,
           p_twin=0, single=(-1, -1)):
    return SynthParams(seed, hq, integer_scores, ident_levels, p_dup, p_missing, p_nohit, p_selfonly, p_twin,
                       single[0], single[1])


# named configurations of BASELINE.json (C2..C5) and their small stand-ins used by the parity tests
CONFIGS = {
    "C2": dict(S=16, G=20000, p=dict(seed=1, hq=50)),
    "C2_small": dict(S=16, G=2000, p=dict(seed=1, hq=50)),
    "C2_tiny": dict(S=6, G=400, p=dict(seed=1, hq=24)),
    "C3": dict(S=64, G=20000, p=dict(seed=2, hq=50)),
    "C4": dict(S=256, G=20000, p=dict(seed=3, hq=8)),
    "C4_tiny": dict(S=48, G=60, p=dict(seed=3, hq=6)),       # many species pairs (2304), few genes
    "C5": dict(S=6, G=5000, n_lengths=0, p=dict(seed=5, hq=12, integer_scores=1, ident_levels=16, p_missing=300,
                                                p_nohit=50, p_selfonly=20, p_twin=50, single=(5, 4))),
    "C5_small": dict(S=5, G=600, p=dict(seed=5, hq=12, integer_scores=1, ident_levels=16, p_missing=300,
                                         p_nohit=50, p_selfonly=20, p_twin=80, single=(4, 3))),
}


def host_lib():
    so = os.path.join(_HERE, "libofg_synth.so")
    src = [os.path.join(_HERE, "csrc", f) for f in ("synth_host.cpp", "synth_hits.h")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src):
        subprocess.check_call(["g++", "-O2", "-fPIC", "-shared", src[0], "-o", so])
    lib = ctypes.CDLL(so)
    lib.ofg_synth_pair_host.restype = ctypes.c_int64
    return lib


def make_lengths(S, G, seed, twins=False, n_distinct=0):
    """clip(round(lognormal(5.8, 0.6)), 30, 5000) per gene, one Philox stream per species.
    twins: odd genes of long proteins get the even neighbour's length +1 / -1 alternately (near-tie stress).
    n_distinct > 0: lengths quantised to that many distinct values (tie-heavy stress)."""
    out = []
    for k in range(S):
        rng = np.random.Generator(np.random.Philox(key=seed * 1000003 + k))
        L = np.clip(np.round(rng.lognormal(5.8, 0.6, G)), 30, 5000)
        if n_distinct:
            grid = np.round(np.exp(np.linspace(np.log(60), np.log(3000), n_distinct)))
            L = grid[np.abs(np.log(L)[:, None] - np.log(grid)[None, :]).argmin(1)]
        if twins:
            even = L[0:G - 1:2]
            odd = L[1:G:2]
            sign = np.where(np.arange(len(odd)) % 2 == 0, 1.0, -1.0)[:len(odd)]
            L[1:G:2] = np.where(even[:len(odd)] >= 800, even[:len(odd)] + sign, odd)
        out.append(L.astype(np.float64))
    return out


def pair_text(lib, P, pi, pj, si, sj, Li, Lj):
    args = (ctypes.byref(P), ctypes.c_int(pi), ctypes.c_int(pj), ctypes.c_int(si), ctypes.c_int(sj),
            ctypes.c_int64(len(Li)), ctypes.c_int64(len(Lj)), Li.ctypes.data_as(ctypes.c_void_p),
            Lj.ctypes.data_as(ctypes.c_void_p))
    nl = ctypes.c_int64(0)
    nb = lib.ofg_synth_pair_host(*args, None, ctypes.c_int64(0), ctypes.byref(nl))
    buf = ctypes.create_string_buffer(max(int(nb), 1))
    nb2 = lib.ofg_synth_pair_host(*args, buf, ctypes.c_int64(nb), ctypes.byref(nl))
    assert nb2 == nb
    return buf.raw[:nb], int(nl.value)


def config_inputs(name):
    """(species_ids, n_seqs, lengths, SynthParams) of a named configuration."""
    c = CONFIGS[name]
    P = params(**c["p"])
    L = make_lengths(c["S"], c["G"], P.seed, twins=P.p_twin_permille > 0,
                     n_distinct=12 if P.ident_levels else 0)
    return list(range(c["S"])), [c["G"]] * c["S"], L, P


def write_working_directory(wd, name=None, S=None, G=None, P=None, lengths=None, species_ids=None,
                            skip_species=()):
    """Writes a WorkingDirectory the reference accepts.  Returns (total_lines, total_bytes)."""
    if name is not None:
        species_ids, n_seqs, lengths, P = config_inputs(name)
        S = len(species_ids)
    else:
        species_ids = list(range(S)) if species_ids is None else list(species_ids)
        if lengths is None:
            lengths = make_lengths(S, G, P.seed, twins=P.p_twin_permille > 0)
    os.makedirs(wd, exist_ok=True)
    lib = host_lib()
    with open(os.path.join(wd, "SpeciesIDs.txt"), "w") as f:
        for k in species_ids:
            f.write("%s%d: species_%d.fa\n" % ("#" if k in skip_species else "", k, k))
    with open(os.path.join(wd, "SequenceIDs.txt"), "w") as f:
        for p, k in enumerate(species_ids):
            for g in range(len(lengths[p])):
                f.write("%d_%d: sp%d_gene%d\n" % (k, g, k, g))
    for p, k in enumerate(species_ids):
        with open(os.path.join(wd, "Species%d.fa" % k), "w") as f:
            for g, L in enumerate(lengths[p]):
                f.write(">%d_%d\n" % (k, g))
                seq = "A" * int(L)
                for o in range(0, len(seq), 60):
                    f.write(seq[o:o + 60] + "\n")
    lines = nbytes = 0
    for p, k in enumerate(species_ids):
        for j, kj in enumerate(species_ids):
            txt, nl = pair_text(lib, P, p, j, k, kj, lengths[p], lengths[j])
            with open(os.path.join(wd, "Blast%d_%d.txt" % (k, kj)), "wb") as f:
                f.write(txt)
            lines += nl
            nbytes += len(txt)
    return lines, nbytes
