"""Host-side ingestion of the hit files (no GPU): file lookup, gzip members, ordered parallel read-ahead."""
import gzip
import os

import pytest

from orthofinder_b200 import waterfall as w


def _line(i, j, k):
    return b"%d_1\t%d_2\t1\t1\t1\t1\t1\t1\t1\t1\t1e-5\t%d\n" % (i, j, k)


def test_read_hit_files_keeps_order_and_inflates_all_gzip_members(tmp_path):
    d = str(tmp_path) + "/"
    texts = {}
    for i in range(3):
        for j in range(3):
            t = _line(i, j, 10 * i + j) * (500 * (i + j + 1))
            texts[(i, j)] = t
            if (i + j) % 2:  # two concatenated gzip members, as `cat a.gz b.gz` produces
                with open(d + "Blast%d_%d.txt.gz" % (i, j), "wb") as f:
                    f.write(gzip.compress(t[:len(t) // 2]) + gzip.compress(t[len(t) // 2:]))
            else:
                with open(d + "Blast%d_%d.txt" % (i, j), "wb") as f:
                    f.write(t)
    req = [(i, j, "tag%d%d" % (i, j)) for i in range(3) for j in range(3)]
    for readers in (1, 4):
        got = list(w.read_hit_files([d], req, n_readers=readers, window=3))
        assert [g[0] for g in got] == req
        assert all(g[1] == texts[(g[0][0], g[0][1])] for g in got)
        assert all(os.path.basename(g[2]) == "Blast%d_%d.txt" % g[0][:2] for g in got)


def test_missing_file_is_an_error_unless_allowed(tmp_path):
    d = str(tmp_path) + "/"
    with open(d + "Blast0_0.txt", "wb") as f:
        f.write(_line(0, 0, 5))
    with pytest.raises(FileNotFoundError):
        list(w.read_hit_files([d], [(0, 0), (0, 1)], n_readers=2))
    got = list(w.read_hit_files([d], [(0, 0), (0, 1)], q_allow_empty=True, n_readers=2))
    assert got[0][1] == _line(0, 0, 5) and got[1][1] is None   # blast_file_processor.py:62-63


def test_first_directory_holding_the_file_wins(tmp_path):
    a, b = str(tmp_path / "a") + "/", str(tmp_path / "b") + "/"
    os.mkdir(a)
    os.mkdir(b)
    with open(b + "Blast1_2.txt", "wb") as f:
        f.write(b"from b\n")
    with open(a + "Blast3_4.txt.gz", "wb") as f:
        f.write(gzip.compress(b"from a\n"))
    with open(b + "Blast3_4.txt", "wb") as f:
        f.write(b"shadowed\n")
    got = {g[0]: g[1] for g in w.read_hit_files([a, b], [(1, 2), (3, 4)], n_readers=2)}
    assert got[(1, 2)] == b"from b\n" and got[(3, 4)] == b"from a\n"   # blast_file_processor.py:59-65


def test_truncated_gzip_raises_like_gzip_open(tmp_path):
    """blast_file_processor.py:63 reads .gz files through gzip.open, which raises EOFError when the stream ends before its
    end-of-stream marker (interrupted DIAMOND run, full disk); the host mirror must not hand the partial text to the parser."""
    d = str(tmp_path) + "/"
    data = b"".join(_line(0, 1, k) for k in range(2000))
    z = gzip.compress(data)
    with open(d + "Blast0_1.txt.gz", "wb") as f:
        f.write(z[:len(z) // 2])
    with pytest.raises(EOFError):
        gzip.decompress(z[:len(z) // 2])   # what the reference's reader does with this file
    with pytest.raises(EOFError) as ei:
        list(w.read_hit_files([d], [(0, 1)], n_readers=1))
    assert ei.value.hit_file[:2] == (0, 1) and 0 < len(ei.value.partial) < len(data)
    with pytest.raises(EOFError):
        list(w.read_hit_files([d], [(0, 1), (0, 1)], n_readers=2))


def test_size_hint_detection_of_gzip_layouts():
    """BuildGraph(device_gunzip="auto") sends a .gz file to the device only if its members can be located from their headers
    (bgzip / BGZF "BC" extra field); everything else is inflated by zlib in the reader threads."""
    import gzip
    from orthofinder_b200 import waterfall
    data = b"0_0\t1_1\t1\t1\t1\t1\t1\t1\t1\t1\t1e-5\t50\n" * 5000
    assert waterfall.gz_has_size_hints(waterfall.gzip_members(data, chunk=30000))
    assert not waterfall.gz_has_size_hints(gzip.compress(data, 1))
    assert not waterfall.gz_has_size_hints(b"")
    assert not waterfall.gz_has_size_hints(data)
    # another extra subfield in front of the size hint, and an extra field without one
    hinted = bytearray(waterfall.gzip_members(data[:100]))
    other = b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x0c\x00XY\x02\x00\x00\x00" + bytes(hinted[12:18]) + bytes(hinted[18:])
    assert waterfall.gz_has_size_hints(other)
    nohint = b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00XY\x02\x00\x00\x00" + bytes(hinted[18:])
    assert not waterfall.gz_has_size_hints(nohint)
    assert gzip.decompress(waterfall.gzip_members(data, chunk=30000)) == data


def test_phylip_writer_matches_the_reference_bytes():
    """dendroblast.phylip_text (host formatting of orthologues.py:413-431 WritePhylipMatrix) on the live reference's completed
    matrices: -inf -> 1.1 * max_og, values in (0, 1e-6) -> 1e-6, "%.6f", zero diagonal, no "-0"."""
    import numpy as np
    from orthofinder_b200 import dendroblast
    from tests.helpers import Golden, DendroGolden
    for name in ("ref_AddTwoSpecies", "synth_C5_small"):
        g, dg = Golden(name), DendroGolden(name)
        n_inf = 0
        for iog, og in enumerate(dg.ogs):
            d = dg.D[iog]
            low = d[np.tril_indices(d.shape[0], -1)]
            max_og = max(-9e99, low.max())
            names = ["%d_%d" % (g.species[p], q) for p, q in og]
            assert dendroblast.phylip_text(d, names, max_og) == dg.phylip[iog], (name, iog)
            n_inf += int(np.isinf(d).sum())
        assert name != "synth_C5_small" or n_inf > 0
    tiny = np.array([[0.5, 2e-7, -0.0], [2e-7, 0.1, -np.inf], [-0.0, -np.inf, 0.3]])
    assert dendroblast.phylip_text(tiny, ["a", "b", "c"], 2.0) == b"3\na 0.000000 0.000001 0.000000\nb 0.000001 0.000000 2.200000\nc 0.000000 2.200000 0.000000\n"


def test_csv_quoted_files_become_the_rows_the_csv_module_yields(tmp_path):
    """Files with csv quoting (accepted by the reference's csv.reader, refused by the device parser) are rewritten by the reader
    threads: the oracle's matrices of the rewritten text == those of the original rows; unquoted files pass through untouched."""
    import numpy as np
    from oracle import waterfall_oracle as wo
    from tests.helpers import Golden, csv_quote_hit_text
    g = Golden("ref_removeFirst")
    d = str(tmp_path) + "/"
    S = len(g.n_seqs)
    for (p, j), t in g.texts.items():
        with open(d + "Blast%d_%d.txt" % (p, j), "wb") as f:
            f.write(csv_quote_hit_text(t, seed=p * S + j) if (p + j) % 2 == 0 else t)
    for (p, j, *_), data, fn in w.read_hit_files([d], [(p, j) for p in range(S) for j in range(S)], n_readers=3):
        if (p + j) % 2:
            assert data == g.texts[(p, j)]
            continue
        assert b'"' not in data and b"x'y" in data              # the literal quote of the doubled-quote column (nobody reads it) is gone
        a = wo.get_blast6_scores(g.texts[(p, j)], g.n_seqs[p], g.n_seqs[j], p == j)
        b = wo.get_blast6_scores(data, g.n_seqs[p], g.n_seqs[j], p == j)
        assert all(np.array_equal(x, y) for x, y in zip(a, b))
    assert w.csv_unquote(b'"0_1"\t"0_2"\t' + b"\t".join([b"1"] * 9) + b'\t" 57.5"\r\n\r\n') == b"0_1\t0_2\t" + b"\t".join([b"1"] * 9) + b"\t 57.5\n"


@pytest.mark.reference
def test_csv_quoted_file_through_the_live_reference(tmp_path):
    """The reference's own GetBLAST6Scores on the quoted file == on the original file (so rewriting the rows is the right thing)."""
    import numpy as np
    from oracle import ref_harness as rh
    if not rh.reference_available():
        pytest.skip("live reference not present (GPU box)")
    from tests.helpers import Golden, csv_quote_hit_text
    gathering, util, files, matrices = rh._import_reference_real()
    from scripts_of import blast_file_processor as bfp
    g = Golden("ref_removeFirst")
    S = len(g.n_seqs)
    starts = np.concatenate(([0], np.cumsum(g.n_seqs)[:-1]))
    si = util.SequencesInfo(nSeqs=sum(g.n_seqs), nSpecies=S, speciesToUse=list(range(S)), seqStartingIndices=[int(x) for x in starts],
                            nSeqsPerSpecies={k: n for k, n in enumerate(g.n_seqs)})
    plain, quoted = str(tmp_path / "plain") + "/", str(tmp_path / "quoted") + "/"
    os.makedirs(plain), os.makedirs(quoted)
    for (p, j) in ((0, 1), (1, 1)):
        with open(plain + "Blast%d_%d.txt" % (p, j), "wb") as f:
            f.write(g.texts[(p, j)])
        with open(quoted + "Blast%d_%d.txt" % (p, j), "wb") as f:
            f.write(csv_quote_hit_text(g.texts[(p, j)], seed=7))
        A = bfp.GetBLAST6Scores(si, [plain], p, j)
        B = bfp.GetBLAST6Scores(si, [quoted], p, j)
        assert (A != B).nnz == 0 and A.nnz > 0
        mine = w.csv_unquote(open(quoted + "Blast%d_%d.txt" % (p, j), "rb").read())
        with open(quoted + "Blast%d_%d.txt" % (p, j), "wb") as f:
            f.write(mine)
        assert (bfp.GetBLAST6Scores(si, [quoted], p, j) != A).nnz == 0


def test_quoted_working_directory_gives_the_reference_graph_through_the_oracle(tmp_path):
    """The host half of tests/test_zz_quoted_gpu.py on the CPU: what the reader threads hand to the device for a working directory
    with quoted (and quoted + gzip-compressed) hit files parses - by the oracle - into the live reference's graph."""
    from oracle import waterfall_oracle as wo
    from orthofinder_b200 import synth
    from tests.helpers import Golden, csv_quote_hit_text
    wd = str(tmp_path) + "/"
    synth.write_working_directory(wd, "C2_tiny")
    g = Golden("synth_C2_tiny")
    for k, fn in enumerate(sorted(f for f in os.listdir(wd) if f.startswith("Blast") and f.endswith(".txt"))):
        if k % 3 == 2:
            continue
        with open(wd + fn, "rb") as f:
            data = csv_quote_hit_text(f.read(), seed=k)
        if k % 3 == 0:
            with open(wd + fn, "wb") as f:
                f.write(data)
        else:
            with open(wd + fn + ".gz", "wb") as f:
                f.write(gzip.compress(data, 1))
            os.remove(wd + fn)
    S = len(g.n_seqs)
    texts = {}
    n_rewritten = 0
    for (p, j), data, fn in w.read_hit_files([wd], [(p, j) for p in range(S) for j in range(S)], n_readers=4, raw_gz="auto"):
        assert not isinstance(data, w.GzBytes) and b'"' not in data
        n_rewritten += b"x'y" in data
        texts[(p, j)] = data
    assert n_rewritten >= S * S // 2
    ref = wo.waterfall(g.n_seqs, g.lengths, lambda p, j: texts[(p, j)])
    assert wo.graph_text(ref["indptr"], ref["indices"], ref["data"], sum(g.n_seqs)) == g.graph
