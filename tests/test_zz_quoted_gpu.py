"""GPU: hit files with csv quoting (missing item 7).  The reference reads hit files through csv.reader(delimiter='\\t')
(blast_file_processor.py:65-66), so quoted ids / scores and quoted columns holding tabs or line breaks are legal input; the device parser
refuses a quote, the reader threads of the host mirror rewrite such files into the plain rows the csv module yields (waterfall.csv_unquote).
Through the working-directory entry point the graph must be the live reference's."""
import gzip
import os

import pytest

from orthofinder_b200 import synth, waterfall
from tests.helpers import Golden, csv_quote_hit_text

pytestmark = pytest.mark.gpu


def test_working_directory_with_csv_quoted_hit_files(tmp_path):
    wd = str(tmp_path) + "/"
    synth.write_working_directory(wd, "C2_tiny")
    g = Golden("synth_C2_tiny")
    for k, fn in enumerate(sorted(f for f in os.listdir(wd) if f.startswith("Blast") and f.endswith(".txt"))):
        if k % 3 == 2:
            continue                                   # a third of the files stays as it is
        with open(wd + fn, "rb") as f:
            data = csv_quote_hit_text(f.read(), seed=k)
        if k % 3 == 0:
            with open(wd + fn, "wb") as f:
                f.write(data)
        else:                                          # quoted AND gzip-compressed (inflated by zlib in the reader threads)
            with open(wd + fn + ".gz", "wb") as f:
                f.write(gzip.compress(data, 1))
            os.remove(wd + fn)
    graph_fn, gb, si = waterfall.DoOrthogroupsGraph(wd, graphFN=wd + "OrthoFinder_graph.txt")
    with open(graph_fn, "rb") as f:
        assert f.read() == g.graph
    gb.close()
