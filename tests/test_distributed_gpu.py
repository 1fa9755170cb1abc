"""Multi-GPU product entry point (needs >= 2 CUDA devices; run with `gpurun --gpus 2`): waterfall.BuildGraphDistributed
over real Blast*.txt[.gz] files, one process per GPU, NCCL + peer inboxes - the written OrthoFinder_graph.txt must be the
single-GPU DoOrthogroupsGraph file (and the live reference's golden) byte for byte."""
import gzip
import os
import socket

import pytest

from tests.helpers import Golden

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _write_wd(g, wd, gz_every=3):
    S = len(g.n_seqs)
    with open(wd + "SpeciesIDs.txt", "w") as f:
        for k in g.species:
            f.write("%d: sp%d.fa\n" % (k, k))
    for p, k in enumerate(g.species):
        with open(wd + "Species%d.fa" % k, "w") as f:
            for q, L in enumerate(g.lengths[p]):
                f.write(">%d_%d\n%s\n" % (k, q, "A" * int(L)))
    n = 0
    for p, kp in enumerate(g.species):
        for j, kj in enumerate(g.species):
            data = g.hit_text(p, j)
            fn = wd + "Blast%d_%d.txt" % (kp, kj)
            if n % gz_every == 0:
                with open(fn + ".gz", "wb") as f:
                    f.write(gzip.compress(data, 1))
            else:
                with open(fn, "wb") as f:
                    f.write(data)
            n += 1


def _worker(rank, world, port, wd, out_fn):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from orthofinder_b200 import waterfall
    sp, n_all, _ = waterfall.GetSpeciesToUse(wd + "SpeciesIDs.txt")
    si = waterfall.GetSeqsInfo([wd], sp, n_all)
    L = waterfall.GetSequenceLengths(si, [wd])
    waterfall.BuildGraphDistributed(si, [wd], L, out_fn, device=rank)
    dist.destroy_process_group()


@pytest.mark.parametrize("name", ["ref_AddTwoSpecies", "synth_C5_small"])
def test_build_graph_distributed_writes_the_single_gpu_file(tmp_path, name):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from orthofinder_b200 import waterfall
    g = Golden(name)
    wd = str(tmp_path) + "/"
    _write_wd(g, wd)
    single, gb, si = waterfall.DoOrthogroupsGraph(wd, graphFN=wd + "single_graph.txt")
    gb.close()
    mp.spawn(_worker, args=(2, _free_port(), wd, wd + "dist_graph.txt"), nprocs=2, join=True)
    a, b = open(single, "rb").read(), open(wd + "dist_graph.txt", "rb").read()
    assert a == b
    assert a == g.graph
    assert not [f for f in os.listdir(wd) if f.startswith("dist_graph.txt_")]   # part files are removed
