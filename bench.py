#!/usr/bin/env python
"""bench.py — hit lines -> MCL graph throughput of the B200 path (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...  # the reference's algorithm on the host cores

A step is one full pass of the hot path (parse -> rows -> length sort -> select -> fit -> scale/BH ->
thresholds -> connect -> emit) over one batch of synthetic hit files.  N=1 runs BASELINE.json configs[1]
("C2": 16 species x 20k genes, ~50 hits/query/pair, ~282 M lines, 15 GB of text).  `value` is whole-job
lines/s with the text resident in HBM; `e2e` is the same job through the public host API with pinned HOST
buffers (H2D of every file + D2H of the graph text inside the timed region).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "hit_lines_per_sec_to_mcl_graph"
UNIT = "lines/s"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ---------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.stop = False
        self.t = None

    def _run(self):
        while not self.stop:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def __enter__(self):
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.t.join(timeout=6)

    def summary(self):
        sm, mx, reasons = [], 0, set()
        for s in self.samples:
            try:
                sm.append(float(s[0]))
                mx = max(mx, float(s[1]))
            except Exception:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], s[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ---------------------------------------------------------------------------------------------- workload
# species per GPU count when --species is not given: N = 1 is BASELINE.json configs[1] (C2, the configuration the metric
# is quoted on), N = 8 is configs[2] (C3: 64 species, 8 query species = 512 species pairs per GPU); N = 2 and 4 keep
# about the same 512 pairs per GPU (weak scaling towards C3)
DEFAULT_SPECIES = {2: 32, 4: 44, 8: 64}


def workload(args, n_gpus, species=None):
    from orthofinder_b200 import synth
    cfg = dict(synth.CONFIGS[args.config])
    S = species or args.species or cfg["S"]
    G = args.genes or cfg["G"]
    pk = dict(cfg["p"])
    if args.hq:
        pk["hq"] = args.hq
    if n_gpus > 1 and not args.species and not species:
        S = DEFAULT_SPECIES.get(n_gpus) or n_gpus * max(1, int(cfg["S"] * n_gpus ** 0.5 / n_gpus))
        if S == synth.CONFIGS["C3"]["S"] and args.config == "C2":
            pk = dict(synth.CONFIGS["C3"]["p"])   # C3's own seed
            if args.hq:
                pk["hq"] = args.hq
    P = synth.params(**pk)
    L = synth.make_lengths(S, G, P.seed, twins=P.p_twin_permille > 0, n_distinct=12 if P.ident_levels else 0)
    return S, G, P, L


def peaks():
    fn = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(fn):
        d = json.load(open(fn))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# algorithmic DRAM bytes per launch of each kernel, given the job's counts
def kernel_alg_bytes(name, c, text_bytes, launches):
    lines, rec, nnz = c["lines"], c["records"], c["nnz"]
    table = {
        "parse_kernel": text_bytes + 16 * lines,          # read text once, write (row, col, score)
        "scatter_rows_kernel": 16 * lines + 12 * rec,     # read COO, write (col, score) into row buckets
        "rowsort_kernel": 12 * rec + 16 * nnz,            # read rows, write sorted CSR + the Lf sort keys
        "radix_hist_kernel": 4 * rec,                     # keys only
        "radix_scatter_kernel": 24 * rec,                 # read + write (Lf, score)
        "bin_cut_kernel": 8 * nnz,                        # scores of every bin
        "bin_select_kernel": 32 * 0.05 * nnz,             # ~5 % of the hits: staged (lx, ly) read + written
        "scale_bh_kernel": 12 * nnz + 12 * nnz,           # read (col, score), write score + flag word
        "rbh_kernel": 12 * nnz,
        "connect_kernel": 12 * nnz,
        "emit_kernel<true>": 12 * nnz,
    }
    v = table.get(name)
    return None if v is None else v / max(launches, 1)   # per launch (waves split a kernel's work over several launches)


def text_hash(buf):
    import hashlib
    return hashlib.sha256(memoryview(buf)).hexdigest()[:16]


# ---------------------------------------------------------------------------------------------- parity
def parity_single(gb, S, G, L, cores):
    """N = 1: the CPU oracle (pinned byte for byte against the live reference) over the WHOLE workload that was timed,
    compared with the device graph: indices exact, weights within 1e-9, fits bit-equal, text byte-identical."""
    import numpy as np
    from oracle import parallel, waterfall_oracle as wo
    from orthofinder_b200 import waterfall
    t0 = time.perf_counter()
    texts = {(p, j): gb.hits_text(p, j) for p in range(S) for j in range(S)}
    n_seqs = [G] * S
    ref = parallel.waterfall_parallel(n_seqs, L, texts, cores, want_text=True)
    t_oracle = time.perf_counter() - t0
    rows, indptr, idx, dat = gb.graph_csr()
    out = {"checked": "whole workload against oracle/ (CPU restatement, %d processes, %.1f s)" % (cores, t_oracle),
           "lines": int(sum(t.count(b"\n") for t in texts.values()))}
    out["indptr_equal"] = bool(np.array_equal(indptr, ref["indptr"]))
    out["indices_equal"] = bool(np.array_equal(idx, ref["indices"]))
    out["edges"] = int(idx.size)
    out["max_rel_weight_err"] = float(np.max(np.abs(dat - ref["data"]) / ref["data"])) if (dat.size and dat.size == ref["data"].size) else None
    nf = nbad = 0
    for (p, j), f in ref["fits"].items():
        g = gb.pair_fit(p, j)
        if f is None:
            nbad += int(g["normalised"])
        else:
            nf += 1
            nbad += int(not (g["normalised"] and g["a"] == f[0] and g["b"] == f[1] and g["m"] == f[2]))
    out["fits_bit_equal"] = "%d/%d" % (nf + (0 if nbad == 0 else -nbad), nf)
    text = waterfall.graph_header(sum(n_seqs)) + gb.graph_text_body().tobytes() + waterfall.GRAPH_FOOTER
    out["text_bytes"] = len(text)
    out["text_identical"] = bool(text == ref["graph"])
    out["ok"] = bool(out["indptr_equal"] and out["indices_equal"] and nbad == 0 and out["text_identical"]
                     and out["max_rel_weight_err"] is not None and out["max_rel_weight_err"] <= 1e-9)
    return out


def parity_multi(args, world, rank, local, make_step):
    """N > 1: a down-scaled job (2 N species x 1500 genes) through the SAME multi-rank path (partition, peer exchange over
    NVLink, per-rank emit); rank 0 gathers the blocks, stitches them and compares with the CPU oracle byte for byte."""
    import numpy as np
    import torch.distributed as dist
    from orthofinder_b200 import synth, waterfall, sharding, _lib
    S, G = 2 * world, 1500
    P = synth.params(seed=11, hq=24)
    L = synth.make_lengths(S, G, P.seed)
    gb = waterfall.GraphBuilder(local)
    gb.set_species([G] * S, L)
    parts = sharding.partition_rows(S, world)
    gb.set_partition(parts[rank], args.exchange == "thresholds")
    gb.synth_hits(P, list(range(S)))
    step, closer = make_step(gb, S, [G] * S, parts)
    step()
    gb.sync()
    body = gb.graph_text_body().tobytes()
    blocks = [None] * world if rank == 0 else None
    dist.gather_object((parts[rank], body), blocks, dst=0)
    out = None
    if rank == 0:
        from oracle import parallel
        lib = synth.host_lib()
        texts = {(p, j): synth.pair_text(lib, P, p, j, p, j, L[p], L[j])[0] for p in range(S) for j in range(S)}
        ref = parallel.waterfall_parallel([G] * S, L, texts, None, want_text=True)
        # stitch: every rank's block holds the rows of its owned species in list order; species are contiguous row ranges
        per_species = {}
        for owned, blk in blocks:
            lines = blk.split(b"\n")[:-1]
            assert len(lines) == G * len(owned)
            for k, p in enumerate(owned):
                per_species[p] = b"\n".join(lines[k * G:(k + 1) * G]) + b"\n"
        text = waterfall.graph_header(S * G) + b"".join(per_species[p] for p in range(S)) + waterfall.GRAPH_FOOTER
        out = {"checked": "%d species x %d genes (%d lines) through the %d-rank path against oracle/" % (
            S, G, sum(t.count(b"\n") for t in texts.values()), world),
            "text_identical": bool(text == ref["graph"]), "edges": int(ref["indices"].size)}
        out["ok"] = out["text_identical"]
    closer()
    gb.close()
    return out


# ---------------------------------------------------------------------------------------------- ours
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from orthofinder_b200 import waterfall, _lib, sharding

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus > 1 and world != args.gpus:
        raise SystemExit("launch with torchrun --nproc-per-node %d (WORLD_SIZE=%d)" % (args.gpus, world))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    S, G, P, L = workload(args, world)
    ids = list(range(S))
    mode = args.exchange

    def make_step(gb, S_, n_seqs_, parts_):
        """step() enqueues one whole job on the builder's stream; closer() releases the exchange buffers."""
        owner = sharding.species_owner(parts_, S_)
        if world == 1:
            return (lambda: gb.build(_lib.STAGE_GRAPH)), (lambda: None)
        if mode == "peer":
            # design B: rows only; the transposed best-hit / connect lists are stored straight into the owners' inboxes
            # over NVLink by the export kernel (CUDA IPC mapped peer memory), one 4-byte all-reduce orders the imports
            gb.set_option("async", 1)
            px = sharding.PeerExchange(gb, owner, rank, world, local, n_seqs_)

            def step():
                gb.build(_lib.STAGE_NORM)
                px.exchange(0)
                gb.build(_lib.STAGE_CONNECT)
                px.exchange(1)
                gb.build(_lib.STAGE_GRAPH)
            return step, px.close
        if mode == "lists":
            def step():
                gb.build(_lib.STAGE_NORM)
                sharding.exchange_lists(gb, 0, owner, rank, world, local)
                gb.build(_lib.STAGE_CONNECT)
                sharding.exchange_lists(gb, 1, owner, rank, world, local)
                gb.build(_lib.STAGE_GRAPH)
            return step, (lambda: None)

        def step():
            # design A (north-star wording): column pairs re-parsed locally, one all-reduce(MIN) of per-gene thresholds
            gb.build(_lib.STAGE_THRESH)
            sharding.exchange_thresholds(gb, local)
            gb.build(_lib.STAGE_GRAPH)
        return step, (lambda: None)

    gb = waterfall.GraphBuilder(local)
    if args.scores_v2:
        gb.set_option("scores_v2", 1)   # variant path, not the headline configuration
    gb.set_option("max_bytes_per_line_inv", 40)   # the synthetic lines are 50-62 bytes
    if args.wave_species:
        gb.set_option("wave_species", args.wave_species)
    gb.set_species([G] * S, L)
    parts = sharding.partition_rows(S, world)
    owned = parts[rank]
    if world > 1:
        gb.set_partition(owned, mode == "thresholds")
    t0 = time.time()
    gb.synth_hits(P, ids)
    log("[rank %d] generated hit text on device in %.1fs" % (rank, time.time() - t0))
    stream = torch.cuda.ExternalStream(gb.lib.ofg_stream(gb.h), device=torch.device("cuda", local))
    step, closer = make_step(gb, S, [G] * S, parts)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
        gb.sync()
    c = gb.counts()
    sizes = {}
    text_bytes_all = 0
    pairs_here = [(p, j) for p in owned for j in range(S)]
    if world > 1 and mode == "thresholds":
        pairs_here += sharding.rank_pairs(owned, S)[1]
    for (p, j) in pairs_here:
        sizes[(p, j)] = gb.hits_text_size(p, j)
        text_bytes_all += sizes[(p, j)]
    hash0 = text_hash(gb.graph_text_body()) if world > 1 else None
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    host_ms = 0.0
    with ClockSampler(local) as clk:
        ev0.record(stream)
        kern = {}
        for _ in range(args.steps):
            th = time.perf_counter()
            step()
            host_ms += (time.perf_counter() - th) * 1e3    # host time to enqueue the step (async: no synchronisation inside)
            for k, (n, ms) in gb.kernel_report().items():  # reads the step's results back (the one host sync of the step)
                a = kern.setdefault(k, [0, 0.0])
                a[0] += n
                a[1] += ms
        ev1.record(stream)
        barrier()
    ms_total = ev0.elapsed_time(ev1)
    kernel_ms = sum(v[1] for v in kern.values()) / args.steps
    lines_parsed = c["lines"]
    if world > 1:
        own_bytes = sum(sizes[(p, j)] for p in owned for j in range(S))
        lines_owned = int(round(lines_parsed * own_bytes / max(text_bytes_all, 1)))
    else:
        lines_owned = lines_parsed
    t = torch.tensor([ms_total, float(lines_owned), float(gb.timing()["launches"]), kernel_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms_total, lines_all, launches, kernel_ms_max = float(tmax[0]), float(tsum[1]), float(tsum[2]), float(tmax[3])
    else:
        lines_all, launches, kernel_ms_max = float(lines_owned), float(t[2]), kernel_ms
    ms_per_step = ms_total / args.steps
    value = lines_all / (ms_per_step / 1e3)

    par_text = {"1": "1 GPU"}
    if world > 1:
        par_text = {"peer": "query species sharded over %d GPUs, rows only; transposed best-hit and connect index lists stored into the "
                            "owners' inboxes over NVLink by the export kernel (CUDA IPC peer memory, design B), one 4-byte all-reduce per "
                            "exchange" % world,
                    "lists": "query species sharded over %d GPUs, rows only; transposed best-hit and connect index lists exchanged "
                             "with NCCL all-to-all (design B, host-sized)" % world,
                    "thresholds": "query species sharded over %d GPUs, column pairs re-parsed, thresholds all-reduced (min) over NCCL "
                                  "(design A)" % world}
    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
           "data": "synthetic",
           "config": {"workload": "%s: %d species x %d genes, hq=%d (%d pairs, %d hit lines, %.2f GB text%s)" % (
               ("C3" if (args.config == "C2" and S == 64 and G == 20000 and P.hq == 50) else args.config), S, G, P.hq, S * S, int(lines_all),
               text_bytes_all * (world if world > 1 else 1) / 1e9,
               ", per-rank text incl. column pairs" if (world > 1 and mode == "thresholds") else ""),
               "scores": "v2 (NormalisedBitScore, percentile-ratio thresholds)" if args.scores_v2 else "default (length-normalised fit)",
               "species": S, "genes_per_species": G, "hits_per_query_per_pair": P.hq, "lines": int(lines_all),
               "pairs_per_gpu": len(pairs_here), "wave_species": args.wave_species,
               "parallelism": par_text["1"] if world == 1 else par_text[mode],
               "l2": "inputs (%.1f GB text per GPU) larger than L2, no flush needed" % (text_bytes_all / 1e9)},
           "gpu_launches": int(launches * args.steps),
           "clocks": clk.summary(), "counts": c, "stage_ms_last_step": gb.timing(),
           "non_kernel_ms_per_step": max(0.0, ms_per_step - kernel_ms_max), "host_enqueue_ms_per_step": host_ms / args.steps}
    # ---- parity of the workload that was timed
    parity = None
    if not args.no_parity:
        if world == 1:
            parity = parity_single(gb, S, G, L, os.cpu_count() or 1)
        else:
            h1 = text_hash(gb.graph_text_body())
            hs = [None] * world
            dist.all_gather_object(hs, (hash0, h1))
            parity = parity_multi(args, world, rank, local, make_step)
            if rank == 0:
                parity["full_size_blocks_identical_across_steps"] = bool(all(a == b for a, b in hs))
                parity["full_size_block_hashes"] = [b for _, b in hs]
                parity["ok"] = bool(parity["ok"] and parity["full_size_blocks_identical_across_steps"])
        if rank == 0:
            out["parity"] = parity
            log("parity:", json.dumps(parity))
    if rank == 0:
        # ---- roofline of the dominant kernel (per-launch CUDA-event time, averaged over the timed steps)
        peak, how = peaks()
        best = None
        rows = []
        for k, (n, ms) in sorted(kern.items(), key=lambda kv: -kv[1][1]):
            ab = kernel_alg_bytes(k, c, text_bytes_all, n / args.steps)
            per = ms / n
            gbs = (ab / (per * 1e-3) / 1e9) if ab else None
            rows.append({"kernel": k, "launches_per_step": n / args.steps, "ms_per_step": ms / args.steps,
                         "alg_GB_per_launch": (ab / 1e9) if ab else None, "GBps": gbs})
            if best is None and ab:
                best = (k, gbs, per, ab)
        out["kernels"] = rows
        traffic, tnote = None, None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            tt = tj.get(best[0]) if best else None
            if tt and abs(tt["workload_lines"] - c["lines"]) <= 0.01 * c["lines"]:
                traffic = tt["dram_bytes_per_launch"]  # from the committed ncu --set full capture of this workload
                tnote = tt.get("source")
        except Exception:
            pass
        if best:
            out["roofline"] = {"bound": "hbm", "kernel": best[0], "achieved": best[1], "peak": peak, "unit": "GB/s",
                               "frac": best[1] / peak, "traffic": traffic, "traffic_source": tnote, "peak_source": how,
                               "ms_per_launch": best[2], "alg_bytes_per_launch": best[3]}
        # whole-path figure of SURVEY.md §8(d): B_alg = T + 124 bytes per line
        T = text_bytes_all / max(c["lines"], 1)
        out["path_roofline"] = {"B_alg_bytes_per_line": T + 124, "achieved_GBps_per_gpu": value / world * (T + 124) / 1e9,
                                "frac_of_measured_peak": value / world * (T + 124) / 1e9 / peak,
                                "frac_of_8TBps": value / world * (T + 124) / 8e12}
    # ---- the consumer: MCL on the graph that sits in device memory; not part of the headline metric
    if args.mcl > 0 and rank == 0 and world == 1:
        try:
            from orthofinder_b200 import mcl as pm
            r = pm.MCL.RunMCLOnBuilder(gb, None, args.mcl)
            r.pop("labels", None)
            out["mcl"] = dict(r, inflation=args.mcl, genes=S * G, edges=c["edges"], unit="ms (device time of the whole process, CUDA events)")
        except Exception as e:
            out["mcl"] = {"error": "%s: %s" % (type(e).__name__, e)}
    # ---- e2e through the host API with pinned host buffers
    if not args.no_e2e:
        try:
            out["e2e"], closer2 = run_e2e(args, gb, world, rank, local, S, G, P, L, parts, sizes, step, make_step, lines_all, barrier, closer)
            if closer2 is not None:
                closer = closer2
        except Exception as e:   # the headline line must survive a failure of the host-buffer leg
            out["e2e"] = {"error": "%s: %s" % (type(e).__name__, e)}
    # ---- CPU baseline: the oracle port on a bounded sample of the same workload (rank 0, N = 1)
    if rank == 0 and world == 1 and not args.no_cpu:
        out["cpu_baseline"] = cpu_sample(gb, S, G, L, cores=1, n_species=min(S, args.cpu_species))
    if rank == 0:
        print(json.dumps(out))
    closer()
    gb.close()
    if world > 1:
        dist.destroy_process_group()


def run_e2e(args, gb, world, rank, local, S, G, P, L, parts, sizes, step, make_step, lines_all, barrier, closer_main):
    """The same metric through the public host API with pinned HOST buffers: H2D of every hit file and D2H of the graph
    text inside the timed region."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from orthofinder_b200 import waterfall, sharding
    e2e_world_note = None
    closer2 = None
    if world > 1:
        # every rank pins its own hit text in HOST memory: the full multi-GPU workload (32 GB of text per rank at C3) does
        # not fit the box's RAM 8 times over, so the e2e leg runs the largest weak-scaled job whose text does
        S2 = S
        budget = args.e2e_host_gb * 1e9
        per_pair = sum(sizes.values()) / max(len(sizes), 1)
        while S2 > world and (S2 * S2 / world) * per_pair > budget / world:
            S2 -= world
        if S2 != S:
            # the same builder (and its device buffers) takes the smaller job
            closer_main()
            L2 = L[:S2]
            gb.set_species([G] * S2, L2)
            parts2 = sharding.partition_rows(S2, world)
            gb.set_partition(parts2[rank], args.exchange == "thresholds")
            gb.synth_hits(P, list(range(S2)))
            step, closer2 = make_step(gb, S2, [G] * S2, parts2)
            pairs2 = [(p, j) for p in parts2[rank] for j in range(S2)]
            if args.exchange == "thresholds":
                pairs2 += sharding.rank_pairs(parts2[rank], S2)[1]
            sizes = {k: gb.hits_text_size(*k) for k in pairs2}
            S, L, parts = S2, L2, parts2
            e2e_world_note = "e2e workload reduced to %d species (host RAM: every rank pins its hit text)" % S2
    total = sum(sizes.values())
    pinned = torch.empty(total + 64, dtype=torch.uint8).pin_memory()
    base = pinned.data_ptr()
    offs, o = {}, 0
    for k, n in sizes.items():
        gb.hits_text_into(k[0], k[1], base + o, n)
        offs[k] = o
        o += n
    step()
    gb.sync()
    lines_e2e = gb.counts()["lines"]
    use_gz = args.e2e_input == "gz"
    gz_total, gz_offs, gz_sizes, gz_pinned, gz_base, t_comp = 0, {}, {}, None, 0, 0.0
    if use_gz:
        # the hit files as DIAMOND leaves them on disk: gzip (config.json:48-52), written as independent 64 KB members with a
        # size hint in every header (BGZF layout, readable by gzip.open) so that the device can inflate the members in parallel
        from concurrent.futures import ThreadPoolExecutor
        pn = pinned.numpy()
        t0c = time.perf_counter()

        def comp(k):
            return k, waterfall.gzip_members(pn[offs[k]:offs[k] + sizes[k]], level=1)
        with ThreadPoolExecutor(max_workers=min(32, (os.cpu_count() or 1) * 2)) as pool:
            blobs = dict(pool.map(comp, list(sizes.keys())))
        t_comp = time.perf_counter() - t0c
        gz_total = sum(len(b) for b in blobs.values())
        gz_pinned = torch.empty(gz_total + 64 * len(blobs) + 64, dtype=torch.uint8).pin_memory()
        gz_base = gz_pinned.data_ptr()
        gn = gz_pinned.numpy()
        o = 0
        for k in sizes:
            b = blobs[k]
            gn[o:o + len(b)] = np.frombuffer(b, np.uint8)
            gz_offs[k], gz_sizes[k] = o, len(b)
            o += (len(b) + 63) & ~63
        del blobs
        log("[rank %d] compressed %.2f GB of hit text to %.2f GB (zlib level 1, %d-byte members) in %.1fs" % (
            rank, total / 1e9, gz_total / 1e9, waterfall.BGZF_CHUNK, t_comp))
    out_pinned = torch.empty(gb.graph_text_size() + (1 << 20), dtype=torch.uint8).pin_memory()
    out_np = out_pinned.numpy()
    want_text = gb.graph_text_body().tobytes() if args.e2e_check else None
    pipelined = world == 1 and args.e2e_parts > 1
    pb = None
    if pipelined:
        # the public pipelined builder: row partitions in separate contexts of this GPU, host->device copies of
        # partition g+1 overlapped with the parse/normalise of partition g
        pb = waterfall.PipelinedGraphBuilder(local, args.e2e_parts)
        pb.set_species([G] * S, L, [sum(sizes[(p, j)] for j in range(S)) for p in range(S)])
        if args.scores_v2:
            pb.set_option("scores_v2", 1)
        pb.set_option("max_bytes_per_line_inv", 40)
        for c_, owned_ in zip(pb.ctxs, pb.parts):
            c_.set_option("reserve_text_bytes", sum(sizes[(p, j)] + 16 for p in owned_ for j in range(S)) + 4096)
            if use_gz:
                c_.set_option("reserve_gz_bytes", sum(gz_sizes[(p, j)] + 80 for p in owned_ for j in range(S)) + 4096)

        def e2e_step():
            if use_gz:
                pb.load(lambda p, j: (gz_base + gz_offs[(p, j)], gz_sizes[(p, j)]), gz=True)
            else:
                pb.load(lambda p, j: (base + offs[(p, j)], sizes[(p, j)]))
            pb.build()
            return pb.graph_text_body(out_np).size
    else:
        gb.set_option("reserve_text_bytes", total + 16 * len(sizes) + 4096)
        if use_gz:
            gb.set_option("reserve_gz_bytes", gz_total + 80 * len(sizes) + 4096)

        def e2e_step():
            gb.clear_hits()
            for k, n in sizes.items():
                if use_gz:
                    gb.add_hits_gz(k[0], k[1], (gz_base + gz_offs[k], gz_sizes[k]))
                else:
                    gb.add_hits_ptr(k[0], k[1], base + offs[k], n)
            step()
            return gb.graph_text_body(out_np).size

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    n_e2e = max(1, min(args.steps, 3))
    d2h = 0
    for _ in range(n_e2e):
        d2h = e2e_step()
    barrier()
    dt = (time.perf_counter() - t0) / n_e2e
    tt = torch.tensor([dt, float(lines_e2e), float(gz_total if use_gz else total), float(d2h)], dtype=torch.float64, device="cuda")
    if world > 1:
        tmax = tt.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(tt, op=dist.ReduceOp.SUM)
        dt, lines_tot, h2d, d2h_all = float(tmax[0]), float(tt[1]), float(tt[2]), float(tt[3])
    else:
        lines_tot, h2d, d2h_all = float(lines_e2e), float(total), float(d2h)
    res = {"value": lines_tot / dt, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h_all),
           "ms_per_step": dt * 1e3, "steps": n_e2e, "lines": int(lines_tot),
           "input": ("gzip: the hit files as multi-member .gz in pinned host memory (%d-byte members with BGZF size hints, zlib level 1, "
                     "%.2f GB for %.2f GB of text on this rank); only the compressed bytes cross PCIe, the device inflates one member per warp"
                     % (waterfall.BGZF_CHUNK, gz_total / 1e9, total / 1e9)) if use_gz else "plain text of the hit files in pinned host memory",
           "note": ("pinned host text -> PipelinedGraphBuilder (%d row partitions in asynchronous contexts of the GPU: the H2D copies, "
                    "the builds and the on-device list hand-over of the partitions overlap) -> graph text D2H" % len(pb.parts)) if pipelined else
                   "pinned host text -> ofg_add_hits_text (H2D) -> ofg_build -> graph text D2H, per rank"}
    if use_gz and world == 1:
        # the same leg with plain text in the host buffers (what round 1 measured): PCIe carries every text byte
        use_gz = False
        e2e_step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(2):
            e2e_step()
        torch.cuda.synchronize()
        dt2 = (time.perf_counter() - t0) / 2
        res["plain_text_input"] = {"value": lines_tot / dt2, "unit": UNIT, "ms_per_step": dt2 * 1e3, "h2d_bytes_per_step": int(total)}
        use_gz = True
    if e2e_world_note:
        res["workload"] = e2e_world_note
    if want_text is not None:
        got = out_np[:d2h].tobytes()
        res["text_equals_device_resident_run"] = bool(got == want_text)
    if pb is not None:
        pb.close()
    del pinned
    return res, closer2


def cpu_sample(gb, S, G, L, cores, n_species):
    """Oracle (kind 'port') on the first n_species species of the workload, texts taken from the device arena."""
    from oracle import waterfall_oracle as wo, parallel
    texts = {(p, j): gb.hits_text(p, j) for p in range(n_species) for j in range(n_species)}
    n_seqs = [G] * n_species
    lines = sum(t.count(b"\n") for t in texts.values())
    t0 = time.perf_counter()
    if cores == 1:
        res = wo.waterfall(n_seqs, L[:n_species], lambda p, j: texts[(p, j)])
        wo.graph_text(res["indptr"], res["indices"], res["data"], sum(n_seqs))
    else:
        parallel.waterfall_parallel(n_seqs, L[:n_species], texts, cores)
    dt = time.perf_counter() - t0
    return {"value": lines / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "first %d species of the workload (%d pairs, %d lines), oracle/waterfall_oracle.py, %.1f s" % (
                n_species, n_species ** 2, lines, dt)}


# ---------------------------------------------------------------------------------------------- reference arm
def run_reference(args):
    """The reference's algorithm on the host cores.  The reference is pure Python (scripts_of/gathering.py) and
    cannot travel to the GPU box, so this arm times the oracle port (numpy + C helpers, pinned byte for byte
    against the live reference) with every host core, one task per species pair."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from orthofinder_b200 import synth
    from oracle import parallel
    world = max(args.gpus, 1)
    S, G, P, L = workload(args, world)
    ns = min(S, args.ref_species)
    cores = os.cpu_count() or 1
    lib = synth.host_lib()
    t0 = time.time()
    texts = {(p, j): synth.pair_text(lib, P, p, j, p, j, L[p], L[j])[0] for p in range(ns) for j in range(ns)}
    lines = sum(t.count(b"\n") for t in texts.values())
    log("reference arm: generated %d lines for %d pairs in %.1fs; %d cores" % (lines, ns * ns, time.time() - t0, cores))
    n_seqs = [G] * ns

    def step():
        parallel.waterfall_parallel(n_seqs, L[:ns], texts, min(cores, ns * ns))

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    value = lines / dt
    sample = "first %d species of the workload (%d pairs, %d lines per step), oracle port, %d processes" % (
        ns, ns * ns, lines, min(cores, ns * ns))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "%s: %d species x %d genes, hq=%d — bounded sample: %s" % (args.config, S, G, P.hq, sample),
                   "species": S, "genes_per_species": G, "hits_per_query_per_pair": P.hq,
                   "lines": lines, "lines_per_step": lines, "sample_species": ns},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": min(cores, ns * ns), "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def wait_for_build(paths, timeout=900.0):
    """Ranks other than local rank 0: poll for the libraries rank 0 builds (published by an atomic rename)."""
    t0 = time.time()
    while time.time() - t0 < timeout:
        if all(os.path.exists(p) for p in paths):
            time.sleep(1.0)
            return
        time.sleep(0.5)
    raise SystemExit("timed out waiting for the native libraries: %s" % ", ".join(paths))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C2")
    ap.add_argument("--species", type=int, default=0)
    ap.add_argument("--genes", type=int, default=0)
    ap.add_argument("--hq", type=int, default=0)
    ap.add_argument("--exchange", default="peer", choices=["peer", "lists", "thresholds"],
                    help="multi-GPU: 'peer' = design B, lists stored into peer inboxes over NVLink by the export kernel; 'lists' = design B "
                         "through NCCL all-to-all; 'thresholds' = design A (column re-parse + all-reduce)")
    ap.add_argument("--wave-species", type=int, default=0, help="RAW + NORM over this many query species at a time (0 = all at once)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the parity check of the timed workload against the CPU oracle")
    ap.add_argument("--scores-v2", action="store_true", help="measure the --scores-v2 variant of the path instead of the default scores")
    ap.add_argument("--e2e-parts", type=int, default=4, help="row partitions of the pipelined e2e run (1 = single context)")
    ap.add_argument("--e2e-check", action="store_true", help="compare the e2e graph text with the device-resident run's")
    ap.add_argument("--e2e-input", default="text", choices=["gz", "text"],
                    help="what the pinned host buffers of the e2e leg hold: the .gz files (inflated on the device) or plain text")
    ap.add_argument("--e2e-host-gb", type=float, default=110.0, help="N > 1: host memory all ranks together may pin for the e2e leg")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--mcl", type=float, default=0.0, metavar="INFLATION",
                    help="N = 1: after the timed loop, cluster the device-resident graph with MCL on the device (SURVEY N3) and add an `mcl` block")
    ap.add_argument("--cpu-species", type=int, default=4, help="species in the cpu_baseline sample")
    ap.add_argument("--ref-species", type=int, default=8, help="species in the reference-arm sample")
    args = ap.parse_args()
    from orthofinder_b200 import _lib
    need = [_lib.LIB_PATH, os.path.join(ROOT, "oracle", "liboracle.so"), os.path.join(ROOT, "orthofinder_b200", "libofg_synth.so")]
    if not all(os.path.exists(p) for p in need):
        if int(os.environ.get("LOCAL_RANK", "0")) == 0:
            import __graft_entry__ as g
            g.build()
        else:
            wait_for_build(need)
    if args.scores_v2 and not args.no_parity and int(os.environ.get("WORLD_SIZE", "1")) == 1:
        args.no_parity = True   # the whole-workload oracle check covers the default scores
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
