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
def workload(args, n_gpus):
    from orthofinder_b200 import synth
    cfg = dict(synth.CONFIGS[args.config])
    S = args.species or cfg["S"]
    G = args.genes or cfg["G"]
    pk = dict(cfg["p"])
    if args.hq:
        pk["hq"] = args.hq
    if n_gpus > 1 and not args.species:
        # weak scaling: the pair grid (S^2) grows with the GPU count so that every GPU keeps ~256 owned pairs; S is
        # the largest multiple of the GPU count not above 16 sqrt(N), so every rank owns the same number of query
        # species and never more text than the single-GPU job (the e2e leg pins every rank's text in host memory)
        S = n_gpus * max(1, int(cfg["S"] * n_gpus ** 0.5 / n_gpus))
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
def kernel_alg_bytes(name, c, text_bytes, n_pass):
    lines, rec, nnz = c["lines"], c["records"], c["nnz"]
    table = {
        "parse_kernel": text_bytes + 16 * lines,          # read text once, write (row, col, score)
        "scatter_rows_kernel": 16 * lines + 12 * rec,     # read COO, write (col, score) into row buckets
        "rowsort_kernel": 12 * rec + 16 * nnz,            # read rows, write sorted CSR + the Lf sort keys
        "radix_hist_kernel": 4 * rec,                     # keys only
        "radix_scatter_kernel": 24 * rec,                 # read + write (Lf, score)
        "bin_cut_kernel": 8 * nnz,                        # scores of every bin
        "bin_select_kernel": 12 * nnz,
        "scale_bh_kernel": 12 * nnz + 12 * nnz,           # read (col, score), write score + flag word
        "rbh_kernel": 12 * nnz,
        "connect_kernel": 12 * nnz,
        "emit_kernel<false>": 12 * nnz,
        "emit_kernel<true>": 12 * nnz,
    }
    return table.get(name)


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
    gb = waterfall.GraphBuilder(local)
    if args.scores_v2:
        gb.set_option("scores_v2", 1)   # variant path, not the headline configuration
    gb.set_species([G] * S, L)
    parts = sharding.partition_rows(S, world)
    owned = parts[rank]
    species_rank = sharding.species_owner(parts, S)
    lists = args.exchange == "lists"
    if world > 1:
        gb.set_partition(owned, not lists)
    t0 = time.time()
    gb.synth_hits(P, ids)
    log("[rank %d] generated hit text on device in %.1fs" % (rank, time.time() - t0))
    stream = torch.cuda.ExternalStream(gb.lib.ofg_stream(gb.h), device=torch.device("cuda", local))

    def step():
        if world == 1:
            gb.build(_lib.STAGE_GRAPH)
        elif lists:
            # design B: rows only; the transposed best-hit and connect index lists travel over NVLink
            gb.build(_lib.STAGE_NORM)
            sharding.exchange_lists(gb, 0, species_rank, rank, world, local)
            gb.build(_lib.STAGE_CONNECT)
            sharding.exchange_lists(gb, 1, species_rank, rank, world, local)
            gb.build(_lib.STAGE_GRAPH)
        else:
            # design A (north star): column pairs re-parsed locally, one all-reduce(MIN) of per-gene thresholds
            gb.build(_lib.STAGE_THRESH)
            sharding.exchange_thresholds(gb, local)
            gb.build(_lib.STAGE_GRAPH)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    phase_ms = None
    if world > 1 and lists and args.phase_timing:
        # one instrumented step (outside the timed region): where does the multi-GPU step spend its time?
        def lap(fn):
            torch.cuda.synchronize(); t = time.perf_counter(); fn(); torch.cuda.synchronize(); return (time.perf_counter() - t) * 1e3
        phase_ms = {
            "build_norm": lap(lambda: gb.build(_lib.STAGE_NORM)),
            "exchange_bh": lap(lambda: sharding.exchange_lists(gb, 0, species_rank, rank, world, local)),
            "build_connect": lap(lambda: gb.build(_lib.STAGE_CONNECT)),
            "exchange_connect": lap(lambda: sharding.exchange_lists(gb, 1, species_rank, rank, world, local)),
            "build_graph": lap(lambda: gb.build(_lib.STAGE_GRAPH))}
    c = gb.counts()
    # owned lines: lines of the rows this rank owns (column-side re-parsing is implementation traffic, not work)
    sizes = {}
    text_bytes_all = 0
    pairs_here = [(p, j) for p in owned for j in range(S)]
    if world > 1 and not lists:
        pairs_here += sharding.rank_pairs(owned, S)[1]
    for (p, j) in pairs_here:
        sizes[(p, j)] = gb.hits_text_size(p, j)
        text_bytes_all += sizes[(p, j)]
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        ev0.record(stream)
        kern = {}
        for _ in range(args.steps):
            step()
            for k, (n, ms) in gb.kernel_report().items():
                a = kern.setdefault(k, [0, 0.0])
                a[0] += n
                a[1] += ms
        ev1.record(stream)
        barrier()
    ms_total = ev0.elapsed_time(ev1)
    # lines owned by this rank = all lines it parsed minus the column-side duplicates
    lines_parsed = c["lines"]
    if world > 1:
        own_bytes = sum(sizes[(p, j)] for p in owned for j in range(S))
        lines_owned = int(round(lines_parsed * own_bytes / max(text_bytes_all, 1)))
    else:
        lines_owned = lines_parsed
    t = torch.tensor([ms_total, float(lines_owned), float(gb.timing()["launches"])], dtype=torch.float64, device="cuda")
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms_total, lines_all, launches = float(tmax[0]), float(tsum[1]), float(tsum[2])
    else:
        lines_all, launches = float(lines_owned), float(t[2])
    ms_per_step = ms_total / args.steps
    value = lines_all / (ms_per_step / 1e3)

    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
           "data": "synthetic",
           "config": {"workload": "%s: %d species x %d genes, hq=%d (%d pairs, %d hit lines, %.2f GB text%s)" % (
               args.config, S, G, P.hq, S * S, int(lines_all), text_bytes_all * (world if world > 1 else 1) / 1e9,
               ", per-rank text incl. column pairs" if (world > 1 and not lists) else ""),
               "scores": "v2 (NormalisedBitScore, percentile-ratio thresholds)" if args.scores_v2 else "default (length-normalised fit)",
               "species": S, "genes_per_species": G, "hits_per_query_per_pair": P.hq, "lines": int(lines_all),
               "parallelism": "1 GPU" if world == 1 else (
                   "query species sharded over %d GPUs, rows only; transposed best-hit and connect index lists "
                   "exchanged with NCCL all-to-all (design B)" % world if lists else
                   "query species sharded over %d GPUs, column pairs re-parsed, thresholds all-reduced (min) over "
                   "NCCL (design A)" % world),
               "l2": "inputs (%.1f GB text per GPU) larger than L2, no flush needed" % (text_bytes_all / 1e9)},
           "gpu_launches": int(launches * args.steps) if world == 1 else int(launches * args.steps),
           "clocks": clk.summary(), "counts": c, "stage_ms_last_step": gb.timing()}
    if phase_ms:
        out["phase_ms_rank0"] = phase_ms
    if rank == 0:
        # ---- roofline of the dominant kernel (per-launch CUDA-event time, averaged over the timed steps)
        peak, how = peaks()
        n_pass = kern.get("radix_scatter_kernel", [0])[0] // max(args.steps, 1)
        best = None
        rows = []
        for k, (n, ms) in sorted(kern.items(), key=lambda kv: -kv[1][1]):
            ab = kernel_alg_bytes(k, c, text_bytes_all, n_pass)
            per = ms / n
            gbs = (ab / (per * 1e-3) / 1e9) if ab else None
            rows.append({"kernel": k, "launches_per_step": n / args.steps, "ms_per_step": ms / args.steps,
                         "alg_GB_per_launch": (ab / 1e9) if ab else None, "GBps": gbs})
            if best is None and ab:
                best = (k, gbs, per, ab)
        out["kernels"] = rows
        traffic = None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            t = tj.get(best[0]) if best else None
            if t and abs(t["workload_lines"] - c["lines"]) <= 0.01 * c["lines"]:
                traffic = t["dram_bytes_per_launch"]  # from the committed ncu --set full capture of this workload
        except Exception:
            pass
        if best:
            out["roofline"] = {"bound": "hbm", "kernel": best[0], "achieved": best[1], "peak": peak, "unit": "GB/s",
                               "frac": best[1] / peak, "traffic": traffic, "peak_source": how,
                               "ms_per_launch": best[2], "alg_bytes_per_launch": best[3],
                               "note": "DRAM traffic equals the algorithmic bytes; the kernel is bound by the integer ALU pipe "
                                       "(ncu --set full: pipe_alu 73 %, issue 74 %, no memory stalls), see "
                                       "profiles/r01_top_kernels_C2_ncu_full.txt and DESIGN.md §4"}
        # whole-path figure of SURVEY.md §8(d): B_alg = T + 124 bytes per line
        T = text_bytes_all / max(c["lines"], 1)
        out["path_roofline"] = {"B_alg_bytes_per_line": T + 124, "achieved_GBps_per_gpu": value / world * (T + 124) / 1e9,
                                "frac_of_measured_peak": value / world * (T + 124) / 1e9 / peak,
                                "frac_of_8TBps": value / world * (T + 124) / 8e12}
    # ---- e2e through the host API with pinned host buffers (N = 1 only; multi-GPU ranks feed their own files)
    if not args.no_e2e:
        total = sum(sizes.values())
        pinned = torch.empty(total + 64, dtype=torch.uint8).pin_memory()
        base = pinned.data_ptr()
        offs, o = {}, 0
        for k, n in sizes.items():
            gb.hits_text_into(k[0], k[1], base + o, n)
            offs[k] = o
            o += n
        step()
        out_pinned = torch.empty(gb.graph_text_size() + (1 << 20), dtype=torch.uint8).pin_memory()
        out_np = out_pinned.numpy()
        want_text = gb.graph_text_body().tobytes() if args.e2e_check else None
        pipelined = world == 1 and args.e2e_parts > 1
        if pipelined:
            # the public pipelined builder: row partitions in separate contexts of this GPU, host->device copies of
            # partition g+1 overlapped with the parse/normalise of partition g
            pb = waterfall.PipelinedGraphBuilder(local, args.e2e_parts)
            pb.set_species([G] * S, L, [sum(sizes[(p, j)] for j in range(S)) for p in range(S)])
            if args.scores_v2:
                pb.set_option("scores_v2", 1)
            for c_, owned_ in zip(pb.ctxs, pb.parts):
                c_.set_option("reserve_text_bytes", sum(sizes[(p, j)] + 16 for p in owned_ for j in range(S)) + 4096)

            def e2e_step():
                pb.load(lambda p, j: (base + offs[(p, j)], sizes[(p, j)]))
                pb.build()
                return pb.graph_text_body(out_np).size
        else:
            gb.set_option("reserve_text_bytes", total + 16 * len(sizes) + 4096)

            def e2e_step():
                gb.clear_hits()
                for k, n in sizes.items():
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
        tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        out["e2e"] = {"value": lines_all / float(tt[0]), "unit": UNIT, "h2d_bytes_per_step": int(total), "d2h_bytes_per_step": int(d2h),
                      "ms_per_step": float(tt[0]) * 1e3, "steps": n_e2e,
                      "note": ("pinned host text -> PipelinedGraphBuilder (%d row partitions on the GPU: H2D of partition g+1 "
                               "overlaps ofg_build of partition g) -> graph text D2H" % len(pb.parts)) if pipelined else
                              "pinned host text -> ofg_add_hits_text (H2D) -> ofg_build -> graph text D2H, per rank"}
        if want_text is not None:
            got = out_np[:d2h].tobytes()
            out["e2e"]["text_equals_device_resident_run"] = bool(got == want_text)
        if pipelined:
            pb.close()
        del pinned
    # ---- CPU baseline: the oracle port on a bounded sample of the same workload (rank 0, N = 1)
    if rank == 0 and world == 1 and not args.no_cpu:
        out["cpu_baseline"] = cpu_sample(gb, S, G, L, cores=1, n_species=min(S, args.cpu_species))
    if rank == 0:
        print(json.dumps(out))
    gb.close()
    if world > 1:
        dist.destroy_process_group()


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
        "config": {"workload": "%s: %d species x %d genes, hq=%d — bounded sample: %s" % (args.config, S, G, P.hq, sample)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": min(cores, ns * ns), "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


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
    ap.add_argument("--exchange", default="lists", choices=["lists", "thresholds"],
                    help="multi-GPU: 'lists' = design B (sparse all-to-all), 'thresholds' = design A (column re-parse + all-reduce)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--scores-v2", action="store_true", help="measure the --scores-v2 variant of the path instead of the default scores")
    ap.add_argument("--phase-timing", action="store_true", help="N > 1: add a per-phase wall-clock breakdown of one extra step")
    ap.add_argument("--e2e-parts", type=int, default=4, help="row partitions of the pipelined e2e run (1 = single context)")
    ap.add_argument("--e2e-check", action="store_true", help="compare the e2e graph text with the device-resident run's")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-species", type=int, default=4, help="species in the cpu_baseline sample")
    ap.add_argument("--ref-species", type=int, default=6, help="species in the reference-arm sample")
    args = ap.parse_args()
    from orthofinder_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH) or not os.path.exists(os.path.join(ROOT, "oracle", "liboracle.so")):
        if int(os.environ.get("LOCAL_RANK", "0")) == 0:
            import __graft_entry__ as g
            g.build()
        else:
            time.sleep(90)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
