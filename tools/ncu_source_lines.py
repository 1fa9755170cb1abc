#!/usr/bin/env python
"""Join an `ncu --page source --csv` SASS listing with nvdisasm's line info and aggregate per CUDA source line.

  cuobjdump -xelf all orthofinder_b200/libofg.so; nvdisasm --print-line-info ofg_api.sm_100a.cubin > all.sass
  ncu -i rep.ncu-rep --page source --csv > k.csv
  python tools/ncu_source_lines.py all.sass k.csv <mangled kernel symbol substring> [top N]
"""
import csv
import re
import sys
from collections import defaultdict


def sass_lines(sass_fn, sym):
    """[(offset, file, line, opcode text)] of one kernel."""
    out, on, cur = [], False, ("?", 0)
    for ln in open(sass_fn, errors="replace"):
        if ln.startswith(".text."):
            on = sym in ln
            continue
        if not on:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            out.append((int(m.group(1), 16), cur[0], cur[1], m.group(2)))
    return out


def main():
    sass_fn, csv_fn, sym = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 60
    sl = sass_lines(sass_fn, sym)
    rows = list(csv.reader(open(csv_fn)))
    # a report with several kernels lists them one after the other ("Kernel Name", <name>): take the one that matches
    want = sys.argv[5] if len(sys.argv) > 5 else None
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] or [0]
    pick = starts[0]
    for i in starts:
        if want and want in rows[i][1]:
            pick = i
            break
    end = next((i for i in starts if i > pick), len(rows))
    rows = rows[pick:end]
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    body = [r for r in rows[hdr_i + 1:] if r and r[0].startswith("0x")]
    ci = hdr.index("Instructions Executed")
    cs = hdr.index("# Samples")
    base = int(body[0][0], 16)
    by_off = {o: (f, l, op) for o, f, l, op in sl}
    agg = defaultdict(lambda: [0, 0, 0])
    tot_i = tot_s = 0
    for r in body:
        off = int(r[0], 16) - base
        f, l, op = by_off.get(off, ("?", 0, ""))
        n, s = int(r[ci] or 0), int(r[cs] or 0)
        a = agg[(f, l)]
        a[0] += n
        a[1] += s
        a[2] += 1
        tot_i += n
        tot_s += s
    print("kernel %s: %d SASS instructions, %d warp instructions executed, %d samples" % (sym, len(body), tot_i, tot_s))
    src_cache = {}
    for (f, l), (n, s, k) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        text = ""
        try:
            if f not in src_cache:
                import glob
                fn = glob.glob("orthofinder_b200/csrc/" + f) or glob.glob("**/" + f, recursive=True)
                src_cache[f] = open(fn[0]).read().splitlines() if fn else []
            text = src_cache[f][l - 1].strip()[:90]
        except Exception:
            pass
        print("%5.2f%% inst %5.2f%% smp  %3d sass  %s:%d  %s" % (100.0 * n / tot_i, 100.0 * s / max(tot_s, 1), k, f, l, text))


if __name__ == "__main__":
    main()
