#!/usr/bin/env python
"""Per-kernel table of one build step: ncu launch list (gpu__time_duration.sum per launch, last build step of the capture) next to
the CUDA-event times of a bench.py line.   python tools/launch_table.py profiles/r02_launches_C2.csv profiles/r02_bench_C2.json"""
import csv
import json
import re
import sys
from collections import OrderedDict


def main():
    rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if r and r[0].isdigit()]
    hdr = next(r for r in csv.reader(open(sys.argv[1], errors="replace")) if r and r[0] == "ID")
    kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    names = [r[kn] for r in rows]
    last = max(i for i, n in enumerate(names) if n.startswith("parse_kernel"))
    step = rows[last:]
    ncu = OrderedDict()
    for r in step:
        name = re.sub(r"\(.*$", "", r[kn]).replace("void ", "").strip()
        v = float(r[mv].replace(",", ""))
        v = v / 1e6 if r[mu] in ("ns", "nsecond") else (v / 1e3 if r[mu] in ("us", "usecond") else v)
        e = ncu.setdefault(name, [0, 0.0])
        e[0] += 1
        e[1] += v
    d = json.loads([l for l in open(sys.argv[2]) if l.strip().startswith("{")][-1])
    ev = {k["kernel"]: k["ms_per_step"] for k in d["kernels"]}

    def event_name(n):
        n2 = re.sub(r"<.*$", "", n)
        for k in ev:
            if re.sub(r"<.*$", "", k) == n2:
                return k
        return None
    tot_n = sum(v[1] for v in ncu.values())
    tot_e = sum(ev.values())
    print("| kernel | launches | ncu ms | ncu share | events ms/step | events share |")
    print("|---|---:|---:|---:|---:|---:|")
    seen = set()
    for name, (cnt, ms) in sorted(ncu.items(), key=lambda kv: -kv[1][1]):
        k = event_name(name)
        if k and k not in seen:
            seen.add(k)
            print("| %s | %d | %.3f | %.1f%% | %.3f | %.1f%% |" % (name, cnt, ms, 100 * ms / tot_n, ev[k], 100 * ev[k] / tot_e))
        else:
            print("| %s | %d | %.3f | %.1f%% | %s | - |" % (name, cnt, ms, 100 * ms / tot_n, "(above)" if k else "-"))
    print()
    print("ncu total %.2f ms; events total %.2f ms (whole step: %.2f ms = %.2f G lines/s)." % (tot_n, tot_e, d["ms_per_step"], d["value"] / 1e9))


if __name__ == "__main__":
    main()
