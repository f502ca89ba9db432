#!/usr/bin/env python
"""Condenses an `ncu --metrics gpu__time_duration.sum --csv` launch list into (1) a per-kernel markdown summary and
(2) a compact per-launch list (id, kernel, grid, us) that is small enough to commit under profiles/.
Usage: python tools/summarize_launches.py gpurun_out/launches_X.csv profiles/r01/ncu_launches_X"""
import collections
import csv
import gzip
import re
import sys


def main(src, dst):
    lines = [l for l in open(src) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    per = []
    for row in csv.DictReader(lines):
        name = re.sub(r"^void ", "", row["Kernel Name"])
        name = re.sub(r"\(.*", "", name).replace("mobi::", "")
        v = float(row["Metric Value"].replace(",", ""))
        v = v / 1e3 if row["Metric Unit"] == "ns" else (v * 1e3 if row["Metric Unit"] == "ms" else v)
        agg[name][0] += 1
        agg[name][1] += v
        per.append((row["ID"], name[:60], row["Grid Size"].replace(" ", ""), "%.2f" % v))
    tot = sum(v[1] for v in agg.values())
    mine = sum(v[1] for k, v in agg.items() if not (k.startswith("at::") or "cutlass" in k or "cublas" in k.lower()))
    with open(dst + ".md", "w") as f:
        f.write("# ncu launch list summary (%s)\n\n" % src)
        f.write("`ncu --metrics gpu__time_duration.sum --clock-control none` over a shortened bench run "
                "(serialised, cold-cache launches: use the SHARES, not the absolute times).\n\n")
        f.write("%d launches, %.1f ms total; %.1f %% of the time in this repo's own kernels "
                "(the rest: torch RNG / copy kernels of input and weight synthesis).\n\n" % (len(per), tot / 1e3, 100 * mine / tot))
        f.write("| kernel | launches | total ms | share |\n|---|---|---|---|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
            f.write("| `%s` | %d | %.2f | %.1f %% |\n" % (k[:90], v[0], v[1] / 1e3, 100 * v[1] / tot))
    with gzip.open(dst + ".tsv.gz", "wt") as f:
        f.write("id\tkernel\tgrid\tus\n")
        for p in per:
            f.write("\t".join(p) + "\n")
    print("wrote %s.md and %s.tsv.gz (%d launches)" % (dst, dst, len(per)))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
