"""Range-view post-processing on one GPU: prints the `range_post` object bench.py embeds.
Usage: python tools/range_bench.py [--samples 32] [--px 512] [--iters 20]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--samples", type=int, default=32)
ap.add_argument("--px", type=int, default=512)
ap.add_argument("--iters", type=int, default=20)
a = ap.parse_args()
print(json.dumps(bench.measure_range_post("cuda:0", a.samples, a.px, iters=a.iters)))
