#!/usr/bin/env python
"""BASELINE config 5: a grid of independent small fits through tame_fit_batch (one GPU).

The grid follows SURVEY.md section 8d (experiments/sensitivity_analysis.py:43,83-89,409-418 widened to 512 fits):
n x T x ar_coefficient x rho_dyadic, r = 2, lr 0.01, max_iter 150, tolerance 1e-4 (the reference default), naive and
good structured mean-field.  Data come from the device generator; the timed region is the tame_fit_batch call alone
(inputs resident, traces to the host).  Prints one JSON line; this is a side measurement, not bench.py's headline.

    python tools/bench_batch.py [--streams 16] [--max-iter 150]
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "python-temporal-ame-svi_b200"))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--streams", type=int, default=0, help="streams of the device-loop path (0 = the library default, 16)")
    ap.add_argument("--max-iter", type=int, default=150)
    ap.add_argument("--tolerance", type=float, default=1e-4)
    args = ap.parse_args()
    if args.streams > 0:
        os.environ["TAME_BATCH_STREAMS"] = str(args.streams)      # read by tame_fit_batch when n_streams <= 0
    import torch
    import bench
    from tame_b200 import _lib
    lib = _lib.load()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    out = bench.extra_config5(lib, _lib, dev, max_iter=args.max_iter, tolerance=args.tolerance)
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
