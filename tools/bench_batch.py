#!/usr/bin/env python
"""BASELINE config 5: a grid of independent small fits through tame_fit_batch (one GPU).

The grid follows SURVEY.md section 8d (experiments/sensitivity_analysis.py:43,83-89,409-418 widened to 512 fits):
n x T x ar_coefficient x rho_dyadic, r = 2, lr 0.01, max_iter 150, tolerance 1e-4 (the reference default), naive and
good structured mean-field.  Data come from the device generator; the timed region is the tame_fit_batch call alone
(inputs resident, traces to the host).  Prints one JSON line; this is a side measurement, not bench.py's headline.

    python tools/bench_batch.py [--streams 8] [--max-iter 150]
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
    ap.add_argument("--streams", type=int, default=8)
    ap.add_argument("--max-iter", type=int, default=150)
    ap.add_argument("--tolerance", type=float, default=1e-4)
    args = ap.parse_args()
    import torch
    import bench
    from tame_b200 import _lib
    lib = _lib.load()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    stream = torch.cuda.current_stream(dev).cuda_stream
    r, lr = 2, 0.01
    d = 2 + 2 * r
    ns = [10, 20, 32, 50, 64, 100, 160, 256]
    Ts = [5, 10, 20, 40]
    ars = [0.5, 0.9]
    rhos = [0.0, 0.3, 0.6, 0.8]
    problems = []
    for n in ns:
        for T in Ts:
            for ar in ars:
                for rho in rhos:
                    problems.append((n, T, ar, rho))
    nf = 2 * len(problems)
    cfgs = (_lib.TameConfig * nf)()
    Yp, Mp, Cp = (C.c_void_p * nf)(), (C.c_void_p * nf)(), (C.c_void_p * nf)()
    keep, units = [], 0.0
    f = 0
    for k, (n, T, ar, rho) in enumerate(problems):
        c = bench.hyper_constants(n, T, r, ar=ar, rho=rho)
        X = bench.gen_latents(c, seed=100 + k).to(dev)
        Y = torch.empty(n, n, T, 2, dtype=torch.float64, device=dev)
        Rflat = np.ascontiguousarray(c["R"].reshape(4))
        _lib.check(lib.tame_generate_Y(n, T, r, _lib.dptr(Rflat), X.data_ptr(), C.c_uint64(100 + k), 0, n, Y.data_ptr(), stream))
        for mode in (_lib.MODE_NAIVE, _lib.MODE_GOOD):
            Xm, Xc = bench.init_state(n, T, d, dev, seed=7 + k)
            cfg, kk = bench.make_cfg(_lib, c, n, T, r, 0, 1, 0)
            cfg.mode, cfg.lr = mode, lr
            cfgs[f] = cfg
            keep.append((kk, Y, Xm, Xc))
            Yp[f], Mp[f], Cp[f] = Y.data_ptr(), Xm.data_ptr(), Xc.data_ptr()
            f += 1
    torch.cuda.synchronize(dev)
    el = np.zeros((nf, args.max_iter))
    ms = np.zeros((nf, args.max_iter))
    nd = (C.c_int32 * nf)()
    launches0 = lib.tame_launch_count()
    t0 = time.time()
    _lib.check(lib.tame_fit_batch(nf, cfgs, Yp, Mp, Cp, args.max_iter, args.tolerance, _lib.dptr(el), _lib.dptr(ms), nd, args.streams))
    torch.cuda.synchronize(dev)
    sec = time.time() - t0
    iters = np.array(list(nd), dtype=np.int64)
    for (n, T, ar, rho), it2 in zip(problems, iters.reshape(-1, 2)):
        units += float(n) * n * T * float(it2.sum())
    ok = bool(np.all(np.isfinite(el[np.arange(nf), iters - 1])))
    print(json.dumps({
        "workload": f"config 5: {nf} independent fits (n in {ns}, T in {Ts}, ar in {ars}, rho in {rhos}, r=2, naive+good), "
                    f"max_iter {args.max_iter}, tolerance {args.tolerance}, lr {lr}",
        "seconds": sec, "fits_per_s": nf / sec, "iterations_total": int(iters.sum()), "iterations_per_s": float(iters.sum()) / sec,
        "dyad_timesteps_per_s": units / sec, "early_stopped": int((iters < args.max_iter).sum()), "streams": args.streams,
        "gpu_launches": int(lib.tame_launch_count() - launches0), "finite": ok}), flush=True)


if __name__ == "__main__":
    main()
