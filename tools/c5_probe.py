#!/usr/bin/env python
"""Probe of the whole-fit kernel: seconds per iteration of ONE small fit (no concurrency) for a few shapes, then the
config-5 grid for both chain-team shapes and two stream counts (numbers: profiles/r02_summary.md).  python tools/c5_probe.py"""
import ctypes as C, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "python-temporal-ame-svi_b200")); sys.path.insert(0, ROOT)
import torch, bench
from tame_b200 import _lib
lib = _lib.load(); dev = torch.device("cuda", 0); torch.cuda.set_device(dev)
stream = torch.cuda.current_stream(dev).cuda_stream
r = 2; d = 2 + 2 * r
for (n, T) in [(10, 5), (10, 40), (64, 20), (256, 5), (256, 40)]:
    for mode in (_lib.MODE_NAIVE, _lib.MODE_GOOD):
        c = bench.hyper_constants(n, T, r, ar=0.9, rho=0.3)
        X = bench.gen_latents(c, seed=3).to(dev)
        Y = torch.empty(n, n, T, 2, dtype=torch.float64, device=dev)
        _lib.check(lib.tame_generate_Y(n, T, r, _lib.dptr(np.ascontiguousarray(c["R"].reshape(4))), X.data_ptr(), C.c_uint64(3), 0, n, Y.data_ptr(), stream))
        cfgs = (_lib.TameConfig * 1)(); cfg, kk = bench.make_cfg(_lib, c, n, T, r, 0, 1, 0); cfg.mode = mode; cfgs[0] = cfg
        Yp, Mp, Cp = (C.c_void_p * 1)(), (C.c_void_p * 1)(), (C.c_void_p * 1)()
        it = 150; el = np.zeros((1, it)); ms = np.zeros((1, it)); nd = (C.c_int32 * 1)()
        best = 1e9
        for rep in range(3):
            Xm, Xc = bench.init_state(n, T, d, dev, seed=7)
            Yp[0], Mp[0], Cp[0] = Y.data_ptr(), Xm.data_ptr(), Xc.data_ptr()
            torch.cuda.synchronize(dev); t0 = time.time()
            _lib.check(lib.tame_fit_batch(1, cfgs, Yp, Mp, Cp, it, 0.0, _lib.dptr(el), _lib.dptr(ms), nd, 0))
            torch.cuda.synchronize(dev); best = min(best, time.time() - t0)
        print(f"n={n} T={T} mode={mode}: {best*1e3:.2f} ms per fit call, {best/it*1e6:.1f} us per iteration (incl. set-up)", flush=True)
for nh in ("2", "1"):                       # wide / narrow chain team
    for s in ("16", "32"):
        os.environ["TAME_NH"] = nh
        os.environ["TAME_BATCH_STREAMS"] = s
        out = bench.extra_config5(lib, _lib, dev)
        print("NH", nh, "streams", s, round(out["seconds"], 3), out["gpu_launches"], out["early_stopped"], flush=True)
