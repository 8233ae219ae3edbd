"""Helpers for the -m gpu tests: call libtame_b200 through its C ABI with numpy buffers."""
import ctypes as C

import numpy as np


def make_config(c, lr, mode, device=0, world=1, rank=0, panel=0):
    """tame_config from an oracle constants dict.  Returns (cfg, keepalive)."""
    from tame_b200 import _lib
    cfg = _lib.TameConfig()
    cfg.n, cfg.T, cfg.r, cfg.mode = int(c["n"]), int(c["T"]), int(c["r"]), int(mode)
    cfg.lr = float(lr)
    for k, v in enumerate(np.asarray(c["R_inv"], dtype=np.float64).reshape(-1)):
        cfg.Rinv[k] = float(v)
    cfg.logdet_R, cfg.logdet_Q, cfg.logdet_S0 = float(c["logdet_R"]), float(c["logdet_Q"]), float(c["logdet_S0"])
    keep = [np.ascontiguousarray(c[k], dtype=np.float64) for k in ("Phi", "Q_inv", "S0_inv")]
    cfg.Phi, cfg.Qinv, cfg.S0inv = (_lib.dptr(a) for a in keep)
    cfg.device, cfg.world, cfg.rank, cfg.panel = device, world, rank, panel
    return cfg, keep


def fit_host(Y, Xm, Xc, c, lr, mode, max_iter, tol=0.0):
    """tame_fit_host on copies of the inputs; returns (elbo_trace, mse_trace, Xm, Xc)."""
    from tame_b200 import _lib
    lib = _lib.load()
    cfg, keep = make_config(c, lr, mode)
    Y = np.ascontiguousarray(Y, dtype=np.float64)
    Xm = np.ascontiguousarray(Xm, dtype=np.float64).copy()
    Xc = np.ascontiguousarray(Xc, dtype=np.float64).copy()
    el = np.zeros(max(max_iter, 1))
    ms = np.zeros(max(max_iter, 1))
    nd = C.c_int32(0)
    _lib.check(lib.tame_fit_host(C.byref(cfg), Y.ctypes.data, Xm.ctypes.data, Xc.ctypes.data, max_iter, float(tol),
                                 _lib.dptr(el), _lib.dptr(ms), C.byref(nd)))
    return el[:nd.value], ms[:nd.value], Xm, Xc


class DeviceFit:
    """Handle-level access (tame_create / bind / sweep / elbo_mse) with torch CUDA tensors as storage."""

    def __init__(self, Y, Xm, Xc, c, lr, mode, device=0):
        import torch
        from tame_b200 import _lib
        self._lib_mod = _lib
        self.lib = _lib.load()
        self.cfg, self._keep = make_config(c, lr, mode, device=device)
        dev = torch.device("cuda", device)
        self.Y = torch.as_tensor(np.ascontiguousarray(Y), dtype=torch.float64).to(dev) if not hasattr(Y, "data_ptr") else Y
        self.Xm = torch.as_tensor(np.ascontiguousarray(Xm), dtype=torch.float64).to(dev) if not hasattr(Xm, "data_ptr") else Xm
        self.Xc = torch.as_tensor(np.ascontiguousarray(Xc), dtype=torch.float64).to(dev) if not hasattr(Xc, "data_ptr") else Xc
        torch.cuda.synchronize(dev)
        self.h = C.c_void_p()
        _lib.check(self.lib.tame_create(C.byref(self.cfg), C.byref(self.h)))
        _lib.check(self.lib.tame_bind_Y(self.h, self.Y.data_ptr()))
        _lib.check(self.lib.tame_bind_state(self.h, self.Xm.data_ptr(), self.Xc.data_ptr()))
        self.out = (C.c_double * 6)()

    def sweep(self):
        self._lib_mod.check(self.lib.tame_sweep(self.h))

    def elbo_mse(self):
        self._lib_mod.check(self.lib.tame_elbo_mse(self.h, self.out))
        return np.array(list(self.out))

    def close(self):
        if self.h:
            self.lib.tame_destroy(self.h)
            self.h = C.c_void_p()
