"""Alignment of estimated states with the truth -- the step that follows every fit in the reference's drivers
(reference: src/utils/alignment.py; demo.py:135-137, tests/test_integration.py:267-279).

Same function names, arguments and return types as the reference module; the work runs in libtame_b200
(csrc/tame_align.cu: cross-covariances, Jacobi polar factor, rotation + per-row sign flips, squared error) on the
device the inputs live on -- CPU tensors are copied to cuda:0 and the result is returned on the caller's device and
dtype.  There is no CPU implementation here: without the CUDA library these functions raise.

Reference quirks that are kept (see oracle/align_oracle.py): R = U Vt of svd(X_true' X_est), `align_signs` on the last
dimension flips whole ROWS, the global mode rotates the 2r-dimensional block at once.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from . import _lib

_MAX_WIDTH = 16


def _dev(x: torch.Tensor) -> torch.device:
    if x.is_cuda:
        return x.device
    if not torch.cuda.is_available():
        raise RuntimeError("tame_b200.alignment needs a CUDA device (no CPU fallback)")
    return torch.device("cuda", 0)


def _prep(x: torch.Tensor, dev: torch.device) -> torch.Tensor:
    return x.detach().to(device=dev, dtype=torch.float64).contiguous()


def _back(y: torch.Tensor, like: torch.Tensor) -> torch.Tensor:
    return y.to(device=like.device, dtype=like.dtype)


def _stream(dev):
    return torch.cuda.current_stream(dev).cuda_stream


def _align_states_device(Xe, Xt, r, each, want_mse):
    """(n, T, d) float64 CUDA tensors -> (aligned, mse or None)."""
    lib = _lib.load()
    n, T, d = Xe.shape
    if d != 2 + 2 * r:
        raise ValueError(f"state dimension {d} does not match latent_dim {r} (expected {2 + 2 * r})")
    out = torch.empty_like(Xe)
    mse = C.c_double(0.0)
    with torch.cuda.device(Xe.device):
        _lib.check(lib.tame_align_states(n, T, r, Xe.data_ptr(), Xt.data_ptr(), 1 if each else 0, out.data_ptr(), None,
                                         C.byref(mse) if want_mse else None, _stream(Xe.device)))
    return out, (mse.value if want_mse else None)


def _signs_device(Xe, Xt, want_mse=False):
    """(rows, width) float64 CUDA tensors."""
    lib = _lib.load()
    out = torch.empty_like(Xe)
    mse = C.c_double(0.0)
    with torch.cuda.device(Xe.device):
        _lib.check(lib.tame_align_signs(Xe.shape[0], Xe.shape[1], Xe.data_ptr(), Xt.data_ptr(), out.data_ptr(),
                                        C.byref(mse) if want_mse else None, _stream(Xe.device)))
    return out, (mse.value if want_mse else None)


def procrustes_alignment(X_est: torch.Tensor, X_true: torch.Tensor, scaling: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
    """alignment.py:31-103: (X_est @ R [* s], R) with R = U Vt of svd(X_true' X_est), det R = +1."""
    if X_est.ndim != 2 or X_est.shape != X_true.shape:
        raise ValueError("procrustes_alignment expects two (n, d) matrices of the same shape")
    n, k = X_est.shape
    if k > _MAX_WIDTH:
        raise ValueError(f"procrustes_alignment on the device supports d <= {_MAX_WIDTH}, got {k}")
    lib = _lib.load()
    dev = _dev(X_est)
    Xe, Xt = _prep(X_est, dev), _prep(X_true, dev)
    out = torch.empty_like(Xe)
    R = torch.empty(k, k, dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.tame_procrustes(n, k, Xe.data_ptr(), Xt.data_ptr(), 1 if scaling else 0, out.data_ptr(), R.data_ptr(),
                                       _stream(dev)))
    return _back(out, X_est), _back(R, X_est)


def align_signs(X_est: torch.Tensor, X_true: torch.Tensor, dim: int = -1) -> torch.Tensor:
    """alignment.py:106-166.  dim == last: every X_est[i] (everything behind the first axis) is negated when that brings
    it closer to X_true[i]; any other dim: every slice along `dim`."""
    dev = _dev(X_est)
    if dim == -1 or dim == X_est.ndim - 1:
        rows = X_est.shape[0]
        Xe, Xt = _prep(X_est, dev).reshape(rows, -1), _prep(X_true, dev).reshape(rows, -1)
        out, _ = _signs_device(Xe, Xt)
        return _back(out.reshape(X_est.shape), X_est)
    Xe = _prep(X_est, dev).movedim(dim, 0)
    Xt = _prep(X_true, dev).movedim(dim, 0)
    shape = Xe.shape
    out, _ = _signs_device(Xe.reshape(shape[0], -1).contiguous(), Xt.reshape(shape[0], -1).contiguous())
    return _back(out.reshape(shape).movedim(0, dim).contiguous(), X_est)


def align_latent_positions(M_est: torch.Tensor, M_true: torch.Tensor, latent_dim: int) -> torch.Tensor:
    """alignment.py:169-221: U and V blocks of (n, 2r): Procrustes each, then row signs."""
    r = latent_dim
    dev = _dev(M_est)
    n = M_est.shape[0]
    pad = torch.zeros(n, 1, 2, dtype=torch.float64, device=dev)
    Xe = torch.cat([pad, _prep(M_est, dev).reshape(n, 1, 2 * r)], dim=2).contiguous()
    Xt = torch.cat([pad, _prep(M_true, dev).reshape(n, 1, 2 * r)], dim=2).contiguous()
    out, _ = _align_states_device(Xe, Xt, r, True, False)
    return _back(out[:, 0, 2:].contiguous(), M_est)


def align_temporal_states(X_est: torch.Tensor, X_true: torch.Tensor, latent_dim: int,
                          align_each_time: bool = True) -> torch.Tensor:
    """alignment.py:224-313: (n, T, d) trajectories, per time step or with one global rotation."""
    dev = _dev(X_est)
    out, _ = _align_states_device(_prep(X_est, dev), _prep(X_true, dev), latent_dim, align_each_time, False)
    return _back(out, X_est)


def compute_alignment_error(X_est: torch.Tensor, X_true: torch.Tensor, latent_dim: Optional[int] = None,
                            align: bool = True) -> Tuple[float, torch.Tensor]:
    """alignment.py:316-385: (mean squared error after alignment, aligned estimate)."""
    if align and X_est.ndim == 3 and latent_dim is None:
        raise ValueError("latent_dim must be provided for temporal alignment")
    if not align or X_est.ndim not in (2, 3):
        return ((X_est - X_true) ** 2).mean().item(), X_est
    dev = _dev(X_est)
    Xe, Xt = _prep(X_est, dev), _prep(X_true, dev)
    if X_est.ndim == 3:
        out, mse = _align_states_device(Xe, Xt, latent_dim, True, True)
    elif latent_dim is not None:
        out, mse = _align_states_device(Xe.unsqueeze(1).contiguous(), Xt.unsqueeze(1).contiguous(), latent_dim, True, True)
        out = out[:, 0]
    else:
        out, mse = _signs_device(Xe, Xt, want_mse=True)
    return mse, _back(out, X_est)


def compute_correlation_after_alignment(X_est: torch.Tensor, X_true: torch.Tensor, latent_dim: Optional[int] = None) -> float:
    """alignment.py:388-435: Pearson correlation of the flattened aligned estimate with the truth."""
    _, Xa = compute_alignment_error(X_est, X_true, latent_dim, align=True)
    a = Xa.flatten().double()
    b = X_true.flatten().to(a.device).double()
    a = a - a.mean()
    b = b - b.mean()
    den = torch.sqrt((a * a).sum() * (b * b).sum())
    if den < 1e-10:
        return 0.0
    return ((a * b).sum() / den).item()
