"""Run tame_align_states a few times at a benchmark shape (for ncu launch lists / event timing)."""
import ctypes as C
import os
import sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "python-temporal-ame-svi_b200"))
from tame_b200 import _lib

n, T, r = (int(x) for x in (sys.argv[1:4] or (8192, 128, 8)))
each = int(sys.argv[4]) if len(sys.argv) > 4 else 1
d = 2 + 2 * r
lib = _lib.load()
g = torch.Generator(device="cuda").manual_seed(1)
Xt = torch.randn(n, T, d, generator=g, dtype=torch.float64, device="cuda")
Xe = 0.7 * Xt + 0.5 * torch.randn(n, T, d, generator=g, dtype=torch.float64, device="cuda")
out = torch.empty_like(Xe)
st = torch.cuda.current_stream().cuda_stream
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for rep in range(8):
    if rep == 3:
        e0.record()
    _lib.check(lib.tame_align_states(n, T, r, Xe.data_ptr(), Xt.data_ptr(), each, out.data_ptr(), None, None, st))
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"align n={n} T={T} r={r} each={each}: {ms:.3f} ms, {5.0 * n * T * d * 8 / ms / 1e6:.0f} GB/s")
