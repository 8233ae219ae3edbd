import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "python-temporal-ame-svi_b200")); sys.path.insert(0, ROOT)
import torch, bench
from tame_b200 import _lib
lib = _lib.load(); dev = torch.device("cuda", 0); torch.cuda.set_device(dev)
for nh in ("1", "2"):
    for s in ("16", "32"):
        os.environ["TAME_NH"] = nh; os.environ["TAME_BATCH_STREAMS"] = s
        out = bench.extra_config5(lib, _lib, dev)
        print("NH", nh, "streams", s, round(out["seconds"], 3), out["gpu_launches"], flush=True)
