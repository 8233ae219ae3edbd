"""N>1 host logic on CPU: the row-ownership map and the sharded schedule, run as 2 gloo ranks.
Each rank holds only its own rows of Y; the owner of a panel updates it and broadcasts the new means; the result must
equal the serial Gauss-Seidel sweep."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import golden_constants, load_golden, rel_err
from oracle import tame_oracle as orc


def test_deal_fits_balances_and_is_deterministic():
    """Independent fits over devices (BASELINE config 5): every fit exactly once, largest first to
    the least loaded device, the caller's order inside a group."""
    from tame_b200 import sharding as sh
    costs = [float(n) * n * T for n in (10, 20, 32, 50, 64, 100, 160, 256) for T in (5, 10, 20, 40) for _ in range(4)]
    for g in (1, 2, 3, 8):
        groups = sh.deal_fits(costs, g)
        assert len(groups) == g and sorted(k for grp in groups for k in grp) == list(range(len(costs)))
        assert all(grp == sorted(grp) for grp in groups)
        load = [sum(costs[k] for k in grp) for grp in groups]
        assert max(load) - min(load) <= max(costs)              # the greedy bound
        assert groups == sh.deal_fits(costs, g)
    assert sh.deal_fits([], 2) == [[], []]
    assert sh.deal_fits([3.0, 1.0, 2.0], 2) == [[0], [1, 2]]
    with pytest.raises(ValueError):
        sh.deal_fits([1.0], 0)


def test_ownership_map_is_a_bijection():
    from tame_b200 import sharding as sh
    for n, panel, world in [(256, 64, 2), (200, 64, 4), (64, 64, 8), (8192, 64, 8), (130, 32, 3)]:
        seen = {}
        for rank in range(world):
            rows = sh.owned_rows(n, panel, world, rank)
            cnt = 0
            for lo, hi in rows:
                for i in range(lo, hi):
                    assert sh.owner_of(i, panel, world) == rank
                    l = sh.local_row(i, panel, world)
                    assert l == cnt, (i, l, cnt)
                    if n % panel == 0:
                        assert sh.global_row(l, panel, world, rank) == i
                    seen[i] = rank
                    cnt += 1
            assert cnt == sh.local_count(n, panel, world, rank)
        assert sorted(seen) == list(range(n))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, case, meth, panel, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from tame_b200 import sharding as sh
        g = load_golden(case)
        c = golden_constants(g)
        mode = orc.MODE_OF[meth]
        lr = float(g["lr"])
        Y_local = sh.shard_rows(g["Y"], panel, world, rank)
        Xm, Xc = g[f"{meth}_init_mean"].copy(), g[f"{meth}_init_cov"].copy()

        def bcast(root, arr):
            t = torch.from_numpy(arr)
            dist.broadcast(t, root)

        for _ in range(2):
            orc.sweep_sharded(Y_local, Xm, Xc, c, lr, mode, world, rank, panel, bcast)
        # gather the covariance rows from their owners (what tame_gather_state does)
        for lo in range(0, c["n"], panel):
            hi = min(c["n"], lo + panel)
            t = torch.from_numpy(np.ascontiguousarray(Xc[lo:hi]))
            dist.broadcast(t, sh.owner_of(lo, panel, world))
            Xc[lo:hi] = t.numpy()
        # every rank must end with the same replicated means
        chk = torch.from_numpy(Xm.copy())
        dist.all_reduce(chk, op=dist.ReduceOp.MAX)
        assert np.array_equal(chk.numpy(), Xm)
        if rank == 0:
            Sm, Sc = g[f"{meth}_init_mean"].copy(), g[f"{meth}_init_cov"].copy()
            for _ in range(2):
                orc.sweep(g["Y"], Sm, Sc, c, lr, mode)
            np.save(out, np.array([rel_err(Xm, Sm), rel_err(Xc, Sc)]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("meth", ["good", "naive"])
def test_sharded_schedule_two_gloo_ranks(tmp_path, meth):
    out = str(tmp_path / "res.npy")
    mp.spawn(_worker, args=(2, _free_port(), "r3_rho08", meth, 4, out), nprocs=2, join=True)
    err = np.load(out)
    assert err[0] == 0.0 and err[1] == 0.0, err      # same arithmetic in the same order -> identical
