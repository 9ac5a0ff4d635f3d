"""The multi-GPU algorithm of pyrecest_b200/sharded.py on CPU: world_size 2 and 3 over gloo, with a numpy
backend built from the oracle standing in for the CUDA kernels.  Checks the collectives, the offspring
boundaries, the all-to-all plan and that the result equals the single-process oracle."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import pf_oracle as O


class NumpyShardBackend:
    """same contract as pyrecest_b200.sharded.CudaShardBackend, arithmetic from the oracle"""

    def scan(self, logw, shift):
        p = np.exp(logw.numpy() - float(shift))
        c = np.cumsum(p)
        return torch.from_numpy(c), torch.tensor([c[-1]], dtype=torch.float64)

    def offspring(self, cdf, offset, total, u0, j_begin, count, n_total, x, out=None, est_kind=None):
        j = np.arange(j_begin, j_begin + count, dtype=np.float64)
        u = (j + np.float64(u0)) / np.float64(n_total)
        v = (np.float64(offset) + cdf.numpy()) / np.float64(total)
        idx = np.minimum(np.searchsorted(v, u, side="right"), cdf.shape[0] - 1)
        res = torch.from_numpy(x.numpy()[idx].copy())
        if out is not None:
            out.copy_(res)
        if est_kind == "sum":
            return res.mean(dim=0) if count > 0 else torch.zeros(x.shape[1], dtype=torch.float64)
        return out if out is not None else res


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, case, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pyrecest_b200.sharded import sharded_sum, sharded_systematic_resample

    data = np.load(case)
    n_loc = data["x"].shape[0] // world
    x = torch.from_numpy(data["x"][rank * n_loc:(rank + 1) * n_loc].copy())
    logw = torch.from_numpy(data["logw"][rank * n_loc:(rank + 1) * n_loc].copy())
    local_max = torch.tensor([float(logw.max())], dtype=torch.float64)
    new, plan = sharded_systematic_resample(NumpyShardBackend(), x, logw, local_max, float(data["u0"]), est_kind="sum")
    s = sharded_sum(new.sum(dim=0))
    assert torch.allclose(sharded_sum(plan["est_sums"]), s, rtol=1e-12)  # sums reduced where the offspring are produced
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), new=new.numpy(), sum=s.numpy(),
             boundaries=np.array(plan["boundaries"]), send=np.array(plan["send"]), recv=np.array(plan["recv"]))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,skew", [(2, 0.0), (3, 0.0), (2, 6.0)])
def test_sharded_systematic_resampling_equals_oracle(tmp_path, world, skew):
    rng = np.random.default_rng(world * 10 + int(skew))
    n_loc, D = 3000, 4
    N = n_loc * world
    x = rng.standard_normal((N, D))
    # skew > 0 puts most of the weight into the first shard: uneven all-to-all pieces
    logw = -0.5 * np.sum((x - 0.3) ** 2, axis=1) - skew * (np.arange(N) >= n_loc)
    u0 = 0.6180339887
    case = str(tmp_path / "case.npz")
    np.savez(case, x=x, logw=logw, u0=u0)
    mp.spawn(_worker, args=(world, _free_port(), case, str(tmp_path)), nprocs=world, join=True)
    res = [np.load(tmp_path / f"rank{r}.npz") for r in range(world)]
    got = np.concatenate([r["new"] for r in res])
    # oracle with the same sharding
    M = logw.max()
    shards = [np.exp(logw[g * n_loc:(g + 1) * n_loc] - M) for g in range(world)]
    idx = O.sharded_systematic_indices(shards, u0)
    assert np.array_equal(got, x[idx])
    v, offs, _ = O.sharded_cdf(shards)
    assert list(res[0]["boundaries"]) == O.offspring_boundaries(offs, u0, N)
    for r in range(world):
        assert res[r]["send"].sum() == res[r]["boundaries"][r + 1] - res[r]["boundaries"][r]
        assert res[r]["recv"].sum() == n_loc
        np.testing.assert_allclose(res[r]["sum"], x[idx].sum(axis=0), rtol=1e-12)
    # and (almost) the unsharded filter: the shard-offset adds can move an index only at an exact tie
    idx1 = O.systematic_indices(np.exp(logw - M), u0)
    assert (idx != idx1).sum() <= 2
    if skew > 0:
        assert res[0]["send"][1] > 0  # shard 0 really feeds the other rank


def test_plan_helpers():
    from pyrecest_b200.sharded import exchange_plan, offspring_boundaries, shard_offsets

    offs = shard_offsets([0.25, 0.5, 0.25])
    assert offs == [0.0, 0.25, 0.75, 1.0]
    B = offspring_boundaries(offs, 0.5, 8)
    assert B == [0, 2, 6, 8]
    # 2 slots per... 8 outputs over 4 ranks would be n_loc = 2; here 3 source shards feed 3 ranks unevenly
    send, recv = exchange_plan([0, 2, 6, 9], 1, 3, 3)
    assert send == [1, 3, 0] and recv == [0, 3, 0]
    assert offspring_boundaries(offs, 0.5, 8) == O.offspring_boundaries(np.array(offs), 0.5, 8)
