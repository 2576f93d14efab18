"""World-size-2 (gloo, CPU) test of the batched mode's multi-GPU partitioning
(SURVEY 8e): problems are split by index, there is NO collective on the data
path; the ranks only meet to gather results.  Each rank solves its shard of
the golden C2 / C3 samples through the host emulation of the kernels (TEST
ONLY) and the gathered result must be bit-identical to the single-process
solve and match the reference's golden values.  On the B200 the same split runs
under torchrun over NCCL (bench.py --gpus N)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import cases
import hostemul


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _solve(lib, name, lo, hi):
    from bounded_lsq_b200 import least_squares_batched, PerProblem
    cfg, method, jac = name.split("_")
    model = cases.MODELS[cfg]()
    z = np.load(os.path.join(cases.GOLDEN, name + ".npz"))
    dev = torch.device("cpu")
    y = cases.T(z["y"][lo:hi], dev)
    X0 = cases.T(np.tile(model.x0, (hi - lo, 1)), dev)
    j = model.jac_t if jac == "exact" else "2-point"
    res = least_squares_batched(model.fun_t, X0, jac=j, bounds=(model.lb, model.ub),
                                method=method, args=(PerProblem(y),), _lib=lib)
    return torch.cat([res.x, res.obj_value[:, None], res.status[:, None].double(),
                      res.nfev[:, None].double(), res.active_mask.double()], dim=1)


def _worker(rank, world, port, name, B, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lib = hostemul.get()
        lo, hi = B * rank // world, B * (rank + 1) // world      # index shard
        mine = _solve(lib, name, lo, hi)
        parts = [torch.empty((B * (r + 1) // world - B * r // world, mine.shape[1]),
                             dtype=torch.float64) for r in range(world)]
        dist.all_gather(parts, mine)          # results only, after the solve
        if rank == 0:
            np.save(os.path.join(out_dir, "gathered.npy"), torch.cat(parts).numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("name", ["c2_trf_exact", "c3_dogbox_2point"])
def test_two_rank_index_sharding(tmp_path, name):
    lib = hostemul.get()                  # build once, before forking
    z = np.load(os.path.join(cases.GOLDEN, name + ".npz"))
    B = z["y"].shape[0]
    mp.spawn(_worker, args=(2, _free_port(), name, B, str(tmp_path)), nprocs=2, join=True)
    got = np.load(tmp_path / "gathered.npy")
    whole = _solve(lib, name, 0, B).numpy()
    assert cases.bits(got, whole)         # the split changes nothing, bit for bit
    n = whole.shape[1] // 2 - 1           # columns: x (n) | obj | status | nfev | mask (n)
    assert np.abs(got[:, :n] - z["x"]).max() / np.abs(z["x"]).max() < 1e-8
    assert (np.abs(got[:, n] - z["obj"]) / z["obj"]).max() < 1e-8
    if name.endswith("exact"):
        assert np.array_equal(got[:, n + 1], z["status"])
        assert np.array_equal(got[:, n + 2], z["nfev"])
    assert np.array_equal(got[:, n + 3:], z["mask"])
