"""World-size-2 (gloo, CPU) test of the row-sharded tall driver
(bounded_lsq_b200/tall.py): two processes own half of the rows each, exchange
the Gram records / ||f||^2 partials with all-gathers and must (a) take
bit-identical steps on both ranks and (b) reproduce the reference's result for
the whole problem.  The kernels are the host emulation (tests/host_emul, TEST
ONLY); on the B200 the same driver runs over NCCL (bench.py --workload c4)."""
import json
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import cases
import hostemul

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, tag, method, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from bounded_lsq_b200 import least_squares
        from bounded_lsq_b200.synthetic import TallLinExp
        lib = hostemul.get()
        z = np.load(os.path.join(cases.GOLDEN, "tall.npz"))
        meta = [m for m in json.loads(str(z["meta"]))
                if m["tag"] == tag and m["method"] == method][0]
        wl = TallLinExp(meta["m"], meta["n"], seed=meta["seed"],
                        x0_tail=meta["x0_tail"])
        # this rank's rows (even split points keep the 16-byte row alignment)
        a = (meta["m"] * rank // world) // 2 * 2
        b = meta["m"] if rank == world - 1 else (meta["m"] * (rank + 1) // world) // 2 * 2
        wl.A, wl.t, wl.y, wl.m = wl.A[a:b], wl.t[a:b], wl.y[a:b], b - a
        dev = torch.device("cpu")
        wl.to_device(dev)
        trials = []
        res = least_squares(
            wl.fun_t, cases.T(wl.x0, dev), jac=wl.jac_t,
            bounds=(cases.T(wl.lb, dev), cases.T(wl.ub, dev)), method=method,
            options=dict(trace=lambda xn, s, i: trials.append(xn.numpy().copy())),
            _lib=lib)
        np.savez(os.path.join(out_dir, f"r{rank}.npz"), x=res.x.numpy(),
                 obj=res.obj_value, status=res.status, nfev=res.nfev,
                 njev=res.njev, mask=res.active_mask.numpy(),
                 trials=np.array(trials), rows=res.fun.shape[0],
                 m_total=res.m_total)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("method", ["trf", "dogbox"])
def test_two_rank_row_sharding(tmp_path, method):
    hostemul.get()                       # build once, before forking
    world, tag = 2, "c"
    mp.spawn(_worker, args=(world, _free_port(), tag, method, str(tmp_path)),
             nprocs=world, join=True)
    r0 = np.load(tmp_path / "r0.npz")
    r1 = np.load(tmp_path / "r1.npz")
    # every rank takes bit-identical steps (replicated n x n work, rank-ordered sums)
    for k in ("x", "obj", "status", "nfev", "njev", "mask", "trials"):
        assert cases.bits(r0[k], r1[k]), k
    assert int(r0["rows"]) + int(r1["rows"]) == int(r0["m_total"]) == 4096
    # ... and they are the reference's steps for the whole problem
    z = np.load(os.path.join(cases.GOLDEN, "tall.npz"))
    pre = f"{tag}_{method}_"
    obj, status, nfev, njev = z[pre + "scalars"][:4]
    assert int(r0["status"]) == int(status)
    assert int(r0["nfev"]) == int(nfev) and int(r0["njev"]) == int(njev)
    assert cases.bits(r0["mask"], z[pre + "mask"])
    gx = z[pre + "x"]
    assert np.abs(r0["x"] - gx).max() / np.abs(gx).max() < 1e-8
    assert abs(float(r0["obj"]) - obj) / obj < 1e-8
    gt = z[pre + "trials"]
    prev = None
    for k in range(3):
        prev = gt[k - 1] if k else np.concatenate([np.full(12, 0.1), [0.8, 1.5, 0.3, 4.0]])
        stepn = np.abs(gt[k] - prev).max()
        assert np.abs(r0["trials"][k] - gt[k]).max() / stepn < 1e-10
