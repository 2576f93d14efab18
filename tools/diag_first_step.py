"""Why is the first timed step after the NCCL barrier slow on every rank at
N = 8?  torchrun script: C2 solves on every rank, a few variants of what
happens between the barrier and the first solve; prints per-rank step times.

    python -m torch.distributed.run --nproc-per-node 8 tools/diag_first_step.py
"""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bounded_lsq_b200 import least_squares_batched, PerProblem, models   # noqa: E402
from bounded_lsq_b200.synthetic import ExpDecay2                         # noqa: E402

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
model = ExpDecay2()
B = 1_000_000
_, yp = model.make_data(262144, seed=10000)
y = torch.from_numpy(np.tile(yp, (4, 1))[:B].copy()).to(dev)
x0 = torch.from_numpy(np.tile(model.x0, (B, 1))).to(dev)
lb, ub = torch.as_tensor(model.lb, device=dev), torch.as_tensor(model.ub, device=dev)
fun, jac = models.callbacks("ExpDecay2", "exact")


def solve():
    return least_squares_batched(fun, x0, jac=jac, bounds=(lb, ub), method="trf",
                                 args=(PerProblem(y),))


def barrier(kind):
    if world > 1:
        if kind == "nccl":
            dist.barrier()
        elif kind == "allreduce":
            t = torch.zeros(1, device=dev)
            dist.all_reduce(t)
    torch.cuda.synchronize()


def run(label, kind="nccl", stagger_ms=0.0, idle_ms=0.0, nsteps=4):
    out = None
    for _ in range(2):
        out = solve()
    barrier(kind)
    if idle_ms:
        time.sleep(idle_ms * 1e-3)
    if stagger_ms:
        time.sleep(rank * stagger_ms * 1e-3)
    marks = [torch.cuda.Event(enable_timing=True)]
    marks[0].record()
    for _ in range(nsteps):
        out = solve()
        marks.append(torch.cuda.Event(enable_timing=True))
        marks[-1].record()
    torch.cuda.synchronize()
    mine = torch.tensor([a.elapsed_time(b) for a, b in zip(marks, marks[1:])],
                        dtype=torch.float64, device=dev)
    allr = torch.empty((world, nsteps), dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_gather_into_tensor(allr.view(-1), mine)
    else:
        allr[0] = mine
    if rank == 0:
        print(json.dumps({"variant": label,
                          "first_step_ms_by_rank": [round(float(v), 2) for v in allr[:, 0]],
                          "later_steps_ms_mean_by_rank": [round(float(v), 2)
                                                          for v in allr[:, 1:].mean(1)]}),
              flush=True)


for _ in range(3):
    solve()
run("dist.barrier + synchronize (as bench.py)")
run("again")
run("all_reduce of one element + synchronize", kind="allreduce")
run("no collective, synchronize only", kind="none")
run("dist.barrier, then rank x 2 ms stagger", stagger_ms=2.0)
run("dist.barrier, then 200 ms idle", idle_ms=200.0)
run("dist.barrier + synchronize (as bench.py), third time")
if world > 1:
    dist.destroy_process_group()
