"""Per-step times of the bench's device-resident C2 loop under variations, to
find what produces sporadic 50-120 ms steps."""
import gc, json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from bounded_lsq_b200 import least_squares_batched, PerProblem, models
from bounded_lsq_b200.synthetic import ExpDecay2
dev = torch.device("cuda:0")
model = ExpDecay2(); B = 1_000_000
_, yp = model.make_data(262144, seed=10000)
y_host = torch.from_numpy(np.tile(yp, (4, 1))[:B].copy()).pin_memory()
x0_host = torch.from_numpy(np.tile(model.x0, (B, 1))).pin_memory()
y, x0 = y_host.to(dev), x0_host.to(dev)
lb, ub = torch.as_tensor(model.lb, device=dev), torch.as_tensor(model.ub, device=dev)
fun, jac = models.callbacks("ExpDecay2", "exact")
def solve():
    return least_squares_batched(fun, x0, jac=jac, bounds=(lb, ub), method="trf", args=(PerProblem(y),))
def loop(label, n=25, sampler=False, nogc=False):
    for _ in range(3): solve()
    torch.cuda.synchronize()
    if nogc: gc.collect(); gc.disable()
    ctx = bench.ClockSampler(0) if sampler else None
    if ctx: ctx.__enter__()
    marks = [torch.cuda.Event(enable_timing=True)]; marks[0].record(); walls = [time.perf_counter()]
    for _ in range(n):
        solve()
        e = torch.cuda.Event(enable_timing=True); e.record(); marks.append(e); walls.append(time.perf_counter())
    torch.cuda.synchronize()
    if ctx: ctx.__exit__(None, None, None)
    if nogc: gc.enable()
    ms = [round(a.elapsed_time(b), 1) for a, b in zip(marks, marks[1:])]
    print(json.dumps(dict(variant=label, ms=ms, host_ms=[round((b - a) * 1e3, 1) for a, b in zip(walls, walls[1:])])), flush=True)
loop("plain")
loop("nvml sampler 20 ms", sampler=True)
loop("gc disabled", nogc=True)
loop("sampler + gc disabled", sampler=True, nogc=True)
os.environ["BLSQ_GRAPH_TAIL"] = "0"
loop("no graph tail + sampler", sampler=True)
