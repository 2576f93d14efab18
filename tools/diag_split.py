"""Does a C2 solve get faster when the batch is split into independently
progressing parts on separate streams (round kernels of one part overlapping
the HBM-bound kernels of another)?  One driver thread per part."""
import os, sys, time, json, threading
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bounded_lsq_b200 import least_squares_batched, PerProblem, models
from bounded_lsq_b200.synthetic import ExpDecay2

dev = torch.device("cuda:0")
model = ExpDecay2()
B = 1_000_000
_, yp = model.make_data(262144, seed=10000)
y = torch.from_numpy(np.tile(yp, (4, 1))[:B].copy()).to(dev)
x0 = torch.from_numpy(np.tile(model.x0, (B, 1))).to(dev)
lb, ub = torch.as_tensor(model.lb, device=dev), torch.as_tensor(model.ub, device=dev)


def solve_part(c0, c1, stream, out, k, cb, opts):
    with torch.cuda.stream(stream):
        out[k] = least_squares_batched(cb[0], x0[c0:c1], jac=cb[1], bounds=(lb, ub),
                                       method="trf", args=(PerProblem(y[c0:c1]),),
                                       options=dict(opts))


def run(label, parts, prio=False, **opts):
    streams = [torch.cuda.Stream(dev, priority=(-1 if (prio and k % 2) else 0))
               for k in range(parts)]
    cbs = [models.callbacks("ExpDecay2", "exact") for _ in range(parts)]
    cuts = [B * k // parts for k in range(parts + 1)]
    ts = []
    for i in range(8):
        torch.cuda.synchronize()
        out = [None] * parts
        t0 = time.perf_counter()
        if parts == 1:
            solve_part(0, B, torch.cuda.current_stream(dev), out, 0, cbs[0], opts)
        else:
            th = [threading.Thread(target=solve_part,
                                   args=(cuts[k], cuts[k + 1], streams[k], out, k, cbs[k], opts))
                  for k in range(parts)]
            for t in th:
                t.start()
            for t in th:
                t.join()
        torch.cuda.synchronize()
        ts.append(round((time.perf_counter() - t0) * 1e3, 2))
    print(json.dumps(dict(variant=label, wall_ms=ts,
                          nfev=float(torch.cat([o.nfev for o in out]).double().mean()))),
          flush=True)


run("1 part, graph tail", 1)
run("1 part, eager tail", 1, graph_tail_rounds=0)
run("2 parts, eager tail", 2, graph_tail_rounds=0)
run("2 parts, eager tail, one high-priority stream", 2, prio=True, graph_tail_rounds=0)
run("4 parts, eager tail", 4, graph_tail_rounds=0)
run("4 parts, eager tail, alternating priority", 4, prio=True, graph_tail_rounds=0)
