"""Two device-resident C2 solves (1M problems, fused callbacks): the target of
the ncu captures (tools/gpu_session.sh ncu=...)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bounded_lsq_b200 import least_squares_batched, PerProblem, models
from bounded_lsq_b200.synthetic import ExpDecay2, GaussPeak

wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
dev = torch.device("cuda:0")
model = ExpDecay2() if wl == "c2" else GaussPeak()
B = int(os.environ.get("B", 1_000_000))
_, yp = model.make_data(65536, seed=10000)
y = torch.from_numpy(np.tile(yp, (B // 65536 + 1, 1))[:B].copy()).to(dev)
x0 = torch.from_numpy(np.tile(model.x0, (B, 1))).to(dev)
lb, ub = torch.as_tensor(model.lb, device=dev), torch.as_tensor(model.ub, device=dev)
fun, jac = models.callbacks(type(model).__name__, "exact" if wl == "c2" else "2-point")
for _ in range(int(os.environ.get("SOLVES", 1))):
    r = least_squares_batched(fun, x0, jac=jac, bounds=(lb, ub),
                              method="trf" if wl == "c2" else "dogbox",
                              args=(PerProblem(y),), options=dict(graph_tail_rounds=0))
torch.cuda.synchronize()
print(r.rounds, int((r.status > 0).sum()))
