#!/usr/bin/env python
"""Does J stay in the 126 MB L2 when a round is run in sub-batches?  The fused
ExpDecay2 callback writes J / f of S problems into ONE reused buffer and
blsq_linearise_batched reads it back at once; the whole sweep over 1M problems
is captured in a CUDA graph (no host time in the number) and timed for several
S.  S = B is today's round (J through HBM)."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bounded_lsq_b200 import _lib as L          # noqa: E402
from bounded_lsq_b200 import synthetic          # noqa: E402


def main():
    lib = L.Lib(sys.argv[1]) if len(sys.argv) > 1 else L.get_lib()
    dev = torch.device("cuda:0")
    model = synthetic.ExpDecay2()
    B, n, m = 1_000_000, 4, 64
    _, ypool = model.make_data(65536, seed=1)
    y = torch.from_numpy(np.tile(ypool, (B // 65536 + 1, 1))[:B].copy()).to(dev)
    X = torch.from_numpy(np.tile(model.x0, (B, 1))).to(dev)
    X += 0.01 * torch.rand_like(X)
    t = torch.as_tensor(model.t, device=dev)
    LS = lib.lin_record_size(n)
    lin = torch.empty((B, LS), dtype=torch.float64, device=dev)
    istate = torch.zeros((B, 8), dtype=torch.int32, device=dev)
    istate[:, 0] = -1                                   # ST_RUNNING
    st = lib.stream(X)
    side = torch.cuda.Stream(dev)
    for S in (B, 500_000, 250_000, 125_000, 62_500, 40_000, 31_250, 20_000, 15_625, 10_000):
        F = torch.empty((S, m), dtype=torch.float64, device=dev)
        J = torch.empty((S, m, n), dtype=torch.float64, device=dev)

        def sweep(stream):
            for c0 in range(0, B, S):
                A = min(S, B - c0)
                lib.call("blsq_model_expdecay2", A, None, m, t.data_ptr(),
                         X.data_ptr() + c0 * n * 8, y.data_ptr() + c0 * m * 8,
                         F.data_ptr(), J.data_ptr(), stream)
                lib.call("blsq_linearise_batched", A, None, m, n, F.data_ptr(),
                         J.data_ptr(), None, None, 0, istate.data_ptr() + c0 * 8 * 4,
                         lin.data_ptr() + c0 * LS * 8, stream)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.stream(side):
            g.capture_begin()
            sweep(side.cuda_stream)
            g.capture_end()
        torch.cuda.synchronize()
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        e0.record()
        for _ in range(reps):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        print(json.dumps({"sub_batch": S, "launch_pairs": -(-B // S),
                          "buffer_MB": round(S * m * (n + 1) * 8 / 1e6, 1),
                          "ms_per_1M_problems": round(ms, 4),
                          "GBps_algorithmic": round(B * m * (n + 1) * 8 * 2 / ms / 1e6, 1)}),
              flush=True)
        del g, F, J


if __name__ == "__main__":
    main()
