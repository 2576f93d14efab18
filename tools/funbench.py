#!/usr/bin/env python
"""Times blsq_model_linexp_fun variants (tools/variants/libmodels_u*.so) at the C4 shape."""
import ctypes as C, glob, json, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
m, n = 1 << 24, 64
dev = torch.device("cuda:0")
J = torch.randn((m, n), dtype=torch.float64, device=dev)
t = torch.rand(m, dtype=torch.float64, device=dev)
y = torch.randn(m, dtype=torch.float64, device=dev)
x = torch.rand(n, dtype=torch.float64, device=dev)
F = torch.empty(m, dtype=torch.float64, device=dev)
ref = None
paths = sys.argv[1:] or sorted(glob.glob(os.path.join(ROOT, "tools/variants/libmodels_*.so")))
for path in paths:
    lib = C.CDLL(path)
    f = lib.blsq_model_linexp_fun
    f.argtypes = [C.c_int64, C.c_int] + [C.c_void_p] * 6
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(3):
        f(m, n, J.data_ptr(), t.data_ptr(), y.data_ptr(), x.data_ptr(), F.data_ptr(), st)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        f(m, n, J.data_ptr(), t.data_ptr(), y.data_ptr(), x.data_ptr(), F.data_ptr(), st)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    if ref is None: ref = F.clone()
    print(json.dumps(dict(lib=os.path.basename(path), ms=ms, gbs=(m * (n - 4 + 3) * 8) / ms / 1e6,
                          maxdiff=float((F - ref).abs().max()))))
