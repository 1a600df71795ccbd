"""Times the split-tf32 tensor-core GEMMs (csrc/tc_gemm3.cu) against the SIMT route on the PPO update's shapes."""
import json
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from apex_b200 import _capi

L = _capi.lib()
s = lambda: torch.cuda.current_stream().cuda_stream


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3  # us


out = {}
for M in (4096, 32768, 65536):
    A = torch.randn(M, 256, device="cuda")
    B = torch.relu(torch.randn(M, 256, device="cuda"))
    W = torch.randn(256, 256, device="cuda") / 16
    b = torch.randn(256, device="cuda")
    C = torch.empty(M, 256, device="cuda")
    G = torch.zeros(256, 256, device="cuda")
    row = {}
    for p in (3, 1):
        row[f"fwd_p{p}_us"] = timeit(lambda: L.apex_tc3_linear(A.data_ptr(), 256, M, 256, W.data_ptr(), 256, 1, b.data_ptr(), 1, None, 0, C.data_ptr(), 256, p, s()))
        row[f"dx_p{p}_us"] = timeit(lambda: L.apex_tc3_linear(A.data_ptr(), 256, M, 256, W.data_ptr(), 1, 256, None, 0, B.data_ptr(), 256, C.data_ptr(), 256, p, s()))
        row[f"dw_p{p}_us"] = timeit(lambda: L.apex_tc3_outer(A.data_ptr(), 256, B.data_ptr(), 256, 256, M, G.data_ptr(), 256, 1, p, s()))
    # whole MLP forward + backward, by mode
    rows, din, hid, dout = M, 50, 256, 10
    x = torch.randn(rows, din, device="cuda")
    w1, b1 = torch.randn(hid, din, device="cuda") / 7, torch.zeros(hid, device="cuda")
    w3, b3 = torch.randn(dout, hid, device="cuda") / 16, torch.zeros(dout, device="cuda")
    h1, h2, y = (torch.empty(rows, n, device="cuda") for n in (hid, hid, dout))
    dy = torch.randn(rows, dout, device="cuda")
    dh2, dh1 = torch.empty(rows, hid, device="cuda"), torch.empty(rows, hid, device="cuda")
    gw1, gb1, gw2, gb2, gw3, gb3 = (torch.zeros_like(t) for t in (w1, b1, W, b, w3, b3))

    def mlp():
        L.apex_mlp_forward(x.data_ptr(), rows, din, hid, dout, w1.data_ptr(), b1.data_ptr(), W.data_ptr(), b.data_ptr(), w3.data_ptr(), b3.data_ptr(),
                           h1.data_ptr(), h2.data_ptr(), y.data_ptr(), s())
        L.apex_mlp_backward(x.data_ptr(), rows, din, hid, dout, W.data_ptr(), w3.data_ptr(), h1.data_ptr(), h2.data_ptr(), dy.data_ptr(), dh2.data_ptr(),
                            dh1.data_ptr(), gw1.data_ptr(), gb1.data_ptr(), gw2.data_ptr(), gb2.data_ptr(), gw3.data_ptr(), gb3.data_ptr(), s())
    for mode in (0, 1, 3):
        L.apex_set_tc_mode(mode)
        row[f"mlp_fwd_bwd_mode{mode}_us"] = timeit(mlp, 10)
    L.apex_set_tc_mode(3)
    out[M] = row
    print(M, json.dumps(row), flush=True)
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], "w"), indent=1)
